"""Generates the committed fixtures under tests/golden/ from the reference's bundled data.

Run in the build container (needs /root/reference and cv2); the GPU box only reads the generated files.
  demo_case.npz   : detect_3d_cuboid demo inputs (detect_3d_cuboid/src/main.cpp:37-68): K, transToWolrd, bbox, the 271 LSD lines of
                    data/edge_detection/LSD/0000_edge.txt, and the float32 distance map of the box ROI computed from
                    data/0000_rgb_raw.jpg with cv2.Canny(80,200) + cv2.distanceTransform(DIST_L2,3) (cv2 4.13).
  tum_ba.npz      : object_slam offline fixture (object_slam/data/{detect_cuboids_saved,pop_cam_poses_saved,truth_cam_poses,
                    output_cam_poses,output_obj_poses}.txt), verbatim numbers.
  kat_ray_plane.npz : the worked example printed in detect_3d_cuboid/src/object_3d_util.cpp:884-905.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REF = "/root/reference"


def main():
    base = REF + "/detect_3d_cuboid/data/"
    gray = cv2.cvtColor(cv2.imread(base + "0000_rgb_raw.jpg", 1), cv2.COLOR_BGR2GRAY)
    H, W = gray.shape
    lines = np.loadtxt(base + "edge_detection/LSD/0000_edge.txt").astype(np.float64)
    K = np.array([[529.5, 0, 365.0], [0, 529.5, 265.0], [0, 0, 1.0]])
    T = np.array([[1, 0.0011, 0.0004, 0], [0, -0.3376, 0.9413, 0], [0.0011, -0.9413, -0.3376, 1.35], [0, 0, 0, 1.0]])
    boxes = np.array([[188 - 1, 189 - 1, 201, 311, 0.88]])
    tasks = O.plan(boxes, W, H)
    assert len(tasks) == 1
    t = tasks[0]
    roi = np.ascontiguousarray(gray[t.top:t.top + t.height, t.left:t.left + t.width])
    dm = cv2.distanceTransform(255 - cv2.Canny(roi, 80, 200), cv2.DIST_L2, 3).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "demo_case.npz"), K=K, T=T, boxes=boxes, lines=lines, dist_map=dm, img_w=W, img_h=H,
                        roi=np.array([t.left, t.top, t.width, t.height]))
    d = REF + "/object_slam/data/"
    np.savez_compressed(os.path.join(HERE, "tum_ba.npz"), det=np.loadtxt(d + "detect_cuboids_saved.txt"), pop=np.loadtxt(d + "pop_cam_poses_saved.txt"),
                        truth=np.loadtxt(d + "truth_cam_poses.txt"), out_cam=np.loadtxt(d + "output_cam_poses.txt"), out_obj=np.loadtxt(d + "output_obj_poses.txt"))
    # object_3d_util.cpp:884-905 (numbers as printed, 6 significant digits); K = TUM calibration of main_obj.cpp:484-486
    np.savez_compressed(os.path.join(HERE, "kat_ray_plane.npz"),
                        pixels=np.array([[344.614, 528.2, 429.72, 255.345], [424, 281.372, 233.359, 340.603]]),
                        rays=np.array([[0.0457866, 0.388681, 0.204744, -0.120948], [0.327151, 0.0626333, -0.026411, 0.172483], [1, 1, 1, 1.0]]),
                        plane=np.array([-0.1053, -0.817599, -0.566077, 1.1019]),
                        pts_sensor=np.array([[0.0601785, 0.650682, 0.398569, -0.191935], [0.429983, 0.104853, -0.0514136, 0.273717], [1.31432, 1.67407, 1.94667, 1.58692]]),
                        pts_world=np.array([[-1.7704, -1.58766, -1.1965, -1.37924], [0.682643, -0.0592668, 0.0370799, 0.778989], [0, 0, 0, 0.0]]),
                        K=np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
