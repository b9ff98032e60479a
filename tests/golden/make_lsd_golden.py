"""Generates the committed LSD fixtures under tests/golden/ (run in the build container: needs /root/reference and cv2 4.13).

  lsd_demo.npz : the reference's own golden vector for the line detector.  gray = cv2 decode + BGR2GRAY of
                 detect_3d_cuboid/data/0000_rgb_raw.jpg; ref_lines = detect_3d_cuboid/data/edge_detection/LSD/0000_edge.txt, the 271 segments the
                 reference's line_lbd node (LSD branch, line_length_thres 15) wrote for that image (6 significant digits).
  lsd_407.npz  : a second golden vector of the reference: gray = cv2 decode + BGR2GRAY of line_lbd/data/407.jpg (the image of
                 line_lbd/launch/line_detect.launch:5, use_LSD_algorithm = true); ref_lines = line_lbd/data/saved_edges.txt, the 295 segments the
                 node wrote for it (line_lbd/src/detect_lines.cpp:94-103).
  lsd_cv2.npz  : pins of the third-party (OpenCV) arithmetic against cv2 4.13 on synthetic frames:
                 * gauss7            cv2.getGaussianKernel(7, 0.6 / 0.8)
                 * atan_y/x/deg      cv2.fastAtan2 samples
                 * img_i             uint8 frames
                 * blur64_i, sc64_i  cv2.GaussianBlur / cv2.resize(INTER_LINEAR) on the CV_64F image (the reference's double pipeline)
                 * sc8_i             cv2's own front end in 4.x (u8 blur + INTER_LINEAR_EXACT), the scaled image its LSD starts from
                 * std_i / none_i    cv2.createLineSegmentDetector(LSD_REFINE_STD / _NONE).detect(img_i): segments, widths
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cube_slam_wu_b200 import synth  # noqa: E402

REF = "/root/reference/detect_3d_cuboid/data/"


def main():
    cv2.setNumThreads(1)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass
    gray = cv2.cvtColor(cv2.imread(REF + "0000_rgb_raw.jpg", 1), cv2.COLOR_BGR2GRAY)
    ref = np.loadtxt(REF + "edge_detection/LSD/0000_edge.txt")
    np.savez_compressed(os.path.join(HERE, "lsd_demo.npz"), gray=gray, ref_lines=ref)
    gray2 = cv2.cvtColor(cv2.imread("/root/reference/line_lbd/data/407.jpg", 1), cv2.COLOR_BGR2GRAY)
    ref2 = np.loadtxt("/root/reference/line_lbd/data/saved_edges.txt")
    np.savez_compressed(os.path.join(HERE, "lsd_407.npz"), gray=gray2, ref_lines=ref2)

    out = {"gauss7": cv2.getGaussianKernel(7, 0.6 / 0.8, cv2.CV_64F).ravel()}
    rng = np.random.default_rng(1)
    ys = np.r_[rng.normal(size=500), 0, 1, -1, 0, 3, -3].astype(np.float32)
    xs = np.r_[rng.normal(size=500), 1, 0, 0, -1, 3, 3].astype(np.float32)
    out["atan_y"], out["atan_x"] = ys, xs
    out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    imgs = [synth.make_lsd_frames(1, 321, 243, seed=41)[0], synth.make_lsd_frames(1, 400, 300, seed=42, texture=1.0, noise_sigma=5.0)[0]]
    for i, img in enumerate(imgs):
        out["img_%d" % i] = img
        b = cv2.GaussianBlur(img.astype(np.float64), (7, 7), 0.6 / 0.8)
        out["sc64_%d" % i] = cv2.resize(b, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
        out["sc8_%d" % i] = cv2.resize(cv2.GaussianBlur(img, (7, 7), 0.6 / 0.8), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
        for name, flag in (("none", cv2.LSD_REFINE_NONE), ("std", cv2.LSD_REFINE_STD)):
            r = cv2.createLineSegmentDetector(flag).detect(img)
            out["%s_%d" % (name, i)] = r[0].reshape(-1, 4)
            out["%s_w_%d" % (name, i)] = r[1].ravel()
    np.savez_compressed(os.path.join(HERE, "lsd_cv2.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
