"""Generates tests/golden/tum_online.npz (run in the build container: needs /root/reference): the inputs of the reference's ONLINE mode on its
bundled TUM sequence (object_slam/src/main_obj.cpp:479-841 with online_detect_mode = true, the node's default):

  jpeg, jpeg_off : the 58 frames object_slam/data/raw_imgs/%04d_rgb_raw.jpg, byte for byte (1.1 MB), back to back + offsets
  boxes          : rows [frame, x, y, w, h, prob] of object_slam/data/filter_2d_obj_txts/%04d_yolo2_0.15.txt (1-based x / y as in the files;
                   frames without a row have no detection)

The expected outputs of that mode are the reference's own committed files output_obj_poses.txt / output_cam_poses.txt, already in
tum_ba.npz (out_obj, out_cam) together with truth_cam_poses.txt (truth)."""
import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
D = "/root/reference/object_slam/data/"


def main():
    blobs, off, rows = [], [0], []
    for f in range(58):
        b = np.fromfile(D + "raw_imgs/%04d_rgb_raw.jpg" % f, np.uint8)
        blobs.append(b)
        off.append(off[-1] + len(b))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            bx = np.loadtxt(D + "filter_2d_obj_txts/%04d_yolo2_0.15.txt" % f).reshape(-1, 5)
        for r in bx:
            rows.append([f] + list(r))
    np.savez_compressed(os.path.join(HERE, "tum_online.npz"), jpeg=np.concatenate(blobs), jpeg_off=np.array(off, np.int64), boxes=np.array(rows, np.float64))
    print(len(blobs), "frames,", off[-1], "bytes,", len(rows), "boxes")


if __name__ == "__main__":
    main()
