"""Generates tests/golden/lbd_cv2.npz (run in the build container: needs cv2 4.13): the OpenCV arithmetic BinaryDescriptor::computeSobel
(line_lbd/libs/binary_descriptor.cpp:347-402) relies on, for two synthetic frames:

  img_i  : uint8 frame
  blur_i : cv2.GaussianBlur(img_i, (5, 5), 1)            (binary_descriptor.cpp:356; cv2's fixed-point CV_8U path, IPP off)
  dx_i   : cv2.Sobel(blur_i, cv2.CV_16S, 1, 0, ksize=3)  (binary_descriptor.cpp:395)
  dy_i   : cv2.Sobel(blur_i, cv2.CV_16S, 0, 1, ksize=3)  (binary_descriptor.cpp:396)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cube_slam_wu_b200 import synth  # noqa: E402


def main():
    cv2.setNumThreads(1)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass
    rng = np.random.default_rng(5)
    imgs = [synth.make_lsd_frames(1, 161, 123, seed=51)[0], rng.integers(0, 256, (37, 53), dtype=np.uint8)]
    out = {}
    for i, img in enumerate(imgs):
        b = cv2.GaussianBlur(img, (5, 5), 1)
        out["img_%d" % i] = img
        out["blur_%d" % i] = b
        out["dx_%d" % i] = cv2.Sobel(b, cv2.CV_16S, 1, 0, ksize=3)
        out["dy_%d" % i] = cv2.Sobel(b, cv2.CV_16S, 0, 1, ksize=3)
    np.savez_compressed(os.path.join(HERE, "lbd_cv2.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
