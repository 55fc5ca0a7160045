#!/usr/bin/env python
"""Golden vectors of cv::resize (INTER_LINEAR, CV_8UC1) from the real OpenCV (cv2) of this container, for oracle/shim/cv_resize_linear_u8.h:
the sizes SmallBlurryImage::MakeFromKF (jni/SmallBlurryImage.cc:22-30) meets (level 3 -> half, exact and inexact) and a few others.
    python tests/golden/make_resize_golden.py        -> tests/golden/resize_linear.npz"""
import os

import cv2
import numpy as np

SIZES = [(240, 135, 120, 67), (80, 60, 40, 30), (60, 33, 30, 16), (135, 240, 67, 120), (33, 17, 16, 8), (100, 75, 50, 37), (480, 270, 240, 135), (61, 61, 30, 30), (64, 48, 21, 16)]

if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    out = {"sizes": np.array(SIZES, dtype=np.int32), "cv2_version": np.array(cv2.__version__)}
    for k, (sw, sh, dw, dh) in enumerate(SIZES):
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        out[f"src{k}"] = src
        out[f"dst{k}"] = cv2.resize(src, (dw, dh))          # default interpolation: INTER_LINEAR
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_linear.npz"), **out)
    print("wrote resize_linear.npz with", len(SIZES), "cases from cv2", cv2.__version__)
