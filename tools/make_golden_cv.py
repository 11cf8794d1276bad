#!/usr/bin/env python
"""Golden vectors for SURVEY 8(f3) from the OpenCV wheel of this image (cv2.remap / cv2.cvtColor = the algorithm the
reference calls through cv::remap / cv::cvtColor).  Writes tests/golden/cv_remap.npz and tests/golden/cv_gray.npz.
usage: python tools/make_golden_cv.py"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def euroc_like_maps(w, h, scale=1.0):
    """initUndistortRectifyMap for an EuRoC-like camera (cam0 intrinsics / distortion of the MH sequences), optionally downscaled."""
    K = np.array([[458.654 * scale, 0, 367.215 * scale], [0, 457.296 * scale, 248.375 * scale], [0, 0, 1]])
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    R = cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))[0]
    P = np.array([[435.2 * scale, 0, 367.4 * scale], [0, 435.2 * scale, 252.2 * scale], [0, 0, 1]])
    return cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32F)


def main():
    rng = np.random.default_rng(2024)
    out = {"cv2_version": np.array(cv2.__version__)}
    # (a) quarter-size rectification, (b) wild maps: far outside, exact integers, halves, 1/64 ties
    src = rng.integers(0, 256, size=(120, 188), dtype=np.uint8)
    m1, m2 = euroc_like_maps(188, 120, 0.25)
    out.update(a_src=src, a_mx=m1, a_my=m2, a_dst=cv2.remap(src, m1, m2, cv2.INTER_LINEAR))
    src = rng.integers(0, 256, size=(61, 83), dtype=np.uint8)
    mx = rng.uniform(-20, 100, size=(48, 70)).astype(np.float32)
    my = rng.uniform(-20, 80, size=(48, 70)).astype(np.float32)
    mx[::3] = np.round(mx[::3]); my[::4] = np.round(my[::4]); mx[::5] += 0.5
    my[1::6] = (np.round(my[1::6] * 64) / 64).astype(np.float32)
    mx[0, :6] = [1e6, -1e6, 40000.3, -40000.7, 82.0, 82.5]
    out.update(b_src=src, b_mx=mx, b_my=my, b_dst=cv2.remap(src, mx, my, cv2.INTER_LINEAR))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cv_remap.npz"), **out)
    img = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cv_gray.npz"), cv2_version=np.array(cv2.__version__), img=img,
                        bgr=cv2.cvtColor(img[..., :3].copy(), cv2.COLOR_BGR2GRAY), rgb=cv2.cvtColor(img[..., :3].copy(), cv2.COLOR_RGB2GRAY),
                        bgra=cv2.cvtColor(img, cv2.COLOR_BGRA2GRAY), rgba=cv2.cvtColor(img, cv2.COLOR_RGBA2GRAY))
    print("written")


if __name__ == "__main__":
    main()
