#!/usr/bin/env python3
"""Golden vectors for oracle/prims.c from the container's OpenCV (cv2 4.13): the integer primitives the hot
path takes from the un-vendored OpenCV 2.4.x library did not change between 2.4 and 4.x, so cv2 pins them bit
for bit (Sobel-7 s16 with replicate border, cross morphological gradient, 2:1 INTER_LINEAR resize, fixed-point
warpPerspective, 3x3 invert).  Writes tests/golden/cv2_prims.npz (committed)."""
import os
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(7)
out = {"cv2_version": np.array(cv2.__version__)}
for i, (w, h) in enumerate([(97, 13), (11, 64), (7, 7), (3, 5)]):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if i == 0:
        img = cv2.GaussianBlur(img, (9, 9), 3)
    out["sobel%d_img" % i] = img
    out["sobel%d_dx" % i] = cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=7, borderType=cv2.BORDER_REPLICATE)
    out["sobel%d_dy" % i] = cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=7, borderType=cv2.BORDER_REPLICATE)
k = cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 3))
for i, (w, h) in enumerate([(408, 1), (61, 27), (19, 27), (2, 2)]):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    out["morph%d_img" % i] = img
    out["morph%d_out" % i] = cv2.morphologyEx(img, cv2.MORPH_GRADIENT, k, borderType=cv2.BORDER_REPLICATE)
img = rng.integers(0, 256, (1, 408), dtype=np.uint8)
out["resize_img"] = img
out["resize_out"] = cv2.resize(img, (204, 1), interpolation=cv2.INTER_LINEAR)
src_img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
out["warp_src"] = src_img
Ms, outs = [], []
for t in range(6):
    quad = np.array([[20, 15], [140, 18], [22, 100], [138, 104]], np.float32) + rng.uniform(-12, 12, (4, 2)).astype(np.float32)
    if t >= 4:
        quad += np.float32(60)  # partly outside the source: exercises BORDER_CONSTANT taps
    dst = np.array([[0, 0], [106, 0], [0, 66], [106, 66]], np.float32)
    M = cv2.getPerspectiveTransform(quad, dst).astype(np.float32)
    Ms.append(M)
    outs.append(cv2.warpPerspective(src_img, M.astype(np.float64), (107, 67), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
out["warp_M"] = np.stack(Ms)
out["warp_out"] = np.stack(outs)
A = rng.normal(size=(8, 3, 3))
out["invert_in"] = A
out["invert_out"] = np.stack([cv2.invert(a)[1] for a in A])
# bilateral filter as prepare_image_for_cat calls it (scan/expiry_categorize.cpp:52-60): d = 3, sigmaColor = 0.95
# (the reference's "space_sigma"), sigmaSpace = 2/3 (its "color_sigma"), replicate border.  Widths stay below cv2 4.x's
# SIMD width on purpose: 4.x's vector body accumulates differently from its own scalar tail (and from 2.4.x);
# the scalar tail is the 2.4.x generic-C summation order the oracle restates, and the path only ever filters
# 11-pixel-wide patches.
for i, (w, h) in enumerate([(11, 16), (11, 16), (7, 30), (3, 3)]):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if i == 1:
        img = (img // 64 * 3).astype(np.uint8)  # small differences: colour weights far from 0
    out["bilateral%d_img" % i] = img
    out["bilateral%d_out" % i] = cv2.bilateralFilter(img, 3, (3 / 2.0 - 1) * 0.3 + 0.8, (3 - 1) / 3.0, borderType=cv2.BORDER_REPLICATE)
path = os.path.join(ROOT, "tests", "golden", "cv2_prims.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
