#!/usr/bin/env python3
"""Generate tests/golden/ref_formats.npz by running the REFERENCE'S OWN SOURCES (oracle/_ref/libdmz_ref.so) on seeded
inputs: dmz_YCbCr_to_RGB (3 and 4 channels), dmz_deinterleave_RGBA_to_R and the three Cython stencils
(dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy).  Run in the build container; the .npz is committed and
pins the plain-C oracle (and, through it, the CUDA path) where the reference is absent."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.binding import Oracle

R = Oracle("ref")
rng = np.random.default_rng(20261017)
out = {}

# colour conversion: every (Cb, Cr) pair against a few Y levels (the arithmetic is per pixel: this is exhaustive in the
# chroma terms and covers both saturation ends), plus a random image of the card's size class
cbg, crg = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
for i, yv in enumerate((0, 1, 77, 128, 254, 255)):
    y = np.full((256, 256), yv, np.uint8)
    rgb = R.ycbcr_to_rgb(y, cbg, crg, 3)
    out["ycc_y%d_check" % yv] = np.uint64((rgb.astype(np.uint64).ravel() * (np.arange(rgb.size, dtype=np.uint64) % 65521 + 1)).sum())
y, cb, cr = (rng.integers(0, 256, (27, 44), dtype=np.uint8) for _ in range(3))
out["ycc_y"], out["ycc_cb"], out["ycc_cr"] = y, cb, cr
out["ycc_rgb"] = R.ycbcr_to_rgb(y, cb, cr, 3)
out["ycc_rgba"] = R.ycbcr_to_rgb(y, cb, cr, 4)

# RGBA -> R: sizes that are multiples of 4 (the reference's stated assumption), with and without a remainder group of four
for n in (16, 20, 1000, 1004):
    src = rng.integers(0, 256, 4 * n, dtype=np.uint8)
    out["rgba%d_src" % n] = src
    out["rgba%d_r" % n] = R.rgba_to_r(src)

# stencils: random, a step edge, and small / thin shapes (the reference asserts width > 8)
imgs = [rng.integers(0, 256, (23, 37), dtype=np.uint8), rng.integers(0, 256, (3, 9), dtype=np.uint8),
        rng.integers(0, 256, (1, 12), dtype=np.uint8), np.repeat((np.arange(40) > 17).astype(np.uint8)[None] * 255, 9, 0)]
for i, img in enumerate(imgs):
    out["st%d_img" % i] = img
    for kind in range(3):
        out["st%d_k%d" % (i, kind)] = R.stencil3(img, kind)
out["n_stencil_imgs"] = np.int32(len(imgs))

path = os.path.join(ROOT, "tests", "golden", "ref_formats.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
