"""oracle/prims.c (restated OpenCV 2.4.x primitives) against golden vectors produced by cv2 4.13
(tools/make_cv2_golden.py).  Bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT


@pytest.fixture(scope="module")
def lib(oracle):
    return oracle.lib


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_prims.npz"))


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_sobel7(lib, g):
    for i in range(4):
        img = np.ascontiguousarray(g["sobel%d_img" % i])
        h, w = img.shape
        for name, (dx, dy) in (("dx", (1, 0)), ("dy", (0, 1))):
            out = np.zeros((h, w), np.int16)
            lib.orc_sobel_u8_s16(P(img), w, w, h, P(out), 2 * w, dx, dy, 7)
            assert np.array_equal(out, g["sobel%d_%s" % (i, name)]), (i, name)


def test_sobel7_saturates(lib):
    img = np.zeros((16, 16), np.uint8)
    img[:, 8:] = 255
    out = np.zeros((16, 16), np.int16)
    lib.orc_sobel_u8_s16(P(img), 16, 16, 16, P(out), 32, 1, 0, 7)
    assert out.max() == 32767  # 255 * 64 * 10 saturates to int16


def test_morph_gradient(lib, g):
    for i in range(4):
        img = np.ascontiguousarray(g["morph%d_img" % i])
        h, w = img.shape
        out = np.zeros_like(img)
        lib.orc_morph_grad_cross3_u8(P(img), w, w, h, P(out), w)
        assert np.array_equal(out, g["morph%d_out" % i]), i


def test_resize_half(lib, g):
    img = np.ascontiguousarray(g["resize_img"])
    out = np.zeros((1, 204), np.uint8)
    lib.orc_resize_half_width_u8(P(img), 408, 408, 1, P(out), 204)
    assert np.array_equal(out, g["resize_out"])


def test_warp_perspective(lib, g):
    src = np.ascontiguousarray(g["warp_src"])
    lib.orc_warp_perspective_u8.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p]
    for M, ref in zip(g["warp_M"], g["warp_out"]):
        M = np.ascontiguousarray(M, np.float32)
        out = np.zeros((67, 107), np.uint8)
        lib.orc_warp_perspective_u8(P(src), 160, 160, 120, P(out), 107, 107, 67, P(M))
        assert np.array_equal(out, ref)


def test_invert3x3(lib, g):
    for a, ref in zip(g["invert_in"], g["invert_out"]):
        a = np.ascontiguousarray(a)
        out = np.zeros((3, 3))
        lib.orc_invert3x3(P(a), P(out))
        assert np.array_equal(out, ref)


def test_bilinear_table_quirk(lib):
    lib.orc_bilinear_tab.restype = C.POINTER(C.c_int16)
    tab = np.ctypeslib.as_array(lib.orc_bilinear_tab(), shape=(32 * 32 * 4,)).reshape(32, 32, 4)
    assert tab[0, 0].tolist() == [32767, 0, 0, 1]  # saturate_cast<short>(32768) compensated on the last tap
    assert (tab.reshape(-1, 4).sum(1) == 32768).all()
    assert tab[16, 16].tolist() == [8192] * 4


def test_bilateral(lib, g):
    """cvSmooth(CV_BILATERAL, 3, 3, 0.95, 2/3) of prepare_image_for_cat (scan/expiry_categorize.cpp:52-60)."""
    lib.orc_bilateral_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    for i in range(4):
        img = np.ascontiguousarray(g["bilateral%d_img" % i])
        h, w = img.shape
        out = np.zeros_like(img)
        lib.orc_bilateral_u8(P(img), w, w, h, P(out), w, 3, (3 / 2.0 - 1) * 0.3 + 0.8, (3 - 1) / 3.0)
        assert np.array_equal(out, g["bilateral%d_out" % i]), i
