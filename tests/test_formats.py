"""Pixel formats either side of the path: dmz_YCbCr_to_RGB (dmz.cpp:58-64, cv/convert.cpp:449-504),
dmz_deinterleave_RGBA_to_R (dmz.cpp:66-109) and the three Cython stencils dmz_scharr3_dx_abs / dmz_scharr3_dy_abs /
dmz_sobel3_dx_dy (dmz.cpp:519-531, cv/sobel.cpp:556-900).  Byte / int16 work: everything is compared bit-exactly.

CPU tests pin the plain-C oracle on tests/golden/ref_formats.npz (outputs of the reference's own sources) and, where
oracle/_ref exists, live against it; the -m gpu tests compare the CUDA kernels with the oracle through the C ABI."""
import numpy as np
import pytest


def weighted(a):
    a = np.ascontiguousarray(a)
    return np.uint64((a.astype(np.uint64).ravel() * (np.arange(a.size, dtype=np.uint64) % 65521 + 1)).sum())


def chroma_grid():
    return np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))


# ---- the oracle against the reference's outputs ------------------------------------------------------------------
def test_oracle_ycbcr_golden(oracle, golden_formats):
    g = golden_formats
    cbg, crg = chroma_grid()
    for yv in (0, 1, 77, 128, 254, 255):  # all 65536 chroma pairs per luma level
        assert weighted(oracle.ycbcr_to_rgb(np.full((256, 256), yv, np.uint8), cbg, crg, 3)) == g["ycc_y%d_check" % yv]
    assert np.array_equal(oracle.ycbcr_to_rgb(g["ycc_y"], g["ycc_cb"], g["ycc_cr"], 3), g["ycc_rgb"])
    assert np.array_equal(oracle.ycbcr_to_rgb(g["ycc_y"], g["ycc_cb"], g["ycc_cr"], 4), g["ycc_rgba"])
    assert (g["ycc_rgba"][..., 3] == 255).all()


def test_oracle_rgba_and_stencils_golden(oracle, golden_formats):
    g = golden_formats
    for n in (16, 20, 1000, 1004):
        assert np.array_equal(oracle.rgba_to_r(g["rgba%d_src" % n]), g["rgba%d_r" % n])
    for i in range(int(g["n_stencil_imgs"])):
        for kind in range(3):
            assert np.array_equal(oracle.stencil3(g["st%d_img" % i], kind), g["st%d_k%d" % (i, kind)]), (i, kind)


def test_oracle_formats_vs_reference_live(ref, oracle):
    rng = np.random.default_rng(5)
    for (h, w) in [(270, 428), (48, 64), (5, 17)]:
        y, cb, cr = (rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(3))
        for ch in (3, 4):
            assert np.array_equal(ref.ycbcr_to_rgb(y, cb, cr, ch), oracle.ycbcr_to_rgb(y, cb, cr, ch))
        for kind in range(3):
            assert np.array_equal(ref.stencil3(y, kind), oracle.stencil3(y, kind))
    src = rng.integers(0, 256, 4 * 4096, dtype=np.uint8)
    assert np.array_equal(ref.rgba_to_r(src), oracle.rgba_to_r(src))


# ---- the CUDA kernels against the oracle (through the C ABI) -------------------------------------------------------
@pytest.mark.gpu
def test_ycbcr_to_rgb_exact(dmz, oracle, golden_formats):
    g = golden_formats
    cbg, crg = chroma_grid()
    ys = np.stack([np.full((256, 256), yv, np.uint8) for yv in (0, 1, 77, 128, 254, 255)])
    rgb = dmz.ycbcr_to_rgb(ys, np.broadcast_to(cbg, ys.shape), np.broadcast_to(crg, ys.shape))
    for i, yv in enumerate((0, 1, 77, 128, 254, 255)):
        assert weighted(rgb[i]) == g["ycc_y%d_check" % yv]
    for ch, key in ((3, "ycc_rgb"), (4, "ycc_rgba")):
        out = dmz.ycbcr_to_rgb(g["ycc_y"][None], g["ycc_cb"][None], g["ycc_cr"][None], channels=ch)
        assert np.array_equal(out[0], g[key])
    rng = np.random.default_rng(8)
    # the card (428 wide: 4-pixel vectors only), a camera frame (16-pixel vectors), odd sizes (byte path), a batch
    for (n, h, w) in [(2, 270, 428), (3, 480, 640), (2, 31, 45), (1, 7, 1), (5, 16, 16)]:
        y, cb, cr = (rng.integers(0, 256, (n, h, w), dtype=np.uint8) for _ in range(3))
        for ch in (3, 4):
            out = dmz.ycbcr_to_rgb(y, cb, cr, channels=ch)
            for k in range(n):
                assert np.array_equal(out[k], oracle.ycbcr_to_rgb(y[k], cb[k], cr[k], ch)), (n, h, w, ch, k)


@pytest.mark.gpu
def test_rgba_to_r_exact(dmz, oracle, golden_formats):
    g = golden_formats
    for n in (16, 20, 1000, 1004):
        assert np.array_equal(dmz.rgba_to_r(g["rgba%d_src" % n]), g["rgba%d_r" % n])
    rng = np.random.default_rng(9)
    for n in (1, 3, 4, 15, 16, 17, 640 * 480, 640 * 480 + 5):
        src = rng.integers(0, 256, 4 * n, dtype=np.uint8)
        assert np.array_equal(dmz.rgba_to_r(src), oracle.rgba_to_r(src)), n
    # device pointers that are not 16-byte aligned take the kernel's byte path
    import torch
    from util import load_pkg
    mem_device = load_pkg().MEM_DEVICE
    src = rng.integers(0, 256, 4 * 5000 + 3, dtype=np.uint8)
    d_src, d_out = torch.from_numpy(src).cuda(), torch.zeros(5000 + 8, dtype=torch.uint8, device="cuda")
    for off_in, off_out in ((1, 0), (0, 1), (3, 2), (0, 0)):
        d_out.zero_()
        dmz._check(dmz.lib.b200_rgba_to_r_batch(dmz.ctx, d_src.data_ptr() + off_in, 5000, mem_device, d_out.data_ptr() + off_out))
        got = d_out.cpu().numpy()
        assert np.array_equal(got[off_out:off_out + 5000], oracle.rgba_to_r(np.ascontiguousarray(src[off_in:off_in + 20000])))
        assert not got[off_out + 5000:].any() and not got[:off_out].any()  # nothing written outside the destination


@pytest.mark.gpu
def test_stencil3_exact(dmz, oracle, golden_formats):
    g = golden_formats
    for i in range(int(g["n_stencil_imgs"])):
        for kind in range(3):
            assert np.array_equal(dmz.stencil3(g["st%d_img" % i][None], kind)[0], g["st%d_k%d" % (i, kind)]), (i, kind)
    rng = np.random.default_rng(10)
    # card, frame, shapes around the tile size (128 x 64) and tiny ones (clamped rows and columns everywhere)
    for (n, h, w) in [(2, 270, 428), (2, 480, 640), (1, 64, 128), (1, 65, 129), (1, 63, 127), (3, 1, 1), (2, 2, 3), (1, 130, 5), (1, 3, 261)]:
        img = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        if n == 2:
            img[1] = (img[1] > 127) * 255  # extremes: the largest magnitudes (4080 / +-510)
        for kind in range(3):
            out = dmz.stencil3(img, kind)
            for k in range(n):
                assert np.array_equal(out[k], oracle.stencil3(img[k], kind)), (n, h, w, kind, k)


@pytest.mark.gpu
def test_formats_strided_device_planes(dmz, oracle):
    """Device pointers with padded rows / frames (B200_MEM_DEVICE): the row-wise 16-pixel, 4-pixel and byte paths of the
    colour conversion (dense host planes always take the flat path) and the stencils' word / byte staging."""
    import torch
    from util import load_pkg
    mem_device = load_pkg().MEM_DEVICE
    rng = np.random.default_rng(12)
    # (w, h, row stride, extra bytes between frames, byte offset of the base pointer)
    shapes = [(64, 9, 80, 0, 0), (64, 9, 80, 160, 0), (60, 7, 64, 0, 0), (60, 7, 64, 64, 4), (61, 5, 70, 3, 0), (32, 4, 32, 16, 0), (48, 3, 48, 0, 1)]
    for (w, h, rs, gap, off) in shapes:
        n = 3
        fs = rs * h + gap
        host = [rng.integers(0, 256, off + n * fs + 16, dtype=np.uint8) for _ in range(3)]
        dev = [torch.from_numpy(a).cuda() for a in host]
        planes = [np.stack([a[off + k * fs: off + k * fs + rs * h].reshape(h, rs)[:, :w] for k in range(n)]) for a in host]
        for ch in (3, 4):
            out = torch.zeros(n * h * w * ch + 32, dtype=torch.uint8, device="cuda")
            dmz._check(dmz.lib.b200_ycbcr_to_rgb_batch(dmz.ctx, dev[0].data_ptr() + off, rs, fs, dev[1].data_ptr() + off, dev[2].data_ptr() + off, rs, fs,
                                                       w, h, n, ch, mem_device, out.data_ptr()))
            got = out.cpu().numpy()
            assert not got[n * h * w * ch:].any()
            got = got[: n * h * w * ch].reshape(n, h, w, ch)
            for k in range(n):
                assert np.array_equal(got[k], oracle.ycbcr_to_rgb(planes[0][k], planes[1][k], planes[2][k], ch)), (w, h, rs, gap, off, ch, k)
            if w % 4 == 0:  # a destination that is only 4-byte aligned (RGBA then takes the 4-pixel kernel's transposed stores)
                out.zero_()
                dmz._check(dmz.lib.b200_ycbcr_to_rgb_batch(dmz.ctx, dev[0].data_ptr() + off, rs, fs, dev[1].data_ptr() + off, dev[2].data_ptr() + off, rs, fs,
                                                           w, h, n, ch, mem_device, out.data_ptr() + 4))
                got4 = out.cpu().numpy()
                assert not got4[:4].any() and not got4[4 + n * h * w * ch:].any()
                assert np.array_equal(got4[4: 4 + n * h * w * ch].reshape(n, h, w, ch), got)
        for kind in range(3):
            out = torch.zeros(n * h * w + 16, dtype=torch.int16, device="cuda")
            dmz._check(dmz.lib.b200_stencil3_batch(dmz.ctx, dev[0].data_ptr() + off, rs, fs, w, h, n, kind, mem_device, out.data_ptr()))
            got = out.cpu().numpy()
            assert not got[n * h * w:].any()
            got = got[: n * h * w].reshape(n, h, w)
            for k in range(n):
                assert np.array_equal(got[k], oracle.stencil3(planes[0][k], kind)), (w, h, rs, gap, off, kind, k)
    # larger strided frames: several tiles per frame, interior (unclamped) tiles included, rows padded to 704
    w, h, rs, n = 640, 200, 704, 2
    host = rng.integers(0, 256, n * rs * h, dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    planes = host.reshape(n, h, rs)[:, :, :w]
    for kind in range(3):
        out = torch.zeros(n * h * w, dtype=torch.int16, device="cuda")
        dmz._check(dmz.lib.b200_stencil3_batch(dmz.ctx, dev.data_ptr(), rs, rs * h, w, h, n, kind, mem_device, out.data_ptr()))
        got = out.cpu().numpy().reshape(n, h, w)
        for k in range(n):
            assert np.array_equal(got[k], oracle.stencil3(np.ascontiguousarray(planes[k]), kind)), (kind, k)


@pytest.mark.gpu
def test_formats_bad_arguments(dmz):
    y = np.zeros((1, 4, 4), np.uint8)
    with pytest.raises(Exception):
        dmz.ycbcr_to_rgb(y, y, y, channels=2)
    with pytest.raises(Exception):
        dmz.stencil3(y, 3)
    out = np.zeros(16, np.uint8)
    assert dmz.lib.b200_rgba_to_r_batch(dmz.ctx, None, 4, 0, out.ctypes.data) != 0
    assert dmz.lib.b200_ycbcr_to_rgb_batch(dmz.ctx, y.ctypes.data, 3, 16, y.ctypes.data, y.ctypes.data, 4, 16, 4, 4, 1, 3, 0, out.ctypes.data) != 0  # row stride < width
