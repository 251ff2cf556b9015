"""BASELINE.json-sized runs (100k frames on one GPU by default; B200_FULLSIZE_FRAMES overrides) checked through
size-independent properties, plus a random sample against the oracle."""
import os

import numpy as np
import pytest

from util import deck_frames_cuda

pytestmark = pytest.mark.gpu
N = int(os.environ.get("B200_FULLSIZE_FRAMES", "100000"))


@pytest.fixture(scope="module")
def full(dmz, pkg):
    import torch
    frames = torch.empty((N, 480, 640), dtype=torch.uint8, device="cuda")
    for f0 in range(0, N, 8192):
        cnt = min(8192, N - f0)
        frames[f0:f0 + cnt] = deck_frames_cuda(f0, cnt)
    recs = torch.zeros((N, 808), dtype=torch.uint8, device="cuda")
    dmz.process_frames_device(frames.data_ptr(), N, 640, 480, recs.data_ptr())
    return frames, recs, recs.cpu().numpy().view(pkg.RECORD_DTYPE).reshape(N)


def test_every_frame_of_the_deck_is_detected(full):
    _, _, r = full
    assert r["all_found"].all()
    assert (r["upside_down"] == 0).all()
    assert (r["usable"] == 1).mean() > 0.6


def test_deterministic_and_batch_invariant(dmz, pkg, full):
    """Running again, and running a slice on its own, give byte-identical records."""
    import torch
    frames, recs, r = full
    again = torch.zeros_like(recs)
    dmz.process_frames_device(frames.data_ptr(), N, 640, 480, again.data_ptr())
    assert bool((again == recs).all())
    lo, cnt = N // 3, min(1000, N - N // 3)
    part = torch.zeros((cnt, 808), dtype=torch.uint8, device="cuda")
    dmz.process_frames_device(frames[lo:].data_ptr(), cnt, 640, 480, part.data_ptr())
    assert bool((part == recs[lo:lo + cnt]).all())


def test_tensor_core_vseg_equals_fp32_vseg(dmz, pkg, full):
    """The tcgen05 form of the vseg hidden layer (exact integer MMAs, vseg_mma.cu) against the FP32 FMA kernel on every
    frame of the deck: every index the scan derives from the row probabilities is identical, the 27-row score sums agree
    to float noise (the two kernels round differently, neither is the reference's summation order)."""
    import torch
    frames, recs, r = full
    os.environ["B200_DMZ_VSEG_FP32"] = "1"
    try:
        other = torch.zeros_like(recs)
        dmz.process_frames_device(frames.data_ptr(), N, 640, 480, other.data_ptr())
    finally:
        del os.environ["B200_DMZ_VSEG_FP32"]
    o = other.cpu().numpy().view(pkg.RECORD_DTYPE).reshape(N)
    for k in ("v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_pattern_offset"):
        assert np.array_equal(r[k], o[k]), k
    assert np.array_equal(r["h_offsets"], o["h_offsets"])
    assert np.abs(r["v_score"] - o["v_score"]).max() <= 2e-4


def test_sessions_read_the_true_number(pkg, full):
    """Each 8-frame session shows one card number; the per-session scanner result must complete for most sessions
    and equal the deck's ground truth (Luhn-valid by construction)."""
    from util import deck_truth
    _, _, r = full
    done = correct = 0
    n_sessions = min(N // 8, 400)
    for s in range(n_sessions):
        sc = pkg.Scanner()
        for k in range(8):
            sc.add_scan(r[s * 8 + k])
        ok, digits = sc.result()
        sc.close()
        if ok:
            done += 1
            truth, _ = deck_truth(s * 8)
            correct += digits.tolist() == truth.tolist()
    assert done >= 0.5 * n_sessions
    assert correct >= 0.95 * done


def test_random_sample_against_oracle(oracle, full):
    frames, _, r = full
    idx = np.sort(np.random.default_rng(0).choice(N, size=min(256, N), replace=False))
    host = frames[idx.tolist()].cpu().numpy()
    want = oracle.process_frames(host)
    got = r[idx]
    for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_offsets"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["corners"].view(np.uint32), want["corners"].view(np.uint32))
    assert np.abs(got["scores"] - want["scores"]).max() <= 1e-4


def test_host_buffer_path_equals_device_path(dmz, pkg, full):
    """The chunked two-lane H2D pipeline returns the same bytes as the device-resident path."""
    frames, recs, r = full
    n = min(N, 5000)
    host = frames[:n].cpu().numpy()
    got = dmz.process_frames(host)
    assert np.array_equal(got.view(np.uint8).reshape(n, 808), recs[:n].cpu().numpy())


def test_categorize_only_at_scale(dmz, oracle):
    """BASELINE configs[2] (n_categorize alone), reduced to 2^18 patches (B200_FULLSIZE_PATCHES overrides): half i.i.d. noise, half crops with digit-like
    structure.  Size-independent properties (rows of every model are distributions, the ensemble is (r0+r1+r2-max)/2 of
    them, results do not depend on how the batch is cut) plus a random sample against the oracle."""
    n = int(os.environ.get("B200_FULLSIZE_PATCHES", str(1 << 18)))
    rng = np.random.default_rng(5)
    patches = rng.integers(0, 256, (n, 27, 19), dtype=np.uint8)
    yy, xx = np.mgrid[0:27, 0:19]
    for i in range(n // 2, n, max(1, n // 1024)):  # some hundred structured patches spread over the second half
        cx, cy, r = rng.uniform(5, 13), rng.uniform(8, 18), rng.uniform(3, 8)
        ring = np.abs(np.hypot(xx - cx, (yy - cy) * 0.8) - r) < 1.5
        patches[i] = np.where(ring, 210, 60).astype(np.uint8) + rng.integers(0, 20, (27, 19)).astype(np.uint8)
    ens, per_model = dmz.categorize_patches(patches)
    assert ens.shape == (n, 10) and np.isfinite(ens).all()
    assert np.abs(per_model.sum(2) - 1).max() < 1e-5
    want = (per_model.sum(1) - per_model.max(1)) / 2.0
    assert np.abs(ens - want).max() < 1e-6
    lo, cnt = n // 3 + 5, 1001  # a ragged slice on its own: byte-identical
    e2, p2 = dmz.categorize_patches(patches[lo:lo + cnt])
    assert np.array_equal(e2, ens[lo:lo + cnt]) and np.array_equal(p2, per_model[lo:lo + cnt])
    idx = np.sort(rng.choice(n, size=96, replace=False))
    for i in idx:
        e, m = oracle.digit_models(oracle.digit_patch_prep(patches[i]))
        assert np.abs(e - ens[i]).max() <= 1e-4 and np.abs(m - per_model[i]).max() <= 1e-4, i


def test_expiry_digits_at_scale(dmz, oracle):
    """E0 at 2^15 crops: run-to-run determinism (the kernel keeps eight crops in flight per CTA), batch invariance,
    and a random sample against the oracle."""
    n = 1 << 15
    rng = np.random.default_rng(9)
    crops = rng.integers(0, 256, (n, 16, 11), dtype=np.uint8)
    crops[: n // 2] = (crops[: n // 2] // 40) * 40  # few grey levels
    a = dmz.expiry_digits(crops)
    assert np.array_equal(a, dmz.expiry_digits(crops))
    assert np.abs(a.sum(1) - 1).max() < 1e-5
    lo, cnt = 12345, 777
    assert np.array_equal(dmz.expiry_digits(crops[lo:lo + cnt]), a[lo:lo + cnt])
    idx = np.sort(rng.choice(n, size=48, replace=False))
    for i in idx:
        assert np.abs(a[i] - oracle.expiry_digit_model(oracle.expiry_patch_prep(crops[i]))).max() <= 1e-4, i
