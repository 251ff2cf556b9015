"""Lazy cards: b200_process_frames_batch without cards_out warps only the card rows scan_card_image reads (68 coarse vseg
rows, the 43-row fine window, and -- rarely -- the final number strip when it leaves that window).  Every record field except
the card checksum must equal the materialised path's and the oracle's; the TMA tile path and the gather path of the warp
kernel must agree everywhere (B200_DMZ_WARP_GATHER selects the latter for a whole process, so that comparison runs in a
child process)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT, deck_frames, oriented_frames

pytestmark = pytest.mark.gpu
FIELDS = ("found", "all_found", "v_y_offset", "v_pattern_type", "v_number_length", "usable", "upside_down", "h_n_offsets", "h_offsets",
          "h_pattern_offset")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def lazy(pkg):
    d = pkg.Dmz(device=0, materialise_cards=False)
    yield d
    d.close()


def same_but_checksum(lazy_recs, full_recs):
    a, b = lazy_recs.copy(), full_recs.copy()
    assert (a["card_check"] == 0).all()
    b["card_check"] = 0
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_lazy_records_equal_materialised_and_oracle(lazy, dmz, oracle):
    frames = deck_frames(500, 64)
    want = oracle.process_frames(frames)
    got = lazy.process_frames(frames)
    full = dmz.process_frames(frames)
    assert (full["card_check"] != 0).any() and same_but_checksum(got, full)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(bits(got["corners"]), bits(want["corners"])) and np.array_equal(bits(got["h_score"]), bits(want["h_score"]))
    assert np.abs(got["scores"] - want["scores"]).max() <= 1e-4
    assert (want["usable"] == 1).sum() > 32


def test_lazy_frames_without_a_card(lazy, dmz, oracle):
    """Frames where detection fails, where the vseg gate fails, and upside-down cards take the early exits of every
    lazy launch; batches mixing them with good frames keep each frame's rows apart."""
    rng = np.random.default_rng(3)
    good = deck_frames(40, 5)
    flipped = good[1][::-1, ::-1].copy()  # card upside down in the frame: scan says upside_down, no hseg
    blank_card = good[2].copy()
    blank_card[120:360, 120:520] = 175    # a card without digits: vseg score below kMinVSegScore
    frames = np.stack([good[0], np.full((480, 640), 90, np.uint8), flipped, rng.integers(0, 256, (480, 640)).astype(np.uint8),
                       blank_card, good[3], good[4]])
    want = oracle.process_frames(frames)
    got = lazy.process_frames(frames)
    assert same_but_checksum(got, dmz.process_frames(frames))
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert want["all_found"].tolist()[:3] == [1, 0, 1] and want["upside_down"][2] == 1 and want["usable"][4] == 0


@pytest.mark.parametrize("orientation", [1, 2, 4])
def test_lazy_other_orientations(lazy, dmz, oracle, orientation):
    recs, cards = oracle.process_frames(deck_frames(300, 8), want_cards=True)
    frames = oriented_frames(cards[recs["usable"] == 1][:4], orientation, oracle.detection_boxes(640, 480, orientation), seed=30 + orientation)
    want = oracle.process_frames(frames, orientation=orientation)
    got = lazy.process_frames(frames, orientation=orientation)
    assert same_but_checksum(got, dmz.process_frames(frames, orientation=orientation))
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert np.abs(got["scores"] - want["scores"]).max() <= 1e-4


def test_lazy_other_resolutions(lazy, dmz, oracle):
    for (w, h, n) in [(1280, 720, 3), (1920, 1080, 2)]:
        fr = deck_frames(0, n, w, h)
        want = oracle.process_frames(fr)
        got = lazy.process_frames(fr)
        assert same_but_checksum(got, dmz.process_frames(fr)), (w, h)
        for f in FIELDS:
            assert np.array_equal(got[f], want[f]), (w, h, f)


def _coarse_y0(oracle, card):
    """The coarse pass of best_n_vseg (n_vseg.cpp:127-137) restated on oracle.vseg_row: rows 0, 4, .., 268 only."""
    vp = np.zeros((270, 2), np.float32)
    for r in range(0, 270, 4):
        vp[r] = oracle.vseg_row(card, r)[1:3]
    vsum = asum = best = np.float32(0)
    by = 0
    for y in range(270):
        vsum, asum = np.float32(vsum + vp[y, 0]), np.float32(asum + vp[y, 1])
        if y >= 26:
            if vsum > best:
                best, by = vsum, y - 26
            if asum > best:
                best, by = asum, y - 26
            vsum, asum = np.float32(vsum - vp[y - 26, 0]), np.float32(asum - vp[y - 26, 1])
    return by


def strip_outside_window_frames(oracle):
    """Cards doctored so that the coarse pass picks a window ten rows below the final one: two coarse rows at the top of
    the number strip blanked, three digit-like rows planted below it at multiples of four.  The final strip then starts
    above the fine window [y0 - 8, y0 + 35) and the lazy path must warp its missing rows on their own (WARP_STRIP).
    The cards reach the path as frames, so detection + resampling shift rows by a fraction of a pixel: a few variants
    per card (row phase -1 / 0 / +1, two background seeds) make sure several of them keep the property."""
    recs, cards = oracle.process_frames(deck_frames(300, 8), want_cards=True)
    ok = recs["usable"] == 1
    boxes = oracle.detection_boxes(640, 480, 3)
    out = []
    for card, D in zip(cards[ok], recs["v_y_offset"][ok]):
        D = int(D)
        probs = [oracle.vseg_row(card, r)[1:3].max() for r in range(D + 10, D + 20)]
        digit_row, bg = card[D + 10 + int(np.argmax(probs))].copy(), card[D - 20].copy()
        for phase in (-1, 0, 1):
            for seed in (1, 2):
                c2, base = card.copy(), (D + 3) // 4 * 4 + phase
                c2[base + 4], c2[base + 8] = bg, bg
                c2[D + 27:min(270, D + 46)] = bg
                for r in (base + 32, base + 36, base + 40):
                    c2[r] = digit_row
                out.append(oriented_frames(c2[None], 3, boxes, jitter=0.0, seed=seed)[0])
    return np.stack(out)


def test_lazy_strip_outside_fine_window(lazy, dmz, oracle):
    frames = strip_outside_window_frames(oracle)
    want, wcards = oracle.process_frames(frames, want_cards=True)
    hits = 0
    for k in range(len(frames)):
        y0, yo = _coarse_y0(oracle, wcards[k]), int(want["v_y_offset"][k])
        lo, hi = max(0, y0 - 8), min(270, y0 + 35)
        hits += bool(want["v_score"][k] > 15 and want["upside_down"][k] == 0 and (yo < lo or yo + 27 > hi))
    assert hits >= 3, "this set should exercise the strip fix-up several times"
    got = lazy.process_frames(frames)
    assert same_but_checksum(got, dmz.process_frames(frames))
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(bits(got["h_score"]), bits(want["h_score"])) and np.abs(got["scores"] - want["scores"]).max() <= 1e-4


def test_lazy_host_crop_path(lazy, oracle):
    """Host buffers: the cropped upload + lazy rows + full-frame redo of quads that leave the crop."""
    frames = deck_frames(300, 24, jitter=14.0)
    want = oracle.process_frames(frames)
    for margin in (0, 2, -1):
        lazy.set_crop_margin(margin)
        got = lazy.process_frames(frames)
        for f in FIELDS:
            assert np.array_equal(got[f], want[f]), (margin, f)
    lazy.set_crop_margin(2)


def test_lazy_at_full_size(lazy, pkg):
    """100k device-resident frames (B200_FULLSIZE_FRAMES overrides): lazy records == materialised records except the
    checksum, byte for byte."""
    import torch
    from util import deck_frames_cuda
    n = int(os.environ.get("B200_FULLSIZE_FRAMES", "100000"))
    frames = torch.empty((n, 480, 640), dtype=torch.uint8, device="cuda")
    for f0 in range(0, n, 8192):
        cnt = min(8192, n - f0)
        frames[f0:f0 + cnt] = deck_frames_cuda(f0, cnt)
    a = torch.zeros((n, 808), dtype=torch.uint8, device="cuda")
    b = torch.zeros((n, 808), dtype=torch.uint8, device="cuda")
    lazy.process_frames_device(frames.data_ptr(), n, 640, 480, a.data_ptr())
    full = pkg.Dmz(device=0, materialise_cards=True)
    try:
        full.process_frames_device(frames.data_ptr(), n, 640, 480, b.data_ptr())
    finally:
        full.close()
    ra = a.cpu().numpy().view(pkg.RECORD_DTYPE).reshape(n)
    rb = b.cpu().numpy().view(pkg.RECORD_DTYPE).reshape(n)
    assert rb["all_found"].all() and (rb["card_check"] != 0).all()
    assert same_but_checksum(ra, rb)


CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from util import load_pkg, deck_frames
pkg = load_pkg()
frames = deck_frames(500, 48)
out = {}
for name, mat in (("full", True), ("lazy", False)):
    d = pkg.Dmz(device=0, materialise_cards=mat)
    recs, cards = d.process_frames(frames, want_cards=True) if mat else (d.process_frames(frames), None)
    out[name] = recs.view(np.uint8).copy()
    if mat:
        out["cards"] = cards
    d.close()
np.savez(%(path)r, **out)
"""


def test_tile_path_equals_gather_path(dmz, lazy, tmp_path):
    """The same frames through a process whose warp kernel is forced onto the gather path (no TMA tile): cards and
    records must be byte-identical to this process's (tile path)."""
    path = str(tmp_path / "gather.npz")
    env = dict(os.environ, B200_DMZ_WARP_GATHER="1")
    subprocess.check_call([sys.executable, "-c", CHILD % {"root": ROOT, "path": path}], env=env)
    other = np.load(path)
    frames = deck_frames(500, 48)
    recs, cards = dmz.process_frames(frames, want_cards=True)
    assert np.array_equal(cards, other["cards"])
    assert np.array_equal(recs.view(np.uint8), other["full"])
    assert np.array_equal(lazy.process_frames(frames).view(np.uint8), other["lazy"])
