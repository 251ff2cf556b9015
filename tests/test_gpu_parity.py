"""Parity of the CUDA path with the CPU oracle, stage by stage, through the C ABI (include/b200_dmz.h).

Bar (BASELINE.json north_star): bit-exact for the integer edge / Hough / warp / segmentation results (and for the
float geometry whose evaluation order is mirrored: line parameters, corners, homography, hseg scores); <= 1e-4
absolute on the CNN / MLP probabilities."""
import json
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, deck_frames, synthetic_strip

pytestmark = pytest.mark.gpu
TOL = 1e-4
G = os.path.join(ROOT, "tests", "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def deck():
    return deck_frames(200, 48)


@pytest.fixture(scope="module")
def orecs(oracle, deck):
    return oracle.process_frames(deck, want_cards=True)


def test_extension_is_the_cuda_library(pkg, dmz):
    assert os.path.basename(pkg.lib_path()) == "libb200dmz.so"
    before = dmz.launches
    dmz.process_frames(deck_frames(0, 1))
    assert dmz.launches > before  # kernels of the in-tree CUDA library actually ran


def test_detect_line_taps_exact(dmz, oracle, deck):
    edges, corners, found, lines = dmz.detect_edges(deck[:16], want_lines=True)
    boxes = oracle.detection_boxes(640, 480)
    for k in range(16):
        for s, (x, y, w, h) in enumerate(boxes):
            ol = oracle.best_line(deck[k][y:y + h, x:x + w], s >= 2)
            gl = lines[k, s]
            for f in ("found", "max_votes", "low", "high", "n_edge_px"):
                assert int(gl[f]) == getattr(ol, f), (k, s, f)
            if ol.found:
                assert (int(gl["r"]), int(gl["n"])) == (ol.r, ol.n)
                assert bits([gl["rho"], gl["theta"]]).tolist() == bits([ol.rho, ol.theta]).tolist()


def test_detect_edges_and_corners_exact(dmz, deck, orecs):
    rec = orecs[0]
    edges, corners, found, _ = dmz.detect_edges(deck)
    assert np.array_equal(found, rec["all_found"].astype(np.uint8))
    assert np.array_equal(edges["found"], rec["found"])
    m = rec["found"] == 1
    assert np.array_equal(bits(edges["rho"])[m], bits(rec["rho"])[m]) and np.array_equal(bits(edges["theta"])[m], bits(rec["theta"])[m])
    assert np.array_equal(bits(corners), bits(rec["corners"]))


def test_detect_edge_cases(dmz, oracle):
    rng = np.random.default_rng(3)
    frames = np.stack([
        np.full((480, 640), 90, np.uint8),                                # flat: no edges at all
        rng.integers(0, 256, (480, 640)).astype(np.uint8),                # uniform noise: high > any magnitude
        deck_frames(7, 1)[0][:, ::-1].copy(),                              # mirrored card
        np.where(np.indices((480, 640))[1] > 320, 200, 20).astype(np.uint8),  # one vertical step through the frame
    ])
    want = oracle.process_frames(frames)
    edges, corners, found, lines = dmz.detect_edges(frames, want_lines=True)
    assert np.array_equal(found, want["all_found"].astype(np.uint8))
    assert np.array_equal(edges["found"], want["found"])
    assert lines["n_edge_px"][1].sum() == 0


def test_detect_textured_frames(dmz, oracle):
    """Dense textures: many NMS candidates / edge pixels per strip (exercises the candidate- and vote-list overflow
    paths of the detect kernel).  Every per-strip tap must still equal the oracle's."""
    yy, xx = np.mgrid[0:480, 0:640]
    rng = np.random.default_rng(21)
    frames = []
    for period in (3, 4, 5, 8):
        frames.append((127 + 120 * np.sin(2 * np.pi * xx / period)).astype(np.uint8))
        frames.append((127 + 120 * np.sin(2 * np.pi * yy / period)).astype(np.uint8))
        frames.append((127 + 60 * np.sin(2 * np.pi * xx / period) + 60 * np.sin(2 * np.pi * yy / (period + 1))).astype(np.uint8))
    frames.append(((xx // 2 + yy // 2) % 2 * 255).astype(np.uint8))
    frames.append(((xx + yy) % 2 * 200 + 20).astype(np.uint8))
    n = rng.integers(0, 256, (480, 640)).astype(np.float32)
    sm = (n + np.roll(n, 1, 0) + np.roll(n, 1, 1) + np.roll(n, -1, 0) + np.roll(n, -1, 1)) / 5
    frames.append(sm.astype(np.uint8))
    frames = np.stack(frames)
    edges, corners, found, lines = dmz.detect_edges(frames, want_lines=True)
    boxes = oracle.detection_boxes(640, 480)
    for k in range(len(frames)):
        for s_, (x, y, w, h) in enumerate(boxes):
            ol = oracle.best_line(frames[k][y:y + h, x:x + w], s_ >= 2)
            gl = lines[k, s_]
            for f in ("found", "max_votes", "low", "high", "n_edge_px"):
                assert int(gl[f]) == getattr(ol, f), (k, s_, f, int(gl[f]), getattr(ol, f))
            if ol.found:
                assert (int(gl["r"]), int(gl["n"])) == (ol.r, ol.n), (k, s_)


def test_chroma_fallback_exact(dmz, oracle):
    """Y plane flat on the left -> the left edge comes from Cb (rho doubled), dmz.cpp:351-367."""
    f = deck_frames(5, 1)[0]
    cb = np.ascontiguousarray(f[::2, ::2])
    cr = np.full((240, 320), 128, np.uint8)
    y = f.copy()
    y[:, :160] = 60
    want = oracle.detect_edges(y, cb, cr)
    edges, corners, found, _ = dmz.detect_edges(y[None], cb[None], cr[None])
    assert list(edges["found"][0]) == list(want.found) and int(found[0]) == want.all_found
    assert bits(edges["rho"][0]).tolist() == bits(list(want.rho)).tolist()
    if want.all_found:
        assert bits(corners[0]).tolist() == bits(list(want.corners)).tolist()


def test_chroma_cr_supplies_edge(dmz, oracle):
    """Y misses the left and the top edge, Cb has the top edge only, Cr has every edge: the top edge comes from Cb and
    the left edge from the THIRD plane (dmz.cpp:351-367, rho multiplier 2 on both chroma planes)."""
    from test_oracle_vs_ref import chroma_case
    y, cb, cr = chroma_case()
    want = oracle.detect_edges(y, cb, cr)
    assert list(want.found) == [1, 1, 1, 1]
    edges, corners, found, _ = dmz.detect_edges(y[None], cb[None], cr[None])
    assert list(edges["found"][0]) == [1, 1, 1, 1] and int(found[0]) == 1
    assert bits(edges["rho"][0]).tolist() == bits(list(want.rho)).tolist()
    assert bits(edges["theta"][0]).tolist() == bits(list(want.theta)).tolist()
    assert bits(corners[0]).tolist() == bits(list(want.corners)).tolist()
    # a flat Cr leaves the left edge missing: the line above really came from the third plane
    flat = np.full_like(cr, 128)
    e2, _, f2, _ = dmz.detect_edges(y[None], cb[None], flat[None])
    assert list(e2["found"][0]) == list(oracle.detect_edges(y, cb, flat).found) == [1, 0, 1, 1] and int(f2[0]) == 0
    # a batch mixing the three situations keeps every frame's planes apart
    ys, cbs, crs = np.stack([y, y, deck_frames(5, 1)[0]]), np.stack([cb, cb, cb]), np.stack([cr, flat, flat])
    e3, c3, f3, _ = dmz.detect_edges(ys, cbs, crs)
    for k in range(3):
        w = oracle.detect_edges(ys[k], cbs[k], crs[k])
        assert list(e3["found"][k]) == list(w.found) and int(f3[k]) == w.all_found, k
        if w.all_found:
            assert bits(c3[k]).tolist() == bits(list(w.corners)).tolist(), k


@pytest.fixture(scope="module")
def upright_cards(oracle):
    recs, cards = oracle.process_frames(deck_frames(300, 8), want_cards=True)
    return cards[recs["usable"] == 1][:4]


@pytest.mark.parametrize("orientation", [1, 2, 4])
def test_detect_other_orientations(dmz, oracle, upright_cards, orientation):
    """Portrait (1, 2) and landscape-left (4): other strip rectangles (dmz.cpp:279-341) and, downstream, another corner
    permutation.  Per-strip taps, edges and corners bit-equal to the oracle."""
    from util import oriented_frames
    boxes = oracle.detection_boxes(640, 480, orientation)
    frames = oriented_frames(upright_cards, orientation, boxes, seed=10 + orientation)
    edges, corners, found, lines = dmz.detect_edges(frames, orientation=orientation, want_lines=True)
    for k in range(len(frames)):
        for s, (x, y, w, h) in enumerate(boxes):
            ol = oracle.best_line(frames[k][y:y + h, x:x + w], s >= 2)
            gl = lines[k, s]
            for f in ("found", "max_votes", "low", "high", "n_edge_px"):
                assert int(gl[f]) == getattr(ol, f), (k, s, f)
            if ol.found:
                assert (int(gl["r"]), int(gl["n"])) == (ol.r, ol.n)
        want = oracle.detect_edges(frames[k], orientation=orientation)
        assert list(edges["found"][k]) == list(want.found) and int(found[k]) == want.all_found == 1
        assert bits(edges["rho"][k]).tolist() == bits(list(want.rho)).tolist()
        assert bits(corners[k]).tolist() == bits(list(want.corners)).tolist()


@pytest.mark.parametrize("orientation", [1, 2, 4])
def test_whole_path_other_orientations(dmz, oracle, upright_cards, orientation):
    """detect -> transform -> scan in the other three orientations against the oracle, and against the reference's own
    sources where that build travelled (oracle/_ref)."""
    from util import oriented_frames
    from oracle.binding import Oracle, available
    frames = oriented_frames(upright_cards, orientation, oracle.detection_boxes(640, 480, orientation), seed=20 + orientation)
    recs, cards = dmz.process_frames(frames, orientation=orientation, want_cards=True)
    checkers = [oracle] + ([Oracle("ref")] if available("ref") else [])
    for chk in checkers:
        want, wcards = chk.process_frames(frames, orientation=orientation, want_cards=True)
        assert want["all_found"].all() and (want["usable"] == 1).any()
        assert np.array_equal(cards, wcards)
        for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_offsets",
                  "h_pattern_offset"):
            assert np.array_equal(recs[f], want[f]), (chk.kind, f)
        assert np.array_equal(bits(recs["corners"]), bits(want["corners"]))
        assert np.array_equal(bits(recs["h_score"]), bits(want["h_score"]))
        assert np.abs(recs["scores"] - want["scores"]).max() <= TOL


def test_transform_card_upsample(dmz, oracle, deck, orecs):
    """upsample = true (dmz.cpp:473-481): a half-size chroma plane is warped with the luma-space corners halved."""
    rec = orecs[0]
    cb = np.ascontiguousarray(deck[:6, ::2, ::2])
    for o in (1, 2, 3, 4):
        got = dmz.transform_card(cb, rec["corners"][:6], orientation=o, upsample=True)
        for k in range(6):
            assert np.array_equal(got[k], oracle.transform_card(cb[k], rec["corners"][k], o, upsample=True)), (o, k)
    plain = dmz.transform_card(cb, rec["corners"][:6], upsample=False)
    assert not np.array_equal(plain, dmz.transform_card(cb, rec["corners"][:6], upsample=True))


def test_detect_1080p_portrait_wide_indices(dmz, oracle, upright_cards):
    """1920x1080 portrait: the side strips are 86 x 897, (w + 2)(h + 2) = 79112 > 65535, so padded pixel indices no
    longer fit 16 bits (the detect kernel's 32-bit work-list variant)."""
    from util import oriented_frames
    boxes = oracle.detection_boxes(1920, 1080, 1)
    assert max((w + 2) * (h + 2) for (_, _, w, h) in boxes) > 65535
    frames = oriented_frames(upright_cards[:2], 1, boxes, w=1920, h=1080, seed=5)
    edges, corners, found, lines = dmz.detect_edges(frames, orientation=1, want_lines=True)
    for k in range(len(frames)):
        for s, (x, y, w, h) in enumerate(boxes):
            ol = oracle.best_line(frames[k][y:y + h, x:x + w], s >= 2)
            gl = lines[k, s]
            for f in ("found", "max_votes", "low", "high", "n_edge_px"):
                assert int(gl[f]) == getattr(ol, f), (k, s, f, int(gl[f]), getattr(ol, f))
            if ol.found:
                assert (int(gl["r"]), int(gl["n"])) == (ol.r, ol.n)
        want = oracle.detect_edges(frames[k], orientation=1)
        assert int(found[k]) == want.all_found == 1
        assert bits(corners[k]).tolist() == bits(list(want.corners)).tolist()
    recs, cards = dmz.process_frames(frames, orientation=1, want_cards=True)
    want, wcards = oracle.process_frames(frames, orientation=1, want_cards=True)
    assert np.array_equal(cards, wcards) and np.array_equal(recs["card_check"], want["card_check"])
    assert np.array_equal(recs["v_y_offset"], want["v_y_offset"]) and np.array_equal(recs["usable"], want["usable"])


def test_homography_bits(dmz, oracle, golden):
    dst = np.array([0, 0, 427, 0, 0, 269, 427, 269], np.float32)
    src = golden["homog_src"]
    M = dmz.calc_persp_transform(src, np.tile(dst, (len(src), 1)))
    assert np.array_equal(M.view(np.uint32), golden["homog_M_bits"])  # reference (Eigen SSE2) bits
    rng = np.random.default_rng(9)
    src = (np.array([106, 105, 533, 105, 106, 374, 533, 374], np.float32) + rng.uniform(-30, 30, (4000, 8))).astype(np.float32)
    Mg = dmz.calc_persp_transform(src, np.tile(dst, (4000, 1)))
    Mo = np.stack([oracle.calc_persp_transform(s, dst) for s in src])
    assert np.array_equal(Mg.view(np.uint32), Mo.view(np.uint32))


def test_transform_card_exact(dmz, oracle, deck, orecs):
    rec, ocards = orecs
    cards = dmz.transform_card(deck, rec["corners"], valid=rec["all_found"].astype(np.uint8))
    assert np.array_equal(cards, ocards)
    # all four orientations and the chroma (upsample) variant on one frame
    for o in (1, 2, 3, 4):
        assert np.array_equal(dmz.transform_card(deck[:1], rec["corners"][:1], orientation=o)[0], oracle.transform_card(deck[0], rec["corners"][0], o))


def test_transform_card_outside_frame(dmz, oracle, deck):
    """Corners partly outside the image: BORDER_CONSTANT 0 taps (cv/warp.cpp:165 CV_WARP_FILL_OUTLIERS)."""
    c = np.array([[-40, -30, -20, 300, 500, -10, 700, 520]], np.float32)
    assert np.array_equal(dmz.transform_card(deck[:1], c)[0], oracle.transform_card(deck[0], c[0]))


def test_warp_identity_is_idempotent(dmz):
    """Size-independent property: warping a 428x270 image with the identity corner set returns it unchanged."""
    img = np.random.default_rng(1).integers(0, 256, (1, 270, 428)).astype(np.uint8)
    c = np.array([[0, 0, 0, 269, 427, 0, 427, 269]], np.float32)  # tl, bl, tr, br
    assert np.array_equal(dmz.transform_card(img, c)[0], img[0])


def test_warp_rounding_ties(dmz, oracle):
    """Unit-scale maps shifted by odd multiples of 1/64 px put (almost) every fixed-point coordinate on or next to a
    rounding tie of cvRound(fX * 32): the kernel's fast coordinates must hand exactly those to the exact sequence."""
    frame = np.random.default_rng(7).integers(0, 256, (1, 480, 640)).astype(np.uint8)
    for s in (1 / 64, 3 / 64, 33 / 64, 1 / 128, 1 / 64 + 2.0 ** -12, 1 / 64 - 2.0 ** -12, 0.5, 0.25):
        for sx, sy in ((s, s), (s, 0.0), (0.0, s)):
            c = np.array([[100 + sx, 100 + sy, 100 + sx, 369 + sy, 527 + sx, 100 + sy, 527 + sx, 369 + sy]], np.float32)  # tl, bl, tr, br
            assert np.array_equal(dmz.transform_card(frame, c)[0], oracle.transform_card(frame[0], c[0])), (s, sx, sy)
    # scale 1/2 and 2 (coordinates exact multiples of 16 resp. 64 thirty-seconds) with half-pixel origins
    for c in ([100.5, 100.5, 100.5, 235.0, 314.0, 100.5, 314.0, 235.0], [10.25, 10.25, 10.25, 470.75, 630.5, 10.25, 630.5, 470.75]):
        c = np.array([c], np.float32)
        assert np.array_equal(dmz.transform_card(frame, c)[0], oracle.transform_card(frame[0], c[0]))


def test_scan_cards(dmz, orecs):
    rec, ocards = orecs
    scans = dmz.scan_cards(ocards, valid=rec["all_found"].astype(np.uint8))
    for f in ("v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_pattern_offset", "h_offsets"):
        assert np.array_equal(scans[f], rec[f]), f
    assert np.abs(scans["v_score"] - rec["v_score"]).max() <= 1e-3  # sum of 27 probabilities
    assert np.array_equal(bits(scans["h_score"]), bits(rec["h_score"])) and np.array_equal(bits(scans["h_number_width"]), bits(rec["h_number_width"]))
    assert np.abs(scans["scores"] - rec["scores"]).max() <= TOL
    assert (rec["usable"] == 1).sum() >= 24, "the deck should mostly be usable"


def test_scan_edge_cases(dmz, oracle):
    rng = np.random.default_rng(4)
    good = oracle.process_frames(deck_frames(0, 1), want_cards=True)[1][0]
    cards = np.stack([np.full((270, 428), 175, np.uint8),              # blank card: vseg gate fails
                      rng.integers(0, 256, (270, 428)).astype(np.uint8),  # noise
                      good[::-1, ::-1].copy(),                           # upside-down card
                      good])
    scans = dmz.scan_cards(cards)
    for k in range(4):
        w = oracle.scan_card_image(cards[k])
        assert (scans["usable"][k], scans["upside_down"][k]) == (w.usable, w.upside_down), k
        assert scans["v_y_offset"][k] == w.vseg.y_offset and scans["v_pattern_type"][k] == w.vseg.pattern_type, k
    assert scans["upside_down"][2] == 1


def test_model_known_answer_vectors(dmz):
    """The reference's embedded KATs (models/generated/*.cpp pass*_()), replayed through the CUDA kernels, at the
    reference's own tolerance 1e-5."""
    def kat(name):
        meta = json.load(open(os.path.join(G, "kat_%s.json" % name)))
        data = np.fromfile(os.path.join(G, "kat_%s.bin" % name), "<f4")
        return {v["label"]: data[v["offset"]:v["offset"] + v["count"]] for v in meta["vectors"]}
    k = kat("modelm_befe75da")
    assert np.abs(dmz.vseg_model(k["test input"])[0] - k["test output"]).max() <= 1e-5
    for i, m in enumerate(("5c241121", "01266c1b", "b00bf70c")):
        k = kat("modelc_" + m)
        _, per_model = dmz.digit_models(k["test input"])
        assert np.abs(per_model[0, i] - k["test output"]).max() <= 1e-5, m


def test_vseg_rows_tensor_core_path(dmz, oracle):
    """vseg_probabilities_for_hstrip (n_vseg.cpp:39-47) through the row kernel the whole path uses -- row preparation +
    hidden layer as exact integer MMAs on the tensor cores (vseg_mma.cu) + logistic layer -- against the oracle row by
    row: real card rows, and rows built to stress the normalisation (flat rows: all x = 0; a two-level row: scale 255;
    rows whose minimum gradient is large; saturated steps; pure noise)."""
    _, cards = oracle.process_frames(deck_frames(32, 6), want_cards=True)
    rng = np.random.default_rng(11)
    synth = np.zeros((4, 270, 428), np.uint8)
    synth[0] = 77                                                     # flat: every row all-equal after the gradient
    synth[1] = rng.integers(0, 256, (270, 428))                      # noise
    synth[2] = (np.arange(428)[None, :] // 3 % 2 * 255).astype(np.uint8)  # saturated steps everywhere: large min gradient
    synth[2, ::2] = np.where(rng.random((135, 428)) < 0.1, 200, synth[2, ::2])
    synth[3] = 100
    synth[3, :, 200] = 101                                            # a single unit step: v in {0, 1}
    synth[3, 8::8, 300:320] = rng.integers(0, 256, (33, 20))
    allc = np.concatenate([cards, synth])
    got = dmz.vseg_rows(allc)
    worst = 0.0
    for c in range(allc.shape[0]):
        for row in range(0, 270, 4):
            want = oracle.vseg_row(allc[c], row)  # (p0, visa-like, amex-like)
            worst = max(worst, float(np.abs(got[c, row] - want[1:]).max()))
    assert worst <= 1e-5, worst
    assert not got[:, 1::4].any() and not got[:, 2::4].any() and not got[:, 3::4].any()  # rows outside the coarse set stay 0
    os.environ["B200_DMZ_VSEG_FP32"] = "1"
    try:
        fp32 = dmz.vseg_rows(allc)
    finally:
        del os.environ["B200_DMZ_VSEG_FP32"]
    assert np.abs(fp32 - got).max() <= 1e-5


def test_tensor_core_cnn_equals_fp32_cnn(dmz, oracle):
    """The tcgen05 CNN path (categorize_mma.cu: exact integer conv + pool, split-fp16 hidden layer) against the FP32 FMA
    kernel: loose patches at counts around the 32-slot group size (1, 31, 32, 33, 100), and whole frames in odd batch
    sizes (1, 3, 7: the last group holds one frame) -- probabilities within the 1e-4 contract, arg-max digits equal."""
    rng = np.random.default_rng(21)
    _, cards = oracle.process_frames(deck_frames(8, 4), want_cards=True)
    pool = np.concatenate([cards[k][150:177, x:x + 19][None] for k in range(4) for x in range(20, 380, 23)])
    for n in (1, 31, 32, 33, 100):
        patches = pool[rng.integers(0, len(pool), n)].copy()
        patches[::3] = rng.integers(0, 256, patches[::3].shape)
        ens, mods = dmz.categorize_patches(patches)
        os.environ["B200_DMZ_CNN_FP32"] = "1"
        try:
            ens32, mods32 = dmz.categorize_patches(patches)
        finally:
            del os.environ["B200_DMZ_CNN_FP32"]
        assert np.abs(ens - ens32).max() <= TOL and np.abs(mods - mods32).max() <= TOL, n
        sure = np.sort(ens32, axis=1)[:, -1] - np.sort(ens32, axis=1)[:, -2] > 1e-3
        assert np.array_equal(ens.argmax(1)[sure], ens32.argmax(1)[sure]), n
    for n in (1, 3, 7):
        frames = deck_frames(40, n)
        got = dmz.process_frames(frames)
        os.environ["B200_DMZ_CNN_FP32"] = "1"
        try:
            ref = dmz.process_frames(frames)
        finally:
            del os.environ["B200_DMZ_CNN_FP32"]
        for k in ("usable", "h_n_offsets", "v_y_offset"):
            assert np.array_equal(got[k], ref[k]), (n, k)
        assert np.abs(got["scores"] - ref["scores"]).max() <= TOL, n
        assert (got["scores"][got["usable"] == 0] == 0).all()


def test_categorize_patches(dmz, oracle, golden):
    patches = golden["card0_patches"]
    ens, mods = dmz.categorize_patches(patches)
    assert np.abs(ens - golden["card0_ensemble"]).max() <= TOL and np.abs(mods - golden["card0_models"]).max() <= TOL
    rng = np.random.default_rng(8)
    noise = rng.integers(0, 256, (100, 27, 19)).astype(np.uint8)
    noise[0] = 0      # all-equal patch: histogram has one bin, lut[0] = 0
    noise[1] = 255
    ens, mods = dmz.categorize_patches(noise)
    for i in range(100):
        e, m = oracle.digit_models(oracle.digit_patch_prep(noise[i]))
        assert np.abs(ens[i] - e).max() <= TOL and np.abs(mods[i] - m).max() <= TOL, i


def test_deinterleave_c2(dmz):
    """dmz_deinterleave_uint8_c2 (dmz.cpp:49-56 = cvSplit): vector path (aligned, w % 8 == 0) and the byte path."""
    rng = np.random.default_rng(2)
    for (n, h, w) in [(3, 240, 320), (2, 7, 13), (1, 1, 1), (5, 360, 640)]:
        planes = rng.integers(0, 256, (n, h, w, 2), dtype=np.uint8)
        c1, c2 = dmz.deinterleave_c2(planes)
        assert np.array_equal(c1, planes[..., 0]) and np.array_equal(c2, planes[..., 1]), (n, h, w)


def test_frame_scores_exact(dmz, oracle, golden, deck):
    """dmz_focus_score / dmz_brightness_score: float bits equal to the oracle and to the reference
    build's golden outputs, at 640x480 and at frame sizes that scale / clip the scoring rectangle."""
    frames = np.concatenate([deck_frames(int(i), 1) for i in golden["deck_idx"]])
    for full in (0, 1):
        f, b = dmz.frame_scores(frames, full)
        assert np.array_equal(bits(f), golden["deck_focus"][full]) and np.array_equal(bits(b), golden["deck_brightness"][full])
    rng = np.random.default_rng(5)
    for (w, h, n) in [(640, 480, 33), (1280, 720, 5), (1920, 1080, 2), (320, 240, 7), (641, 479, 3), (100, 50, 4)]:
        imgs = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        imgs[0] = 0            # flat frame: variance exactly 0
        if n > 2:
            imgs[1] = 255
            imgs[2] = (np.arange(w)[None, :] * 3 + np.arange(h)[:, None] * 5) & 255
        for full in (0, 1):
            f, b = dmz.frame_scores(imgs, full)
            wf = np.array([oracle.focus_score(x, full) for x in imgs], np.float32)
            wb = np.array([oracle.brightness_score(x, full) for x in imgs], np.float32)
            assert np.array_equal(bits(f), bits(wf)) and np.array_equal(bits(b), bits(wb)), (w, h, full)
    f, _ = dmz.frame_scores(deck[:8])
    assert f.min() > 5.0  # the synthetic deck is in focus by the reference's own measure


def test_expiry_digit_known_answer_and_parity(dmz, oracle, golden):
    """E0: prepare_image_for_cat + applyc_bf4dd6c8 (scan/expiry_categorize.cpp:37-109).  The
    reference's embedded KAT at its own 1e-5; random / structured character crops against the oracle at 1e-4."""
    meta = json.load(open(os.path.join(G, "kat_modelc_bf4dd6c8.json")))
    data = np.fromfile(os.path.join(G, "kat_modelc_bf4dd6c8.bin"), "<f4")
    k = {v["label"]: data[v["offset"]:v["offset"] + v["count"]] for v in meta["vectors"]}
    assert np.abs(dmz.expiry_digit_models(k["test input"])[0] - k["test output"]).max() <= 1e-5
    got = dmz.expiry_digits(golden["expiry_patches"])  # outputs of the reference's own SCAN_EXPIRY=1 build
    assert np.abs(got - golden["expiry_probs"]).max() <= TOL
    got = dmz.expiry_digit_models(golden["expiry_prep_bits"].view(np.float32).reshape(-1, 176))
    assert np.abs(got - golden["expiry_probs"]).max() <= TOL
    rng = np.random.default_rng(21)
    n = 203  # not a multiple of the 4 digits a CTA takes per iteration
    patches = rng.integers(0, 256, (n, 16, 11)).astype(np.uint8)
    patches[0] = 0       # flat patch: gradient 0 everywhere, single histogram bin
    patches[1] = 255
    patches[2:60] = (patches[2:60] // 32) * 9       # few grey levels: colour weights of the bilateral filter matter
    yy, xx = np.mgrid[0:16, 0:11]
    for i in range(60, 120):                         # stroke-like shapes
        cx, cy, r = rng.uniform(3, 8), rng.uniform(4, 12), rng.uniform(2, 5)
        ring = np.abs(np.hypot(xx - cx, (yy - cy) * 0.7) - r) < 1.0
        patches[i] = np.where(ring, 200 + rng.integers(0, 40), 30 + rng.integers(0, 30, (16, 11))).astype(np.uint8)
    got = dmz.expiry_digits(patches)
    assert got.shape == (n, 10) and np.abs(got.sum(1) - 1).max() < 1e-5
    for i in range(n):
        want = oracle.expiry_digit_model(oracle.expiry_patch_prep(patches[i]))
        assert np.abs(got[i] - want).max() <= TOL, i
    # prepared-input entry, batched, against the oracle's model alone
    prepared = rng.random((37, 176)).astype(np.float32)
    got = dmz.expiry_digit_models(prepared)
    for i in range(37):
        assert np.abs(got[i] - oracle.expiry_digit_model(prepared[i])).max() <= TOL, i
    assert got.argmax(1).tolist() == [int(oracle.expiry_digit_model(prepared[i]).argmax()) for i in range(37)]


def test_tensor_core_expiry_cnn_equals_fp32(dmz):
    """E0 with layer 2 on the tensor cores (expiry_mma.cu, split-fp16 operands) against the FP32 kernel: crop counts around the
    seven-crop batch (1, 6, 7, 8, 50: the last batch is partial), noise and flat crops; probabilities within 1e-4."""
    rng = np.random.default_rng(5)
    for n in (1, 6, 7, 8, 50):
        crops = rng.integers(0, 256, (n, 16, 11)).astype(np.uint8)
        crops[::4] = (rng.integers(0, 2, (len(crops[::4]), 16, 11)) * 180 + 40).astype(np.uint8)
        if n > 2:
            crops[2] = 99
        got = dmz.expiry_digits(crops)
        os.environ["B200_DMZ_EXPIRY_FP32"] = "1"
        try:
            ref = dmz.expiry_digits(crops)
        finally:
            del os.environ["B200_DMZ_EXPIRY_FP32"]
        assert np.abs(got - ref).max() <= TOL, (n, float(np.abs(got - ref).max()))
        assert np.abs(got.sum(1) - 1).max() < 1e-5


def test_best_expiry_seg(dmz, golden):
    """best_expiry_seg on synthetic expiry cards -- groups identical to the golden outputs of the
    reference's SCAN_EXPIRY=1 build (and to that build itself when oracle/_ref travelled here); the |Scharr| plane
    against the golden checksum."""
    from util import expiry_card
    base, yo0 = golden["deck_card0"], int(golden["deck_records"]["v_y_offset"][0])
    seeds = [int(s) for s in golden["expiry_seg_seeds"]]
    made = [expiry_card(base, yo0, sd) for sd in seeds]
    cards, yos = np.stack([m[0] for m in made]), np.array([m[1] for m in made], np.uint16)
    groups, counts, dropped, sob = dmz.best_expiry_seg(cards, yos, max_groups=16, want_sobel=True)
    assert not dropped.any() and np.array_equal(counts, golden["expiry_seg_counts"])
    flat = lambda g: np.concatenate([[g["top"], g["left"], g["width"], g["height"], g["character_width"], g["pattern"], g["n_rects"]],
                                     np.stack([g["rect_top"], g["rect_left"]], 1).ravel()]).astype(np.int32)
    got = [flat(groups[i, k]) for i in range(len(seeds)) for k in range(counts[i])]
    assert np.array_equal(np.array(got, np.int32).reshape(-1, 17), golden["expiry_seg_groups"])
    i = seeds.index(1001)
    part = sob[i, int(yos[i]) + 27:].astype(np.uint64).ravel()
    assert np.uint64((part * np.arange(1, part.size + 1, dtype=np.uint64)).sum()) == golden["expiry_scharr_check"]
    # max_groups smaller than what a card yields: the surplus is counted, not written past the end
    many = int(np.argmax(counts))
    if counts[many] > 1:
        g1, c1, d1 = dmz.best_expiry_seg(cards[many:many + 1], yos[many:many + 1], max_groups=1)
        assert c1[0] == 1 and d1[0] == counts[many] - 1 and flat(g1[0, 0]).tolist() == flat(groups[many, 0]).tolist()
    # live against the reference build on a larger, fresh sample (ragged vs the 2048-card chunking is covered by n = 2500)
    from oracle.binding import Oracle, available
    if available("refx"):
        rx = Oracle("refx")
        made = [expiry_card(base, yo0, sd) for sd in range(20000, 22500)]
        cards, yos = np.stack([m[0] for m in made]), np.array([m[1] for m in made], np.uint16)
        groups, counts, dropped = dmz.best_expiry_seg(cards, yos, max_groups=16)
        bad = 0
        for j in range(0, len(made), 5):
            want = rx.best_expiry_seg(cards[j], int(yos[j]))
            gotj = np.array([flat(groups[j, k]) for k in range(counts[j])], np.int32).reshape(-1, 17)
            bad += not (gotj.shape == want.shape and np.array_equal(gotj, want))
        assert bad == 0
        assert (counts > 0).sum() > 500


def test_whole_path_records(dmz, deck, orecs):
    rec, ocards = orecs
    got, cards = dmz.process_frames(deck, want_cards=True)
    assert np.array_equal(cards, ocards)
    for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_offsets", "h_pattern_offset"):
        assert np.array_equal(got[f], rec[f]), f
    assert np.array_equal(bits(got["corners"]), bits(rec["corners"]))
    assert np.abs(got["scores"] - rec["scores"]).max() <= TOL


def test_whole_path_against_reference_fixture(dmz, golden):
    """Golden records produced by the reference's own sources (tools/make_ref_golden.py)."""
    idx = golden["deck_idx"]
    frames = np.concatenate([deck_frames(int(i), 1) for i in idx])
    got, cards = dmz.process_frames(frames, want_cards=True)
    want = golden["deck_records"]
    assert np.array_equal(cards[0], golden["deck_card0"])
    for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_offsets"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(bits(got["corners"]), bits(want["corners"]))
    assert np.abs(got["scores"] - want["scores"]).max() <= TOL


def test_strided_input_and_ragged_batches(dmz, deck, orecs):
    """Row / frame strides and batch sizes that are not multiples of anything; records must not depend on batching."""
    rec = orecs[0]
    for n in (1, 3, 17):
        got = dmz.process_frames(deck[:n])
        assert np.array_equal(got["card_check"], rec["card_check"][:n])
    padded = np.zeros((5, 500, 704), np.uint8)
    padded[:, :480, :640] = deck[:5]
    import ctypes as C
    recs = np.zeros(5, got.dtype)
    rc = dmz.lib.b200_process_frames_batch(dmz.ctx, padded.ctypes.data_as(C.c_void_p), 704, 704 * 500, 640, 480, 5, 3, 0,
                                           recs.ctypes.data_as(C.c_void_p), None)
    assert rc == 0 and np.array_equal(recs["card_check"], rec["card_check"][:5])


def test_cropped_upload_falls_back_to_full_frames(pkg, oracle):
    """Host-buffer path: with a zero margin some quads of a jitter-14 deck reach outside the uploaded rectangle;
    those frames must be redone from the whole frame and every record must still equal the oracle's."""
    frames = deck_frames(300, 24, jitter=14.0)
    want = oracle.process_frames(frames)
    d = pkg.Dmz(device=0)
    try:
        for margin in (0, 8, -1):
            d.set_crop_margin(margin)
            before = d.full_frame_redos
            got = d.process_frames(frames)
            for f in ("found", "all_found", "card_check", "v_y_offset", "usable", "h_offsets"):
                assert np.array_equal(got[f], want[f]), (margin, f)
            assert np.array_equal(bits(got["corners"]), bits(want["corners"]))
            if margin == 0:
                assert d.full_frame_redos > before, "the test deck should exercise the fallback"
            if margin < 0:
                assert d.full_frame_redos == before
    finally:
        d.close()


def test_720p_frames(dmz, oracle):
    fr = deck_frames(0, 4, 1280, 720)
    want = oracle.process_frames(fr)
    got = dmz.process_frames(fr)
    for f in ("found", "all_found", "card_check", "v_y_offset", "usable"):
        assert np.array_equal(got[f], want[f]), f
    assert want["all_found"].all()


def test_1080p_detect_and_path(dmz, oracle):
    """BASELINE configs[3] sizes: the 1080p strips (875x64, 86x543) exceed shared memory for dx/dy, so the detect
    kernel keeps the gradients in a global scratch slab; results must not change."""
    fr = deck_frames(0, 2, 1920, 1080)
    edges, corners, found, lines = dmz.detect_edges(fr, want_lines=True)
    boxes = oracle.detection_boxes(1920, 1080)
    for k in range(2):
        for s_, (x, y, w, h) in enumerate(boxes):
            ol = oracle.best_line(fr[k][y:y + h, x:x + w], s_ >= 2)
            gl = lines[k, s_]
            for f in ("found", "max_votes", "low", "high", "n_edge_px"):
                assert int(gl[f]) == getattr(ol, f), (k, s_, f)
            if ol.found:
                assert (int(gl["r"]), int(gl["n"])) == (ol.r, ol.n)
    want = oracle.process_frames(fr)
    got = dmz.process_frames(fr)
    for f in ("found", "all_found", "card_check", "v_y_offset", "usable"):
        assert np.array_equal(got[f], want[f]), f


def test_second_device_in_one_process(pkg, dmz, deck, orecs):
    """Contexts on two devices in one process (kernel attributes such as the dynamic shared-memory opt-in are per device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rec, _ = orecs
    other = pkg.Dmz(device=1)
    got = other.process_frames(deck)
    for f in ("all_found", "card_check", "v_y_offset", "usable", "h_offsets"):
        assert np.array_equal(got[f], rec[f]), f


def test_bad_arguments_fail_cleanly(dmz, pkg):
    with pytest.raises(pkg.B200Error):
        dmz.process_frames(np.zeros((1, 16, 16), np.uint8))  # frame too small for detection strips


def _run_dropin_caller(exe, oracle, tmp_path):
    frames = deck_frames(16, 8)
    fin, fout = str(tmp_path / "frames.bin"), str(tmp_path / "out.bin")
    frames.tofile(fin)
    subprocess.check_call([exe, fin, "8", "640", "480", fout])
    dt = np.dtype([("rec", "<i4", 8), ("corners", "<f4", 8), ("scores", "<f4", 160), ("digits", "u1", 16), ("fb", "<f4", 2), ("fmt", "<u4", 6)])
    got = np.fromfile(fout, dt)
    want, cards = oracle.process_frames(frames, want_cards=True)

    def wsum(a):  # compat_main's weighted_sum over the bytes of a
        b = np.ascontiguousarray(a).view(np.uint8).ravel().astype(np.uint64)
        return int((b * np.arange(1, b.size + 1, dtype=np.uint64)).sum() & 0xFFFFFFFF)

    # the colour card (dmz_YCbCr_to_RGB on the card and its upsampled chroma), RGBA -> R, and the Cython stencil taps
    for k in range(8):
        if not want["all_found"][k]:
            continue
        cb, cr = np.ascontiguousarray(frames[k][::2, ::2]), np.ascontiguousarray(255 - frames[k][::2, ::2])
        cbc = oracle.transform_card(cb, want["corners"][k], upsample=True)
        crc = oracle.transform_card(cr, want["corners"][k], upsample=True)
        rgb, rgba = oracle.ycbcr_to_rgb(cards[k], cbc, crc, 3), oracle.ycbcr_to_rgb(cards[k], cbc, crc, 4)
        rows = sum((r + 1) * wsum(rgb[r]) for r in range(270)) & 0xFFFFFFFF
        assert got["fmt"][k, 0] == rows, k
        assert got["fmt"][k, 1] == wsum(rgba) and got["fmt"][k, 2] == wsum(rgba[..., 0]), k
        for kind in range(3):
            assert got["fmt"][k, 3 + kind] == wsum(oracle.stencil3(cards[k], kind)), (k, kind)
    assert np.array_equal(bits(got["fb"][:, 0]), bits(np.array([oracle.focus_score(f) for f in frames], np.float32)))
    assert np.array_equal(bits(got["fb"][:, 1]), bits(np.array([oracle.brightness_score(f) for f in frames], np.float32)))
    assert np.array_equal(got["rec"][:, 0], want["all_found"])
    assert np.array_equal(bits(got["corners"]), bits(want["corners"]))
    assert np.array_equal(got["rec"][:, 7].astype(np.uint32), want["card_check"])
    assert np.array_equal(got["rec"][:, 2], want["usable"]) and np.array_equal(got["rec"][:, 4], want["v_y_offset"])
    s = oracle.scanner_new()
    done = False
    for k in range(8):
        if not done:  # once the number is complete the reference stops collecting scores (scan.cpp:43, frame.cpp:49)
            assert np.abs(got["scores"][k] - want["scores"][k]).max() <= TOL
        oracle.scanner_add_frame(s, cards[k])
        done, digits = oracle.scanner_result(s)
        assert got["rec"][k, 5] == int(done)
        if done:
            assert got["digits"][k][: len(digits)].tolist() == digits.tolist()
    assert done
    oracle.scanner_free(s)


def test_cxx_dropin_layer(dmz, oracle, tmp_path):
    """A caller written against the reference's dmz.h / scan.h shape, linked with libb200dmz.so."""
    exe = str(tmp_path / "compat_main")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "compat_main.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "card.io-dmz_b200"), "-lb200dmz",
                           "-Wl,-rpath," + os.path.join(ROOT, "card.io-dmz_b200"), "-ldl"])
    _run_dropin_caller(exe, oracle, tmp_path)


def test_cxx_dropin_reference_headers(dmz, oracle, tmp_path):
    """The same caller compiled against the reference's own UNMODIFIED dmz.h + scan/scan.h (vendored opencv2 / Eigen
    headers; oracle/Makefile target `dropin`, built where /root/reference exists and shipped under oracle/_ref): its
    structs, inline Eigen accessors and mangled call sites are the reference's, the library behind them is this repo's."""
    exe = os.path.join(ROOT, "oracle", "_ref", "compat_main_refhdr")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/compat_main_refhdr not present (built only where /root/reference exists)")
    _run_dropin_caller(exe, oracle, tmp_path)


def test_cxx_dropin_expiry(dmz, oracle, tmp_path):
    """scanner_add_frame_with_expiry(scan_expiry = true) through the C++ drop-in layer: GPU segmentation + expiry digit CNN,
    cross-frame aggregation in the caller's ScannerState -- against the reference's SCAN_EXPIRY=1 build on the same cards
    (needs oracle/_ref/libdmz_ref_expiry.so, which travels with the repo snapshot)."""
    from oracle.binding import Oracle, available
    from util import expiry_glyph
    if not available("refx"):
        pytest.skip("oracle/_ref/libdmz_ref_expiry.so not present")
    rx = Oracle("refx")
    exe = str(tmp_path / "compat_expiry_main")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "compat_expiry_main.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "card.io-dmz_b200"), "-lb200dmz",
                           "-Wl,-rpath," + os.path.join(ROOT, "card.io-dmz_b200"), "-ldl"])
    recs, cards = oracle.process_frames(deck_frames(16, 8), want_cards=True)
    dt = np.dtype([("head", "<i4", 8), ("groups", [("meta", "<i4", 4), ("rows", "<f4", 40)], 8)])
    for case, (txt, fg, jump) in enumerate((("08/27", 40, 0), ("11/29", 250, 0), ("03/30", 40, 7), ("05/28", 250, 3))):
        stamped = []
        for k in range(8):
            yo = int(recs["v_y_offset"][k])
            c = cards[k].copy()
            x0 = 70 + (jump if k % 2 else 0)
            for i, ch in enumerate(txt):
                reg = c[min(yo + 67, 250):min(yo + 67, 250) + 15, x0 + 12 * i:x0 + 12 * i + 9]
                reg[expiry_glyph(ch) > 0] = fg
            stamped.append(c)
        stamped = np.ascontiguousarray(np.stack(stamped))
        fin, fout = str(tmp_path / ("cards%d.bin" % case)), str(tmp_path / ("out%d.bin" % case))
        stamped.tofile(fin)
        subprocess.check_call([exe, fin, "8", fout])
        got = np.fromfile(fout, dt)
        rs = rx.scanner_new()
        seen = 0
        for k in range(8):
            scan, _ = rx.scanner_add_frame_with_expiry(rs, stamped[k], True)
            done, _, rm, ry = rx.scanner_result_expiry(rs)
            (m, y), meta, scores = rx.scanner_expiry_peek(rs)
            h = got["head"][k]
            assert (h[0], h[1]) == (scan.usable, scan.upside_down), (case, k)
            assert (h[2], h[3]) == (m, y) and h[7] == len(meta), (case, k, h, m, y, len(meta))
            assert (h[4], h[5], h[6]) == (int(done), rm, ry), (case, k)
            for g in range(min(len(meta), 8)):
                assert np.array_equal(got["groups"][k]["meta"][g], meta[g]), (case, k, g)
                assert np.abs(got["groups"][k]["rows"][g].reshape(4, 10) - scores[g]).max() <= TOL, (case, k, g)
            seen = max(seen, len(meta))
        rx.scanner_free(rs)
        assert seen >= 1
