"""The plain-C oracle against (a) the reference's embedded model known-answer vectors and (b) golden fixtures
produced by running the reference's own sources (tools/make_ref_golden.py -> tests/golden/ref_golden.npz).
Integer / index results bit-exact; float probabilities within 1e-5 (the reference's own KAT tolerance)."""
import json
import os

import numpy as np
import pytest

from util import ROOT, deck_frames

G = os.path.join(ROOT, "tests", "golden")


def kat(name):
    meta = json.load(open(os.path.join(G, "kat_%s.json" % name)))
    data = np.fromfile(os.path.join(G, "kat_%s.bin" % name), "<f4")
    return {v["label"]: data[v["offset"]:v["offset"] + v["count"]] for v in meta["vectors"]}, meta["tolerance_abs"]


def test_kat_vseg_mlp(oracle):
    k, tol = kat("modelm_befe75da")  # models/generated/modelm_befe75da.cpp:1793-1832
    assert np.abs(oracle.vseg_model(k["test input"]) - k["test output"]).max() <= tol


@pytest.mark.parametrize("idx,model", list(enumerate(["5c241121", "01266c1b", "b00bf70c"])))
def test_kat_digit_cnn(oracle, idx, model):
    k, tol = kat("modelc_" + model)  # models/generated/modelc_*.cpp:1944-2036
    _, per_model = oracle.digit_models(k["test input"])
    assert np.abs(per_model[idx] - k["test output"]).max() <= tol


def test_kat_expiry_cnn(oracle):
    """E0: applyc_bf4dd6c8's embedded per-layer vectors (models/expiry/modelc_bf4dd6c8.cpp:13507-13560)."""
    k, tol = kat("modelc_bf4dd6c8")
    out, l1, l2, hid = oracle.expiry_digit_model(k["test input"], taps=True)
    assert np.abs(l1.ravel() - k["test output layer 1"]).max() <= tol
    assert np.abs(l2.ravel() - k["test output layer 2"]).max() <= tol
    assert np.abs(hid.ravel() - k["test output layer 3"]).max() <= tol
    assert np.abs(out - k["test output"]).max() <= tol


def test_expiry_digit_golden(oracle, golden):
    """E0 against outputs of the reference's own prepare_image_for_cat / applyc_bf4dd6c8 (SCAN_EXPIRY=1 build)."""
    for i, patch in enumerate(golden["expiry_patches"]):
        prep = oracle.expiry_patch_prep(patch)
        assert np.array_equal(prep.view(np.uint32), golden["expiry_prep_bits"][i]), i
        assert np.abs(oracle.expiry_digit_model(prep) - golden["expiry_probs"][i]).max() <= 1e-5, i


def test_frame_scores(oracle, golden):
    """dmz_focus_score / dmz_brightness_score (dmz.cpp:114-195) against the reference build's outputs, bit for bit."""
    frames = np.concatenate([deck_frames(int(i), 1) for i in golden["deck_idx"]])
    for full in (0, 1):
        f = np.array([oracle.focus_score(x, full) for x in frames], np.float32).view(np.uint32)
        b = np.array([oracle.brightness_score(x, full) for x in frames], np.float32).view(np.uint32)
        assert np.array_equal(f, golden["deck_focus"][full]) and np.array_equal(b, golden["deck_brightness"][full])
    srng = np.random.default_rng(99)
    for (w, h) in [(1280, 720), (320, 240), (641, 479)]:
        img = srng.integers(0, 256, (h, w), dtype=np.uint8)
        for full in (0, 1):
            assert np.array_equal(oracle.scoring_rect(w, h, full), golden["score_rect_%dx%d" % (w, h)][full])
            got = np.array([oracle.focus_score(img, full), oracle.brightness_score(img, full)], np.float32).view(np.uint32)
            assert np.array_equal(got, golden["score_%dx%d" % (w, h)][full]), (w, h, full)


def test_detection_boxes(oracle, golden):
    for key in golden.files:
        if key.startswith("boxes_"):
            w, h, o = key[6:].replace("x", "_").replace("o", "").split("_")
            assert np.array_equal(oracle.detection_boxes(int(w), int(h), int(o)), golden[key]), key


def test_detect_strips(oracle, golden):
    fields = [str(f) for f in golden["line_fields"]]
    for i in range(int(golden["n_strips"])):
        img = golden["strip%d_img" % i]
        vert = int(golden["strip%d_vertical" % i])
        dx, dy = oracle.sobel7(img)
        assert np.array_equal(dx, golden["strip%d_dx" % i]) and np.array_equal(dy, golden["strip%d_dy" % i]), i
        edges, lo, hi = oracle.adaptive_canny(img, dx, dy)
        assert np.array_equal(edges, golden["strip%d_edges" % i]), i
        l = oracle.best_line(img, vert)
        want = dict(zip(fields, golden["strip%d_line" % i]))
        for f in ("found", "low", "high", "n_edge_px"):
            assert getattr(l, f) == int(want[f]), (i, f)
        if l.found:
            assert (l.r, l.n) == (int(want["r"]), int(want["n"])), i
        bits = np.array([l.rho, l.theta], np.float32).view(np.uint32)
        assert np.array_equal(bits, golden["strip%d_rho_theta_bits" % i]), i


def test_homography_bits(oracle, golden):
    dst = np.array([0, 0, 427, 0, 0, 269, 427, 269], np.float32)
    for src, bits in zip(golden["homog_src"], golden["homog_M_bits"]):
        assert np.array_equal(oracle.calc_persp_transform(src, dst).view(np.uint32), bits)


def test_whole_path_on_deck(oracle, golden):
    idx = golden["deck_idx"]
    frames = np.concatenate([deck_frames(int(i), 1) for i in idx])
    assert np.array_equal(frames[0], golden["deck_frame0"]), "deck generator drifted from the committed fixture"
    recs, cards = oracle.process_frames(frames, want_cards=True)
    want = golden["deck_records"]
    assert np.array_equal(cards[0], golden["deck_card0"])
    for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets",
              "h_offsets", "h_pattern_offset"):
        assert np.array_equal(recs[f], want[f]), f
    for f in ("rho", "theta", "corners", "h_score", "h_number_width"):  # mirrored float order: bit-exact
        assert np.array_equal(recs[f].view(np.uint32), want[f].view(np.uint32)), f
    assert np.abs(recs["v_score"] - want["v_score"]).max() <= 1e-3
    assert np.abs(recs["scores"] - want["scores"]).max() <= 1e-5


def test_card0_stage_taps(oracle, golden):
    card = golden["deck_card0"]
    rows = np.stack([oracle.vseg_row(card, r) for r in range(0, 270, 9)])
    assert np.abs(rows - golden["card0_vseg_rows"]).max() <= 1e-5
    for p, prep, ens, mod in zip(golden["card0_patches"], golden["card0_patch_prep"], golden["card0_ensemble"], golden["card0_models"]):
        assert np.array_equal(oracle.digit_patch_prep(p).view(np.uint32), prep.view(np.uint32))  # integer ops + one float multiply
        e, m = oracle.digit_models(prep)
        assert np.abs(e - ens).max() <= 1e-5 and np.abs(m - mod).max() <= 1e-5


def test_transform_orientations(oracle, golden):
    frame = golden["deck_frame0"]
    corners = golden["deck_records"]["corners"][0]
    for o in (1, 2, 3, 4):
        card = oracle.transform_card(frame, corners, o)
        chk = np.uint32((card.astype(np.uint64).ravel() * np.arange(1, 428 * 270 + 1, dtype=np.uint64)).sum() & 0xFFFFFFFF)
        assert chk == golden["card0_orient%d_check" % o], o


def test_scanner_session(oracle, golden):
    sess = deck_frames(16, 8)
    _, cards = oracle.process_frames(sess, want_cards=True)
    s = oracle.scanner_new()
    flags = []
    for k in range(8):
        oracle.scanner_add_frame(s, cards[k])
        done, digits = oracle.scanner_result(s)
        flags.append(int(done))
    a15, a16, cnt = oracle.scanner_peek(s)
    oracle.scanner_free(s)
    assert flags == golden["session_complete"].tolist()
    assert digits.tolist() == golden["session_digits"].tolist()
    assert np.array_equal(cnt, golden["session_counts"])
    assert np.abs(a16 - golden["session_agg16"]).max() <= 1e-5


def test_edge_cases(oracle):
    # flat frame: nothing detected, record stays empty
    flat = np.full((1, 480, 640), 90, np.uint8)
    r = oracle.process_frames(flat)
    assert r["all_found"][0] == 0 and r["found"].sum() == 0 and r["card_check"][0] == 0
    # uniform noise: 'high' exceeds every magnitude -> no edges
    noise = np.random.default_rng(5).integers(0, 256, (28, 389)).astype(np.uint8)
    l = oracle.best_line(noise, 0)
    assert l.found == 0 and l.n_edge_px == 0
    # upside-down card: digit row in the top half
    card = np.full((270, 428), 175, np.uint8)
    s = oracle.scan_card_image(card)
    assert s.usable == 0
    assert oracle.luhn([4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]) and not oracle.luhn([4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2])
    assert oracle.card_type([4] + [0] * 15) == 4 and oracle.card_type([3, 4] + [0] * 13) == 2 and oracle.card_type([9] * 16) == 0
