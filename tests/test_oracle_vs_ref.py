"""Live, stage-by-stage comparison of the plain-C oracle with the reference's own sources (oracle/_ref).  Skipped
where oracle/_ref is not built.  This is what pins the oracle beyond the committed fixtures: fresh random inputs."""
import numpy as np
import pytest

from util import deck_frames, synthetic_strip


def fields(s):
    out = {}
    for n, _ in s._fields_:
        v = getattr(s, n)
        out[n] = list(v) if hasattr(v, "__len__") else v
    return out


def test_reference_kats(ref):
    assert ref.lib.ref_run_kats() == 0b1111  # passm_befe75da, passc_5c241121, passc_01266c1b, passc_b00bf70c


def test_struct_sizes(ref):
    # NVerticalSegmentation, NHorizontalSegmentation, NumberScores, dmz_edges, dmz_corner_points
    assert [ref.lib.ref_sizeof(i) for i in (0, 1, 2, 6, 7)] == [28, 48, 640, 48, 32]


def test_detect_stage_random_strips(ref, oracle):
    rng = np.random.default_rng(11)
    for t in range(24):
        vert = t % 2
        w, h = ((38, 241) if vert else (389, 28)) if t < 16 else (int(rng.integers(8, 80)), int(rng.integers(8, 80)))
        img = synthetic_strip(rng, w, h, vert, ["edge", "edge", "noise"][t % 3])
        dxr, dyr = ref.sobel7(img)
        dxo, dyo = oracle.sobel7(img)
        assert np.array_equal(dxr, dxo) and np.array_equal(dyr, dyo)
        er, eo = ref.adaptive_canny(img, dxr, dyr), oracle.adaptive_canny(img, dxr, dyr)
        assert er[1:] == eo[1:] and np.array_equal(er[0], eo[0])
        lr, lo = fields(ref.best_line(img, vert)), fields(oracle.best_line(img, vert))
        lr.pop("max_votes"); lo.pop("max_votes")
        if not lr["found"]:
            lr.pop("r"), lr.pop("n"), lo.pop("r"), lo.pop("n")
        assert lr == lo


def test_homography_bits_random(ref, oracle):
    rng = np.random.default_rng(12)
    dst = np.array([0, 0, 427, 0, 0, 269, 427, 269], np.float32)
    for _ in range(3000):
        src = (np.array([106, 105, 533, 105, 106, 374, 533, 374], np.float32) + rng.uniform(-30, 30, 8)).astype(np.float32)
        assert np.array_equal(ref.calc_persp_transform(src, dst).view(np.uint32), oracle.calc_persp_transform(src, dst).view(np.uint32))


def test_whole_path_deck(ref, oracle):
    frames = deck_frames(100, 12)
    rr, rc = ref.process_frames(frames, want_cards=True)
    po, pc = oracle.process_frames(frames, want_cards=True)
    assert np.array_equal(rc, pc)
    for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_offsets", "h_pattern_offset"):
        assert np.array_equal(rr[f], po[f]), f
    for f in ("corners", "h_score", "h_number_width"):
        assert np.array_equal(rr[f].view(np.uint32), po[f].view(np.uint32)), f
    ok = rr["all_found"] == 1
    assert np.array_equal(rr["rho"][ok].view(np.uint32), po["rho"][ok].view(np.uint32))
    assert np.abs(rr["scores"] - po["scores"]).max() <= 1e-5


def test_chroma_fallback(ref, oracle):
    """Y plane flat on the left -> the left edge must come from Cb (rho doubled), dmz.cpp:351-367."""
    f = deck_frames(5, 1)[0]
    cb = np.ascontiguousarray(f[::2, ::2])  # half-resolution copy with real structure
    cr = np.full((240, 320), 128, np.uint8)
    y = f.copy()
    y[:, :160] = 60  # wipe the left edge from Y only
    dr, do = ref.detect_edges(y, cb, cr), oracle.detect_edges(y, cb, cr)
    assert fields(dr)["found"] == fields(do)["found"]
    assert dr.all_found == do.all_found
    if dr.all_found:
        assert np.array_equal(np.array(dr.corners, np.float32).view(np.uint32), np.array(do.corners, np.float32).view(np.uint32))


def test_expiry_digit_against_reference_build(refx, oracle):
    """E0: prepare_image_for_cat + applyc_bf4dd6c8 of the SCAN_EXPIRY=1 reference build vs the restatement: the
    prepared patch bit for bit, the probabilities within 1e-5."""
    assert refx.lib.ref_run_kats() == 0b1111
    rng = np.random.default_rng(1)
    for t in range(200):
        patch = rng.integers(0, 256, (16, 11), dtype=np.uint8)
        if t % 3 == 1:
            patch = (patch // 32 * 9).astype(np.uint8)
        if t % 3 == 2 and t % 2:
            patch[:] = rng.integers(0, 256)
        a, b = refx.expiry_patch_prep(patch), oracle.expiry_patch_prep(patch)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), t
        assert np.abs(refx.expiry_digit_model(a) - oracle.expiry_digit_model(b)).max() <= 1e-5, t


def _upright_cards(oracle, first=300, n=6):
    recs, cards = oracle.process_frames(deck_frames(first, n), want_cards=True)
    return cards[recs["usable"] == 1]


def test_whole_path_all_orientations(ref, oracle):
    """FrameOrientation 1, 2, 4 change the strip geometry (dmz.cpp:279-341) and the corner permutation
    (dmz.cpp:446-471): the restatement must follow the reference in every one of them."""
    from util import oriented_frames
    cards = _upright_cards(oracle)[:3]
    for o in (1, 2, 3, 4):
        assert np.array_equal(ref.detection_boxes(640, 480, o), oracle.detection_boxes(640, 480, o))
        frames = oriented_frames(cards, o, oracle.detection_boxes(640, 480, o), seed=o)
        rr, rc = ref.process_frames(frames, orientation=o, want_cards=True)
        po, pc = oracle.process_frames(frames, orientation=o, want_cards=True)
        assert rr["all_found"].all(), o
        assert np.array_equal(rc, pc), o
        for f in ("found", "all_found", "card_check", "v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_offsets"):
            assert np.array_equal(rr[f], po[f]), (o, f)
        assert np.array_equal(rr["corners"].view(np.uint32), po["corners"].view(np.uint32)), o
        assert np.abs(rr["scores"] - po["scores"]).max() <= 1e-5


def test_transform_card_upsample(ref, oracle):
    """upsample = true: the sample is a half-size chroma plane and the corners are halved (dmz.cpp:473-481)."""
    frames = deck_frames(40, 2)
    rec = oracle.process_frames(frames)
    for k in range(2):
        cb = np.ascontiguousarray(frames[k][::2, ::2])
        for o in (1, 2, 3, 4):
            a = ref.transform_card(cb, rec["corners"][k], o, upsample=True)
            b = oracle.transform_card(cb, rec["corners"][k], o, upsample=True)
            assert np.array_equal(a, b), (k, o)
            assert not np.array_equal(a, oracle.transform_card(cb, rec["corners"][k], o, upsample=False))


def chroma_case():
    """Y misses the left and the top edge; Cb has the top edge only; Cr has every edge: top comes from Cb, left from Cr."""
    f = deck_frames(5, 1)[0]
    half = np.ascontiguousarray(f[::2, ::2])
    y = f.copy()
    y[:, :160] = 60
    y[:100, :] = 60
    cb = half.copy()
    cb[:, :80] = 60
    return y, cb, half.copy()


def test_chroma_cr_supplies_edge(ref, oracle):
    y, cb, cr = chroma_case()
    dr, do = ref.detect_edges(y, cb, cr), oracle.detect_edges(y, cb, cr)
    assert fields(dr)["found"] == [1, 1, 1, 1] == fields(do)["found"]
    assert np.array_equal(np.array(dr.rho, np.float32).view(np.uint32), np.array(do.rho, np.float32).view(np.uint32))
    assert np.array_equal(np.array(dr.corners, np.float32).view(np.uint32), np.array(do.corners, np.float32).view(np.uint32))
    # without Cr the left edge stays missing: it really is the third plane that supplies it
    flat = np.full_like(cr, 128)
    assert fields(ref.detect_edges(y, cb, flat))["found"] == [1, 0, 1, 1]
