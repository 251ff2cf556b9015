"""best_expiry_seg -- the decision logic of card.io-dmz_b200/csrc/expiry_seg_core.h, compiled here
with g++ (tests/expiry_host.cpp) so it can be checked on a machine without a GPU: against golden outputs of the
reference's SCAN_EXPIRY=1 build (tests/golden/ref_golden.npz, tools/make_ref_golden.py), live against that build when
it is present, and its std::sort restatement against the real std::sort (ties included).  The GPU kernels run this same
header one thread per card (tests/test_gpu_parity.py::test_best_expiry_seg)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, expiry_card

GROUP_DT = np.dtype([("h", "<i4", 7), ("rt", "<i4", 5), ("rl", "<i4", 5)])


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("xh") / "libxh.so")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-ffp-contract=off", "-shared",
                           "-I" + os.path.join(ROOT, "card.io-dmz_b200", "csrc"), os.path.join(ROOT, "tests", "expiry_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.xh_best_expiry_seg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.xh_sort_check.argtypes = [C.c_void_p, C.c_int]
    lib.xh_scharr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def slash_w():
    return np.fromfile(os.path.join(ROOT, "card.io-dmz_b200", "weights", "modelm_730c4cbd.bin"), "<f4")


def flat_groups(out, k):
    rows = []
    for i in range(k):
        row = out["h"][i].tolist()
        for j in range(5):
            row += [int(out["rt"][i][j]), int(out["rl"][i][j])]
        rows.append(row)
    return np.array(rows, np.int32).reshape(k, 17)


def host_seg(host, slash_w, card, yo):
    out = np.zeros(64, GROUP_DT)
    ov = C.c_int(0)
    k = host.xh_best_expiry_seg(card.ctypes.data, int(yo), slash_w.ctypes.data, out.ctypes.data, 64, C.byref(ov))
    assert ov.value == 0
    return flat_groups(out, k)


def test_std_sort_restatement(host):
    """libstdc++'s introsort order for equal keys (what the reference's std::sort calls produce)."""
    rng = np.random.default_rng(0)
    for t in range(1500):
        n = int(rng.integers(0, 600))
        keys = rng.integers(0, int(rng.choice([2, 5, 50, 10 ** 6])), n).astype(np.int64)
        if t % 7 == 0:
            keys = np.sort(keys)
        if t % 11 == 0:
            keys = np.sort(keys)[::-1].copy()
        assert host.xh_sort_check(keys.ctypes.data, n) == 0, t
    for n in (17, 33, 100, 420, 1000, 5000):
        pipe = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.int64)
        assert host.xh_sort_check(pipe.ctypes.data, len(pipe)) == 0
        flat = np.zeros(n, np.int64)
        assert host.xh_sort_check(flat.ctypes.data, n) == 0


def test_against_reference_golden(host, slash_w, golden):
    base, yo0 = golden["deck_card0"], int(golden["deck_records"]["v_y_offset"][0])
    pos = 0
    for sd, cnt in zip(golden["expiry_seg_seeds"], golden["expiry_seg_counts"]):
        card, yo = expiry_card(base, yo0, int(sd))
        got = host_seg(host, slash_w, card, yo)
        want = golden["expiry_seg_groups"][pos:pos + cnt]
        pos += cnt
        assert got.shape == want.shape and np.array_equal(got, want), int(sd)
    assert (golden["expiry_seg_counts"] > 0).sum() >= 40  # the fixture does exercise the whole chain
    card, yo = expiry_card(base, yo0, 1001)
    sch = np.zeros((270, 428), np.int16)
    host.xh_scharr(card.ctypes.data, yo, sch.ctypes.data)
    part = sch[yo + 27:].astype(np.uint64).ravel()
    assert np.uint64((part * np.arange(1, part.size + 1, dtype=np.uint64)).sum()) == golden["expiry_scharr_check"]
    assert not sch[:yo + 27].any()


def test_against_reference_build(host, slash_w, golden, refx):
    base, yo0 = golden["deck_card0"], int(golden["deck_records"]["v_y_offset"][0])
    found = 0
    for sd in range(5000, 5400):
        card, yo = expiry_card(base, yo0, sd)
        want = refx.best_expiry_seg(card, yo)
        got = host_seg(host, slash_w, card, yo)
        assert got.shape == want.shape and np.array_equal(got, want), sd
        found += len(want) > 0
    assert found > 100


def _stamp(card, txt, x, y, pitch=12, fg=60):
    from util import expiry_glyph
    c = card.copy()
    for i, ch in enumerate(txt):
        xs = x + i * pitch
        reg = c[y:y + 15, xs:xs + 9]
        reg[expiry_glyph(ch) > 0] = fg
    return np.ascontiguousarray(c)


def test_session_expiry_logic_against_reference_build(refx, oracle, pkg):
    """expiry_extract's session half (aggregation across frames, stability, date rules) in scanner.cpp -- host logic, no
    GPU -- fed with the reference's own per-frame groups and digit probabilities, against the reference session."""
    import datetime
    from util import deck_frames
    now = datetime.datetime.now()
    recs, cards = oracle.process_frames(deck_frames(16, 8), want_cards=True)
    for txt, fg, drift in (("08/27", 40, 0), ("11/29", 250, 0), ("03/30", 40, 3), ("05/28", 250, 7)):
        rs = refx.scanner_new()
        mine = pkg.Scanner()
        for k in range(8):
            yo = int(recs["v_y_offset"][k])
            x = 70 + (drift * k if drift == 3 else (drift if k % 2 else 0))  # slow drift / jumps beyond the 5-px tolerance
            card = _stamp(cards[k], txt, x, min(yo + 27 + 40, 250), 12, fg)
            groups = refx.best_expiry_seg(card, yo)  # what scan_card_image hands to expiry_extract
            scan, _ = refx.scanner_add_frame_with_expiry(rs, card, True)
            if scan.usable and len(groups):
                g = np.zeros(len(groups), pkg.EXPIRY_GROUP_DTYPE)
                sc = np.zeros((len(groups), 4, 10), np.float32)
                for i, row in enumerate(groups):
                    for f, name in enumerate(("top", "left", "width", "height", "character_width", "pattern", "n_rects")):
                        g[i][name] = row[f]
                    rects = row[7:].reshape(5, 2)
                    g[i]["rect_top"], g[i]["rect_left"] = rects[:, 0], rects[:, 1]
                    for r, ci in enumerate((0, 1, 3, 4)):
                        t, l = rects[ci]
                        sc[i, r] = refx.expiry_digit_model(refx.expiry_patch_prep(card[t:t + 16, l:l + 11]))
                mine.add_expiry(g, sc, now.year, now.month, allow_past_dates=True)
            (m, y), meta, scores = refx.scanner_expiry_peek(rs)
            mmeta, mscores = mine.expiry_peek()
            assert mine.expiry() == (m, y), (txt, k)
            assert np.array_equal(meta, mmeta), (txt, k, meta, mmeta)
            assert np.array_equal(scores.view(np.uint32), mscores.view(np.uint32)), (txt, k)
        refx.scanner_free(rs)
        mine.close()


def test_expiry_date_rules_against_reference_build(refx, pkg):
    """get_stable_expiry_month_and_year + expiry_string_to_expiry_month_and_year on crafted score rows (CYTHON_DMZ build:
    past dates allowed), including unstable digits, YY/MM swapping, out-of-range months and the 'later date wins' rule."""
    import datetime
    now = datetime.datetime.now()
    rng = np.random.default_rng(4)
    accepted = 0
    for t in range(1500):
        digits = [int(rng.integers(0, 10)) for _ in range(5)]
        if t % 3 == 0:
            digits[0], digits[1] = divmod(int(rng.integers(1, 13)), 10)
        if t % 5 == 0:
            digits[3], digits[4] = divmod(int(rng.integers(now.year % 100 - 2, now.year % 100 + 7)) % 100, 10)
        if t % 7 == 0:  # YY/MM order
            digits[3], digits[4] = divmod(int(rng.integers(1, 13)), 10)
            digits[0], digits[1] = divmod(int(rng.integers(13, 40)), 10)
        sc = rng.random((5, 10)).astype(np.float32) * 0.05
        for i in range(5):
            sc[i, digits[i]] = 1.0 if rng.random() > 0.1 else 0.12  # some rows fall below the 0.7 stability bar
        if t % 11 == 0:
            sc[1, (digits[1] + 1) % 10] = sc[1, digits[1]]  # exact tie: first maximum wins, stability 0.5
        m0, y0 = (0, 0) if t % 4 else (int(rng.integers(1, 13)), int(now.year + rng.integers(-1, 4)))
        want = refx.expiry_month_year(sc, m0, y0)
        got = pkg.expiry_month_year_from_scores(sc, now.year, now.month, True, m0, y0)
        assert got == want, (t, digits, got, want)
        accepted += want != (m0, y0)
    assert accepted > 100
