#!/usr/bin/env python3
"""Generate tests/golden/ref_golden.npz by running the REFERENCE'S OWN SOURCES (oracle/_ref/libdmz_ref.so =
/root/reference unity build + oracle/cvshim) on seeded inputs.  Run in the build container (needs
/root/reference to have built oracle/_ref); the .npz is committed and pins the plain-C oracle (and, through it,
the CUDA path) on machines where the reference is absent."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import deck_frames, synthetic_strip
from oracle.binding import Oracle, Line

R = Oracle("ref")
assert R.lib.ref_run_kats() == 0b1111, "reference KATs failed"
rng = np.random.default_rng(20260925)
out = {}

# ---- detect stage on isolated strips (full-size strips of both orientations + tiny / degenerate ones)
strips = []
for (w, h, vert, kind) in [(389, 28, 0, "edge"), (38, 241, 1, "edge"), (389, 28, 0, "noise"), (38, 241, 1, "flat"),
                           (61, 23, 0, "edge"), (19, 77, 1, "edge"), (9, 8, 0, "edge"), (8, 9, 1, "noise")]:
    strips.append((synthetic_strip(rng, w, h, vert, kind), vert))
# two strips cut from a deck frame
f = deck_frames(3, 1)[0]
b = R.detection_boxes(640, 480)
strips.append((f[b[0][1]:b[0][1] + b[0][3], b[0][0]:b[0][0] + b[0][2]].copy(), 0))
strips.append((f[b[3][1]:b[3][1] + b[3][3], b[3][0]:b[3][0] + b[3][2]].copy(), 1))
line_fields = [n for n, _ in Line._fields_ if n != "max_votes"]
for i, (img, vert) in enumerate(strips):
    dx, dy = R.sobel7(img)
    edges, lo, hi = R.adaptive_canny(img, dx, dy)
    l = R.best_line(img, vert)
    out["strip%d_img" % i] = img
    out["strip%d_vertical" % i] = np.int32(vert)
    out["strip%d_dx" % i] = dx
    out["strip%d_dy" % i] = dy
    out["strip%d_edges" % i] = edges
    out["strip%d_line" % i] = np.array([getattr(l, n) for n in line_fields], np.float64)
    out["strip%d_rho_theta_bits" % i] = np.array([l.rho, l.theta], np.float32).view(np.uint32)
out["n_strips"] = np.int32(len(strips))
out["line_fields"] = np.array(line_fields)

# ---- detection boxes
for (w, h, o) in [(640, 480, 3), (1280, 720, 3), (1920, 1080, 4), (480, 640, 1), (480, 640, 2)]:
    out["boxes_%dx%d_o%d" % (w, h, o)] = R.detection_boxes(w, h, o)

# ---- homography (Eigen householderQr().solve in float)
src = (np.array([106, 105, 533, 105, 106, 374, 533, 374], np.float32) + rng.uniform(-25, 25, (64, 8))).astype(np.float32)
dst = np.tile(np.array([0, 0, 427, 0, 0, 269, 427, 269], np.float32), (64, 1))
out["homog_src"] = src
out["homog_M_bits"] = np.stack([R.calc_persp_transform(src[i], dst[i]) for i in range(64)]).view(np.uint32)

# ---- whole path on deck frames (regenerated from the seed by the tests)
idx = np.array([0, 1, 8, 9, 17, 42], np.uint32)
frames = np.concatenate([deck_frames(int(i), 1) for i in idx])
recs, cards = R.process_frames(frames, want_cards=True)
out["deck_idx"] = idx
out["deck_records"] = recs
out["deck_frame0"] = frames[0]
out["deck_card0"] = cards[0]
# per-row vseg probabilities and per-digit model outputs of card 0
out["card0_vseg_rows"] = np.stack([R.vseg_row(cards[0], r) for r in range(0, 270, 9)])
vs = R.best_n_vseg(cards[0]); hs = R.best_n_hseg(cards[0], vs)
patches = np.stack([cards[0][vs.y_offset:vs.y_offset + 27, hs.offsets[d]:hs.offsets[d] + 19] for d in range(hs.n_offsets)])
prep = np.stack([R.digit_patch_prep(p) for p in patches])
models = [R.digit_models(p) for p in prep]
out["card0_patches"] = patches
out["card0_patch_prep"] = prep
out["card0_ensemble"] = np.stack([m[0] for m in models])
out["card0_models"] = np.stack([m[1] for m in models])
# orientation variants of dmz_transform_card on frame 0
c = recs["corners"][0]
for o in (1, 2, 3, 4):
    out["card0_orient%d_check" % o] = np.uint32((R.transform_card(frames[0], c, o).astype(np.uint64).ravel() * np.arange(1, 428 * 270 + 1, dtype=np.uint64)).sum() & 0xFFFFFFFF)

# ---- scanner session over 8 frames of one deck session
s = R.scanner_new()
sess = deck_frames(16, 8)
srecs, scards = R.process_frames(sess, want_cards=True)
results = []
for k in range(8):
    R.scanner_add_frame(s, scards[k])
    done, digits = R.scanner_result(s)
    results.append((int(done), digits.tolist()))
a15, a16, cnt = R.scanner_peek(s)
R.scanner_free(s)
out["session_agg16"] = a16; out["session_agg15"] = a15; out["session_counts"] = cnt
out["session_complete"] = np.array([r[0] for r in results], np.int32)
out["session_digits"] = np.array(results[-1][1], np.uint8)

# ---- frame scoring: dmz_focus_score / dmz_brightness_score on the deck frames above and on noise at other sizes
out["deck_focus"] = np.array([[R.focus_score(f, full) for f in frames] for full in (0, 1)], np.float32).view(np.uint32)
out["deck_brightness"] = np.array([[R.brightness_score(f, full) for f in frames] for full in (0, 1)], np.float32).view(np.uint32)
srng = np.random.default_rng(99)
for (w, h) in [(1280, 720), (320, 240), (641, 479)]:
    img = srng.integers(0, 256, (h, w), dtype=np.uint8)
    out["score_%dx%d_img_seed" % (w, h)] = np.int32(99)
    out["score_%dx%d" % (w, h)] = np.array([[R.focus_score(img, full), R.brightness_score(img, full)] for full in (0, 1)], np.float32).view(np.uint32)
    out["score_rect_%dx%d" % (w, h)] = np.stack([R.scoring_rect(w, h, full) for full in (0, 1)])

# ---- E0 (expiry digit): prepare_image_for_cat + applyc_bf4dd6c8 of the SCAN_EXPIRY=1 build on seeded character crops
RX = Oracle("refx")
erng = np.random.default_rng(77)
ep = erng.integers(0, 256, (48, 16, 11), dtype=np.uint8)
ep[16:32] = (ep[16:32] // 32) * 9
ep[32] = 0
ep[33] = 255
eprep = np.stack([RX.expiry_patch_prep(p) for p in ep])
out["expiry_patches"] = ep
out["expiry_prep_bits"] = eprep.view(np.uint32)
out["expiry_probs"] = np.stack([RX.expiry_digit_model(p) for p in eprep])

# ---- best_expiry_seg of the SCAN_EXPIRY=1 build on synthetic expiry cards (tests/util.expiry_card)
from util import expiry_card
yo0 = int(recs["v_y_offset"][0])
seg_seeds = np.arange(1000, 1160)
seg_out, seg_counts = [], []
for sd in seg_seeds:
    c, yo = expiry_card(cards[0], yo0, int(sd))
    gr = RX.best_expiry_seg(c, yo)
    seg_counts.append(len(gr))
    seg_out.append(gr)
out["expiry_seg_seeds"] = seg_seeds
out["expiry_seg_counts"] = np.array(seg_counts, np.int32)
out["expiry_seg_groups"] = np.concatenate(seg_out).astype(np.int32) if sum(seg_counts) else np.zeros((0, 17), np.int32)
c, yo = expiry_card(cards[0], yo0, 1001)
sch = RX.scharr3_dx_abs(c[yo + 27:])
out["expiry_scharr_check"] = np.uint64((sch.astype(np.uint64).ravel() * np.arange(1, sch.size + 1, dtype=np.uint64)).sum())

path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes; session complete flags", out["session_complete"], "digits", out["session_digits"])
