#!/usr/bin/env python3
"""First-contact GPU shake-out: run every stage through the C ABI on a small deck and print how each compares
with the CPU oracle.  Not a test (tests/ has the real parity suite); a debugging aid for gpurun sessions."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import load_pkg, deck_frames
from oracle.binding import Oracle

pkg = load_pkg()
O = Oracle("port")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames = deck_frames(0, n)
t = time.time(); orec, ocards = O.process_frames(frames, want_cards=True); print("oracle %.1f ms/frame" % ((time.time() - t) / n * 1e3))
d = pkg.Dmz()

# ---- detect
edges, corners, found, lines = d.detect_edges(frames, want_lines=True)
boxes = O.detection_boxes(640, 480)
bad = 0
for k in range(n):
    for s, (x, y, w, h) in enumerate(boxes):
        ol = O.best_line(frames[k][y:y + h, x:x + w], s >= 2)
        gl = lines[k, s]
        for f in ("found", "r", "n", "max_votes", "low", "high", "n_edge_px"):
            if int(gl[f]) != getattr(ol, f):
                bad += 1
                if bad < 12: print("line mismatch frame", k, "strip", s, f, int(gl[f]), getattr(ol, f))
print("detect line tap mismatches:", bad)
print("all_found gpu/oracle:", found.sum(), orec["all_found"].sum())
slot = np.array([0, 1, 2, 3])
print("edge found eq:", np.array_equal(edges["found"], orec["found"]),
      " rho bits eq:", np.array_equal(edges["rho"].view(np.uint32)[orec["found"] == 1], orec["rho"].view(np.uint32)[orec["found"] == 1]),
      " theta bits eq:", np.array_equal(edges["theta"].view(np.uint32)[orec["found"] == 1], orec["theta"].view(np.uint32)[orec["found"] == 1]))
print("corner bits eq:", np.array_equal(corners.view(np.uint32), orec["corners"].view(np.uint32)), "max abs diff", np.abs(corners - orec["corners"]).max())

# ---- homography
rng = np.random.default_rng(0)
src = (np.array([106, 105, 533, 105, 106, 374, 533, 374], np.float32) + rng.uniform(-20, 20, (2000, 8))).astype(np.float32)
dst = np.tile(np.array([0, 0, 427, 0, 0, 269, 427, 269], np.float32), (2000, 1))
Mg = d.calc_persp_transform(src, dst)
Mo = np.stack([O.calc_persp_transform(src[i], dst[i]) for i in range(2000)])
print("homography bit mismatches:", int((Mg.view(np.uint32) != Mo.view(np.uint32)).any(axis=(1, 2)).sum()), "/ 2000")

# ---- warp
cards = d.transform_card(frames, orec["corners"], valid=orec["all_found"].astype(np.uint8))
print("card pixel mismatches:", int((cards != ocards).sum()), "of", cards.size)

# ---- scan on oracle cards
scans = d.scan_cards(ocards, valid=orec["all_found"].astype(np.uint8))
for f in ("v_y_offset", "v_pattern_type", "usable", "upside_down", "h_n_offsets", "h_pattern_offset"):
    print("scan", f, "mismatches:", int((scans[f] != orec[f]).sum()))
print("scan v_score max diff", np.abs(scans["v_score"] - orec["v_score"]).max())
ok = (orec["v_score"] > 15) & (orec["upside_down"] == 0)
print("hseg offsets mismatches (gated):", int((scans["h_offsets"][ok] != orec["h_offsets"][ok]).any(axis=1).sum()), "of", int(ok.sum()))
print("hseg score bits eq:", np.array_equal(scans["h_score"][ok].view(np.uint32), orec["h_score"][ok].view(np.uint32)),
      "width bits eq:", np.array_equal(scans["h_number_width"][ok].view(np.uint32), orec["h_number_width"][ok].view(np.uint32)))
print("scores max abs diff:", np.abs(scans["scores"][ok] - orec["scores"][ok]).max() if ok.any() else None)

# ---- stage taps: KATs
g = os.path.join(ROOT, "tests", "golden")
k = np.fromfile(os.path.join(g, "kat_modelm_befe75da.bin"), "<f4")
print("mlp KAT diff", np.abs(d.vseg_model(k[:204])[0] - k[204:207]).max())

# ---- whole path
t = time.time(); recs = d.process_frames(frames); print("gpu process_frames (host buffers) %.3f ms/frame" % ((time.time() - t) / n * 1e3))
for f in ("all_found", "v_y_offset", "v_pattern_type", "usable", "upside_down", "card_check"):
    print("record", f, "mismatches:", int((recs[f] != orec[f]).sum()))
okk = (orec["usable"] == 1)
print("record scores max abs diff:", np.abs(recs["scores"][ok] - orec["scores"][ok]).max() if ok.any() else None)
print("launches", d.launches)
