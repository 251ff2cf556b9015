#!/bin/bash
# compute-sanitizer (memcheck + racecheck + initcheck) over one small pass of every kernel
mkdir -p gpurun_out
: > gpurun_out/sanitizer.log
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from util import load_pkg, deck_frames
pkg = load_pkg(); d = pkg.Dmz()
fr = deck_frames(0, 6)
r, c = d.process_frames(fr, want_cards=True)
d.detect_edges(fr[:2], np.ascontiguousarray(fr[:2, ::2, ::2]), np.full((2, 240, 320), 128, np.uint8))
d.scan_cards(c[:3]); d.categorize_patches(c[0][150:177, 30:49][None]); d.vseg_model(np.zeros((3, 204), np.float32)); d.vseg_rows(c[:2])
d.process_frames(fr)  # lazy card rows (no cards asked for)
d.set_crop_margin(-1); d.process_frames(fr[:3])
big = deck_frames(0, 1, 1920, 1080); d.process_frames(big)
rng = np.random.default_rng(0)
d.expiry_digits(rng.integers(0, 256, (11, 16, 11)).astype(np.uint8))
d.frame_scores(fr[:2])
d.best_expiry_seg(c[:2], r["v_y_offset"][:2].astype(np.uint16))
print("ok", r["all_found"].sum())
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> gpurun_out/sanitizer.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py >> gpurun_out/sanitizer.log 2>&1
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== |ok " gpurun_out/sanitizer.log
