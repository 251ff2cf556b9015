#!/bin/bash
# compute-sanitizer (memcheck + racecheck + initcheck) over one small pass of every formats kernel variant
mkdir -p gpurun_out
: > gpurun_out/sanitizer_formats.log
cat > /tmp/san_fmt.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from util import load_pkg
pkg = load_pkg(); d = pkg.Dmz()
rng = np.random.default_rng(0)
for (n, h, w) in [(2, 270, 428), (1, 96, 160), (2, 31, 45), (1, 7, 1)]:           # flat 16-pixel path + byte tail
    y, cb, cr = (rng.integers(0, 256, (n, h, w), dtype=np.uint8) for _ in range(3))
    for ch in (3, 4):
        d.ycbcr_to_rgb(y, cb, cr, channels=ch)
    for kind in range(3):
        d.stencil3(y, kind)                                                        # word staging, byte staging, interior + border tiles
d.stencil3(rng.integers(0, 256, (1, 200, 300), dtype=np.uint8), 0)
for npx in (5, 512, 5000):
    d.rgba_to_r(rng.integers(0, 256, 4 * npx, dtype=np.uint8))
for (w, h, rs, off, doff) in [(64, 9, 80, 0, 0), (60, 7, 64, 4, 0), (60, 7, 64, 0, 4), (61, 5, 70, 0, 0)]:  # row-wise 16 / 4 / 4-staged / byte kernels
    fs = rs * h
    dev = [torch.from_numpy(rng.integers(0, 256, off + 2 * fs, dtype=np.uint8)).cuda() for _ in range(3)]
    for ch in (3, 4):
        out = torch.zeros(2 * h * w * ch + 16, dtype=torch.uint8, device="cuda")
        d._check(d.lib.b200_ycbcr_to_rgb_batch(d.ctx, dev[0].data_ptr() + off, rs, fs, dev[1].data_ptr() + off, dev[2].data_ptr() + off, rs, fs, w, h, 2, ch,
                                               pkg.MEM_DEVICE, out.data_ptr() + doff))
    o16 = torch.zeros(2 * h * w, dtype=torch.int16, device="cuda")
    for kind in range(3):
        d._check(d.lib.b200_stencil3_batch(d.ctx, dev[0].data_ptr() + off, rs, fs, w, h, 2, kind, pkg.MEM_DEVICE, o16.data_ptr()))
src = torch.from_numpy(rng.integers(0, 256, 4 * 600 + 3, dtype=np.uint8)).cuda(); dst = torch.zeros(608, dtype=torch.uint8, device="cuda")
d._check(d.lib.b200_rgba_to_r_batch(d.ctx, src.data_ptr() + 1, 600, pkg.MEM_DEVICE, dst.data_ptr() + 1))
torch.cuda.synchronize()
print("ok formats")
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> gpurun_out/sanitizer_formats.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_fmt.py >> gpurun_out/sanitizer_formats.log 2>&1
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== |ok " gpurun_out/sanitizer_formats.log
