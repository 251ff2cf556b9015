#!/usr/bin/env python3
"""Device-resident rates of the stage entry points that bench.py's headline does not time (SURVEY 8a row E0 and the
8f rows): expiry digits, n_categorize alone (BASELINE configs[2]), expiry segmentation, frame scores, de-interleave.
Prints one JSON line; run on a GPU box:  python tools/gpu_side_bench.py [n]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import load_pkg, deck_frames_cuda

pkg = load_pkg()
dmz = pkg.Dmz(device=0)
lib, ctx = dmz.lib, dmz.ctx
MEM_DEVICE = 1 if not hasattr(pkg, "MEM_DEVICE") else pkg.MEM_DEVICE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
g = torch.Generator(device="cuda").manual_seed(1)
out = {}


def timed(label, fn, units, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    out[label] = {"per_s": units / dt, "ms": dt * 1e3, "units": units}


# E0: expiry digit crops 16 x 11
patches = torch.randint(0, 256, (n, 176), dtype=torch.uint8, device="cuda", generator=g)
probs = torch.zeros((n, 10), dtype=torch.float32, device="cuda")
timed("expiry_digits", lambda: dmz._check(lib.b200_expiry_digits_batch(ctx, C.c_void_p(patches.data_ptr()), n, MEM_DEVICE, C.c_void_p(probs.data_ptr()))), n)
# n_categorize alone (configs[2]): 27 x 19 patches
p2 = torch.randint(0, 256, (n, 513), dtype=torch.uint8, device="cuda", generator=g)
o2 = torch.zeros((n, 40), dtype=torch.float32, device="cuda")
timed("categorize_patches", lambda: dmz._check(lib.b200_categorize_patches_batch(ctx, C.c_void_p(p2.data_ptr()), n, MEM_DEVICE, C.c_void_p(o2.data_ptr()))), n)
# frame scores on deck frames
F = int(os.environ.get("SIDE_BENCH_CARDS", "8192"))
frames = torch.cat([deck_frames_cuda(f0, min(8192, F - f0), 640, 480, 8.0, 0xCA2D10) for f0 in range(0, F, 8192)])
fo = torch.zeros(F, dtype=torch.float32, device="cuda"); br = torch.zeros(F, dtype=torch.float32, device="cuda")
timed("frame_scores", lambda: dmz._check(lib.b200_frame_scores_batch(ctx, C.c_void_p(frames.data_ptr()), 640, 640 * 480, 640, 480, F, 0, MEM_DEVICE,
                                                                        C.c_void_p(fo.data_ptr()), C.c_void_p(br.data_ptr()))), F)
# best_expiry_seg on warped cards of the deck (8f rank 4): Scharr + row sums + the one-thread-per-card group search
recs = torch.zeros((F, 808), dtype=torch.uint8, device="cuda")
cards = torch.zeros((F, 270, 428), dtype=torch.uint8, device="cuda")
dmz.process_frames_device(frames.data_ptr(), F, 640, 480, recs.data_ptr(), cards.data_ptr())
r = recs.cpu().numpy().view(pkg.RECORD_DTYPE).reshape(F)
yo = torch.from_numpy(r["v_y_offset"].astype("uint16").view("int16")).cuda()
MAXG = 16
groups = torch.zeros((F, MAXG, 68), dtype=torch.uint8, device="cuda")
ng = torch.zeros(F, dtype=torch.int32, device="cuda"); nd = torch.zeros(F, dtype=torch.int32, device="cuda")
timed("best_expiry_seg", lambda: dmz._check(lib.b200_best_expiry_seg_batch(ctx, C.c_void_p(cards.data_ptr()), C.c_void_p(yo.data_ptr()), F, MEM_DEVICE,
                                                                          C.c_void_p(groups.data_ptr()), MAXG, C.c_void_p(ng.data_ptr()), C.c_void_p(nd.data_ptr()), None)), F, reps=3)
out["best_expiry_seg"]["cards_with_groups"] = int((ng > 0).sum().item())
print(json.dumps(out))
