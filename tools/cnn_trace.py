#!/usr/bin/env python3
"""Timeline of the CNN kernel's pipeline (library built with -DB200_CNN_TRACE): clock64 stamps of block 0."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import load_pkg, deck_frames
pkg = load_pkg(); d = pkg.Dmz()
fr = deck_frames(0, 512)
d.process_frames(fr)
buf = np.zeros(3 * 2048, np.int64)
d.lib.b200_cnn_trace.argtypes = [C.c_void_p, C.c_int]
d.lib.b200_cnn_trace(buf.ctypes.data, 0)   # discard the first call's trace
d.process_frames(fr)
d.lib.b200_cnn_trace(buf.ctypes.data, 0)
ev = []
for who in range(3):
    for i in range(1024):
        tag, t = int(buf[who * 2048 + 2 * i]), int(buf[who * 2048 + 2 * i + 1])
        if t: ev.append((t, tag))
ev.sort()
t0 = ev[0][0]
names = {100: "I empty0 ok", 101: "I empty1 ok", 110: "I commit0", 111: "I commit1", 120: "I hidden start", 121: "I hidden commit",
         200: "C0 full0", 201: "C0 full1", 250: "  C15 full0", 251: "  C15 full1", 210: "C0 ld0 done", 211: "C0 ld1 done", 220: "C0 model done", 230: "C0 wait hfull", 231: "C0 hfull ok"}
prev = t0
for t, tag in ev[:400]:
    print("%8d (+%5d)  %s" % (t - t0, t - prev, names.get(tag, tag)))
    prev = t
