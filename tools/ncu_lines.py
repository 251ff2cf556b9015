#!/usr/bin/env python3
"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

ncu's CLI source page has no line column, so the instruction order of `ncu --page source --csv` is zipped with
`nvdisasm -g` (which interleaves '//## File ..., line N' markers) of the same cubin.

    tools/ncu_lines.py prof.ncu-rep detect_strips detect [top]
"""
import csv, re, subprocess, sys, collections, os, tempfile

rep, kernel_re, cubin_stem = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "card.io-dmz_b200", "libb200dmz.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_stem)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# split per function
funcs, cur, line = {}, None, None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1); funcs[cur] = []; line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        funcs[cur].append(line)
name = [f for f in funcs if re.search(os.environ.get("NCU_LINES_FUNC", kernel_re), f)][0]
lines = funcs[name]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_re], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ii, si = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
inst = [r for r in rows[h + 1:] if len(r) > si][: len(lines)]
# first kernel instance only (captures may hold several launches)
agg = collections.defaultdict(lambda: [0, 0])
for r, ln in zip(inst, lines):
    try:
        agg[ln][0] += int(r[ii]); agg[ln][1] += int(r[si])
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
src = {}
print("%s: %d SASS instructions, %d executed warp-instructions, %d stall samples" % (name[:60], len(lines), ti, ts))
for ln, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    if ln:
        path = os.path.join(os.path.dirname(lib), "csrc", ln[0])
        if os.path.exists(path):
            src.setdefault(path, open(path).read().splitlines())
            text = src[path][ln[1] - 1].strip()[:100] if ln[1] - 1 < len(src[path]) else ""
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * a / ti, 100 * b / ts, ln[0] if ln else "?", ln[1] if ln else "?", text))
