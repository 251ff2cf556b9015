#!/usr/bin/env python3
"""profiles/<tag>_sass_evidence.md: which Blackwell / Hopper-class opcodes (TMA, tcgen05, TMEM, mbarrier, packed min/max ...) each
kernel of the shipped objects contains.   python tools/sass_evidence.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
PAT = re.compile(r"\b(UTMALDG|UTMASTG|UTMAPF|UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UTCCP|SYNCS|LDGSTS|VIMNMX3?|FFMA2|REDUX|IDP|I2IP|VABSDIFF4)\b[\.\w]*")
SHOW = ("UTMALDG", "UTCIMMA", "UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS")
out = ["# %s SASS evidence (cuobjdump -sass of the objects linked into card.io-dmz_b200/libb200dmz.so; sm_100a)\n" % tag,
       "Counts of the Blackwell / Hopper-class opcodes per kernel, then the instruction lines themselves.\n"]
build = os.path.join(ROOT, "card.io-dmz_b200", "build")
for obj in sorted(f for f in os.listdir(build) if f.endswith(".cu.o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build, obj)], capture_output=True, text=True).stdout
    cur, cnt, lines = None, collections.defaultdict(collections.Counter), collections.defaultdict(list)
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        m = PAT.search(l)
        if m and cur:
            cnt[cur][m.group(0)] += 1
            if m.group(1) in SHOW and len(lines[cur]) < 24:
                lines[cur].append(l.strip()[:150])
    for fn, c in cnt.items():
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", dem).split("(")[0][:90]
        out.append("\n## %s  %s\n" % (obj[:-2], short))
        out.append("  " + ", ".join("%s x%d" % (o, n) for o, n in sorted(c.items())) + "\n")
        out.extend("    " + l + "\n" for l in lines[fn])
open(os.path.join(ROOT, "profiles", tag + "_sass_evidence.md"), "w").write("".join(out))
print("wrote profiles/%s_sass_evidence.md" % tag)
