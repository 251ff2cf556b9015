#!/usr/bin/env python3
"""profiles/<tag>_formats_ncu_summary.md from gpurun_out/<tag>_formats_ncu.ncu-rep (tools/gpu_formats.sh: one
`ncu --set full --clock-control none` capture of the seven formats legs at 2048 frames / 4096 cards per launch), and
copies of the bench lines of the same visit.   usage: python tools/formats_profile_summary.py <tag>"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", os.path.join(OUT, tag + "_formats_ncu.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %")]
px = {"ycbcr_to_rgb_kernel<16, 3": 2048 * 640 * 480, "ycbcr_to_rgb_kernel<16, 4": 2048 * 640 * 480, "rgba_to_r": 2048 * 640 * 480, "stencil3": 2048 * 640 * 480}
lines = ["# formats kernels: `ncu --set full --clock-control none`, one launch per leg (%s)" % tag, "",
         "Capture: `tools/gpu_formats.sh` (2048 frames of 640x480 per launch; the third colour launch is 4096 cards of 428x270, which the",
         "launcher runs as one flat run of 16-pixel items).  Times under ncu are cold-cache and serialised: the bench line",
         "(`profiles/%s_formats.json`) holds the CUDA-event numbers." % tag, ""]
seen = 0
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    seen += 1
    pixels = 4096 * 428 * 270 if (name.startswith("ycbcr") and seen == 3) else 2048 * 640 * 480
    lines.append("## %d. `%s`  (%d pixels)" % (seen, name, pixels))
    lines.append("")
    lines.append("| metric | value |")
    lines.append("|---|---|")
    vals = {}
    for m, label in M:
        if m in hdr:
            v, u = r[hdr.index(m)], units[hdr.index(m)]
            vals[label] = (float(v.replace(",", "")), u)
            lines.append("| %s | %s %s |" % (label, v, u))
    if "warp instructions" in vals:
        lines.append("| thread instructions per pixel | %.1f |" % (vals["warp instructions"][0] * 32 / pixels))
    if "DRAM read" in vals and "DRAM written" in vals:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        b = vals["DRAM read"][0] * scale[vals["DRAM read"][1]] + vals["DRAM written"][0] * scale[vals["DRAM written"][1]]
        lines.append("| DRAM bytes per pixel (read + written) | %.2f |" % (b / pixels))
    lines.append("")
open(os.path.join(PROF, tag + "_formats_ncu_summary.md"), "w").write("\n".join(lines))
for suffix in ("_formats.json", "_formats_ref.json"):
    src = os.path.join(OUT, tag + suffix)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, tag + suffix))
print("\n".join(lines[:40]))
