#!/usr/bin/env python3
"""Turn one GPU-box visit's ncu output (gpurun_out/launches.csv + gpurun_out/prof.ncu-rep, written by
tools/gpu_round.sh) into the tracked summaries under profiles/:

    profiles/<tag>_launches.csv        the raw launch list (gpu__time_duration.sum per launch)
    profiles/<tag>_launch_shares.md    per-kernel share of the step, beside bench.py's live stage timers
    profiles/<tag>_ncu_full_summary.md selected `ncu --set full` metrics per kernel
    profiles/traffic.json              DRAM bytes per frame per stage (bench.py's roofline.traffic source)

usage: python tools/profile_summary.py <tag> [frames_per_launch_in_full_capture]
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]
STAGE_OF = {"detect_strips": "detect", "warp_kernel": "warp", "warp_rows": "warp", "vseg_rows": "vseg", "vseg_mma": "vseg", "hseg_kernel": "hseg",
            "categorize_kernel": "categorize", "categorize_mma": "categorize", "digit_prep": "categorize", "finalize_records": "finalize"}
SUMMED = ("vseg", "warp", "categorize")  # stages made of several launches per step: their traffic adds up


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def launch_shares(tag):
    src = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(src):
        src = os.path.join(OUT, "launches.csv")
    shutil.copy(src, os.path.join(PROF, tag + "_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        if k.startswith("deck_") or "deck" in k:
            continue  # synthetic-input generator, outside the timed region
        t = per.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r["Metric Value"]) / 1e3
    total = sum(v[1] for v in per.values())
    cmd = "python bench.py --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-materialised"
    md = ["# ncu launch list (gpu__time_duration.sum --clock-control none), `%s`" % cmd, "",
          "| kernel | launches | total us | share of step |", "|---|---|---|---|"]
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        md.append("| %s | %d | %.1f | %.3f |" % (k, n, us, us / total))
    bench = os.path.join(OUT, tag + "_bench.json")
    if not os.path.exists(bench):
        bench = os.path.join(OUT, "bench.json")
    if os.path.exists(bench):
        try:
            b = json.loads(open(bench).read().strip().splitlines()[-1])
            st = {k: v["ms_per_100k_frames"] for k, v in (b.get("stages") or {}).items()}
            if st:
                tot = sum(st.values())
                md += ["", "Live CUDA-event stage shares of the same build (bench.py, %s frames/step, value %.0f %s):" %
                       (b["config"].get("frames_per_gpu_per_step"), b["value"], b["unit"]), ""]
                md += ["- %s: %.2f ms / 100k frames (share %.3f)" % (k, v, v / tot) for k, v in st.items()]
        except Exception as e:  # the summary is still useful without it
            md.append("(bench.json not parsed: %s)" % e)
    open(os.path.join(PROF, tag + "_launch_shares.md"), "w").write("\n".join(md) + "\n")


def full_summary(tag, frames):
    rep = os.path.join(OUT, tag + "_prof.ncu-rep")
    if not os.path.exists(rep):
        rep = os.path.join(OUT, "prof.ncu-rep")
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    md = ["# ncu --set full capture (%s), %d frames per launch, B200, --clock-control none" % (tag, frames), "",
          "Command: see tools/gpu_final.sh.  Per-launch values (cold cache, serialised: compare shares, not absolutes).", ""]
    traffic = {}
    insts = {}
    seen = set()
    for r in rows[2:]:
        d = dict(zip(head, r))
        name = short(d["Kernel Name"])
        md.append("## %s (grid %s, block %s)" % (name, d.get("Grid Size", "?"), d.get("Block Size", "?")))
        for m in METRICS:
            if m in d and d[m] != "":
                md.append("- %s = %s %s" % (m, d[m], units[head.index(m)]))
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * scale[units[head.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * scale[units[head.index("dram__bytes_write.sum")]]
            per_frame = (rd + wr) / frames
            md.append("- derived: DRAM traffic per frame = %.1f KB" % (per_frame / 1e3))
            ipf = float(d.get("smsp__inst_executed.sum", 0) or 0) / frames
            for key, stage in STAGE_OF.items():
                if key in name:
                    if stage in SUMMED or stage not in seen:
                        insts[stage] = insts.get(stage, 0) + int(ipf)
                    if stage in SUMMED:  # several launches per step add up
                        traffic[stage] = traffic.get(stage, 0) + int(per_frame)
                    elif stage not in seen:
                        traffic[stage] = int(per_frame)
                    seen.add(stage)
        except (KeyError, ValueError):
            pass
        md.append("")
    open(os.path.join(PROF, tag + "_ncu_full_summary.md"), "w").write("\n".join(md))
    json.dump({"source": "ncu --set full --clock-control none, %d frames per launch (profiles/%s_ncu_full_summary.md)" % (frames, tag),
               "unit": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per frame", "per_frame": traffic,
               "warp_instructions_per_frame": insts},
              open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    return traffic


if __name__ == "__main__":
    tag = sys.argv[1]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    launch_shares(tag)
    print(full_summary(tag, frames))
