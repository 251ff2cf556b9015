#!/usr/bin/env python3
"""bench.py -- frames/s of the card.io-dmz detect -> warp -> OCR hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

One "step" = one pass of the whole hot path (b200_process_frames_batch: Sobel-7 / adaptive Canny / Hough in the
four detection strips -> corners + homography -> fixed-point warp to 428x270 -> vseg -> hseg -> 3-CNN digit
ensemble) over one batch of synthetic 640x480 Y frames (BASELINE.json configs[1]: 100k frames per GPU, generated
on the device by tools/deck).  `value` is timed with the frames already resident in HBM; `e2e` goes through the
same C-ABI call with HOST (pinned) buffers, H2D/D2H inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "frames/sec end-to-end detect+warp+OCR on 640x480 Y"
W, H = 640, 480
FRAME_BYTES = W * H
RECORD_BYTES = 808
DECK_SEED = 0xCA2D10
JITTER = 8.0

# algorithmic bytes per frame (DESIGN.md "Roofline arithmetic")
STRIP_BYTES = 2 * 389 * 28 + 2 * 38 * 241  # 40 100 px in the four detection strips
ALG_BYTES = {
    "detect": STRIP_BYTES + 48,
    "geometry": 4 * 36 + 264,
    "warp": 465 * 297 + 428 * 270,                # source bounding rectangle read + card written
    "vseg": 111 * 408 + 111 * 8,                  # ~68 coarse + ~43 fine rows of 408 px, 2 floats out per row
    "hseg": 428 * 27 + 48,
    "categorize": 16 * (513 + 40),
    "finalize": 428 * 270 + RECORD_BYTES,
    "pipeline_fused": 138905,                      # unique source bytes + record (never materialising the card)
    "pipeline_materialised": 370025,               # + card written once and read once (what this build does)
}


def usable_cpus():
    """Host threads this process may really use: affinity mask, clipped by a cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / period + 0.5)))
        except Exception:
            pass
    return n


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: move this rank onto the CPUs next to its GPU before any pinned host memory is allocated, so the
    e2e leg's host buffers sit on the GPU's own NUMA node (with all ranks on one socket half of the H2D traffic crosses
    the inter-socket link).  Best effort: a cpuset that does not include those CPUs leaves the affinity as it was."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus_before": before, "cpus_after": len(os.sched_getaffinity(0))}
    except Exception as e:  # noqa: BLE001 -- reported in the JSON line, never fatal
        return {"unchanged": str(e)[:120]}


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle.binding import Oracle, available
    from util import deck_frames
    kind = "ref" if available("ref") else "port"
    orc = Oracle(kind)
    cores = usable_cpus()
    sample = args.cpu_sample
    frames = deck_frames(0, sample, W, H, JITTER, DECK_SEED, threads=min(cores, 64))
    for _ in range(args.warmup):
        orc.bench_frames(frames[: max(cores, 64)], cores)
    t = 0.0
    for _ in range(args.steps):
        secs, _ = orc.bench_frames(frames, cores)
        t += secs
    fps = sample * args.steps / t
    desc = "%d deck frames per step on %d threads (%s)" % (sample, cores, "oracle/_ref: reference sources + cvshim" if kind == "ref" else "oracle port")
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32 + f32", "data": "synthetic",
        "config": {"workload": "100k synthetic 640x480 frames, full detect->warp->OCR pipeline (bounded CPU sample)",
                   "frames_per_step": sample, "width": W, "height": H},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": desc},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=100000, help="frames per GPU per step (BASELINE configs[1]: 100k)")
    ap.add_argument("--e2e-frames", type=int, default=16384, help="frames per step on the host-buffer (e2e) path")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="frames in the bounded CPU-baseline sample")
    ap.add_argument("--card-mode", default="lazy", choices=["lazy", "full"],
                    help="lazy (library default): no cards_out -> only the card rows the scan reads are warped; full: every card materialised")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    from util import load_pkg, deck_frames_cuda
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = load_pkg()
    dmz = pkg.Dmz(device=local_rank, materialise_cards=args.card_mode == "full")
    F = args.frames

    # ---- synthetic deck of this rank, generated on the device (frames [rank*F, (rank+1)*F))
    frames = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
    gen_chunk = 8192
    for f0 in range(0, F, gen_chunk):
        cnt = min(gen_chunk, F - f0)
        frames[f0:f0 + cnt] = deck_frames_cuda(rank * F + f0, cnt, W, H, JITTER, DECK_SEED)
    records = torch.zeros((F, RECORD_BYTES), dtype=torch.uint8, device="cuda")
    dmz.reserve(F, W, H)
    ext = torch.cuda.ExternalStream(dmz.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def digit_strings():
        """16 digits + n + flags per frame (32 B) from the device records; gathered to rank 0 over NCCL."""
        rec = records
        scores = rec[:, 84:84 + 640].contiguous().view(torch.float32).view(F, 16, 10)
        digits = scores.argmax(dim=2).to(torch.uint8)
        out = torch.zeros((F, 32), dtype=torch.uint8, device="cuda")
        out[:, :16] = digits
        out[:, 16] = rec[:, 84 + 640]          # hseg.n_offsets
        out[:, 17] = rec[:, 84 + 640 + 48 + 28]  # usable
        return out

    def step():
        dmz.process_frames_device(frames.data_ptr(), F, W, H, records.data_ptr())
        if dist is not None:
            ds = digit_strings()
            gathered = [torch.empty_like(ds) for _ in range(world)] if rank == 0 else None
            dist.gather(ds, gathered, dst=0)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dmz.set_profiling(True)
    launches0 = dmz.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = dmz.launches - launches0
    stage_ms, stage_frames = dmz.stage_times()
    dmz.set_profiling(False)
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    value = F * world * args.steps / (t_ms * 1e-3)

    # ---- e2e: host (pinned) buffers through the same C-ABI call
    e2e = None
    if not args.no_e2e:
        Fe = min(args.e2e_frames, F)
        h_frames = torch.empty((Fe, H, W), dtype=torch.uint8).pin_memory()
        h_frames.copy_(frames[:Fe])
        h_records = torch.zeros((Fe, RECORD_BYTES), dtype=torch.uint8).pin_memory()
        for _ in range(max(1, min(args.warmup, 2))):
            dmz.process_frames_host_ptr(h_frames.data_ptr(), Fe, W, H, h_records.data_ptr())
        barrier()
        h2d0, d2h0 = dmz.transfer_bytes()
        te = time.perf_counter()
        for _ in range(args.steps):
            dmz.process_frames_host_ptr(h_frames.data_ptr(), Fe, W, H, h_records.data_ptr())
            if dist is not None:
                pass  # records are already on the host of each rank; digit strings are a slice of them
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - te
        t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e_s = float(t_e.item())
        same = bool((h_records.cuda() == records[:Fe]).all().item())
        h2d1, d2h1 = dmz.transfer_bytes()
        e2e = {"value": Fe * world * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": (h2d1 - h2d0) // args.steps,
               "d2h_bytes_per_step": (d2h1 - d2h0) // args.steps, "frames_per_step": Fe, "records_equal_device_path": same,
               "host_frame_bytes_per_step": Fe * FRAME_BYTES, "full_frame_redos": dmz.full_frame_redos,
               "note": "host frames are full 640x480 planes in pinned memory; the library uploads only the detection-region rectangle (+2 px) of each and re-uploads a whole frame if its card quad reaches outside it",
               "timing": "wall clock around the synchronous C-ABI calls, max over ranks", "rank0_cpu_binding": numa}

    # ---- BASELINE configs[0]: ONE frame through the whole path (latency of a batch of one through the C ABI, host
    # buffers, copies and the final synchronisation inside) -- outside the throughput timing, rank 0 only
    single = None
    if rank == 0 and not args.no_e2e:
        h1 = torch.empty((64, H, W), dtype=torch.uint8).pin_memory()
        h1.copy_(frames[:64])
        r1 = torch.zeros((1, RECORD_BYTES), dtype=torch.uint8).pin_memory()
        lat = []
        for i in range(264):
            t1 = time.perf_counter()
            dmz.process_frames_host_ptr(h1[i % 64].data_ptr(), 1, W, H, r1.data_ptr())
            lat.append(time.perf_counter() - t1)
        lat = np.sort(np.array(lat[64:])) * 1e6
        single = {"gpu_call_us_median": float(np.median(lat)), "gpu_call_us_p90": float(lat[int(0.9 * len(lat))]), "calls": int(len(lat)),
                  "what": "b200_process_frames_batch(n=1, host buffers): H2D, 13 kernels, D2H, sync"}

    # ---- CPU baseline on the host cores (rank 0 only, bounded sample)
    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        from oracle.binding import Oracle, available
        kind = "ref" if available("ref") else "port"
        orc = Oracle(kind)
        cores = usable_cpus()
        S = min(args.cpu_sample, F)
        sample = frames[:S].cpu().numpy()
        orc.bench_frames(sample[: max(cores, 64)], cores)
        secs, orecs = orc.bench_frames(sample, cores)
        grecs = records[:S].cpu().numpy().view(pkg.RECORD_DTYPE).reshape(S)
        agree = {k: int((grecs[k] != orecs[k]).sum()) for k in ("all_found", "card_check", "v_y_offset", "usable", "h_pattern_offset")}
        ok = (orecs["usable"] == 1)
        agree["scores_max_abs_diff"] = float(np.abs(grecs["scores"][ok] - orecs["scores"][ok]).max()) if ok.any() else 0.0
        agree["digit_string_mismatches"] = int((grecs["scores"][ok].reshape(-1, 16, 10).argmax(2) != orecs["scores"][ok].reshape(-1, 16, 10).argmax(2)).any(1).sum())
        if single is not None:
            s1, _ = orc.bench_frames(sample[:256], 1)
            single["cpu_reference_us_per_frame_1_thread"] = 1e6 * s1 / min(256, S)
        cpu = {"value": S / secs, "unit": "frames/s", "cores": cores, "kind": "reference" if kind == "ref" else "port",
               "sample": "%d frames of the same deck, %d threads, %.2f s wall (%s)" % (S, cores, secs, "oracle/_ref" if kind == "ref" else "oracle port"),
               "parity_vs_gpu_on_sample": agree}

    if rank == 0:
        peak, peak_src = measured_peaks()
        per_stage = {}
        for k, ms in stage_ms.items():
            if ms > 0 and stage_frames:
                per_stage[k] = {"ms_per_100k_frames": ms / stage_frames * 1e5, "alg_GBps": ALG_BYTES[k] * stage_frames / (ms * 1e-3) / 1e9}
        dom = max(stage_ms, key=lambda k: stage_ms[k]) if stage_frames else None
        roof = None
        if dom:
            ach = ALG_BYTES[dom] * stage_frames / (stage_ms[dom] * 1e-3) / 1e9
            pipe = ALG_BYTES["pipeline_materialised"] * stage_frames / (sum(stage_ms.values()) * 1e-3) / 1e9
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = {"bytes_per_launch": tj["per_frame"][dom] * F, "bytes_per_frame": tj["per_frame"][dom], "source": tj["source"]}
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame": ALG_BYTES[dom],
                    "share_of_step": stage_ms[dom] / sum(stage_ms.values()),
                    "pipeline": {"alg_bytes_per_frame": ALG_BYTES["pipeline_materialised"], "achieved": pipe, "frac": pipe / peak},
                    "note": "per-stage CUDA-event times on the launching stream; stages other than detect/warp are FP32/latency bound (DESIGN.md)"}
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32 (detect, warp, hseg) + f32 (vseg, digit CNNs)", "data": "synthetic",
            "config": {"workload": "100k synthetic 640x480 frames, full detect->warp->OCR pipeline on 1xB200 (per GPU)",
                       "frames_per_gpu_per_step": F, "width": W, "height": H, "deck_seed": DECK_SEED,
                       "l2": "inputs (%.1f GB per step) are larger than L2; no flush needed" % (F * FRAME_BYTES / 1e9),
                       "parallelism": "frames sharded across %d GPU(s), NCCL gather of 32-byte digit strings to rank 0%s" % (world, "" if world > 1 else " (n/a at 1 GPU)")},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "stages": per_stage,
            "cpu_baseline": cpu, "single_frame": single, "wall_s_timed_region": wall,
        }
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
