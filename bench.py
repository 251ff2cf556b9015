#!/usr/bin/env python3
"""bench.py -- the card.io-dmz detect -> warp -> OCR hot path on B200 (BASELINE.json metric and configs).

    python bench.py --gpus N --steps K --warmup W                      # configs[1] / [4]: frames/s of the whole path
    python bench.py --config categorize ...                            # configs[2]: n_categorize only, 10 M patches
    python bench.py --config detect-sweep ...                          # configs[3]: Canny + Hough at 480p / 720p / 1080p
    python bench.py --impl reference [--config ...] ...                # the reference's own CPU path on the host cores

Default config: one "step" = one pass of the whole hot path (b200_process_frames_batch: Sobel-7 / adaptive Canny / Hough in
the four detection strips -> corners + homography -> fixed-point warp -> vseg -> hseg -> 3-CNN digit ensemble) over one
batch of synthetic 640x480 Y frames (100k frames per GPU, generated on the device by tools/deck).  `value` is timed with the
frames already resident in HBM; `e2e` goes through the same C-ABI call with HOST (pinned) buffers, H2D / D2H inside the
timed region.  The caller does not ask for the 428x270 cards, so the library never materialises them (only the card rows
the scan reads are warped); `materialised` reports the same step with every card warped in full.  BASELINE configs[0]
(one frame) is the `single_frame` block.  Prints ONE JSON line on rank 0.
"""
import argparse
import importlib.util
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
LAZY_ROWS = 68 + 43 - 11                    # coarse rows + fine window - the fine rows that are coarse rows too
ALG_BYTES = {
    "detect": STRIP_BYTES + 48,
    "geometry": 4 * 36 + 264,
    "warp": 465 * 297 + 428 * 270,                # source bounding rectangle read + card written (materialised card)
    "warp_lazy": 465 * 297 + 428 * (68 + 43),     # the same source rectangle + the ~111 rows that are written
    "vseg": 111 * 408 + 111 * 8,                  # ~68 coarse + ~43 fine rows of 408 px, 2 floats out per row
    "hseg": 428 * 27 + 48,
    "categorize": 16 * (513 + 40),
    "finalize": RECORD_BYTES,
    "pipeline_fused": 138905,                      # unique source bytes + record (the card never exists in HBM)
    "pipeline_materialised": 370025,               # + card written once and read once
}
# FP32 work of the two network stages: flop per scored row / per digit patch
VSEG_FLOP_PER_ROW = 2 * (50 * 204 + 3 * 50)
CNN_FLOP_PER_PATCH = 218880
DETECT_SIZES = {"480p": (640, 480), "720p": (1280, 720), "1080p": (1920, 1080)}


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


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def fp32_peak_tflops(sm_count, sm_mhz):
    """FP32 CUDA-core peak: SMs x 128 lanes x 2 flop (FMA) x clock."""
    return sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12


def load_sharding():
    spec = importlib.util.spec_from_file_location("cardio_dmz_b200_sharding", os.path.join(ROOT, "card.io-dmz_b200", "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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


def cpu_checker():
    from oracle.binding import Oracle, available
    kind = "ref" if available("ref") else "port"
    return Oracle(kind), kind, ("oracle/_ref: reference sources + cvshim" if kind == "ref" else "oracle port")


def synthetic_patches_host(n, seed=7):
    """BASELINE configs[2] input on the host (CPU legs): half i.i.d. noise, half crops of deck digits."""
    from util import deck_frames
    orc, _, _ = cpu_checker()
    recs, cards = orc.process_frames(deck_frames(0, 64, W, H, JITTER, DECK_SEED), want_cards=True)
    crops = []
    for r, c in zip(recs, cards):
        if r["usable"]:
            for d in range(int(r["h_n_offsets"])):
                y, x = int(r["v_y_offset"]), int(r["h_offsets"][d])
                crops.append(c[y:y + 27, x:x + 19])
    crops = np.stack(crops)
    out = np.random.default_rng(seed).integers(0, 256, (n, 27, 19), dtype=np.uint8)
    out[n // 2:] = crops[np.arange(n - n // 2) % len(crops)]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation, all host threads, bounded sample per step
# ---------------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    if rank != 0:
        return
    from util import deck_frames
    orc, kind, kind_desc = cpu_checker()
    cores = usable_cpus()
    extra = {}
    if args.config == "pipeline":
        sample = args.cpu_sample
        frames = deck_frames(0, sample, W, H, JITTER, DECK_SEED, threads=min(cores, 64))
        run = lambda: orc.bench_frames(frames, cores)[0]  # noqa: E731
        warm = lambda: orc.bench_frames(frames[: max(cores, 64)], cores)  # noqa: E731
        metric, unit, units = METRIC, "frames/s", sample
        workload = "100k synthetic 640x480 frames, full detect->warp->OCR pipeline (bounded CPU sample)"
        desc = "%d deck frames per step on %d threads (%s)" % (sample, cores, kind_desc)
    elif args.config == "categorize":
        sample = 1 << 19
        patches = synthetic_patches_host(sample)
        run = lambda: orc.bench_patches(patches, cores)[0]  # noqa: E731
        warm = lambda: orc.bench_patches(patches[: 64 * cores], cores)  # noqa: E731
        metric, unit, units = "digit patches/sec through n_categorize (3-CNN ensemble)", "patches/s", sample
        workload = "n_categorize only: 10M synthetic digit patches (bounded CPU sample)"
        desc = "%d patches per step on %d threads (%s)" % (sample, cores, kind_desc)
    elif args.config == "formats":
        from concurrent.futures import ThreadPoolExecutor
        sample = 32 * cores
        rng = np.random.default_rng(3)
        planes = rng.integers(0, 256, (3, sample, H, W), dtype=np.uint8)
        pool = ThreadPoolExecutor(cores)  # the reference function is single-threaded: one frame per call, ctypes drops the GIL

        def run():
            t0 = time.perf_counter()
            list(pool.map(lambda k: orc.ycbcr_to_rgb(planes[0, k], planes[1, k], planes[2, k], 3), range(sample)))
            return time.perf_counter() - t0
        warm = run
        metric, unit, units = "pixels/sec through dmz_YCbCr_to_RGB (640x480 planes -> interleaved RGB)", "pixels/s", sample * W * H
        workload = "pixel formats around the path: dmz_YCbCr_to_RGB on 640x480 planes (bounded CPU sample)"
        desc = "%d frames per step, one frame per call on %d threads (%s)" % (sample, cores, kind_desc)
    else:
        per = {"480p": 4096, "720p": 2048, "1080p": 1024}
        decks = {k: deck_frames(0, per[k], w, h, JITTER, DECK_SEED, threads=min(cores, 64)) for k, (w, h) in DETECT_SIZES.items()}
        run = lambda: sum(orc.bench_detect(decks[k], cores)[0] for k in decks)  # noqa: E731
        warm = lambda: [orc.bench_detect(decks[k][: max(cores, 64)], cores) for k in decks]  # noqa: E731
        metric, unit, units = "frames/sec Canny+Hough (best_line_for_sample x 4 strips), 480p/720p/1080p sweep", "frames/s", sum(per.values())
        workload = "Canny+Hough sweep 480p/720p/1080p x 100k frames each (bounded CPU sample)"
        desc = "%s frames per step on %d threads (%s)" % (per, cores, kind_desc)
        extra["sweep"] = {k: {"frames/s": per[k] / orc.bench_detect(decks[k], cores)[0]} for k in decks}
    for _ in range(args.warmup):
        warm()
    t = sum(run() for _ in range(args.steps))
    value = units * args.steps / t
    out = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32 + f32", "data": "synthetic",
        "config": {"workload": workload, "units_per_step": units, "width": W, "height": H},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.update(extra)
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# shared GPU set-up
# ---------------------------------------------------------------------------------------------------------------------
class Gpu:
    def __init__(self, args):
        import torch
        from util import load_pkg
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else None
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.pkg = load_pkg()
        self.dmz = self.pkg.Dmz(device=self.local_rank, materialise_cards=args.card_mode == "full")
        self.ext = torch.cuda.ExternalStream(self.dmz.stream, device=torch.device("cuda", self.local_rank))
        self.sm_count = torch.cuda.get_device_properties(self.local_rank).multi_processor_count

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup):
        """W warm-up steps, then K steps between CUDA events on the library's stream (+ the current stream, where a
        collective may trail), barrier + synchronize on both sides; returns (max-over-ranks ms, clocks, wall s)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record(self.ext)
        for _ in range(steps):
            step()
        e1.record(self.ext)
        e2.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        clocks = sampler.stop() if self.rank == 0 else None
        return self.max_over_ranks(ms), clocks, wall

    def deck(self, first, n, w=W, h=H):
        from util import deck_frames_cuda
        frames = self.torch.empty((n, h, w), dtype=self.torch.uint8, device="cuda")
        chunk = max(256, (8192 * W * H) // (w * h))
        for f0 in range(0, n, chunk):
            cnt = min(chunk, n - f0)
            frames[f0:f0 + cnt] = deck_frames_cuda(first + f0, cnt, w, h, JITTER, DECK_SEED)
        return frames

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def base_line(g, args, metric, value, unit, t_ms, dtype, workload, extra_cfg):
    cfg = {"workload": workload}
    cfg.update(extra_cfg)
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": g.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": cfg}


def host_h2d_ceiling(g, nbytes=1 << 30, reps=4):
    """What this host can deliver to this GPU: a plain contiguous pinned-memory H2D copy, every rank at the same time."""
    torch = g.torch
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    g.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = g.max_over_ranks(time.perf_counter() - t0)
    del src, dst
    return nbytes * reps / dt / 1e9


# ---------------------------------------------------------------------------------------------------------------------
# configs[1] / [4] (+ [0] as single_frame): the whole path
# ---------------------------------------------------------------------------------------------------------------------
def pipeline_config(args):
    g = Gpu(args)
    torch, dmz, dist, rank, world = g.torch, g.dmz, g.dist, g.rank, g.world
    sh = load_sharding()
    F = args.frames
    frames = g.deck(rank * F, F)
    records = torch.zeros((F, RECORD_BYTES), dtype=torch.uint8, device="cuda")
    dmz.reserve(F, W, H)

    def step():
        dmz.process_frames_device(frames.data_ptr(), F, W, H, records.data_ptr())
        if dist is not None:  # the path's only collective: 32-byte digit strings to rank 0
            sh.gather_equal_shards_torch(sh.digit_records_torch(records), dist, rank, world)

    for _ in range(args.warmup):
        step()
    dmz.set_profiling(True)  # (zeroes the per-stage accumulators)
    launches0 = dmz.launches
    t_ms, clocks, wall = g.timed(step, args.steps, 0)
    launches = dmz.launches - launches0
    stage_ms, stage_frames = dmz.stage_times()
    dmz.set_profiling(False)
    value = F * world * args.steps / (t_ms * 1e-3)

    # ---- the same step with every card materialised (what a caller passing cards_out pays), device-resident
    materialised = None
    full_records = records
    if args.card_mode == "lazy" and not args.no_materialised:
        lazy_records = records.clone()
        dmz.set_card_mode(True)
        step()
        dmz.set_profiling(True)
        m_ms, _, _ = g.timed(step, args.steps, 0)
        m_stage, m_frames = dmz.stage_times()
        dmz.set_profiling(False)
        full_records = records.clone()
        dmz.set_card_mode(False)
        b = full_records.clone()
        b[:, 804:808] = 0  # card_check: 0 on the lazy path by definition
        materialised = {"value": F * world * args.steps / (m_ms * 1e-3), "unit": "frames/s", "ms_per_step": m_ms / args.steps,
                        "stages_ms_per_100k_frames": {k: v / max(m_frames, 1) * 1e5 for k, v in m_stage.items() if v > 0},
                        "records_equal_lazy_except_card_check": bool((lazy_records == b).all().item()),
                        "card_checks_nonzero": bool((full_records[:, 804:808] != 0).any(dim=1).all().item())}
        del b
        records.copy_(lazy_records)
        del lazy_records

    # ---- e2e: host (pinned) buffers through the same C-ABI call, digit strings gathered to rank 0 at N > 1
    e2e = None
    if not args.no_e2e:
        Fe = min(args.e2e_frames, F)
        avail = mem_available_bytes()
        budget = int(0.6 * avail / max(world, 1)) if avail else Fe * FRAME_BYTES
        Fe = max(1024, min(Fe, budget // FRAME_BYTES) // 1024 * 1024) if Fe >= 1024 else Fe
        wc_ptr = None
        if os.environ.get("B200_BENCH_WC"):  # experiment: write-combined pinned upload buffer (cudaHostAllocWriteCombined)
            import ctypes as C
            rt = C.CDLL("libcudart.so")
            wc_ptr = C.c_void_p()
            assert rt.cudaHostAlloc(C.byref(wc_ptr), C.c_size_t(Fe * FRAME_BYTES), C.c_uint(4)) == 0
            assert rt.cudaMemcpy(wc_ptr, C.c_void_p(frames.data_ptr()), C.c_size_t(Fe * FRAME_BYTES), C.c_int(2)) == 0

            class _Wc:
                def data_ptr(self):
                    return wc_ptr.value
            h_frames = _Wc()
        else:
            h_frames = torch.empty((Fe, H, W), dtype=torch.uint8).pin_memory()
            h_frames.copy_(frames[:Fe])
        h_records = torch.zeros((Fe, RECORD_BYTES), dtype=torch.uint8).pin_memory()
        h_np = h_records.numpy().view(g.pkg.RECORD_DTYPE).reshape(Fe)

        def e2e_step():
            dmz.process_frames_host_ptr(h_frames.data_ptr(), Fe, W, H, h_records.data_ptr())
            if dist is not None:  # digit strings of this rank's records -> rank 0 (records go back up: 80 MB, ~1.6 ms; arg-max on the GPU)
                sh.gather_equal_shards_torch(sh.digit_records_torch(h_records.to("cuda", non_blocking=True)), dist, rank, world)

        e2e_step()
        g.barrier()
        h2d0, d2h0 = dmz.transfer_bytes()
        te = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = g.max_over_ranks(time.perf_counter() - te)
        same = bool((h_records.cuda() == records[:Fe]).all().item())
        h2d1, d2h1 = dmz.transfer_bytes()
        del h_frames
        if wc_ptr is not None:
            rt.cudaFreeHost(wc_ptr)
        ceiling = host_h2d_ceiling(g)
        h2d_per_frame = (h2d1 - h2d0) / args.steps / Fe
        e2e_value = Fe * world * args.steps / e2e_s
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": (h2d1 - h2d0) // args.steps,
               "d2h_bytes_per_step": (d2h1 - d2h0) // args.steps, "frames_per_step": Fe, "records_equal_device_path": same,
               "host_frame_bytes_per_step": Fe * FRAME_BYTES, "full_frame_redos": dmz.full_frame_redos,
               "h2d_GBps_per_gpu": e2e_value / world * h2d_per_frame / 1e9,
               "host_h2d_ceiling_GBps_per_gpu": ceiling,
               "fraction_of_host_h2d_ceiling": e2e_value / world * h2d_per_frame / 1e9 / ceiling,
               "note": "host frames are full 640x480 planes in pinned memory; the library uploads only the detection-region rectangle (+2 px) of each "
                       "(%d B per frame) and re-uploads a whole frame if its card quad reaches outside it; the ceiling is a contiguous pinned H2D copy "
                       "run by all %d rank(s) at once on this host; the same link delivers 47.95 GB/s for this rectangle of 480-of-640-byte rows whatever "
                       "moves it (pitched DMA on 1 / 2 / 4 streams, a zero-copy kernel, hybrids: profiles/r04_h2d_rectangle_microbench.txt)" % (int(h2d_per_frame), world),
               "digit_strings_gathered": dist is not None,
               "timing": "wall clock around the synchronous C-ABI calls (+ the gather), max over ranks", "rank0_cpu_binding": g.numa}

    # ---- BASELINE configs[0]: ONE frame -- the whole-path batch call with n = 1, and the SDK's three-call sequence
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
                  "what": "b200_process_frames_batch(n=1, host buffers): H2D, kernels, D2H, sync"}
        exe = os.path.join(ROOT, "card.io-dmz_b200", "build", "dropin_latency")
        if os.path.exists(exe):
            try:
                path = "/tmp/b200_dropin_frames_%d.bin" % os.getpid()
                h1.numpy().tofile(path)
                out = subprocess.run([exe, path, "64", str(W), str(H), "200"], capture_output=True, text=True, timeout=120).stdout
                os.unlink(path)
                single["dropin_sequence"] = json.loads(out.strip().splitlines()[-1])
            except Exception as e:  # noqa: BLE001
                single["dropin_sequence"] = {"error": str(e)[:200]}

    # ---- CPU baseline on the host cores (rank 0 only, bounded sample) + parity of the timed GPU records against it
    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        from oracle.binding import Oracle, available
        orc, kind, kind_desc = cpu_checker()
        cores = usable_cpus()
        S = min(args.cpu_sample, F)
        sample = frames[:S].cpu().numpy()
        orc.bench_frames(sample[: max(cores, 64)], cores)
        secs, orecs = orc.bench_frames(sample, cores)
        grecs = records[:S].cpu().numpy().view(g.pkg.RECORD_DTYPE).reshape(S)
        frecs = full_records[:S].cpu().numpy().view(g.pkg.RECORD_DTYPE).reshape(S)
        agree = {k: int((grecs[k] != orecs[k]).sum()) for k in ("all_found", "v_y_offset", "usable", "h_pattern_offset")}
        if materialised is not None or args.card_mode == "full":
            agree["card_check (materialised records)"] = int((frecs["card_check"] != orecs["card_check"]).sum())
        ok = (orecs["usable"] == 1)
        agree["scores_max_abs_diff"] = float(np.abs(grecs["scores"][ok] - orecs["scores"][ok]).max()) if ok.any() else 0.0
        agree["digit_string_mismatches"] = int((grecs["scores"][ok].reshape(-1, 16, 10).argmax(2) != orecs["scores"][ok].reshape(-1, 16, 10).argmax(2)).any(1).sum())
        stage_secs, stage_counts = orc.bench_stages(sample[:256])
        per_stage_us = {k: 1e6 * s / max(c, 1) for k, s, c in zip(("detect", "transform", "scan"), stage_secs, stage_counts)}
        if single is not None:
            s1, _ = orc.bench_frames(sample[:256], 1)
            single["cpu_reference_us_per_frame_1_thread"] = 1e6 * s1 / min(256, S)
            single["cpu_reference_us_per_stage_1_thread"] = per_stage_us
        cpu = {"value": S / secs, "unit": "frames/s", "cores": cores, "kind": "reference" if kind == "ref" else "port",
               "sample": "%d frames of the same deck, %d threads, %.2f s wall (%s, -O2 x86-64 baseline ISA)" % (S, cores, secs, kind_desc),
               "us_per_stage_1_thread": per_stage_us, "parity_vs_gpu_on_sample": agree}
        if available("refo3"):
            o3 = Oracle("refo3")
            o3.bench_frames(sample[: max(cores, 64)], cores)
            s3, _ = o3.bench_frames(sample, cores)
            cpu["o3_avx2_build"] = {"value": S / s3, "unit": "frames/s", "cores": cores,
                                    "what": "the same sources at -O3 -march=x86-64-v3 (AVX2+FMA code generation, no FP contraction); "
                                            "-march=native cannot travel from the build container to this host"}

    if rank == 0:
        peak, peak_src = measured_peaks()
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = fp32_peak_tflops(g.sm_count, sm_mhz)
        rec_np = records.cpu().numpy().view(g.pkg.RECORD_DTYPE).reshape(F)
        ran_cnn = np.abs(rec_np["scores"]).sum(axis=1) > 0
        digits_per_frame = float(rec_np["h_n_offsets"][ran_cnn].sum()) / F
        rows_per_frame = float(rec_np["all_found"].mean()) * LAZY_ROWS
        lazy = args.card_mode == "lazy"
        alg = dict(ALG_BYTES)
        if lazy:
            alg["warp"] = alg["warp_lazy"]
        per_stage = {}
        for k, ms in stage_ms.items():
            if ms > 0 and stage_frames:
                per_stage[k] = {"ms_per_100k_frames": ms / stage_frames * 1e5, "alg_GBps": alg[k] * stage_frames / (ms * 1e-3) / 1e9}
        for k, flop in (("vseg", rows_per_frame * VSEG_FLOP_PER_ROW), ("categorize", digits_per_frame * CNN_FLOP_PER_PATCH)):
            if k in per_stage:
                tf = flop * stage_frames / (stage_ms[k] * 1e-3) / 1e12
                per_stage[k].update({"bound": "fp32", "flop_per_frame": flop, "TFLOPs": tf, "fp32_peak_TFLOPs": fp32_peak, "fp32_frac": tf / fp32_peak})
        dom = max(stage_ms, key=lambda k: stage_ms[k]) if stage_frames else None
        roof = None
        if dom:
            total_ms = sum(stage_ms.values())
            ach = alg[dom] * stage_frames / (stage_ms[dom] * 1e-3) / 1e9
            fused = ALG_BYTES["pipeline_fused"] * stage_frames / (total_ms * 1e-3) / 1e9
            mat = ALG_BYTES["pipeline_materialised"] * stage_frames / (total_ms * 1e-3) / 1e9
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = {"bytes_per_launch": tj["per_frame"][dom] * F, "bytes_per_frame": tj["per_frame"][dom], "source": tj["source"]}
            except Exception:
                pass
            issue = None
            try:
                wi = tj["warp_instructions_per_frame"][dom]
                rate = wi * stage_frames / (stage_ms[dom] * 1e-3)
                issue_peak = g.sm_count * 4 * sm_mhz * 1e6  # one warp instruction per scheduler per cycle
                issue = {"warp_instructions_per_frame": wi, "achieved_per_s": rate, "peak_per_s": issue_peak, "frac": rate / issue_peak,
                         "what": "the limiter of this kernel: warp instructions (ncu smsp__inst_executed.sum of the same build, "
                                 "profiles/traffic.json) / the kernel's CUDA-event time, against SMs x 4 schedulers x clock"}
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "issue_slots": issue,
                    "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame": alg[dom],
                    "share_of_step": stage_ms[dom] / total_ms,
                    "pipeline": {"alg_bytes_per_frame": ALG_BYTES["pipeline_fused"], "achieved": fused, "frac": fused / peak,
                                 "what": ("fused denominator (SURVEY 8d): unique source bytes + record; the cards are never materialised on this path "
                                          "(%d of 270 card rows per frame are written to and re-read from HBM, see DESIGN.md)" % (68 + 43)) if lazy else
                                         "materialised cards", "materialised_denominator_frac": mat / peak},
                    "note": "per-stage CUDA-event times on the launching stream.  No stage of this path is HBM-bound: detect / warp are "
                            "instruction-issue bound integer kernels, vseg / categorize are FP32-FMA bound (their fp32_frac is in `stages`)"}
        out = base_line(g, args, METRIC, value, "frames/s", t_ms, "u8/int32 (detect, warp, hseg) + f32 (vseg, digit CNNs)",
                        "100k synthetic 640x480 frames per GPU, full detect->warp->OCR pipeline on %dxB200" % world,
                        {"frames_per_gpu_per_step": F, "width": W, "height": H, "deck_seed": DECK_SEED, "card_mode": args.card_mode,
                         "l2": "inputs (%.1f GB per step) are larger than L2; no flush needed" % (F * FRAME_BYTES / 1e9),
                         "parallelism": "frames sharded across %d GPU(s), NCCL gather of 32-byte digit strings to rank 0%s" % (world, "" if world > 1 else " (n/a at 1 GPU)")})
        out.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "stages": per_stage,
                    "materialised": materialised, "cpu_baseline": cpu, "single_frame": single, "wall_s_timed_region": wall})
        print(json.dumps(out), flush=True)
    g.finish()


# ---------------------------------------------------------------------------------------------------------------------
# configs[2]: n_categorize only
# ---------------------------------------------------------------------------------------------------------------------
def categorize_config(args):
    g = Gpu(args)
    torch, dmz, rank, world = g.torch, g.dmz, g.rank, g.world
    n = args.patches
    # half i.i.d. uniform noise, half digit crops cut from deck cards at the offsets the path itself found
    nf = 4096
    frames = g.deck(0, nf)
    recs = torch.zeros((nf, RECORD_BYTES), dtype=torch.uint8, device="cuda")
    cards = torch.zeros((nf, 270, 428), dtype=torch.uint8, device="cuda")
    dmz.process_frames_device(frames.data_ptr(), nf, W, H, recs.data_ptr(), d_cards=cards.data_ptr())
    r = recs.cpu().numpy().view(g.pkg.RECORD_DTYPE).reshape(nf)
    hc = cards.cpu().numpy()
    crops = [hc[k, int(r["v_y_offset"][k]):int(r["v_y_offset"][k]) + 27, int(x):int(x) + 19]
             for k in np.nonzero(r["usable"] == 1)[0] for x in r["h_offsets"][k][: int(r["h_n_offsets"][k])]]
    crops = torch.from_numpy(np.stack(crops)).cuda()
    del frames, cards, recs
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7 + rank)
    patches = torch.randint(0, 256, (n, 27, 19), dtype=torch.uint8, device="cuda", generator=gen)
    half = n // 2
    for lo in range(half, n, 1 << 20):  # (chunked: index tensors of 5 M int64 x 513 would not be small)
        hi = min(n, lo + (1 << 20))
        patches[lo:hi] = crops[torch.arange(lo - half, hi - half, device="cuda") % crops.shape[0]]
    out = torch.zeros((n, 40), dtype=torch.float32, device="cuda")

    def step():
        dmz.categorize_patches_device(patches.data_ptr(), n, out.data_ptr())

    step()
    l0 = dmz.launches
    t_ms, clocks, wall = g.timed(step, args.steps, args.warmup)
    launches = (dmz.launches - l0) * args.steps // max(args.steps + args.warmup, 1)
    value = n * world * args.steps / (t_ms * 1e-3)

    # e2e: host patches in, 40 floats per patch out, through the same entry point (1 M patches per step)
    e2e = None
    if not args.no_e2e:
        ne = min(n, 1 << 20)
        hp = patches[:ne].cpu().numpy()
        dmz.categorize_patches(hp[:4096])
        te = time.perf_counter()
        for _ in range(args.steps):
            dmz.categorize_patches(hp)
        e2e_s = g.max_over_ranks(time.perf_counter() - te)
        e2e = {"value": ne * world * args.steps / e2e_s, "unit": "patches/s", "h2d_bytes_per_step": ne * 513, "d2h_bytes_per_step": ne * 160,
               "patches_per_step": ne, "note": "pageable numpy buffers through b200_categorize_patches_batch(B200_MEM_HOST)"}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        orc, kind, kind_desc = cpu_checker()
        cores = usable_cpus()
        S = min(1 << 18, n // 2 * 2)
        sample = torch.cat([patches[:S // 2], patches[half:half + S // 2]]).cpu().numpy()
        gpu_out = torch.cat([out[:S // 2], out[half:half + S // 2]]).cpu().numpy()
        orc.bench_patches(sample[: 64 * cores], cores)
        secs, cpu_out = orc.bench_patches(sample, cores)
        cpu = {"value": S / secs, "unit": "patches/s", "cores": cores, "kind": "reference" if kind == "ref" else "port",
               "sample": "%d patches (half noise, half digit crops), %d threads, %.2f s wall (%s)" % (S, cores, secs, kind_desc),
               "parity_vs_gpu_on_sample": {"max_abs_diff_40_outputs": float(np.abs(gpu_out - cpu_out).max()),
                                           "argmax_mismatches": int((gpu_out[:, :10].argmax(1) != cpu_out[:, :10].argmax(1)).sum())}}
    if rank == 0:
        peak, peak_src = measured_peaks()
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = fp32_peak_tflops(g.sm_count, sm_mhz)
        tf = CNN_FLOP_PER_PATCH * n * args.steps / (t_ms * 1e-3) / 1e12
        gbps = 553.0 * n * args.steps / (t_ms * 1e-3) / 1e9
        line = base_line(g, args, "digit patches/sec through n_categorize (3-CNN ensemble)", value, "patches/s", t_ms, "f32",
                         "n_categorize only: 10M synthetic digit patches through models/generated CNN, direct-conv kernel, 1xB200",
                         {"patches_per_gpu_per_step": n, "mix": "half i.i.d. U{0..255}, half 19x27 digit crops from the deck",
                          "l2": "inputs (%.1f GB per step) are larger than L2; no flush needed" % (n * 513 / 1e9)})
        line.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                     "roofline": {"bound": "fp32", "kernel": "digit_prep_kernel + categorize_mma_kernel", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s",
                                  "frac": tf / fp32_peak, "traffic": None,
                                  "peak_source": "FP32 CUDA-core peak = %d SMs x 128 lanes x 2 x %.0f MHz (clock sampled during the run)" % (g.sm_count, sm_mhz),
                                  "flop_per_patch": CNN_FLOP_PER_PATCH,
                                  "hbm": {"alg_bytes_per_patch": 553, "achieved_GBps": gbps, "peak_GBps": peak, "frac": gbps / peak, "peak_source": peak_src},
                                  "note": "218 880 flop per 553 B: ~400 flop/B, far right of the FP32 ridge.  The figure is the reference's FP32 flop count over the "
                                          "CUDA-core FP32 peak (what SURVEY 8d asks for); since round 2 the conv and hidden contractions run on the tensor cores "
                                          "(tcgen05 kind::i8 on exact base-128 weight digits, split-fp16 hidden layer; categorize_mma.cu), which is how the "
                                          "fraction can pass what FFMA alone would reach"},
                     "cpu_baseline": cpu, "wall_s_timed_region": wall})
        print(json.dumps(line), flush=True)
    g.finish()


# ---------------------------------------------------------------------------------------------------------------------
# configs[3]: Canny + Hough sweep
# ---------------------------------------------------------------------------------------------------------------------
def detect_sweep_config(args):
    g = Gpu(args)
    torch, dmz, rank, world = g.torch, g.dmz, g.rank, g.world
    peak, peak_src = measured_peaks()
    per_res, total_frames, total_ms, launches, clocks_all = {}, 0, 0.0, 0, None
    cpu = {}
    for name, (w, h) in DETECT_SIZES.items():
        if args.sizes and name not in args.sizes.split(","):
            continue
        n = args.detect_frames
        free, _ = torch.cuda.mem_get_info()
        resident = max(1, min(n, int(0.55 * free) // (w * h)))   # frames kept in HBM (1080p: 100k frames would need 207 GB)
        call = min(resident, 25000 if w * h > 1000000 else resident)  # frames per call (1080p keeps its gradients in a global scratch)
        frames = g.deck(rank * n, resident, w, h)
        lines = torch.zeros((call, 4, 36), dtype=torch.uint8, device="cuda")
        passes = [(f0 % resident, min(call, resident - f0 % resident, n - f0)) for f0 in range(0, n, call)]

        def step():
            for lo, cnt in passes:
                dmz.detect_lines_device(frames[lo:].data_ptr(), cnt, w, h, lines.data_ptr())

        done = sum(c for _, c in passes)
        step()
        l0 = dmz.launches
        t_ms, clocks, _ = g.timed(step, args.steps, args.warmup)
        launches += (dmz.launches - l0) * args.steps // max(args.steps + args.warmup, 1)
        clocks_all = clocks if clocks_all is None else clocks_all
        if rank == 0:
            orc, kind, kind_desc = cpu_checker()
            boxes = orc.detection_boxes(w, h)
            strip_px = int(sum(int(b[2]) * int(b[3]) for b in boxes))
            alg = strip_px + 48
            fps = done * world * args.steps / (t_ms * 1e-3)
            gbps = alg * fps / world / 1e9
            first = np.frombuffer(lines[:min(call, 1000)].cpu().numpy().tobytes(), dtype=g.pkg.LINE_DTYPE)
            per_res[name] = {"frames/s": fps, "ms_per_100k_frames": t_ms / args.steps / done * 1e5, "frames_per_step": done, "resident_frames": resident,
                             "strip_px": strip_px, "alg_bytes_per_frame": alg, "achieved_GBps": gbps, "frac_of_hbm": gbps / peak,
                             "lines_found_of_first_%d" % len(first): int(first["found"].sum()),
                             "l2": "resident set %.1f GB >> L2" % (resident * w * h / 1e9)}
            if not args.no_cpu and world == 1:
                cores = usable_cpus()
                S = min({"480p": 4096, "720p": 2048, "1080p": 1024}[name], resident)
                sample = frames[:S].cpu().numpy()
                orc.bench_detect(sample[: max(cores, 64)], cores)
                secs, nfound = orc.bench_detect(sample, cores)
                k64 = min(64, S)
                gl = dmz.detect_lines(sample[:k64])
                mism = 0
                for k in range(k64):
                    for s_, (x, y, ww, hh) in enumerate(boxes):
                        o = orc.best_line(sample[k][y:y + hh, x:x + ww], s_ >= 2)
                        got = gl[k, s_]
                        mism += int(got["found"]) != o.found or bool(o.found and (int(got["r"]), int(got["n"])) != (o.r, o.n))
                cpu[name] = {"frames/s": S / secs, "sample_frames": S, "cores": cores, "frames_with_all_edges": nfound,
                             "line_mismatches_vs_gpu_first_%d_frames" % k64: int(mism)}
        total_frames += done
        total_ms += t_ms
        del frames, lines
        torch.cuda.empty_cache()
    if rank == 0:
        value = total_frames * world * args.steps / (total_ms * 1e-3)
        dom = max(per_res, key=lambda k: per_res[k]["achieved_GBps"]) if per_res else None
        cpu_block = None
        if cpu:
            tot_s = sum(v["sample_frames"] / v["frames/s"] for v in cpu.values())
            cpu_block = {"value": sum(v["sample_frames"] for v in cpu.values()) / tot_s, "unit": "frames/s", "cores": usable_cpus(),
                         "kind": "reference" if cpu_checker()[1] == "ref" else "port",
                         "sample": "dmz_detect_edges on 4096 / 2048 / 1024 frames (480p / 720p / 1080p), all host threads", "per_resolution": cpu}
        line = base_line(g, args, "frames/sec Canny+Hough (best_line_for_sample x 4 strips), 480p/720p/1080p sweep", value, "frames/s", total_ms,
                         "u8/int32", "Canny+Hough sweep 480p/720p/1080p x 100k frames each, achieved HBM GB/s vs roofline, 1xB200",
                         {"frames_per_resolution_per_step": args.detect_frames, "value_is": "all frames of the sweep / total time"})
        line.update({"e2e": None, "gpu_launches": int(launches), "clocks": clocks_all, "sweep": per_res,
                     "roofline": None if not dom else {"bound": "hbm", "kernel": "detect_strips_kernel (%s)" % dom, "achieved": per_res[dom]["achieved_GBps"], "peak": peak,
                                                        "unit": "GB/s", "frac": per_res[dom]["frac_of_hbm"], "traffic": None, "peak_source": peak_src,
                                                        "note": "algorithmic bytes = strip pixels + 48 B (SURVEY 8d); the kernel is instruction-issue bound integer work in "
                                                                "shared memory, see DESIGN.md"},
                     "cpu_baseline": cpu_block})
        print(json.dumps(line), flush=True)
    g.finish()


# ---------------------------------------------------------------------------------------------------------------------
# pixel formats either side of the path (widening rows): dmz_YCbCr_to_RGB, dmz_deinterleave_RGBA_to_R, Cython stencils
# ---------------------------------------------------------------------------------------------------------------------
def formats_config(args):
    g = Gpu(args)
    torch, dmz, rank, world = g.torch, g.dmz, g.rank, g.world
    lib, ctx, MEM_DEVICE = dmz.lib, dmz.ctx, g.pkg.MEM_DEVICE
    n = args.format_frames
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11 + rank)
    planes = torch.randint(0, 256, (3, n, H, W), dtype=torch.uint8, device="cuda", generator=gen)
    rgb = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")  # also the RGBA source of the R extraction
    st = torch.empty((n, H, W), dtype=torch.int16, device="cuda")
    nc = n * 2  # warped cards: 428 wide, so only 4-pixel vectors apply
    cplanes = torch.randint(0, 256, (3, nc, 270, 428), dtype=torch.uint8, device="cuda", generator=gen)
    crgb = torch.empty((nc, 270, 428, 3), dtype=torch.uint8, device="cuda")
    px, cpx = n * W * H, nc * 428 * 270

    def ycc(ch):
        return lambda: dmz._check(lib.b200_ycbcr_to_rgb_batch(ctx, planes[0].data_ptr(), W, W * H, planes[1].data_ptr(), planes[2].data_ptr(), W, W * H,
                                                              W, H, n, ch, MEM_DEVICE, rgb.data_ptr()))

    legs = {  # name -> (call, algorithmic bytes per launch, pixels)
        "ycbcr_to_rgb_640x480": (ycc(3), px * 6, px),
        "ycbcr_to_rgba_640x480": (ycc(4), px * 7, px),
        "ycbcr_to_rgb_card_428x270": (lambda: dmz._check(lib.b200_ycbcr_to_rgb_batch(ctx, cplanes[0].data_ptr(), 428, 428 * 270, cplanes[1].data_ptr(),
                                                                                    cplanes[2].data_ptr(), 428, 428 * 270, 428, 270, nc, 3, MEM_DEVICE,
                                                                                    crgb.data_ptr())), cpx * 6, cpx),
        "rgba_to_r": (lambda: dmz._check(lib.b200_rgba_to_r_batch(ctx, rgb.data_ptr(), px, MEM_DEVICE, st.data_ptr())), px * 5, px),
    }
    for kind, name in enumerate(("scharr3_dx_abs", "scharr3_dy_abs", "sobel3_dx_dy")):
        legs[name + "_640x480"] = ((lambda k: lambda: dmz._check(lib.b200_stencil3_batch(ctx, planes[0].data_ptr(), W, W * H, W, H, n, k, MEM_DEVICE,
                                                                                         st.data_ptr())))(kind), px * 3, px)
    peak, peak_src = measured_peaks()
    results, clocks, launches, wall_total = {}, None, 0, 0.0
    for name, (call, nbytes, pixels) in legs.items():
        l0 = dmz.launches
        t_ms, ck, wall = g.timed(call, args.steps, args.warmup)
        launches += (dmz.launches - l0) * args.steps // max(args.steps + args.warmup, 1)
        wall_total += wall
        gbps = nbytes * args.steps / (t_ms * 1e-3) / 1e9
        results[name] = {"ms_per_launch": t_ms / args.steps, "pixels/s": pixels * world * args.steps / (t_ms * 1e-3), "alg_bytes_per_pixel": nbytes // pixels,
                         "achieved_GBps": gbps, "frac_of_hbm_peak": gbps / peak}
        if name == "ycbcr_to_rgb_640x480":
            clocks, head_ms = ck, t_ms

    e2e = None
    if not args.no_e2e:
        ne = min(n, 256)
        hp = planes[:, :ne].cpu().numpy()
        dmz.ycbcr_to_rgb(hp[0][:8], hp[1][:8], hp[2][:8])
        te = time.perf_counter()
        for _ in range(args.steps):
            dmz.ycbcr_to_rgb(hp[0], hp[1], hp[2])
        e2e_s = g.max_over_ranks(time.perf_counter() - te)
        e2e = {"value": ne * W * H * world * args.steps / e2e_s, "unit": "pixels/s", "h2d_bytes_per_step": 3 * ne * W * H, "d2h_bytes_per_step": 3 * ne * W * H,
               "frames_per_step": ne, "note": "pageable numpy planes through b200_ycbcr_to_rgb_batch(B200_MEM_HOST)"}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        orc, kind, kind_desc = cpu_checker()
        S = 48
        hp = planes[:, :S].cpu().numpy()
        dmz._check(lib.b200_ycbcr_to_rgb_batch(ctx, planes[0].data_ptr(), W, W * H, planes[1].data_ptr(), planes[2].data_ptr(), W, W * H, W, H, S, 3,
                                               MEM_DEVICE, rgb.data_ptr()))
        gpu_rgb = rgb.view(-1)[: S * H * W * 3].cpu().numpy().reshape(S, H, W, 3)
        t0 = time.perf_counter()
        cpu_rgb = np.stack([orc.ycbcr_to_rgb(hp[0, k], hp[1, k], hp[2, k], 3) for k in range(S)])
        secs = time.perf_counter() - t0
        cpu = {"value": S * W * H / secs, "unit": "pixels/s", "cores": 1, "kind": "reference" if kind == "ref" else "port",
               "sample": "%d random 640x480 frames, one thread, %.2f s wall (%s)" % (S, secs, kind_desc),
               "parity_vs_gpu_on_sample": {"mismatching_bytes": int((gpu_rgb != cpu_rgb).sum())}}
    if rank == 0:
        head = results["ycbcr_to_rgb_640x480"]
        line = base_line(g, args, "pixels/sec through dmz_YCbCr_to_RGB (640x480 planes -> interleaved RGB)", head["pixels/s"], "pixels/s", head_ms, "u8",
                         "pixel formats around the path: dmz_YCbCr_to_RGB on 640x480 planes, 1xB200 (+ RGBA->R and the Cython stencils as side legs)",
                         {"frames_per_gpu_per_launch": n, "cards_per_gpu_per_launch": nc, "data_note": "i.i.d. U{0..255} planes (the arithmetic is data independent)",
                          "l2": "every leg streams %.1f GB or more per launch: larger than L2, no flush needed" % (px * 3 / 1e9)})
        line.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                     "roofline": {"bound": "hbm", "kernel": "ycbcr_to_rgb_kernel<16, 3>", "achieved": head["achieved_GBps"], "peak": peak, "unit": "GB/s",
                                  "frac": head["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                                  "alg_bytes": "3 B read + 3 B written per pixel (4 written for RGBA; RGBA->R 4 + 1; stencils 1 + 2)"},
                     "legs": results, "cpu_baseline": cpu, "wall_s_timed_region": wall_total})
        print(json.dumps(line), flush=True)
    g.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="pipeline", choices=["pipeline", "categorize", "detect-sweep", "formats"])
    ap.add_argument("--frames", type=int, default=100000, help="frames per GPU per step (BASELINE configs[1]: 100k)")
    ap.add_argument("--e2e-frames", type=int, default=100000, help="frames per step on the host-buffer (e2e) path (clipped to what host memory allows)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="frames in the bounded CPU-baseline sample")
    ap.add_argument("--patches", type=int, default=10000000, help="--config categorize: patches per GPU per step (BASELINE configs[2]: 10 M)")
    ap.add_argument("--detect-frames", type=int, default=100000, help="--config detect-sweep: frames per resolution per step")
    ap.add_argument("--format-frames", type=int, default=8192, help="--config formats: 640x480 frames per GPU per launch")
    ap.add_argument("--sizes", default="", help="--config detect-sweep: comma-separated subset of 480p,720p,1080p")
    ap.add_argument("--card-mode", default="lazy", choices=["lazy", "full"],
                    help="lazy (library default): no cards_out -> only the card rows the scan reads are warped; full: every card materialised")
    ap.add_argument("--no-materialised", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    elif args.config == "categorize":
        categorize_config(args)
    elif args.config == "formats":
        formats_config(args)
    elif args.config == "detect-sweep":
        detect_sweep_config(args)
    else:
        pipeline_config(args)


if __name__ == "__main__":
    main()
