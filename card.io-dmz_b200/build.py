"""Build the in-tree native libraries (sm_100a only).  Used by __graft_entry__.build() and by hand:

    python card.io-dmz_b200/build.py

Outputs (git-ignored, but they travel with gpurun snapshots):
    card.io-dmz_b200/libb200dmz.so     the product: CUDA kernels + C ABI (include/b200_dmz.h)
    tools/deck/libdeck_cuda.so          bench/test support: synthetic deck generator on the GPU
    tools/deck/libdeck_cpu.so           the same generator for the CPU checkers
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.join(CSRC, "b200_internal.h"), os.path.join(CSRC, "expiry_seg_core.h"), os.path.join(CSRC, "umma.cuh"), os.path.join(ROOT, "include", "b200_dmz.h")]
    units = [
        # (source, extra flags).  exact.cu / detect.cu: bit-exact float stages -> no FMA contraction.
        ("detect.cu", ["-fmad=false"]),
        ("exact.cu", ["-fmad=false"]),
        ("warp.cu", ["-fmad=false"]),
        ("expiry_seg.cu", ["-fmad=false"]),
        ("nets.cu", []),
        ("vseg_mma.cu", []),
        ("categorize_mma.cu", []),
        ("expiry_mma.cu", []),
        ("formats.cu", []),
        ("api.cu", ["-fmad=false"]),
        ("b200_tables.cpp", ["-Xcompiler", "-ffp-contract=off"]),
        ("scanner.cpp", ["-Xcompiler", "-ffp-contract=off"]),
        ("dmz_compat.cpp", ["-Xcompiler", "-ffp-contract=off"]),
    ]
    objs = []
    for src, extra in units:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(obj_dir, src + ".o")
        if force or newer(o, [s] + hdrs):
            run([nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o])
        objs.append(o)
    lib = os.path.join(HERE, "libb200dmz.so")
    if force or newer(lib, objs):
        run([nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart", "-ldl"])
    # the SDK call sequence timed through the C++ drop-in layer (bench.py single_frame.dropin_sequence)
    lat = os.path.join(obj_dir, "dropin_latency")
    lat_src = os.path.join(ROOT, "tools", "dropin_latency.cpp")
    if force or newer(lat, [lat_src, lib, os.path.join(ROOT, "include", "dmz_b200_compat.h")]):
        run(["g++", "-std=c++14", "-O2", "-I" + os.path.join(ROOT, "include"), lat_src, "-o", lat, "-L" + HERE, "-lb200dmz",
             "-Wl,-rpath,$ORIGIN/..", "-ldl"])
    # device-side checks of the primitives (tools/microbench: tcgen05 kind::i8 against a host product; FFMA / FFMA2 rates;
    # h2d_pitched: what the host-buffer path's upload rectangle can reach over PCIe -- DMA, several streams, zero-copy, hybrids)
    mb = os.path.join(ROOT, "tools", "microbench")
    for name, extra in (("umma_i8", ["-I" + CSRC]), ("ffma2", []), ("h2d_pitched", ["-std=c++17"])):
        src, exe = os.path.join(mb, name + ".cu"), os.path.join(mb, name)
        if os.path.exists(src) and (force or newer(exe, [src, os.path.join(CSRC, "umma.cuh")])):
            run([nvcc] + ARCH + ["-O2", "-o", exe, src] + extra)
    # deck generators (support code)
    deck = os.path.join(ROOT, "tools", "deck")
    dsrc = [os.path.join(deck, f) for f in ("deck_gen.h", "glyphs.h")]
    lib_cpu = os.path.join(deck, "libdeck_cpu.so")
    if force or newer(lib_cpu, dsrc + [os.path.join(deck, "deck_cpu.c")]):
        run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", lib_cpu,
             os.path.join(deck, "deck_cpu.c"), "-lpthread"])
    lib_gpu = os.path.join(deck, "libdeck_cuda.so")
    cu = os.path.join(deck, "deck_cuda.cu")
    if os.path.exists(cu) and (force or newer(lib_gpu, dsrc + [cu])):
        run([nvcc] + ARCH + ["-O3", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-o", lib_gpu, cu, "-lcudart"])
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv)
