"""Shared helpers for the test-suite: package loader, deck generator bindings, comparison utilities."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DECK_SEED = 0xCA2D10


def load_pkg():
    """Import card.io-dmz_b200 (the directory name is not a Python identifier)."""
    name = "cardio_dmz_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "card.io-dmz_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "card.io-dmz_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_deck_cpu = None


def deck_cpu_lib():
    global _deck_cpu
    if _deck_cpu is None:
        lib = C.CDLL(os.path.join(ROOT, "tools", "deck", "libdeck_cpu.so"))
        lib.deck_render_cpu.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]
        lib.deck_truth.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _deck_cpu = lib
    return _deck_cpu


def deck_frames(first, n, w=640, h=480, jitter=8.0, seed=DECK_SEED, threads=8):
    out = np.zeros((n, h, w), np.uint8)
    deck_cpu_lib().deck_render_cpu(seed, first, n, w, h, jitter, out.ctypes.data, threads)
    return out


def deck_truth(frame, w=640, h=480, jitter=8.0, seed=DECK_SEED):
    digits = np.zeros(16, np.uint8)
    n = C.c_int32()
    quad = np.zeros(8)
    deck_cpu_lib().deck_truth(seed, frame, w, h, jitter, digits.ctypes.data, C.byref(n), quad.ctypes.data)
    return digits[: n.value].copy(), quad


def deck_frames_cuda(first, n, w=640, h=480, jitter=8.0, seed=DECK_SEED):
    """Render on the GPU into a torch uint8 tensor (n, h, w) on cuda:0."""
    import torch
    lib = C.CDLL(os.path.join(ROOT, "tools", "deck", "libdeck_cuda.so"))
    lib.deck_render_cuda.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    out = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rc = lib.deck_render_cuda(seed, first, n, w, h, jitter, out.data_ptr(), None)
    if rc != 0:
        raise RuntimeError("deck_render_cuda failed: %d" % rc)
    return out


def synthetic_strip(rng, w, h, vertical, kind="edge"):
    """Small structured test images for the detect stage."""
    img = np.clip(rng.normal(60, 8, (h, w)), 0, 255)
    if kind == "noise":
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == "flat":
        return np.full((h, w), 77, np.uint8)
    # a tilted step edge through the strip
    yy, xx = np.mgrid[0:h, 0:w]
    ang = np.deg2rad(rng.uniform(-4, 4))
    if vertical:
        d = (xx - w / 2 - rng.uniform(-w / 4, w / 4)) * np.cos(ang) + (yy - h / 2) * np.sin(ang)
    else:
        d = (yy - h / 2 - rng.uniform(-h / 4, h / 4)) * np.cos(ang) + (xx - w / 2) * np.sin(ang)
    img = np.where(d > 0, 175 + rng.normal(0, 6, (h, w)), img)
    return np.clip(img, 0, 255).astype(np.uint8)
