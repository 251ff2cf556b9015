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


# ---- synthetic expiry text (SURVEY 8f rank 4 tests): seven-segment digits and a two-pixel slash stamped on a warped card.
# Plain numpy so the same cards can be rebuilt wherever the tests run; the reference's segmentation does find MM/YY groups
# on them (about 40 % of the random layouts below), which is all the parity tests need.
_SEGMENTS = {0: "abcdef", 1: "bc", 2: "abged", 3: "abgcd", 4: "fgbc", 5: "afgcd", 6: "afgedc", 7: "abc", 8: "abcdefg", 9: "abfgcd"}


def expiry_glyph(ch, w=9, h=15, t=2):
    g = np.zeros((h, w), np.uint8)
    if ch == " ":
        return g
    if ch == "/":
        for y in range(h):
            x = int(round((w - 2) * (1 - y / (h - 1))))
            g[y, max(0, x):x + 2] = 1
        return g
    s, m = _SEGMENTS[int(ch)], h // 2
    if "a" in s: g[0:t, :] = 1
    if "g" in s: g[m - t // 2:m - t // 2 + t, :] = 1
    if "d" in s: g[h - t:h, :] = 1
    if "f" in s: g[0:m + 1, 0:t] = 1
    if "b" in s: g[0:m + 1, w - t:w] = 1
    if "e" in s: g[m:h, 0:t] = 1
    if "c" in s: g[m:h, w - t:w] = 1
    return g


def expiry_card(base_card, y_offset0, seed):
    """A 428x270 card with 1-3 lines of digits / MM/YY text below the number row; returns (card, y_offset)."""
    rng = np.random.default_rng(seed)
    c = base_card.copy()
    if seed % 4 == 3:
        c = np.clip(c.astype(np.int32) + rng.integers(-12, 13, c.shape), 0, 255).astype(np.uint8)
    yo = y_offset0 + int(rng.integers(-20, 10))
    for _ in range(int(rng.integers(1, 4))):
        txt = "".join(rng.choice(list("0123456789"), int(rng.integers(0, 6)))) + " " * int(rng.integers(0, 2))
        txt += "%02d/%02d" % (rng.integers(1, 13), rng.integers(15, 40)) + " " * int(rng.integers(0, 2))
        txt += "".join(rng.choice(list("0123456789 "), int(rng.integers(0, 8))))
        x, y = int(rng.integers(5, 250)), yo + 27 + int(rng.integers(3, 70))
        if y + 16 > 270:
            continue
        w, h, t = [(9, 15, 2), (8, 14, 2), (9, 15, 3), (7, 13, 2)][int(rng.integers(0, 4))]
        pitch, fg = int(rng.integers(11, 14)), int(rng.choice([30, 60, 240, 255]))
        for i, ch in enumerate(txt):
            xs = x + i * pitch
            if xs + w > 428:
                break
            reg = c[y:y + h, xs:xs + w]
            reg[expiry_glyph(ch, w, h, t) > 0] = fg
    return np.ascontiguousarray(c), yo


# ---- frames for the other three FrameOrientations (dmz_olm.h:16-22).  The deck generator renders landscape-right
# cards only; here a warped 428x270 card is projected back into a frame so that its quad sits in the guide rectangle of
# the requested orientation and dmz_transform_card's corner permutation (dmz.cpp:446-471) turns it upright again.
def _homography(src, dst):
    """3x3 H with H @ (x, y, 1) ~ (u, v, 1) for the four correspondences src[i] -> dst[i] (float64, numpy solve)."""
    a, b = [], []
    for (x, y), (u, v) in zip(src, dst):
        a.append([x, y, 1, 0, 0, 0, -x * u, -y * u]); b.append(u)
        a.append([0, 0, 0, x, y, 1, -x * v, -y * v]); b.append(v)
    h = np.linalg.solve(np.array(a, np.float64), np.array(b, np.float64))
    return np.append(h, 1.0).reshape(3, 3)


def guide_quad(boxes):
    """Centre lines of the four detection strips (top, bottom, left, right rows of detection_boxes) -> tl, bl, tr, br."""
    top = boxes[0][1] + boxes[0][3] / 2.0
    bot = boxes[1][1] + boxes[1][3] / 2.0
    left = boxes[2][0] + boxes[2][2] / 2.0
    right = boxes[3][0] + boxes[3][2] / 2.0
    return np.array([[left, top], [left, bot], [right, top], [right, bot]], np.float64)


def oriented_frames(cards, orientation, boxes, w=640, h=480, jitter=None, seed=1):
    """cards: (n, 270, 428) u8 upright cards.  Returns (n, h, w) u8 frames showing each card as a camera held in
    `orientation` sees it, corners jittered inside the detection strips given by `boxes` (oracle.detection_boxes)."""
    rng = np.random.default_rng(seed)
    quad = guide_quad(boxes)  # tl, bl, tr, br in the frame
    if jitter is None:
        jitter = 0.3 * min(boxes[0][3], boxes[2][2])
    # card corner (0,0),(427,0),(0,269),(427,269) <- frame corner, per dmz_transform_card's permutation
    perm = {1: (1, 0, 3, 2), 4: (3, 1, 2, 0), 2: (2, 3, 0, 1), 3: (0, 2, 1, 3)}[orientation]
    card_pts = [(0, 0), (427, 0), (0, 269), (427, 269)]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.zeros((len(cards), h, w), np.uint8)
    for k, card in enumerate(cards):
        q = quad + rng.uniform(-jitter, jitter, quad.shape)
        H = _homography([tuple(q[i]) for i in perm], card_pts)  # frame -> card
        den = H[2, 0] * xx + H[2, 1] * yy + H[2, 2]
        u = (H[0, 0] * xx + H[0, 1] * yy + H[0, 2]) / den
        v = (H[1, 0] * xx + H[1, 1] * yy + H[1, 2]) / den
        inside = (u >= 0) & (u <= 427) & (v >= 0) & (v <= 269)
        u0 = np.clip(np.floor(u), 0, 426).astype(np.int64); v0 = np.clip(np.floor(v), 0, 268).astype(np.int64)
        fu = np.clip(u - u0, 0, 1); fv = np.clip(v - v0, 0, 1)
        c = card.astype(np.float64)
        val = (c[v0, u0] * (1 - fu) + c[v0, u0 + 1] * fu) * (1 - fv) + (c[v0 + 1, u0] * (1 - fu) + c[v0 + 1, u0 + 1] * fu) * fv
        bg = np.clip(rng.normal(60, 8, (h, w)), 0, 255)
        out[k] = np.where(inside, np.clip(np.rint(val), 0, 255), bg).astype(np.uint8)
    return out
