"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  * ``Oracle("port")``  -> oracle/liboracle.so        (plain-C restatement, oracle/dmz_oracle.c)
  * ``Oracle("ref")``   -> oracle/_ref/libdmz_ref.so  (the reference's own sources + cvshim)

Both expose the same stage taps (prefix ``orc_`` / ``ref_``), so a test can run either.  Only
tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product path never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WEIGHTS_DIR = os.path.join(ROOT, "card.io-dmz_b200", "weights")


class VSeg(C.Structure):
    _fields_ = [("score", C.c_float), ("y_offset", C.c_uint16), ("pattern_type", C.c_uint8),
                ("number_pattern", C.c_uint8 * 19), ("number_pattern_length", C.c_uint8),
                ("number_length", C.c_uint8)]


class HSeg(C.Structure):
    _fields_ = [("n_offsets", C.c_uint8), ("offsets", C.c_uint16 * 16), ("score", C.c_float),
                ("number_width", C.c_float), ("pattern_offset", C.c_uint16)]


class Line(C.Structure):
    _fields_ = [("found", C.c_int32), ("r", C.c_int32), ("n", C.c_int32), ("max_votes", C.c_int32),
                ("low", C.c_int32), ("high", C.c_int32), ("n_edge_px", C.c_int32),
                ("rho", C.c_float), ("theta", C.c_float)]


class Detect(C.Structure):
    _fields_ = [("found", C.c_int32 * 4), ("rho", C.c_float * 4), ("theta", C.c_float * 4),
                ("corners", C.c_float * 8), ("all_found", C.c_int32)]


class Scan(C.Structure):
    _fields_ = [("scores", C.c_float * 160), ("hseg", HSeg), ("vseg", VSeg),
                ("usable", C.c_uint8), ("upside_down", C.c_uint8), ("pad", C.c_uint8 * 2)]


class FrameRecord(C.Structure):
    _fields_ = [("detect", Detect), ("scan", Scan), ("card_check", C.c_uint32)]


assert C.sizeof(VSeg) == 28 and C.sizeof(HSeg) == 48 and C.sizeof(Scan) == 720, (
    C.sizeof(VSeg), C.sizeof(HSeg), C.sizeof(Scan))

RECORD_DTYPE = np.dtype([
    ("found", "<i4", 4), ("rho", "<f4", 4), ("theta", "<f4", 4), ("corners", "<f4", 8), ("all_found", "<i4"),
    ("scores", "<f4", 160),
    ("h_n_offsets", "u1"), ("_p0", "u1"), ("h_offsets", "<u2", 16), ("_p1", "u1", 2), ("h_score", "<f4"),
    ("h_number_width", "<f4"), ("h_pattern_offset", "<u2"), ("_p2", "u1", 2),
    ("v_score", "<f4"), ("v_y_offset", "<u2"), ("v_pattern_type", "u1"), ("v_number_pattern", "u1", 19),
    ("v_number_pattern_length", "u1"), ("v_number_length", "u1"),
    ("usable", "u1"), ("upside_down", "u1"), ("_p4", "u1", 2),
    ("card_check", "<u4"),
])
assert RECORD_DTYPE.itemsize == C.sizeof(FrameRecord), (RECORD_DTYPE.itemsize, C.sizeof(FrameRecord))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lib_path(kind):
    """port: the plain-C restatement; ref: the reference's sources (SCAN_EXPIRY=0, re-entrant); refx: the same sources with
    SCAN_EXPIRY=1 (adds the expiry taps; function statics make it single-threaded)."""
    return {"port": os.path.join(HERE, "liboracle.so"), "ref": os.path.join(HERE, "_ref", "libdmz_ref.so"),
            "refx": os.path.join(HERE, "_ref", "libdmz_ref_expiry.so"),
            # timing only: the reference's sources at -O3 -march=x86-64-v3 (needs AVX2 + FMA on the host)
            "refo3": os.path.join(HERE, "_ref", "libdmz_ref_o3.so")}[kind]


def available(kind):
    if kind == "refo3":
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            return False
        if " avx2" not in flags or " fma" not in flags or " bmi2" not in flags:
            return False
    return os.path.exists(lib_path(kind))


class Oracle:
    """Uniform front-end over either checker library."""

    def __init__(self, kind="port"):
        assert kind in ("port", "ref", "refx", "refo3")
        self.kind = kind
        self.prefix = "orc_" if kind == "port" else "ref_"
        self.lib = C.CDLL(lib_path(kind))
        if kind == "port":
            self.lib.orc_load_weights.argtypes = [C.c_char_p]
            rc = self.lib.orc_load_weights(WEIGHTS_DIR.encode())
            if rc != 0:
                raise RuntimeError("oracle: cannot load weights from " + WEIGHTS_DIR)
        f = self._f
        vp, i, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
        f("detection_boxes", [i, i, i, vp])
        f("sobel7", [vp, i, i, i, vp, vp])
        f("adaptive_canny", [vp, i, i, i, vp, vp, vp, vp, vp])
        f("best_line", [vp, i, i, i, i, C.POINTER(Line)])
        f("detect_edges", [vp, i, i, i, vp, vp, i, i, C.POINTER(Detect)], i)
        f("calc_persp_transform", [vp, vp, vp])
        f("transform_card", [vp, i, i, i, vp, i, vp])
        f("transform_card_up", [vp, i, i, i, vp, i, i, vp])
        f("vseg_row", [vp, i, vp])
        f("vseg_model", [vp, vp])
        f("best_n_vseg", [vp, C.POINTER(VSeg)])
        f("best_n_hseg", [vp, C.POINTER(VSeg), C.POINTER(HSeg)])
        f("number_scores", [vp, i, C.POINTER(HSeg), vp])
        f("digit_patch_prep", [vp, i, vp])
        f("digit_models", [vp, vp])
        f("scan_card_image", [vp, C.POINTER(Scan)])
        f("process_frame", [vp, i, i, i, vp, vp, i, i, vp, vp])
        f("scanner_new", [], vp)
        f("scanner_free", [vp])
        f("scanner_reset", [vp])
        f("scanner_add_frame", [vp, vp, C.POINTER(Scan)])
        f("scanner_peek", [vp, vp, vp, vp])
        f("scanner_result", [vp, vp, C.POINTER(C.c_int32)], i)
        f("focus_score", [vp, i, i, i, i], C.c_float)
        f("brightness_score", [vp, i, i, i, i], C.c_float)
        f("scoring_rect", [i, i, i, vp])
        f("ycbcr_to_rgb", [vp, i, vp, vp, i, i, i, i, vp, i])
        f("rgba_to_r", [vp, vp, C.c_size_t])
        f("stencil3", [vp, i, i, i, i, vp])
        f("luhn", [vp, i], i)
        f("card_type", [vp, i], i)
        f("bench_frames", [vp, i, i, i, vp, vp, i, i, vp], C.c_double)
        f("bench_stages", [vp, i, i, i, vp, vp, i, vp, vp])
        f("bench_detect", [vp, i, i, i, vp, vp, i, i, vp], C.c_double)
        f("bench_patches", [vp, i, i, vp], C.c_double)
        if kind in ("ref", "refx", "refo3"):
            self.lib.ref_run_kats.restype = i
            self.lib.ref_sizeof.argtypes = [i]
            if kind == "refx":
                self.lib.ref_expiry_patch_prep.argtypes = [vp, i, i, i, i, i, vp]
                self.lib.ref_expiry_digit_model.argtypes = [vp, vp]
                self.lib.ref_slash_model.argtypes = [vp, vp]
                self.lib.ref_scharr3_dx_abs.argtypes = [vp, i, i, i, vp]
                self.lib.ref_best_expiry_seg.argtypes = [vp, i, vp, i, vp]
                self.lib.ref_scanner_add_frame_with_expiry.argtypes = [vp, vp, i, vp, vp]
                self.lib.ref_scanner_expiry_peek.argtypes = [vp, vp, vp, vp, vp, i]
                self.lib.ref_scanner_result_expiry.argtypes = [vp, vp, vp, vp, vp]
                self.lib.ref_expiry_month_year.argtypes = [vp, vp, vp]
        else:
            self.lib.orc_expiry_patch_prep.argtypes = [vp, i, vp]
            self.lib.orc_expiry_digit_model.argtypes = [vp, vp, vp, vp, vp]
            self.lib.orc_scanner_add_scan.argtypes = [vp, C.POINTER(Scan)]
            self.lib.orc_card_check.argtypes = [vp, C.c_size_t]
            self.lib.orc_card_check.restype = C.c_uint32

    def _f(self, name, argtypes, restype=None):
        fn = getattr(self.lib, self.prefix + name)
        fn.argtypes = argtypes
        fn.restype = restype
        setattr(self, "_" + name, fn)

    # ---- stage taps -------------------------------------------------------------------------
    def detection_boxes(self, w, h, orientation=3):
        out = np.zeros(16, np.int32)
        self._detection_boxes(w, h, orientation, _p(out))
        return out.reshape(4, 4)  # top, bottom, left, right  x  (x, y, w, h)

    def sobel7(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        dx = np.zeros((h, w), np.int16)
        dy = np.zeros((h, w), np.int16)
        self._sobel7(_p(img), w, w, h, _p(dx), _p(dy))
        return dx, dy

    def adaptive_canny(self, img, dx, dy):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        dx = np.ascontiguousarray(dx, np.int16)
        dy = np.ascontiguousarray(dy, np.int16)
        edges = np.zeros((h, w), np.uint8)
        lo, hi = C.c_int32(), C.c_int32()
        self._adaptive_canny(_p(img), w, w, h, _p(dx), _p(dy), _p(edges), C.addressof(lo), C.addressof(hi))
        return edges, lo.value, hi.value

    def best_line(self, img, vertical):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = Line()
        self._best_line(_p(img), w, w, h, int(vertical), C.byref(out))
        return out

    def detect_edges(self, y, cb=None, cr=None, orientation=3):
        y = np.ascontiguousarray(y, np.uint8)
        h, w = y.shape
        if cb is None:
            cb = np.full((h // 2, w // 2), 128, np.uint8)
        if cr is None:
            cr = np.full((h // 2, w // 2), 128, np.uint8)
        cb = np.ascontiguousarray(cb, np.uint8)
        cr = np.ascontiguousarray(cr, np.uint8)
        out = Detect()
        self._detect_edges(_p(y), w, h, w, _p(cb), _p(cr), w // 2, orientation, C.byref(out))
        return out

    def calc_persp_transform(self, src, dst):
        src = np.ascontiguousarray(src, np.float32).reshape(8)
        dst = np.ascontiguousarray(dst, np.float32).reshape(8)
        m = np.zeros(9, np.float32)
        self._calc_persp_transform(_p(src), _p(dst), _p(m))
        return m.reshape(3, 3)

    def transform_card(self, y, corners, orientation=3, upsample=False):
        y = np.ascontiguousarray(y, np.uint8)
        h, w = y.shape
        corners = np.ascontiguousarray(corners, np.float32).reshape(8)
        card = np.zeros((270, 428), np.uint8)
        self._transform_card_up(_p(y), w, h, w, _p(corners), orientation, int(upsample), _p(card))
        return card

    def vseg_row(self, card, row):
        card = np.ascontiguousarray(card, np.uint8)
        out = np.zeros(3, np.float32)
        self._vseg_row(_p(card), int(row), _p(out))
        return out

    def vseg_model(self, x204):
        x = np.ascontiguousarray(x204, np.float32)
        out = np.zeros(3, np.float32)
        self._vseg_model(_p(x), _p(out))
        return out

    def best_n_vseg(self, card):
        card = np.ascontiguousarray(card, np.uint8)
        out = VSeg()
        self._best_n_vseg(_p(card), C.byref(out))
        return out

    def best_n_hseg(self, card, vseg):
        card = np.ascontiguousarray(card, np.uint8)
        out = HSeg()
        self._best_n_hseg(_p(card), C.byref(vseg), C.byref(out))
        return out

    def number_scores(self, card, y_offset, hseg):
        card = np.ascontiguousarray(card, np.uint8)
        out = np.zeros((16, 10), np.float32)
        self._number_scores(_p(card), int(y_offset), C.byref(hseg), _p(out))
        return out

    def digit_patch_prep(self, img27x19):
        img = np.ascontiguousarray(img27x19, np.uint8)
        assert img.shape == (27, 19)
        out = np.zeros((27, 19), np.float32)
        self._digit_patch_prep(_p(img), 19, _p(out))
        return out

    def digit_models(self, patch):
        patch = np.ascontiguousarray(patch, np.float32).reshape(27 * 19)
        out = np.zeros(40, np.float32)
        self._digit_models(_p(patch), _p(out))
        return out[:10].copy(), out[10:].reshape(3, 10).copy()

    # ---- frame scoring (dmz_focus_score / dmz_brightness_score) ---------------------------------------------
    def focus_score(self, y, use_full_image=False):
        y = np.ascontiguousarray(y, np.uint8)
        return float(self._focus_score(_p(y), y.shape[1], y.shape[1], y.shape[0], int(use_full_image)))

    def brightness_score(self, y, use_full_image=False):
        y = np.ascontiguousarray(y, np.uint8)
        return float(self._brightness_score(_p(y), y.shape[1], y.shape[1], y.shape[0], int(use_full_image)))

    def scoring_rect(self, w, h, use_full_image=False):
        out = np.zeros(4, np.int32)
        self._scoring_rect(w, h, int(use_full_image), _p(out))
        return out

    # ---- pixel formats either side of the path (dmz_YCbCr_to_RGB, dmz_deinterleave_RGBA_to_R, Cython stencils) ----
    def ycbcr_to_rgb(self, y, cb, cr, channels=3):
        y, cb, cr = (np.ascontiguousarray(a, np.uint8) for a in (y, cb, cr))
        h, w = y.shape
        out = np.zeros((h, w, channels), np.uint8)
        self._ycbcr_to_rgb(_p(y), w, _p(cb), _p(cr), w, w, h, channels, _p(out), w * channels)
        return out

    def rgba_to_r(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.uint8).reshape(-1)
        out = np.zeros(rgba.size // 4, np.uint8)
        self._rgba_to_r(_p(rgba), _p(out), out.size)
        return out

    def stencil3(self, img, kind):
        """kind 0 / 1 / 2 = llcv_scharr3_dx_abs / llcv_scharr3_dy_abs / llcv_sobel3_dx_dy on a whole u8 image -> int16."""
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(img.shape, np.int16)
        self._stencil3(_p(img), img.shape[1], img.shape[1], img.shape[0], kind, _p(out))
        return out

    # ---- E0: expiry digit (port only; the reference build here has SCAN_EXPIRY off) -----------------------
    def expiry_patch_prep(self, img16x11):
        img = np.ascontiguousarray(img16x11, np.uint8)
        assert img.shape == (16, 11)
        out = np.zeros((16, 11), np.float32)
        if self.kind == "refx":  # prepare_image_for_cat with the patch as the whole image, rect at (0, 0)
            self.lib.ref_expiry_patch_prep(_p(img), 11, 11, 16, 0, 0, _p(out))
        else:
            self.lib.orc_expiry_patch_prep(_p(img), 11, _p(out))
        return out

    def expiry_digit_model(self, x, taps=False):
        x = np.ascontiguousarray(x, np.float32).reshape(176)
        out = np.zeros(10, np.float32)
        if self.kind == "refx":
            self.lib.ref_expiry_digit_model(_p(x), _p(out))
            return out
        l1, l2, hid = np.zeros(3500, np.float32), np.zeros(120, np.float32), np.zeros(176, np.float32)
        self.lib.orc_expiry_digit_model(_p(x), _p(out), _p(l1), _p(l2), _p(hid))
        return (out, l1, l2, hid) if taps else out

    def best_expiry_seg(self, card, y_offset):
        """refx only: best_expiry_seg (scan/expiry_seg.cpp:706-903).  Returns an (n_groups, 17) int32 array:
        top, left, width, height, character_width, pattern, n_rects, then (rect.top, rect.left) x 5."""
        card = np.ascontiguousarray(card, np.uint8)
        out = np.zeros(8192, np.int32)
        n = C.c_int(0)
        k = self.lib.ref_best_expiry_seg(_p(card), int(y_offset), _p(out), out.size, C.byref(n))
        assert k >= 0
        return out[:n.value].reshape(k, 17).copy()

    # ---- refx only: the session with its expiry branch live (scan/scan.cpp:41-86, expiry_categorize.cpp:448-497)
    def scanner_add_frame_with_expiry(self, s, card, scan_expiry=True):
        card = np.ascontiguousarray(card, np.uint8)
        out, n = Scan(), C.c_int32(0)
        self.lib.ref_scanner_add_frame_with_expiry(s, _p(card), int(scan_expiry), C.byref(out), C.byref(n))
        return out, n.value

    def scanner_expiry_peek(self, s, cap=64):
        m, y = C.c_int32(0), C.c_int32(0)
        meta, scores = np.zeros((cap, 4), np.int32), np.zeros((cap, 4, 10), np.float32)
        n = self.lib.ref_scanner_expiry_peek(s, C.byref(m), C.byref(y), _p(meta), _p(scores), cap)
        return (m.value, y.value), meta[:n].copy(), scores[:n].copy()

    def scanner_result_expiry(self, s):
        d, n, m, y = np.zeros(16, np.uint8), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        done = self.lib.ref_scanner_result_expiry(s, _p(d), C.byref(n), C.byref(m), C.byref(y))
        return bool(done), d[:n.value].copy(), m.value, y.value

    def expiry_month_year(self, scores5x10, month=0, year=0):
        sc = np.ascontiguousarray(scores5x10, np.float32).reshape(50)
        m, y = C.c_int32(month), C.c_int32(year)
        self.lib.ref_expiry_month_year(_p(sc), C.byref(m), C.byref(y))
        return m.value, y.value

    def scharr3_dx_abs(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(img.shape, np.int16)
        self.lib.ref_scharr3_dx_abs(_p(img), img.shape[1], img.shape[1], img.shape[0], _p(out))
        return out

    def scan_card_image(self, card):
        card = np.ascontiguousarray(card, np.uint8)
        out = Scan()
        self._scan_card_image(_p(card), C.byref(out))
        return out

    def process_frames(self, frames, orientation=3, want_cards=False):
        """frames: (n, h, w) u8. Returns a RECORD_DTYPE array (and the cards if asked)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        cb = np.full((h // 2, w // 2), 128, np.uint8)
        recs = np.zeros(n, RECORD_DTYPE)
        cards = np.zeros((n, 270, 428), np.uint8) if want_cards else None
        for k in range(n):
            self._process_frame(_p(frames[k]), w, h, w, _p(cb), _p(cb), w // 2, orientation,
                                recs[k:k + 1].ctypes.data_as(C.c_void_p),
                                _p(cards[k]) if want_cards else None)
        return (recs, cards) if want_cards else recs

    def bench_frames(self, frames, nthreads, orientation=3):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        cb = np.full((h // 2, w // 2), 128, np.uint8)
        recs = np.zeros(n, RECORD_DTYPE)
        secs = self._bench_frames(_p(frames), n, w, h, _p(cb), _p(cb), orientation, int(nthreads), _p(recs))
        return secs, recs

    def bench_stages(self, frames, orientation=3):
        """One thread: seconds spent in detect / transform / scan over the frames, and how many frames reached each."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        cb = np.full((h // 2, w // 2), 128, np.uint8)
        secs, counts = np.zeros(3, np.float64), np.zeros(3, np.int32)
        self._bench_stages(_p(frames), n, w, h, _p(cb), _p(cb), orientation, _p(secs), _p(counts))
        return secs, counts

    def bench_detect(self, frames, nthreads, orientation=3):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        cb = np.full((h // 2, w // 2), 128, np.uint8)
        found = C.c_int32(0)
        secs = self._bench_detect(_p(frames), n, w, h, _p(cb), _p(cb), orientation, int(nthreads), C.byref(found))
        return secs, found.value

    def bench_patches(self, patches, nthreads):
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, 27 * 19)
        out = np.zeros((patches.shape[0], 40), np.float32)
        secs = self._bench_patches(_p(patches), patches.shape[0], int(nthreads), _p(out))
        return secs, out

    # ---- scanner session --------------------------------------------------------------------
    def scanner_new(self):
        return self._scanner_new()

    def scanner_free(self, s):
        self._scanner_free(s)

    def scanner_add_frame(self, s, card):
        card = np.ascontiguousarray(card, np.uint8)
        out = Scan()
        self._scanner_add_frame(s, _p(card), C.byref(out))
        return out

    def scanner_peek(self, s):
        a15 = np.zeros((16, 10), np.float32)
        a16 = np.zeros((16, 10), np.float32)
        cnt = np.zeros(2, np.int32)
        self._scanner_peek(s, _p(a15), _p(a16), _p(cnt))
        return a15, a16, cnt

    def scanner_result(self, s):
        digits = np.zeros(16, np.uint8)
        n = C.c_int32()
        complete = self._scanner_result(s, _p(digits), C.byref(n))
        return bool(complete), digits[: n.value].copy()

    def luhn(self, digits):
        d = np.ascontiguousarray(digits, np.uint8)
        return bool(self._luhn(_p(d), len(d)))

    def card_type(self, digits):
        d = np.ascontiguousarray(digits, np.uint8)
        return int(self._card_type(_p(d), len(d)))
