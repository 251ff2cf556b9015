"""ctypes mirror of include/b200_dmz.h.  Raises if the CUDA extension is missing or fails: there is no
CPU fallback on the product path."""
import ctypes as C
import os

import numpy as np

__all__ = ["Dmz", "Scanner", "B200Error", "lib_path", "Edges", "CornerPoints", "VSeg", "HSeg", "Scan", "Line",
           "FrameRecord", "RECORD_DTYPE", "SCAN_DTYPE", "LINE_DTYPE", "EXPIRY_GROUP_DTYPE", "expiry_month_year_from_scores",
           "MEM_HOST", "MEM_DEVICE", "CARD_W", "CARD_H"]

HERE = os.path.dirname(os.path.abspath(__file__))
MEM_HOST, MEM_DEVICE = 0, 1
CARD_W, CARD_H = 428, 270


class B200Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(HERE, "libb200dmz.so")


class FoundEdge(C.Structure):
    _fields_ = [("found", C.c_int32), ("rho", C.c_float), ("theta", C.c_float)]


class Edges(C.Structure):
    _fields_ = [("top", FoundEdge), ("left", FoundEdge), ("bottom", FoundEdge), ("right", FoundEdge)]


class CornerPoints(C.Structure):
    _fields_ = [("top_left", C.c_float * 2), ("bottom_left", C.c_float * 2), ("top_right", C.c_float * 2),
                ("bottom_right", C.c_float * 2)]


class VSeg(C.Structure):
    _fields_ = [("score", C.c_float), ("y_offset", C.c_uint16), ("pattern_type", C.c_uint8),
                ("number_pattern", C.c_uint8 * 19), ("number_pattern_length", C.c_uint8), ("number_length", C.c_uint8)]


class HSeg(C.Structure):
    _fields_ = [("n_offsets", C.c_uint8), ("offsets", C.c_uint16 * 16), ("score", C.c_float),
                ("number_width", C.c_float), ("pattern_offset", C.c_uint16)]


class Scan(C.Structure):
    _fields_ = [("scores", C.c_float * 160), ("hseg", HSeg), ("vseg", VSeg), ("usable", C.c_uint8),
                ("upside_down", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Line(C.Structure):
    _fields_ = [("found", C.c_int32), ("r", C.c_int32), ("n", C.c_int32), ("max_votes", C.c_int32), ("low", C.c_int32),
                ("high", C.c_int32), ("n_edge_px", C.c_int32), ("rho", C.c_float), ("theta", C.c_float)]


class FrameRecord(C.Structure):
    _fields_ = [("found", C.c_int32 * 4), ("rho", C.c_float * 4), ("theta", C.c_float * 4), ("corners", C.c_float * 8),
                ("all_found", C.c_int32), ("scan", Scan), ("card_check", C.c_uint32)]


_SCAN_FIELDS = [
    ("scores", "<f4", 160),
    ("h_n_offsets", "u1"), ("_p0", "u1"), ("h_offsets", "<u2", 16), ("_p1", "u1", 2), ("h_score", "<f4"),
    ("h_number_width", "<f4"), ("h_pattern_offset", "<u2"), ("_p2", "u1", 2),
    ("v_score", "<f4"), ("v_y_offset", "<u2"), ("v_pattern_type", "u1"), ("v_number_pattern", "u1", 19),
    ("v_number_pattern_length", "u1"), ("v_number_length", "u1"),
    ("usable", "u1"), ("upside_down", "u1"), ("_p4", "u1", 2),
]
SCAN_DTYPE = np.dtype(_SCAN_FIELDS)
RECORD_DTYPE = np.dtype([("found", "<i4", 4), ("rho", "<f4", 4), ("theta", "<f4", 4), ("corners", "<f4", 8),
                         ("all_found", "<i4")] + _SCAN_FIELDS + [("card_check", "<u4")])
LINE_DTYPE = np.dtype([("found", "<i4"), ("r", "<i4"), ("n", "<i4"), ("max_votes", "<i4"), ("low", "<i4"),
                       ("high", "<i4"), ("n_edge_px", "<i4"), ("rho", "<f4"), ("theta", "<f4")])
assert SCAN_DTYPE.itemsize == C.sizeof(Scan) == 720
assert RECORD_DTYPE.itemsize == C.sizeof(FrameRecord) == 808
assert LINE_DTYPE.itemsize == C.sizeof(Line) == 36

_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise B200Error("CUDA extension %s is not built (run __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.b200_ctx_create.argtypes = [C.POINTER(vp), i, C.c_char_p]
    lib.b200_ctx_destroy.argtypes = [vp]
    lib.b200_last_error.argtypes = [vp]
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_ctx_reserve.argtypes = [vp, i, i, i]
    lib.b200_launch_count.argtypes = [vp]
    lib.b200_launch_count.restype = C.c_uint64
    lib.b200_ctx_stream.argtypes = [vp]
    lib.b200_ctx_stream.restype = vp
    lib.b200_detect_edges_batch.argtypes = [vp, vp, i, sz, vp, vp, i, sz, i, i, i, i, i, vp, vp, vp, vp]
    lib.b200_detect_lines_batch.argtypes = [vp, vp, i, sz, i, i, i, i, i, vp]
    lib.b200_transform_card_batch.argtypes = [vp, vp, i, sz, i, i, i, vp, vp, i, i, i, vp]
    lib.b200_scan_cards_batch.argtypes = [vp, vp, i, vp, i, vp]
    lib.b200_process_frames_batch.argtypes = [vp, vp, i, sz, i, i, i, i, i, vp, vp]
    lib.b200_calc_persp_transform_batch.argtypes = [vp, vp, vp, i, vp]
    lib.b200_categorize_patches_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_vseg_model_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_vseg_rows_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_digit_models_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_best_expiry_seg_batch.argtypes = [vp, vp, vp, i, i, vp, i, vp, vp, vp]
    lib.b200_deinterleave_c2_batch.argtypes = [vp, vp, i, C.c_size_t, i, i, i, i, vp, vp]
    lib.b200_frame_scores_batch.argtypes = [vp, vp, i, C.c_size_t, i, i, i, i, i, vp, vp]
    lib.b200_ycbcr_to_rgb_batch.argtypes = [vp, vp, i, C.c_size_t, vp, vp, i, C.c_size_t, i, i, i, i, i, vp]
    lib.b200_rgba_to_r_batch.argtypes = [vp, vp, C.c_size_t, i, vp]
    lib.b200_stencil3_batch.argtypes = [vp, vp, i, C.c_size_t, i, i, i, i, i, vp]
    lib.b200_expiry_digits_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_expiry_digits_at_batch.argtypes = [vp, vp, i, vp, i, i, vp]
    lib.b200_scanner_add_expiry.argtypes = [vp, vp, vp, i, i, i, i]
    lib.b200_scanner_expiry.argtypes = [vp, vp, vp]
    lib.b200_expiry_month_year_from_scores.argtypes = [vp, i, i, i, i, vp, vp]
    lib.b200_scanner_expiry_peek.argtypes = [vp, vp, vp, i]
    lib.b200_expiry_digit_models_batch.argtypes = [vp, vp, i, i, vp]
    lib.b200_set_profiling.argtypes = [vp, i]
    lib.b200_set_crop_margin.argtypes = [vp, i]
    lib.b200_set_card_mode.argtypes = [vp, i]
    lib.b200_full_frame_redos.argtypes = [vp]
    lib.b200_full_frame_redos.restype = C.c_uint64
    lib.b200_transfer_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.b200_stage_times.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    lib.b200_scanner_new.restype = vp
    lib.b200_scanner_free.argtypes = [vp]
    lib.b200_scanner_reset.argtypes = [vp]
    lib.b200_scanner_add_scan.argtypes = [vp, vp]
    lib.b200_scanner_result.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    lib.b200_scanner_peek.argtypes = [vp, vp, vp, vp]
    _lib = lib
    return lib


EXPIRY_GROUP_DTYPE = np.dtype([("top", "<i4"), ("left", "<i4"), ("width", "<i4"), ("height", "<i4"), ("character_width", "<i4"),
                               ("pattern", "<i4"), ("n_rects", "<i4"), ("rect_top", "<i4", 5), ("rect_left", "<i4", 5)])


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Dmz:
    """One b200_ctx (device, stream, weights, scratch).  NumPy arrays = host buffers (B200_MEM_HOST);
    integer addresses = device pointers (B200_MEM_DEVICE)."""

    def __init__(self, device=0, weights_dir=None, materialise_cards=True):
        """materialise_cards: b200_set_card_mode -- True (the default of this mirror, so that records always carry the
        card checksum the parity tests compare) warps every card in full; False is the library's own default: calls that do
        not ask for the cards warp only the rows scan_card_image reads and leave card_check 0."""
        self.lib = _load()
        self.ctx = C.c_void_p()
        rc = self.lib.b200_ctx_create(C.byref(self.ctx), device, weights_dir.encode() if weights_dir else None)
        if rc != 0:
            msg = self.lib.b200_last_error(self.ctx).decode() if self.ctx else "context allocation failed"
            if self.ctx:
                self.lib.b200_ctx_destroy(self.ctx)
                self.ctx = None
            raise B200Error("b200_ctx_create failed (%d): %s" % (rc, msg))
        self.set_card_mode(materialise_cards)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.b200_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B200Error("b200 call failed (%d): %s" % (rc, self.lib.b200_last_error(self.ctx).decode()))

    @property
    def launches(self):
        return int(self.lib.b200_launch_count(self.ctx))

    @property
    def stream(self):
        return self.lib.b200_ctx_stream(self.ctx)

    def reserve(self, n, w, h):
        self._check(self.lib.b200_ctx_reserve(self.ctx, n, w, h))

    # ---- host-buffer API (numpy) -----------------------------------------------------------------
    def detect_edges(self, y, cb=None, cr=None, orientation=3, want_lines=False):
        """y: (n,h,w) u8.  Returns (edges[n] as structured array view, corners (n,8), all_found (n,), lines (n,4))."""
        y = np.ascontiguousarray(y, np.uint8)
        n, h, w = y.shape
        if cb is not None:
            cb = np.ascontiguousarray(cb, np.uint8)
            cr = np.ascontiguousarray(cr, np.uint8)
        edges = np.zeros((n, 4, 3), np.float32)  # placeholder raw view: (found:int32, rho, theta) per edge
        edges_raw = np.zeros(n * 12, np.int32)
        corners = np.zeros((n, 8), np.float32)
        found = np.zeros(n, np.uint8)
        lines = np.zeros((n, 4), LINE_DTYPE) if want_lines else None
        self._check(self.lib.b200_detect_edges_batch(self.ctx, _ptr(y), w, w * h, _ptr(cb), _ptr(cr), w // 2,
                                                     (w // 2) * (h // 2), w, h, n, orientation, MEM_HOST,
                                                     _ptr(edges_raw), _ptr(corners), _ptr(found), _ptr(lines)))
        er = edges_raw.reshape(n, 4, 3)
        out = {"found": er[:, :, 0].copy(), "rho": er[:, :, 1].copy().view(np.float32), "theta": er[:, :, 2].copy().view(np.float32)}
        return out, corners, found, lines

    def transform_card(self, frames, corners, valid=None, orientation=3, upsample=False):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        corners = np.ascontiguousarray(corners, np.float32).reshape(n, 8)
        if valid is not None:
            valid = np.ascontiguousarray(valid, np.uint8)
        cards = np.zeros((n, CARD_H, CARD_W), np.uint8)
        self._check(self.lib.b200_transform_card_batch(self.ctx, _ptr(frames), w, w * h, w, h, n, _ptr(corners), _ptr(valid),
                                                       orientation, int(upsample), MEM_HOST, _ptr(cards)))
        return cards

    def scan_cards(self, cards, valid=None):
        cards = np.ascontiguousarray(cards, np.uint8)
        n = cards.shape[0]
        assert cards.shape[1:] == (CARD_H, CARD_W)
        if valid is not None:
            valid = np.ascontiguousarray(valid, np.uint8)
        scans = np.zeros(n, SCAN_DTYPE)
        self._check(self.lib.b200_scan_cards_batch(self.ctx, _ptr(cards), n, _ptr(valid), MEM_HOST, _ptr(scans)))
        return scans

    def process_frames(self, frames, orientation=3, want_cards=False):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        recs = np.zeros(n, RECORD_DTYPE)
        cards = np.zeros((n, CARD_H, CARD_W), np.uint8) if want_cards else None
        self._check(self.lib.b200_process_frames_batch(self.ctx, _ptr(frames), w, w * h, w, h, n, orientation, MEM_HOST,
                                                       _ptr(recs), _ptr(cards)))
        return (recs, cards) if want_cards else recs

    def calc_persp_transform(self, src, dst):
        src = np.ascontiguousarray(src, np.float32).reshape(-1, 8)
        dst = np.ascontiguousarray(dst, np.float32).reshape(-1, 8)
        n = src.shape[0]
        m = np.zeros((n, 9), np.float32)
        self._check(self.lib.b200_calc_persp_transform_batch(self.ctx, _ptr(src), _ptr(dst), n, _ptr(m)))
        return m.reshape(n, 3, 3)

    def categorize_patches(self, patches):
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, 27, 19)
        n = patches.shape[0]
        out = np.zeros((n, 40), np.float32)
        self._check(self.lib.b200_categorize_patches_batch(self.ctx, _ptr(patches), n, MEM_HOST, _ptr(out)))
        return out[:, :10].copy(), out[:, 10:].reshape(n, 3, 10).copy()

    def digit_models(self, patches):
        """patches: (n, 27, 19) float32, already prepared.  Returns (ensemble (n,10), per-model probs (n,3,10))."""
        patches = np.ascontiguousarray(patches, np.float32).reshape(-1, 27 * 19)
        n = patches.shape[0]
        out = np.zeros((n, 40), np.float32)
        self._check(self.lib.b200_digit_models_batch(self.ctx, _ptr(patches), n, MEM_HOST, _ptr(out)))
        return out[:, :10].copy(), out[:, 10:].reshape(n, 3, 10).copy()

    def best_expiry_seg(self, cards, y_offsets, max_groups=16, want_sobel=False):
        """cards: (n, 270, 428) u8; y_offsets: n.  Returns (groups[n, max_groups] EXPIRY_GROUP_DTYPE, counts[n], dropped[n]) and
        the |Scharr| planes if asked (best_expiry_seg, scan/expiry_seg.cpp:706-903)."""
        cards = np.ascontiguousarray(cards, np.uint8)
        n = cards.shape[0]
        yo = np.ascontiguousarray(y_offsets, np.uint16)
        groups = np.zeros((n, max_groups), EXPIRY_GROUP_DTYPE)
        counts, dropped = np.zeros(n, np.int32), np.zeros(n, np.int32)
        sob = np.zeros((n, 270, 428), np.int16) if want_sobel else None
        self._check(self.lib.b200_best_expiry_seg_batch(self.ctx, _ptr(cards), _ptr(yo), n, MEM_HOST, _ptr(groups), max_groups,
                                                        _ptr(counts), _ptr(dropped), _ptr(sob) if want_sobel else None))
        return (groups, counts, dropped, sob) if want_sobel else (groups, counts, dropped)

    def deinterleave_c2(self, planes):
        """planes: (n, h, w, 2) u8 interleaved CbCr.  Returns (channel1, channel2), each (n, h, w) (dmz_deinterleave_uint8_c2)."""
        planes = np.ascontiguousarray(planes, np.uint8)
        n, h, w, _ = planes.shape
        c1, c2 = np.zeros((n, h, w), np.uint8), np.zeros((n, h, w), np.uint8)
        self._check(self.lib.b200_deinterleave_c2_batch(self.ctx, _ptr(planes), 2 * w, 2 * w * h, w, h, n, MEM_HOST, _ptr(c1), _ptr(c2)))
        return c1, c2

    def ycbcr_to_rgb(self, y, cb, cr, channels=3):
        """y, cb, cr: (n, h, w) u8 planes of the same size.  Returns (n, h, w, channels) R, G, B (, 255) (dmz_YCbCr_to_RGB)."""
        y, cb, cr = (np.ascontiguousarray(a, np.uint8) for a in (y, cb, cr))
        n, h, w = y.shape
        assert cb.shape == y.shape and cr.shape == y.shape
        out = np.zeros((n, h, w, channels), np.uint8)
        self._check(self.lib.b200_ycbcr_to_rgb_batch(self.ctx, _ptr(y), w, w * h, _ptr(cb), _ptr(cr), w, w * h, w, h, n, channels,
                                                     MEM_HOST, _ptr(out)))
        return out

    def rgba_to_r(self, rgba):
        """rgba: u8 array of 4 * n_pixels bytes (any shape).  Returns the n_pixels R bytes (dmz_deinterleave_RGBA_to_R)."""
        if not (rgba.dtype == np.uint8 and rgba.ndim == 1 and rgba.strides[0] == 1):  # keep a caller's (mis)alignment
            rgba = np.ascontiguousarray(rgba, np.uint8).reshape(-1)
        n = rgba.size // 4
        out = np.zeros(n, np.uint8)
        self._check(self.lib.b200_rgba_to_r_batch(self.ctx, _ptr(rgba), n, MEM_HOST, _ptr(out)))
        return out

    def stencil3(self, planes, kind):
        """planes: (n, h, w) u8.  kind 0 / 1 / 2 = dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy.  Returns int16 (n, h, w)."""
        planes = np.ascontiguousarray(planes, np.uint8)
        n, h, w = planes.shape
        out = np.zeros((n, h, w), np.int16)
        self._check(self.lib.b200_stencil3_batch(self.ctx, _ptr(planes), w, w * h, w, h, n, int(kind), MEM_HOST, _ptr(out)))
        return out

    def frame_scores(self, frames, use_full_image=False):
        """frames: (n, h, w) u8 luma.  Returns (focus, brightness) float32 arrays (dmz_focus_score / dmz_brightness_score)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        focus, bright = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._check(self.lib.b200_frame_scores_batch(self.ctx, _ptr(frames), w, w * h, w, h, n, int(use_full_image), MEM_HOST,
                                                     _ptr(focus), _ptr(bright)))
        return focus, bright

    def expiry_digits(self, patches):
        """patches: (n, 16, 11) u8 character crops.  Returns (n, 10) digit probabilities (E0)."""
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, 176)
        out = np.zeros((patches.shape[0], 10), np.float32)
        self._check(self.lib.b200_expiry_digits_batch(self.ctx, _ptr(patches), patches.shape[0], MEM_HOST, _ptr(out)))
        return out

    def expiry_digits_at(self, cards, where):
        """cards: (n, 270, 428) u8; where: (m, 3) int32 rows of (card index, top, left).  Returns (m, 10) probabilities."""
        cards = np.ascontiguousarray(cards, np.uint8)
        where = np.ascontiguousarray(where, np.int32).reshape(-1, 3)
        out = np.zeros((where.shape[0], 10), np.float32)
        self._check(self.lib.b200_expiry_digits_at_batch(self.ctx, _ptr(cards), cards.shape[0], _ptr(where), where.shape[0], MEM_HOST, _ptr(out)))
        return out

    def expiry_digit_models(self, prepared):
        prepared = np.ascontiguousarray(prepared, np.float32).reshape(-1, 176)
        out = np.zeros((prepared.shape[0], 10), np.float32)
        self._check(self.lib.b200_expiry_digit_models_batch(self.ctx, _ptr(prepared), prepared.shape[0], MEM_HOST, _ptr(out)))
        return out

    STAGES = ("detect", "geometry", "warp", "vseg", "hseg", "categorize", "finalize")

    def set_card_mode(self, always_materialise):
        self.lib.b200_set_card_mode(self.ctx, int(bool(always_materialise)))

    def set_crop_margin(self, margin):
        self.lib.b200_set_crop_margin(self.ctx, int(margin))

    @property
    def full_frame_redos(self):
        return int(self.lib.b200_full_frame_redos(self.ctx))

    def transfer_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.b200_transfer_bytes(self.ctx, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def set_profiling(self, on=True):
        self.lib.b200_set_profiling(self.ctx, int(on))

    def stage_times(self):
        """{stage: accumulated ms}, frames covered (device-pointer process_frames calls while profiling)."""
        ms = np.zeros(7, np.float64)
        frames = C.c_uint64()
        self._check(self.lib.b200_stage_times(self.ctx, _ptr(ms), C.byref(frames)))
        return dict(zip(self.STAGES, ms.tolist())), int(frames.value)

    def vseg_model(self, rows):
        rows = np.ascontiguousarray(rows, np.float32).reshape(-1, 204)
        n = rows.shape[0]
        out = np.zeros((n, 3), np.float32)
        self._check(self.lib.b200_vseg_model_batch(self.ctx, _ptr(rows), n, MEM_HOST, _ptr(out)))
        return out

    def vseg_rows(self, cards):
        """(visa-like, amex-like) probability of the coarse rows 0, 4, .., 268 of each card: (n, 270, 2), 0 elsewhere."""
        cards = np.ascontiguousarray(cards, np.uint8).reshape(-1, 270, 428)
        n = cards.shape[0]
        out = np.zeros((n, 270, 2), np.float32)
        self._check(self.lib.b200_vseg_rows_batch(self.ctx, _ptr(cards), n, MEM_HOST, _ptr(out)))
        return out

    # ---- device-pointer API (bench: inputs resident in HBM) ----------------------------------------
    def process_frames_device(self, d_frames, n, w, h, d_records, d_cards=None, orientation=3, row_stride=None, frame_stride=None):
        self._check(self.lib.b200_process_frames_batch(self.ctx, C.c_void_p(d_frames), row_stride or w, frame_stride or w * h,
                                                       w, h, n, orientation, MEM_DEVICE, C.c_void_p(d_records),
                                                       C.c_void_p(d_cards) if d_cards else None))

    def detect_lines_device(self, d_frames, n, w, h, d_lines, orientation=3):
        """D1-D4 only on device-resident planes; d_lines: 4 * n LINE_DTYPE entries in device memory."""
        self._check(self.lib.b200_detect_lines_batch(self.ctx, C.c_void_p(d_frames), w, w * h, w, h, n, orientation, MEM_DEVICE,
                                                     C.c_void_p(d_lines)))

    def detect_lines(self, y, orientation=3):
        y = np.ascontiguousarray(y, np.uint8)
        n, h, w = y.shape
        lines = np.zeros((n, 4), LINE_DTYPE)
        self._check(self.lib.b200_detect_lines_batch(self.ctx, _ptr(y), w, w * h, w, h, n, orientation, MEM_HOST, _ptr(lines)))
        return lines

    def categorize_patches_device(self, d_patches, n, d_out):
        """n_categorize only on device-resident 27x19 u8 patches; d_out: n x 40 floats in device memory."""
        self._check(self.lib.b200_categorize_patches_batch(self.ctx, C.c_void_p(d_patches), n, MEM_DEVICE, C.c_void_p(d_out)))

    def process_frames_host_ptr(self, h_frames, n, w, h, h_records, orientation=3):
        """Host pointers given as integers (e.g. pinned torch tensors): the e2e path, copies inside the call."""
        self._check(self.lib.b200_process_frames_batch(self.ctx, C.c_void_p(h_frames), w, w * h, w, h, n, orientation,
                                                       MEM_HOST, C.c_void_p(h_records), None))


def expiry_month_year_from_scores(scores5x10, current_year, current_month, allow_past_dates=False, month=0, year=0):
    """get_stable_expiry_month_and_year (expiry_categorize.cpp:398-441); returns (month, year)."""
    lib = _load()
    sc = np.ascontiguousarray(scores5x10, np.float32).reshape(5, 10)
    m, y = C.c_int32(month), C.c_int32(year)
    lib.b200_expiry_month_year_from_scores(_ptr(sc), 5, int(current_year), int(current_month), int(allow_past_dates), C.byref(m), C.byref(y))
    return m.value, y.value


class Scanner:
    """scanner_* session (scan/scan.h:50-72) over b200_scan records."""

    def __init__(self):
        self.lib = _load()
        self.s = self.lib.b200_scanner_new()

    def close(self):
        if self.s:
            self.lib.b200_scanner_free(self.s)
            self.s = None

    def reset(self):
        self.lib.b200_scanner_reset(self.s)

    def add_scan(self, scan_record):
        """scan_record: one element of a SCAN_DTYPE array, or the scan part of a RECORD_DTYPE element."""
        buf = np.zeros(1, SCAN_DTYPE)
        for name in SCAN_DTYPE.names:
            buf[0][name] = scan_record[name]
        self.lib.b200_scanner_add_scan(self.s, _ptr(buf))

    def peek(self):
        a15 = np.zeros((16, 10), np.float32)
        a16 = np.zeros((16, 10), np.float32)
        cnt = np.zeros(2, np.int32)
        self.lib.b200_scanner_peek(self.s, _ptr(a15), _ptr(a16), _ptr(cnt))
        return a15, a16, cnt

    def add_expiry(self, groups, scores, current_year, current_month, allow_past_dates=False):
        """groups: EXPIRY_GROUP_DTYPE array of one frame; scores: (len(groups), 4, 10) digit probabilities."""
        groups = np.ascontiguousarray(groups, EXPIRY_GROUP_DTYPE)
        scores = np.ascontiguousarray(scores, np.float32)
        self.lib.b200_scanner_add_expiry(self.s, _ptr(groups), _ptr(scores), len(groups), int(current_year), int(current_month),
                                         int(allow_past_dates))

    def expiry(self):
        m, y = C.c_int32(), C.c_int32()
        self.lib.b200_scanner_expiry(self.s, C.byref(m), C.byref(y))
        return m.value, y.value

    def expiry_peek(self, cap=64):
        meta, scores = np.zeros((cap, 4), np.int32), np.zeros((cap, 4, 10), np.float32)
        n = self.lib.b200_scanner_expiry_peek(self.s, _ptr(meta), _ptr(scores), cap)
        return meta[:n].copy(), scores[:n].copy()

    def result(self):
        digits = np.zeros(16, np.uint8)
        n = C.c_int32()
        complete = self.lib.b200_scanner_result(self.s, _ptr(digits), C.byref(n))
        return bool(complete), digits[: n.value].copy()
