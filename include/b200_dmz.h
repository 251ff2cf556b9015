/*
 * include/b200_dmz.h -- C ABI of the B200-native card.io-dmz hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/OpenCV types.  The reference's
 * C++ entry points (dmz.h / scan/scan.h) are re-exported with their original signatures by the thin
 * C++ layer in include/dmz_b200_compat.h + card.io-dmz_b200/csrc/dmz_compat.cpp, which forwards here
 * with a batch of one.  Throughput callers use the *_batch functions directly.
 *
 * Every image argument is a dense-or-strided 8-bit single-channel plane.  `mem` says where the
 * caller's buffers (inputs AND outputs) live: B200_MEM_HOST (copies to/from the device are done
 * inside the call on the context's stream) or B200_MEM_DEVICE (pointers are device pointers on the
 * context's device; no copies, the call returns after enqueueing and synchronising the stream).
 *
 * All functions return 0 on success, a negative B200_E* code on failure (never throw, never abort);
 * b200_last_error() gives the text.  A context is thread-compatible (one thread at a time), like the
 * reference (SURVEY 8b "Threading").
 */
#ifndef B200_DMZ_H
#define B200_DMZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MEM_HOST 0
#define B200_MEM_DEVICE 1

#define B200_OK 0
#define B200_EINVAL (-1)   /* bad argument */
#define B200_ECUDA (-2)    /* CUDA runtime error (text in b200_last_error) */
#define B200_ENOMEM (-3)
#define B200_EUNSUPPORTED (-4)

/* FrameOrientation, dmz_olm.h:16-22 */
#define B200_ORIENT_PORTRAIT 1
#define B200_ORIENT_PORTRAIT_UPSIDE_DOWN 2
#define B200_ORIENT_LANDSCAPE_RIGHT 3
#define B200_ORIENT_LANDSCAPE_LEFT 4

#define B200_CARD_W 428 /* kCreditCardTargetWidth,  dmz_constants.h:7 */
#define B200_CARD_H 270 /* kCreditCardTargetHeight, dmz_constants.h:8 */

typedef struct b200_ctx b200_ctx;

/* == dmz_found_edge / dmz_edges, dmz.h:22-37 (member order top, left, bottom, right) */
typedef struct {
  int32_t found;
  float rho, theta;
} b200_found_edge;
typedef struct {
  b200_found_edge top, left, bottom, right;
} b200_edges;

/* == dmz_corner_points, dmz_olm.h:37-42 */
typedef struct {
  float top_left[2], bottom_left[2], top_right[2], bottom_right[2];
} b200_corner_points;

/* == NVerticalSegmentation, scan/n_vseg.h:14-21 */
typedef struct {
  float score;
  uint16_t y_offset;
  uint8_t pattern_type;
  uint8_t number_pattern[19];
  uint8_t number_pattern_length;
  uint8_t number_length;
} b200_vseg;

/* == NHorizontalSegmentation, scan/n_hseg.h:13-19 */
typedef struct {
  uint8_t n_offsets;
  uint16_t offsets[16];
  float score;
  float number_width;
  uint16_t pattern_offset;
} b200_hseg;

/* What scan_card_image (scan/frame.cpp:24-81) leaves in a FrameScanResult. */
typedef struct {
  float scores[160]; /* NumberScores: 16 x 10 row-major, rows >= hseg.n_offsets are 0 */
  b200_hseg hseg;
  b200_vseg vseg;
  uint8_t usable;
  uint8_t upside_down;
  uint8_t pad[2];
} b200_scan;

/* Per-strip integer taps of best_line_for_sample (dmz.cpp:224-271); used by the parity tests. */
typedef struct {
  int32_t found, r, n, max_votes, low, high, n_edge_px;
  float rho, theta; /* ROI-local; FLT_MAX when !found */
} b200_line;

/* Per-frame record of the whole path. */
typedef struct {
  int32_t found[4]; /* top, left, bottom, right */
  float rho[4];     /* full-frame (rho, theta) per edge; unspecified where !found */
  float theta[4];
  float corners[8]; /* tl, bl, tr, br (x,y) */
  int32_t all_found;
  b200_scan scan;      /* valid iff all_found */
  uint32_t card_check; /* sum_i (i+1)*card[i] mod 2^32 over the 428x270 card; 0 if !all_found or if the cards were
                          not materialised (b200_set_card_mode) */
} b200_frame_record;

/* ---- life cycle (replaces dmz_context_create / destroy + mz_create, dmz.h:45-51, mz.h:19-25) ---- */

/* weights_dir: directory with the model blobs (card.io-dmz_b200/weights); NULL = $B200_DMZ_WEIGHTS or
 * the directory next to the shared library. */
int b200_ctx_create(b200_ctx **ctx, int device_ordinal, const char *weights_dir);
void b200_ctx_destroy(b200_ctx *ctx);
const char *b200_last_error(const b200_ctx *ctx);
/* Optional: pre-size device scratch for batches of up to max_frames frames of width x height. */
int b200_ctx_reserve(b200_ctx *ctx, int max_frames, int width, int height);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t b200_launch_count(const b200_ctx *ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
void *b200_ctx_stream(const b200_ctx *ctx);
/* Host-buffer path of b200_process_frames_batch: only the bounding rectangle of the four detection strips plus
 * `margin` pixels is copied to the device (margin < 0: whole frames; default 2 or $B200_DMZ_CROP_MARGIN).  Frames
 * whose detected card quad reaches outside that rectangle are transparently redone from a full-frame upload;
 * b200_full_frame_redos counts them.  Results never depend on the margin. */
void b200_set_crop_margin(b200_ctx *ctx, int margin);
/* b200_process_frames_batch without cards_out: by default the 428 x 270 cards are never materialised -- only the ~111
 * card rows scan_card_image reads (68 coarse vseg rows, the 43-row fine window, the 27-row number strip) are warped, and
 * b200_frame_record.card_check is 0.  always_materialise != 0 (or $B200_DMZ_CARD_MODE=1) warps every card in full, as a
 * call with cards_out does; every other field of the records is identical either way. */
void b200_set_card_mode(b200_ctx *ctx, int always_materialise);
uint64_t b200_full_frame_redos(const b200_ctx *ctx);
/* Bytes copied host->device / device->host so far by b200_process_frames_batch(B200_MEM_HOST) calls. */
void b200_transfer_bytes(const b200_ctx *ctx, uint64_t *h2d, uint64_t *d2h);
/* Per-stage device time of b200_process_frames_batch (B200_MEM_DEVICE calls), measured with CUDA events on the
 * context's stream.  Stages: 0 detect (Sobel/Canny/Hough), 1 geometry, 2 warp, 3 vseg, 4 hseg, 5 categorize,
 * 6 finalize.  b200_set_profiling(ctx, 1) zeroes the accumulators; b200_stage_times returns the accumulated
 * milliseconds and the number of frames they cover. */
void b200_set_profiling(b200_ctx *ctx, int on);
int b200_stage_times(const b200_ctx *ctx, double ms[7], uint64_t *frames);

/* ---- dmz_detect_edges (dmz.h:86-87, dmz.cpp:371-439), batched ----
 * y: n planes width x height; cb, cr: n planes (width/2) x (height/2), or both NULL (then only the Y
 * plane is searched; with real chroma the reference falls back Y -> Cb -> Cr per edge, dmz.cpp:351-367).
 * edges / corners / all_found: n entries each; lines (optional, may be NULL): 4*n Y-plane strip taps in
 * detection order top, bottom, left, right. */
int b200_detect_edges_batch(b200_ctx *ctx, const uint8_t *y, int y_row_stride, size_t y_frame_stride,
                            const uint8_t *cb, const uint8_t *cr, int c_row_stride, size_t c_frame_stride,
                            int width, int height, int n, int orientation, int mem,
                            b200_edges *edges, b200_corner_points *corners, uint8_t *all_found, b200_line *lines);

/* ---- best_line_for_sample (dmz.cpp:224-271) for the four detection strips of n Y planes: llcv_sobel7 x2 (cv/sobel.cpp:500),
 * llcv_adaptive_canny7_precomputed_sobel (cv/canny.cpp:568), llcv_hough (cv/hough.cpp:52).  lines: 4*n taps in detection
 * order top, bottom, left, right.  This is BASELINE.json configs[3]'s unit of work (Canny+Hough alone). */
int b200_detect_lines_batch(b200_ctx *ctx, const uint8_t *y, int y_row_stride, size_t y_frame_stride, int width, int height,
                            int n, int orientation, int mem, b200_line *lines);

/* ---- dmz_transform_card (dmz.h:96, dmz.cpp:443-497), batched ----
 * valid (optional): frames with valid[i] == 0 are skipped (their card is zero-filled).
 * cards: n dense 428 x 270 planes. */
int b200_transform_card_batch(b200_ctx *ctx, const uint8_t *sample, int row_stride, size_t frame_stride, int width,
                              int height, int n, const b200_corner_points *corners, const uint8_t *valid,
                              int orientation, int upsample, int mem, uint8_t *cards);

/* ---- scan_card_image (scan/frame.cpp:24; the arithmetic inside scanner_add_frame_*), batched ----
 * cards: n dense 428 x 270 planes.  scans: n entries. */
int b200_scan_cards_batch(b200_ctx *ctx, const uint8_t *cards, int n, const uint8_t *valid, int mem, b200_scan *scans);

/* ---- whole path, intermediates stay in HBM: detect -> transform -> scan ----
 * records: n entries.  cards_out (optional): n dense 428 x 270 planes. */
int b200_process_frames_batch(b200_ctx *ctx, const uint8_t *y, int y_row_stride, size_t y_frame_stride, int width,
                              int height, int n, int orientation, int mem, b200_frame_record *records,
                              uint8_t *cards_out);

/* ---- stage taps for the parity tests (device work, host-visible results; mem as above) ---- */
/* llcv_calc_persp_transform (cv/warp.cpp:34): pts = n x 8 floats (x0,y0..x3,y3) src and dst; M = n x 9. */
int b200_calc_persp_transform_batch(b200_ctx *ctx, const float *src_pts, const float *dst_pts, int n, float *M);
/* scores_for_number_image + patch prep (scan/n_categorize.cpp:45-108): patches = n x 27 x 19 u8 taken from a
 * card; out = n x 40 floats (10 ensemble scores + 3 x 10 model probabilities). */
int b200_categorize_patches_batch(b200_ctx *ctx, const uint8_t *patches, int n, int mem, float *out);
/* applyc_* ensemble on already-prepared float patches (n x 27 x 19 f32); out as above.  This is the entry the
 * reference's embedded model known-answer tests (modelc_*.cpp:2039-2051) are replayed through. */
int b200_digit_models_batch(b200_ctx *ctx, const float *patches, int n, int mem, float *out);
/* applym_befe75da on prepared rows: in = n x 204 f32, out = n x 3 f32 */
int b200_vseg_model_batch(b200_ctx *ctx, const float *rows, int n, int mem, float *out);
/* vseg_probabilities_for_hstrip (scan/n_vseg.cpp:39-47) on the coarse rows 0, 4, .., 268 of n warped cards (428 x 270 u8,
 * dense): out = n x 270 x 2 f32, (visa-like, amex-like) probability per card row, 0 for the rows not scored.  The tap of
 * the row kernel the whole path uses (row preparation + hidden layer on the tensor cores + logistic layer). */
int b200_vseg_rows_batch(b200_ctx *ctx, const uint8_t *cards, int n, int mem, float *out);

/* ---- chroma de-interleave ----
 * dmz_deinterleave_uint8_c2 (dmz.h:61, dmz.cpp:49-56): n interleaved CbCr planes (width x height pixels, two bytes per
 * pixel, row_stride >= 2 * width bytes) into two dense width x height planes each (channel1 = even bytes). */
int b200_deinterleave_c2_batch(b200_ctx *ctx, const uint8_t *interleaved, int row_stride, size_t frame_stride, int width, int height,
                               int n, int mem, uint8_t *channel1, uint8_t *channel2);

/* ---- pixel formats either side of the path (the widening after SURVEY 8f's four rows) ----
 * dmz_YCbCr_to_RGB (dmz.h:72, dmz.cpp:58-64 -> llcv_YCbCr2RGB_u8_c, cv/convert.cpp:449-504): n triples of SAME-SIZED
 * u8 planes (the SDK converts the warped card: Cb / Cr come out of dmz_transform_card with upsample = true) into
 * interleaved R, G, B bytes, or R, G, B, 255 when channels == 4 (the reference keys that on dst->nChannels).
 * rgb = n x height x width x channels, dense.  Cb and Cr share their strides. */
int b200_ycbcr_to_rgb_batch(b200_ctx *ctx, const uint8_t *y, int y_row_stride, size_t y_frame_stride, const uint8_t *cb,
                            const uint8_t *cr, int c_row_stride, size_t c_frame_stride, int width, int height, int n, int channels,
                            int mem, uint8_t *rgb);
/* dmz_deinterleave_RGBA_to_R (dmz.h:67, dmz.cpp:66-109): r[i] = rgba[4 i] for n_pixels pixels (frames of a batch are
 * just more pixels).  The reference assumes n_pixels % 4 == 0 and writes past `dest` otherwise; this one never does. */
int b200_rgba_to_r_batch(b200_ctx *ctx, const uint8_t *rgba, size_t n_pixels, int mem, uint8_t *r);
/* dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy (dmz.h:105-107, dmz.cpp:519-531 -> cv/sobel.cpp:556-900),
 * kind = B200_STENCIL_*: n u8 planes -> n x height x width int16, dense; rows and columns clamp at the plane's edge.
 * (The "abs" pair takes |difference| BEFORE the 3-10-3 smoothing, as the reference does: it is not |Scharr|.) */
#define B200_STENCIL_SCHARR_DX_ABS 0
#define B200_STENCIL_SCHARR_DY_ABS 1
#define B200_STENCIL_SOBEL_DX_DY 2
int b200_stencil3_batch(b200_ctx *ctx, const uint8_t *img, int row_stride, size_t frame_stride, int width, int height, int n,
                        int kind, int mem, int16_t *out);

/* ---- frame scoring ----
 * dmz_focus_score / dmz_brightness_score (dmz.h:77-80, dmz.cpp:114-195) for n luma planes: focus = stddev of
 * |sobel3 dx.dy| and brightness = mean, both over the reference's scoring rectangle (the centred card-sized rectangle,
 * or its central ninth when use_full_image is 0).  Either output pointer may be NULL.  Outputs live where `mem` says. */
int b200_frame_scores_batch(b200_ctx *ctx, const uint8_t *y, int y_row_stride, size_t y_frame_stride, int width, int height,
                            int n, int use_full_image, int mem, float *focus, float *brightness);

/* ---- E0: expiry digit (the SCAN_EXPIRY-only branch of scanner_add_frame_with_expiry) ----
 * prepare_image_for_cat + applyc_bf4dd6c8 (scan/expiry_categorize.cpp:37-109, models/expiry/modelc_bf4dd6c8.cpp):
 * patches = n x 16 rows x 11 cols u8 cut from the card at a character rect; out = n x 10 digit probabilities.
 * The *_models_ variant takes already prepared 16 x 11 float inputs (the reference's embedded KAT entry). */
int b200_expiry_digits_batch(b200_ctx *ctx, const uint8_t *patches, int n, int mem, float *out);
int b200_expiry_digit_models_batch(b200_ctx *ctx, const float *prepared, int n, int mem, float *out);
/* the same, cropping on the device: where = m x {card index, top, left}; crop i is the 16 x 11 window at (top, left) of
 * that 428x270 card (the character rectangles b200_best_expiry_seg_batch returns).  out = m x 10. */
int b200_expiry_digits_at_batch(b200_ctx *ctx, const uint8_t *cards, int n_cards, const int32_t *where, int m, int mem, float *out);

/* ---- expiry segmentation ----
 * best_expiry_seg (scan/expiry_seg.h:12, scan/expiry_seg.cpp:706-903; dmz_best_expiry_seg dmz.h:111 is its Cython
 * wrapper): for each warped 428x270 card and the number row's y offset (NVerticalSegmentation.y_offset), the MM/YY
 * candidates below the number -- five character rectangles each, the middle one accepted by the slash classifier.
 * One record mirrors the reference's GroupedRects fields that survive the call. */
typedef struct b200_expiry_group {
  int32_t top, left, width, height, character_width;
  int32_t pattern;   /* ExpiryPattern, always ExpiryPatternMMsYY (0) here: expiry_types.h:36-45 */
  int32_t n_rects;   /* always 5 */
  int32_t rect_top[5], rect_left[5];
} b200_expiry_group;
/* groups: n x max_groups records, n_groups: n counts (groups beyond max_groups are dropped and counted in n_dropped,
 * which may be NULL).  sobel_out (may be NULL): n x 270 x 428 int16, the |Scharr-dx| image the search ran on (rows
 * above y_offset + 24 are not written).  All pointers live where `mem` says. */
int b200_best_expiry_seg_batch(b200_ctx *ctx, const uint8_t *cards, const uint16_t *y_offsets, int n, int mem,
                               b200_expiry_group *groups, int max_groups, int32_t *n_groups, int32_t *n_dropped, int16_t *sobel_out);

/* ---- scanner session (scan/scan.h:50-72): host-side aggregation, no GPU work ---- */
typedef struct b200_scanner b200_scanner;
b200_scanner *b200_scanner_new(void);
void b200_scanner_free(b200_scanner *s);
void b200_scanner_reset(b200_scanner *s);
/* scanner_add_frame's bookkeeping given the frame's scan result (scan.cpp:41-86). */
void b200_scanner_add_scan(b200_scanner *s, const b200_scan *scan);
/* scanner_result (scan.cpp:88-194): returns complete; digits[16], *n_numbers filled when complete
 * (and, as in the reference, with the current best guess while checks are still failing). */
int b200_scanner_result(b200_scanner *s, uint8_t digits[16], int32_t *n_numbers);
/* expiry_extract's session half (scan/expiry_categorize.cpp:258-330, 334-441, 448-497): aggregate one frame's MM/YY
 * groups -- `scores` = n x 4 x 10 digit probabilities of characters 0, 1, 3, 4 of each group -- with the groups seen on
 * earlier frames (position tolerance 8 / 5 px, 0.7 decay, three-frame presence), then pick month / year from groups seen
 * at least three times.  The reference reads the wall clock here; the caller passes the date instead.  allow_past_dates
 * selects the reference's DMZ_DEBUG / CYTHON_DMZ variant that also accepts expired cards. */
void b200_scanner_add_expiry(b200_scanner *s, const b200_expiry_group *groups, const float *scores, int n, int current_year,
                             int current_month, int allow_past_dates);
void b200_scanner_expiry(const b200_scanner *s, int32_t *month, int32_t *year);
/* get_stable_expiry_month_and_year (expiry_categorize.cpp:398-441) on its own: scores = n_chars x 10 rows (n_chars = 5,
 * row 2 is the slash and ignored); month / year are in-out: a date is taken only if it is later than the one passed in. */
void b200_expiry_month_year_from_scores(const float *scores, int n_chars, int current_year, int current_month, int allow_past_dates,
                                        int32_t *month, int32_t *year);
/* aggregated groups: meta = {top, left, recently_seen, total_seen} x count, scores = count x 4 x 10; returns count */
int b200_scanner_expiry_peek(const b200_scanner *s, int32_t *meta, float *scores, int cap);
void b200_scanner_peek(const b200_scanner *s, float agg15[160], float agg16[160], int32_t counts[2]);

#ifdef __cplusplus
}
#endif
#endif
