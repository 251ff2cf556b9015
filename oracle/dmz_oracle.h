/*
 * oracle/dmz_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of card.io-dmz's per-frame detect -> warp -> OCR path (SURVEY section 8a rows
 * D0..S1).  Each function cites the reference file:line it restates.  It exists so the CUDA path can
 * be checked on machines where /root/reference is absent (the GPU box).  It is itself pinned
 * against (a) the reference's embedded model known-answer vectors (tests/golden/kat_*.bin),
 * (b) golden fixtures produced by running the reference's own sources (oracle/_ref, see
 * tools/make_ref_golden.py) and (c) live, stage by stage, against oracle/_ref when that library is
 * present (tests/test_oracle_vs_ref.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * this.  The product path (card.io-dmz_b200/csrc) never does.
 */
#ifndef DMZ_ORACLE_H
#define DMZ_ORACLE_H

#include <stddef.h>
#include "oracle_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Load the generated-model weights (flat f32 blobs written by tools/extract_weights.py) from dir.
 * Returns 0 on success. Must be called before any vseg / categorize function. */
int orc_load_weights(const char *weights_dir);

void orc_detection_boxes(int w, int h, int orientation, int32_t out[16]);
void orc_sobel7(const uint8_t *img, int step, int w, int h, int16_t *dx, int16_t *dy);
void orc_adaptive_canny(const uint8_t *img, int step, int w, int h, const int16_t *dx, const int16_t *dy,
                        uint8_t *edges, int32_t *low, int32_t *high);
void orc_best_line(const uint8_t *img, int step, int w, int h, int vertical, orc_line *out);
int orc_detect_edges(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                     int orientation, orc_detect *out);
void orc_calc_persp_transform(const float src_pts[8], const float dst_pts[8], float M[9]);
void orc_transform_card(const uint8_t *y, int w, int h, int ystep, const float corners[8], int orientation,
                        uint8_t *card);
void orc_transform_card_up(const uint8_t *y, int w, int h, int ystep, const float corners[8], int orientation, int upsample,
                           uint8_t *card);
void orc_vseg_row(const uint8_t *card, int row, float probs[3]);
void orc_vseg_model(const float *in204, float probs[3]);
void orc_best_n_vseg(const uint8_t *card, orc_vseg *out);
void orc_best_n_hseg(const uint8_t *card, const orc_vseg *vseg, orc_hseg *out);
void orc_number_scores(const uint8_t *card, int y_offset, const orc_hseg *hseg, float *scores);
void orc_digit_patch_prep(const uint8_t *img, int step, float *patch);
void orc_digit_models(const float *patch, float *out40);
void orc_scan_card_image(const uint8_t *card, orc_scan *out);
/* E0 (expiry digit): patch = 16 rows x 11 cols */
void orc_expiry_patch_prep(const uint8_t *img, int step, float *out176);
void orc_expiry_digit_model(const float *in176, float *out10, float *l1_3500, float *l2_120, float *hid_176);
void orc_process_frame(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                       int orientation, orc_frame_record *rec, uint8_t *card_out);

/* scanner session (scan/scan.cpp) */
void *orc_scanner_new(void);
void orc_scanner_free(void *s);
void orc_scanner_reset(void *s);
void orc_scanner_add_frame(void *state, const uint8_t *card, orc_scan *out);
void orc_scanner_add_scan(void *state, const orc_scan *scan); /* aggregation only, from a precomputed scan */
void orc_scanner_peek(void *state, float agg15[160], float agg16[160], int32_t counts[2]);
int orc_scanner_result(void *state, uint8_t digits[16], int32_t *n_numbers);
int orc_luhn(const uint8_t *digits, int n);
int orc_card_type(const uint8_t *digits, int n);

/* ---- frame scoring: dmz_focus_score / dmz_brightness_score (dmz.cpp:114-195) ----
 * rect = {x, y, w, h} of dmz_set_roi_for_scoring for a w x h frame. */
void orc_scoring_rect(int w, int h, int use_full_image, int rect[4]);
float orc_focus_score(const uint8_t *y, int ystep, int w, int h, int use_full_image);
float orc_brightness_score(const uint8_t *y, int ystep, int w, int h, int use_full_image);

/* ---- pixel formats either side of the path: dmz_YCbCr_to_RGB (dmz.cpp:58-64, cv/convert.cpp:449-504),
 * dmz_deinterleave_RGBA_to_R (dmz.cpp:66-109), and the 3x3 stencils of the Cython layer (dmz.cpp:519-531; kind 0 / 1 / 2 =
 * llcv_scharr3_dx_abs / llcv_scharr3_dy_abs / llcv_sobel3_dx_dy, cv/sobel.cpp:556-900) ---- */
void orc_ycbcr_to_rgb(const uint8_t *y, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep, int w, int h, int channels,
                      uint8_t *dst, int dstep);
void orc_rgba_to_r(const uint8_t *source, uint8_t *dest, size_t size);
void orc_stencil3(const uint8_t *img, int step, int w, int h, int kind, int16_t *out);

uint32_t orc_card_check(const uint8_t *p, size_t n);

/* multi-threaded timing of the whole path (bench.py cpu_baseline kind "port"); returns wall seconds */
double orc_bench_frames(const uint8_t *frames, int n, int w, int h, const uint8_t *cb, const uint8_t *cr,
                        int orientation, int nthreads, orc_frame_record *recs);

#ifdef __cplusplus
}
#endif
#endif
