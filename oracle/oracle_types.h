/*
 * oracle/oracle_types.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat POD records shared by the two checkers (oracle/_ref/libdmz_ref.so = the reference's own
 * sources + cvshim, and oracle/liboracle.so = the plain-C restatement) and by the ctypes bindings in
 * tests/.  The first two mirror the byte layout of the reference's own structs so a memcpy converts
 * between them.
 */
#ifndef ORACLE_TYPES_H
#define ORACLE_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == NVerticalSegmentation, scan/n_vseg.h:14-21 (28 bytes) */
typedef struct {
  float score;
  uint16_t y_offset;
  uint8_t pattern_type; /* 0 unknown, 1 visa-like (16), 2 amex-like (15) */
  uint8_t number_pattern[19];
  uint8_t number_pattern_length;
  uint8_t number_length;
} orc_vseg;

/* == NHorizontalSegmentation, scan/n_hseg.h:13-19 (48 bytes) */
typedef struct {
  uint8_t n_offsets;
  uint16_t offsets[16];
  float score;
  float number_width;
  uint16_t pattern_offset;
} orc_hseg;

/* Result of best_line_for_sample (dmz.cpp:224-271) for one strip, with the integer taps the parity
 * tests compare bit for bit. */
typedef struct {
  int32_t found;     /* !is_null */
  int32_t r, n;      /* Hough argmax cell: rho index, angle index (valid iff found) */
  int32_t max_votes; /* accumulator value at the argmax (always valid) */
  int32_t low, high; /* adaptive Canny thresholds (canny.cpp:573-579) */
  int32_t n_edge_px; /* number of 255 pixels in the Canny map */
  float rho, theta;  /* ROI-local line (FLT_MAX, FLT_MAX when not found) */
} orc_line;

/* dmz_edges (dmz.h:22-37) + dmz_corner_points (dmz_olm.h:37-42) + return value of dmz_detect_edges. */
typedef struct {
  int32_t found[4];  /* top, left, bottom, right  (dmz_edges member order) */
  float rho[4];
  float theta[4];
  float corners[8];  /* top_left.xy, bottom_left.xy, top_right.xy, bottom_right.xy */
  int32_t all_found; /* the bool dmz_detect_edges returns */
} orc_detect;

/* What scan_card_image (scan/frame.cpp:24-81) leaves in FrameScanResult, flattened. */
typedef struct {
  float scores[160]; /* NumberScores, 16x10 row-major; rows >= n_offsets are 0 */
  orc_hseg hseg;
  orc_vseg vseg;
  uint8_t usable;
  uint8_t upside_down;
  uint8_t pad[2];
} orc_scan;

/* Per-frame record of the whole path (detect -> transform -> scan_card_image). */
typedef struct {
  orc_detect detect;
  orc_scan scan;      /* valid iff detect.all_found */
  uint32_t card_check; /* sum_i (i+1)*card[i] mod 2^32 over the 428x270 card image, 0 if not detected */
} orc_frame_record;

#ifdef __cplusplus
}
#endif
#endif
