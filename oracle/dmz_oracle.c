/*
 * oracle/dmz_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See dmz_oracle.h.
 *
 * CPU restatement (plain C, scalar) of card.io-dmz's detect -> warp -> OCR hot path.  Citations are
 * relative to /root/reference.  Where the reference evaluates float reductions through Eigen 3.2.4's
 * SSE2 kernels the evaluation ORDER is restated as well (helpers eig_*), because integer results
 * downstream (warp pixels, hseg offsets, gating decisions) depend on the exact bits.
 *
 * Must be compiled for x86-64 baseline with -ffp-contract=off (oracle/Makefile).
 */
#include "dmz_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "prims.h"

#define CV_PI 3.1415926535897932384626433832795

/* ------------------------------------------------------------------------------------------------
 * constants (dmz_constants.h:7-27, dmz.cpp:199-208, scan/frame.cpp:20-22, scan/scan.cpp:16-17)
 * ------------------------------------------------------------------------------------------------ */
enum { kCardW = 428, kCardH = 270, kNumberW = 19, kNumberH = 27 };
#define kPortraitVerticalPercentInset ((float)((640 - 270) / 2) / (float)640)
#define kPortraitHorizontalPercentInset ((float)((480 - 428) / 2) / (float)480)
#define kLandscapeVerticalPercentInset ((float)((480 - 270) / 2) / (float)480)
#define kLandscapeHorizontalPercentInset ((float)((640 - 428) / 2) / (float)640)
#define kVerticalPercentSlop 0.03f
#define kHorizontalPercentSlop 0.03f
#define kHorizontalAngle ((float)(CV_PI / 2.0f))
#define kVerticalAngle ((float)CV_PI)
#define kMaxAngleDeviationAllowed ((float)(5.0f * (CV_PI / 180.0f)))
#define kHoughGradientAngleThreshold 10
#define kHoughThresholdLengthDivisor 6

/* ------------------------------------------------------------------------------------------------
 * Eigen 3.2.4 evaluation-order helpers
 * ------------------------------------------------------------------------------------------------ */

/* Redux.h:192-246, LinearVectorizedTraversal / NoUnrolling with alignedStart == 0 (any expression
 * without DirectAccessBit, or an aligned plain object), Packet4f, SSE2 predux = (a0+a2)+(a1+a3)
 * (arch/SSE/PacketMath.h, no SSE3). v[] are the already-evaluated expression coefficients. */
static float eig_redux_sum(const float *v, int n) {
  int aligned2 = (n / 8) * 8, aligned = (n / 4) * 4, i, l;
  float res;
  if (n == 0) return 0.0f;
  if (aligned) {
    float p0[4], p1[4];
    for (l = 0; l < 4; l++) p0[l] = v[l];
    if (aligned > 4) {
      for (l = 0; l < 4; l++) p1[l] = v[4 + l];
      for (i = 8; i < aligned2; i += 8)
        for (l = 0; l < 4; l++) {
          p0[l] = p0[l] + v[i + l];
          p1[l] = p1[l] + v[i + 4 + l];
        }
      for (l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
      if (aligned > aligned2)
        for (l = 0; l < 4; l++) p0[l] = p0[l] + v[aligned2 + l];
    }
    res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
    for (i = aligned; i < n; i++) res = res + v[i];
  } else {
    res = v[0];
    for (i = 1; i < n; i++) res = res + v[i];
  }
  return res;
}

/* Redux.h:96-118 redux_novec_unroller: balanced binary tree over [start, start+len) (fixed-size
 * expressions without PacketAccessBit, completely unrolled). */
static float eig_tree_sum(const float *v, int start, int len) {
  if (len == 1) return v[start];
  return eig_tree_sum(v, start, len / 2) + eig_tree_sum(v, start + len / 2, len - len / 2);
}

/* ------------------------------------------------------------------------------------------------
 * D0  detection_boxes_for_sample  (dmz.cpp:279-341)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int x, y, w, h;
} rect_t;

static rect_t inset_rect(rect_t r, int hi, int vi) { /* geometry.h:10-15 */
  rect_t o = {r.x + hi, r.y + vi, r.w - 2 * hi, r.h - 2 * vi};
  return o;
}

static void detection_boxes(int img_w, int img_h, int orientation, rect_t boxes[4] /* top,bottom,left,right */) {
  int inset_v, slop_v, inset_h, slop_h;
  int width = (img_h * 4) / 3; /* central 4:3 region, dmz.cpp:286-288 */
  int left_margin = (img_w - width) / 2;
  rect_t image_rect, outer, inner;
  switch (orientation) {
    case 1:
    case 2: /* portrait: note the reference pairs the *horizontal* percent with the height (dmz.cpp:293-296) */
      inset_v = (int)roundf(kPortraitHorizontalPercentInset * img_h);
      slop_v = (int)roundf(kHorizontalPercentSlop * img_h);
      inset_h = (int)roundf(kPortraitVerticalPercentInset * width);
      slop_h = (int)roundf(kVerticalPercentSlop * width);
      break;
    case 3:
    case 4:
      inset_v = (int)roundf(kLandscapeVerticalPercentInset * img_h);
      slop_v = (int)roundf(kHorizontalPercentSlop * img_h);
      inset_h = (int)roundf(kLandscapeHorizontalPercentInset * width);
      slop_h = (int)roundf(kVerticalPercentSlop * width);
      break;
    default: inset_v = slop_v = inset_h = slop_h = 0; break;
  }
  image_rect.x = left_margin;
  image_rect.y = 0;
  image_rect.w = width - 1;
  image_rect.h = img_h - 1;
  outer = inset_rect(image_rect, inset_h - slop_h, inset_v - slop_v);
  inner = inset_rect(image_rect, inset_h + slop_h, inset_v + slop_v);
  boxes[0].x = inner.x, boxes[0].y = outer.y, boxes[0].w = inner.w, boxes[0].h = 2 * slop_v;
  boxes[1].x = inner.x, boxes[1].y = inner.y + inner.h, boxes[1].w = inner.w, boxes[1].h = 2 * slop_v;
  boxes[2].x = outer.x, boxes[2].y = inner.y, boxes[2].w = 2 * slop_h, boxes[2].h = inner.h;
  boxes[3].x = inner.x + inner.w, boxes[3].y = inner.y, boxes[3].w = 2 * slop_h, boxes[3].h = inner.h;
}

void orc_detection_boxes(int w, int h, int orientation, int32_t out[16]) {
  rect_t b[4];
  int i;
  detection_boxes(w, h, orientation, b);
  for (i = 0; i < 4; i++) {
    out[4 * i] = b[i].x;
    out[4 * i + 1] = b[i].y;
    out[4 * i + 2] = b[i].w;
    out[4 * i + 3] = b[i].h;
  }
}

/* ------------------------------------------------------------------------------------------------
 * D1  llcv_sobel7  (cv/sobel.cpp:476-478 -> cvSobel(.., 7))
 * ------------------------------------------------------------------------------------------------ */
void orc_sobel7(const uint8_t *img, int step, int w, int h, int16_t *dx, int16_t *dy) {
  orc_sobel_u8_s16(img, step, w, h, dx, w * 2, 1, 0, 7);
  orc_sobel_u8_s16(img, step, w, h, dy, w * 2, 0, 1, 7);
}

/* ------------------------------------------------------------------------------------------------
 * D2 + D3  adaptive Canny  (cv/canny.cpp:568-580 thresholds; canny.cpp:58-336 NMS + hysteresis)
 *
 * The reference's stack-based hysteresis marks exactly the candidate pixels (m > low, local maximum
 * along the quantised gradient direction) that are 8-connected to a seed (candidate with m > high);
 * its prev_flag / "upper neighbour already pushed" shortcuts only avoid redundant pushes.  This
 * restatement computes the same set with an explicit flood fill.
 * ------------------------------------------------------------------------------------------------ */
#define CANNY_SHIFT 15
#define TG22 ((int)(0.4142135623730950488016887242097 * (1 << CANNY_SHIFT) + 0.5))

void orc_adaptive_canny(const uint8_t *img, int step, int w, int h, const int16_t *dx, const int16_t *dy,
                        uint8_t *edges, int32_t *low_out, int32_t *high_out) {
  double mean = (orc_sum_abs_s16(dx, w * 2, w, h) + orc_sum_abs_s16(dy, w * 2, w, h)) / (w * h);
  double low_thresh = mean, high_thresh = 3.0f * mean;
  int low = orc_cv_floor(low_thresh), high = orc_cv_floor(high_thresh);
  int W = w + 2, i, j, top = 0;
  int *mag = (int *)calloc((size_t)W * (h + 2), sizeof(int)); /* zero border, canny.cpp:113,146,186 */
  uint8_t *map = (uint8_t *)malloc((size_t)W * (h + 2));      /* 0 candidate, 1 no edge, 2 edge */
  int *stack = (int *)malloc(sizeof(int) * (size_t)W * (h + 2));
  (void)img;
  (void)step;
  if (low_out) *low_out = low;
  if (high_out) *high_out = high;
  memset(map, 1, (size_t)W * (h + 2));
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) mag[(i + 1) * W + j + 1] = abs((int)dx[i * w + j]) + abs((int)dy[i * w + j]);
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) {
      const int *m0 = mag + (i + 1) * W + j + 1;
      int64_t x = dx[i * w + j], y = dy[i * w + j];
      int s = (x ^ y) < 0 ? -1 : 1;
      int m = *m0, is_max = 0;
      x = llabs(x);
      y = llabs(y);
      if (m > low) {
        int64_t tg22x = x * TG22;
        int64_t tg67x = tg22x + ((x + x) << CANNY_SHIFT);
        y <<= CANNY_SHIFT;
        if (y < tg22x) is_max = m > m0[-1] && m >= m0[1];
        else if (y > tg67x) is_max = m > m0[-W] && m >= m0[W];
        else is_max = m > m0[-W - s] && m > m0[W + s];
      }
      if (is_max) {
        if (m > high) {
          map[(i + 1) * W + j + 1] = 2;
          stack[top++] = (i + 1) * W + j + 1;
        } else {
          map[(i + 1) * W + j + 1] = 0;
        }
      }
    }
  while (top > 0) {
    static const int dyx[8][2] = {{0, -1}, {0, 1}, {-1, -1}, {-1, 0}, {-1, 1}, {1, -1}, {1, 0}, {1, 1}};
    int p = stack[--top], k;
    for (k = 0; k < 8; k++) {
      int q = p + dyx[k][0] * W + dyx[k][1];
      if (!map[q]) {
        map[q] = 2;
        stack[top++] = q;
      }
    }
  }
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) edges[i * w + j] = (uint8_t) - (map[(i + 1) * W + j + 1] >> 1);
  free(mag);
  free(map);
  free(stack);
}

/* ------------------------------------------------------------------------------------------------
 * D4  llcv_hough  (cv/hough.cpp:52-196) as called from best_line_for_sample (dmz.cpp:246-259)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int numangle;
  int tab_sin[16], tab_cos[16];
  float slope_a, slope_b, theta, theta_min;
} hough_consts;

static void hough_setup(int vertical, hough_consts *hc) {
  float base_angle = vertical ? kVerticalAngle : kHorizontalAngle;
  float theta_min = base_angle - kMaxAngleDeviationAllowed;
  float theta_max = base_angle + kMaxAngleDeviationAllowed;
  float theta = (float)CV_PI / 180.0f, irho = 1 / 1.0f, ang;
  float gat = kHoughGradientAngleThreshold;
  int n;
  hc->theta = theta;
  hc->theta_min = theta_min;
  hc->numangle = orc_cv_round((theta_max - theta_min) / theta);
  for (ang = theta_min, n = 0; n < hc->numangle; ang += theta, n++) {
    hc->tab_sin[n] = (int)floorf(1024 * sinf(ang) * irho);
    hc->tab_cos[n] = (int)floorf(1024 * cosf(ang) * irho);
  }
  if (vertical) {
    hc->slope_a = tanf((float)(CV_PI * (180 - gat) / 180.0f));
    hc->slope_b = tanf((float)(CV_PI * (180 + gat) / 180.0f));
  } else {
    hc->slope_a = tanf((float)(CV_PI * (90 - gat) / 180.0f));
    hc->slope_b = tanf((float)(CV_PI * (90 + gat) / 180.0f));
  }
}

static void hough_best_line(const uint8_t *edges, const int16_t *dx, const int16_t *dy, int w, int h, int vertical,
                            int threshold, orc_line *out) {
  hough_consts hc;
  int numrho = orc_cv_round(((w + h) * 2 + 1) / 1.0f);
  int i, j, n, r, max_val = 0, max_base = 0;
  int *accum;
  hough_setup(vertical, &hc);
  accum = (int *)calloc((size_t)(hc.numangle + 2) * (numrho + 2), sizeof(int));
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) {
      int use = 0;
      int16_t del_x, del_y;
      if (!edges[i * w + j]) continue;
      del_x = dx[i * w + j];
      del_y = dy[i * w + j];
      if (del_x != 0) {
        float slope = (float)del_y / (float)del_x;
        if (vertical) use = slope >= hc.slope_a && slope <= hc.slope_b;
        else use = slope >= hc.slope_a || slope <= hc.slope_b;
      } else {
        use = !vertical;
      }
      if (use)
        for (n = 0; n < hc.numangle; n++) {
          r = (j * hc.tab_cos[n] + i * hc.tab_sin[n]) >> 10;
          r += (numrho - 1) / 2;
          accum[(n + 1) * (numrho + 2) + r + 1]++;
        }
    }
  for (r = 0; r < numrho; r++)
    for (n = 0; n < hc.numangle; n++) {
      int base = (n + 1) * (numrho + 2) + r + 1;
      if (accum[base] > max_val) {
        max_val = accum[base];
        max_base = base;
      }
    }
  out->max_votes = max_val;
  out->found = 0;
  out->rho = FLT_MAX; /* ParametricLineNone, geometry.h:24-29 */
  out->theta = FLT_MAX;
  out->r = out->n = 0;
  if (max_val > threshold) {
    float scale = 1.0f / (numrho + 2);
    int nn = orc_cv_floor(max_base * scale) - 1;
    int rr = max_base - (nn + 1) * (numrho + 2) - 1;
    out->found = 1;
    out->r = rr;
    out->n = nn;
    out->rho = (rr - (numrho - 1) * 0.5f) * 1.0f;
    out->theta = nn * hc.theta + hc.theta_min;
  }
  free(accum);
}

/* best_line_for_sample (dmz.cpp:224-271) */
void orc_best_line(const uint8_t *img, int step, int w, int h, int vertical, orc_line *out) {
  int16_t *dx = (int16_t *)malloc((size_t)w * h * 2), *dy = (int16_t *)malloc((size_t)w * h * 2);
  uint8_t *edges = (uint8_t *)malloc((size_t)w * h);
  int i, threshold = (w > h ? w : h) / kHoughThresholdLengthDivisor;
  memset(out, 0, sizeof(*out));
  orc_sobel7(img, step, w, h, dx, dy);
  orc_adaptive_canny(img, step, w, h, dx, dy, edges, &out->low, &out->high);
  for (i = 0; i < w * h; i++) out->n_edge_px += edges[i] != 0;
  hough_best_line(edges, dx, dy, w, h, vertical, threshold, out);
  free(dx);
  free(dy);
  free(edges);
}

/* ------------------------------------------------------------------------------------------------
 * D5  geometry  (geometry.cpp:14-43) and dmz_detect_edges (dmz.cpp:346-439)
 * ------------------------------------------------------------------------------------------------ */
static void line_by_shifting_origin(float rho, float theta, int x_off, int y_off, float *new_rho, float *new_theta) {
  /* atan(float) resolves to the float overload (atanf) in the reference's C++ */
  double offset_angle = x_off == 0 ? CV_PI / 2.0f : (double)atanf((float)y_off / (float)x_off);
  double delta_angle = theta - offset_angle + CV_PI / 2.0f;
  double offset_magnitude = sqrt((double)(x_off * x_off + y_off * y_off));
  double delta_rho = offset_magnitude * cos(CV_PI / 2 - delta_angle);
  *new_theta = theta;
  *new_rho = (float)(rho + delta_rho);
}

/* parametricIntersect: Eigen Matrix2f determinant / inverse (LU/Inverse.h:70-89), lazy 2x2 * 2x1 product */
static int parametric_intersect(float rho1, float theta1, float rho2, float theta2, float *x, float *y) {
  float t00, t01, t10, t11, det, invdet, i00, i10, i01, i11;
  if (theta1 == FLT_MAX || theta2 == FLT_MAX) return 0;
  t00 = cosf(theta1), t01 = sinf(theta1), t10 = cosf(theta2), t11 = sinf(theta2);
  det = t00 * t11 - t10 * t01;
  if (det < 1e-10) return 0;
  invdet = 1.0f / det;
  i00 = t11 * invdet;
  i10 = -t10 * invdet;
  i01 = -t01 * invdet;
  i11 = t00 * invdet;
  *x = i00 * rho1 + i01 * rho2;
  *y = i10 * rho1 + i11 * rho2;
  return 1;
}

int orc_detect_edges(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                     int orientation, orc_detect *out) {
  const uint8_t *planes[3] = {y, cb, cr};
  int pw[3] = {w, w / 2, w / 2}, ph[3] = {h, h / 2, h / 2}, ps[3] = {ystep, cstep, cstep};
  float rho_mult[3] = {1.0f, 2.0f, 2.0f};
  rect_t boxes[3][4];
  /* dmz_edges member order is top,left,bottom,right; detection order is top,bottom,left,right */
  static const int box_of_edge[4] = {0, 2, 1, 3};
  static const int detect_order[4] = {0, 2, 1, 3}; /* edge slots visited: top, bottom, left, right */
  int e, i;
  memset(out, 0, sizeof(*out));
  for (i = 0; i < 3; i++) detection_boxes(pw[i], ph[i], orientation, boxes[i]);
  for (e = 0; e < 4; e++) {
    int slot = detect_order[e];  /* index into out->found (top,left,bottom,right) */
    int box = box_of_edge[slot]; /* index into boxes (top,bottom,left,right) */
    int vertical = (box >= 2);
    out->found[slot] = 0;
    for (i = 0; i < 3 && !out->found[slot]; i++) {
      rect_t r = boxes[i][box];
      orc_line l;
      float nr, nt;
      orc_best_line(planes[i] + (size_t)r.y * ps[i] + r.x, ps[i], r.w, r.h, vertical, &l);
      line_by_shifting_origin(l.rho, l.theta, r.x, r.y, &nr, &nt);
      nr *= rho_mult[i];
      out->rho[slot] = nr;
      out->theta[slot] = nt;
      out->found[slot] = !(nt == FLT_MAX);
    }
  }
  out->all_found = 0;
  if (out->found[0] && out->found[1] && out->found[2] && out->found[3]) {
    /* slots: 0 top, 1 left, 2 bottom, 3 right; corners: tl, bl, tr, br */
    float c[8];
    int tl = parametric_intersect(out->rho[0], out->theta[0], out->rho[1], out->theta[1], &c[0], &c[1]);
    int bl = parametric_intersect(out->rho[2], out->theta[2], out->rho[1], out->theta[1], &c[2], &c[3]);
    int tr = parametric_intersect(out->rho[0], out->theta[0], out->rho[3], out->theta[3], &c[4], &c[5]);
    int br = parametric_intersect(out->rho[2], out->theta[2], out->rho[3], out->theta[3], &c[6], &c[7]);
    if (tl && bl && tr && br) {
      memcpy(out->corners, c, sizeof(c));
      out->all_found = 1;
    }
  }
  return out->all_found;
}

/* ------------------------------------------------------------------------------------------------
 * W1  llcv_calc_persp_transform  (cv/warp.cpp:34-125): x = A.householderQr().solve(b), float.
 * Eigen 3.2.4 order: Householder/HouseholderQR.h unblocked (block size 8 == whole matrix),
 * Householder.h makeHouseholder / applyHouseholderOnTheLeft, HouseholderSequence applyThisOnTheLeft,
 * TriangularSolverVector.h (col-major, Upper).  squaredNorm, the per-column inner products of the
 * factorisation and the 1x1 inner products of the solve all go through the vectorised redux with
 * alignedStart == 0 (eig_redux_sum); everything else is element-wise.
 * ------------------------------------------------------------------------------------------------ */
#define A_(i, j) a[(i) + 8 * (j)]

static float *g_dbg_hcoef = NULL, *g_dbg_c = NULL; /* test taps (orc_dbg_qr) */

static void householder_qr_solve8(float *a /* col-major 8x8, overwritten */, const float *bvec, float *xout) {
  float hcoef[8], tmp[8], c[8], v[8];
  int k, i, j;
  for (k = 0; k < 8; k++) {
    int rem_rows = 8 - k, n = rem_rows - 1;
    float c0 = A_(k, k), tail_sq, beta, tau;
    for (i = 0; i < n; i++) v[i] = A_(k + 1 + i, k) * A_(k + 1 + i, k);
    tail_sq = rem_rows == 1 ? 0.0f : eig_redux_sum(v, n);
    if (tail_sq == 0.0f) {
      tau = 0.0f;
      beta = c0;
      for (i = 0; i < n; i++) A_(k + 1 + i, k) = 0.0f;
    } else {
      float denom;
      beta = sqrtf(c0 * c0 + tail_sq);
      if (c0 >= 0.0f) beta = -beta;
      denom = c0 - beta;
      for (i = 0; i < n; i++) A_(k + 1 + i, k) = A_(k + 1 + i, k) / denom;
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    A_(k, k) = beta;
    if (rem_rows == 1) {
      /* bottomRightCorner(1, 0): empty */
    } else {
      for (j = k + 1; j < 8; j++) {
        /* essential.adjoint() * bottom: each coefficient is (lhs.row.transpose().cwiseProduct(rhs.col)).sum(),
         * i.e. the vectorised redux (verified against the vendored Eigen: 8400/8400 coefficients) */
        for (i = 0; i < n; i++) v[i] = A_(k + 1 + i, k) * A_(k + 1 + i, j);
        tmp[j] = eig_redux_sum(v, n);
      }
      for (j = k + 1; j < 8; j++) tmp[j] += A_(k, j);
      for (j = k + 1; j < 8; j++) A_(k, j) -= tau * tmp[j];
      for (j = k + 1; j < 8; j++)
        for (i = 0; i < n; i++) A_(k + 1 + i, j) -= (A_(k + 1 + i, k) * tau) * tmp[j];
    }
  }
  /* c = Q^T b */
  for (i = 0; i < 8; i++) c[i] = bvec[i];
  for (k = 0; k < 8; k++) {
    int n = 7 - k;
    float tau = hcoef[k], t;
    if (n == 0) {
      c[k] *= (1.0f - tau);
    } else {
      for (i = 0; i < n; i++) v[i] = A_(k + 1 + i, k) * c[k + 1 + i];
      t = eig_redux_sum(v, n);
      t += c[k];
      c[k] -= tau * t;
      for (i = 0; i < n; i++) c[k + 1 + i] -= (A_(k + 1 + i, k) * tau) * t;
    }
  }
  if (g_dbg_hcoef) memcpy(g_dbg_hcoef, hcoef, sizeof(hcoef));
  if (g_dbg_c) memcpy(g_dbg_c, c, sizeof(c));
  /* back substitution, column-major, one panel of 8 */
  for (k = 0; k < 8; k++) {
    i = 7 - k;
    c[i] /= A_(i, i);
    for (j = 0; j < i; j++) c[j] -= c[i] * A_(j, i);
  }
  for (i = 0; i < 8; i++) xout[i] = c[i];
}

/* test tap: factor + solve a caller-supplied column-major system, returning the intermediates */
void orc_dbg_qr(const float *a_colmajor, const float *b, float *qr, float *hc, float *x, float *c_after_q) {
  memcpy(qr, a_colmajor, 64 * sizeof(float));
  g_dbg_hcoef = hc;
  g_dbg_c = c_after_q;
  householder_qr_solve8(qr, b, x);
  g_dbg_hcoef = g_dbg_c = NULL;
}

void orc_calc_persp_transform(const float s[8], const float d[8], float M[9]) {
  float a[64], b[8], x[8];
  int i;
  memset(a, 0, sizeof(a));
  for (i = 0; i < 4; i++) {
    float sx = s[2 * i], sy = s[2 * i + 1], dx = d[2 * i], dy = d[2 * i + 1];
    A_(i, 0) = sx, A_(i, 1) = sy, A_(i, 2) = 1, A_(i, 3) = 0, A_(i, 4) = 0, A_(i, 5) = 0;
    A_(i, 6) = -sx * dx;
    A_(i, 7) = -sy * dx;
    A_(i + 4, 0) = 0, A_(i + 4, 1) = 0, A_(i + 4, 2) = 0, A_(i + 4, 3) = sx, A_(i + 4, 4) = sy, A_(i + 4, 5) = 1;
    A_(i + 4, 6) = -sx * dy;
    A_(i + 4, 7) = -sy * dy;
    b[i] = dx;
    b[i + 4] = dy;
  }
  householder_qr_solve8(a, b, x);
  M[0] = x[0], M[1] = x[1], M[2] = x[2];
  M[3] = x[3], M[4] = x[4], M[5] = x[5];
  M[6] = x[6], M[7] = x[7], M[8] = 1.0f;
}

/* W0 + W2: dmz_transform_card (dmz.cpp:443-497) -> llcv_unwarp (cv/warp.cpp:130-167) */
void orc_transform_card(const uint8_t *y, int w, int h, int ystep, const float c[8], int orientation, uint8_t *card) {
  orc_transform_card_up(y, w, h, ystep, c, orientation, 0, card);
}

/* upsample: the sample is a half-size chroma plane, so the (luma-space) corners are halved (dmz.cpp:473-481;
 * llcv_warp_auto_upsamples() is false outside IOS_DMZ, cv/warp.cpp:26-32) */
void orc_transform_card_up(const uint8_t *y, int w, int h, int ystep, const float c[8], int orientation, int upsample, uint8_t *card) {
  /* corner order in c: tl, bl, tr, br */
  const float *tl = c, *bl = c + 2, *tr = c + 4, *br = c + 6;
  const float *sp[4];
  float src[8], dst[8] = {0, 0, 427, 0, 0, 269, 427, 269}, M[9];
  int i;
  switch (orientation) {
    case 1: sp[0] = bl, sp[1] = tl, sp[2] = br, sp[3] = tr; break;
    case 4: sp[0] = br, sp[1] = bl, sp[2] = tr, sp[3] = tl; break;
    case 2: sp[0] = tr, sp[1] = br, sp[2] = tl, sp[3] = bl; break;
    case 3:
    default: sp[0] = tl, sp[1] = tr, sp[2] = bl, sp[3] = br; break;
  }
  for (i = 0; i < 4; i++) src[2 * i] = sp[i][0], src[2 * i + 1] = sp[i][1];
  if (upsample)
    for (i = 0; i < 8; i++) src[i] /= 2.0f;
  orc_calc_persp_transform(src, dst, M);
  orc_warp_perspective_u8(y, ystep, w, h, card, kCardW, kCardW, kCardH, M);
}

/* ------------------------------------------------------------------------------------------------
 * model weights
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  float conv_w[8][9], conv_b[8], hid_w[32][320], hid_b[32], log_w[10][32], log_b[10];
} cnn_weights;
typedef struct {
  float hid_w[50][204], hid_b[50], log_w[3][50], log_b[3];
} mlp_weights;

static mlp_weights g_vseg;
static cnn_weights g_cnn[3];
static int g_weights_loaded = 0;

/* expiry digit CNN, models/expiry/modelc_bf4dd6c8.cpp (blob order = table order in the generated file) */
typedef struct {
  float c1w[50][25], c1b[50], c2w[40][50][25], c2b[40], hw[176][120], hb[176], lw[10][176], lb[10];
} expiry_weights;
static expiry_weights g_exp;
static int g_exp_loaded = 0;

static int read_blob(const char *dir, const char *name, void *dst, size_t bytes) {
  char path[1024];
  FILE *f;
  size_t got;
  snprintf(path, sizeof(path), "%s/%s", dir, name);
  f = fopen(path, "rb");
  if (!f) return -1;
  got = fread(dst, 1, bytes, f);
  fclose(f);
  return got == bytes ? 0 : -1;
}

int orc_load_weights(const char *dir) {
  static const char *cnn_names[3] = {"modelc_5c241121.bin", "modelc_01266c1b.bin", "modelc_b00bf70c.bin"};
  int i;
  if (read_blob(dir, "modelm_befe75da.bin", &g_vseg, sizeof(g_vseg))) return -1;
  for (i = 0; i < 3; i++)
    if (read_blob(dir, cnn_names[i], &g_cnn[i], sizeof(cnn_weights))) return -1;
  g_weights_loaded = 1;
  g_exp_loaded = read_blob(dir, "modelc_bf4dd6c8.bin", &g_exp, sizeof(g_exp)) == 0;
  return 0;
}

static void need_weights(void) {
  if (!g_weights_loaded) {
    fprintf(stderr, "dmz_oracle: orc_load_weights() was not called\n");
    abort();
  }
}

/* ------------------------------------------------------------------------------------------------
 * V1 + V2  (scan/n_vseg.cpp:39-47, models/generated/modelm_befe75da.cpp:1770-1786)
 * ------------------------------------------------------------------------------------------------ */
void orc_vseg_model(const float *x, float probs[3]) {
  float hid[50], o[3], sum;
  int i, j;
  need_weights();
  for (i = 0; i < 50; i++) {
    float acc = 0.0f;
    for (j = 0; j < 204; j++) acc += g_vseg.hid_w[i][j] * x[j];
    hid[i] = tanhf(acc + g_vseg.hid_b[i]);
  }
  for (i = 0; i < 3; i++) {
    float acc = 0.0f;
    for (j = 0; j < 50; j++) acc += g_vseg.log_w[i][j] * hid[j];
    o[i] = expf(acc + g_vseg.log_b[i]);
  }
  sum = (o[0] + o[1]) + o[2];
  for (i = 0; i < 3; i++) probs[i] = o[i] / sum;
}

static void vseg_row_input(const uint8_t *card, int row, float x[204]) {
  uint8_t grad[408], down[204];
  /* ROI (10, row, 408, 1) treated as an isolated 408x1 image */
  orc_morph_grad_cross3_u8(card + (size_t)row * kCardW + 10, kCardW, 408, 1, grad, 408);
  orc_resize_half_width_u8(grad, 408, 408, 1, down, 204);
  orc_convert_scale_u8_f32(down, 204, 204, 1, x, 204 * 4, 1.0f / 255.0f);
  orc_normalize_minmax_f32(x, 204 * 4, 204, 1);
}

void orc_vseg_row(const uint8_t *card, int row, float probs[3]) {
  float x[204];
  vseg_row_input(card, row, x);
  orc_vseg_model(x, probs);
}

/* best_segmentation_for_vseg_scores (scan/n_vseg.cpp:49-92) */
static void best_segmentation(const float *visa, const float *amex, orc_vseg *best) {
  float vsum = 0.0f, asum = 0.0f, vring[27], aring[27];
  int y;
  best->score = 0.0f;
  best->pattern_type = 0;
  best->y_offset = 0;
  for (y = 0; y < 270; y++) {
    int bi = y % 27;
    vsum += visa[y];
    asum += amex[y];
    vring[bi] = visa[y];
    aring[bi] = amex[y];
    if (y >= 26) {
      int ni = (y + 1) % 27;
      if (vsum > best->score) {
        best->score = vsum;
        best->pattern_type = 1;
        best->y_offset = (uint16_t)(y - 26);
      }
      if (asum > best->score) {
        best->score = asum;
        best->pattern_type = 2;
        best->y_offset = (uint16_t)(y - 26);
      }
      vsum -= vring[ni];
      asum -= aring[ni];
    }
  }
}

/* best_n_vseg (scan/n_vseg.cpp:94-168) */
void orc_best_n_vseg(const uint8_t *card, orc_vseg *best) {
  static const uint8_t pat[3][19] = {{0},
                                     {1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1},
                                     {1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 0}};
  static const uint8_t pat_len[3] = {0, 19, 17}, num_len[3] = {0, 16, 15};
  float visa[270], amex[270], p[3];
  int y, lo, hi;
  memset(best, 0, sizeof(*best));
  memset(visa, 0, sizeof(visa));
  memset(amex, 0, sizeof(amex));
  for (y = 0; y < 270; y += 4) {
    orc_vseg_row(card, y, p);
    visa[y] = p[1];
    amex[y] = p[2];
  }
  best_segmentation(visa, amex, best);
  lo = best->y_offset < 8 ? 0 : best->y_offset - 8;
  if (lo > 270) lo = 270;
  hi = best->y_offset + 27 + 8;
  if (hi > 270) hi = 270;
  for (y = lo; y < hi; y++)
    if (visa[y] == 0 && amex[y] == 0) {
      orc_vseg_row(card, y, p);
      visa[y] = p[1];
      amex[y] = p[2];
    }
  best_segmentation(visa, amex, best);
  best->number_pattern_length = pat_len[best->pattern_type];
  memcpy(best->number_pattern, pat[best->pattern_type], 19);
  best->number_length = num_len[best->pattern_type];
}

/* ------------------------------------------------------------------------------------------------
 * H0 + H1  (scan/n_hseg.cpp:39-152)
 * ------------------------------------------------------------------------------------------------ */
static const float k_number_grad_sum_pattern[19] = {
    0.26228655f, 0.30289554f, 0.34632607f, 0.38725636f, 0.42745813f, 0.45875135f, 0.46498017f,
    0.45258447f, 0.43045216f, 0.42430462f, 0.44796554f, 0.47726529f, 0.48471646f, 0.46457738f,
    0.42799847f, 0.38851183f, 0.33966308f, 0.28802608f, 0.25377602f};

static void hseg_constrained(const float *grad, const orc_vseg *vseg, orc_hseg *best, float wmin, float wmax,
                             float wstep, unsigned omin, unsigned omax, unsigned ostep) {
  float pattern[428], diff[428], width;
  uint16_t temp_offsets[16];
  for (width = wmin; width < wmax; width += wstep) {
    float pattern_width = vseg->number_pattern_length * width;
    uint16_t pattern_offset_max = (uint16_t)omax;
    uint16_t maximum = (uint16_t)(428 - lrintf(pattern_width));
    uint16_t offset;
    if (pattern_offset_max == 0xFFFF || pattern_offset_max > maximum) pattern_offset_max = maximum;
    for (offset = (uint16_t)omin; offset < pattern_offset_max; offset = (uint16_t)(offset + ostep)) {
      int in_bounds = 1, oi = 0, pi, i;
      memset(pattern, 0, sizeof(pattern));
      memset(temp_offsets, 0, sizeof(temp_offsets));
      for (pi = 0; pi < vseg->number_pattern_length; pi++) {
        if (vseg->number_pattern[pi]) {
          uint16_t center = (uint16_t)(offset + lrintf(pi * width));
          if (center + 19 < 428) memcpy(pattern + center, k_number_grad_sum_pattern, sizeof(k_number_grad_sum_pattern));
          else in_bounds = 0;
          if (oi < 16) temp_offsets[oi] = center;
          oi++;
        }
      }
      if (in_bounds) {
        float score;
        for (i = 0; i < 428; i++) diff[i] = fabsf(grad[i] - pattern[i]);
        score = eig_redux_sum(diff, 428); /* (a - b).cwiseAbs().sum() on 1x428, vectorised redux */
        if (score < best->score) {
          memcpy(best->offsets, temp_offsets, sizeof(temp_offsets));
          best->score = score;
          best->number_width = width;
          best->pattern_offset = offset;
        }
      }
    }
  }
}

void orc_best_n_hseg(const uint8_t *card, const orc_vseg *vseg, orc_hseg *best) {
  uint8_t grad[428 * 27];
  float gsum[428];
  memset(best, 0, sizeof(*best));
  orc_morph_grad_cross3_u8(card + (size_t)vseg->y_offset * kCardW, kCardW, 428, 27, grad, 428);
  orc_reduce_cols_sum_u8_f32(grad, 428, 428, 27, gsum);
  orc_normalize_minmax_f32(gsum, 428 * 4, 428, 1);
  best->n_offsets = vseg->number_length;
  best->score = 428.0f;
  best->number_width = 0.0f;
  hseg_constrained(gsum, vseg, best, 17.1f, 19.7f, 0.5f, 0, 0xFFFF, 10);
  hseg_constrained(gsum, vseg, best, best->number_width - 0.5f, best->number_width + 0.5f, 0.2f,
                   best->pattern_offset < 10 ? 0 : best->pattern_offset - 10, (uint16_t)(best->pattern_offset + 10), 1);
  hseg_constrained(gsum, vseg, best, best->number_width - 0.2f, best->number_width + 0.2f, 0.1f,
                   best->pattern_offset < 3 ? 0 : best->pattern_offset - 3, (uint16_t)(best->pattern_offset + 3), 1);
  hseg_constrained(gsum, vseg, best, best->number_width - 0.1f, best->number_width + 0.1f, 0.05f,
                   best->pattern_offset < 3 ? 0 : best->pattern_offset - 3, (uint16_t)(best->pattern_offset + 3), 1);
}

/* ------------------------------------------------------------------------------------------------
 * C0..C2  (scan/n_categorize.cpp:45-108, cv/stats.cpp:116-159, models/generated/modelc_*.cpp:1844-1937)
 * ------------------------------------------------------------------------------------------------ */
static void equalize_hist_u8(uint8_t *img, int step, int w, int h) {
  int hist[256], x, y, i, sum = 0;
  uint8_t lut[256];
  float scale = 255.f / (w * h);
  memset(hist, 0, sizeof(hist));
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) hist[img[y * step + x]]++;
  for (i = 0; i < 256; i++) {
    int val;
    sum += hist[i];
    val = orc_cv_round(sum * scale);
    lut[i] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
  }
  lut[0] = 0;
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) img[y * step + x] = lut[img[y * step + x]];
}

void orc_digit_patch_prep(const uint8_t *img, int step, float *patch) {
  uint8_t g[27 * 20];
  orc_morph_grad_cross3_u8(img, step, 19, 27, g, 20);
  equalize_hist_u8(g, 20, 19, 27);
  orc_convert_scale_u8_f32(g, 20, 19, 27, patch, 19 * 4, 1.0f / 255.0f);
}

static void cnn_apply(const cnn_weights *m, const float *x /*27x19*/, float out[10]) {
  float feat[320], hid[32], o[10], sum;
  int k, r, c, i, j;
  for (k = 0; k < 8; k++) {
    float conv[24][15];
    for (r = 0; r < 24; r++)
      for (c = 0; c < 15; c++) {
        float prod[9];
        for (i = 0; i < 3; i++)
          for (j = 0; j < 3; j++) prod[i * 3 + j] = m->conv_w[k][i * 3 + j] * x[(r + i) * 19 + c + j];
        conv[r][c] = eig_tree_sum(prod, 0, 9); /* Matrix3f::sum(): unrolled binary tree */
      }
    for (r = 0; r < 8; r++)
      for (c = 0; c < 5; c++) {
        float mx = conv[r * 3][c * 3];
        for (i = 0; i < 3; i++)
          for (j = 0; j < 3; j++)
            if (conv[r * 3 + i][c * 3 + j] > mx) mx = conv[r * 3 + i][c * 3 + j];
        feat[k * 40 + r * 5 + c] = tanhf(mx + m->conv_b[k]);
      }
  }
  for (i = 0; i < 32; i++) {
    float acc = 0.0f;
    for (j = 0; j < 320; j++) acc += m->hid_w[i][j] * feat[j];
    hid[i] = tanhf(acc + m->hid_b[i]);
  }
  for (i = 0; i < 10; i++) {
    float acc = 0.0f;
    for (j = 0; j < 32; j++) acc += m->log_w[i][j] * hid[j];
    o[i] = expf(acc + m->log_b[i]);
  }
  sum = eig_tree_sum(o, 0, 10);
  for (i = 0; i < 10; i++) out[i] = o[i] / sum;
}

void orc_digit_models(const float *patch, float *out) {
  int j;
  need_weights();
  cnn_apply(&g_cnn[0], patch, out + 10);
  cnn_apply(&g_cnn[1], patch, out + 20);
  cnn_apply(&g_cnn[2], patch, out + 30);
  for (j = 0; j < 10; j++) {
    float r0 = out[10 + j], r1 = out[20 + j], r2 = out[30 + j];
    float mx = r0 > r1 ? r0 : r1;
    mx = mx > r2 ? mx : r2;
    out[j] = (((r0 + r1) + r2) - mx) / 2.0f;
  }
}

void orc_number_scores(const uint8_t *card, int y_offset, const orc_hseg *hseg, float *scores) {
  int d;
  memset(scores, 0, 160 * sizeof(float));
  for (d = 0; d < hseg->n_offsets && d < 16; d++) {
    float patch[27 * 19], out[40];
    orc_digit_patch_prep(card + (size_t)y_offset * kCardW + hseg->offsets[d], kCardW, patch);
    orc_digit_models(patch, out);
    memcpy(scores + d * 10, out, 10 * sizeof(float));
  }
}

/* ------------------------------------------------------------------------------------------------
 * E0  expiry digit: prepare_image_for_cat (scan/expiry_categorize.cpp:37-73) + applyc_bf4dd6c8
 *     (models/expiry/modelc_bf4dd6c8.cpp:12500-13505)
 * ------------------------------------------------------------------------------------------------ */
void orc_expiry_patch_prep(const uint8_t *img, int step, float *out /* 16 x 11 */) {
  uint8_t g[16 * 12], b[16 * 12];
  const int aperture = 3;
  const double space_sigma = (aperture / 2.0 - 1) * 0.3 + 0.8, color_sigma = (aperture - 1) / 3.0;
  orc_morph_grad_cross3_u8(img, step, 11, 16, g, 12);       /* cvMorphologyEx(GRADIENT, 3x3 cross) on the 11x16 ROI */
  equalize_hist_u8(g, 12, 11, 16);                           /* llcv_equalize_hist, scale 255.f / 176 */
  /* cvSmooth(.., CV_BILATERAL, 3, 3, space_sigma, color_sigma): param3 -> sigmaColor, param4 -> sigmaSpace */
  orc_bilateral_u8(g, 12, 11, 16, b, 12, aperture, space_sigma, color_sigma);
  orc_convert_scale_u8_f32(b, 12, 11, 16, out, 11 * 4, 1.0f / 255.0f);
}

static float relu(float v) { return v > 0.0f ? v : 0.0f; }

/* in: 16 x 11 floats.  Optional taps: l1 50 x 70, l2 40 x 3, hid 176 (the layers the reference's KAT checks). */
void orc_expiry_digit_model(const float *in, float *out10, float *l1_out, float *l2_out, float *hid_out) {
  float x[16][11], l1[50][10][7], l2[40][3], hid[176], o[10], mean = 0.0f, sum;
  int f, r, c, i, j, k;
  if (!g_exp_loaded) {
    fprintf(stderr, "dmz_oracle: expiry weights (modelc_bf4dd6c8.bin) not loaded\n");
    abort();
  }
  /* normalized_input = input - input.mean(): Eigen's vectorised linear redux over the 176 coefficients, / 176 */
  mean = eig_redux_sum(in, 176) / 176.0f;
  for (r = 0; r < 16; r++)
    for (c = 0; c < 11; c++) x[r][c] = in[r * 11 + c] - mean;
  /* layer 1: 50 x "full" 5x5 cross-correlation with zero padding (20 x 14), 2x2 max pool (10 x 7), + bias, ReLU */
  for (f = 0; f < 50; f++) {
    float conv[20][14];
    for (r = 0; r < 20; r++)
      for (c = 0; c < 14; c++) {
        float prod[25];
        for (i = 0; i < 5; i++)
          for (j = 0; j < 5; j++) {
            int rr = r - 4 + i, cc = c - 4 + j;
            prod[i * 5 + j] = g_exp.c1w[f][i * 5 + j] * ((rr >= 0 && rr < 16 && cc >= 0 && cc < 11) ? x[rr][cc] : 0.0f);
          }
        conv[r][c] = 0.0f + eig_tree_sum(prod, 0, 25); /* kernel.cwiseProduct(sub).sum(): unrolled tree; += onto Zero() */
      }
    for (r = 0; r < 10; r++)
      for (c = 0; c < 7; c++) {
        float m = conv[2 * r][2 * c];
        if (conv[2 * r][2 * c + 1] > m) m = conv[2 * r][2 * c + 1];
        if (conv[2 * r + 1][2 * c] > m) m = conv[2 * r + 1][2 * c];
        if (conv[2 * r + 1][2 * c + 1] > m) m = conv[2 * r + 1][2 * c + 1];
        l1[f][r][c] = relu(m + g_exp.c1b[f]);
      }
  }
  /* layer 2: 40 maps, each the sum over 50 input maps of a valid 5x5 correlation (6 x 3), 2x3 max pool (3 x 1) */
  for (f = 0; f < 40; f++) {
    float acc[6][3];
    memset(acc, 0, sizeof(acc));
    for (k = 0; k < 50; k++)
      for (r = 0; r < 6; r++)
        for (c = 0; c < 3; c++) {
          float prod[25];
          for (i = 0; i < 5; i++)
            for (j = 0; j < 5; j++) prod[i * 5 + j] = g_exp.c2w[f][k][i * 5 + j] * l1[k][r + i][c + j];
          acc[r][c] += eig_tree_sum(prod, 0, 25);
        }
    for (r = 0; r < 3; r++) {
      float m = acc[2 * r][0];
      for (i = 0; i < 2; i++)
        for (j = 0; j < 3; j++)
          if (acc[2 * r + i][j] > m) m = acc[2 * r + i][j];
      l2[f][r] = relu(m + g_exp.c2b[f]);
    }
  }
  for (i = 0; i < 176; i++) {
    float a = 0.0f;
    for (j = 0; j < 120; j++) a += g_exp.hw[i][j] * (&l2[0][0])[j];
    hid[i] = relu(a + g_exp.hb[i]);
  }
  for (i = 0; i < 10; i++) {
    float a = 0.0f;
    for (j = 0; j < 176; j++) a += g_exp.lw[i][j] * hid[j];
    o[i] = expf(a + g_exp.lb[i]);
  }
  sum = eig_tree_sum(o, 0, 10);
  for (i = 0; i < 10; i++) out10[i] = o[i] / sum;
  if (l1_out) memcpy(l1_out, l1, sizeof(l1));
  if (l2_out) memcpy(l2_out, l2, sizeof(l2));
  if (hid_out) memcpy(hid_out, hid, sizeof(hid));
}

/* ------------------------------------------------------------------------------------------------
 * S0  scan_card_image  (scan/frame.cpp:24-81)
 * ------------------------------------------------------------------------------------------------ */
void orc_scan_card_image(const uint8_t *card, orc_scan *out) {
  float number_score;
  memset(out, 0, sizeof(*out));
  orc_best_n_vseg(card, &out->vseg);
  if (out->vseg.y_offset < (kCardH - kNumberH) / 2) {
    out->upside_down = 1;
    return;
  }
  out->usable = out->vseg.score > 15;
  if (!out->usable) return;
  orc_best_n_hseg(card, &out->vseg, &out->hseg);
  orc_number_scores(card, out->vseg.y_offset, &out->hseg, out->scores);
  number_score = out->hseg.n_offsets - eig_redux_sum(out->scores, 160);
  out->usable = number_score < 3;
}

uint32_t orc_card_check(const uint8_t *p, size_t n) {
  uint32_t c = 0;
  size_t i;
  for (i = 0; i < n; i++) c += (uint32_t)(i + 1) * p[i];
  return c;
}

void orc_process_frame(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                       int orientation, orc_frame_record *rec, uint8_t *card_out) {
  uint8_t *card;
  memset(rec, 0, sizeof(*rec));
  if (!orc_detect_edges(y, w, h, ystep, cb, cr, cstep, orientation, &rec->detect)) return;
  card = card_out ? card_out : (uint8_t *)malloc(kCardW * kCardH);
  orc_transform_card(y, w, h, ystep, rec->detect.corners, orientation, card);
  rec->card_check = orc_card_check(card, kCardW * kCardH);
  orc_scan_card_image(card, &rec->scan);
  if (!card_out) free(card);
}

/* ------------------------------------------------------------------------------------------------
 * S1  scanner session  (scan/scan.cpp:19-194, dmz_olm.cpp luhn / prefix table)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  uint16_t count15, count16;
  float agg15[160], agg16[160];
  int complete;
  uint8_t digits[16];
  int n_numbers;
} scanner_t;

void *orc_scanner_new(void) {
  scanner_t *s = (scanner_t *)calloc(1, sizeof(scanner_t));
  return s;
}
void orc_scanner_free(void *s) { free(s); }
void orc_scanner_reset(void *s) { memset(s, 0, sizeof(scanner_t)); }

void orc_scanner_add_scan(void *state, const orc_scan *r) {
  scanner_t *s = (scanner_t *)state;
  float *agg;
  int i;
  if (s->complete) return; /* still_need_to_collect_card_number == false: scores are not collected */
  if (r->upside_down || !r->usable) return;
  if (r->hseg.n_offsets == 15) agg = s->agg15, s->count15++;
  else if (r->hseg.n_offsets == 16) agg = s->agg16, s->count16++;
  else return;
  for (i = 0; i < 160; i++) agg[i] *= 0.8f;
  for (i = 0; i < 160; i++) agg[i] += r->scores[i] * (1 - 0.8f);
}

void orc_scanner_add_frame(void *state, const uint8_t *card, orc_scan *out) {
  orc_scan r;
  scanner_t *s = (scanner_t *)state;
  if (s->complete) {
    /* scan_card_image(collect_card_number = false): vseg only (frame.cpp:49) */
    memset(&r, 0, sizeof(r));
    orc_best_n_vseg(card, &r.vseg);
    if (r.vseg.y_offset < (kCardH - kNumberH) / 2) r.upside_down = 1;
    else r.usable = r.vseg.score > 15;
  } else {
    orc_scan_card_image(card, &r);
    orc_scanner_add_scan(state, &r);
  }
  if (out) *out = r;
}

void orc_scanner_peek(void *state, float agg15[160], float agg16[160], int32_t counts[2]) {
  scanner_t *s = (scanner_t *)state;
  memcpy(agg15, s->agg15, sizeof(s->agg15));
  memcpy(agg16, s->agg16, sizeof(s->agg16));
  counts[0] = s->count15;
  counts[1] = s->count16;
}

/* ---- frame scoring ---------------------------------------------------------------------------------------
 * dmz_card_rect_for_screen + dmz_set_roi_for_scoring (dmz.cpp:136-181): the card-sized (or card/3-sized) rectangle
 * centred in the frame, scaled by min(w/640, h/480) in float with (int) truncation when the frame is not 640x480.
 * cvSetImageROI then clips it to the image. */
void orc_scoring_rect(int w, int h, int use_full_image, int rect[4]) {
  int cw = use_full_image ? 428 : 428 / 3, ch = use_full_image ? 270 : 270 / 3;
  int rw, rh, x, y, x1, y1;
  if (w == 0 || h == 0) {
    rect[0] = rect[1] = rect[2] = rect[3] = 0;
    return;
  }
  if (w == 640 && h == 480) {
    rw = cw, rh = ch;
  } else {
    float rx = ((float)w) / ((float)640), ry = ((float)h) / ((float)480);
    float ratio = rx < ry ? rx : ry;
    rw = (int)(cw * ratio);
    rh = (int)(ch * ratio);
  }
  x = (w - rw) / 2, y = (h - rh) / 2;
  x1 = x + rw, y1 = y + rh; /* cvSetImageROI clipping */
  if (x < 0) x = 0;
  if (y < 0) y = 0;
  if (x1 > w) x1 = w;
  if (y1 > h) y1 = h;
  rect[0] = x, rect[1] = y, rect[2] = x1 - x, rect[3] = y1 - y;
}

/* dmz_focus_score_for_image (dmz.cpp:114-127): llcv_sobel3_dx_dy's scalar path (cv/sobel.cpp:556-628: the diagonal
 * kernel [1 0 -1; 0 0 0; -1 0 1] with rows and columns clamped at the ROI), then llcv_stddev_of_abs_c
 * (cv/stats.cpp:87-92: cvAbs, cvAvgSdv). */
float orc_focus_score(const uint8_t *img, int ystep, int w, int h, int use_full_image) {
  int rc[4], x, y;
  int16_t *g;
  double sd;
  orc_scoring_rect(w, h, use_full_image, rc);
  if (rc[2] < 2 || rc[3] < 1) return 0.0f;
  g = (int16_t *)malloc((size_t)rc[2] * rc[3] * sizeof(int16_t));
  for (y = 0; y < rc[3]; y++) {
    const uint8_t *r1 = img + (size_t)(rc[1] + (y == 0 ? 0 : y - 1)) * ystep + rc[0];
    const uint8_t *r2 = img + (size_t)(rc[1] + (y == rc[3] - 1 ? y : y + 1)) * ystep + rc[0];
    for (x = 0; x < rc[2]; x++) {
      int xl = x == 0 ? 0 : x - 1, xr = x == rc[2] - 1 ? x : x + 1;
      int v = r1[xl] - r1[xr] - r2[xl] + r2[xr];
      g[(size_t)y * rc[2] + x] = (int16_t)(v < 0 ? -v : v); /* cvAbs; |v| <= 510 */
    }
  }
  orc_mean_stddev_s16(g, rc[2] * (int)sizeof(int16_t), rc[2], rc[3], NULL, &sd);
  free(g);
  return (float)sd;
}

/* dmz_brightness_score_for_image (dmz.cpp:129-134): (float)cvAvg over the scoring ROI */
float orc_brightness_score(const uint8_t *img, int ystep, int w, int h, int use_full_image) {
  int rc[4];
  orc_scoring_rect(w, h, use_full_image, rc);
  if (rc[2] < 1 || rc[3] < 1) return 0.0f;
  return (float)orc_mean_u8(img + (size_t)rc[1] * ystep + rc[0], ystep, rc[2], rc[3]);
}

/* ---- pixel formats either side of the path (widening rows) --------------------------------------- */

/* llcv_YCbCr2RGB_u8_c (cv/convert.cpp:449-504, behind dmz_YCbCr_to_RGB dmz.cpp:58-64): three same-sized planes ->
 * interleaved R, G, B (, A = 255) bytes.  Fixed point, 14 fractional bits, round-half-up by the arithmetic shift
 * (which rounds negative products towards minus infinity, as gcc's >> on int does), then saturation. */
void orc_ycbcr_to_rgb(const uint8_t *y, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep, int w, int h, int channels,
                      uint8_t *dst, int dstep) {
  int r, c;
  for (r = 0; r < h; r++)
    for (c = 0; c < w; c++) {
      const int py = y[(size_t)r * ystep + c];
      const int scb = (int)cb[(size_t)r * cstep + c] - 128, scr = (int)cr[(size_t)r * cstep + c] - 128; /* int8_t range */
      const int pb = py + ((scb * 29049 + (1 << 13)) >> 14);
      const int pg = py + ((scb * -5636 + scr * -11698 + (1 << 13)) >> 14);
      const int pr = py + ((scr * 22987 + (1 << 13)) >> 14);
      uint8_t *o = dst + (size_t)r * dstep + (size_t)c * channels;
      o[0] = (uint8_t)(pr < 0 ? 0 : pr > 255 ? 255 : pr);
      o[1] = (uint8_t)(pg < 0 ? 0 : pg > 255 ? 255 : pg);
      o[2] = (uint8_t)(pb < 0 ? 0 : pb > 255 ? 255 : pb);
      if (channels == 4) o[3] = 0xff;
    }
}

/* dmz_deinterleave_RGBA_to_R (dmz.cpp:66-109): dest[i] = source[4 i] for the `size` pixels (the scalar path's groups of
 * eight and of four cover every index when size % 4 == 0, which the reference assumes). */
void orc_rgba_to_r(const uint8_t *source, uint8_t *dest, size_t size) {
  size_t i;
  for (i = 0; i < size; i++) dest[i] = source[4 * i];
}

/* The three 3x3 stencils the Cython layer exposes (dmz.cpp:519-531), all with rows and columns clamped at the image:
 *   kind 0  llcv_scharr3_dx_abs (cv/sobel.cpp:706-799): t = |p(x+1) - p(x-1)| per row, then 3 t(y-1) + 10 t(y) + 3 t(y+1)
 *   kind 1  llcv_scharr3_dy_abs (cv/sobel.cpp:826-900): t = |p(y+1) - p(y-1)| per column, then 3 t(x-1) + 10 t(x) + 3 t(x+1)
 *   kind 2  llcv_sobel3_dx_dy   (cv/sobel.cpp:556-628): p(x-1,y-1) - p(x+1,y-1) - p(x-1,y+1) + p(x+1,y+1)
 * (the absolute value is taken BEFORE the smoothing pass: these are not |Scharr|). */
void orc_stencil3(const uint8_t *img, int step, int w, int h, int kind, int16_t *out) {
  int x, y;
  for (y = 0; y < h; y++) {
    const uint8_t *r0 = img + (size_t)(y == 0 ? 0 : y - 1) * step, *r1 = img + (size_t)y * step,
                  *r2 = img + (size_t)(y == h - 1 ? y : y + 1) * step;
    for (x = 0; x < w; x++) {
      const int xl = x == 0 ? 0 : x - 1, xr = x == w - 1 ? x : x + 1;
      int v;
      if (kind == 0)
        v = 3 * (abs(r0[xr] - r0[xl]) + abs(r2[xr] - r2[xl])) + 10 * abs(r1[xr] - r1[xl]);
      else if (kind == 1)
        v = 3 * (abs(r2[xl] - r0[xl]) + abs(r2[xr] - r0[xr])) + 10 * abs(r2[x] - r0[x]);
      else
        v = r0[xl] - r0[xr] - r2[xl] + r2[xr];
      out[(size_t)y * w + x] = (int16_t)v;
    }
  }
}

int orc_luhn(const uint8_t *d, int n) { /* dmz_olm.cpp dmz_passes_luhn_checksum */
  int even = 0, sum = 0, i;
  for (i = n - 1; i >= 0; i--) {
    int addend = d[i] * (1 << (even++ & 1));
    sum += addend % 10 + addend / 10;
  }
  return sum % 10 == 0;
}

int orc_card_type(const uint8_t *d, int n) { /* dmz_card_info_for_prefix_and_length(.., allow_incomplete=false).card_type */
  static const struct {
    int type, len, plen;
    long lo, hi;
  } t[] = {{5, 16, 4, 2221, 2720}, {6, 14, 3, 300, 305},  {6, 14, 3, 309, 309}, {2, 15, 2, 34, 34},   {3, 16, 4, 3528, 3589},
           {6, 14, 2, 36, 36},     {6, 14, 2, 38, 39},    {2, 15, 2, 37, 37},   {4, 16, 1, 4, 4},     {7, 16, 2, 50, 50},
           {5, 16, 2, 51, 55},     {7, 16, 2, 56, 59},    {6, 16, 4, 6011, 6011}, {7, 16, 2, 61, 61}, {6, 16, 2, 62, 62},
           {7, 16, 2, 63, 63},     {6, 16, 3, 644, 649},  {6, 16, 2, 65, 65},   {7, 16, 2, 66, 69},   {6, 16, 2, 88, 88}};
  int compatible = 0, type = 0 /* unrecognized */;
  size_t i;
  if (n <= 0) return 0;
  for (i = 0; i < sizeof(t) / sizeof(t[0]); i++) {
    int plen = t[i].plen, factor = 1, j;
    long prefix = 0;
    if (n != t[i].len) continue;
    while (plen > n) factor *= 10, plen--;
    for (j = 0; j < plen; j++) prefix = prefix * 10 + d[j];
    if (prefix >= t[i].lo / factor && prefix <= t[i].hi / factor) compatible++, type = t[i].type;
  }
  if (compatible == 1) return type;
  if (compatible > 1) return 1; /* ambiguous */
  return 0;
}

int orc_scanner_result(void *state, uint8_t digits[16], int32_t *n_numbers) {
  scanner_t *s = (scanner_t *)state;
  memset(digits, 0, 16);
  *n_numbers = 0;
  if (!s->complete) {
    int maxc = s->count15 > s->count16 ? s->count15 : s->count16;
    int minc = s->count15 < s->count16 ? s->count15 : s->count16;
    const float *agg;
    uint8_t num[16];
    int i, j, n, type;
    if (maxc - minc < 3) return 0;
    if (minc * 2 > maxc) return 0;
    if (s->count15 > s->count16) n = 15, agg = s->agg15;
    else n = 16, agg = s->agg16;
    *n_numbers = n;
    for (i = 0; i < n; i++) {
      const float *row = agg + i * 10;
      float mx = row[0], sum, stability;
      int arg = 0;
      for (j = 1; j < 10; j++)
        if (row[j] > mx) mx = row[j], arg = j;
      sum = eig_tree_sum(row, 0, 10);
      num[i] = (uint8_t)arg;
      digits[i] = (uint8_t)arg;
      stability = mx / sum;
      if (stability < 0.7f) return 0;
    }
    type = orc_card_type(num, n);
    if (type != 1 && type != 0 && orc_luhn(num, n)) {
      s->complete = 1;
      s->n_numbers = n;
      memcpy(s->digits, num, 16);
    }
  }
  if (s->complete) {
    memcpy(digits, s->digits, 16);
    *n_numbers = s->n_numbers;
    return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------ timing */
typedef struct {
  const uint8_t *frames, *cb, *cr;
  int w, h, lo, hi, orientation;
  orc_frame_record *recs;
} bench_job;

static void *bench_worker(void *arg) {
  bench_job *j = (bench_job *)arg;
  int i;
  for (i = j->lo; i < j->hi; i++)
    orc_process_frame(j->frames + (size_t)i * j->w * j->h, j->w, j->h, j->w, j->cb, j->cr, j->w / 2, j->orientation,
                      &j->recs[i], NULL);
  return NULL;
}

double orc_bench_frames(const uint8_t *frames, int n, int w, int h, const uint8_t *cb, const uint8_t *cr,
                        int orientation, int nthreads, orc_frame_record *recs) {
  pthread_t *th;
  bench_job *jobs;
  struct timespec t0, t1;
  int t;
  if (nthreads < 1) nthreads = 1;
  th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  jobs = (bench_job *)malloc(sizeof(bench_job) * nthreads);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (t = 0; t < nthreads; t++) {
    bench_job j = {frames, cb, cr, w, h, (int)((long)n * t / nthreads), (int)((long)n * (t + 1) / nthreads), orientation, recs};
    jobs[t] = j;
    pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

#define BT(x) orc_##x
#include "bench_taps.inc"
#undef BT
