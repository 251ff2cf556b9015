// card.io-dmz_b200/csrc/b200_tables.cpp -- host-side constant tables of the detect stage.
//
// Everything in the reference's edge detection that touches libm (sinf/cosf/tanf/atanf/cos/sqrt) or
// float rounding of constants takes only a handful of distinct values per (resolution, orientation):
// 10 Hough angles per line orientation and 4 strip origins per plane.  They are evaluated here ONCE on
// the host, with the same C library and float/double expressions as the reference, and handed to the
// kernels as constants; the device code then needs integer arithmetic and IEEE +,-,*,/ only.
// Compiled with -ffp-contract=off (no FMA), x86-64 baseline.
//
//   detection_boxes_for_sample   dmz.cpp:279-341
//   best_line_for_sample         dmz.cpp:224-271   (theta window, threshold)
//   llcv_hough tables            cv/hough.cpp:98-124
//   lineByShiftingOrigin         geometry.cpp:34-43
//   parametricIntersect          geometry.cpp:14-32 (cosf / sinf of the 10 + 10 possible thetas)
#include <float.h>
#include <math.h>
#include <string.h>

#include "b200_internal.h"

#define CV_PI 3.1415926535897932384626433832795

namespace {

struct Rect {
  int x, y, w, h;
};

Rect inset(Rect r, int hi, int vi) { return Rect{r.x + hi, r.y + vi, r.w - 2 * hi, r.h - 2 * vi}; }

// dmz_constants.h:16-27 (integer division inside the macros is intentional)
const float kPortraitVerticalPercentInset = (float)((640 - 270) / 2) / (float)640;
const float kPortraitHorizontalPercentInset = (float)((480 - 428) / 2) / (float)480;
const float kLandscapeVerticalPercentInset = (float)((480 - 270) / 2) / (float)480;
const float kLandscapeHorizontalPercentInset = (float)((640 - 428) / 2) / (float)640;

void detection_boxes(int img_w, int img_h, int orientation, Rect boxes[4]) {
  int inset_v = 0, slop_v = 0, inset_h = 0, slop_h = 0;
  int width = (img_h * 4) / 3;  // central 4:3 region
  int left_margin = (img_w - width) / 2;
  switch (orientation) {
    case B200_ORIENT_PORTRAIT:
    case B200_ORIENT_PORTRAIT_UPSIDE_DOWN:
      inset_v = (int)roundf(kPortraitHorizontalPercentInset * img_h);
      slop_v = (int)roundf(0.03f * img_h);
      inset_h = (int)roundf(kPortraitVerticalPercentInset * width);
      slop_h = (int)roundf(0.03f * width);
      break;
    case B200_ORIENT_LANDSCAPE_RIGHT:
    case B200_ORIENT_LANDSCAPE_LEFT:
      inset_v = (int)roundf(kLandscapeVerticalPercentInset * img_h);
      slop_v = (int)roundf(0.03f * img_h);
      inset_h = (int)roundf(kLandscapeHorizontalPercentInset * width);
      slop_h = (int)roundf(0.03f * width);
      break;
    default: break;
  }
  Rect image_rect{left_margin, 0, width - 1, img_h - 1};
  Rect outer = inset(image_rect, inset_h - slop_h, inset_v - slop_v);
  Rect inner = inset(image_rect, inset_h + slop_h, inset_v + slop_v);
  boxes[0] = Rect{inner.x, outer.y, inner.w, 2 * slop_v};            // top
  boxes[1] = Rect{inner.x, inner.y + inner.h, inner.w, 2 * slop_v};  // bottom
  boxes[2] = Rect{outer.x, inner.y, 2 * slop_h, inner.h};            // left
  boxes[3] = Rect{inner.x + inner.w, inner.y, 2 * slop_h, inner.h};  // right
}

int cv_round(double v) { return (int)lrint(v); }

const float kHorizontalAngle = (float)(CV_PI / 2.0f);
const float kVerticalAngle = (float)CV_PI;
const float kMaxAngleDeviationAllowed = (float)(5.0f * (CV_PI / 180.0f));

void hough_tables(int vertical, int tab_sin[], int tab_cos[], float theta_out[], float *slope_a, float *slope_b) {
  float base_angle = vertical ? kVerticalAngle : kHorizontalAngle;
  float theta_min = base_angle - kMaxAngleDeviationAllowed;
  float theta_max = base_angle + kMaxAngleDeviationAllowed;
  float theta = (float)CV_PI / 180.0f, irho = 1 / 1.0f, ang;
  float gat = 10;  // kHoughGradientAngleThreshold
  int numangle = cv_round((theta_max - theta_min) / theta), n;
  if (numangle != B200_NUMANGLE) numangle = B200_NUMANGLE;  // 9.999998f rounds to 10 on every IEEE platform
  for (ang = theta_min, n = 0; n < numangle; ang += theta, n++) {
    tab_sin[n] = (int)floorf(1024 * sinf(ang) * irho);
    tab_cos[n] = (int)floorf(1024 * cosf(ang) * irho);
    theta_out[n] = n * theta + theta_min;  // line.angle = n * theta + theta_min, hough.cpp:190
  }
  if (vertical) {
    *slope_a = tanf((float)(CV_PI * (180 - gat) / 180.0f));
    *slope_b = tanf((float)(CV_PI * (180 + gat) / 180.0f));
  } else {
    *slope_a = tanf((float)(CV_PI * (90 - gat) / 180.0f));
    *slope_b = tanf((float)(CV_PI * (90 + gat) / 180.0f));
  }
}

double delta_rho_for(float theta, int x_off, int y_off) {  // geometry.cpp:34-43, everything but the final add
  double offset_angle = x_off == 0 ? CV_PI / 2.0f : (double)atanf((float)y_off / (float)x_off);
  double delta_angle = theta - offset_angle + CV_PI / 2.0f;
  double offset_magnitude = sqrt((double)(x_off * x_off + y_off * y_off));
  return offset_magnitude * cos(CV_PI / 2 - delta_angle);
}

}  // namespace

void b200_build_detect_params(int width, int height, int orientation, int block_threads, DetectParams *out) {
  Rect boxes[4];
  memset(out, 0, sizeof(*out));
  detection_boxes(width, height, orientation, boxes);
  for (int s = 0; s < 4; s++) {
    StripDesc &d = out->strip[s];
    d.x = boxes[s].x, d.y = boxes[s].y, d.w = boxes[s].w, d.h = boxes[s].h;
    d.vertical = s >= 2;
    d.numrho = cv_round(((d.w + d.h) * 2 + 1) / 1.0f);
    d.half = (d.numrho - 1) / 2;
    d.threshold = (d.w > d.h ? d.w : d.h) / 6;
    hough_tables(d.vertical, d.tab_sin, d.tab_cos, d.theta, &d.slope_a, &d.slope_b);
    // compact accumulator: r = ((j cos + i sin) >> 10) + half is monotone in j and in i, so its extremes
    // over the strip are attained at the four corner pixels
    int base = 0;
    for (int n = 0; n < B200_NUMANGLE; n++) {
      int lo = 1 << 30, hi = -(1 << 30);
      for (int c = 0; c < 4; c++) {
        int j = (c & 1) ? d.w - 1 : 0, i = (c & 2) ? d.h - 1 : 0;
        int r = ((j * d.tab_cos[n] + i * d.tab_sin[n]) >> 10) + d.half;
        lo = r < lo ? r : lo;
        hi = r > hi ? r : hi;
      }
      d.rlo[n] = lo;
      d.rcount[n] = hi - lo + 1;
      d.cell_base[n] = base;
      base += d.rcount[n];
    }
    d.ncells = base;
    d.nchunks = d.w > 0 ? block_threads / d.w : 1;
    if (d.nchunks < 1) d.nchunks = 1;
    if (d.nchunks > d.h) d.nchunks = d.h > 0 ? d.h : 1;
    d.chunk_rows = (d.h + d.nchunks - 1) / d.nchunks;
    d.nchunks = d.chunk_rows > 0 ? (d.h + d.chunk_rows - 1) / d.chunk_rows : 1;
  }
}

void b200_build_geom_params(int width, int height, int orientation, int n_planes, GeomParams *out) {
  memset(out, 0, sizeof(*out));
  out->orientation = orientation;
  out->n_planes = n_planes;
  for (int v = 0; v < 2; v++) {
    int ts[B200_NUMANGLE], tc[B200_NUMANGLE];
    float sa, sb;
    hough_tables(v, ts, tc, out->theta[v], &sa, &sb);
    for (int n = 0; n < B200_NUMANGLE; n++) {
      out->cos_t[v][n] = cosf(out->theta[v][n]);
      out->sin_t[v][n] = sinf(out->theta[v][n]);
    }
  }
  for (int p = 0; p < 3; p++) {
    Rect boxes[4];
    int pw = p == 0 ? width : width / 2, ph = p == 0 ? height : height / 2;
    detection_boxes(pw, ph, orientation, boxes);
    for (int s = 0; s < 4; s++)
      for (int n = 0; n < B200_NUMANGLE; n++)
        out->delta_rho[p][s][n] = delta_rho_for(out->theta[s >= 2][n], boxes[s].x, boxes[s].y);
  }
}

// prepare_image_for_cat's bilateral filter (scan/expiry_categorize.cpp:52-60): cvSmooth(CV_BILATERAL, 3, 3,
// space_sigma, color_sigma) == cv::bilateralFilter(d = 3, sigmaColor = space_sigma, sigmaSpace = color_sigma)
// -- the reference's variable names are swapped relative to OpenCV's parameter order.  Colour LUT and the
// five in-circle spatial weights (mask order N, W, C, E, S) as imgproc/smooth.cpp builds them: (float)exp(double).
// llcv_norm_convert_1d_u8_to_f32 (cv/convert.cpp:380-383) = cvConvertScale(1/255) then cvNormalize(MINMAX, 0, 1): the float
// scale and shift cv::normalize derives (in double) from the row's min and max.  A row's 8-bit values only enter through
// their min mn and max mx, so the pair is tabulated for every (mn, mx) here -- the same IEEE operations as the reference,
// on the host -- and the row kernel needs no double-precision arithmetic.
void b200_build_minmax_norm_table(float *table) {
  const float k255 = 1.0f / 255.0f;
  for (int mn = 0; mn < 256; mn++)
    for (int mx = 0; mx < 256; mx++) {
      const float fmn = (float)mn * k255, fmx = (float)mx * k255;
      const double smin = (double)fmn, smax = (double)fmx;
      const double scale = (smax - smin > DBL_EPSILON) ? 1. / (smax - smin) : 0.0;
      const double shift = 0.0 - smin * scale;
      table[(mn * 256 + mx) * 2 + 0] = (float)scale;
      table[(mn * 256 + mx) * 2 + 1] = (float)shift;
    }
}

// Tensor-core form of the vseg hidden layer (vseg_mma.cu).  modelm_befe75da blob: W1 50 x 204 row-major, b1 50, W2 3 x 50, b2 3.
//   W1[u][k] = cu * Q[u][k],  Q = llround(W1 / S_u * 2^27)  (|Q| <= 2^27),  Q = ((q0 * 128 + q1) * 128 + q2) * 128 + q3 with
//   balanced digits q1..q3 in [-64, 63] and q0 in [-64, 64]: four s8 operand matrices, K-major in 16-byte chunks:
//   digit t, K chunk c, unit u at ((t * 14 + c) * 64 + u) * 16.
//   (s, d0) per (mn, mx): the reference's x_k = fl(fl(fl(v_k * k255) * scale) + shift) is, before its three roundings,
//   (v_k - mn) * s + d0 with s = k255 * scale and d0 = shift + mn * s -- evaluated here in double from the same float
//   scale / shift cv::normalize produces (b200_build_minmax_norm_table).
void b200_build_vseg_mma_tables(const float *blob, int8_t *wq, VsegUnit *units, float *sd) {
  memset(wq, 0, (size_t)4 * 14 * 64 * 16);
  memset(units, 0, sizeof(VsegUnit) * 64);
  for (int u = 0; u < 50; u++) {
    const float *w = blob + (size_t)u * 204;
    double smax = 0.0, sum = 0.0;
    for (int k = 0; k < 204; k++) {
      smax = fabs((double)w[k]) > smax ? fabs((double)w[k]) : smax;
      sum += (double)w[k];
    }
    if (smax == 0.0) smax = 1.0;
    for (int k = 0; k < 204; k++) {
      long long q = llround((double)w[k] / smax * 134217728.0);  // 2^27
      int digit[4];
      for (int t = 3; t >= 1; t--) {
        long long d = ((q + 64) % 128 + 128) % 128 - 64;  // balanced remainder in [-64, 63]
        digit[t] = (int)d;
        q = (q - d) / 128;
      }
      digit[0] = (int)q;  // |q| <= 64
      for (int t = 0; t < 4; t++) wq[(((size_t)t * 14 + k / 16) * 64 + u) * 16 + k % 16] = (int8_t)digit[t];
    }
    units[u].cu = (float)(smax / 134217728.0);
    units[u].sumw = (float)sum;
    units[u].b1 = blob[10200 + u];
    units[u].w20 = blob[10250 + u], units[u].w21 = blob[10250 + 50 + u], units[u].w22 = blob[10250 + 100 + u];
  }
  const float k255 = 1.0f / 255.0f;
  for (int mn = 0; mn < 256; mn++)
    for (int mx = 0; mx < 256; mx++) {
      const float fmn = (float)mn * k255, fmx = (float)mx * k255;
      const double smin = (double)fmn, smax = (double)fmx;
      const double scale = (smax - smin > DBL_EPSILON) ? 1. / (smax - smin) : 0.0;
      const float fs = (float)scale, fb = (float)(0.0 - smin * scale);
      const double s = (double)k255 * (double)fs;
      sd[(mn * 256 + mx) * 2 + 0] = (float)s;
      sd[(mn * 256 + mx) * 2 + 1] = (float)((double)fb + (double)mn * s);
    }
}

// float -> IEEE half, round to nearest even (subnormals kept, overflow -> inf)
static uint16_t half_bits_rn(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | (x > 0x7F800000u ? 0x7E00u : 0x7C00u));
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);  // rounds to >= 65520: inf
  if (x < 0x38800000u) {                                      // below the smallest normal half: subnormal
    if (x < 0x33000000u) return (uint16_t)sign;               // < 2^-25: zero
    const int shift = 126 - (int)(x >> 23);                   // 14 .. 24
    const uint32_t mant = (x & 0x7FFFFFu) | 0x800000u;
    uint32_t h = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (h & 1u))) h++;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((x >> 13) - (112u << 10));
  const uint32_t rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
  return (uint16_t)(sign | h);
}
static float half_bits_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
  float out;
  if (e == 0) {
    out = (float)m * 5.9604644775390625e-08f;  // 2^-24
    return sign ? -out : out;
  }
  const uint32_t x = sign | ((e + 112u) << 23) | (m << 13);
  memcpy(&out, &x, 4);
  return out;
}

// Tensor-core form of the digit CNNs (categorize_mma.cu).  modelc blob: conv W 8 x 9, conv b 8, hidden W 32 x 320, hidden b 32,
// logistic W 10 x 32, logistic b 10.
//   conv: per (model, kernel) w = Q / F with integer |Q| <= 2^20 and 255 * sum |Q| < 2^31 (so that sum_taps Q q fits 32 bits for
//   any bytes q); Q = (d0 << 14) + (d1 << 7) + d2 with balanced digits.  Operand column n = kernel * 9 + (pr * 3 + pc) (the conv
//   position inside the 3 x 3 pool window), row k = wy * 5 + wx (byte of the cell's 5 x 5 window): tap (wy - pr, wx - pc);
//   the three digit matrices are stacked along N (column 80 j + n), so one N = 240 MMA fills the three accumulators.
//   convf: (1/255 as float, in double) / F per (model, kernel), then the biases (added after the pool maximum).
//   hidden: W[u][kernel * 40 + cell] as fp16 hi + fp16 lo, operand chunk = cell, row = unit, element = kernel.
void b200_build_cnn_mma_tables(const float *const blobs[3], int8_t *convb, float *convf, uint16_t *hidb) {
  memset(convb, 0, (size_t)3 * 3 * 2 * 80 * 16);
  const double k255 = (double)(1.0f / 255.0f);
  for (int m = 0; m < 3; m++) {
    const float *b = blobs[m];
    for (int k = 0; k < 8; k++) {
      const float *w = b + k * 9;
      double smax = 0.0;
      for (int t = 0; t < 9; t++) smax = fabs((double)w[t]) > smax ? fabs((double)w[t]) : smax;
      if (smax == 0.0) smax = 1.0;
      double F = 1048576.0 / smax;  // 2^20 / max |w|
      long long Q[9];
      for (;;) {
        long long sum = 0;
        for (int t = 0; t < 9; t++) {
          Q[t] = llround((double)w[t] * F);
          sum += Q[t] < 0 ? -Q[t] : Q[t];
        }
        if (255 * sum < 2147483647LL - 64 * 1024 * 1024) break;
        F *= 0.5;
      }
      convf[m * 8 + k] = (float)(k255 / F);
      convf[24 + m * 8 + k] = b[72 + k];
      for (int t = 0; t < 9; t++) {
        long long q = Q[t];
        int digit[3];
        for (int j = 2; j >= 1; j--) {
          const long long d = ((q + 64) % 128 + 128) % 128 - 64;
          digit[j] = (int)d;
          q = (q - d) / 128;
        }
        digit[0] = (int)q;
        const int ti = t / 3, tj = t % 3;  // tap row / column
        for (int pr = 0; pr < 3; pr++)
          for (int pc = 0; pc < 3; pc++) {
            const int n = k * 9 + pr * 3 + pc, kk = (pr + ti) * 5 + (pc + tj);
            for (int j = 0; j < 3; j++) convb[((((size_t)m * 2 + kk / 16) * 240) + 80 * j + n) * 16 + kk % 16] = (int8_t)digit[j];
          }
      }
    }
    const float *hw = b + 80;  // [32][320]
    for (int c = 0; c < 40; c++)
      for (int u = 0; u < 32; u++)
        for (int k = 0; k < 8; k++) {
          const float v = hw[(size_t)u * 320 + k * 40 + c];
          const uint16_t hi = half_bits_rn(v);
          const uint16_t lo = half_bits_rn(v - half_bits_to_float(hi));
          const size_t at = (((size_t)c * 32) + u) * 8 + k;
          hidb[((size_t)m * 2 + 0) * (40 * 32 * 8) + at] = hi;
          hidb[((size_t)m * 2 + 1) * (40 * 32 * 8) + at] = lo;
        }
  }
}

// Layer-2 weights of the expiry CNN as fp16 operand slices (expiry_mma.cu): slice s = taps 2s, 2s + 1 (tap = i * 5 + j),
// part hi / lo, 14 chunks (tap slot * 7 + k) of 48 rows (output map f; rows 40 .. 47 zero) x 8 halfs (input map 8k + e).
// Blob: conv2 kernels at 1300 + f * 1250 + map * 25 + tap.
void b200_build_expiry_c2_halfs(const float *blob, uint16_t *out) {
  memset(out, 0, (size_t)13 * 2 * 14 * 48 * 8 * sizeof(uint16_t));
  for (int s = 0; s < 13; s++)
    for (int ts = 0; ts < 2; ts++) {
      const int tap = 2 * s + ts;
      if (tap >= 25) continue;
      for (int k = 0; k < 7; k++)
        for (int f = 0; f < 40; f++)
          for (int e = 0; e < 8; e++) {
            const int m = 8 * k + e;
            if (m >= 50) continue;
            const float w = blob[1300 + (size_t)f * 1250 + m * 25 + tap];
            const uint16_t hi = half_bits_rn(w), lo = half_bits_rn(w - half_bits_to_float(hi));
            const size_t at = (((size_t)(ts * 7 + k)) * 48 + f) * 8 + e;
            out[((size_t)s * 2 + 0) * (14 * 48 * 8) + at] = hi;
            out[((size_t)s * 2 + 1) * (14 * 48 * 8) + at] = lo;
          }
    }
}

void b200_build_bilateral_tables(float *color256, float *space5) {
  const int aperture = 3;
  const double sigma_color = (aperture / 2.0 - 1) * 0.3 + 0.8;  // the reference's "space_sigma"
  const double sigma_space = (aperture - 1) / 3.0;              // the reference's "color_sigma"
  const double gcc = -0.5 / (sigma_color * sigma_color), gsc = -0.5 / (sigma_space * sigma_space);
  for (int i = 0; i < 256; i++) color256[i] = (float)exp(i * i * gcc);
  const double r2[5] = {1.0, 1.0, 0.0, 1.0, 1.0};
  for (int k = 0; k < 5; k++) space5[k] = (float)exp(r2[k] * gsc);
}

// dmz_set_roi_for_scoring + dmz_card_rect_for_screen (dmz.cpp:136-181): card-sized (or card / 3) rectangle centred in
// the frame, scaled by min(w / 640, h / 480) in float with truncation unless the frame is 640x480; then clipped to
// the image as cvSetImageROI does.
void b200_scoring_rect(int w, int h, int use_full_image, int rect[4]) {
  const int card_w = use_full_image ? 428 : 428 / 3, card_h = use_full_image ? 270 : 270 / 3;
  int rw = card_w, rh = card_h;
  if (!(w == 640 && h == 480)) {
    const float ratio_w = (float)w / (float)640, ratio_h = (float)h / (float)480;
    const float ratio = ratio_w < ratio_h ? ratio_w : ratio_h;
    rw = (int)(card_w * ratio);
    rh = (int)(card_h * ratio);
  }
  int x0 = (w - rw) / 2, y0 = (h - rh) / 2, x1 = x0 + rw, y1 = y0 + rh;
  x0 = x0 < 0 ? 0 : x0, y0 = y0 < 0 ? 0 : y0;
  x1 = x1 > w ? w : x1, y1 = y1 > h ? h : y1;
  rect[0] = x0, rect[1] = y0, rect[2] = x1 - x0, rect[3] = y1 - y0;
}
