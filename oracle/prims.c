/*
 * oracle/prims.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See prims.h.
 *
 * Restates OpenCV 2.4.x (core + imgproc; third-party, un-vendored: headers pinned at 2.4.5 in
 * /root/reference/opencv2/core/version.hpp:50-53) for exactly the calls the reference's hot path
 * makes.  No code is taken from the reference; the algorithms are the library's published ones.
 * Build WITHOUT -march / -ffast-math / FMA contraction (x86-64 baseline SSE2): float results of the
 * 2.4.x library are mul-then-add.
 */
#include "prims.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* opencv2/core/types_c.h:305-334: cvRound == lrint under the default rounding mode (half-to-even). */
int orc_cv_round(double v) { return (int)lrint(v); }

/* opencv2/core/types_c.h:337-354. */
int orc_cv_floor(double v) {
  int i = (int)v;
  return i - (i > v);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int16_t sat_s16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* cv::getSobelKernels (imgproc/deriv.cpp, 2.4.x): binomial smoothing followed by `order` finite differences. */
void orc_sobel_kernel(int order, int ksize, int *out) {
  int ker[32];
  int i, j;
  memset(ker, 0, sizeof(ker));
  ker[0] = 1;
  for (i = 0; i < ksize - order - 1; i++) {
    int oldval = ker[0];
    for (j = 1; j <= ksize; j++) {
      int newval = ker[j] + ker[j - 1];
      ker[j - 1] = oldval;
      oldval = newval;
    }
  }
  for (i = 0; i < order; i++) {
    int oldval = -ker[0];
    for (j = 1; j <= ksize; j++) {
      int newval = ker[j - 1] - ker[j];
      ker[j - 1] = oldval;
      oldval = newval;
    }
  }
  for (i = 0; i < ksize; i++) out[i] = ker[i];
}

void orc_sobel_u8_s16(const uint8_t *src, int sstep, int w, int h, int16_t *dst, int dstep,
                      int xorder, int yorder, int ksize) {
  int kx[32], ky[32];
  int r = ksize / 2;
  int x, y, k;
  int *rowbuf = (int *)malloc(sizeof(int) * (size_t)w * (size_t)h);
  orc_sobel_kernel(xorder, ksize, kx);
  orc_sobel_kernel(yorder, ksize, ky);
  /* row pass: u8 -> int32 (FilterEngine row filter with integer kernel, bits = 0) */
  for (y = 0; y < h; y++) {
    const uint8_t *s = src + (size_t)y * sstep;
    int *d = rowbuf + (size_t)y * w;
    for (x = 0; x < w; x++) {
      int acc = 0;
      for (k = 0; k < ksize; k++) acc += kx[k] * (int)s[clampi(x + k - r, 0, w - 1)];
      d[x] = acc;
    }
  }
  /* column pass: int32 -> saturate_cast<short> */
  for (y = 0; y < h; y++) {
    int16_t *d = (int16_t *)((uint8_t *)dst + (size_t)y * dstep);
    for (x = 0; x < w; x++) {
      int acc = 0;
      for (k = 0; k < ksize; k++) acc += ky[k] * rowbuf[(size_t)clampi(y + k - r, 0, h - 1) * w + x];
      d[x] = sat_s16(acc);
    }
  }
  free(rowbuf);
}

double orc_sum_abs_s16(const int16_t *src, int sstep, int w, int h) {
  double total = 0.0;
  int x, y;
  for (y = 0; y < h; y++) {
    const int16_t *s = (const int16_t *)((const uint8_t *)src + (size_t)y * sstep);
    for (x = 0; x < w; x++) {
      int v = s[x];
      v = v < 0 ? -v : v;
      if (v > 32767) v = 32767; /* saturate_cast<short>(abs(-32768)) */
      total += v;
    }
  }
  return total;
}

void orc_morph_grad_cross3_u8(const uint8_t *src, int sstep, int w, int h, uint8_t *dst, int dstep) {
  int x, y;
  for (y = 0; y < h; y++) {
    const uint8_t *c = src + (size_t)y * sstep;
    const uint8_t *n = src + (size_t)(y > 0 ? y - 1 : y) * sstep;
    const uint8_t *s = src + (size_t)(y < h - 1 ? y + 1 : y) * sstep;
    uint8_t *d = dst + (size_t)y * dstep;
    for (x = 0; x < w; x++) {
      int xl = x > 0 ? x - 1 : x, xr = x < w - 1 ? x + 1 : x;
      int v[5] = {n[x], c[xl], c[x], c[xr], s[x]};
      int mx = v[0], mn = v[0], i;
      for (i = 1; i < 5; i++) {
        if (v[i] > mx) mx = v[i];
        if (v[i] < mn) mn = v[i];
      }
      d[x] = (uint8_t)(mx - mn);
    }
  }
}

void orc_resize_half_width_u8(const uint8_t *src, int sstep, int w, int h, uint8_t *dst, int dstep) {
  int x, y;
  for (y = 0; y < h; y++) {
    const uint8_t *s = src + (size_t)y * sstep;
    uint8_t *d = dst + (size_t)y * dstep;
    for (x = 0; x < w / 2; x++) d[x] = (uint8_t)((s[2 * x] + s[2 * x + 1] + 1) >> 1);
  }
}

void orc_convert_scale_u8_f32(const uint8_t *src, int sstep, int w, int h, float *dst, int dstep, float scale) {
  int x, y;
  for (y = 0; y < h; y++) {
    const uint8_t *s = src + (size_t)y * sstep;
    float *d = (float *)((uint8_t *)dst + (size_t)y * dstep);
    for (x = 0; x < w; x++) {
      volatile float p = (float)s[x] * scale; /* volatile: forbid contraction with the +0 shift */
      d[x] = p + 0.0f;
    }
  }
}

void orc_normalize_minmax_f32(float *data, int step, int w, int h) {
  double smin = DBL_MAX, smax = -DBL_MAX, scale, shift;
  float fscale, fshift;
  int x, y;
  for (y = 0; y < h; y++) {
    const float *s = (const float *)((const uint8_t *)data + (size_t)y * step);
    for (x = 0; x < w; x++) {
      if (s[x] < smin) smin = s[x];
      if (s[x] > smax) smax = s[x];
    }
  }
  scale = (1.0 - 0.0) * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0);
  shift = 0.0 - smin * scale;
  fscale = (float)scale;
  fshift = (float)shift;
  for (y = 0; y < h; y++) {
    float *s = (float *)((uint8_t *)data + (size_t)y * step);
    for (x = 0; x < w; x++) {
      volatile float p = s[x] * fscale;
      s[x] = p + fshift;
    }
  }
}

void orc_reduce_cols_sum_u8_f32(const uint8_t *src, int sstep, int w, int h, float *dst) {
  int x, y;
  for (x = 0; x < w; x++) {
    int acc = 0;
    for (y = 0; y < h; y++) acc += src[(size_t)y * sstep + x];
    dst[x] = (float)acc;
  }
}

int orc_invert3x3(const double m[9], double t[9]) {
#define S(i, j) m[(i)*3 + (j)]
  double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
             S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
  if (d == 0.) {
    memset(t, 0, 9 * sizeof(double));
    return 0;
  }
  d = 1. / d;
  t[0] = (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d;
  t[1] = (S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d;
  t[2] = (S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d;
  t[3] = (S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * d;
  t[4] = (S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * d;
  t[5] = (S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * d;
  t[6] = (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * d;
  t[7] = (S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * d;
  t[8] = (S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * d;
#undef S
  return 1;
}

/* cv::initInterTab2D(INTER_LINEAR, fixpt=true) from imgproc/imgwarp.cpp (2.4.x). */
static int16_t g_bilinear_tab[32 * 32 * 4];
static int g_bilinear_tab_ready = 0;

const int16_t *orc_bilinear_tab(void) {
  if (!g_bilinear_tab_ready) {
    float tab1d[32][2];
    const float scale = 1.f / 32;
    int i, j, k1, k2;
    memset(g_bilinear_tab, 0, sizeof(g_bilinear_tab));
    for (i = 0; i < 32; i++) {
      float x = i * scale;
      tab1d[i][0] = 1.f - x;
      tab1d[i][1] = x;
    }
    for (i = 0; i < 32; i++)
      for (j = 0; j < 32; j++) {
        int16_t *itab = g_bilinear_tab + (i * 32 + j) * 4;
        int isum = 0;
        for (k1 = 0; k1 < 2; k1++) {
          float vy = tab1d[i][k1];
          for (k2 = 0; k2 < 2; k2++) {
            float v = vy * tab1d[j][k2];
            int iv = orc_cv_round(v * 32768);
            itab[k1 * 2 + k2] = sat_s16(iv);
            isum += itab[k1 * 2 + k2];
          }
        }
        if (isum != 32768) {
          /* 2.4.x compensates on the window k1,k2 in [ksize/2, ksize/2+2) which for ksize=2 runs one
           * element past the 2x2 block into the (still zero) next entry; reproduce that scan. */
          int diff = isum - 32768;
          int Mk = 1 * 2 + 1, mk = 1 * 2 + 1;
          for (k1 = 1; k1 < 3; k1++)
            for (k2 = 1; k2 < 3; k2++) {
              int idx = k1 * 2 + k2;
              int16_t cur = (i * 32 + j) * 4 + idx < 32 * 32 * 4 ? itab[idx] : 0;
              if (cur < itab[mk]) mk = idx;
              else if (cur > itab[Mk]) Mk = idx;
            }
          if (diff < 0) itab[Mk] = (int16_t)(itab[Mk] - diff);
          else itab[mk] = (int16_t)(itab[mk] - diff);
        }
      }
    g_bilinear_tab_ready = 1;
  }
  return g_bilinear_tab;
}

void orc_warp_perspective_u8(const uint8_t *src, int sstep, int sw, int sh, uint8_t *dst, int dstep, int dw, int dh,
                             const float Mf[9]) {
  double M0[9], M[9];
  const int16_t *wtab = orc_bilinear_tab();
  int i, x, y, x1, y1;
  int bh0, bw0;
  for (i = 0; i < 9; i++) M0[i] = (double)Mf[i];
  orc_invert3x3(M0, M);

  /* cv::warpPerspective block traversal (BLOCK_SZ = 32): bh0 = min(16, h); bw0 = min(1024/bh0, w); bh0 = min(1024/bw0, h) */
  bh0 = 16 < dh ? 16 : dh;
  bw0 = (32 * 32 / bh0) < dw ? (32 * 32 / bh0) : dw;
  bh0 = (32 * 32 / bw0) < dh ? (32 * 32 / bw0) : dh;

  for (y = 0; y < dh; y += bh0)
    for (x = 0; x < dw; x += bw0) {
      int bw = bw0 < dw - x ? bw0 : dw - x;
      int bh = bh0 < dh - y ? bh0 : dh - y;
      for (y1 = 0; y1 < bh; y1++) {
        volatile double X0 = M[0] * x + M[1] * (y + y1) + M[2];
        volatile double Y0 = M[3] * x + M[4] * (y + y1) + M[5];
        volatile double W0 = M[6] * x + M[7] * (y + y1) + M[8];
        uint8_t *D = dst + (size_t)(y + y1) * dstep + x;
        for (x1 = 0; x1 < bw; x1++) {
          double W = W0 + M[6] * x1;
          double fX, fY;
          int X, Y, sx, sy, fxy;
          const int16_t *wt;
          W = W ? 32. / W : 0;
          fX = (X0 + M[0] * x1) * W;
          fY = (Y0 + M[3] * x1) * W;
          fX = fX < (double)INT_MIN ? (double)INT_MIN : (fX > (double)INT_MAX ? (double)INT_MAX : fX);
          fY = fY < (double)INT_MIN ? (double)INT_MIN : (fY > (double)INT_MAX ? (double)INT_MAX : fY);
          X = orc_cv_round(fX);
          Y = orc_cv_round(fY);
          sx = sat_s16(X >> 5);
          sy = sat_s16(Y >> 5);
          fxy = (Y & 31) * 32 + (X & 31);
          wt = wtab + fxy * 4;
          if ((unsigned)sx < (unsigned)(sw - 1 > 0 ? sw - 1 : 0) && (unsigned)sy < (unsigned)(sh - 1 > 0 ? sh - 1 : 0)) {
            const uint8_t *S = src + (size_t)sy * sstep + sx;
            D[x1] = sat_u8((S[0] * wt[0] + S[1] * wt[1] + S[sstep] * wt[2] + S[sstep + 1] * wt[3] + (1 << 14)) >> 15);
          } else if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
            D[x1] = 0;
          } else {
            int v0 = (sx >= 0 && sy >= 0 && sx < sw && sy < sh) ? src[(size_t)sy * sstep + sx] : 0;
            int v1 = (sx + 1 >= 0 && sy >= 0 && sx + 1 < sw && sy < sh) ? src[(size_t)sy * sstep + sx + 1] : 0;
            int v2 = (sx >= 0 && sy + 1 >= 0 && sx < sw && sy + 1 < sh) ? src[(size_t)(sy + 1) * sstep + sx] : 0;
            int v3 = (sx + 1 >= 0 && sy + 1 >= 0 && sx + 1 < sw && sy + 1 < sh) ? src[(size_t)(sy + 1) * sstep + sx + 1] : 0;
            D[x1] = sat_u8((v0 * wt[0] + v1 * wt[1] + v2 * wt[2] + v3 * wt[3] + (1 << 14)) >> 15);
          }
        }
      }
    }
}

void orc_bilateral_tables(int d, double sigma_color, double sigma_space, float *color_lut, float *space_w) {
  int radius = d / 2, i, j, k = 0;
  double gcc, gsc;
  if (sigma_color <= 0) sigma_color = 1;
  if (sigma_space <= 0) sigma_space = 1;
  gcc = -0.5 / (sigma_color * sigma_color);
  gsc = -0.5 / (sigma_space * sigma_space);
  if (radius < 1) radius = 1;
  for (i = 0; i < 256; i++) color_lut[i] = (float)exp(i * i * gcc);
  for (i = -radius; i <= radius; i++)
    for (j = -radius; j <= radius; j++, k++) {
      double r = sqrt((double)i * i + (double)j * j);
      space_w[k] = r > radius ? 0.0f : (float)exp(r * r * gsc);
    }
}

void orc_bilateral_u8(const uint8_t *src, int sstep, int w, int h, uint8_t *dst, int dstep, int d, double sigma_color,
                      double sigma_space) {
  float color_lut[256], space_w[81];
  int radius = d / 2 < 1 ? 1 : d / 2, dd = 2 * radius + 1, x, y, i, j;
  orc_bilateral_tables(d, sigma_color, sigma_space, color_lut, space_w);
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      float sum = 0, wsum = 0;
      int val0 = src[(size_t)y * sstep + x];
      for (i = -radius; i <= radius; i++)
        for (j = -radius; j <= radius; j++) {
          float sw = space_w[(i + radius) * dd + (j + radius)], wgt;
          int yy, xx, val;
          if (sqrt((double)i * i + (double)j * j) > radius) continue;
          yy = clampi(y + i, 0, h - 1);
          xx = clampi(x + j, 0, w - 1);
          val = src[(size_t)yy * sstep + xx];
          {
            volatile float wv = sw * color_lut[abs(val - val0)];
            volatile float pv = val * wv;
            wgt = wv;
            sum = sum + pv;
            wsum = wsum + wgt;
          }
        }
      dst[(size_t)y * dstep + x] = (uint8_t)orc_cv_round(sum / wsum);
    }
}

double orc_mean_u8(const uint8_t *src, int sstep, int w, int h) {
  /* cv::mean (core/stat.cpp, 2.4.x): integer block sums flushed into a double, then s * (1. / total) */
  double total = 0;
  int x, y;
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) total += src[(size_t)y * sstep + x];
  return total * (1. / ((double)w * h));
}

void orc_mean_stddev_s16(const int16_t *src, int sstep, int w, int h, double *mean, double *stddev) {
  /* cv::meanStdDev (core/stat.cpp, 2.4.x) for CV_16S: sum in int blocks, sqsum in double -- both exact for
   * anything this path feeds it -- then scale = 1. / total; s *= scale; sd = sqrt(max(sq * scale - s * s, 0)) */
  int64_t sum = 0, sq = 0;
  int x, y;
  for (y = 0; y < h; y++) {
    const int16_t *row = (const int16_t *)((const uint8_t *)src + (size_t)y * sstep);
    for (x = 0; x < w; x++) {
      sum += row[x];
      sq += (int64_t)row[x] * row[x];
    }
  }
  {
    double scale = 1. / ((double)w * h), s = (double)sum * scale, v = (double)sq * scale - s * s;
    if (mean) *mean = s;
    if (stddev) *stddev = sqrt(v > 0. ? v : 0.);
  }
}
