/*
 * tools/deck/deck_gen.h -- synthetic camera-frame generator (SURVEY section 8d "synthetic deck").
 *
 * BENCH / TEST SUPPORT, NOT PRODUCT CODE.  One implementation compiled twice: by gcc for the CPU
 * checkers (tools/deck/deck_cpu.c) and by nvcc for the GPU bench (tools/deck/deck_cuda.cu).  Only
 * integer arithmetic and IEEE double +,-,*,/ are used, with fused multiply-add disabled on both
 * sides (-ffp-contract=off / -fmad=false), so the two builds produce bit-identical frames
 * (tests/test_deck.py checks that).
 *
 * Frame k of a deck (seed S, W x H):
 *   - session = k / 8: all 8 frames of a session show the same card number (so scanner_result can
 *     complete); even sessions are Visa-like 4-4-4-4 (16 digits, prefix 4), odd sessions Amex-like
 *     4-6-5 (15 digits, prefix 34 / 37); the last digit is the Luhn check digit.
 *   - the card quad is the landscape guide rectangle (dmz_constants.h:7-27: corners (106,105),
 *     (533,105),(106,374),(533,374) at 640x480, scaled to the central 4:3 region otherwise) with every
 *     corner jittered by U(-jitter, +jitter) px, so each edge stays inside its detection strip.
 *   - card face: grey 175 + noise, 3 px darker rim, embossed digit row at y = 150 (tools/deck/glyphs.h).
 *   - background: grey 60 + noise.  Noise is a counter-based hash of (seed, frame, x, y).
 */
#ifndef DECK_GEN_H
#define DECK_GEN_H

#include <stdint.h>

#ifdef __CUDACC__
#define DECK_HD __device__ static inline
#define DECK_CONST static __device__ __constant__ const
#else
#define DECK_HD static inline
#define DECK_CONST static const
#endif

#include "glyphs.h"

typedef struct {
  double hinv[9];  /* frame (x,y,1) -> card (u,v,w), card in 428x270 pixel units */
  double quad[8];  /* card corners in the frame: tl, tr, bl, br (x,y) */
  uint8_t digits[16];
  int32_t n_digits;   /* 16 or 15 */
  int32_t row_y;      /* top of the digit row in card space */
  int32_t slot_x0;    /* x of the first digit slot */
  uint32_t frame;     /* frame index */
  uint64_t seed;
} deck_params;

DECK_HD uint64_t deck_mix(uint64_t x) { /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

DECK_HD uint64_t deck_hash3(uint64_t seed, uint64_t a, uint64_t b) { return deck_mix(deck_mix(seed ^ (a * 0xD1342543DE82EF95ull)) + b); }

/* uniform in [-1, 1) with 24 bits */
DECK_HD double deck_u11(uint64_t h) { return (double)(int64_t)((h >> 40) & 0xFFFFFF) / 8388608.0 - 1.0; }

/* Solve the 8x8 system for the projective map taking src[i] -> dst[i] (Gaussian elimination with
 * partial pivoting, plain double ops in a fixed order). out = row-major 3x3 with out[8] = 1. */
DECK_HD void deck_homography(const double src[8], const double dst[8], double out[9]) {
  double a[8][9];
  int i, j, k;
  for (i = 0; i < 4; i++) {
    double sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
    a[i][0] = sx, a[i][1] = sy, a[i][2] = 1, a[i][3] = 0, a[i][4] = 0, a[i][5] = 0, a[i][6] = -sx * dx, a[i][7] = -sy * dx, a[i][8] = dx;
    a[i + 4][0] = 0, a[i + 4][1] = 0, a[i + 4][2] = 0, a[i + 4][3] = sx, a[i + 4][4] = sy, a[i + 4][5] = 1;
    a[i + 4][6] = -sx * dy, a[i + 4][7] = -sy * dy, a[i + 4][8] = dy;
  }
  for (k = 0; k < 8; k++) {
    int piv = k;
    double best = a[k][k] < 0 ? -a[k][k] : a[k][k];
    for (i = k + 1; i < 8; i++) {
      double v = a[i][k] < 0 ? -a[i][k] : a[i][k];
      if (v > best) best = v, piv = i;
    }
    if (piv != k)
      for (j = 0; j < 9; j++) {
        double t = a[k][j];
        a[k][j] = a[piv][j];
        a[piv][j] = t;
      }
    for (i = k + 1; i < 8; i++) {
      double f = a[i][k] / a[k][k];
      for (j = k; j < 9; j++) a[i][j] = a[i][j] - f * a[k][j];
    }
  }
  for (k = 7; k >= 0; k--) {
    double s = a[k][8];
    for (j = k + 1; j < 8; j++) s = s - a[k][j] * out[j];
    out[k] = s / a[k][k];
  }
  out[8] = 1.0;
}

DECK_HD void deck_frame_params(uint64_t seed, uint32_t frame, int W, int H, double jitter, deck_params *p) {
  uint32_t session = frame / 8;
  int amex = (int)(session & 1);
  /* guide rectangle in the central 4:3 region (dmz.cpp:286-288), scaled from the 640x480 constants */
  int width43 = (H * 4) / 3, left = (W - width43) / 2;
  double sc = (double)H / 480.0;
  double gx0 = left + 106.0 * sc, gx1 = left + 533.0 * sc, gy0 = 105.0 * sc, gy1 = 374.0 * sc;
  double card[8] = {0, 0, 427, 0, 0, 269, 427, 269};
  double base[8];
  int i, even, sum;
  base[0] = gx0, base[1] = gy0, base[2] = gx1, base[3] = gy0, base[4] = gx0, base[5] = gy1, base[6] = gx1, base[7] = gy1;
  p->seed = seed;
  p->frame = frame;
  for (i = 0; i < 8; i++) p->quad[i] = base[i] + jitter * sc * deck_u11(deck_hash3(seed, frame, 100 + i));
  deck_homography(p->quad, card, p->hinv);
  p->n_digits = amex ? 15 : 16;
  for (i = 0; i < 16; i++) p->digits[i] = (uint8_t)(deck_hash3(seed, 0x5E55 + session, i) % 10);
  if (amex) {
    p->digits[0] = 3;
    p->digits[1] = (deck_hash3(seed, 0x5E55 + session, 77) & 1) ? 4 : 7;
    p->digits[15] = 0;
  } else {
    p->digits[0] = 4;
  }
  /* Luhn check digit (dmz_olm.cpp dmz_passes_luhn_checksum): choose the last digit so the sum is 0 mod 10 */
  sum = 0;
  even = 1; /* position n-2 is doubled */
  for (i = p->n_digits - 2; i >= 0; i--) {
    int addend = p->digits[i] * (1 << (even & 1));
    sum += addend % 10 + addend / 10;
    even++;
  }
  p->digits[p->n_digits - 1] = (uint8_t)((10 - sum % 10) % 10);
  p->row_y = 150 + (int)(deck_hash3(seed, 0x5E55 + session, 55) % 7) - 3;
  p->slot_x0 = 33 + (int)(deck_hash3(seed, 0x5E55 + session, 56) % 5) - 2;
}

/* digit index -> slot index in the 19-slot strip (scan/n_vseg.cpp:28-29 number patterns) */
DECK_HD int deck_slot_of_digit(int n_digits, int d) {
  if (n_digits == 16) return d + d / 4;             /* 1111 0 1111 0 1111 0 1111 */
  return d < 4 ? d : (d < 10 ? d + 1 : d + 2);      /* 1111 0 111111 0 11111 */
}

/* signed grey-level delta of the digit row at card position (cu, cv), bilinear in the glyph table */
DECK_HD double deck_digit_delta(const deck_params *p, double cu, double cv) {
  double gy = cv - (double)p->row_y, gx;
  int d, slot, ix, iy;
  double fx, fy, v00, v01, v10, v11;
  if (gy < 0.0 || gy >= 26.0) return 0.0;
  gx = cu - (double)p->slot_x0;
  if (gx < 0.0) return 0.0;
  slot = (int)(gx / 19.0);
  if (slot >= 19) return 0.0;
  gx = gx - 19.0 * slot;
  /* which digit occupies this slot? */
  d = -1;
  {
    int k;
    for (k = 0; k < p->n_digits; k++)
      if (deck_slot_of_digit(p->n_digits, k) == slot) d = k;
  }
  if (d < 0 || gx >= 18.0) return 0.0;
  ix = (int)gx;
  iy = (int)gy;
  fx = gx - ix;
  fy = gy - iy;
  v00 = deck_glyphs[p->digits[d]][iy][ix];
  v01 = deck_glyphs[p->digits[d]][iy][ix + 1];
  v10 = deck_glyphs[p->digits[d]][iy + 1][ix];
  v11 = deck_glyphs[p->digits[d]][iy + 1][ix + 1];
  return (v00 * (1.0 - fx) + v01 * fx) * (1.0 - fy) + (v10 * (1.0 - fx) + v11 * fx) * fy;
}

/* approx N(0, sigma) from four hash bytes (Irwin-Hall), integer */
DECK_HD int deck_noise(uint64_t h, int sigma_x16) {
  int s = (int)(h & 255) + (int)((h >> 8) & 255) + (int)((h >> 16) & 255) + (int)((h >> 24) & 255) - 510; /* sd ~147.8 */
  return (s * sigma_x16) / (148 * 16);
}

DECK_HD uint8_t deck_pixel(const deck_params *p, int x, int y) {
  const double *m = p->hinv;
  double w = m[6] * x + m[7] * y + m[8];
  double cu = (m[0] * x + m[1] * y + m[2]) / w;
  double cv = (m[3] * x + m[4] * y + m[5]) / w;
  uint64_t h = deck_hash3(p->seed ^ 0xC0FFEEull, p->frame, ((uint64_t)(uint32_t)y << 32) | (uint32_t)x);
  int v;
  if (cu >= 0.0 && cu <= 427.0 && cv >= 0.0 && cv <= 269.0) {
    double face = 175.0;
    if (cu < 3.0 || cu > 424.0 || cv < 3.0 || cv > 266.0) face = 120.0; /* rim */
    face = face + deck_digit_delta(p, cu, cv);
    v = (int)(face + 0.5) + deck_noise(h, 6 * 16);
  } else {
    v = 60 + deck_noise(h, 8 * 16);
  }
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

#endif
