// card.io-dmz_b200/csrc/exact.cu -- the bit-exact float / fixed-point stages.  Compiled with -fmad=false:
// every float and double operation below is a single IEEE-754 round-to-nearest op in exactly the order
// the reference's x86-64 SSE2 build performs it (SURVEY section 7 "hard parts").
//
//   geometry_kernel        find_line_in_detection_rects tail + corner intersections   dmz.cpp:346-438, geometry.cpp
//   householder_qr_solve8  llcv_calc_persp_transform = Eigen 3.2.4 householderQr().solve cv/warp.cpp:34-125
//   (the warp itself, cvWarpPerspective, lives in warp.cu)
//   vseg_select_kernel     best_segmentation_for_vseg_scores + gating                  scan/n_vseg.cpp:49-92, frame.cpp:38-47
//   hseg_kernel            best_n_hseg / best_n_hseg_constrained                        scan/n_hseg.cpp:39-152
//   scan_finish_kernel     number_score gate                                            scan/frame.cpp:63-64
//   finalize_records_kernel  flat per-frame record + card checksum
#include <float.h>

#include <stdlib.h>

#include "b200_internal.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Eigen 3.2.4 vectorised redux order (Core/Redux.h:192-246, Packet4f, SSE2 predux (a0+a2)+(a1+a3)),
// alignedStart == 0.  v(i) yields the i-th coefficient of the reduced expression.
// ------------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ float eig_redux_sum(F v, int n) {
  if (n == 0) return 0.0f;
  const int aligned2 = (n / 8) * 8, aligned = (n / 4) * 4;
  float res;
  if (aligned) {
    float p0[4], p1[4];
#pragma unroll
    for (int l = 0; l < 4; l++) p0[l] = v(l);
    if (aligned > 4) {
#pragma unroll
      for (int l = 0; l < 4; l++) p1[l] = v(4 + l);
      for (int i = 8; i < aligned2; i += 8) {
#pragma unroll
        for (int l = 0; l < 4; l++) {
          p0[l] = p0[l] + v(i + l);
          p1[l] = p1[l] + v(i + 4 + l);
        }
      }
#pragma unroll
      for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
      if (aligned > aligned2) {
#pragma unroll
        for (int l = 0; l < 4; l++) p0[l] = p0[l] + v(aligned2 + l);
      }
    }
    res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
    for (int i = aligned; i < n; i++) res = res + v(i);
  } else {
    res = v(0);
    for (int i = 1; i < n; i++) res = res + v(i);
  }
  return res;
}

// ------------------------------------------------------------------------------------------------
// W1: x = A.householderQr().solve(b) for the 8x8 float system (column-major a[i + 8 j]).
// Householder/HouseholderQR.h (unblocked; block size 8 covers the matrix), Householder.h,
// HouseholderSequence.h, products/TriangularSolverVector.h -- operation by operation.
// ------------------------------------------------------------------------------------------------
#define A_(i, j) a[(i) + 8 * (j)]

__device__ void householder_qr_solve8(float *a, const float *bvec, float *xout) {
  float hcoef[8], tmp[8], c[8];
  for (int k = 0; k < 8; k++) {
    const int rem_rows = 8 - k, n = rem_rows - 1;
    const float c0 = A_(k, k);
    float tail_sq = 0.0f, beta, tau;
    if (rem_rows != 1) tail_sq = eig_redux_sum([&](int i) { return A_(k + 1 + i, k) * A_(k + 1 + i, k); }, n);
    if (tail_sq == 0.0f) {
      tau = 0.0f;
      beta = c0;
      for (int i = 0; i < n; i++) A_(k + 1 + i, k) = 0.0f;
    } else {
      beta = sqrtf(c0 * c0 + tail_sq);
      if (c0 >= 0.0f) beta = -beta;
      const float denom = c0 - beta;
      for (int i = 0; i < n; i++) A_(k + 1 + i, k) = A_(k + 1 + i, k) / denom;
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    A_(k, k) = beta;
    if (rem_rows != 1) {
      for (int j = k + 1; j < 8; j++) tmp[j] = eig_redux_sum([&](int i) { return A_(k + 1 + i, k) * A_(k + 1 + i, j); }, n);
      for (int j = k + 1; j < 8; j++) tmp[j] = tmp[j] + A_(k, j);
      for (int j = k + 1; j < 8; j++) A_(k, j) = A_(k, j) - tau * tmp[j];
      for (int j = k + 1; j < 8; j++)
        for (int i = 0; i < n; i++) A_(k + 1 + i, j) = A_(k + 1 + i, j) - (A_(k + 1 + i, k) * tau) * tmp[j];
    }
  }
  for (int i = 0; i < 8; i++) c[i] = bvec[i];
  for (int k = 0; k < 8; k++) {
    const int n = 7 - k;
    const float tau = hcoef[k];
    if (n == 0) {
      c[k] = c[k] * (1.0f - tau);
    } else {
      float t = eig_redux_sum([&](int i) { return A_(k + 1 + i, k) * c[k + 1 + i]; }, n);
      t = t + c[k];
      c[k] = c[k] - tau * t;
      for (int i = 0; i < n; i++) c[k + 1 + i] = c[k + 1 + i] - (A_(k + 1 + i, k) * tau) * t;
    }
  }
  for (int k = 0; k < 8; k++) {
    const int i = 7 - k;
    c[i] = c[i] / A_(i, i);
    for (int j = 0; j < i; j++) c[j] = c[j] - c[i] * A_(j, i);
  }
  for (int i = 0; i < 8; i++) xout[i] = c[i];
}

__device__ void calc_persp_transform(const float *s, const float *d, float *M) {
  float a[64], b[8], x[8];
  for (int i = 0; i < 64; i++) a[i] = 0.0f;
  for (int i = 0; i < 4; i++) {
    const float sx = s[2 * i], sy = s[2 * i + 1], dx = d[2 * i], dy = d[2 * i + 1];
    A_(i, 0) = sx, A_(i, 1) = sy, A_(i, 2) = 1.0f;
    A_(i, 6) = -sx * dx;
    A_(i, 7) = -sy * dx;
    A_(i + 4, 3) = sx, A_(i + 4, 4) = sy, A_(i + 4, 5) = 1.0f;
    A_(i + 4, 6) = -sx * dy;
    A_(i + 4, 7) = -sy * dy;
    b[i] = dx;
    b[i + 4] = dy;
  }
  householder_qr_solve8(a, b, x);
  for (int i = 0; i < 8; i++) M[i] = x[i];
  M[8] = 1.0f;
}
#undef A_

// cv::invert of the 3x3 matrix promoted to double (closed-form adjugate path of lapack.cpp, 2.4.x)
__device__ void invert3x3(const float *Mf, double *t) {
  double m[9];
  for (int i = 0; i < 9; i++) m[i] = (double)Mf[i];
#define S(i, j) m[(i)*3 + (j)]
  double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
             S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
  if (d == 0.) {
    for (int i = 0; i < 9; i++) t[i] = 0.0;
    return;
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
}

// dmz_transform_card corner permutation (dmz.cpp:446-481) + homography + inverse
__device__ void corners_to_transform(const float *c /* tl, bl, tr, br */, int orientation, int upsample, FrameGeom *g) {
  const float *tl = c, *bl = c + 2, *tr = c + 4, *br = c + 6;
  const float *sp[4];
  switch (orientation) {
    case B200_ORIENT_PORTRAIT: sp[0] = bl, sp[1] = tl, sp[2] = br, sp[3] = tr; break;
    case B200_ORIENT_LANDSCAPE_LEFT: sp[0] = br, sp[1] = bl, sp[2] = tr, sp[3] = tl; break;
    case B200_ORIENT_PORTRAIT_UPSIDE_DOWN: sp[0] = tr, sp[1] = br, sp[2] = tl, sp[3] = bl; break;
    default: sp[0] = tl, sp[1] = tr, sp[2] = bl, sp[3] = br; break;
  }
  float src[8];
  const float dst[8] = {0.0f, 0.0f, 427.0f, 0.0f, 0.0f, 269.0f, 427.0f, 269.0f};
  for (int i = 0; i < 4; i++) {
    src[2 * i] = sp[i][0];
    src[2 * i + 1] = sp[i][1];
    if (upsample) {
      src[2 * i] = src[2 * i] / 2.0f;
      src[2 * i + 1] = src[2 * i + 1] / 2.0f;
    }
  }
  calc_persp_transform(src, dst, g->M);
  invert3x3(g->M, g->Minv);
}

// parametricIntersect (geometry.cpp:14-32): Eigen Matrix2f determinant / inverse (LU/Inverse.h:70-89)
__device__ bool parametric_intersect(float rho1, float c1, float s1, float rho2, float c2, float s2, float *x, float *y) {
  const float det = c1 * s2 - c2 * s1;
  if ((double)det < 1e-10) return false;
  const float invdet = 1.0f / det;
  const float i00 = s2 * invdet, i10 = -c2 * invdet, i01 = -s1 * invdet, i11 = c1 * invdet;
  *x = i00 * rho1 + i01 * rho2;
  *y = i10 * rho1 + i11 * rho2;
  return true;
}

__global__ void geometry_kernel(const __grid_constant__ GeomParams G, const b200_line *__restrict__ lines,
                                size_t plane_stride, int n, FrameGeom *__restrict__ geom) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  FrameGeom g;
  // dmz_edges slot order top, left, bottom, right; strips are stored top, bottom, left, right
  const int box_of_slot[4] = {0, 2, 1, 3};
  const float mult[3] = {1.0f, 2.0f, 2.0f};
  for (int slot = 0; slot < 4; slot++) {
    const int box = box_of_slot[slot];
    g.found[slot] = 0, g.rho[slot] = 0.0f, g.theta[slot] = 0.0f, g.n_idx[slot] = 0;
    for (int p = 0; p < G.n_planes && !g.found[slot]; p++) {
      const b200_line l = lines[(size_t)p * plane_stride + (size_t)f * 4 + box];
      if (l.found) {
        float rho = (float)((double)l.rho + G.delta_rho[p][box][l.n]);  // lineByShiftingOrigin
        rho = rho * mult[p];
        g.found[slot] = 1;
        g.rho[slot] = rho;
        g.theta[slot] = G.theta[box >= 2][l.n];
        g.n_idx[slot] = l.n;
      }
    }
  }
  g.all_found = 0;
  for (int i = 0; i < 8; i++) g.corners[i] = 0.0f;
  for (int i = 0; i < 9; i++) g.M[i] = 0.0f, g.Minv[i] = 0.0;
  g.pad = 0;
  if (g.found[0] && g.found[1] && g.found[2] && g.found[3]) {
    // slots: 0 top (horizontal), 1 left (vertical), 2 bottom, 3 right
    const float ct = G.cos_t[0][g.n_idx[0]], st = G.sin_t[0][g.n_idx[0]];
    const float cl = G.cos_t[1][g.n_idx[1]], sl = G.sin_t[1][g.n_idx[1]];
    const float cb = G.cos_t[0][g.n_idx[2]], sb = G.sin_t[0][g.n_idx[2]];
    const float cr = G.cos_t[1][g.n_idx[3]], sr = G.sin_t[1][g.n_idx[3]];
    float c[8];
    const bool tl = parametric_intersect(g.rho[0], ct, st, g.rho[1], cl, sl, &c[0], &c[1]);
    const bool bl = parametric_intersect(g.rho[2], cb, sb, g.rho[1], cl, sl, &c[2], &c[3]);
    const bool tr = parametric_intersect(g.rho[0], ct, st, g.rho[3], cr, sr, &c[4], &c[5]);
    const bool br = parametric_intersect(g.rho[2], cb, sb, g.rho[3], cr, sr, &c[6], &c[7]);
    if (tl && bl && tr && br) {
      for (int i = 0; i < 8; i++) g.corners[i] = c[i];
      g.all_found = 1;
      corners_to_transform(g.corners, G.orientation, 0, &g);
    }
  }
  geom[f] = g;
}

__global__ void corners_to_geom_kernel(const b200_corner_points *__restrict__ corners, const uint8_t *__restrict__ valid,
                                       int n, int orientation, int upsample, FrameGeom *__restrict__ geom) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  FrameGeom g;
  for (int i = 0; i < 4; i++) g.found[i] = 1, g.rho[i] = 0.0f, g.theta[i] = 0.0f, g.n_idx[i] = 0;
  const float *c = reinterpret_cast<const float *>(corners + f);
  for (int i = 0; i < 8; i++) g.corners[i] = c[i];
  g.all_found = valid ? (valid[f] != 0) : 1;
  g.pad = 0;
  for (int i = 0; i < 9; i++) g.M[i] = 0.0f, g.Minv[i] = 0.0;
  if (g.all_found) corners_to_transform(g.corners, orientation, upsample, &g);
  geom[f] = g;
}

__global__ void homography_only_kernel(const float *__restrict__ src, const float *__restrict__ dst, int n, float *__restrict__ M) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  float s[8], d[8], m[9];
  for (int i = 0; i < 8; i++) s[i] = src[(size_t)f * 8 + i], d[i] = dst[(size_t)f * 8 + i];
  calc_persp_transform(s, d, m);
  for (int i = 0; i < 9; i++) M[(size_t)f * 9 + i] = m[i];
}

// ------------------------------------------------------------------------------------------------
// V0 tail: best_segmentation_for_vseg_scores (n_vseg.cpp:49-92), sequential float running sums.
// pass 0: coarse best -> window for the fine rows.  pass 1: final vseg + gating (frame.cpp:38-47).
// ------------------------------------------------------------------------------------------------
__device__ void best_segmentation(const float *__restrict__ vp /* [270][2] visa, amex */, float *score, int *ptype, int *yoff) {
  float vsum = 0.0f, asum = 0.0f, best = 0.0f;
  int bt = 0, by = 0;
  for (int y = 0; y < 270; y++) {
    vsum = vsum + vp[2 * y];
    asum = asum + vp[2 * y + 1];
    if (y >= 26) {
      if (vsum > best) best = vsum, bt = 1, by = y - 26;
      if (asum > best) best = asum, bt = 2, by = y - 26;
      // ring_buffer[(y + 1) % 27] holds the score of row y - 26
      vsum = vsum - vp[2 * (y - 26)];
      asum = asum - vp[2 * (y - 26) + 1];
    }
  }
  *score = best, *ptype = bt, *yoff = by;
}

// One thread per frame runs the sequential scan, but out of SHARED memory: a CTA first stages the score rows of its 32
// frames with coalesced 128-bit loads (row stride 541 floats: the 32 lanes of the scanning warp hit 32 different banks),
// and clears the 32 scan records cooperatively.  (One thread per frame reading global memory directly touched a new
// cache line per lane per step: ~1 ms per pass and 100k frames; staged: a few tens of microseconds.)
constexpr int kSelFrames = 32, kSelThreads = 128, kSelStride = 541;

__global__ void __launch_bounds__(kSelThreads)
vseg_select_kernel(const float *__restrict__ vprob, const uint8_t *__restrict__ gate, int n, int pass, b200_scan *__restrict__ scans,
                   uint16_t *__restrict__ coarse_y) {
  extern __shared__ float sel_rows[];  // [kSelFrames][kSelStride]
  const int f0 = blockIdx.x * kSelFrames, tid = threadIdx.x;
  const int nf = min(kSelFrames, n - f0);
  {
    const float4 *src = reinterpret_cast<const float4 *>(vprob + (size_t)f0 * 540);  // 2160-byte rows: 16-byte aligned
    for (int i = tid; i < nf * 135; i += kSelThreads) {
      const int fr = i / 135, k = (i - fr * 135) * 4;
      const float4 v = __ldg(src + i);
      float *d = sel_rows + fr * kSelStride + k;
      d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    }
    if (pass == 1) {  // sizeof(b200_scan) == 720: padding bytes must be 0 too
      unsigned int *words = reinterpret_cast<unsigned int *>(scans + f0);
      for (int i = tid; i < nf * 180; i += kSelThreads) words[i] = 0u;
    }
  }
  __syncthreads();
  if (tid >= nf) return;
  const int f = f0 + tid;
  b200_scan *sc = scans + f;
  if (gate && !gate[f]) {
    if (pass == 0) {
      sc->vseg.y_offset = 0xFFFF;  // no fine rows
      if (coarse_y) coarse_y[f] = 0xFFFF;
    }
    return;
  }
  float score;
  int pt, yo;
  best_segmentation(sel_rows + tid * kSelStride, &score, &pt, &yo);
  if (pass == 0) {
    sc->vseg.y_offset = (uint16_t)yo;
    if (coarse_y) coarse_y[f] = (uint16_t)yo;  // kept for the lazy warp: pass 1 overwrites the record
    return;
  }
  const uint8_t pat[3][19] = {{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                              {1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1},
                              {1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 0}};
  const uint8_t pat_len[3] = {0, 19, 17}, num_len[3] = {0, 16, 15};
  sc->vseg.score = score;
  sc->vseg.y_offset = (uint16_t)yo;
  sc->vseg.pattern_type = (uint8_t)pt;
  for (int i = 0; i < 19; i++) sc->vseg.number_pattern[i] = pat[pt][i];
  sc->vseg.number_pattern_length = pat_len[pt];
  sc->vseg.number_length = num_len[pt];
  if (yo < (B200_CARD_H - 27) / 2) sc->upside_down = 1;  // kFlipVSegYOffsetCutoff
  else sc->usable = score > 15.0f;                       // kMinVSegScore
}

// ------------------------------------------------------------------------------------------------
// H0 / H1: best_n_hseg.  One CTA per frame.
// ------------------------------------------------------------------------------------------------
constexpr int kHsegThreads = 256;
constexpr int kMaxWidths = 8;   // widths per pass: the four (min, max, step) triples of n_hseg.cpp:110-147 give at most 6
constexpr int kMaxCands = 1024;
constexpr int kPatPad = 192;  // zeros in front of a pattern row: a valid candidate's offset is <= 428 - 17 * 16.3 = 151

__device__ __constant__ float kNumberGradSumPattern[19] = {
    0.26228655f, 0.30289554f, 0.34632607f, 0.38725636f, 0.42745813f, 0.45875135f, 0.46498017f,
    0.45258447f, 0.43045216f, 0.42430462f, 0.44796554f, 0.47726529f, 0.48471646f, 0.46457738f,
    0.42799847f, 0.38851183f, 0.33966308f, 0.28802608f, 0.25377602f};

struct HsegPass {
  int nwidths;
  float width[kMaxWidths];
  int omin[kMaxWidths], count[kMaxWidths], start[kMaxWidths + 1];
  int ostep;
};

__global__ void __launch_bounds__(kHsegThreads)
hseg_kernel(const uint8_t *__restrict__ cards, int n, b200_scan *__restrict__ scans) {
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  b200_scan *sc = scans + f;
  if (!sc->usable) return;  // upside_down / vseg gate (block-uniform)

  // the source strip is dead once the column sums exist, two barriers before the first pattern row is written: the two
  // share storage (36 -> 25 KB per CTA: eight resident CTAs per SM instead of six for this barrier-bound kernel)
  constexpr int kPatRow = kPatPad + B200_CARD_W + 4;
  __shared__ __align__(16) uint8_t s_buf[kMaxWidths * kPatRow * 4 > 27 * B200_CARD_W ? kMaxWidths * kPatRow * 4 : 27 * B200_CARD_W];
  uint8_t (*s_strip)[B200_CARD_W] = reinterpret_cast<uint8_t (*)[B200_CARD_W]>(s_buf);
  float (*s_patw)[kPatRow] = reinterpret_cast<float (*)[kPatRow]>(s_buf);
  __shared__ float s_g[B200_CARD_W];
  __shared__ int s_isum[B200_CARD_W];
  __shared__ int s_mn, s_mx;
  __shared__ HsegPass s_pass;
  __shared__ unsigned long long s_red[kHsegThreads / 32];
  __shared__ unsigned int s_best_words[12];  // a b200_hseg whose padding bytes stay zero
  b200_hseg &s_best = *reinterpret_cast<b200_hseg *>(s_best_words);
  __shared__ uint8_t s_pat[19];
  __shared__ int s_npl;
  // The pattern of a candidate (width w, offset o) is the pattern of (w, 0) shifted right by o (every digit centre is
  // o + lrintf(index * w)), so each pass builds ONE 428-float pattern row per width -- kPatPad zeros in front -- and a
  // candidate reads it at [i - o].
  __shared__ short s_last_center[kMaxWidths];           // centre of the last digit at offset 0: validity is o + centre + 19 < 428

  const int y_off = sc->vseg.y_offset;
  const uint8_t *card = cards + (size_t)f * (B200_CARD_W * B200_CARD_H) + (size_t)y_off * B200_CARD_W;
  for (int i = tid; i < 27 * B200_CARD_W / 4; i += kHsegThreads)
    reinterpret_cast<unsigned int *>(&s_strip[0][0])[i] = __ldg(reinterpret_cast<const unsigned int *>(card) + i);
  if (tid < 19) s_pat[tid] = sc->vseg.number_pattern[tid];
  if (tid == 0) {
    s_npl = sc->vseg.number_pattern_length;
    s_mn = 0x7fffffff, s_mx = 0;
    for (int i = 0; i < 12; i++) s_best_words[i] = 0u;
    s_best.n_offsets = sc->vseg.number_length;
    s_best.score = 428.0f;
    s_best.number_width = 0.0f;
  }
  __syncthreads();

  // llcv_morph_grad3_2d_cross_u8 on the isolated 428 x 27 strip (cv/morph.cpp:177-255), then cvReduce column sums
  int cmn = 0x7fffffff, cmx = 0;
  for (int x = tid; x < B200_CARD_W; x += kHsegThreads) {
    const int xl = x > 0 ? x - 1 : x, xr = x < B200_CARD_W - 1 ? x + 1 : x;
    int acc = 0;
    for (int y = 0; y < 27; y++) {
      const int yu = y > 0 ? y - 1 : y, yd = y < 26 ? y + 1 : y;
      const int a = s_strip[yu][x], b = s_strip[y][xl], c = s_strip[y][x], d = s_strip[y][xr], e = s_strip[yd][x];
      const int mx = max(a, max(b, max(c, max(d, e)))), mn = min(a, min(b, min(c, min(d, e))));
      acc += mx - mn;
    }
    s_isum[x] = acc;
    cmn = min(cmn, acc), cmx = max(cmx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cmn = min(cmn, __shfl_xor_sync(0xffffffffu, cmn, o));
    cmx = max(cmx, __shfl_xor_sync(0xffffffffu, cmx, o));
  }
  if ((tid & 31) == 0) atomicMin(&s_mn, cmn), atomicMax(&s_mx, cmx);  // one pair of atomics per warp
  __syncthreads();
  {
    // cvNormalize(MINMAX, 0, 1): scale / shift in double, applied as float multiply then float add
    const double smin = (double)(float)s_mn, smax = (double)(float)s_mx;
    const double scale = (1.0 - 0.0) * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    const float fs = (float)scale, fb = (float)shift;
    for (int x = tid; x < B200_CARD_W; x += kHsegThreads) s_g[x] = (float)s_isum[x] * fs + fb;
  }
  __syncthreads();

  // thread 0, after a pass: fold the pass's best candidate (s_red) into s_best.  Runs right before the same thread
  // sets up the next pass, so the two single-thread sections share one barrier interval.
  auto update_best = [&]() {
      unsigned long long k = ~0ull;
      for (int i = 0; i < kHsegThreads / 32; i++) k = s_red[i] < k ? s_red[i] : k;
      if (k != ~0ull) {
        const float score = __uint_as_float((unsigned)(k >> 32));
        const int c = (int)(k & 0xFFFFFFFFu);
        if (score < s_best.score) {
          int wi = 0;
          while (c >= s_pass.start[wi + 1]) wi++;
          const float width = s_pass.width[wi];
          const int offset = s_pass.omin[wi] + (c - s_pass.start[wi]) * s_pass.ostep;
          int oi = 0;
          for (int i = 0; i < 16; i++) s_best.offsets[i] = 0;
          for (int pi = 0; pi < s_npl; pi++)
            if (s_pat[pi]) {
              if (oi < 16) s_best.offsets[oi] = (uint16_t)(offset + __float2int_rn((float)pi * width));
              oi++;
            }
          s_best.score = score;
          s_best.number_width = width;
          s_best.pattern_offset = (uint16_t)offset;
        }
      }
  };
  for (int pass = 0; pass < 4; pass++) {
    if (tid == 0) {
      if (pass > 0) update_best();
      float wmin, wmax, wstep;
      unsigned omin, omax, ostep;
      const b200_hseg &b = s_best;
      if (pass == 0) {
        wmin = 17.1f, wmax = 19.7f, wstep = 0.5f, omin = 0, omax = 0xFFFF, ostep = 10;
      } else if (pass == 1) {
        wmin = b.number_width - 0.5f, wmax = b.number_width + 0.5f, wstep = 0.2f;
        omin = b.pattern_offset < 10 ? 0 : b.pattern_offset - 10, omax = (uint16_t)(b.pattern_offset + 10), ostep = 1;
      } else if (pass == 2) {
        wmin = b.number_width - 0.2f, wmax = b.number_width + 0.2f, wstep = 0.1f;
        omin = b.pattern_offset < 3 ? 0 : b.pattern_offset - 3, omax = (uint16_t)(b.pattern_offset + 3), ostep = 1;
      } else {
        wmin = b.number_width - 0.1f, wmax = b.number_width + 0.1f, wstep = 0.05f;
        omin = b.pattern_offset < 3 ? 0 : b.pattern_offset - 3, omax = (uint16_t)(b.pattern_offset + 3), ostep = 1;
      }
      int nw = 0, total = 0;
      for (float width = wmin; width < wmax && nw < kMaxWidths; width += wstep) {
        const float pattern_width = (float)s_npl * width;
        unsigned pom = omax;
        const unsigned maximum = (uint16_t)(428 - (long)__float2int_rn(pattern_width));  // lrintf: current rounding mode = nearest-even
        if (pom == 0xFFFFu || pom > maximum) pom = maximum;
        int cnt = pom > omin ? (int)((pom - omin + ostep - 1) / ostep) : 0;
        if (total + cnt > kMaxCands) cnt = kMaxCands - total;
        s_pass.width[nw] = width;
        s_pass.omin[nw] = (int)omin;
        s_pass.count[nw] = cnt;
        s_pass.start[nw] = total;
        total += cnt;
        nw++;
      }
      s_pass.start[nw] = total;
      s_pass.nwidths = nw;
      s_pass.ostep = (int)ostep;
    }
    __syncthreads();
    const int total = s_pass.start[s_pass.nwidths];
    // pattern rows of this pass: warp wi builds width wi (zeros, then the 19-tap template at every digit centre in digit
    // order -- a later digit overwrites the overlapping tail of the previous one, as the reference's sequential copy does)
    for (int wi = tid >> 5; wi < s_pass.nwidths; wi += kHsegThreads / 32) {
      const int lane = tid & 31;
      float *row = s_patw[wi];
      for (int i = lane; i < kPatPad + B200_CARD_W + 4; i += 32) row[i] = 0.0f;
      __syncwarp();
      const float width = s_pass.width[wi];
      int nd = 0, last = 0;
      for (int pi = 0; pi < s_npl; pi++) {
        if (s_pat[pi]) {
          if (nd < 16) {
            const int center = __float2int_rn((float)pi * width);
            if (lane < 19 && center + lane < B200_CARD_W + 4) row[kPatPad + center + lane] = kNumberGradSumPattern[lane];
            last = center;
            __syncwarp();
          }
          nd++;
        }
      }
      if (lane == 0) s_last_center[wi] = (short)last;
    }
    __syncthreads();
    unsigned long long best_key = ~0ull;
    // eight lanes per candidate reproduce Eigen's two 4-lane packet accumulators over the 428 coefficients:
    // lane8 takes i = lane8, lane8 + 8, ...
    const int lane8 = tid & 7;
    const float *g = s_g + lane8;
    for (int c0 = 0; c0 < total; c0 += kHsegThreads / 8) {
      const int c = c0 + (tid >> 3);
      float score = 0.0f;
      bool valid = c < total;
      const float *pw = s_patw[0] + kPatPad;
      if (valid) {
        int wi = 0;
        while (c >= s_pass.start[wi + 1]) wi++;
        const int offset = s_pass.omin[wi] + (c - s_pass.start[wi]) * s_pass.ostep;
        // every centre must satisfy centre + 19 < 428 (n_hseg.cpp:57); centres increase, so the last one decides
        // (an offset beyond kPatPad fails that test anyway: the last centre alone is >= 16 * 16.3)
        valid = ((offset + s_last_center[wi]) & 0xFFFF) + 19 < 428 && offset <= kPatPad;
        pw = s_patw[wi] + (kPatPad - offset) + lane8;
      }
      // (valid is uniform across the eight lanes of a candidate; the shuffles stay outside any branch)
      float acc = 0.0f;
      if (valid) {
        acc = fabsf(g[0] - pw[0]);
#pragma unroll
        for (int k = 1; k < 53; k++) acc = acc + fabsf(g[8 * k] - pw[8 * k]);
      }
      const float hi = __shfl_down_sync(0xffffffffu, acc, 4, 8);
      float r = acc + hi;                                         // packet_res0 + packet_res1
      if (valid && lane8 < 4) r = r + fabsf(g[424] - pw[424]);    // the 107th packet
      const float r2 = __shfl_down_sync(0xffffffffu, r, 2, 8);
      const float t = r + r2;                                     // lanes 0,1: (r0 + r2), (r1 + r3)
      const float t1 = __shfl_down_sync(0xffffffffu, t, 1, 8);
      score = t + t1;                                             // lane 0: (r0 + r2) + (r1 + r3)
      if (valid && lane8 == 0) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)c;
        best_key = key < best_key ? key : best_key;
      }
    }
    // block argmin (smallest score, then smallest candidate index = first in the reference's loop order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, best_key, o);
      best_key = t < best_key ? t : best_key;
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = best_key;
    __syncthreads();
  }
  if (tid == 0) update_best();
  __syncthreads();
  if (tid < 12) reinterpret_cast<unsigned int *>(&sc->hseg)[tid] = s_best_words[tid];
}

// S0 tail: usable = (n_offsets - scores.sum()) < kMaxNumberScoreDelta, frame.cpp:63-64.  scores.sum() on the
// 16x10 fixed matrix is the vectorised linear redux.
__global__ void scan_finish_kernel(int n, b200_scan *__restrict__ scans) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  b200_scan *sc = scans + f;
  if (!sc->usable) return;
  const float *s = sc->scores;
  const float sum = eig_redux_sum([&](int i) { return s[i]; }, 160);
  const float number_score = (float)sc->hseg.n_offsets - sum;
  sc->usable = number_score < 3.0f;
}

// Flat per-frame record; the card checksum was accumulated by the warp kernel.  One thread per frame.
__global__ void finalize_records_kernel(const FrameGeom *__restrict__ geom, const b200_scan *__restrict__ scans,
                                        const unsigned int *__restrict__ card_check, int n, b200_frame_record *__restrict__ recs,
                                        uint8_t *__restrict__ needs_full, int cx0, int cy0, int cx1, int cy1) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const FrameGeom &g = geom[f];
  if (needs_full != nullptr) {
    // Cropped upload: the warp's taps lie in the convex hull of the four source corners (+1 px for the second tap,
    // +1 for the 1/32-px rounding).  If that box leaves the uploaded crop [cx0, cx1) x [cy0, cy1) the frame is
    // redone from a full upload by the host (b200_process_frames_batch).
    uint8_t flag = 0;
    if (g.all_found) {
      float lox = g.corners[0], hix = g.corners[0], loy = g.corners[1], hiy = g.corners[1];
      for (int i = 1; i < 4; i++) {
        lox = fminf(lox, g.corners[2 * i]), hix = fmaxf(hix, g.corners[2 * i]);
        loy = fminf(loy, g.corners[2 * i + 1]), hiy = fmaxf(hiy, g.corners[2 * i + 1]);
      }
      flag = !(lox - 2.0f >= (float)cx0 && hix + 3.0f < (float)cx1 && loy - 2.0f >= (float)cy0 && hiy + 3.0f < (float)cy1);
    }
    needs_full[f] = flag;
  }
  b200_frame_record *r = recs + f;
  for (int i = 0; i < 4; i++) r->found[i] = g.found[i], r->rho[i] = g.rho[i], r->theta[i] = g.theta[i];
  for (int i = 0; i < 8; i++) r->corners[i] = g.corners[i];
  r->all_found = g.all_found;
  const uint4 *src = reinterpret_cast<const uint4 *>(scans + f);  // sizeof(b200_scan) == 720 == 45 x 16
  uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(r) + 84);  // offsetof(b200_frame_record, scan), 4-aligned only
  unsigned int *dw = reinterpret_cast<unsigned int *>(dst);
  const unsigned int *sw_ = reinterpret_cast<const unsigned int *>(src);
  if (g.all_found) {
    for (int i = 0; i < 180; i++) dw[i] = sw_[i];
    r->card_check = card_check ? card_check[f] : 0u;  // lazy cards: no card, no checksum
  } else {
    for (int i = 0; i < 180; i++) dw[i] = 0u;
    r->card_check = 0u;
  }
}

__global__ void scan_gate_kernel(const FrameGeom *__restrict__ geom, const uint8_t *__restrict__ valid, int n, uint8_t *__restrict__ gate) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  uint8_t g = 1;
  if (geom) g = geom[f].all_found != 0;
  if (valid) g = g && valid[f];
  gate[f] = g;
}

}  // namespace

static int blocks_for(int n, int t) { return (n + t - 1) / t; }

int launch_geometry(const GeomParams &g, const b200_line *lines, size_t plane_stride, int n, FrameGeom *geom, cudaStream_t s) {
  geometry_kernel<<<blocks_for(n, 64), 64, 0, s>>>(g, lines, plane_stride, n, geom);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_homography_only(const float *src_pts, const float *dst_pts, int n, float *M, cudaStream_t s) {
  homography_only_kernel<<<blocks_for(n, 64), 64, 0, s>>>(src_pts, dst_pts, n, M);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_corners_to_geom(const b200_corner_points *corners, const uint8_t *valid, int n, int orientation, int upsample,
                           FrameGeom *geom, cudaStream_t s) {
  corners_to_geom_kernel<<<blocks_for(n, 64), 64, 0, s>>>(corners, valid, n, orientation, upsample, geom);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------
// Frame scoring: dmz_focus_score / dmz_brightness_score (dmz.cpp:114-195).
// One CTA per frame over the scoring rectangle (rows and columns clamp at the rectangle: the reference sets an
// image ROI first).  Integer sums are exact, so the only floating point is the final double arithmetic of
// cv::meanStdDev / cv::mean, done by one thread in the reference's operation order (no FMA in this file).
// ------------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 256;

__global__ void __launch_bounds__(kScoreThreads)
frame_scores_kernel(const uint8_t *__restrict__ frames, int row_stride, size_t frame_stride, int n, int rx, int ry, int rw,
                    int rh, float *__restrict__ focus, float *__restrict__ brightness) {
  __shared__ unsigned long long s_acc[kScoreThreads / 32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int f = blockIdx.x; f < n; f += gridDim.x) {
    const uint8_t *roi = frames + (size_t)f * frame_stride + (size_t)ry * row_stride + rx;
    unsigned int sum_px = 0, sum_abs = 0;
    unsigned long long sum_sq = 0;
    for (int y = warp; y < rh; y += kScoreThreads / 32) {
      const uint8_t *r0 = roi + (size_t)y * row_stride;
      const uint8_t *r1 = roi + (size_t)(y == 0 ? 0 : y - 1) * row_stride;
      const uint8_t *r2 = roi + (size_t)(y == rh - 1 ? y : y + 1) * row_stride;
      unsigned int sq = 0;
      for (int x = lane; x < rw; x += 32) {
        const int xl = x == 0 ? 0 : x - 1, xr = x == rw - 1 ? x : x + 1;
        const int v = (int)__ldg(r1 + xl) - (int)__ldg(r1 + xr) - (int)__ldg(r2 + xl) + (int)__ldg(r2 + xr);
        const unsigned int a = (unsigned int)abs(v);
        sum_px += __ldg(r0 + x);
        sum_abs += a;
        sq += a * a;  // <= 510^2 * ceil(rw / 32): fits 32 bits for any row a frame can have
      }
      sum_sq += sq;
    }
    unsigned long long v0 = sum_px, v1 = sum_abs, v2 = sum_sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    __syncthreads();  // previous frame's s_acc fully consumed
    if (lane == 0) s_acc[warp][0] = v0, s_acc[warp][1] = v1, s_acc[warp][2] = v2;
    __syncthreads();
    if (tid == 0) {
      unsigned long long t0 = 0, t1 = 0, t2 = 0;
      for (int w = 0; w < kScoreThreads / 32; w++) t0 += s_acc[w][0], t1 += s_acc[w][1], t2 += s_acc[w][2];
      const double scale = 1. / ((double)rw * (double)rh);
      if (brightness) brightness[f] = (float)((double)t0 * scale);  // cv::mean: s * (1. / total)
      if (focus) {  // cv::meanStdDev: s *= scale; sd = sqrt(max(sq * scale - s * s, 0.))
        const double m = (double)t1 * scale;
        const double var = (double)t2 * scale - m * m;
        focus[f] = (float)sqrt(var > 0. ? var : 0.);
      }
    }
  }
}

int launch_frame_scores(const uint8_t *frames, int row_stride, size_t frame_stride, int n, int rx, int ry, int rw, int rh,
                        float *focus, float *brightness, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = n < sms * 8 ? n : sms * 8;
  frame_scores_kernel<<<grid, kScoreThreads, 0, s>>>(frames, row_stride, frame_stride, n, rx, ry, rw, rh, focus, brightness);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------
// dmz_deinterleave_uint8_c2 (dmz.cpp:49-56 -> llcv_split_u8, cv/convert.cpp:74-76 = cvSplit): the CbCr plane of a
// biplanar camera frame into two dense planes.  Pure streaming: 16 interleaved bytes in, 8 + 8 bytes out per thread
// step when everything is 16-byte aligned, a byte loop otherwise.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
deinterleave_c2_kernel(const uint8_t *__restrict__ src, int row_stride, size_t frame_stride, int w, int h, size_t n,
                       uint8_t *__restrict__ ch1, uint8_t *__restrict__ ch2, int vec_ok) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  if (vec_ok) {  // w % 8 == 0, all bases and strides 16-byte (src) / 8-byte (dst) aligned
    const int wv = w >> 3;
    const size_t total = n * (size_t)h * wv;
    for (size_t i = tid; i < total; i += nthreads) {
      const size_t f = i / ((size_t)h * wv);
      const int rem = (int)(i - f * ((size_t)h * wv)), y = rem / wv, xv = rem - y * wv;
      const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src + f * frame_stride + (size_t)y * row_stride) + xv);
      uint2 a, b;  // even bytes -> channel 1, odd bytes -> channel 2
      a.x = __byte_perm(v.x, v.y, 0x6420), a.y = __byte_perm(v.z, v.w, 0x6420);
      b.x = __byte_perm(v.x, v.y, 0x7531), b.y = __byte_perm(v.z, v.w, 0x7531);
      const size_t o = (f * h + y) * (size_t)w + (size_t)xv * 8;
      __stcs(reinterpret_cast<uint2 *>(ch1 + o), a);
      __stcs(reinterpret_cast<uint2 *>(ch2 + o), b);
    }
  } else {
    const size_t total = n * (size_t)h * w;
    for (size_t i = tid; i < total; i += nthreads) {
      const size_t f = i / ((size_t)h * w);
      const int rem = (int)(i - f * ((size_t)h * w)), y = rem / w, x = rem - y * w;
      const uint8_t *p = src + f * frame_stride + (size_t)y * row_stride + 2 * x;
      ch1[i] = p[0], ch2[i] = p[1];
    }
  }
}

int launch_deinterleave_c2(const uint8_t *src, int row_stride, size_t frame_stride, int w, int h, int n, uint8_t *ch1, uint8_t *ch2,
                           cudaStream_t s) {
  const bool vec_ok = (w % 8 == 0) && (row_stride % 16 == 0) && (frame_stride % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                      ((uintptr_t)ch1 % 8 == 0) && ((uintptr_t)ch2 % 8 == 0);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t items = (size_t)n * h * (vec_ok ? w / 8 : w);
  size_t blocks = (items + 255) / 256;
  if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;  // grid-stride: 16 CTAs x 256 threads per SM
  if (blocks < 1) blocks = 1;
  deinterleave_c2_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, row_stride, frame_stride, w, h, (size_t)n, ch1, ch2, vec_ok ? 1 : 0);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_finalize_records(const FrameGeom *geom, const b200_scan *scans, const unsigned int *card_check, int n,
                            b200_frame_record *recs, cudaStream_t s, uint8_t *needs_full, int cx0, int cy0, int cx1, int cy1) {
  finalize_records_kernel<<<blocks_for(n, 128), 128, 0, s>>>(geom, scans, card_check, n, recs, needs_full, cx0, cy0, cx1, cy1);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// exported to nets.cu's launch_scan through these thin wrappers
int launch_scan_gate(const FrameGeom *geom, const uint8_t *valid, int n, uint8_t *gate, cudaStream_t s) {
  scan_gate_kernel<<<blocks_for(n, 256), 256, 0, s>>>(geom, valid, n, gate);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_vseg_select(const float *vprob, const uint8_t *gate, int n, int pass, b200_scan *scans, cudaStream_t s, uint16_t *coarse_y) {
  const size_t smem = sizeof(float) * kSelFrames * kSelStride;
  static PerDeviceOnce once;
  if (!once.ensure([smem] { return cudaFuncSetAttribute(vseg_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; }))
    return -1;
  vseg_select_kernel<<<blocks_for(n, kSelFrames), kSelThreads, smem, s>>>(vprob, gate, n, pass, scans, coarse_y);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_hseg(const uint8_t *cards, int n, b200_scan *scans, cudaStream_t s) {
  hseg_kernel<<<n, kHsegThreads, 0, s>>>(cards, n, scans);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_scan_finish(int n, b200_scan *scans, cudaStream_t s) {
  scan_finish_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, scans);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
