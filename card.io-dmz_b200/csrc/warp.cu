// card.io-dmz_b200/csrc/warp.cu -- W2: the fixed-point perspective warp, TMA-tiled.  Compiled with -fmad=false.
//
//   cvWarpPerspective(src, dst, M, CV_INTER_LINEAR + CV_WARP_FILL_OUTLIERS, 0)         cv/warp.cpp:153-166
//
// cv::warpPerspective walks the destination in 64 x 16 blocks and evaluates X0 = M0*x_block + M1*y + M2 once per block
// row, then (X0 + M0*x1) * (32 / W) per pixel; the coordinates are rounded (half-to-even) to 1/32 px and the four taps are
// blended with 15-bit integer weights, out-of-image taps being 0 (BORDER_CONSTANT).
//
// Work decomposition.  A CTA produces a SEGMENT (a run of four-pixel groups) of a SET OF ROWS of one card:
//   WARP_FULL    rows [b R, b R + R)                      the materialised card (dmz_transform_card's output)
//   WARP_COARSE  rows 4 (b R + i)                          the 68 rows the vseg coarse pass scores (n_vseg.cpp:127-137)
//   WARP_FINE    rows [y0 - 8, y0 + 35) around coarse y0   the rows the vseg fine pass scores (n_vseg.cpp:140-166)
//   WARP_STRIP   rows [y_offset, y_offset + 27)            only if the final number strip left the FINE window
// The last three are the lazy path of b200_process_frames_batch when the caller does not ask for the cards: only the
// ~111 rows that scan_card_image ever reads are warped (into their places in the card buffer), the other ~159 never exist.
//
// Source staging.  The source pixels a CTA can touch lie in the bounding box of the images of its destination
// rectangle's corners (a projective map sends the rectangle to a convex quadrilateral as long as W keeps its sign).
// One elected thread computes that box in double precision and fetches it with ONE cp.async.bulk.tensor (TMA) into a
// 256 x 64 (landscape) or 64 x 256 (portrait) shared-memory tile, completion on an mbarrier; elements outside the source
// plane are zero-filled by the TMA unit, which IS BORDER_CONSTANT 0 -- the tile path has no in-image test and no global
// address arithmetic: a tap is LDS.U8 [tile + ty * pitch + tx (+1, +pitch, +pitch + 1)].  While the copy is in flight the
// other threads set up their linear forms.  If the box does not fit the tile (a caller-supplied quad at another scale),
// W changes sign, or the plane cannot be described by a tensor map (unaligned base / strides), the CTA takes the
// gather path: the same arithmetic with byte loads through L1 and explicit bounds tests.
//
// Coordinates.  The reference computes W' = 32 / W (IEEE divide), fX = (X0 + M0 x1) W', X = cvRound(fX).  Here X is
// first computed the cheap way: fused multiply-adds for the three linear forms (stepped from row to row by addition),
// rcp.approx + ONE Newton step for 1/W (relative error ~1e-12), and a single FMA
//     t = fX * 2^14 + (1.5 * 2^52 + 2^30 + 2^13)
// whose low word u then holds round(fX * 2^14) + 2^30 + 2^13: X + 65536 = u >> 14 is fX rounded to nearest, (u >> 19) the
// source pixel + 2048, (u >> 14) & 31 the 1/32 fraction.  The fast value differs from the real-number one by < 1e-3 of
// u's unit, so X is the reference's unless fX * 2^14 sits within that distance of a rounding tie, i.e. unless the low
// 14 bits of u are 0 (the tie itself).  Low bits 0 OR 1 send the PIXEL (not the quad) to the exact reference sequence
// -- one call in ~0.1 % of the quads -- so the result is bit-identical for every pixel.
//
// Tried and dropped (B200, ms per 100 k frames, lazy rows / materialised card): persistent CTAs with a producer warp that plans
// and TMA-fetches one task ahead into a second tile buffer (tools/experiments/warp_persistent.cu.txt: bit-exact, all tests
// green) 11.0 / 26.4 with three CTAs per SM, 11.4 / 23.2 with two, against 9.7 / 18.9 for this one-task-per-CTA kernel: the
// kernel is issue-bound, four resident CTAs already cover each other's plan and TMA latency, and the second tile buffer
// costs a quarter of the resident consumer warps.
#include <cuda.h>
#include <float.h>
#include <stdlib.h>

#include "b200_internal.h"

namespace {

constexpr int kQuadsPerRow = B200_CARD_W / 4;  // 107 four-pixel groups per destination row
constexpr int kWarpThreads = 224;

struct WarpArgs {
  const uint8_t *src;   // buffer pixel (0, 0) of frame 0 (the whole plane, or the uploaded crop)
  int row_stride;
  size_t frame_stride;
  int bw, bh;           // buffer size in pixels
  int ox, oy;           // position of buffer pixel (0, 0) in frame coordinates (crop origin; 0, 0 for whole planes)
  const FrameGeom *geom;
  const b200_scan *scans;        // FINE / STRIP: per-frame vseg state
  const uint16_t *coarse_y;      // STRIP: the coarse y0 the FINE pass used
  uint8_t *cards;
  unsigned int *card_check;      // FULL only (may be null)
  int mode, rows_per_cta, nseg, nqseg, use_tma, frame0;
};

// ---- exact reference sequence for destination pixel (xb + x1, y), xb = origin of its 64-wide block
// (separate multiply and add: this file is compiled with -fmad=false)
__device__ __noinline__ int2 warp_coords_exact(const double *M, int xb, int y, int x1) {
  const double X0 = M[0] * xb + M[1] * y + M[2];
  const double Y0 = M[3] * xb + M[4] * y + M[5];
  const double W0 = M[6] * xb + M[7] * y + M[8];
  const double Wr = W0 + M[6] * x1;
  const double nx = X0 + M[0] * x1, ny = Y0 + M[3] * x1;
  double W = Wr != 0.0 ? 32. / Wr : 0.0;
  const double gx = nx * W, gy = ny * W;
  // cvt.rni.s32.f64: round-half-even, saturating == saturate_cast<int>(clamp(.))
  return make_int2(__double2int_rn(gx), __double2int_rn(gy));
}

__device__ __forceinline__ double rcp_newton(double w) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
  return __fma_rn(r, __fma_rn(-w, r, 1.0), r);
}

// The 15-bit weights are (32-fx)(32-fy)*32, fx(32-fy)*32, (32-fx)fy*32, fx*fy*32, so
//   (sum w_i v_i + 2^14) >> 15  ==  ((v0 (32-fx) + v1 fx)(32-fy) + (v2 (32-fx) + v3 fx) fy + 2^9) >> 10   exactly.
// OpenCV's table holds {32767, 0, 0, 1} at (0,0) (saturate_cast<short>(32768) + compensation); that entry also
// evaluates to v0 for every 8-bit v0, v3, as does this formula, so no special case is needed.
__device__ __forceinline__ int warp_blend(int fx, int fy, int v0, int v1, int v2, int v3) {
  const int ax = 32 - fx;
  const int top = v0 * ax + v1 * fx, bot = v2 * ax + v3 * fx;
  return (top * (32 - fy) + bot * fy + 512) >> 10;  // always in [0, 255]
}

// gather path: the four taps of the pixel at fixed-point (X, Y); taps outside [bx0, bx1) x [by0, by1) read as 0
__device__ __forceinline__ void gather_taps(const uint8_t *__restrict__ s, int row_stride, int bx0, int by0, int bx1, int by1, int X, int Y,
                                            int v[4]) {
  // saturate_cast<short>(X >> 5) only matters beyond +-32767 px, where every tap is outside the image anyway
  const int sx = X >> 5, sy = Y >> 5;
  if (sx >= bx0 && sx + 1 < bx1 && sy >= by0 && sy + 1 < by1) {
    const uint8_t *p = s + ((ptrdiff_t)sy * row_stride + sx);
    v[0] = __ldg(p), v[1] = __ldg(p + 1), v[2] = __ldg(p + row_stride), v[3] = __ldg(p + row_stride + 1);
  } else if (sx >= bx1 || sx + 1 < bx0 || sy >= by1 || sy + 1 < by0) {
    v[0] = v[1] = v[2] = v[3] = 0;
  } else {
    const bool x0ok = sx >= bx0 && sx < bx1, x1ok = sx + 1 >= bx0 && sx + 1 < bx1;
    const bool y0ok = sy >= by0 && sy < by1, y1ok = sy + 1 >= by0 && sy + 1 < by1;
    v[0] = (x0ok && y0ok) ? __ldg(s + ((ptrdiff_t)sy * row_stride + sx)) : 0;
    v[1] = (x1ok && y0ok) ? __ldg(s + ((ptrdiff_t)sy * row_stride + sx + 1)) : 0;
    v[2] = (x0ok && y1ok) ? __ldg(s + ((ptrdiff_t)(sy + 1) * row_stride + sx)) : 0;
    v[3] = (x1ok && y1ok) ? __ldg(s + ((ptrdiff_t)(sy + 1) * row_stride + sx + 1)) : 0;
  }
}

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

enum { kSkip = 0, kZero = 1, kTile = 2, kGather = 3 };

struct CtaPlan {  // written by thread 0, read by everyone after the barrier
  int what;       // kSkip .. kGather
  int first, nrows, rstride;  // destination rows first + i * rstride, i < nrows
  int tx0, ty0;   // frame coordinates of tile element (0, 0)
};

// TW x TH = shared-memory tile (256 x 64 or 64 x 256).
template <int TW, int TH>
__global__ void __launch_bounds__(kWarpThreads, 4)
warp_rows_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ WarpArgs A) {
  constexpr int kTileBytes = TW * TH;
  static_assert(kTileBytes <= 32768 && TW % 16 == 0 && TW <= 256 && TH <= 256, "tile shape");
  __shared__ __align__(128) uint8_t s_tile[kTileBytes];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ double sM[9];
  __shared__ double sF[3];  // M0, M3, M6 * 2^-19
  __shared__ CtaPlan s_plan;

  const int tid = threadIdx.x;
  const int seg = blockIdx.x % A.nseg, band = blockIdx.x / A.nseg;
  const int frame = blockIdx.y;  // within this launch; A.frame0 + frame within the batch
  const FrameGeom *g = A.geom + frame;
  const int q0 = seg * A.nqseg;
  const int nq = min(A.nqseg, kQuadsPerRow - q0);

  if (tid >= 32 && tid < 41) sM[tid - 32] = g->Minv[tid - 32];
  if (tid >= 64 && tid < 67) {
    const int i = tid - 64;
    sF[i] = g->Minv[3 * i] * (i == 2 ? 1.0 / 524288.0 : 1.0);
  }
  if (tid < 32) {
    // warp 0 plans the CTA: every lane derives the row set (uniform), lanes 0..3 each project one corner of the destination
    // rectangle, two shuffle rounds give the bounding box, lane 0 launches the copy
    CtaPlan P;
    P.what = kSkip, P.first = 0, P.nrows = 0, P.rstride = 1, P.tx0 = 0, P.ty0 = 0;
    const bool ok = g->all_found != 0;
    const int R = A.rows_per_cta;
    if (A.mode == WARP_FULL) {
      P.first = band * R, P.nrows = min(R, B200_CARD_H - P.first);
      P.what = ok ? kGather : kZero;  // frames without a card get a zero card (b200_transform_card_batch's `valid`)
    } else if (A.mode == WARP_COARSE) {
      P.first = 4 * band * R, P.nrows = min(R, 68 - band * R), P.rstride = 4;  // rows 0, 4, .., 268
      if (ok) P.what = kGather;
    } else if (ok) {
      const b200_scan *sc = A.scans + frame;
      if (A.mode == WARP_FINE) {
        const int y0 = sc->vseg.y_offset;  // coarse best, written by vseg_select pass 0 (0xFFFF: no fine rows)
        if (y0 != 0xFFFF) {
          const int lo = y0 < 8 ? 0 : y0 - 8, hi = min(B200_CARD_H, y0 + 27 + 8);
          P.first = lo + band * R, P.nrows = min(R, hi - P.first);
          P.what = kGather;
        }
      } else {  // WARP_STRIP: the final strip, only if it is not inside the rows FINE produced
        const int y0 = A.coarse_y[frame], yo = sc->vseg.y_offset;
        const int lo = y0 < 8 ? 0 : y0 - 8, hi = min(B200_CARD_H, y0 + 27 + 8);
        if (sc->usable && y0 != 0xFFFF && (yo < lo || yo + 27 > hi)) {
          P.first = yo + band * R, P.nrows = min(R, min(B200_CARD_H, yo + 27) - P.first);
          P.what = kGather;
        }
      }
    }
    if (P.nrows <= 0) P.what = kSkip;
    if (P.what == kGather && A.use_tma) {  // (warp-uniform)
      // bounding box of the source taps: images of the four corners of the destination rectangle
      const double *M = g->Minv;
      const double cx = (tid & 1) ? (double)(4 * (q0 + nq) - 1) : (double)(4 * q0);
      const double cy = (tid & 2) ? (double)(P.first + (P.nrows - 1) * P.rstride) : (double)P.first;
      const double w = M[6] * cx + M[7] * cy + M[8];
      const double fx = (M[0] * cx + M[1] * cy + M[2]) / w, fy = (M[3] * cx + M[4] * cy + M[5]) / w;
      // NaNs and far-away corners fail the range test; W must keep its sign over the rectangle
      bool fits = fx > -1500.0 && fx < 5500.0 && fy > -1500.0 && fy < 5500.0;
      int sgn = w > 1e-9 ? 1 : (w < -1e-9 ? 2 : 4);
      // taps of a pixel at real position f: columns floor(f - 1/64) .. floor(f + 1/64) + 1  (1/32-px rounding)
      int x_lo = fits ? (int)floor(fx) - 1 : 0, x_hi = fits ? (int)floor(fx) + 2 : 1 << 20;
      int y_lo = fits ? (int)floor(fy) - 1 : 0, y_hi = fits ? (int)floor(fy) + 2 : 1 << 20;
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        x_lo = min(x_lo, __shfl_xor_sync(0xffffffffu, x_lo, o)), x_hi = max(x_hi, __shfl_xor_sync(0xffffffffu, x_hi, o));
        y_lo = min(y_lo, __shfl_xor_sync(0xffffffffu, y_lo, o)), y_hi = max(y_hi, __shfl_xor_sync(0xffffffffu, y_hi, o));
        sgn |= __shfl_xor_sync(0xffffffffu, sgn, o);
      }
      const int tx0 = x_lo & ~15;  // 16-byte aligned box start (in frame coordinates; ox is a multiple of 16 as well)
      if ((sgn == 1 || sgn == 2) && x_hi - tx0 < TW && y_hi - y_lo < TH) {
        P.what = kTile, P.tx0 = tx0, P.ty0 = y_lo;
        if (tid == 0) {
          const unsigned int bar = smem_u32(&s_bar), dst = smem_u32(s_tile);
          asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTileBytes) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
              "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(tx0 - A.ox), "r"(y_lo - A.oy), "r"(A.frame0 + frame), "r"(bar)
              : "memory");
        }
      }
    }
    if (tid == 0) s_plan = P;
  }
  __syncthreads();
  const int what = s_plan.what;
  if (what == kSkip) return;
  const int first = s_plan.first, nrows = s_plan.nrows, rstride = s_plan.rstride;

  // thread -> (row slot rr, quad q) of the segment; rp row slots per pass
  const int rp = kWarpThreads / nq;
  const int rr = tid / nq, q = tid - rr * nq;
  const bool active = rr < rp;
  const int x = 4 * (q0 + q);
  uint8_t *dst = A.cards + (size_t)frame * (B200_CARD_W * B200_CARD_H);
  unsigned int sum = 0;

  if (what == kZero) {
    if (active)
      for (int i = rr; i < nrows; i += rp) *reinterpret_cast<unsigned int *>(dst + (first + i * rstride) * B200_CARD_W + x) = 0u;
    return;
  }

  // fast-path linear forms of this thread's first pixel; stepped rp * rstride rows per pass
  const double kWScale = 1.0 / 524288.0;  // 2^-19: 1 / (W * 2^-19) = 32 * 2^14 / W
  const int yfirst = first + rr * rstride;
  double tX = __fma_rn(sM[0], (double)x, __fma_rn(sM[1], (double)yfirst, sM[2]));
  double tY = __fma_rn(sM[3], (double)x, __fma_rn(sM[4], (double)yfirst, sM[5]));
  double tW = __fma_rn(sM[6], (double)x, __fma_rn(sM[7], (double)yfirst, sM[8])) * kWScale;
  const double step = (double)(rp * rstride);
  const double dX = sM[1] * step, dY = sM[4] * step, dW = sM[7] * step * kWScale;
  const int xblk = x & ~63;  // the reference's 64-wide block origin (exact path)
  unsigned int doff = (unsigned)(yfirst * B200_CARD_W + x);
  const unsigned int dstep = (unsigned)(rp * rstride * B200_CARD_W);

  const int nr = active ? nrows : 0;  // threads beyond the last whole row slot idle (and stay for the reduction below)
  if (what == kTile) {
    // wait for the tile (phase 0 of the barrier)
    {
      const unsigned int bar = smem_u32(&s_bar);
      unsigned int done;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
      } while (!done);
    }
    // u >> 19 = source pixel + 2048: fold the tile origin and that bias into the base pointer
    const uint8_t *tb = s_tile - ((s_plan.ty0 + 2048) * TW + (s_plan.tx0 + 2048));
    const double kMagic = 6755399441055744.0 + 1073741824.0 + 8192.0;  // 1.5 * 2^52 + 2^30 + 2^13
#pragma unroll 1
    for (int i = rr; i < nr; i += rp) {
      const volatile double *F = sF;  // re-read from shared memory every pass: cheaper than holding six registers
      const double m0 = F[0], m3 = F[1], m6s = F[2];
      unsigned int ux[4], uy[4];
      bool tie = false;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double kd = (double)k;  // immediate operand
        const double Wf = k ? __fma_rn(m6s, kd, tW) : tW, nxf = k ? __fma_rn(m0, kd, tX) : tX, nyf = k ? __fma_rn(m3, kd, tY) : tY;
        const double r = rcp_newton(Wf);
        ux[k] = (unsigned)__double2loint(__fma_rn(nxf, r, kMagic));
        uy[k] = (unsigned)__double2loint(__fma_rn(nyf, r, kMagic));
        tie = tie || (ux[k] & 0x3FFEu) == 0u || (uy[k] & 0x3FFEu) == 0u;
      }
      tX += dX, tY += dY, tW += dW;
      if (tie) {
        const int y = first + i * rstride;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if ((ux[k] & 0x3FFEu) == 0u || (uy[k] & 0x3FFEu) == 0u) {
            const int2 e = warp_coords_exact(sM, xblk, y, x + k - xblk);
            ux[k] = ((unsigned)(e.x + 65536) << 14) | 2u, uy[k] = ((unsigned)(e.y + 65536) << 14) | 2u;
          }
      }
      unsigned int packed = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint8_t *p = tb + ((uy[k] >> 19) * TW + (ux[k] >> 19));
        const unsigned int px = (unsigned)warp_blend((int)((ux[k] >> 14) & 31u), (int)((uy[k] >> 14) & 31u), p[0], p[1], p[TW], p[TW + 1]);
        packed |= px << (8 * k);
        sum += (doff + 1u + k) * px;
      }
      *reinterpret_cast<unsigned int *>(dst + doff) = packed;
      doff += dstep;
    }
  } else {
    // gather path: byte taps through L1 with explicit bounds (the crop, or the whole plane)
    const int bx0 = A.ox, by0 = A.oy, bx1 = A.ox + A.bw, by1 = A.oy + A.bh;
    const uint8_t *s = A.src + (size_t)(A.frame0 + frame) * A.frame_stride - ((ptrdiff_t)A.oy * A.row_stride + A.ox);  // virtual pixel (0, 0)
    const double kMagic = 6755399441055744.0 + 1073741824.0;  // 1.5 * 2^52 + 2^30
#pragma unroll 1
    for (int i = rr; i < nr; i += rp) {
      const volatile double *F = sF;
      const double m0 = F[0], m3 = F[1], m6s = F[2];
      int X[4], Y[4];
      unsigned bad = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double kd = (double)k;
        const double Wf = k ? __fma_rn(m6s, kd, tW) : tW, nxf = k ? __fma_rn(m0, kd, tX) : tX, nyf = k ? __fma_rn(m3, kd, tY) : tY;
        const double r = rcp_newton(Wf);
        const double tx = __fma_rn(nxf, r, kMagic), ty = __fma_rn(nyf, r, kMagic);
        const unsigned ux = (unsigned)__double2loint(tx), uy = (unsigned)__double2loint(ty);
        // valid while the high word is still that of the constant: -65536 <= fX < 196608 (W ~ 0 and far-away coordinates fail)
        bad |= ((unsigned)__double2hiint(tx) ^ 0x43380000u) | ((unsigned)__double2hiint(ty) ^ 0x43380000u);
        const unsigned nearx = (ux & 0x3FFFu) - (0x2000u - 1u), neary = (uy & 0x3FFFu) - (0x2000u - 1u);  // <= 2: within 1/16384 of .5
        bad |= (unsigned)(min(nearx, neary) <= 2u);
        X[k] = (int)((ux + 0x2000u) >> 14) - 65536, Y[k] = (int)((uy + 0x2000u) >> 14) - 65536;
      }
      tX += dX, tY += dY, tW += dW;
      if (bad) {
        const int y = first + i * rstride;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int2 e = warp_coords_exact(sM, xblk, y, x + k - xblk);
          X[k] = e.x, Y[k] = e.y;
        }
      }
      unsigned int packed = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int v[4];
        gather_taps(s, A.row_stride, bx0, by0, bx1, by1, X[k], Y[k], v);
        const unsigned int px = (unsigned)warp_blend(X[k] & 31, Y[k] & 31, v[0], v[1], v[2], v[3]);
        packed |= px << (8 * k);
        sum += (doff + 1u + k) * px;
      }
      *reinterpret_cast<unsigned int *>(dst + doff) = packed;
      doff += dstep;
    }
  }
  if (A.card_check != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0 && sum) atomicAdd(&A.card_check[frame], sum);
  }
}

// ---- tensor map over the source planes: u8 [n][bh][bw], box TW x TH x 1, zero fill outside
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool make_tensor_map(CUtensorMap *map, const WarpSource &S, int tw, int th) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const size_t fs = S.n > 1 ? S.frame_stride : (size_t)S.row_stride * S.bh;
  if ((reinterpret_cast<uintptr_t>(S.base) & 15u) || (S.row_stride & 15) || (fs & 15u) || (S.ox & 15)) return false;
  if (S.row_stride < S.bw || fs < (size_t)S.row_stride * S.bh) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)S.bw, (cuuint64_t)S.bh, (cuuint64_t)S.n};
  const cuuint64_t strides[2] = {(cuuint64_t)S.row_stride, (cuuint64_t)fs};
  const cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 1u}, estr[3] = {1u, 1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(S.base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

}  // namespace

// Segments per row / rows per CTA: chosen so that the source box of a CTA fits the tile for a card that fills the guide
// rectangle of a frame of this size (a CTA whose box does not fit falls back to the gather path, so this is tuning only).
int launch_warp(const WarpSource &S, const FrameGeom *geom, uint8_t *cards, unsigned int *card_check, int mode, const b200_scan *scans,
                const uint16_t *coarse_y, int portrait, cudaStream_t s) {
  if (card_check && cudaMemsetAsync(card_check, 0, sizeof(unsigned int) * (size_t)S.n, s) != cudaSuccess) return -1;
  static const int force_gather = env_int("B200_DMZ_WARP_GATHER", 0);
  static const int force_rows = env_int("B200_DMZ_WARP_ROWS", 0);
  static const int force_segs = env_int("B200_DMZ_WARP_SEGS", 0);
  WarpArgs A;
  A.src = S.base, A.row_stride = S.row_stride, A.frame_stride = S.frame_stride, A.bw = S.bw, A.bh = S.bh, A.ox = S.ox, A.oy = S.oy;
  A.geom = geom, A.scans = scans, A.coarse_y = coarse_y, A.cards = cards, A.card_check = card_check, A.mode = mode;
  // scale of the guide rectangle relative to the 428 x 270 card: 640 x 480 frames show the card at ~1:1
  const double scale = (double)(S.frame_h > 0 ? S.frame_h : S.bh) / 480.0;
  // tile: the long side (256) runs along the destination rows (frame columns in landscape, frame rows in portrait); the
  // short side is 128 (measured on B200, ms per 100 k frames, 128 against 64: materialised card 18.9 / 21.3, lazy rows
  // 9.7 / 11.4 -- fewer, larger CTAs amortise the per-CTA set-up).  B200_DMZ_WARP_TILE=64|128 overrides.
  static const int force_tile = env_int("B200_DMZ_WARP_TILE", 0);
  // The 43-row fine window and the 27-row strip need ~60 source lines: a 256 x 64 box fetches half the bytes of the 256 x 128
  // one the full card / the coarse rows use (B200_DMZ_WARP_FINE_TILE overrides).
  static const int fine_tile = env_int("B200_DMZ_WARP_FINE_TILE", 64);
  const int tshort = force_tile == 64 || force_tile == 128 ? force_tile : ((mode == WARP_FINE || mode == WARP_STRIP) && fine_tile == 64 ? 64 : 128);
  const int tw = portrait ? tshort : 256, th = portrait ? 256 : tshort;
  // along a destination row the source advances `scale` pixels per pixel; keep ~8 % slack for corner jitter and 24 for
  // the aligned start and the tap margins
  int nseg = (int)(428.0 * scale * 1.08 / (256 - 24)) + 1;
  if (nseg < 2) nseg = 2;  // two segments of 54 / 53 quads: 4 x 54 = 216 of the 224 threads busy
  if (force_segs > 0) nseg = force_segs;
  if (nseg > kQuadsPerRow) nseg = kQuadsPerRow;
  int nqseg = (kQuadsPerRow + nseg - 1) / nseg;
  nseg = (kQuadsPerRow + nqseg - 1) / nqseg;
  const int rp = kWarpThreads / nqseg;
  // across rows: rows * rstride * scale source lines + the tilt of the segment (16 px of corner jitter over the card
  // width, shared out over the segments) + 6 must fit the short tile side
  const int rstride = mode == WARP_COARSE ? 4 : 1;
  int rows = (int)(((double)tshort - 6.0 - 18.0 * scale / nseg) / (scale * 1.06 * rstride));
  if (rstride > 1) rows += 1;  // (rows - 1) * rstride lines between the first and the last row
  rows -= rows % rp;           // whole passes
  if (rows < rp) rows = rp;
  if (force_rows > 0) rows = force_rows;
  const int total_rows = mode == WARP_FULL ? B200_CARD_H : (mode == WARP_COARSE ? 68 : (mode == WARP_FINE ? 43 : 27));
  if (rows > total_rows) rows = total_rows;
  const int bands = (total_rows + rows - 1) / rows;
  A.rows_per_cta = rows, A.nseg = nseg, A.nqseg = nqseg;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  A.use_tma = !force_gather && make_tensor_map(&map, S, tw, th) ? 1 : 0;
  int launches = 0;
  for (int f0 = 0; f0 < S.n; f0 += 65535) {
    const int cnt = S.n - f0 < 65535 ? S.n - f0 : 65535;
    WarpArgs B = A;
    B.frame0 = f0;
    B.geom = geom + f0, B.scans = scans ? scans + f0 : nullptr, B.coarse_y = coarse_y ? coarse_y + f0 : nullptr;
    B.cards = cards + (size_t)f0 * (B200_CARD_W * B200_CARD_H), B.card_check = card_check ? card_check + f0 : nullptr;
    const dim3 grid(bands * nseg, cnt);
    if (tw == 256 && th == 64) warp_rows_kernel<256, 64><<<grid, kWarpThreads, 0, s>>>(map, B);
    else if (tw == 256) warp_rows_kernel<256, 128><<<grid, kWarpThreads, 0, s>>>(map, B);
    else if (tw == 64) warp_rows_kernel<64, 256><<<grid, kWarpThreads, 0, s>>>(map, B);
    else warp_rows_kernel<128, 256><<<grid, kWarpThreads, 0, s>>>(map, B);
    launches++;
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}
