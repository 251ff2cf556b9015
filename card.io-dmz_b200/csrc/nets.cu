// card.io-dmz_b200/csrc/nets.cu -- the generated networks of the number path as FP32 CUDA-core kernels (FMA; contract:
// <= 1e-4 on the probabilities, scan/../models KATs at 1e-5), the digit patch preparation, the expiry digit CNN, and the
// launch sequence of scan_card_image.
//
//   vseg_rows_kernel    vseg_probabilities_for_hstrip = llcv_morph_grad3_1d_u8 -> llcv_lineardown2_1d_u8 ->
//                       llcv_norm_convert_1d_u8_to_f32 -> applym_befe75da      scan/n_vseg.cpp:39-47,
//                                                                               models/generated/modelm_befe75da.cpp:1770-1786
//   digit_prep_kernel   per digit ROI -> llcv_morph_grad3_2d_cross_u8 -> llcv_equalize_hist     scan/n_categorize.cpp:75-108
//   categorize_kernel   applyc_{5c241121,01266c1b,b00bf70c} -> (r0+r1+r2-max)/2                 scan/n_categorize.cpp:45-73,
//                                                                               models/generated/modelc_*.cpp:1844-1937
//
// Since round 2 the card rows and the prepared byte patches of the whole path go through the tensor-core kernels
// (vseg_mma.cu, categorize_mma.cu: batched over frames the contractions do fill MMA tiles, and with byte activations they
// are exact integer products).  The FP32 kernels of this file remain the route of FLOAT inputs -- the reference's embedded
// model known-answer vectors are float rows / float patches -- and the A / B reference (B200_DMZ_VSEG_FP32 /
// B200_DMZ_CNN_FP32): convolution weights are warp-uniform kernel-parameter operands, FC weights sit in shared memory for
// the lifetime of a persistent CTA.
#include <float.h>

#include <stdlib.h>

#include "b200_internal.h"

// exact.cu
int launch_scan_gate(const FrameGeom *geom, const uint8_t *valid, int n, uint8_t *gate, cudaStream_t s);
int launch_vseg_select(const float *vprob, const uint8_t *gate, int n, int pass, b200_scan *scans, cudaStream_t s, uint16_t *coarse_y = nullptr);
int launch_hseg(const uint8_t *cards, int n, b200_scan *scans, cudaStream_t s);
int launch_scan_finish(int n, b200_scan *scans, cudaStream_t s);

namespace {

constexpr int kCardBytes = B200_CARD_W * B200_CARD_H;

// tanh x = 1 - 2 / (exp(2x) + 1) on the special-function unit: ex2.approx and rcp.approx are good to ~2^-22 relative,
// the result to ~2e-7 ABSOLUTE (the same order as tanhf's own 2 ulp near +-1); saturates to +-1 for large |x|.
// tanhf costs ~22 instructions, this 6; used by the digit CNNs (1e-4 contract on probabilities), not by vseg (index).
__device__ __forceinline__ float tanh_sfu(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// (the conv kernels and post-pool biases of the three digit CNNs arrive as kernel parameters: ConvConsts, b200_internal.h)

// ------------------------------------------------------------------------------------------------
// V1 + V2.  Warp-autonomous tiles: every warp of a persistent CTA owns tiles of kVTile = 16 (frame,row) work items
// and takes each through preparation, hidden layer and output layer by itself -- no block barrier inside the tile
// loop, so the preparation of one warp overlaps the FMAs of another.  W1 (50 x 204, rows padded to 212 floats, units
// padded to 56 with zeros) stays in shared memory for the CTA's lifetime.
//
// Hidden layer: lane = (row lane lr = lane / 8, unit lane lu = lane % 8); a thread accumulates rows lr + 4i (i < 4)
// x units lu + 8j (j < 7) = 28 sums in registers.  Per 4-float K step it issues 4 + 7 LDS.128 (each row address is
// shared by 8 lanes, each unit address by 4: 64 B resp. 128 B unique per request) for 112 FMAs, so the shared-memory
// crossbar stays at ~0.4 of the FMA issue rate (the earlier 1 unit x 7 rows tiling was crossbar-bound at ~54 % issue).
// ------------------------------------------------------------------------------------------------
constexpr int kVWarps = 12;
constexpr int kVThreads = kVWarps * 32;
constexpr int kVTile = 16;     // work items per warp tile
constexpr int kVStride = 212;  // 53 16-byte units per row (odd): the 4 row / 8 unit addresses of a request hit distinct bank groups
constexpr int kVUnits = 56;    // 50 hidden units padded to 7 x 8
constexpr int kFineSlots = 43; // rows [y0 - 8, y0 + 35)

struct VsegWarp {
  alignas(16) float x[kVTile][kVStride];  // prepared rows of the current tile
  alignas(4) uint8_t raw[412];            // one staged card row (408 px)
  int item_frame[kVTile];
  int item_row[kVTile];
  int pad[3];
};
struct VsegSmem {
  alignas(16) float w1[kVUnits * kVStride];
  float b1[kVUnits];
  float w2[3][kVUnits];
  float b2[4];
  VsegWarp warp[kVWarps];
};

// mode 0: coarse rows 0,4,..,268 of every gated frame.  mode 1: fine rows around the coarse best.
// mode 2: raw prepared rows (stage tap): in = n x 204 floats, out = n x 3.
__global__ void __launch_bounds__(kVThreads, 1)
vseg_rows_kernel(const float *__restrict__ wts, const float *__restrict__ norm_tab, const uint8_t *__restrict__ cards,
                 const uint8_t *__restrict__ gate, const b200_scan *__restrict__ scans, int n, int mode, float *__restrict__ vprob,
                 const float *__restrict__ raw_rows, float *__restrict__ raw_out) {
  extern __shared__ __align__(16) uint8_t vs_raw[];
  VsegSmem &S = *reinterpret_cast<VsegSmem *>(vs_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kVUnits * kVStride; i += kVThreads) {
    const int u = i / kVStride, k = i - u * kVStride;
    S.w1[i] = (u < 50 && k < 204) ? __ldg(wts + u * 204 + k) : 0.0f;
  }
  for (int i = tid; i < kVUnits; i += kVThreads) S.b1[i] = i < 50 ? __ldg(wts + 10200 + i) : 0.0f;
  for (int i = tid; i < 3 * kVUnits; i += kVThreads) {
    const int c = i / kVUnits, u = i - c * kVUnits;
    S.w2[c][u] = u < 50 ? __ldg(wts + 10250 + c * 50 + u) : 0.0f;
  }
  if (tid < 3) S.b2[tid] = __ldg(wts + 10400 + tid);
  __syncthreads();  // the only block barrier: weights visible

  VsegWarp &Wp = S.warp[warp];
  const int per_frame = mode == 0 ? 68 : (mode == 1 ? kFineSlots : 1);
  const long long total = (long long)n * per_frame;
  const long long ntiles = (total + kVTile - 1) / kVTile;
  const int lr = lane >> 3, lu = lane & 7;

  for (long long tile = (long long)blockIdx.x * kVWarps + warp; tile < ntiles; tile += (long long)gridDim.x * kVWarps) {
    // ---- resolve the work items of this tile (lanes 0..15)
    int fr = -1, row = -1;
    if (lane < kVTile) {
      const long long item = tile * kVTile + lane;
      if (item < total) {
        const int f = (int)(item / per_frame), j = (int)(item % per_frame);
        if (mode == 2) {
          fr = f, row = 0;
        } else if (!gate || gate[f]) {
          if (mode == 0) {
            fr = f, row = 4 * j;
          } else {
            const int y0 = scans[f].vseg.y_offset;  // coarse best (vseg_select pass 0)
            const int lo = y0 < 8 ? 0 : y0 - 8;     // n_vseg.cpp:140-142
            const int hi = min(270, y0 + 27 + 8);
            const int r = lo + j;
            if (y0 != 0xFFFF && r < hi && (r & 3) != 0) fr = f, row = r;  // rows r % 4 == 0 were scored by the coarse pass
          }
        }
      }
    }
    const unsigned active = __ballot_sync(0xffffffffu, fr >= 0);
    if (active == 0u) continue;  // warp-uniform: nothing to score in this tile
    __syncwarp();                // the previous tile's readers of x / item_* are done
    if (lane < kVTile) Wp.item_frame[lane] = fr, Wp.item_row[lane] = row;
    __syncwarp();

    // ---- V1: row preparation, one row at a time; the 16-bit loads of the next active row are issued before the
    // current row is processed (software prefetch: one global-latency exposure per tile, not per row)
    if (mode == 2) {
      for (int r = 0; r < kVTile; r++) {
        const int f = Wp.item_frame[r];
        if (f < 0) continue;
        for (int k = lane; k < 204; k += 32) Wp.x[r][k] = __ldg(raw_rows + (size_t)f * 204 + k);
      }
    } else {
      unsigned short pre[7];
      auto fetch = [&](int r) {
        const uint8_t *src = cards + (size_t)Wp.item_frame[r] * kCardBytes + (size_t)Wp.item_row[r] * B200_CARD_W + 10;
        // the row starts at an even, not 4-aligned, offset: coalesced 16-bit loads
#pragma unroll
        for (int q = 0; q < 7; q++) {
          const int k = lane + 32 * q;
          pre[q] = k < 204 ? __ldg(reinterpret_cast<const unsigned short *>(src) + k) : (unsigned short)0;
        }
      };
      unsigned todo = active;
      int r = __ffs(todo) - 1;
      fetch(r);
      while (todo) {
        todo &= todo - 1;
        uint8_t *raw = Wp.raw;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 7; q++) {
          const int k = lane + 32 * q;
          if (k < 204) reinterpret_cast<unsigned short *>(raw)[k] = pre[q];
        }
        __syncwarp();
        const int rnext = todo ? __ffs(todo) - 1 : -1;
        if (rnext >= 0) fetch(rnext);
        // ROI (10, row, 408, 1): 3-tap max - min with replicate at the ROI edge, then (a + b + 1) >> 1
        int vals[7];
        int mn = 255, mx = 0;
#pragma unroll
        for (int q = 0; q < 7; q++) {
          const int k = lane + 32 * q;  // output index 0..203
          int v = 0;
          if (k < 204) {
            const int j0 = 2 * k, j1 = 2 * k + 1;
            const int a = raw[j0 > 0 ? j0 - 1 : 0], b = raw[j0], c = raw[j1], d = raw[j1 < 407 ? j1 + 1 : 407];
            const int g0 = max(a, max(b, c)) - min(a, min(b, c));
            const int g1 = max(b, max(c, d)) - min(b, min(c, d));
            v = (g0 + g1 + 1) >> 1;
            mn = min(mn, v);
            mx = max(mx, v);
          }
          vals[q] = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        // cvConvertScale(1/255) then cvNormalize(MINMAX 0..1): float multiply, then float multiply + float add
        // (scale and shift as cv::normalize derives them in double from min and max: host-built table, b200_tables.cpp)
        const float k255 = 1.0f / 255.0f;
        const float2 nrm = __ldg(reinterpret_cast<const float2 *>(norm_tab) + (mn * 256 + mx));
        const float fs = nrm.x, fb = nrm.y;
#pragma unroll
        for (int q = 0; q < 7; q++) {
          const int k = lane + 32 * q;
          if (k < 204) Wp.x[r][k] = __fadd_rn(__fmul_rn(__fmul_rn((float)vals[q], k255), fs), fb);
        }
        r = rnext;
      }
    }
    __syncwarp();

    // ---- V2 hidden layer: 4 rows x 7 units per thread (rows of inactive items hold stale data; their sums are dropped)
    float acc[4][7];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 7; j++) acc[i][j] = 0.0f;
    {
      const float4 *x4 = reinterpret_cast<const float4 *>(&Wp.x[lr][0]);       // row lr + 4i at x4 + i * 4 * 53
      const float4 *w4 = reinterpret_cast<const float4 *>(S.w1 + lu * kVStride);  // unit lu + 8j at w4 + j * 8 * 53
#pragma unroll 3
      for (int k4 = 0; k4 < 51; k4++) {
        float4 xv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) xv[i] = x4[i * 4 * (kVStride / 4) + k4];
#pragma unroll
        for (int j = 0; j < 7; j++) {
          const float4 wv = w4[j * 8 * (kVStride / 4) + k4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            acc[i][j] = fmaf(wv.x, xv[i].x, acc[i][j]);
            acc[i][j] = fmaf(wv.y, xv[i].y, acc[i][j]);
            acc[i][j] = fmaf(wv.z, xv[i].z, acc[i][j]);
            acc[i][j] = fmaf(wv.w, xv[i].w, acc[i][j]);
          }
        }
      }
    }
    // ---- logistic layer: per-thread partial sums over its 7 units, butterfly over the 8 unit lanes, then softmax
    // (expf / sum, no max shift -- as the generated model does)
    float o[4][3];
#pragma unroll
    for (int i = 0; i < 4; i++) o[i][0] = o[i][1] = o[i][2] = 0.0f;
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const int u = lu + 8 * j;
      const float b = S.b1[u], w20 = S.w2[0][u], w21 = S.w2[1][u], w22 = S.w2[2][u];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float hv = tanhf(acc[i][j] + b);  // feeds an arg-max index: keep the accurate version
        o[i][0] = fmaf(w20, hv, o[i][0]);
        o[i][1] = fmaf(w21, hv, o[i][1]);
        o[i][2] = fmaf(w22, hv, o[i][2]);
      }
    }
#pragma unroll
    for (int sft = 1; sft < 8; sft <<= 1)
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) o[i][c] += __shfl_xor_sync(0xffffffffu, o[i][c], sft);
    // unit lane i (< 4) finishes row lr + 4i
    if (lu < 4) {
      float z0 = o[0][0], z1 = o[0][1], z2 = o[0][2];
#pragma unroll
      for (int i = 1; i < 4; i++)
        if (lu == i) z0 = o[i][0], z1 = o[i][1], z2 = o[i][2];
      const int r = lr + 4 * lu;
      const int f = Wp.item_frame[r];
      if (f >= 0) {
        const float e0 = expf(z0 + S.b2[0]), e1 = expf(z1 + S.b2[1]), e2 = expf(z2 + S.b2[2]);
        const float sum = (e0 + e1) + e2;
        if (mode == 2) {
          raw_out[(size_t)f * 3 + 0] = e0 / sum;
          raw_out[(size_t)f * 3 + 1] = e1 / sum;
          raw_out[(size_t)f * 3 + 2] = e2 / sum;
        } else {
          float *dst = vprob + ((size_t)f * 270 + Wp.item_row[r]) * 2;
          dst[0] = e1 / sum;  // visa-like
          dst[1] = e2 / sum;  // amex-like
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C0..C2.  One CTA (16 warps) processes the 16 digit slots of a "group" (= one frame, or 16 raw patches);
// persistent over groups so the 123 KB of transposed hidden weights are staged once per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kCThreads = 512;

constexpr int kQStride = 528;  // bytes per prepared digit patch in global memory (27 x 19 = 513, padded to 33 x 16)

struct CatPrepWarp {  // patch-preparation scratch of one warp (digit_prep_kernel)
  unsigned int hist[256];
  uint8_t raw[27 * 24];
  uint8_t g8[27 * 20];
  uint8_t lut[256];
};
constexpr int kFeatStride = 324;  // 81 16-byte units (odd): the four digit rows of a hidden-layer request hit distinct bank groups
struct CatWork {  // network activations of the current group
  alignas(16) float feat[16][kFeatStride];  // tanh(pool + bias) of the current model
  alignas(16) float part[16][16][32];       // hidden-layer partial sums of the sixteen K slices: [slice][digit][unit]
  float hid[16][3][32];
  float prob[16][3][10];
};
struct CatSmem {
  float hwT[3][320][32];  // hidden W transposed: [model][feature][unit]
  float hb[3][32];
  float lw[3][10][32];
  float lb[3][10];
  alignas(16) float patch[16][kQStride];  // normalised digit images (513 floats used per row)
  CatWork work;
};

// four of the eight kernels of model m on one pooled cell: conv 3x3 over the 5x5 window, 3x3 max, + bias, tanh
template <int K0>
__device__ __forceinline__ void conv_pool_four(const ConvConsts &C, int m, const float (&win)[5][5], float *feat_cell /* stride 40 per kernel */) {
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    const int k = K0 + kk;
    float best = -FLT_MAX;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) acc = fmaf(C.w[m][k][i * 3 + j], win[r + i][c + j], acc);
        best = fmaxf(best, acc);
      }
    feat_cell[k * 40] = tanh_sfu(best + C.b[m][k]);
  }
}

// C0 + C1: digit patch preparation, one warp per (frame, digit slot): ROI (offsets[d], y_off, 19, 27) ->
// llcv_morph_grad3_2d_cross_u8 on the isolated ROI -> llcv_equalize_hist.  Writes the equalised bytes (the CNN kernel
// applies cvConvertScale's * 1/255 when it expands them), kQStride bytes per digit.  A kernel of its own so that it runs
// at full occupancy (2.5 KB of scratch per warp) instead of inside the one-CTA-per-SM CNN kernel between barriers.
constexpr int kPrepWarps = 8;

template <bool kRaw>
__global__ void __launch_bounds__(kPrepWarps * 32)
digit_prep_kernel(const uint8_t *__restrict__ cards, const b200_scan *__restrict__ scans, const uint8_t *__restrict__ raw_patches,
                  int n_digits /* frames * 16, or raw patches */, uint8_t *__restrict__ q8) {
  __shared__ CatPrepWarp s_prep[kPrepWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * kPrepWarps + warp;
  if (item >= n_digits) return;
  CatPrepWarp &P = s_prep[warp];
  uint8_t *raw = P.raw;  // [27][24], patch column c at byte c + a of each row
  int a = 0;
  if (kRaw) {
    const uint8_t *src = raw_patches + (size_t)item * (27 * 19);
    for (int i = lane; i < 27 * 19; i += 32) raw[(i / 19) * 24 + (i % 19)] = __ldg(src + i);
  } else {
    const int f = item >> 4, d = item & 15;
    const b200_scan *sc = scans + f;
    if (!sc->usable || d >= (int)sc->hseg.n_offsets) return;  // warp-uniform: upside-down / vseg gate (frame.cpp:38-47)
    // rows of a card are 428 = 4 * 107 bytes apart, so every row of the patch has the same word alignment:
    // fetch the <= 6 aligned words covering the 19 bytes of each row
    const uint8_t *src = cards + (size_t)f * kCardBytes + (size_t)sc->vseg.y_offset * B200_CARD_W + sc->hseg.offsets[d];
    a = (int)(reinterpret_cast<uintptr_t>(src) & 3u);
    const unsigned int *wsrc = reinterpret_cast<const unsigned int *>(src - a);
    const int words = (a + 19 + 3) >> 2;
    // all six loads of a lane are issued before the first store: one global-latency exposure per digit, not six
    unsigned int wv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int i = lane + 32 * k, row = i / 6, q = i - row * 6;
      wv[k] = (i < 27 * 6 && q < words) ? __ldg(wsrc + row * (B200_CARD_W / 4) + q) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int i = lane + 32 * k;
      if (i < 27 * 6) reinterpret_cast<unsigned int *>(raw)[i] = wv[k];
    }
  }
  for (int i = lane; i < 256; i += 32) P.hist[i] = 0;
  __syncwarp();
  // 5-point cross max - min with replicate at the PATCH edge (cv/morph.cpp:177-255 on the 19x27 ROI)
  for (int i = lane; i < 27 * 19; i += 32) {
    const int y = i / 19, x = i - y * 19;
    const int yu = y > 0 ? y - 1 : y, yd = y < 26 ? y + 1 : y, xl = x > 0 ? x - 1 : x, xr = x < 18 ? x + 1 : x;
    const int p = raw[yu * 24 + x + a], q = raw[y * 24 + xl + a], c = raw[y * 24 + x + a];
    const int e = raw[y * 24 + xr + a], f = raw[yd * 24 + x + a];
    const int v = max(p, max(q, max(c, max(e, f)))) - min(p, min(q, min(c, min(e, f))));
    P.g8[y * 20 + x] = (uint8_t)v;
    atomicAdd(&P.hist[v], 1u);
  }
  __syncwarp();
  // llcv_equalize_hist (cv/stats.cpp:116-159): lut[i] = sat8(cvRound(cum(i) * (255.f / 513))), lut[0] = 0
  {
    unsigned int local[8], run = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      run += P.hist[lane * 8 + q];
      local[q] = run;
    }
    unsigned int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned int excl = incl - run;
    const float scale = 255.f / (19 * 27);
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int val = __float2int_rn(__fmul_rn((float)(int)(excl + local[q]), scale));
      P.lut[lane * 8 + q] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
    }
    __syncwarp();
    if (lane == 0) P.lut[0] = 0;
    __syncwarp();
  }
  uint8_t *dst = q8 + (size_t)item * kQStride;
  for (int i = lane; i < 27 * 19; i += 32) {
    const int y = i / 19, x = i - y * 19;
    dst[i] = P.lut[P.g8[y * 20 + x]];
  }
  if (lane < kQStride - 27 * 19) dst[27 * 19 + lane] = 0;  // the padding is fetched (and ignored) by the CNN kernel
}

// C2 on prepared patches.  One CTA (16 warps) per SM, persistent over groups (= one frame's 16 digit slots, or 16 raw
// patches) so the 123 KB of transposed hidden weights are staged once per CTA.  The 8.4 KB of prepared bytes of the NEXT
// group are requested (one 128-bit load per thread) before the current group is computed and expanded to floats after it.
template <bool kRaw>
__global__ void __launch_bounds__(kCThreads, 1)
categorize_kernel(const __grid_constant__ NetWeights W, const uint8_t *__restrict__ q8, b200_scan *__restrict__ scans,
                  const float *__restrict__ raw_float, int n_items /* frames, or raw patches */, float *__restrict__ raw_out) {
  extern __shared__ __align__(16) uint8_t cs_raw[];
  CatSmem &S = *reinterpret_cast<CatSmem *>(cs_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  (void)lane;

  for (int i = tid; i < 3 * 320 * 32; i += kCThreads) (&S.hwT[0][0][0])[i] = __ldg(W.cnn_hwT + i);
  for (int m = 0; m < 3; m++) {
    const float *b = W.cnn[m];
    for (int i = tid; i < 32; i += kCThreads) S.hb[m][i] = __ldg(b + 80 + 10240 + i);
    for (int i = tid; i < 320; i += kCThreads) S.lw[m][i / 32][i % 32] = __ldg(b + 80 + 10240 + 32 + i);
    for (int i = tid; i < 10; i += kCThreads) S.lb[m][i] = __ldg(b + 80 + 10240 + 32 + 320 + i);
  }

  const int n_groups = kRaw ? (n_items + 15) / 16 : n_items;
  const size_t q8_limit = (size_t)n_items * (kRaw ? kQStride : 16 * kQStride);  // bytes of q8 that exist
  // 16 digits x kQStride bytes = 528 uint4 per group: thread t holds uint4 t, threads 0..15 also uint4 512 + t
  uint4 pre0 = make_uint4(0, 0, 0, 0), pre1 = make_uint4(0, 0, 0, 0);
  auto prefetch = [&](int grp) {
    if (q8 == nullptr || grp >= n_groups) return;
    const size_t base = (size_t)grp * 16 * kQStride;
    const size_t o0 = base + (size_t)tid * 16, o1 = base + (size_t)(512 + tid) * 16;
    if (o0 + 16 <= q8_limit) pre0 = __ldg(reinterpret_cast<const uint4 *>(q8 + o0));
    if (tid < 16 && o1 + 16 <= q8_limit) pre1 = __ldg(reinterpret_cast<const uint4 *>(q8 + o1));
  };
  auto expand = [&](const uint4 &v, int u4 /* uint4 index inside the group */) {
    const int d = u4 / (kQStride / 16), k0 = (u4 - d * (kQStride / 16)) * 16;
    float *dst = &S.patch[d][k0];
    const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int b = 0; b < 4; b++) dst[q * 4 + b] = __fmul_rn((float)((w[q] >> (8 * b)) & 0xFFu), 1.0f / 255.0f);  // cvConvertScale
  };
  prefetch(blockIdx.x);
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    __syncthreads();  // previous group's readers of patch / work are done
    int nd;
    bool skip = false;
    if (kRaw) {
      nd = min(16, n_items - grp * 16);
    } else {
      const b200_scan *sc = scans + grp;
      skip = !sc->usable;  // block-uniform: upside-down / vseg gate (frame.cpp:38-47)
      nd = min(16, (int)sc->hseg.n_offsets);
    }
    if (kRaw && raw_float != nullptr) {  // already-prepared float patches (model known-answer tests)
      if (warp < nd)
        for (int i = lane; i < 27 * 19; i += 32) S.patch[warp][i] = __ldg(raw_float + (size_t)(grp * 16 + warp) * (27 * 19) + i);
    } else if (!skip) {
      expand(pre0, tid);
      if (tid < 16) expand(pre1, 512 + tid);
    }
    prefetch(grp + gridDim.x);
    if (skip) continue;
    __syncthreads();  // patches complete
    // ---- C2, model by model
    for (int m = 0; m < 3; m++) {
      // conv 3x3 (valid, 24 x 15 computed) -> 3x3/3 max pool (8 x 5) -> + bias -> tanh.  Work item = (kernel half,
      // digit, pooled cell); the 5x5 input window of a cell is loaded once and feeds four kernels whose weights are
      // warp-uniform __constant__ operands.
      const int half_items = nd * 40;
      for (int it = tid; it < 2 * half_items; it += kCThreads) {
        const int kh = it >= half_items;
        const int dc = it - kh * half_items;
        const int d = dc / 40, cell = dc - d * 40;
        const int pr = cell / 5, pc = cell - pr * 5;
        const float *p = &S.patch[d][(pr * 3) * 19 + pc * 3];
        float win[5][5];
#pragma unroll
        for (int i = 0; i < 5; i++)
#pragma unroll
          for (int j = 0; j < 5; j++) win[i][j] = p[i * 19 + j];
        if (kh == 0) conv_pool_four<0>(W.conv, m, win, &S.work.feat[d][cell]);
        else conv_pool_four<4>(W.conv, m, win, &S.work.feat[d][cell]);
      }
      __syncthreads();
      // hidden layer 320 -> 32 as a [16 digits x 320] . [320 x 32] product: warp = K slice of 20 features, lane =
      // (digit lane dg = lane / 8, unit lane ug = lane % 8), thread = digits dg + 4i x units 4ug .. 4ug+3 (16 sums).
      // Per four features a thread issues 4 + 4 LDS.128 (weights: 128 B unique per request, features: 64 B) for 64 FMAs.
      {
        const int dg = lane >> 3, ug = lane & 7;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
        const float (*wT)[32] = S.hwT[m];
#pragma unroll
        for (int js = 0; js < 5; js++) {
          const int j = warp * 20 + js * 4;
          float4 f[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; i++) f[i] = *reinterpret_cast<const float4 *>(&S.work.feat[dg + 4 * i][j]);
#pragma unroll
          for (int t = 0; t < 4; t++) wv[t] = *reinterpret_cast<const float4 *>(&wT[j + t][4 * ug]);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float fv[4] = {f[i].x, f[i].y, f[i].z, f[i].w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
              acc[i][0] = fmaf(wv[t].x, fv[t], acc[i][0]);
              acc[i][1] = fmaf(wv[t].y, fv[t], acc[i][1]);
              acc[i][2] = fmaf(wv[t].z, fv[t], acc[i][2]);
              acc[i][3] = fmaf(wv[t].w, fv[t], acc[i][3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<float4 *>(&S.work.part[warp][dg + 4 * i][4 * ug]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
      __syncthreads();
      {
        const int u = tid & 31, d = tid >> 5;  // 16 digits x 32 units
        if (d < nd) {
          float sum = 0.0f;
#pragma unroll
          for (int q = 0; q < 16; q++) sum += S.work.part[q][d][u];
          S.work.hid[d][m][u] = tanh_sfu(sum + S.hb[m][u]);
        }
      }
      // (feat is rewritten by the next model's conv only after the barrier below; part after the one above)
      __syncthreads();
    }
    // logistic layer 32 -> 10 and softmax, thread -> (digit, model, class)
    for (int it = tid; it < nd * 30; it += kCThreads) {
      const int d = it / 30, r = it - d * 30, m = r / 10, c = r - m * 10;
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; j++) acc = fmaf(S.lw[m][c][j], S.work.hid[d][m][j], acc);
      S.work.prob[d][m][c] = expf(acc + S.lb[m][c]);
    }
    __syncthreads();
    for (int it = tid; it < 16 * 10; it += kCThreads) {
      const int d = it / 10, c = it - d * 10;
      float e = 0.0f, pm[3] = {0.0f, 0.0f, 0.0f};
      if (d < nd) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          float sum = 0.0f;
#pragma unroll
          for (int j = 0; j < 10; j++) sum += S.work.prob[d][m][j];
          pm[m] = S.work.prob[d][m][c] / sum;
        }
        const float mx = fmaxf(pm[0], fmaxf(pm[1], pm[2]));
        e = (((pm[0] + pm[1]) + pm[2]) - mx) / 2.0f;  // n_categorize.cpp:69-70
      }
      if (kRaw) {
        if (d < nd) {
          float *o = raw_out + (size_t)(grp * 16 + d) * 40;
          o[c] = e;
          o[10 + c] = pm[0];
          o[20 + c] = pm[1];
          o[30 + c] = pm[2];
        }
      } else {
        scans[grp].scores[d * 10 + c] = e;  // rows >= n_offsets stay 0 (NumberScores::Zero())
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// E0: expiry digit.  prepare_image_for_cat (scan/expiry_categorize.cpp:37-73) + applyc_bf4dd6c8
// (models/expiry/modelc_bf4dd6c8.cpp:12500-13505).  Secondary path (SCAN_EXPIRY builds only); four digits per CTA
// iteration, layer-2 weights (200 KB) streamed through shared memory one input map at a time.
// ------------------------------------------------------------------------------------------------
constexpr int kEDigits = 8;
constexpr int kEStage = 5;      // input maps per staged block of layer-2 kernels
constexpr int kEThreads = 320;  // layer 2: 8 digits x 40 output maps; layer 1: 8 x 70 pooled cells in two rounds

__constant__ float c_bil_color[256];  // bilateral colour LUT / spatial weights, computed on the host (b200_tables.cpp)
__constant__ float c_bil_space[5];    // mask order N, W, C, E, S

struct ExpirySmem {
  float c1w[50][25];
  float c1b[50];
  float xpad[kEDigits][24][20];   // mean-subtracted input with a 4-pixel zero border (full correlation), row stride 20
  float l1[kEDigits][50][70];     // ReLU(pool(conv1) + b)
  alignas(16) float wk[2][kEStage][40][28];  // layer-2 kernels of kEStage input maps, double-buffered (25 taps padded to 28)
  float c2[kEDigits][40][18];
  float l2[kEDigits][120];
  float hid[kEDigits][176];
  float o[kEDigits][10];
  float x[kEDigits][176];
  unsigned int hist[kEDigits][256];
  uint8_t raw[kEDigits][16 * 12], g8[kEDigits][16 * 12], lut[kEDigits][256];
};

__global__ void __launch_bounds__(kEThreads, 1)
expiry_kernel(const float *__restrict__ W /* modelc_bf4dd6c8 blob */, const uint8_t *__restrict__ patches,
              const float *__restrict__ prepared, int n, float *__restrict__ out, const int32_t *__restrict__ where) {
  // where != nullptr: `patches` holds whole 428x270 cards and crop i is the 16x11 window at (where[3i+1], where[3i+2]) of
  // card where[3i] -- prepare_image_for_cat's cvSetImageROI(rect->left, rect->top, 11, 16), expiry_categorize.cpp:41
  extern __shared__ __align__(16) uint8_t es_raw[];
  ExpirySmem &S = *reinterpret_cast<ExpirySmem *>(es_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *c2b = W + 51300, *hw = W + 51340, *hb = W + 72460, *lw = W + 72636, *lb = W + 74396;
  for (int i = tid; i < 1250; i += kEThreads) (&S.c1w[0][0])[i] = __ldg(W + i);
  for (int i = tid; i < 50; i += kEThreads) S.c1b[i] = __ldg(W + 1250 + i);

  const int n_groups = (n + kEDigits - 1) / kEDigits;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int nd = min(kEDigits, n - grp * kEDigits);
    __syncthreads();
    for (int i = tid; i < kEDigits * 24 * 20; i += kEThreads) (&S.xpad[0][0][0])[i] = 0.0f;
    __syncthreads();  // the preparation warps write the interior of xpad: every zero must have landed first
    // ---- patch preparation: one warp per digit
    if (warp < nd) {
      const int d = warp;
      const size_t idx = (size_t)grp * kEDigits + d;
      if (prepared != nullptr) {
        for (int i = lane; i < 176; i += 32) S.x[d][i] = __ldg(prepared + idx * 176 + i);
      } else {
        uint8_t *raw = S.raw[d], *g8 = S.g8[d];
        unsigned int *hist = S.hist[d];
        for (int i = lane; i < 256; i += 32) hist[i] = 0;
        if (where != nullptr) {
          const int ci = where[3 * idx], top = where[3 * idx + 1], left = where[3 * idx + 2];
          const bool inside = top >= 0 && top + 16 <= B200_CARD_H && left >= 0 && left + 11 <= B200_CARD_W;
          const uint8_t *card = patches + (size_t)ci * (B200_CARD_W * B200_CARD_H);
          for (int i = lane; i < 176; i += 32)
            raw[(i / 11) * 12 + (i % 11)] = inside ? __ldg(card + (top + i / 11) * B200_CARD_W + left + (i % 11)) : (uint8_t)0;
        } else {
          for (int i = lane; i < 176; i += 32) raw[(i / 11) * 12 + (i % 11)] = __ldg(patches + idx * 176 + i);
        }
        __syncwarp();
        for (int i = lane; i < 176; i += 32) {  // cvMorphologyEx(GRADIENT, 3x3 cross), replicate at the ROI edge
          const int y = i / 11, x = i - y * 11;
          const int yu = y > 0 ? y - 1 : y, yd = y < 15 ? y + 1 : y, xl = x > 0 ? x - 1 : x, xr = x < 10 ? x + 1 : x;
          const int a = raw[yu * 12 + x], b = raw[y * 12 + xl], c = raw[y * 12 + x], e = raw[y * 12 + xr], f = raw[yd * 12 + x];
          const int v = max(a, max(b, max(c, max(e, f)))) - min(a, min(b, min(c, min(e, f))));
          g8[y * 12 + x] = (uint8_t)v;
          atomicAdd(&hist[v], 1u);
        }
        __syncwarp();
        {  // llcv_equalize_hist: lut[i] = sat8(cvRound(cum(i) * (255.f / 176))), lut[0] = 0
          unsigned int local[8], run = 0;
#pragma unroll
          for (int q = 0; q < 8; q++) {
            run += hist[lane * 8 + q];
            local[q] = run;
          }
          unsigned int incl = run;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const unsigned int excl = incl - run;
          const float scale = 255.f / (11 * 16);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int val = __float2int_rn(__fmul_rn((float)(int)(excl + local[q]), scale));
            S.lut[d][lane * 8 + q] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
          }
          __syncwarp();
          if (lane == 0) S.lut[d][0] = 0;
          __syncwarp();
        }
        for (int i = lane; i < 176; i += 32) raw[(i / 11) * 12 + (i % 11)] = S.lut[d][g8[(i / 11) * 12 + (i % 11)]];
        __syncwarp();
        for (int i = lane; i < 176; i += 32) {  // cv::bilateralFilter(d = 3): mask N, W, C, E, S; float accumulation in that order
          const int y = i / 11, x = i - y * 11;
          const int yu = y > 0 ? y - 1 : y, yd = y < 15 ? y + 1 : y, xl = x > 0 ? x - 1 : x, xr = x < 10 ? x + 1 : x;
          const int v0 = raw[y * 12 + x];
          const int vals[5] = {raw[yu * 12 + x], raw[y * 12 + xl], v0, raw[y * 12 + xr], raw[yd * 12 + x]};
          float sum = 0.0f, wsum = 0.0f;
#pragma unroll
          for (int k = 0; k < 5; k++) {
            const float w = __fmul_rn(c_bil_space[k], c_bil_color[abs(vals[k] - v0)]);
            sum = __fadd_rn(sum, __fmul_rn((float)vals[k], w));
            wsum = __fadd_rn(wsum, w);
          }
          const int r = __float2int_rn(__fdiv_rn(sum, wsum));
          S.x[d][i] = __fmul_rn((float)(r < 0 ? 0 : (r > 255 ? 255 : r)), 1.0f / 255.0f);  // cvConvertScale
        }
      }
      __syncwarp();
      // normalized_input = input - input.mean()
      float part = 0.0f;
      for (int i = lane; i < 176; i += 32) part += S.x[d][i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const float mean = part / 176.0f;
      __syncwarp();
      for (int i = lane; i < 176; i += 32) S.xpad[d][4 + i / 11][4 + i % 11] = S.x[d][i] - mean;
    }
    __syncthreads();
    // ---- layer 1: item = (digit, pooled cell); the 6x6 input window feeds all 50 kernels (weights broadcast from smem)
    for (int it = tid; it < nd * 70; it += kEThreads) {
      const int d = it / 70, cell = it - d * 70, pr = cell / 7, pc = cell - pr * 7;
      float win[6][6];
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) win[i][j] = S.xpad[d][2 * pr + i][2 * pc + j];  // conv output (r, c) reads xpad rows r .. r+4
      for (int f = 0; f < 50; f++) {
        float a00 = 0.0f, a01 = 0.0f, a10 = 0.0f, a11 = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; i++)
#pragma unroll
          for (int j = 0; j < 5; j++) {
            const float w = S.c1w[f][i * 5 + j];
            a00 = fmaf(w, win[i][j], a00);
            a01 = fmaf(w, win[i][j + 1], a01);
            a10 = fmaf(w, win[i + 1][j], a10);
            a11 = fmaf(w, win[i + 1][j + 1], a11);
          }
        S.l1[d][f][cell] = fmaxf(fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + S.c1b[f], 0.0f);
      }
    }
    // ---- layer 2 (40 x 50 x 5 x 5 valid correlation on the 10 x 7 pooled maps -> 6 x 3): thread = (output map f, digit d)
    // with all 18 position sums in registers.  The kernels are staged kEStage input maps at a time with cp.async from the
    // [k][f][28] copy of the weights (double buffer: the next block is in flight during the FMAs of the current one, two
    // barriers per block).  Per input map a thread pulls its 25 weights (7 conflict-free LDS.128) and the digit's 70 inputs
    // (35 LDS.64) into registers and issues 450 FMAs on them.  (The first cut re-read a weight from shared memory for
    // every FMA; reading the weights straight from global memory touched 32 cache lines per request and was L1-bound.)
    {
      const int f = tid % 40, d = tid / 40;
      float acc[18];
#pragma unroll
      for (int q = 0; q < 18; q++) acc[q] = 0.0f;
      const float *c2k = W + B200_EXPIRY_C2K_OFFSET;
      auto stage = [&](int blk) {
        const float4 *src = reinterpret_cast<const float4 *>(c2k + (size_t)blk * kEStage * 40 * 28);
        float4 *dst = reinterpret_cast<float4 *>(&S.wk[blk & 1][0][0][0]);
        for (int i = tid; i < kEStage * 40 * 7; i += kEThreads) {
          const unsigned int sa = (unsigned int)__cvta_generic_to_shared(dst + i);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      stage(0);
      for (int blk = 0; blk < 50 / kEStage; blk++) {
        if (blk + 1 < 50 / kEStage) {
          stage(blk + 1);  // the other buffer: its readers passed the barrier that ended block blk - 1
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // block blk staged by every thread (and, for blk == 0, layer 1 complete)
        if (d < nd) {
#pragma unroll 1
          for (int kk = 0; kk < kEStage; kk++) {
            const int k = blk * kEStage + kk;
            float w[28], in[70];
            const float4 *wsrc = reinterpret_cast<const float4 *>(&S.wk[blk & 1][kk][f][0]);
#pragma unroll
            for (int t = 0; t < 7; t++) {
              const float4 v = wsrc[t];
              w[4 * t] = v.x, w[4 * t + 1] = v.y, w[4 * t + 2] = v.z, w[4 * t + 3] = v.w;
            }
            const float2 *src = reinterpret_cast<const float2 *>(&S.l1[d][k][0]);  // (d * 50 + k) * 70 floats: 8-byte aligned
#pragma unroll
            for (int t = 0; t < 35; t++) {
              const float2 v = src[t];
              in[2 * t] = v.x, in[2 * t + 1] = v.y;
            }
            // tap-major order: the 18 position sums are independent, so consecutive FMAs never wait on each other
#pragma unroll
            for (int i = 0; i < 5; i++)
#pragma unroll
              for (int j = 0; j < 5; j++)
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                  for (int c = 0; c < 3; c++) acc[r * 3 + c] = fmaf(w[i * 5 + j], in[(r + i) * 7 + c + j], acc[r * 3 + c]);
          }
        }
        __syncthreads();  // block blk consumed
      }
      if (d < nd) {
#pragma unroll
        for (int q = 0; q < 18; q++) S.c2[d][f][q] = acc[q];
      }
    }
    __syncthreads();
    for (int it = tid; it < nd * 120; it += kEThreads) {  // 2x3 max pool -> + bias -> ReLU; feature order [map][3]
      const int d = it / 120, q = it - d * 120, f = q / 3, r = q - f * 3;
      const float *p = &S.c2[d][f][2 * r * 3];
      const float m = fmaxf(fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3])), fmaxf(p[4], p[5]));
      S.l2[d][q] = fmaxf(m + __ldg(c2b + f), 0.0f);
    }
    __syncthreads();
    for (int it = tid; it < nd * 176; it += kEThreads) {  // hidden 120 -> 176, ReLU
      const int d = it / 176, i = it - d * 176;
      const float *w = hw + (size_t)i * 120;
      float a = 0.0f;
      for (int j = 0; j < 120; j++) a = fmaf(__ldg(w + j), S.l2[d][j], a);
      S.hid[d][i] = fmaxf(a + __ldg(hb + i), 0.0f);
    }
    __syncthreads();
    for (int it = tid; it < nd * 10; it += kEThreads) {  // logistic 176 -> 10
      const int d = it / 10, i = it - d * 10;
      const float *w = lw + (size_t)i * 176;
      float a = 0.0f;
      for (int j = 0; j < 176; j++) a = fmaf(__ldg(w + j), S.hid[d][j], a);
      S.o[d][i] = expf(a + __ldg(lb + i));
    }
    __syncthreads();
    for (int it = tid; it < nd * 10; it += kEThreads) {
      const int d = it / 10, i = it - d * 10;
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < 10; j++) sum += S.o[d][j];
      out[((size_t)grp * kEDigits + d) * 10 + i] = S.o[d][i] / sum;
    }
  }
}

int g_num_sms[64] = {0};  // per device

int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  int &n = g_num_sms[dev & 63];
  if (!n) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

template <typename K>
bool ensure_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

}  // namespace

void fill_conv_constants(const float *cnn_blobs[3], ConvConsts *out) {
  for (int m = 0; m < 3; m++) {
    for (int i = 0; i < 72; i++) (&out->w[m][0][0])[i] = cnn_blobs[m][i];
    for (int i = 0; i < 8; i++) out->b[m][i] = cnn_blobs[m][72 + i];
  }
}

static int launch_vseg_rows(const NetWeights &wts, const uint8_t *cards, const uint8_t *gate, const b200_scan *scans, int n,
                            int mode, float *vprob, const float *raw_rows, float *raw_out, cudaStream_t s) {
  // card rows go through the tensor-core kernel (vseg_mma.cu); raw float rows (the model tap) and B200_DMZ_VSEG_FP32=1
  // (A / B measurements) through the FP32 kernel of this file
  const char *fp32_env = getenv("B200_DMZ_VSEG_FP32");  // (read per call: tests flip it between calls)
  const bool force_fp32 = fp32_env && *fp32_env && atoi(fp32_env) != 0;
  if (mode != 2 && !force_fp32) return launch_vseg_rows_mma(wts, cards, gate, scans, n, mode, vprob, s);
  static PerDeviceOnce once;
  if (!once.ensure([] { return ensure_smem(vseg_rows_kernel, sizeof(VsegSmem)); })) return -1;
  const int per_frame = mode == 0 ? 68 : (mode == 1 ? kFineSlots : 1);
  const long long tiles = ((long long)n * per_frame + kVTile - 1) / kVTile;
  long long grid = (long long)num_sms();  // one persistent CTA (12 autonomous warps) per SM
  if (grid > (tiles + kVWarps - 1) / kVWarps) grid = (tiles + kVWarps - 1) / kVWarps;
  if (grid < 1) grid = 1;
  vseg_rows_kernel<<<(int)grid, kVThreads, sizeof(VsegSmem), s>>>(wts.vseg, wts.vseg_norm, cards, gate, scans, n, mode, vprob, raw_rows, raw_out);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_vseg_coarse_rows(const NetWeights &wts, const uint8_t *cards, int n, float *vprob, cudaStream_t s) {
  if (cudaMemsetAsync(vprob, 0, (size_t)n * 540 * sizeof(float), s) != cudaSuccess) return -1;
  return launch_vseg_rows(wts, cards, nullptr, nullptr, n, 0, vprob, nullptr, nullptr, s);
}

int launch_vseg_model(const NetWeights &wts, const float *rows, int n, float *out, cudaStream_t s) {
  return launch_vseg_rows(wts, nullptr, nullptr, nullptr, n, 2, nullptr, rows, out, s);
}

static int launch_categorize(const NetWeights &wts, const uint8_t *cards, b200_scan *scans, const uint8_t *raw,
                             const float *raw_float, int n, float *raw_out, uint8_t *q8, cudaStream_t s) {
  static PerDeviceOnce once;
  if (!once.ensure([] { return ensure_smem(categorize_kernel<false>, sizeof(CatSmem)) && ensure_smem(categorize_kernel<true>, sizeof(CatSmem)); }))
    return -1;
  const bool is_raw = raw != nullptr || raw_float != nullptr;
  const int groups = is_raw ? (n + 15) / 16 : n;
  int grid = num_sms();
  if (grid > groups) grid = groups;
  if (grid < 1) grid = 1;
  int launches = 1;
  if (raw_float == nullptr) {  // C0 + C1 into q8
    if (q8 == nullptr) return -1;
    const long long digits = is_raw ? n : (long long)n * 16;
    const int pgrid = (int)((digits + kPrepWarps - 1) / kPrepWarps);
    if (is_raw) digit_prep_kernel<true><<<pgrid, kPrepWarps * 32, 0, s>>>(nullptr, nullptr, raw, (int)digits, q8);
    else digit_prep_kernel<false><<<pgrid, kPrepWarps * 32, 0, s>>>(cards, scans, nullptr, (int)digits, q8);
    launches++;
  }
  // prepared byte patches go through the tensor-core kernel (categorize_mma.cu); float patches (the model tap) and
  // B200_DMZ_CNN_FP32=1 (A / B measurements) through the FP32 kernel of this file
  const char *fp32_env = getenv("B200_DMZ_CNN_FP32");
  if (raw_float == nullptr && !(fp32_env && *fp32_env && atoi(fp32_env) != 0)) {
    const int rc = launch_categorize_mma(wts, q8, scans, n, is_raw, raw_out, s);
    return rc < 0 ? -1 : launches - 1 + rc;
  }
  if (is_raw) categorize_kernel<true><<<grid, kCThreads, sizeof(CatSmem), s>>>(wts, raw_float ? nullptr : q8, nullptr, raw_float, n, raw_out);
  else categorize_kernel<false><<<grid, kCThreads, sizeof(CatSmem), s>>>(wts, q8, scans, nullptr, n, nullptr);
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int upload_bilateral_tables(const float *color256, const float *space5) {
  if (cudaMemcpyToSymbol(c_bil_color, color256, 256 * sizeof(float)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_bil_space, space5, 5 * sizeof(float)) != cudaSuccess) return -1;
  return 0;
}

int launch_expiry_digits(const float *weights, const uint8_t *patches, const float *prepared, int n, float *out, cudaStream_t s,
                         const int32_t *where) {
  // byte crops go through the tensor-core kernel (expiry_mma.cu); prepared float inputs (the per-layer KAT tap) and
  // B200_DMZ_EXPIRY_FP32=1 (A / B measurements) through the FP32 kernel of this file
  const char *fp32_env = getenv("B200_DMZ_EXPIRY_FP32");
  if (prepared == nullptr && !(fp32_env && *fp32_env && atoi(fp32_env) != 0)) return launch_expiry_digits_mma(weights, patches, n, out, s, where);
  static PerDeviceOnce once;
  if (!once.ensure([] { return ensure_smem(expiry_kernel, sizeof(ExpirySmem)); })) return -1;
  int grid = num_sms();
  const int groups = (n + kEDigits - 1) / kEDigits;
  if (grid > groups) grid = groups;
  if (grid < 1) grid = 1;
  expiry_kernel<<<grid, kEThreads, sizeof(ExpirySmem), s>>>(weights, patches, prepared, n, out, where);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_categorize_patches(const NetWeights &wts, const uint8_t *patches, const float *float_patches, int n, float *out, uint8_t *q8,
                              cudaStream_t s) {
  return launch_categorize(wts, nullptr, nullptr, patches, float_patches, n, out, q8, s);
}

// scan_card_image for a batch: gate -> vseg (coarse, select, fine, select) -> hseg -> categorize -> finish.
// vprob doubles as scratch: its tail holds the per-frame gate bytes.
// Lazy cards (lazy != nullptr): nobody asked for the 428 x 270 cards, so only the rows this sequence reads are warped, each
// set as soon as it is known -- the 68 coarse rows up front, the fine window once the coarse pass has chosen y0, and (rarely)
// the final number strip if it left that window.  lazy_ev (optional): 6 events bracketing the three warp launches.
int launch_scan(const NetWeights &wts, const uint8_t *cards, int n, const FrameGeom *geom, const uint8_t *valid,
                float *vprob, uint8_t *q8, b200_scan *scans, cudaStream_t s, cudaEvent_t ev_vseg, cudaEvent_t ev_hseg, cudaEvent_t ev_cat,
                cudaEvent_t ev_fin, const LazyWarp *lazy, uint8_t *lazy_cards, cudaEvent_t *lazy_ev) {
  int launches = 0, rc;
  uint8_t *gate = reinterpret_cast<uint8_t *>(vprob + (size_t)n * 540);
  uint16_t *coarse_y = reinterpret_cast<uint16_t *>(gate) + n;  // bytes [2n, 4n) of the 16n-byte tail; gate uses [0, n)
  auto lazy_warp = [&](int mode, int k) -> int {
    if (lazy_ev && cudaEventRecord(lazy_ev[2 * k], s) != cudaSuccess) return -1;
    const int r = launch_warp(lazy->src, geom, lazy_cards, nullptr, mode, scans, coarse_y, lazy->portrait, s);
    if (lazy_ev && cudaEventRecord(lazy_ev[2 * k + 1], s) != cudaSuccess) return -1;
    return r;
  };
#define STEP(call)          \
  rc = (call);              \
  if (rc < 0) return -1;    \
  launches += rc;
  if (ev_vseg) cudaEventRecord(ev_vseg, s);
  STEP(launch_scan_gate(geom, valid, n, gate, s));
  if (lazy) { STEP(lazy_warp(WARP_COARSE, 0)); }
  if (cudaMemsetAsync(vprob, 0, (size_t)n * 540 * sizeof(float), s) != cudaSuccess) return -1;
  STEP(launch_vseg_rows(wts, cards, gate, scans, n, 0, vprob, nullptr, nullptr, s));
  STEP(launch_vseg_select(vprob, gate, n, 0, scans, s, coarse_y));
  if (lazy) { STEP(lazy_warp(WARP_FINE, 1)); }
  STEP(launch_vseg_rows(wts, cards, gate, scans, n, 1, vprob, nullptr, nullptr, s));
  STEP(launch_vseg_select(vprob, gate, n, 1, scans, s));
  if (lazy) { STEP(lazy_warp(WARP_STRIP, 2)); }
  if (ev_hseg) cudaEventRecord(ev_hseg, s);
  STEP(launch_hseg(cards, n, scans, s));
  if (ev_cat) cudaEventRecord(ev_cat, s);
  STEP(launch_categorize(wts, cards, scans, nullptr, nullptr, n, nullptr, q8, s));
  if (ev_fin) cudaEventRecord(ev_fin, s);
  STEP(launch_scan_finish(n, scans, s));
#undef STEP
  return launches;
}
