// card.io-dmz_b200/csrc/expiry_mma.cu -- E0, the expiry digit CNN, with its one large contraction on the tensor cores.
//
//   prepare_image_for_cat + applyc_bf4dd6c8      scan/expiry_categorize.cpp:37-109, models/expiry/modelc_bf4dd6c8.cpp:12500-13505
//
// Layer 2 (40 maps x 50 input maps x 5x5, valid correlation on the 10 x 7 pooled maps -> 6 x 3) is 0.9 M of the network's
// 1.3 M multiply-adds per crop.  Batched it is a contraction with M = (crop, output position), N = 40 maps, K = (tap, input
// map) = 25 x 50: SEVEN crops fill one M = 128 tile (7 x 18 = 126 rows, every row a wanted output).  The activations are
// floats (ReLU outputs), so the operands are fp16 pairs: value = hi + lo, weight = Whi + Wlo, and
// hi Whi + lo Whi + hi Wlo accumulate in fp32 in tensor memory (the dropped lo Wlo is 2^-22 relative).
// K is walked in 13 slices of two taps (2 x 56 halfs = 7 K steps of tcgen05.mma kind::f16, 21 MMAs per slice): a slice of
// the operand is an im2col copy -- row (crop, position) <- the 112-byte map vector of pixel (position + tap) of the
// layer-1 output, which layer 1 writes pixel-major as fp16 hi / lo -- built by all threads with 16-byte copies while the
// weights of the next slice stream in (cp.async, double buffered).  No garbage rows, no float conversion in the loop.
//
// Everything else stays on the CUDA cores, as in expiry_kernel (nets.cu): patch preparation (cross gradient, histogram
// equalisation, 3x3 bilateral: exact bytes), layer 1 (50 maps, 5x5 full correlation + 2x2 pool: one 6x6 window per
// (crop, pooled cell)), the 2x3 pool after layer 2, hidden 120 -> 176, logistic 176 -> 10, softmax.
// One persistent CTA per SM, 512 threads, seven crops per iteration.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200_internal.h"
#include "umma.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kCrops = 7;               // per iteration: 7 x 18 output positions = 126 rows of one M = 128 tile
constexpr int kRows = kCrops * 70;      // layer-1 output pixels of a batch
constexpr int kMapHalfs = 56;           // 50 input maps padded to 7 x 8 halfs (112 bytes per pixel and part)
constexpr int kSlices = 13;             // two taps per slice (the 26th tap has zero weights)
constexpr int kSliceChunks = 14;        // 2 taps x 7 16-byte chunks
constexpr int kN = 48;                  // 40 output maps padded to a multiple of 16
constexpr int kAPart = kSliceChunks * 128 * 16;   // 28 672: one fp16 part of an operand slice
constexpr int kBPart = kSliceChunks * kN * 16;    // 10 752: one fp16 part of a weight slice

__constant__ float c_bil_color2[256];  // (the bilateral tables are input independent; this unit keeps its own copy)
__constant__ float c_bil_space2[5];

struct PrepScratch {  // per crop, during preparation only
  unsigned int hist[256];
  uint8_t raw[16 * 12], g8[16 * 12], lut[256];
};
struct Smem {
  alignas(128) uint8_t a[2 * kAPart];        // operand slice: hi part, lo part.  Before layer 2: preparation scratch + the
                                             // zero-padded inputs; after it: the 126 x 40 layer-2 sums
  alignas(128) uint8_t b[2][2 * kBPart];     // weight slices (Whi, Wlo), double buffered
  alignas(16) __half l1h[kRows][kMapHalfs];  // layer-1 output, pixel-major: hi part
  alignas(16) __half l1l[kRows][kMapHalfs];  // lo part
  float c1w[50][25];
  float c1b[50];
  float x[kCrops][176];
  float l2[kCrops][120];
  float hid[kCrops][176];
  float o[kCrops][10];
  alignas(8) unsigned long long bar;
  uint32_t tmem;
};
struct Early {  // lives in Smem::a until layer 1 is done
  float xpad[kCrops][24][20];  // mean-subtracted input with a 4-pixel zero border (full correlation), row stride 20
  PrepScratch prep[kCrops];
};
static_assert(sizeof(Early) <= 2 * kAPart, "early-phase scratch must fit the operand slice");
static_assert(126 * 40 * 4 <= 24576 && 24576 + 2 * kCrops * 176 * 4 <= 2 * kAPart, "layer-2 sums and hidden partial sums must fit the operand slice");

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

__global__ void __launch_bounds__(kThreads, 1)
expiry_mma_kernel(const float *__restrict__ W /* modelc_bf4dd6c8 blob */, const __half *__restrict__ w2h /* layer-2 weight slices */,
                  const uint8_t *__restrict__ patches, int n, float *__restrict__ out, const int32_t *__restrict__ where) {
  // where != nullptr: `patches` holds whole 428x270 cards and crop i is the 16x11 window at (where[3i+1], where[3i+2]) of
  // card where[3i] -- prepare_image_for_cat's cvSetImageROI(rect->left, rect->top, 11, 16), expiry_categorize.cpp:41
  extern __shared__ __align__(128) uint8_t es_raw[];
  Smem &S = *reinterpret_cast<Smem *>(es_raw);
  Early &E = *reinterpret_cast<Early *>(S.a);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *c2b = W + 51300, *hb = W + 72460, *lw = W + 72636, *lb = W + 74396;
  const float *hwT = W + B200_EXPIRY_HWT_OFFSET;  // hidden weights transposed: [120][176]
  for (int i = tid; i < 1250; i += kThreads) (&S.c1w[0][0])[i] = __ldg(W + i);
  for (int i = tid; i < 50; i += kThreads) S.c1b[i] = __ldg(W + 1250 + i);
  // map padding of the layer-1 output (halfs 50 .. 55) stays zero for the kernel's lifetime
  for (int i = tid; i < kRows * kMapHalfs; i += kThreads) (&S.l1h[0][0])[i] = __float2half_rn(0.0f), (&S.l1l[0][0])[i] = __float2half_rn(0.0f);
  if (tid == 0) umma::mbar_init(&S.bar, 1), umma::mbar_fence_init();
  if (warp == 0) umma::tmem_alloc(&S.tmem, 64);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = S.tmem;
  const uint32_t idesc = umma::instr_desc(umma::kAccF32, umma::kFmtF16, umma::kFmtF16, 128, kN);
  uint32_t phase = 0;

  auto load_weights = [&](int slice) {  // 21.5 KB of fp16 weights of one slice into buffer slice & 1
    const uint32_t dst = umma::smem_addr(S.b[slice & 1]);
    const uint8_t *src = reinterpret_cast<const uint8_t *>(w2h) + (size_t)slice * (2 * kBPart);
    for (int i = tid; i < 2 * kBPart / 16; i += kThreads) cp_async16(dst + 16u * i, src + (size_t)i * 16);
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per slice
  };

  const int n_groups = (n + kCrops - 1) / kCrops;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int nd = min(kCrops, n - grp * kCrops);
    __syncthreads();  // the previous iteration's readers of a / l2 / hid / o are done
    load_weights(0);  // (arrives long before layer 2 needs it)
    for (int i = tid; i < kCrops * 24 * 20; i += kThreads) (&E.xpad[0][0][0])[i] = 0.0f;
    __syncthreads();  // the preparation warps write the interior of xpad: every zero must have landed first
    // ---- patch preparation: one warp per crop (as expiry_kernel, nets.cu)
    if (warp < nd) {
      const int d = warp;
      const size_t idx = (size_t)grp * kCrops + d;
      uint8_t *raw = E.prep[d].raw, *g8 = E.prep[d].g8, *lut = E.prep[d].lut;
      unsigned int *hist = E.prep[d].hist;
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
          lut[lane * 8 + q] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
        }
        __syncwarp();
        if (lane == 0) lut[0] = 0;
        __syncwarp();
      }
      for (int i = lane; i < 176; i += 32) raw[(i / 11) * 12 + (i % 11)] = lut[g8[(i / 11) * 12 + (i % 11)]];
      __syncwarp();
      for (int i = lane; i < 176; i += 32) {  // cv::bilateralFilter(d = 3): mask N, W, C, E, S; float accumulation in that order
        const int y = i / 11, x = i - y * 11;
        const int yu = y > 0 ? y - 1 : y, yd = y < 15 ? y + 1 : y, xl = x > 0 ? x - 1 : x, xr = x < 10 ? x + 1 : x;
        const int v0 = raw[y * 12 + x];
        const int vals[5] = {raw[yu * 12 + x], raw[y * 12 + xl], v0, raw[y * 12 + xr], raw[yd * 12 + x]};
        float sum = 0.0f, wsum = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; k++) {
          const float w = __fmul_rn(c_bil_space2[k], c_bil_color2[abs(vals[k] - v0)]);
          sum = __fadd_rn(sum, __fmul_rn((float)vals[k], w));
          wsum = __fadd_rn(wsum, w);
        }
        const int r = __float2int_rn(__fdiv_rn(sum, wsum));
        S.x[d][i] = __fmul_rn((float)(r < 0 ? 0 : (r > 255 ? 255 : r)), 1.0f / 255.0f);  // cvConvertScale
      }
      __syncwarp();
      // normalized_input = input - input.mean()
      float part = 0.0f;
      for (int i = lane; i < 176; i += 32) part += S.x[d][i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const float mean = part / 176.0f;
      __syncwarp();
      for (int i = lane; i < 176; i += 32) E.xpad[d][4 + i / 11][4 + i % 11] = S.x[d][i] - mean;
    }
    __syncthreads();
    // ---- layer 1: item = (crop, pooled cell); the 6x6 input window feeds all 50 kernels (weights broadcast from smem).
    // Output pixel-major as fp16 hi + lo: the operand rows of layer 2 are copies of these 112-byte map vectors.
    for (int it = tid; it < nd * 70; it += kThreads) {
      const int d = it / 70, cell = it - d * 70, pr = cell / 7, pc = cell - pr * 7;
      float win[6][6];
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) win[i][j] = E.xpad[d][2 * pr + i][2 * pc + j];  // conv output (r, c) reads xpad rows r .. r+4
      __half2 *oh = reinterpret_cast<__half2 *>(&S.l1h[d * 70 + cell][0]), *ol = reinterpret_cast<__half2 *>(&S.l1l[d * 70 + cell][0]);
      for (int f = 0; f < 50; f += 2) {
        float v[2];
#pragma unroll
        for (int ff = 0; ff < 2; ff++) {
          float a00 = 0.0f, a01 = 0.0f, a10 = 0.0f, a11 = 0.0f;
#pragma unroll
          for (int i = 0; i < 5; i++)
#pragma unroll
            for (int j = 0; j < 5; j++) {
              const float w = S.c1w[f + ff][i * 5 + j];
              a00 = fmaf(w, win[i][j], a00);
              a01 = fmaf(w, win[i][j + 1], a01);
              a10 = fmaf(w, win[i + 1][j], a10);
              a11 = fmaf(w, win[i + 1][j + 1], a11);
            }
          v[ff] = fmaxf(fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + S.c1b[f + ff], 0.0f);
        }
        const __half h0 = __float2half_rn(v[0]), h1 = __float2half_rn(v[1]);
        oh[f >> 1] = __halves2half2(h0, h1);
        ol[f >> 1] = __halves2half2(__float2half_rn(v[0] - __half2float(h0)), __float2half_rn(v[1] - __half2float(h1)));
      }
    }
    __syncthreads();  // layer 1 complete; the early-phase scratch in `a` is dead
    // rows 126, 127 of the operand (no crop) and the rows of absent crops read as zero
    for (int i = tid; i < 2 * kAPart / 16; i += kThreads) reinterpret_cast<uint4 *>(S.a)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // ---- layer 2 on the tensor cores: 13 K slices of two taps
    const __half *task_src = nullptr;  // this thread's copy task: map vector of its row's pixel at tap (0, 0), hi or lo part
    uint8_t *task_dst = nullptr;
    int task_ts = 0;
    if (tid < 126 * 4) {
      const int r = tid >> 2, part = (tid >> 1) & 1, d = r / 18, pos = r - d * 18, pr = pos / 3, pc = pos - pr * 3;
      task_ts = tid & 1;
      if (d < nd) {
        task_src = part ? &S.l1l[d * 70 + pr * 7 + pc][0] : &S.l1h[d * 70 + pr * 7 + pc][0];
        task_dst = S.a + part * kAPart + (task_ts * 7) * (128 * 16) + r * 16;
      }
    }
    for (int slice = 0; slice < kSlices; slice++) {
      if (slice + 1 < kSlices) load_weights(slice + 1);  // (buffer (slice + 1) & 1: its MMAs of slice - 1 completed before the previous wait)
      // operand slice: row r = (crop, position (pr, pc)) <- map vector of pixel (pr + i, pc + j) for the slice's two taps;
      // thread = (row, part, tap slot) for the whole batch (504 of the 512 threads), seven 16-byte copies per slice
      if (task_src != nullptr) {
        const int tap = 2 * slice + task_ts;
        if (tap < 25) {  // (slice 12: its second tap has zero weights; the stale operand bytes there are finite)
          const int ti = tap / 5, tj = tap - ti * 5;
          const uint4 *src = reinterpret_cast<const uint4 *>(task_src + (ti * 7 + tj) * kMapHalfs);
          uint4 v[7];
#pragma unroll
          for (int k = 0; k < 7; k++) v[k] = src[k];
#pragma unroll
          for (int k = 0; k < 7; k++) *reinterpret_cast<uint4 *>(task_dst + k * (128 * 16)) = v[k];
        }
      }
      // this slice's weights have landed (its group was committed one slice ago); the next slice's may still be in flight
      if (slice + 1 < kSlices) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      umma::fence_async_smem();  // operand and weight bytes -> visible to the tensor core
      umma::fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        umma::fence_after_sync();
        const uint64_t ah = umma::smem_desc(umma::smem_addr(S.a), 128 * 16, 128), al = umma::desc_advance(ah, kAPart);
        const uint64_t bh = umma::smem_desc(umma::smem_addr(S.b[slice & 1]), kN * 16, 128), bl = umma::desc_advance(bh, kBPart);
#pragma unroll 1
        for (int term = 0; term < 3; term++) {
          const uint64_t ad = term == 1 ? al : ah, bd = term == 2 ? bl : bh;
#pragma unroll
          for (int ks = 0; ks < 7; ks++)
            umma::mma_f16(tmem, umma::desc_advance(ad, 2 * ks * (128 * 16)), umma::desc_advance(bd, 2 * ks * (kN * 16)), idesc,
                          (uint32_t)((slice | term | ks) != 0));
        }
        umma::mma_commit(&S.bar);
      }
      umma::mbar_wait(&S.bar, phase);  // the slice's MMAs are done: the operand buffer may be rebuilt
      phase ^= 1u;
      umma::fence_after_sync();
    }
    // ---- layer-2 sums out of tensor memory: thread = row = (crop, position), 40 maps
    float *c2 = reinterpret_cast<float *>(S.a);  // [126][40]
    if (warp < 4) {
      const int r = 32 * warp + lane;
#pragma unroll
      for (int c0 = 0; c0 < 40; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        umma::tmem_ld_wait();
        if (r < 126) {
#pragma unroll
          for (int q = 0; q < 8; q++) c2[r * 40 + c0 + q] = __uint_as_float(v[q]);
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    for (int it = tid; it < nd * 120; it += kThreads) {  // 2x3 max pool -> + bias -> ReLU; feature order [map][3]
      const int d = it / 120, q = it - d * 120, f = q / 3, r = q - f * 3;
      const float *p = c2 + (d * 18 + 2 * r * 3) * 40 + f;  // positions (2r, 0..2), (2r + 1, 0..2)
      const float m = fmaxf(fmaxf(fmaxf(p[0], p[40]), fmaxf(p[80], p[120])), fmaxf(p[160], p[200]));
      S.l2[d][q] = fmaxf(m + __ldg(c2b + f), 0.0f);
    }
    __syncthreads();
    // hidden 120 -> 176, ReLU: thread = (unit i, half of the 120 inputs) for all seven crops; the weights are read from the
    // transposed copy [j][176] (one coalesced 128-byte line per warp and j) and used seven times
    float (*hpart)[kCrops][176] = reinterpret_cast<float (*)[kCrops][176]>(S.a + 24576);  // (behind the layer-2 sums, which are dead by now)
    if (tid < 352) {
      const int i = tid % 176, half = tid / 176;
      float acc[kCrops];
#pragma unroll
      for (int d = 0; d < kCrops; d++) acc[d] = 0.0f;
      const float *w = hwT + (size_t)(60 * half) * 176 + i;
#pragma unroll 6
      for (int j = 0; j < 60; j++) {
        const float wv = __ldg(w + (size_t)j * 176);
#pragma unroll
        for (int d = 0; d < kCrops; d++) acc[d] = fmaf(wv, S.l2[d][60 * half + j], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < kCrops; d++) hpart[half][d][i] = acc[d];
    }
    __syncthreads();
    for (int it = tid; it < nd * 176; it += kThreads) {
      const int d = it / 176, i = it - d * 176;
      S.hid[d][i] = fmaxf((hpart[0][d][i] + hpart[1][d][i]) + __ldg(hb + i), 0.0f);
    }
    __syncthreads();
    // logistic 176 -> 10: one warp per (crop, class), lanes over the 176 inputs
    for (int it = warp; it < nd * 10; it += kThreads / 32) {
      const int d = it / 10, i = it - d * 10;
      const float *w = lw + (size_t)i * 176;
      float a = 0.0f;
      for (int j = lane; j < 176; j += 32) a = fmaf(__ldg(w + j), S.hid[d][j], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) S.o[d][i] = expf(a + __ldg(lb + i));
    }
    __syncthreads();
    for (int it = tid; it < nd * 10; it += kThreads) {
      const int d = it / 10, i = it - d * 10;
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < 10; j++) sum += S.o[d][j];
      out[((size_t)grp * kCrops + d) * 10 + i] = S.o[d][i] / sum;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(S.tmem, 64);
}

}  // namespace

int upload_bilateral_tables_mma(const float *color256, const float *space5) {
  if (cudaMemcpyToSymbol(c_bil_color2, color256, 256 * sizeof(float)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_bil_space2, space5, 5 * sizeof(float)) != cudaSuccess) return -1;
  return 0;
}

int launch_expiry_digits_mma(const float *weights, const uint8_t *patches, int n, float *out, cudaStream_t s, const int32_t *where) {
  static PerDeviceOnce once;
  if (!once.ensure([] {
        return cudaFuncSetAttribute(expiry_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) == cudaSuccess;
      }))
    return -1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = (n + kCrops - 1) / kCrops;
  int grid = sms < groups ? sms : groups;
  if (grid < 1) grid = 1;
  const __half *w2h = reinterpret_cast<const __half *>(weights + B200_EXPIRY_C2H_OFFSET);
  expiry_mma_kernel<<<grid, kThreads, sizeof(Smem), s>>>(weights, w2h, patches, n, out, where);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
