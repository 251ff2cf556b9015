// card.io-dmz_b200/csrc/vseg_mma.cu -- V1 + V2 for card rows with the hidden layer on the tensor cores (tcgen05, sm_100a).
//
//   vseg_probabilities_for_hstrip (scan/n_vseg.cpp:39-47): llcv_morph_grad3_1d_u8 -> llcv_lineardown2_1d_u8 ->
//   llcv_norm_convert_1d_u8_to_f32 -> applym_befe75da (models/generated/modelm_befe75da.cpp:1770)
//
// Once frames are batched the hidden layer is a real contraction: [rows x 204] . [204 x 50], 68 + ~32 rows per frame.
// The reference feeds it x_k = fl(fl(fl(v_k / 255) * scale) + shift), v_k the 8-bit down-sampled gradient of the row and
// (scale, shift) derived from the row's min mn and max mx; in real numbers x_k = (v_k - mn) * s + d0 with
//     s = (1/255) * scale,   d0 = shift + mn * s   (what the float rounding of `shift` leaves over; ~1e-8)
// so   W1 x = s * sum_k W1[u][k] (v_k - mn) + d0 * sum_k W1[u][k].
// The sum over k has INTEGER activations (0 .. 255), and each weight is written as S_u * 2^-27 * Q with a 28-bit integer Q
// (S_u = max_k |W1[u][k]|) = four signed base-128 digits, so the contraction is four exact u8 x s8 -> s32 MMAs
// (tcgen05.mma kind::i8, M = 128 rows, N = 64 units, K = 224) whose accumulators recombine to sum_k Q_k (v_k - mn)
// with no rounding at all: the only float roundings left are the three of the recombination and the final scale.  That is
// closer to the real-number value than any FP32 summation order, the reference's included (weights are represented to
// 2^-28 of the unit's largest one).
//
// One persistent CTA per SM, 256 threads = two independent groups of 128 (thread = card row), each with its own A tile,
// accumulator columns, mbarrier and named barrier, so one group's MMA wait is covered by the other group's arithmetic:
//   stage   the warp copies its 32 rows (412 bytes each, card columns 8 .. 419) into shared memory with coalesced cp.async
//   prep    thread = row: two pixels per 16-bit lane pair (VIMNMX3.U16x2), 204 outputs kept in 51 registers, min / max on
//           the fly, subtract mn, thirteen conflict-free STS.128 into the K-major operand tile (umma.cuh layout)
//   mma     one thread: 4 digits x 7 K steps of tcgen05.mma, tcgen05.commit -> mbarrier; the next tile's rows are staged
//   finish  thread = row = TMEM lane: tcgen05.ld the four digit sums of 16 units at a time, recombine, scale, tanh,
//           logistic layer, softmax; two floats per row to vprob
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200_internal.h"
#include "umma.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTileRows = 128;
constexpr int kTerms = 4;
constexpr int kChunks = 14;                           // K = 224 bytes = 204 outputs + 20 zeros
constexpr int kUnits = 64;                            // 50 hidden units + 14 zero columns (N % 16 == 0 for M = 128)
constexpr int kWBytes = kTerms * kChunks * kUnits * 16;  // 57 344
constexpr int kATileBytes = kChunks * kTileRows * 16;    // 28 672
constexpr int kRawPitch = 424;                        // staged row: 412 bytes used; 8-byte aligned, 53 8-byte units (odd)
constexpr int kFineSlots = 33;                        // rows of [y0 - 8, y0 + 35) that are not multiples of four: <= 33
constexpr size_t kCardBytes = (size_t)B200_CARD_W * B200_CARD_H;

struct Smem {
  alignas(128) uint8_t w[kWBytes];
  alignas(128) uint8_t a[2][kATileBytes];
  alignas(16) uint8_t raw[kThreads * kRawPitch];
  alignas(16) VsegUnit unit[kUnits];
  float b2[4];
  alignas(8) unsigned long long bar[2];
  uint32_t tmem;
};

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kTileRows) : "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}

// the (frame, card row) of work item `item`; frame = -1 if there is nothing to score
__device__ __forceinline__ void resolve_item(long long item, long long total, int mode, const uint8_t *__restrict__ gate,
                                             const b200_scan *__restrict__ scans, int &f, int &row) {
  f = -1, row = 0;
  if (item >= total) return;
  if (mode == 0) {
    const int fr = (int)(item / 68), j = (int)(item - (long long)fr * 68);
    if (!gate || gate[fr]) f = fr, row = 4 * j;  // coarse rows 0, 4, .., 268 (n_vseg.cpp:127-137)
  } else {
    const int fr = (int)(item / kFineSlots), j = (int)(item - (long long)fr * kFineSlots);
    if (gate && !gate[fr]) return;
    const int y0 = scans[fr].vseg.y_offset;  // coarse best (vseg_select pass 0); 0xFFFF: none
    if (y0 == 0xFFFF) return;
    const int lo = y0 < 8 ? 0 : y0 - 8, hi = min(B200_CARD_H, y0 + 27 + 8);  // n_vseg.cpp:140-142
    // the j-th row >= lo that is not a multiple of four (those were scored by the coarse pass)
    const int base = lo & ~3, skip = max(0, (lo & 3) - 1), i = j + skip;
    const int r = base + 4 * (i / 3) + 1 + (i % 3);
    if (r < hi) f = fr, row = r;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
vseg_mma_kernel(const int8_t *__restrict__ wq, const VsegUnit *__restrict__ unitf, const float *__restrict__ b2, const float2 *__restrict__ sd_tab,
                const uint8_t *__restrict__ cards, const uint8_t *__restrict__ gate, const b200_scan *__restrict__ scans, int n, int mode,
                float *__restrict__ vprob) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 7, gt = tid & 127, gw = warp & 3;

  // ---- set-up: weights (already in the operand layout), unit constants, zero K padding, barriers, tensor memory
  for (int i = tid; i < kWBytes / 16; i += kThreads) reinterpret_cast<uint4 *>(S.w)[i] = __ldg(reinterpret_cast<const uint4 *>(wq) + i);
  for (int i = tid; i < kUnits * 2; i += kThreads) reinterpret_cast<uint4 *>(S.unit)[i] = __ldg(reinterpret_cast<const uint4 *>(unitf) + i);
  if (tid < 3) S.b2[tid] = __ldg(b2 + tid);
  reinterpret_cast<uint4 *>(S.a[g] + 13 * (kTileRows * 16))[gt] = make_uint4(0u, 0u, 0u, 0u);  // K bytes 208 .. 223 (never rewritten)
  if (tid == 0) {
    umma::mbar_init(&S.bar[0], 1);
    umma::mbar_init(&S.bar[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&S.tmem, 512);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = S.tmem + 256u * (uint32_t)g;  // this group's accumulator columns: digit t at + 64 t

  const int per_frame = mode == 0 ? 68 : kFineSlots;
  const long long total = (long long)n * per_frame;
  const long long ntiles = (total + kTileRows - 1) / kTileRows;
  const long long tstride = (long long)gridDim.x * 2;
  uint8_t *my_raw = S.raw + (size_t)tid * kRawPitch;
  const uint32_t raw_warp = umma::smem_addr(S.raw + (size_t)(tid - lane) * kRawPitch);

  int f, row;  // this thread's row of the tile being staged / processed
  // stage: the warp copies its 32 rows, one row per iteration, 256 contiguous bytes per request
  auto stage = [&](long long tile) {
    resolve_item(tile * kTileRows + gt, total, mode, gate, scans, f, row);
#pragma unroll 4
    for (int rr = 0; rr < 32; rr++) {
      const int ff = __shfl_sync(0xffffffffu, f, rr), rw = __shfl_sync(0xffffffffu, row, rr);
      if (ff < 0) continue;  // warp-uniform
      const uint8_t *src = cards + (size_t)ff * kCardBytes + (size_t)rw * B200_CARD_W + 8;  // card column 8: 2 before the ROI, 4-byte aligned
      const uint32_t dst = raw_warp + (uint32_t)rr * kRawPitch;
      if ((rw & 1) == 0) {  // even rows start 8-byte aligned: 52 x 8 bytes
        cp_async8(dst + 8u * lane, src + 8 * lane);
        if (lane < 20) cp_async8(dst + 8u * (lane + 32), src + 8 * (lane + 32));
      } else {              // 104 x 4 bytes
        cp_async4(dst + 4u * lane, src + 4 * lane);
        cp_async4(dst + 4u * (lane + 32), src + 4 * (lane + 32));
        cp_async4(dst + 4u * (lane + 64), src + 4 * (lane + 64));
        if (lane < 8) cp_async4(dst + 4u * (lane + 96), src + 4 * (lane + 96));
      }
    }
  };

  const long long t0 = (long long)blockIdx.x * 2 + g;
  uint32_t phase = 0;
  if (t0 < ntiles) stage(t0);
  for (long long tile = t0; tile < ntiles; tile += tstride) {
    const int my_f = f, my_row = row;
    cp_async_wait_all();
    __syncwarp();
    // ---- prep.  ROI pixel j (card column 10 + j) is staged byte j + 2.  With E[m] = px 2m, O[m] = px 2m + 1, staged word i
    // holds (E[2i-1], O[2i-1], E[2i], O[2i]); output k needs a = O[k-1], b = E[k], c = O[k], d = E[k+1]:
    //   g0 = max(a,b,c) - min(a,b,c), g1 = max(b,c,d) - min(b,c,d)  (3-tap max - min, replicate at the ROI edge),  v = (g0 + g1 + 1) >> 1
    // Outputs (2i, 2i+1) ride in the two 16-bit lanes of one register.
    float s_row = 0.0f, d0_row = 0.0f;
    if (my_f >= 0) {
      uint32_t out[51];
      uint32_t mn2 = 0x00FF00FFu, mx2 = 0u;
      const uint2 *rp = reinterpret_cast<const uint2 *>(my_raw);
      uint2 cur = rp[0];
      cur.x = __byte_perm(cur.x, 0u, 0x3220);  // px -1 := px 0
      uint32_t e0 = cur.x & 0x00FF00FFu, o0 = (cur.x >> 8) & 0x00FF00FFu;
#pragma unroll
      for (int j = 0; j < 51; j++) {
        uint2 nxt = rp[j + 1];                                  // words 2j + 2, 2j + 3
        if (j == 50) nxt.x = __byte_perm(nxt.x, 0u, 0x3110);    // px 408 := px 407
        const uint32_t e1 = cur.y & 0x00FF00FFu, o1 = (cur.y >> 8) & 0x00FF00FFu;
        const uint32_t e2 = nxt.x & 0x00FF00FFu, o2 = (nxt.x >> 8) & 0x00FF00FFu;
        uint32_t v[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t a = h ? o1 : o0, d = h ? e2 : e1;
          const uint32_t b = __byte_perm(h ? e1 : e0, h ? e2 : e1, 0x5432), c = __byte_perm(h ? o1 : o0, h ? o2 : o1, 0x5432);
          const uint32_t g0 = __vimax3_u16x2(a, b, c) - __vimin3_u16x2(a, b, c);
          const uint32_t g1 = __vimax3_u16x2(b, c, d) - __vimin3_u16x2(b, c, d);
          v[h] = ((g0 + g1 + 0x00010001u) >> 1) & 0x00FF00FFu;
        }
        mn2 = __vimin3_u16x2(mn2, v[0], v[1]);
        mx2 = __vimax3_u16x2(mx2, v[0], v[1]);
        out[j] = __byte_perm(v[0], v[1], 0x6420);
        cur = nxt, e0 = e2, o0 = o2;
      }
      const uint32_t mn = min(mn2 & 0xFFFFu, mn2 >> 16), mx = max(mx2 & 0xFFFFu, mx2 >> 16);
      const float2 sd = __ldg(sd_tab + (mn * 256u + mx));
      s_row = sd.x, d0_row = sd.y;
      const uint32_t sub = mn * 0x01010101u;  // every byte >= mn: no borrows
      uint8_t *dst = S.a[g] + gt * 16;
#pragma unroll
      for (int c = 0; c < 13; c++) {
        const uint32_t w3 = c < 12 ? out[4 * c + 3] - sub : 0u;  // outputs 204 .. 207 are K padding
        *reinterpret_cast<uint4 *>(dst + c * (kTileRows * 16)) = make_uint4(out[4 * c] - sub, out[4 * c + 1] - sub, out[4 * c + 2] - sub, w3);
      }
    }
    umma::fence_async_smem();  // operand tile -> visible to the tensor core's (async proxy) reads
    umma::fence_before_sync();
    group_barrier(g);          // tile complete; staged rows consumed; previous accumulators read
    if (gt == 0) {
      umma::fence_after_sync();
      const uint32_t idesc = umma::instr_desc(umma::kAccS32, umma::kFmtU8, umma::kFmtS8, kTileRows, kUnits);
      const uint32_t a_addr = umma::smem_addr(S.a[g]), w_addr = umma::smem_addr(S.w);
#pragma unroll 1
      for (int t = 0; t < kTerms; t++)
#pragma unroll
        for (int ks = 0; ks < kChunks / 2; ks++)
          umma::mma_i8(tmem + 64u * t, umma::smem_desc(a_addr + 2 * ks * (kTileRows * 16), kTileRows * 16, 128),
                       umma::smem_desc(w_addr + (t * kChunks + 2 * ks) * (kUnits * 16), kUnits * 16, 128), idesc, ks > 0);
      umma::mma_commit(&S.bar[g]);
    }
    if (tile + tstride < ntiles) stage(tile + tstride);  // overlaps the MMAs and the finish below
    umma::mbar_wait(&S.bar[g], phase);
    phase ^= 1u;
    umma::fence_after_sync();
    // ---- finish: thread = row = TMEM lane
    {
      const uint32_t tbase = tmem + ((uint32_t)(gw * 32) << 16);
      float z0 = 0.0f, z1 = 0.0f, z2 = 0.0f;
      auto unit_step = [&](int u, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
        // sum_k Q_k (v_k - mn) = ((q0 * 128 + q1) * 128 + q2) * 128 + q3; |q0| <= 204 * 255 * 64 < 2^22: exact in float
        const float T = fmaf(fmaf(fmaf((float)(int)q0, 128.0f, (float)(int)q1), 128.0f, (float)(int)q2), 128.0f, (float)(int)q3);
        const float4 ua = *reinterpret_cast<const float4 *>(&S.unit[u].cu);
        const float2 ub = *reinterpret_cast<const float2 *>(&S.unit[u].w21);
        const float zin = fmaf(T, ua.x * s_row, fmaf(d0_row, ua.y, ua.z));
#ifndef B200_VSEG_TANH_SFU
        const float hv = tanhf(zin);  // feeds an arg-max index: accurate tanh (the SFU form below measured 3.07 against 3.15 ms: not worth it)
#else
        // tanh on the special-function unit: 1 - 2 / (e^(2|z|) + 1) with the sign copied back; absolute error <= ~1.5e-7
        // everywhere (ex2.approx is 2^-22 relative, the quotient is <= 1), the size of tanhf's own 1 - 2 ulp near +-1.  The
        // row probabilities feed an arg-max: index-identical to the FP32 kernel (which uses tanhf) on the 100 000-frame deck
        // (test_tensor_core_vseg_equals_fp32_vseg) and <= 1e-5 from the oracle row by row.
        const float hv = copysignf(1.0f - __fdividef(2.0f, __expf(2.0f * fabsf(zin)) + 1.0f), zin);
#endif
        z0 = fmaf(ua.w, hv, z0), z1 = fmaf(ub.x, hv, z1), z2 = fmaf(ub.y, hv, z2);
      };
#pragma unroll 1
      for (int ug = 0; ug < 3; ug++) {
        uint32_t q0[16], q1[16], q2[16], q3[16];
        umma::tmem_ld16(tbase + 16u * ug, q0);
        umma::tmem_ld16(tbase + 64u + 16u * ug, q1);
        umma::tmem_ld16(tbase + 128u + 16u * ug, q2);
        umma::tmem_ld16(tbase + 192u + 16u * ug, q3);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; q++) unit_step(16 * ug + q, q0[q], q1[q], q2[q], q3[q]);
      }
      {
        uint32_t q0[2], q1[2], q2[2], q3[2];  // units 48, 49
        tmem_ld2(tbase + 48u, q0), tmem_ld2(tbase + 64u + 48u, q1), tmem_ld2(tbase + 128u + 48u, q2), tmem_ld2(tbase + 192u + 48u, q3);
        umma::tmem_ld_wait();
        unit_step(48, q0[0], q1[0], q2[0], q3[0]);
        unit_step(49, q0[1], q1[1], q2[1], q3[1]);
      }
      if (my_f >= 0) {
        // softmax as the generated model does it: expf / sum, no max shift
        const float e0 = expf(z0 + S.b2[0]), e1 = expf(z1 + S.b2[1]), e2 = expf(z2 + S.b2[2]);
        const float sum = (e0 + e1) + e2;
        *reinterpret_cast<float2 *>(vprob + ((size_t)my_f * 270 + my_row) * 2) = make_float2(e1 / sum, e2 / sum);  // visa-like, amex-like
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(S.tmem, 512);
}

}  // namespace

int launch_vseg_rows_mma(const NetWeights &wts, const uint8_t *cards, const uint8_t *gate, const b200_scan *scans, int n, int mode, float *vprob,
                         cudaStream_t s) {
  static PerDeviceOnce once;
  if (!once.ensure([] {
        return cudaFuncSetAttribute(vseg_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) == cudaSuccess;
      }))
    return -1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long total = (long long)n * (mode == 0 ? 68 : kFineSlots);
  const long long tiles = (total + kTileRows - 1) / kTileRows;
  long long grid = sms;  // one persistent CTA per SM, two tiles in flight each
  if (grid > (tiles + 1) / 2) grid = (tiles + 1) / 2;
  if (grid < 1) grid = 1;
  vseg_mma_kernel<<<(int)grid, kThreads, sizeof(Smem), s>>>(wts.vseg_wq, reinterpret_cast<const VsegUnit *>(wts.vseg_unit), wts.vseg + 10400,
                                                           reinterpret_cast<const float2 *>(wts.vseg_sd), cards, gate, scans, n, mode, vprob);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
