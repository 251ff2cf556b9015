// card.io-dmz_b200/csrc/categorize_mma.cu -- C2 (the three digit CNNs) with both contractions on the tensor cores (tcgen05, sm_100a).
//
//   scores_for_number_image + applyc_{5c241121,01266c1b,b00bf70c}   scan/n_categorize.cpp:45-73, models/generated/modelc_*.cpp
//
// Input: the prepared (gradient + equalised) digit patches of digit_prep_kernel, 27 x 19 bytes q per digit; the reference feeds
// the networks x = fl(q * (1/255)).  Per model: conv 3x3 (8 kernels, 24 x 15 positions) -> 3x3/3 max pool (8 x 5 cells) -> + bias
// -> tanh -> 320 features -> hidden 320 -> 32, tanh -> logistic 32 -> 10, softmax; ensemble (r0 + r1 + r2 - max) / 2.
//
// conv + pool as ONE exact integer contraction per (tile of 128 pooled cells, model):
//   A  row = pooled cell, K = the 5 x 5 byte window of the patch that the cell's nine conv positions read (25 bytes + 7 zeros):
//      the window is the operand as it is, no float conversion;
//   B  column n = kernel * 9 + pool position, row k = window byte: the kernel's tap for that (position, byte) or 0 -- each
//      weight an integer Q = round(w * F), |Q| <= 2^20, written as three signed base-128 digits (three s8 matrices);
//   D  three s32 accumulators; sum_taps Q q = (d0 << 14) + (d1 << 7) + d2 exactly, the pool maximum is taken on these
//      integers (the scale is positive), then ONE multiply-add with (1/255) / F and the bias, and tanh.
// hidden layer as an fp16 contraction with split operands: feature = hi + lo, weight = Whi + Wlo (fp16 each, the low parts
// unscaled: their absolute precision, 3e-8, is that of an fp32 value near 0.5); hi Whi + lo Whi + hi Wlo accumulate in fp32
// (K = 3 x 320) -- the dropped lo Wlo term is 2^-24 relative.  M = 128 accumulator rows of which 32 are used (two frames x 16
// digit slots); the tensor pipe has nothing else to do.
//
// One persistent CTA per SM, 512 threads, a group of two frames per iteration:
//   stage the group's 32 x 528 prepared bytes (cp.async), build the <= 10 window tiles, then per model: conv units double
//   buffered in tensor memory (unit u + 1 runs while all 16 warps finish unit u: thread = (cell, kernel pair)), hidden MMAs,
//   hidden epilogue; finally the logistic layers and the ensemble on the CUDA cores.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "b200_internal.h"
#include "umma.cuh"

namespace {

#ifdef B200_CNN_TRACE
__device__ long long g_trace[3][2048];  // three tracer threads of block 0, private regions: no atomics in the timed path
#define TRACE(tag, cond)                                                                         \
  do {                                                                                           \
    if ((cond) && blockIdx.x == 0) {                                                             \
      const int who_ = threadIdx.x == 0 ? 0 : (threadIdx.x == 480 ? 1 : 2);                      \
      if (trace_n < 1000) g_trace[who_][2 * trace_n] = (tag), g_trace[who_][2 * trace_n + 1] = clock64(); \
      trace_n++;                                                                                 \
    }                                                                                            \
  } while (0)
#else
#define TRACE(tag, cond) do {} while (0)
#endif

constexpr int kConsumers = 512;            // 16 warps: window tiles, conv / hidden epilogues, logistic layers
constexpr int kThreads = kConsumers + 64;  // + warp 16: MMA issuer, warp 17: hidden-weight loader
constexpr int kG = 2;                      // frames per group
constexpr int kSlots = kG * 16;            // digit slots per group
constexpr int kQStride = B200_Q8_STRIDE;   // 528
constexpr int kConvN = 80;                 // 8 kernels x 9 pool positions = 72 columns, padded to a multiple of 16
constexpr int kFeatPart = 40 * kSlots * 16;   // 20 480: one fp16 part of the feature operand: [40 cells][32 slots][8 kernels]
constexpr int kHidBPart = 40 * 32 * 16;       // 20 480: one fp16 part of a model's hidden weights: [40 cells][32 units][8 kernels]
constexpr int kConvBBytes = 3 * 2 * (3 * kConvN) * 16;  // 23 040: [model][2 K chunks][3 digits x 80][16]

struct Smem {
  alignas(128) uint8_t feat[2][2 * kFeatPart];       // double buffered by model parity; each: hi part, lo part
  alignas(128) uint8_t hidb[2 * kHidBPart];          // Whi, Wlo of the current model (the MMA's unused rows 32 .. 127 of the
                                                     // feature operand read on into here: harmless)
  alignas(128) uint8_t conva[kG * 5 * 2 * 128 * 16];  // window tiles: [frame][tile][2 K chunks][128 cells][16]
  alignas(128) uint8_t convb[kConvBBytes];           // [model][2 K chunks][digit * 80 + column][16]
  alignas(16) uint8_t q8[kSlots * kQStride];
  float hid[3][32][kSlots];                          // [model][unit][slot]
  float prob[kSlots][3][10];
  float hb[3][32];
  float lw[3][32][10];                               // [model][unit][class]
  float lb[3][10];
  float cs[3][8], cbias[3][8];
  // mbarriers.  full / empty: conv accumulator slots; feat_ready: a model's features written (16 warps); hfull: hidden MMAs
  // done; hread: hidden accumulator read back (4 warps); a_ready: the group's window tiles built; hidb_ready: weights loaded
  alignas(8) unsigned long long full[2], empty[2], feat_ready, hfull, hread, a_ready, hidb_ready;
  uint32_t tmem;
};

__device__ __forceinline__ float tanh_sfu(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }  // as nets.cu
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory"); }
__device__ __forceinline__ void mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// kRaw: groups of 32 loose patches (stage tap: 10 ensemble scores + 3 x 10 model probabilities per patch) instead of two frames
template <bool kRaw>
__global__ void __launch_bounds__(kThreads, 1)
categorize_mma_kernel(const int8_t *__restrict__ convb_g, const float *__restrict__ convf_g, const __half *__restrict__ hidb_g,
                      const float *__restrict__ cnn0, const float *__restrict__ cnn1, const float *__restrict__ cnn2,
                      const uint8_t *__restrict__ q8, b200_scan *__restrict__ scans, int n_items /* frames, or raw patches */,
                      float *__restrict__ raw_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int trace_n = 0;
  (void)trace_n;
  const int lq = warp & 3, kq = (warp >> 2) & 3;  // consumers: TMEM lane quarter; kernel pair (conv) / unit octet (hidden)

  // ---- set-up
  for (int i = tid; i < kConvBBytes / 16; i += kThreads) reinterpret_cast<uint4 *>(S.convb)[i] = __ldg(reinterpret_cast<const uint4 *>(convb_g) + i);
  for (int i = tid; i < 48; i += kThreads) (i < 24 ? &S.cs[0][0] : &S.cbias[0][0])[i % 24] = __ldg(convf_g + i);
  for (int m = 0; m < 3; m++) {
    const float *b = m == 0 ? cnn0 : (m == 1 ? cnn1 : cnn2);
    for (int i = tid; i < 32; i += kThreads) S.hb[m][i] = __ldg(b + 80 + 10240 + i);
    for (int i = tid; i < 320; i += kThreads) S.lw[m][i % 32][i / 32] = __ldg(b + 80 + 10240 + 32 + i);
    for (int i = tid; i < 10; i += kThreads) S.lb[m][i] = __ldg(b + 80 + 10240 + 32 + 320 + i);
  }
  for (int i = tid; i < (int)sizeof(S.conva) / 16; i += kThreads) reinterpret_cast<uint4 *>(S.conva)[i] = make_uint4(0u, 0u, 0u, 0u);  // K padding stays 0
  for (int i = tid; i < (int)sizeof(S.feat) / 16; i += kThreads) reinterpret_cast<uint4 *>(&S.feat[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    umma::mbar_init(&S.full[0], 1), umma::mbar_init(&S.full[1], 1);
    umma::mbar_init(&S.empty[0], 16), umma::mbar_init(&S.empty[1], 16);
    umma::mbar_init(&S.feat_ready, 16), umma::mbar_init(&S.hfull, 1), umma::mbar_init(&S.hread, 4);
    umma::mbar_init(&S.a_ready, 1), umma::mbar_init(&S.hidb_ready, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&S.tmem, 512);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = S.tmem;  // conv unit slot s: columns [240 s, 240 s + 240) (digit j at + 80 j); hidden: [480, 512)

  const int per_group = kRaw ? kSlots : kG;  // items per group
  const int n_groups = (n_items + per_group - 1) / per_group;
  // digits to score in the two frames of a group (0: frame absent, not usable, or no patches) -- every role derives the same
  auto group_digits = [&](int grp, int &nd0, int &nd1) {
    int nd[2];
#pragma unroll
    for (int fi = 0; fi < kG; fi++) {
      nd[fi] = 0;
      if (kRaw) {
        nd[fi] = max(0, min(16, n_items - (grp * kSlots + fi * 16)));
      } else {
        const int f = grp * kG + fi;
        if (f < n_items && scans[f].usable) nd[fi] = min(16, (int)scans[f].hseg.n_offsets);  // upside-down / vseg gate (frame.cpp:38-47)
      }
    }
    nd0 = nd[0], nd1 = nd[1];
  };

  if (warp == 16) {
    // ================= MMA issuer (one thread) =================
    // Measured on B200 (ms per 100 k frames for the categorize stage): conv units double buffered + each model's 60 hidden
    // MMAs as ONE burst 11.1; hidden MMAs sliced between the next model's conv units 12.1; a warp-uniform issue loop with
    // an elected lane 14.1; hidden MMAs issued one at a time whenever the conv slots are busy 17.7 -- alternating between the
    // integer and the fp16 MMA kinds is what costs, so the two kinds are kept in long runs.  Conv units split into kernel
    // halves (N = 112, four accumulator slots instead of two, thread = (cell, one kernel)): 14.6 -- twice the barrier traffic
    // and MMAs for the same arithmetic; fewer, larger units win.
    if (lane == 0) {
      const uint32_t idesc_c = umma::instr_desc(umma::kAccS32, umma::kFmtU8, umma::kFmtS8, 128, 3 * kConvN);
      const uint32_t idesc_h = umma::instr_desc(umma::kAccF32, umma::kFmtF16, umma::kFmtF16, 128, 32);
      const uint64_t conva_d = umma::smem_desc(umma::smem_addr(S.conva), 2048, 128);
      const uint64_t convb_d = umma::smem_desc(umma::smem_addr(S.convb), 3 * kConvN * 16, 128);
      const uint64_t feat_d = umma::smem_desc(umma::smem_addr(&S.feat[0][0]), kSlots * 16, 128);
      const uint64_t hidb_d = umma::smem_desc(umma::smem_addr(S.hidb), 32 * 16, 128);
      uint32_t pe0 = 1, pe1 = 1, pfr = 0, phr = 1, pa = 0, phb = 0;  // (waiting on parity 1 of a fresh barrier passes at once)
      for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        int nd0, nd1;
        group_digits(grp, nd0, nd1);
        if (nd0 + nd1 == 0) continue;
        const int nt0 = (nd0 * 40 + 127) >> 7, nunits = nt0 + ((nd1 * 40 + 127) >> 7);
        umma::mbar_wait(&S.a_ready, pa), pa ^= 1u;
        umma::fence_after_sync();
        for (int m = 0; m < 3; m++) {
          const uint64_t bd = umma::desc_advance(convb_d, (m * 2) * (3 * kConvN * 16));
          for (int u = 0; u < nunits; u++) {
            if (u & 1) umma::mbar_wait(&S.empty[1], pe1), pe1 ^= 1u;
            else umma::mbar_wait(&S.empty[0], pe0), pe0 ^= 1u;
            umma::fence_after_sync();
            TRACE(100 + (u & 1), true);
            const int fi = u >= nt0, t = u - (fi ? nt0 : 0);
            umma::mma_i8(tmem + 240u * (u & 1), umma::desc_advance(conva_d, ((fi * 5 + t) * 2) * 2048), bd, idesc_c, 0u);
            umma::mma_commit(&S.full[u & 1]);
            TRACE(110 + (u & 1), true);
          }
          // hidden layer of model m: [32 slots x 960] . [960 x 32] into columns 480 .. 511
          umma::mbar_wait(&S.feat_ready, pfr), pfr ^= 1u;
          umma::mbar_wait(&S.hread, phr), phr ^= 1u;
          umma::mbar_wait(&S.hidb_ready, phb), phb ^= 1u;
          umma::fence_after_sync();
          TRACE(120, true);
          const uint64_t fa = umma::desc_advance(feat_d, (m & 1) * (2 * kFeatPart));
#pragma unroll 1
          for (int part = 0; part < 3; part++) {
            const uint64_t ab = umma::desc_advance(fa, part == 1 ? kFeatPart : 0), bb = umma::desc_advance(hidb_d, part == 2 ? kHidBPart : 0);
#pragma unroll 4
            for (int st = 0; st < 20; st++)
              umma::mma_f16(tmem + 480u, umma::desc_advance(ab, 2 * st * (kSlots * 16)), umma::desc_advance(bb, 2 * st * (32 * 16)), idesc_h,
                            (uint32_t)((part | st) != 0));
          }
          umma::mma_commit(&S.hfull);
          TRACE(121, true);
        }
      }
    }
  } else if (warp == 17) {
    // ================= hidden-weight loader =================
    uint32_t ph = 0;
    bool first = true;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      int nd0, nd1;
      group_digits(grp, nd0, nd1);
      if (nd0 + nd1 == 0) continue;
      for (int m = 0; m < 3; m++) {
        if (!first) umma::mbar_wait(&S.hfull, ph), ph ^= 1u;  // the previous hidden layer has read its weights
        first = false;
        const uint32_t dst = umma::smem_addr(S.hidb);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(hidb_g) + (size_t)m * 2 * kHidBPart;
        for (int i = lane; i < 2 * kHidBPart / 16; i += 32) cp_async16(dst + 16u * i, src + (size_t)i * 16);
        cp_async_wait_all();
        umma::fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.hidb_ready);
      }
    }
  } else {
    // ================= consumers =================
    const size_t q8_limit = (size_t)n_items * (kRaw ? kQStride : 16 * kQStride);
    auto prefetch = [&](int grp) {  // the group's 32 x 528 prepared bytes
      if (grp >= n_groups) return;
      const size_t base = (size_t)grp * kSlots * kQStride;
      const uint32_t dst = umma::smem_addr(S.q8);
      for (int i = tid; i < kSlots * kQStride / 16; i += kConsumers)
        if (base + (size_t)i * 16 + 16 <= q8_limit) cp_async16(dst + 16u * i, q8 + base + (size_t)i * 16);
    };
    uint32_t pf0 = 0, pf1 = 0, ph = 0;
    // hidden epilogue of model m: accumulator rows 0 .. 31 = TMEM lanes 0 .. 31 -> warps 0, 4, 8, 12 take eight units each
    auto hidden_finish = [&](int m) {
      TRACE(230, tid == 0);
      umma::mbar_wait(&S.hfull, ph), ph ^= 1u;  // (every consumer warp follows the barrier's phases)
      TRACE(231, tid == 0);
      if (lq != 0) return;
      umma::fence_after_sync();
      uint32_t v[8];
      tmem_ld8(tmem + 480u + 8u * kq, v);
      umma::tmem_ld_wait();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.hread);
#pragma unroll
      for (int i = 0; i < 8; i++) S.hid[m][8 * kq + i][lane] = tanh_sfu(__uint_as_float(v[i]) + S.hb[m][8 * kq + i]);
    };

    prefetch(blockIdx.x);
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      int nd0, nd1;
      group_digits(grp, nd0, nd1);
      cp_async_wait_all();
      consumer_barrier();  // q8 visible; the previous group's readers of hid / prob are done
      if (nd0 + nd1 == 0) {
        prefetch(grp + gridDim.x);
        continue;
      }
      // ---- window tiles: thread = (frame, cell index ci = digit * 40 + cell)
      for (int task = tid; task < kG * 640; task += kConsumers) {
        const int fi = task / 640, ci = task - fi * 640;
        if (ci >= (fi ? nd1 : nd0) * 40) continue;
        const int d = ci / 40, cell = ci - d * 40, cy = cell / 5, cx = cell - cy * 5;
        const uint8_t *p = S.q8 + (fi * 16 + d) * kQStride + (3 * cy) * 19 + 3 * cx;
        uint32_t wv[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int wy = 0; wy < 5; wy++)
#pragma unroll
          for (int wx = 0; wx < 5; wx++) {
            const int kk = wy * 5 + wx;
            wv[kk >> 2] |= (uint32_t)p[wy * 19 + wx] << (8 * (kk & 3));
          }
        uint8_t *dst = S.conva + ((fi * 5 + (ci >> 7)) * 2) * 2048 + (ci & 127) * 16;
        *reinterpret_cast<uint4 *>(dst) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        *reinterpret_cast<uint4 *>(dst + 2048) = make_uint4(wv[4], wv[5], wv[6], 0u);
      }
      umma::fence_async_smem();
      consumer_barrier();  // tiles complete, q8 consumed
      if (tid == 0) mbar_arrive(&S.a_ready);
      prefetch(grp + gridDim.x);
      const int nt0 = (nd0 * 40 + 127) >> 7, nunits = nt0 + ((nd1 * 40 + 127) >> 7);

      for (int m = 0; m < 3; m++) {
        if (m == 2) hidden_finish(0);  // also: model 0's hidden MMAs are done with feat[0], which model 2 now rewrites
        uint8_t *featm = &S.feat[m & 1][0];
        for (int u = 0; u < nunits; u++) {
          if (u & 1) umma::mbar_wait(&S.full[1], pf1), pf1 ^= 1u;
          else umma::mbar_wait(&S.full[0], pf0), pf0 ^= 1u;
          umma::fence_after_sync();
          TRACE(200 + (u & 1), tid == 0);
          TRACE(250 + (u & 1), tid == 480);
          // conv epilogue: thread = (cell = TMEM lane, kernels 2 kq and 2 kq + 1)
          const int fi = u >= nt0, t = u - (fi ? nt0 : 0);
          const int ci = 128 * t + 32 * lq + lane, d = ci / 40, cell = ci - d * 40;
          const uint32_t ta = tmem + ((uint32_t)(32 * lq) << 16) + 240u * (u & 1) + 18u * kq;
          uint32_t a0[16], a1[16], a2[16], b0[2], b1[2], b2[2];
          umma::tmem_ld16(ta, a0), tmem_ld2(ta + 16u, b0);
          umma::tmem_ld16(ta + kConvN, a1), tmem_ld2(ta + kConvN + 16u, b1);
          umma::tmem_ld16(ta + 2 * kConvN, a2), tmem_ld2(ta + 2 * kConvN + 16u, b2);
          umma::tmem_ld_wait();
          TRACE(210 + (u & 1), tid == 0);
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.empty[u & 1]);  // this warp has read its part of the slot
          if (d < (fi ? nd1 : nd0)) {
            int best[2];
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
              int mx = INT_MIN;
#pragma unroll
              for (int p = 0; p < 9; p++) {
                const int c = kk * 9 + p;
                const int q0 = (int)(c < 16 ? a0[c & 15] : b0[c & 1]), q1 = (int)(c < 16 ? a1[c & 15] : b1[c & 1]),
                          q2 = (int)(c < 16 ? a2[c & 15] : b2[c & 1]);
                mx = max(mx, (q0 << 14) + (q1 << 7) + q2);
              }
              best[kk] = mx;
            }
            const int k = 2 * kq;
            const float v0 = tanh_sfu(fmaf((float)best[0], S.cs[m][k], S.cbias[m][k]));
            const float v1 = tanh_sfu(fmaf((float)best[1], S.cs[m][k + 1], S.cbias[m][k + 1]));
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
            uint8_t *dst = featm + cell * (kSlots * 16) + (fi * 16 + d) * 16 + k * 2;
            *reinterpret_cast<__half2 *>(dst) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2 *>(dst + kFeatPart) = __halves2half2(l0, l1);
          }
        }
        TRACE(220, tid == 0);
        umma::fence_async_smem();  // this warp's features -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.feat_ready);
      }
      hidden_finish(1);
      hidden_finish(2);
      consumer_barrier();  // hid complete
      // ---- logistic layer 32 -> 10 and softmax, thread -> (slot, model, class)
      for (int it = tid; it < kSlots * 30; it += kConsumers) {
        const int s = it / 30, r = it - s * 30, m = r / 10, c = r - m * 10;
        if ((s & 15) >= (s < 16 ? nd0 : nd1)) continue;
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; j++) acc = fmaf(S.lw[m][j][c], S.hid[m][j][s], acc);
        S.prob[s][m][c] = expf(acc + S.lb[m][c]);
      }
      consumer_barrier();
      for (int it = tid; it < kSlots * 10; it += kConsumers) {
        const int s = it / 10, c = it - s * 10, fi = s >> 4, d = s & 15;
        const int nd = fi ? nd1 : nd0;
        if (nd == 0) continue;  // frame not scored: its record keeps the zeros it was initialised with
        float e = 0.0f, pm[3] = {0.0f, 0.0f, 0.0f};
        if (d < nd) {
#pragma unroll
          for (int m = 0; m < 3; m++) {
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 10; j++) sum += S.prob[s][m][j];
            pm[m] = S.prob[s][m][c] / sum;
          }
          const float mx = fmaxf(pm[0], fmaxf(pm[1], pm[2]));
          e = (((pm[0] + pm[1]) + pm[2]) - mx) / 2.0f;  // n_categorize.cpp:69-70
        }
        if (kRaw) {
          if (d < nd) {
            float *o = raw_out + (size_t)(grp * kSlots + s) * 40;
            o[c] = e, o[10 + c] = pm[0], o[20 + c] = pm[1], o[30 + c] = pm[2];
          }
        } else {
          scans[grp * kG + fi].scores[d * 10 + c] = e;  // rows >= n_offsets stay 0 (NumberScores::Zero())
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(S.tmem, 512);
}

}  // namespace

#ifdef B200_CNN_TRACE
extern "C" int b200_cnn_trace(long long *out, int cap) {  // out: 3 x 2048 stamps (tag, clock); unused entries 0
  (void)cap;
  cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 3 * 2048);
  cudaMemset(nullptr, 0, 0);
  static long long zeros[3 * 2048];
  cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
  return 3 * 1024;
}
#endif

int launch_categorize_mma(const NetWeights &wts, const uint8_t *q8, b200_scan *scans, int n, bool raw, float *raw_out, cudaStream_t s) {
  static PerDeviceOnce once;
  if (!once.ensure([] {
        return cudaFuncSetAttribute(categorize_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) == cudaSuccess &&
               cudaFuncSetAttribute(categorize_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) == cudaSuccess;
      }))
    return -1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = raw ? (n + kSlots - 1) / kSlots : (n + kG - 1) / kG;
  int grid = sms < groups ? sms : groups;
  if (grid < 1) grid = 1;
  const __half *hb = reinterpret_cast<const __half *>(wts.cnn_hidb);
  if (raw) categorize_mma_kernel<true><<<grid, kThreads, sizeof(Smem), s>>>(wts.cnn_convb, wts.cnn_convf, hb, wts.cnn[0], wts.cnn[1], wts.cnn[2], q8, nullptr, n, raw_out);
  else categorize_mma_kernel<false><<<grid, kThreads, sizeof(Smem), s>>>(wts.cnn_convb, wts.cnn_convf, hb, wts.cnn[0], wts.cnn[1], wts.cnn[2], q8, scans, n, nullptr);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
