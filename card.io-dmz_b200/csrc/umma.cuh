// card.io-dmz_b200/csrc/umma.cuh -- the few tcgen05 / TMEM / mbarrier primitives the batched contractions of this
// library need (sm_100a inline PTX; no CUTLASS).  Used where a weight x activation product is large enough to fill an
// MMA tile once frames are batched: the vseg hidden layer (M = 128 card rows, N = 64 units, K = 224 pixels).
//
// Operand layout (both operands K-major, no swizzle): a matrix of R rows x K bytes is stored as 16-byte chunks
//     chunk c (= 16 consecutive K bytes) of row r at   base + c * (R * 16) + r * 16
// i.e. "core matrices" of 8 rows x 16 bytes are 128 contiguous bytes, 8-row groups follow each other directly
// (stride-dimension byte offset SBO = 128) and the next K chunk starts R * 16 bytes later (leading-dimension byte
// offset LBO = R * 16).  A warp whose lanes hold consecutive rows therefore writes a chunk as one conflict-free 512-byte
// STS.128, and an MMA K step (32 bytes of K for 8-bit operands) reads chunks 2s and 2s + 1.
#ifndef B200_UMMA_CUH
#define B200_UMMA_CUH

#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!done);
}

// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory: one warp allocates (power of two >= 32 columns) and later frees
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot_in_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t tmem, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}

// ---- descriptors
// shared-memory matrix descriptor, K-major, no swizzle (layout above): rows = R
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, bits [0, 14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading-dimension byte offset, bits [16, 30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride-dimension byte offset, bits [32, 46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  return d;                                           // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor: dense, K-major A and B
enum { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2, kFmtU8 = 0, kFmtS8 = 1, kAccF32 = 1, kAccS32 = 2 };
__host__ __device__ constexpr uint32_t instr_desc(int acc_fmt, int a_fmt, int b_fmt, int m, int n) {
  return ((uint32_t)acc_fmt << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// One lane of a fully converged warp (warp-uniform control flow around the issue keeps descriptors and addresses in uniform
// registers: an `if (lane == 0)` region would pay an R2UR round trip per operand).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
// descriptor of the same matrix `bytes` further on (start-address field only; no carry out of its 14 bits for shared memory)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// ---- MMA issue (ONE thread): D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(void *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- accumulator read-back: warp w reads TMEM lanes 32 (w % 4) .. + 31 (lane = thread), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace umma

#endif
