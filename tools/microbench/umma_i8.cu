// tools/microbench/umma_i8.cu -- checks card.io-dmz_b200/csrc/umma.cuh on the device: D_j[128 x 64] (s32) = A[128 x 224] (u8) . B_j[64 x 224]^T (s8)
// for four weight terms j, through tcgen05.mma kind::i8 with no-swizzle K-major descriptors, read back with tcgen05.ld and
// compared with the host product.  Prints PASS / FAIL.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../card.io-dmz_b200/csrc
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "umma.cuh"

constexpr int M = 128, N = 64, K = 224, T = 4, KC = K / 16;

__global__ void __launch_bounds__(128) k(const uint8_t *__restrict__ A, const int8_t *__restrict__ B, int *__restrict__ D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sA = smem;                    // [KC][M][16]
  uint8_t *sB = smem + KC * M * 16;      // [T][KC][N][16]
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) umma::mbar_init(&bar, 1), umma::mbar_fence_init();
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  // operands into the canonical layout
  for (int i = tid; i < M * KC; i += 128) {
    const int r = i % M, c = i / M;
    *reinterpret_cast<uint4 *>(sA + (c * M + r) * 16) = *reinterpret_cast<const uint4 *>(A + r * K + c * 16);
  }
  for (int i = tid; i < T * N * KC; i += 128) {
    const int n = i % N, c = (i / N) % KC, t = i / (N * KC);
    *reinterpret_cast<uint4 *>(sB + ((t * KC + c) * N + n) * 16) = *reinterpret_cast<const uint4 *>(B + (t * N + n) * K + c * 16);
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma::instr_desc(umma::kAccS32, umma::kFmtU8, umma::kFmtS8, M, N);
    for (int t = 0; t < T; t++)
      for (int s = 0; s < K / 32; s++) {
        const uint64_t ad = umma::smem_desc(umma::smem_addr(sA) + 2 * s * (M * 16), M * 16, 128);
        const uint64_t bd = umma::smem_desc(umma::smem_addr(sB) + (t * KC + 2 * s) * (N * 16), N * 16, 128);
        umma::mma_i8(tmem + t * N, ad, bd, idesc, s > 0);
      }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int t = 0; t < T; t++)
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + t * N + c0, v);
      umma::tmem_ld_wait();
      for (int j = 0; j < 16; j++) D[(t * M + tid) * N + c0 + j] = (int)v[j];
    }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 256);
}

int main() {
  std::vector<uint8_t> A(M * K);
  std::vector<int8_t> B(T * N * K);
  srand(7);
  for (auto &v : A) v = (uint8_t)(rand() & 255);
  for (auto &v : B) v = (int8_t)((rand() & 255) - 128);
  uint8_t *dA;
  int8_t *dB;
  int *dD;
  cudaMalloc(&dA, A.size()), cudaMalloc(&dB, B.size()), cudaMalloc(&dD, T * M * N * 4);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice), cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, T * M * N * 4);
  const int smem = KC * M * 16 + T * KC * N * 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<int> D(T * M * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int t = 0; t < T; t++)
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        int ref = 0;
        for (int kk = 0; kk < K; kk++) ref += (int)A[m * K + kk] * (int)B[(t * N + n) * K + kk];
        if (ref != D[(t * M + m) * N + n]) {
          if (bad < 5) printf("mismatch t=%d m=%d n=%d: got %d want %d\n", t, m, n, D[(t * M + m) * N + n], ref);
          bad++;
        }
      }
  printf("%s: %ld mismatches of %d\n", bad ? "FAIL" : "PASS", bad, T * M * N);
  return bad != 0;
}
