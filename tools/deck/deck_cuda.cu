// tools/deck/deck_cuda.cu -- GPU build of the synthetic deck generator (bench / test support, NOT product code).
// Compiled with -fmad=false so frames are bit-identical to tools/deck/deck_cpu.c (tests/test_deck.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "deck_gen.h"

__global__ void deck_params_kernel(uint64_t seed, uint32_t first, int n, int W, int H, double jitter, deck_params *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) deck_frame_params(seed, first + (uint32_t)i, W, H, jitter, &out[i]);
}

__global__ void deck_render_kernel(const deck_params *__restrict__ params, int W, int H, uint8_t *__restrict__ out) {
  __shared__ deck_params p;
  const int frame = blockIdx.y;
  if (threadIdx.x == 0) p = params[frame];
  __syncthreads();
  const int quads = W * H / 4;
  uint8_t *dst = out + (size_t)frame * W * H;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
    const int idx = q * 4, y = idx / W, x = idx - y * W;  // W % 4 == 0
    unsigned int v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v |= (unsigned)deck_pixel(&p, x + k, y) << (8 * k);
    reinterpret_cast<unsigned int *>(dst)[q] = v;
  }
}

// Render frames [first, first+n) of deck `seed` into the DEVICE buffer d_out (n dense W*H planes) on `stream`.
extern "C" int deck_render_cuda(uint64_t seed, uint32_t first, int n, int W, int H, double jitter, uint8_t *d_out, void *stream) {
  if (W % 4) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  deck_params *dp = nullptr;
  if (cudaMalloc(&dp, sizeof(deck_params) * (size_t)n) != cudaSuccess) return -2;
  deck_params_kernel<<<(n + 63) / 64, 64, 0, s>>>(seed, first, n, W, H, jitter, dp);
  for (int f0 = 0; f0 < n; f0 += 32768) {
    int cnt = n - f0 < 32768 ? n - f0 : 32768;
    deck_render_kernel<<<dim3(30, cnt), 256, 0, s>>>(dp + f0, W, H, d_out + (size_t)f0 * W * H);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(dp);
  return e == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -3;
}
