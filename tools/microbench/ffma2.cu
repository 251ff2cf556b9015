// tools/microbench/ffma2.cu -- FP32 FMA issue rates on sm_100a: scalar FFMA (3 registers / uniform-register operand) against the packed
// FFMA2 (fma.rn.f32x2), which is what reaches the quoted 128 FMA / clk / SM.  Decides how the FP32-bound kernels (digit CNNs, vseg MLP,
// expiry CNN) should issue their multiply-adds.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int kAcc = 16, kIters = 4096;

// MODE 0: scalar FFMA, all operands in registers; 1: scalar FFMA, multiplier warp-uniform (kernel parameter -> UR / constant operand)
// MODE 2: FFMA2, register pairs; 3: FFMA2 with a uniform broadcast multiplier
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float u0, float u1, float u2, float u3) {
  float a[kAcc];
  float2 p[kAcc / 2];
  const float x = (float)threadIdx.x * 1e-3f, y = x + 0.5f;
  for (int i = 0; i < kAcc; i++) a[i] = x + i;
  for (int i = 0; i < kAcc / 2; i++) p[i] = make_float2(x + i, y + i);
  float m0 = x * 0.999f, m1 = y * 0.998f;
  const float us[4] = {u0, u1, u2, u3};
#pragma unroll 1
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < kAcc; i++) a[i] = fmaf(a[i], (i & 1) ? m0 : m1, m1);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < kAcc; i++) a[i] = fmaf(a[i], us[r], m1);
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < kAcc / 2; i++) p[i] = ffma2(p[i], make_float2(m0, m1), make_float2(m1, m0));
      } else {
#pragma unroll
        for (int i = 0; i < kAcc / 2; i++) p[i] = ffma2(p[i], make_float2(us[r], us[r]), make_float2(m1, m0));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < kAcc; i++) s += a[i];
  for (int i = 0; i < kAcc / 2; i++) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *d) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(d, 0.5f, 0.25f, 0.125f, 0.75f);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) k<MODE><<<blocks, 256>>>(d, 0.5f, 0.25f, 0.125f, 0.75f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = 5.0 * blocks * 256.0 * kIters * 4.0 * kAcc;
  printf("%-46s %8.2f TFLOP/s  %6.1f FMA/clk/SM (at %d MHz)\n", name, 2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
}

int main() {
  float *d;
  cudaMalloc(&d, 148 * 8 * 256 * 4 * 2);
  run<0>("FFMA  r, r, r", d);
  run<1>("FFMA  r, uniform, r", d);
  run<2>("FFMA2 rr, rr, rr", d);
  run<3>("FFMA2 rr, uniform broadcast, rr", d);
  return cudaDeviceSynchronize() != cudaSuccess;
}
