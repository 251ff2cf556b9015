// h2d_pitched.cu -- how fast can the upload rectangle of the host-buffer path (480 x 301 bytes out of every 640 x 480
// frame in pinned host memory) reach HBM?  Compares, on the same buffers:
//   contiguous   one cudaMemcpyAsync of the same number of bytes (the link's ceiling for this host)
//   pitched x1   one cudaMemcpy3DAsync per 1024-frame chunk on one stream (what api.cu does today)
//   pitched xS   the same chunks dealt round-robin to S streams (do several copy engines help?)
//   band         whole rows [y0, y1) of every frame: one contiguous 640 x 301 segment per frame (2-D copy)
//   zero-copy    a kernel that reads the rectangle straight out of the mapped pinned buffer (16-byte loads) and writes
//                it densely to HBM: no DMA descriptors at all
// usage: h2d_pitched [frames]      prints GB/s of useful (rectangle) bytes and frames/s for each
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);     \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int W = 640, H = 480, X0 = 80, Y0 = 89, CW = 480, CH = 301;

// one warp per row segment: 30 lanes move 16 bytes each (480 bytes), rows dealt to warps grid-stride
__global__ void __launch_bounds__(256) gather_rect_kernel(const uint8_t *__restrict__ host_frames, uint8_t *__restrict__ dst, int n) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t rows = (size_t)n * CH;
  for (size_t r = warp; r < rows; r += warps) {
    const size_t f = r / CH;
    const int y = (int)(r - f * CH);
    if (lane < CW / 16) {
      const uint4 v = *(reinterpret_cast<const uint4 *>(host_frames + f * (size_t)(W * H) + (size_t)(Y0 + y) * W + X0) + lane);
      *(reinterpret_cast<uint4 *>(dst + r * CW) + lane) = v;
    }
  }
}

// rows [yb, yb + yc) of the rectangle of n frames, read from mapped host memory; dst is the dense CW x CH rectangle per frame
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t *__restrict__ host_frames, uint8_t *__restrict__ dst, int n, int yb, int yc) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t rows = (size_t)n * yc;
  for (size_t r = warp; r < rows; r += warps) {
    const size_t f = r / yc;
    const int y = yb + (int)(r - f * yc);
    if (lane < CW / 16) {
      const uint4 v = *(reinterpret_cast<const uint4 *>(host_frames + f * (size_t)(W * H) + (size_t)(Y0 + y) * W + X0) + lane);
      *(reinterpret_cast<uint4 *>(dst + (f * CH + y) * (size_t)CW) + lane) = v;
    }
  }
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8192, chunk = 1024, reps = 3;
  uint8_t *h = nullptr, *d = nullptr;
  CK(cudaHostAlloc(&h, (size_t)n * W * H, cudaHostAllocDefault));
  memset(h, 1, (size_t)n * W * H);
  CK(cudaMalloc(&d, (size_t)n * W * H));
  cudaStream_t st[4];
  for (int i = 0; i < 4; i++) CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double rect_bytes = (double)n * CW * CH;
  auto report = [&](const char *name, double ms, double bytes) {
    printf("%-14s %8.2f ms  %6.2f GB/s moved  -> %8.0f frames/s\n", name, ms, bytes / ms / 1e6, n / (ms * 1e-3));
  };
  auto timed = [&](auto &&body) {
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, st[0]));
      body();
      for (int i = 1; i < 4; i++) {  // st[0] waits for the others so that e1 closes the whole batch
        cudaEvent_t j;
        CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
        CK(cudaEventRecord(j, st[i]));
        CK(cudaStreamWaitEvent(st[0], j, 0));
        CK(cudaEventDestroy(j));
      }
      CK(cudaEventRecord(e1, st[0]));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0 && ms < best) best = ms;
    }
    return (double)best;
  };
  auto pitched = [&](int f0, int cnt, cudaStream_t s) {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(h + (size_t)f0 * W * H, W, W, H);
    p.srcPos = make_cudaPos(X0, Y0, 0);
    p.dstPtr = make_cudaPitchedPtr(d + (size_t)f0 * CW * CH, CW, CW, CH);
    p.extent = make_cudaExtent(CW, CH, cnt);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3DAsync(&p, s));
  };
  report("contiguous", timed([&] { CK(cudaMemcpyAsync(d, h, (size_t)rect_bytes, cudaMemcpyHostToDevice, st[0])); }), rect_bytes);
  for (int S = 1; S <= 4; S *= 2) {
    char name[32];
    snprintf(name, sizeof(name), "pitched x%d", S);
    report(name, timed([&] {
             for (int f0 = 0, k = 0; f0 < n; f0 += chunk, k++) pitched(f0, n - f0 < chunk ? n - f0 : chunk, st[k % S]);
           }), rect_bytes);
  }
  report("band 640x301", timed([&] {
           CK(cudaMemcpy2DAsync(d, (size_t)W * CH, h + (size_t)Y0 * W, (size_t)W * H, (size_t)W * CH, n, cudaMemcpyHostToDevice, st[0]));
         }), (double)n * W * CH);
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
    char name[32];
    snprintf(name, sizeof(name), "zero-copy x%d", per_sm);
    report(name, timed([&] { gather_rect_kernel<<<sms * per_sm, 256, 0, st[0]>>>(h, d, n); CK(cudaGetLastError()); }), rect_bytes);
  }
  // does the rectangle's start alignment / segment length matter?  (x0, width) variants that all contain the columns the path needs
  {
    const int variants[][2] = {{80, 480}, {64, 496}, {64, 512}, {0, 576}, {64, 576}, {80, 496}, {85, 469}, {84, 472}};
    for (auto &v : variants) {
      const int x0 = v[0], cw = v[1];
      char name[32];
      snprintf(name, sizeof(name), "x0=%d w=%d", x0, cw);
      report(name, timed([&] {
               for (int f0 = 0; f0 < n; f0 += chunk) {
                 const int cnt = n - f0 < chunk ? n - f0 : chunk;
                 cudaMemcpy3DParms p;
                 memset(&p, 0, sizeof(p));
                 p.srcPtr = make_cudaPitchedPtr(h + (size_t)f0 * W * H, W, W, H);
                 p.srcPos = make_cudaPos(x0, Y0, 0);
                 p.dstPtr = make_cudaPitchedPtr(d + (size_t)f0 * cw * CH, cw, cw, CH);
                 p.extent = make_cudaExtent(cw, CH, cnt);
                 p.kind = cudaMemcpyHostToDevice;
                 CK(cudaMemcpy3DAsync(&p, st[0]));
               }
             }), (double)n * cw * CH);
    }
  }
  // hybrid: the DMA engine is bound by rows per second (above: 469 .. 512-byte rows all take 10 ns each), the link is not
  // full -- let SMs fetch the last k rows of every rectangle straight from the mapped host buffer while the DMA moves the rest
  for (int k : {0, 16, 24, 32, 40, 48, 64, 96, 150}) {
    for (int ctas : {32, 148}) {
      if (k == 0 && ctas != 32) continue;
      char name[32];
      snprintf(name, sizeof(name), "hybrid k=%d c=%d", k, ctas);
      report(name, timed([&] {
               for (int f0 = 0; f0 < n; f0 += chunk) {
                 const int cnt = n - f0 < chunk ? n - f0 : chunk;
                 cudaMemcpy3DParms p;
                 memset(&p, 0, sizeof(p));
                 p.srcPtr = make_cudaPitchedPtr(h + (size_t)f0 * W * H, W, W, H);
                 p.srcPos = make_cudaPos(X0, Y0, 0);
                 p.dstPtr = make_cudaPitchedPtr(d + (size_t)f0 * CW * CH, CW, CW, CH);
                 p.extent = make_cudaExtent(CW, CH - k, cnt);
                 p.kind = cudaMemcpyHostToDevice;
                 CK(cudaMemcpy3DAsync(&p, st[0]));
                 if (k) gather_rows_kernel<<<ctas, 256, 0, st[1]>>>(h + (size_t)f0 * W * H, d + (size_t)f0 * CW * CH, cnt, CH - k, k);
               }
             }), rect_bytes);
    }
  }
  // zero-copy split over two streams (two kernels at once, half the frames each)
  report("zero-copy 2str", timed([&] {
           gather_rect_kernel<<<sms * 2, 256, 0, st[0]>>>(h, d, n / 2);
           gather_rect_kernel<<<sms * 2, 256, 0, st[1]>>>(h + (size_t)(n / 2) * W * H, d + (size_t)(n / 2) * CW * CH, n - n / 2);
         }), rect_bytes);
  return 0;
}
