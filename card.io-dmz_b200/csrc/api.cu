// card.io-dmz_b200/csrc/api.cu -- the extern "C" boundary (include/b200_dmz.h): context, device memory,
// host<->device staging and the per-call kernel sequences.  No CPU fallback exists: if CUDA is
// unavailable every entry point fails with B200_ECUDA.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <exception>
#include <new>
#include <string>
#include <vector>

#include "b200_internal.h"

namespace {
constexpr int kDetectThreads = 416;  // must match detect.cu
constexpr size_t kCardBytes = (size_t)B200_CARD_W * B200_CARD_H;
}  // namespace

// Device scratch of one in-flight batch.  A context owns two lanes so that the host-buffer path can overlap
// the H2D copy of chunk k+1 with the kernels of chunk k (separate streams).
struct Lane {
  cudaStream_t stream = nullptr;
  int cap = 0, cap_w = 0, cap_h = 0;
  size_t frame_bytes = 0;   // size of d_frames (host-buffer staging)
  size_t chroma_bytes = 0;  // size of d_cb and of d_cr (tracked on their own: (w/2)(h/2) is not w*h/4 for odd sizes)
  uint8_t *d_frames = nullptr, *d_cb = nullptr, *d_cr = nullptr;
  b200_line *d_lines = nullptr;  // 3 * cap * 4
  FrameGeom *d_geom = nullptr;
  uint8_t *d_cards = nullptr;
  float *d_vprob = nullptr;
  uint8_t *d_q8 = nullptr;  // prepared digit patches, 16 x B200_Q8_STRIDE bytes per frame
  b200_scan *d_scan = nullptr;
  b200_frame_record *d_records = nullptr;
  unsigned int *d_check = nullptr;
  uint8_t *d_flags = nullptr;
  int16_t *d_grad = nullptr;
  size_t grad_elems = 0;
};

enum { ST_DETECT = 0, ST_GEOMETRY, ST_WARP, ST_VSEG, ST_HSEG, ST_CATEGORIZE, ST_FINALIZE, ST_COUNT };

struct b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // == lane[0].stream
  std::string error;
  uint64_t launches = 0;
  uint64_t full_frame_redos = 0;           // host-buffer frames that had to be re-uploaded whole (crop too small)
  uint64_t h2d_bytes = 0, d2h_bytes = 0;  // bytes moved by the host-buffer entry points
  Lane lane[2];
  int host_chunk = 1024;  // frames per pipelined chunk on the host-buffer path (measured: 1024 > 2048 > 4096)
  int crop_margin = 2;    // host-buffer path uploads only the detection region + this margin (< 0: whole frames)
  int card_mode = 0;      // 0: cards are materialised only when the caller asks for them (lazy rows otherwise); 1: always
  cudaEvent_t lazy_ev[6] = {nullptr};  // brackets of the three lazy warp launches (profiling)
  // per-stage device time (CUDA events on the lane stream), accumulated while profiling is on
  int profiling = 0;
  double stage_ms[ST_COUNT] = {0};
  uint64_t stage_frames = 0;
  bool lazy_timed = false;
  cudaEvent_t ev[ST_COUNT + 1] = {nullptr};

  // weights
  float *d_vseg = nullptr;
  float *d_cnn[3] = {nullptr, nullptr, nullptr};
  float *d_hwT = nullptr;
  float *d_vnorm = nullptr;   // (min, max) -> scale / shift table of the vseg row normalisation
  int8_t *d_vseg_wq = nullptr;  // tensor-core form of the vseg hidden layer: weight digits, per-unit constants, (s, d0) table
  float *d_vseg_unit = nullptr, *d_vseg_sd = nullptr;
  int8_t *d_cnn_convb = nullptr;  // tensor-core form of the digit CNNs: conv tap digits, per-kernel scale / bias, fp16 hidden weights
  float *d_cnn_convf = nullptr;
  uint16_t *d_cnn_hidb = nullptr;
  float *d_expiry = nullptr;  // modelc_bf4dd6c8 blob (optional: E0 entry points need it)
  float *d_slash = nullptr;   // modelm_730c4cbd blob (optional: b200_best_expiry_seg_batch needs it)
  NetWeights wts{};

  // geometry cache
  int cfg_w = 0, cfg_h = 0, cfg_orient = 0, cfg_planes = 0;
  DetectParams dp[2];  // Y, chroma
  GeomParams gp;

  // small staging buffers for host-mode outputs
  void *d_misc = nullptr;
  size_t misc_bytes = 0;
  void *h_pinned = nullptr;
  size_t pinned_bytes = 0;
};

namespace {

int fail(b200_ctx *c, int code, const char *fmt, ...) noexcept {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) {
    try {
      c->error = buf;
    } catch (...) {  // (out of memory while recording the message: the code still says what happened)
    }
  }
  return code;
}

// Every extern "C" entry point is a function-try-block closed by this: no C++ exception (std::bad_alloc from the host
// vectors and strings, anything else) crosses the C ABI -- include/b200_dmz.h promises "never throw".
#define B200_GUARD(ctx_expr)                                                                              \
  catch (const std::bad_alloc &) { return fail((ctx_expr), B200_ENOMEM, "out of host memory"); }          \
  catch (const std::exception &e_) { return fail((ctx_expr), B200_ECUDA, "internal error: %s", e_.what()); } \
  catch (...) { return fail((ctx_expr), B200_ECUDA, "internal error (unknown exception)"); }

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(ctx, B200_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));    \
  } while (0)

#define LAUNCH(call)                                                                                  \
  do {                                                                                                \
    int rc_ = (call);                                                                                 \
    if (rc_ < 0) return fail(ctx, B200_ECUDA, "%s: %s", #call, cudaGetErrorString(cudaGetLastError())); \
    ctx->launches += (uint64_t)rc_;                                                                   \
  } while (0)

bool read_blob(const std::string &path, std::vector<float> *out, size_t floats) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  out->resize(floats);
  size_t got = fread(out->data(), sizeof(float), floats, f);
  fclose(f);
  return got == floats;
}

std::string default_weights_dir() {
  const char *env = getenv("B200_DMZ_WEIGHTS");
  if (env && *env) return env;
  Dl_info info;
  if (dladdr((void *)&default_weights_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t slash = p.find_last_of('/');
    std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
    return dir + "/weights";
  }
  return "weights";
}

void free_lane(Lane *l) {
  cudaFree(l->d_frames), cudaFree(l->d_cb), cudaFree(l->d_cr), cudaFree(l->d_lines), cudaFree(l->d_geom);
  cudaFree(l->d_q8), l->d_q8 = nullptr;
  cudaFree(l->d_cards), cudaFree(l->d_vprob), cudaFree(l->d_scan), cudaFree(l->d_records), cudaFree(l->d_grad), cudaFree(l->d_check), cudaFree(l->d_flags);
  l->d_frames = l->d_cb = l->d_cr = nullptr;
  l->d_lines = nullptr, l->d_geom = nullptr, l->d_cards = nullptr, l->d_vprob = nullptr, l->d_scan = nullptr;
  l->d_records = nullptr, l->d_grad = nullptr, l->d_check = nullptr, l->d_flags = nullptr;
  l->grad_elems = 0;
  l->frame_bytes = 0;
  l->chroma_bytes = 0;
  l->cap = 0;
}

int ensure_config(b200_ctx *ctx, int w, int h, int orientation, int planes) {
  if (w < 32 || h < 32 || orientation < 1 || orientation > 4) return fail(ctx, B200_EINVAL, "bad frame size / orientation");
  if (ctx->cfg_w == w && ctx->cfg_h == h && ctx->cfg_orient == orientation && ctx->cfg_planes == planes) return B200_OK;
  b200_build_detect_params(w, h, orientation, kDetectThreads, &ctx->dp[0]);
  b200_build_detect_params(w / 2, h / 2, orientation, kDetectThreads, &ctx->dp[1]);
  b200_build_geom_params(w, h, orientation, planes, &ctx->gp);
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
  for (int p = 0; p < 2; p++) {
    ctx->dp[p].use_global_grad = 0;
    // the shared-memory variant stores padded pixel indices in 16 bits (detect.cu): larger strips (1080p portrait,
    // 88 x 899) take the global-gradient variant with 32-bit work lists even where their gradients would fit
    bool wide_index = false;
    for (int s = 0; s < 4; s++)
      wide_index = wide_index || (size_t)((ctx->dp[p].strip[s].w + 2 + 3) & ~3) * (ctx->dp[p].strip[s].h + 2) > 65535;
    if (wide_index || detect_smem_bytes(ctx->dp[p]) > (size_t)max_smem - 1024) {
      ctx->dp[p].use_global_grad = 1;
      if (detect_smem_bytes(ctx->dp[p]) > (size_t)max_smem - 1024)
        return fail(ctx, B200_EUNSUPPORTED, "detection strips of a %dx%d frame do not fit in shared memory", w, h);
    }
  }
  for (int s = 0; s < 4; s++)
    if (ctx->dp[0].strip[s].w < 1 || ctx->dp[0].strip[s].h < 1) return fail(ctx, B200_EINVAL, "degenerate detection strip");
  ctx->cfg_w = w, ctx->cfg_h = h, ctx->cfg_orient = orientation, ctx->cfg_planes = planes;
  return B200_OK;
}

int ensure_capacity(b200_ctx *ctx, Lane *l, int n, int w, int h, bool need_frames, bool need_chroma = false) {
  // per-frame records / cards: sized by frame COUNT
  if (n > l->cap) {
    uint8_t *keep_frames = l->d_frames, *keep_cb = l->d_cb, *keep_cr = l->d_cr;
    const size_t keep_bytes = l->frame_bytes, keep_chroma = l->chroma_bytes;
    l->d_frames = l->d_cb = l->d_cr = nullptr;  // staging buffers are managed below
    free_lane(l);
    l->d_frames = keep_frames, l->d_cb = keep_cb, l->d_cr = keep_cr;
    l->frame_bytes = keep_bytes, l->chroma_bytes = keep_chroma;
    CU(cudaMalloc(&l->d_lines, sizeof(b200_line) * 3 * (size_t)n * 4));
    CU(cudaMalloc(&l->d_geom, sizeof(FrameGeom) * (size_t)n));
    CU(cudaMemset(l->d_geom, 0, sizeof(FrameGeom) * (size_t)n));  // struct padding is copied to the host with the records
    CU(cudaMalloc(&l->d_cards, kCardBytes * (size_t)n));
    CU(cudaMalloc(&l->d_vprob, (size_t)n * (540 * sizeof(float) + 16)));
    CU(cudaMalloc(&l->d_q8, (size_t)n * 16 * B200_Q8_STRIDE));
    CU(cudaMemset(l->d_q8, 0, (size_t)n * 16 * B200_Q8_STRIDE));  // the CNN kernel prefetches all 16 slots of a frame, written or not
    CU(cudaMalloc(&l->d_scan, sizeof(b200_scan) * (size_t)n));
    CU(cudaMalloc(&l->d_records, sizeof(b200_frame_record) * (size_t)n));
    CU(cudaMalloc(&l->d_check, sizeof(unsigned int) * (size_t)n));
    CU(cudaMalloc(&l->d_flags, (size_t)n));
    l->cap = n;
  }
  // input staging (host-buffer calls only): sized in BYTES for this call's n x w x h
  if (need_frames) {
    const size_t need = (size_t)n * w * h;
    if (need > l->frame_bytes) {
      cudaFree(l->d_frames);
      l->d_frames = nullptr;
      l->frame_bytes = 0;
      CU(cudaMalloc(&l->d_frames, need));
      l->frame_bytes = need;
    }
    const size_t need_c = (size_t)n * (w / 2) * (h / 2) + 16;
    if (need_chroma && need_c > l->chroma_bytes) {
      cudaFree(l->d_cb), cudaFree(l->d_cr);
      l->d_cb = l->d_cr = nullptr;
      l->chroma_bytes = 0;
      CU(cudaMalloc(&l->d_cb, need_c));
      CU(cudaMalloc(&l->d_cr, need_c));
      l->chroma_bytes = need_c;
    }
  }
  if ((size_t)w * h > (size_t)l->cap_w * l->cap_h) l->cap_w = w, l->cap_h = h;
  return B200_OK;
}

int ensure_grad(b200_ctx *ctx, Lane *l, int n) {
  size_t need = 0;
  for (int p = 0; p < 2; p++) {
    if (!ctx->dp[p].use_global_grad) continue;
    size_t mx = 0;
    for (int s = 0; s < 4; s++) {
      size_t v = (size_t)((ctx->dp[p].strip[s].w + 2 + 3) & ~3) * (ctx->dp[p].strip[s].h + 2);  // detect_pad_pitch
      mx = v > mx ? v : mx;
    }
    size_t e = (size_t)n * 4 * mx * 2;
    need = e > need ? e : need;
  }
  if (need > l->grad_elems) {
    cudaFree(l->d_grad);
    l->d_grad = nullptr;
    l->grad_elems = 0;
    CU(cudaMalloc(&l->d_grad, need * sizeof(int16_t)));
    l->grad_elems = need;
  }
  return B200_OK;
}

int ensure_misc(b200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->misc_bytes) return B200_OK;
  cudaFree(ctx->d_misc);
  ctx->d_misc = nullptr, ctx->misc_bytes = 0;
  CU(cudaMalloc(&ctx->d_misc, bytes));
  ctx->misc_bytes = bytes;
  return B200_OK;
}

// Copy n strided host planes into a dense device buffer (or hand back the caller's device pointer).
int stage_planes(b200_ctx *ctx, cudaStream_t stream, const uint8_t *src, int row_stride, size_t frame_stride, int w, int h,
                 int n, int mem, uint8_t *d_dense, const uint8_t **out_ptr, int *out_row_stride, size_t *out_frame_stride) {
  if (mem == B200_MEM_DEVICE) {
    *out_ptr = src, *out_row_stride = row_stride, *out_frame_stride = frame_stride;
    return B200_OK;
  }
  if (row_stride == w && frame_stride == (size_t)w * h) {
    CU(cudaMemcpyAsync(d_dense, src, (size_t)n * w * h, cudaMemcpyHostToDevice, stream));
  } else if (frame_stride == (size_t)row_stride * h) {
    CU(cudaMemcpy2DAsync(d_dense, w, src, row_stride, w, (size_t)h * n, cudaMemcpyHostToDevice, stream));
  } else {
    for (int i = 0; i < n; i++)
      CU(cudaMemcpy2DAsync(d_dense + (size_t)i * w * h, w, src + (size_t)i * frame_stride, row_stride, w, h,
                           cudaMemcpyHostToDevice, stream));
  }
  *out_ptr = d_dense, *out_row_stride = w, *out_frame_stride = (size_t)w * h;
  return B200_OK;
}

// a plane description is usable when rows do not overlap and frames do not overlap (n == 1 callers may pass any frame stride)
bool strides_ok(int row_stride, size_t frame_stride, int w, int h) {
  return w > 0 && h > 0 && row_stride >= w && frame_stride >= (size_t)row_stride * (size_t)(h - 1) + (size_t)w;
}

struct Crop {  // uploaded sub-rectangle [x0, x1) x [y0, y1) of the frame; x0 == x1 means "whole frame"
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  bool active() const { return x1 > x0; }
};

int detect_sequence(b200_ctx *ctx, Lane *l, const uint8_t *y, int yrs, size_t yfs, const uint8_t *cb, const uint8_t *cr,
                    int crs, size_t cfs, int n, bool timed, Crop crop = Crop()) {
  const size_t plane_stride = (size_t)l->cap * 4;
  int rc = ensure_grad(ctx, l, n);
  if (rc) return rc;
  if (timed) CU(cudaEventRecord(ctx->ev[ST_DETECT], l->stream));
  LAUNCH(launch_detect(ctx->dp[0], y, yrs, yfs, n, nullptr, nullptr, l->d_lines, l->d_grad, l->stream, crop.x0, crop.y0));
  if (cb && cr) {
    // Cb is searched only where Y produced no line, Cr only where neither Y nor Cb did (dmz.cpp:351)
    LAUNCH(launch_detect(ctx->dp[1], cb, crs, cfs, n, l->d_lines, nullptr, l->d_lines + plane_stride, l->d_grad, l->stream));
    LAUNCH(launch_detect(ctx->dp[1], cr, crs, cfs, n, l->d_lines, l->d_lines + plane_stride, l->d_lines + 2 * plane_stride,
                         l->d_grad, l->stream));
  }
  if (timed) CU(cudaEventRecord(ctx->ev[ST_GEOMETRY], l->stream));
  LAUNCH(launch_geometry(ctx->gp, l->d_lines, plane_stride, n, l->d_geom, l->stream));
  return B200_OK;
}

WarpSource warp_source(const uint8_t *base, int row_stride, size_t frame_stride, int width, int height, int n, const Crop &crop) {
  WarpSource S;
  S.base = base, S.row_stride = row_stride, S.frame_stride = frame_stride, S.frame_w = width, S.frame_h = height, S.n = n;
  S.ox = crop.active() ? crop.x0 : 0, S.oy = crop.active() ? crop.y0 : 0;
  S.bw = crop.active() ? crop.x1 - crop.x0 : width, S.bh = crop.active() ? crop.y1 - crop.y0 : height;
  return S;
}

// detect -> warp -> scan -> records for n frames already visible to the device on lane l.  lazy: the caller does not
// want the cards, so only the rows scan_card_image reads are warped (inside launch_scan) and card_check stays 0.
int pipeline_on_lane(b200_ctx *ctx, Lane *l, const uint8_t *dy, int drs, size_t dfs, int width, int height, int n,
                     uint8_t *dcards, b200_frame_record *drec, bool timed, bool lazy, Crop crop = Crop()) {
  int rc = detect_sequence(ctx, l, dy, drs, dfs, nullptr, nullptr, 0, 0, n, timed, crop);
  if (rc) return rc;
  LazyWarp lw;
  lw.src = warp_source(dy, drs, dfs, width, height, n, crop);
  lw.portrait = ctx->cfg_orient == B200_ORIENT_PORTRAIT || ctx->cfg_orient == B200_ORIENT_PORTRAIT_UPSIDE_DOWN;
  if (timed) CU(cudaEventRecord(ctx->ev[ST_WARP], l->stream));
  if (!lazy) LAUNCH(launch_warp(lw.src, l->d_geom, dcards, l->d_check, WARP_FULL, nullptr, nullptr, lw.portrait, l->stream));
  cudaEvent_t *ev = timed ? ctx->ev : nullptr;
  LAUNCH(launch_scan(ctx->wts, dcards, n, l->d_geom, nullptr, l->d_vprob, l->d_q8, l->d_scan, l->stream,
                     ev ? ev[ST_VSEG] : nullptr, ev ? ev[ST_HSEG] : nullptr, ev ? ev[ST_CATEGORIZE] : nullptr,
                     ev ? ev[ST_FINALIZE] : nullptr, lazy ? &lw : nullptr, dcards, timed && lazy ? ctx->lazy_ev : nullptr));
  ctx->lazy_timed = timed && lazy;
  LAUNCH(launch_finalize_records(l->d_geom, l->d_scan, lazy ? nullptr : l->d_check, n, drec, l->stream, crop.active() ? l->d_flags : nullptr,
                                 crop.x0, crop.y0, crop.x1, crop.y1));
  if (timed) CU(cudaEventRecord(ctx->ev[ST_COUNT], l->stream));
  return B200_OK;
}

int collect_stage_times(b200_ctx *ctx, int n) {
  for (int s = 0; s < ST_COUNT; s++) {
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, ctx->ev[s], ctx->ev[s + 1]));
    ctx->stage_ms[s] += ms;
  }
  if (ctx->lazy_timed) {  // the lazy warp launches run inside the vseg bracket: book them under "warp"
    for (int k = 0; k < 3; k++) {
      float ms = 0.0f;
      CU(cudaEventElapsedTime(&ms, ctx->lazy_ev[2 * k], ctx->lazy_ev[2 * k + 1]));
      ctx->stage_ms[ST_WARP] += ms;
      ctx->stage_ms[ST_VSEG] -= ms;
    }
  }
  ctx->stage_frames += (uint64_t)n;
  return B200_OK;
}

}  // namespace

extern "C" {

int b200_ctx_create(b200_ctx **out, int device_ordinal, const char *weights_dir) try {
  if (!out) return B200_EINVAL;
  *out = nullptr;
  b200_ctx *ctx = new b200_ctx();
  ctx->device = device_ordinal;
  *out = ctx;  // returned even on failure so the caller can read b200_last_error, then destroy
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(ctx, B200_ECUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  CU(cudaSetDevice(device_ordinal));
  for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&ctx->lane[i].stream, cudaStreamNonBlocking));
  ctx->stream = ctx->lane[0].stream;
  for (int i = 0; i <= ST_COUNT; i++) CU(cudaEventCreate(&ctx->ev[i]));
  for (int i = 0; i < 6; i++) CU(cudaEventCreate(&ctx->lazy_ev[i]));
  const char *mode_env = getenv("B200_DMZ_CARD_MODE");
  if (mode_env && *mode_env) ctx->card_mode = atoi(mode_env) != 0;
  const char *chunk_env = getenv("B200_DMZ_HOST_CHUNK");
  if (chunk_env && atoi(chunk_env) > 0) ctx->host_chunk = atoi(chunk_env);
  const char *crop_env = getenv("B200_DMZ_CROP_MARGIN");
  if (crop_env && *crop_env) ctx->crop_margin = atoi(crop_env);

  const std::string dir = weights_dir && *weights_dir ? weights_dir : default_weights_dir();
  std::vector<float> blob;
  if (!read_blob(dir + "/modelm_befe75da.bin", &blob, 10403)) return fail(ctx, B200_EINVAL, "cannot read %s/modelm_befe75da.bin", dir.c_str());
  CU(cudaMalloc(&ctx->d_vseg, blob.size() * sizeof(float)));
  CU(cudaMemcpy(ctx->d_vseg, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  static const char *names[3] = {"modelc_5c241121.bin", "modelc_01266c1b.bin", "modelc_b00bf70c.bin"};
  std::vector<float> cnn[3], hwT(3 * 320 * 32);
  const float *ptrs[3];
  for (int m = 0; m < 3; m++) {
    if (!read_blob(dir + "/" + names[m], &cnn[m], 10682)) return fail(ctx, B200_EINVAL, "cannot read %s/%s", dir.c_str(), names[m]);
    CU(cudaMalloc(&ctx->d_cnn[m], cnn[m].size() * sizeof(float)));
    CU(cudaMemcpy(ctx->d_cnn[m], cnn[m].data(), cnn[m].size() * sizeof(float), cudaMemcpyHostToDevice));
    for (int u = 0; u < 32; u++)
      for (int j = 0; j < 320; j++) hwT[((size_t)m * 320 + j) * 32 + u] = cnn[m][80 + (size_t)u * 320 + j];
    ptrs[m] = cnn[m].data();
  }
  CU(cudaMalloc(&ctx->d_hwT, hwT.size() * sizeof(float)));
  CU(cudaMemcpy(ctx->d_hwT, hwT.data(), hwT.size() * sizeof(float), cudaMemcpyHostToDevice));
  fill_conv_constants(ptrs, &ctx->wts.conv);
  {
    std::vector<int8_t> convb((size_t)3 * 3 * 2 * 80 * 16);
    std::vector<float> convf(48);
    std::vector<uint16_t> hidb((size_t)3 * 2 * 40 * 32 * 8);
    b200_build_cnn_mma_tables(ptrs, convb.data(), convf.data(), hidb.data());
    CU(cudaMalloc(&ctx->d_cnn_convb, convb.size()));
    CU(cudaMemcpy(ctx->d_cnn_convb, convb.data(), convb.size(), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&ctx->d_cnn_convf, convf.size() * sizeof(float)));
    CU(cudaMemcpy(ctx->d_cnn_convf, convf.data(), convf.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&ctx->d_cnn_hidb, hidb.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(ctx->d_cnn_hidb, hidb.data(), hidb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    ctx->wts.cnn_convb = ctx->d_cnn_convb, ctx->wts.cnn_convf = ctx->d_cnn_convf, ctx->wts.cnn_hidb = ctx->d_cnn_hidb;
  }
  {  // E0 (expiry digit): optional weights + bilateral tables
    std::vector<float> eb;
    if (read_blob(dir + "/modelc_bf4dd6c8.bin", &eb, 74406)) {
      // appended for expiry_kernel's layer 2: the 40 x 50 x 25 kernels regrouped by INPUT map, [k][f][28] (25 taps padded to
      // seven 16-byte units), so that the kernels of a few input maps are one contiguous block to stage into shared memory
      eb.resize(B200_EXPIRY_C2K_OFFSET + 50 * 40 * 28, 0.0f);
      for (int f = 0; f < 40; f++)
        for (int k = 0; k < 50; k++)
          for (int t = 0; t < 25; t++) eb[B200_EXPIRY_C2K_OFFSET + ((size_t)k * 40 + f) * 28 + t] = eb[1300 + (size_t)f * 1250 + k * 25 + t];
      {  // + the layer-2 kernels as fp16 hi / lo operand slices for the tensor-core kernel (expiry_mma.cu)
        std::vector<uint16_t> halfs((size_t)13 * 2 * 14 * 48 * 8);
        b200_build_expiry_c2_halfs(eb.data(), halfs.data());
        eb.resize(B200_EXPIRY_C2H_OFFSET + B200_EXPIRY_C2H_FLOATS, 0.0f);
        memcpy(eb.data() + B200_EXPIRY_C2H_OFFSET, halfs.data(), halfs.size() * sizeof(uint16_t));
        eb.resize(B200_EXPIRY_HWT_OFFSET + 120 * 176, 0.0f);  // + the hidden weights [176][120] transposed to [120][176] (coalesced reads)
        for (int i = 0; i < 176; i++)
          for (int j = 0; j < 120; j++) eb[B200_EXPIRY_HWT_OFFSET + (size_t)j * 176 + i] = eb[51340 + (size_t)i * 120 + j];
      }
      CU(cudaMalloc(&ctx->d_expiry, eb.size() * sizeof(float)));
      CU(cudaMemcpy(ctx->d_expiry, eb.data(), eb.size() * sizeof(float), cudaMemcpyHostToDevice));
      float color[256], space[5];
      b200_build_bilateral_tables(color, space);
      if (upload_bilateral_tables(color, space) != 0 || upload_bilateral_tables_mma(color, space) != 0)
        return fail(ctx, B200_ECUDA, "cudaMemcpyToSymbol (bilateral tables)");
    }
    if (read_blob(dir + "/modelm_730c4cbd.bin", &eb, 14322)) {
      CU(cudaMalloc(&ctx->d_slash, eb.size() * sizeof(float)));
      CU(cudaMemcpy(ctx->d_slash, eb.data(), eb.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  ctx->wts.vseg = ctx->d_vseg;
  for (int m = 0; m < 3; m++) ctx->wts.cnn[m] = ctx->d_cnn[m];
  ctx->wts.cnn_hwT = ctx->d_hwT;
  {
    std::vector<float> tab(256 * 256 * 2);
    b200_build_minmax_norm_table(tab.data());
    CU(cudaMalloc(&ctx->d_vnorm, tab.size() * sizeof(float)));
    CU(cudaMemcpy(ctx->d_vnorm, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  ctx->wts.vseg_norm = ctx->d_vnorm;
  {
    std::vector<int8_t> wq((size_t)4 * 14 * 64 * 16);
    std::vector<VsegUnit> units(64);
    std::vector<float> sd((size_t)256 * 256 * 2);
    b200_build_vseg_mma_tables(blob.data(), wq.data(), units.data(), sd.data());
    CU(cudaMalloc(&ctx->d_vseg_wq, wq.size()));
    CU(cudaMemcpy(ctx->d_vseg_wq, wq.data(), wq.size(), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&ctx->d_vseg_unit, sizeof(VsegUnit) * units.size()));
    CU(cudaMemcpy(ctx->d_vseg_unit, units.data(), sizeof(VsegUnit) * units.size(), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&ctx->d_vseg_sd, sd.size() * sizeof(float)));
    CU(cudaMemcpy(ctx->d_vseg_sd, sd.data(), sd.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->wts.vseg_wq = ctx->d_vseg_wq, ctx->wts.vseg_unit = ctx->d_vseg_unit, ctx->wts.vseg_sd = ctx->d_vseg_sd;
  }
  return B200_OK;
} B200_GUARD(out ? *out : nullptr)

void b200_ctx_destroy(b200_ctx *ctx) {
  if (!ctx) return;
  for (int i = 0; i < 2; i++) {
    if (ctx->lane[i].stream) cudaStreamSynchronize(ctx->lane[i].stream);
    free_lane(&ctx->lane[i]);
    if (ctx->lane[i].stream) cudaStreamDestroy(ctx->lane[i].stream);
  }
  for (int i = 0; i <= ST_COUNT; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 6; i++)
    if (ctx->lazy_ev[i]) cudaEventDestroy(ctx->lazy_ev[i]);
  cudaFree(ctx->d_misc);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  cudaFree(ctx->d_vnorm), cudaFree(ctx->d_vseg_wq), cudaFree(ctx->d_vseg_unit), cudaFree(ctx->d_vseg_sd);
  cudaFree(ctx->d_cnn_convb), cudaFree(ctx->d_cnn_convf), cudaFree(ctx->d_cnn_hidb);
  cudaFree(ctx->d_vseg), cudaFree(ctx->d_hwT), cudaFree(ctx->d_expiry), cudaFree(ctx->d_slash);
  for (int m = 0; m < 3; m++) cudaFree(ctx->d_cnn[m]);
  delete ctx;
}

const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }
uint64_t b200_launch_count(const b200_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *b200_ctx_stream(const b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t b200_full_frame_redos(const b200_ctx *ctx) { return ctx ? ctx->full_frame_redos : 0; }
void b200_transfer_bytes(const b200_ctx *ctx, uint64_t *h2d, uint64_t *d2h) {
  if (h2d) *h2d = ctx ? ctx->h2d_bytes : 0;
  if (d2h) *d2h = ctx ? ctx->d2h_bytes : 0;
}
void b200_set_crop_margin(b200_ctx *ctx, int margin) {
  if (ctx) ctx->crop_margin = margin;
}
void b200_set_card_mode(b200_ctx *ctx, int always_materialise) {
  if (ctx) ctx->card_mode = always_materialise != 0;
}

int b200_ctx_reserve(b200_ctx *ctx, int max_frames, int width, int height) try {
  if (!ctx || max_frames < 1) return B200_EINVAL;
  CU(cudaSetDevice(ctx->device));
  return ensure_capacity(ctx, &ctx->lane[0], max_frames, width, height, false);
} B200_GUARD(ctx)

void b200_set_profiling(b200_ctx *ctx, int on) {
  if (!ctx) return;
  ctx->profiling = on;
  for (int s = 0; s < ST_COUNT; s++) ctx->stage_ms[s] = 0.0;
  ctx->stage_frames = 0;
}

int b200_stage_times(const b200_ctx *ctx, double ms[7], uint64_t *frames) {
  if (!ctx) return B200_EINVAL;
  for (int s = 0; s < ST_COUNT; s++) ms[s] = ctx->stage_ms[s];
  if (frames) *frames = ctx->stage_frames;
  return B200_OK;
}

int b200_detect_edges_batch(b200_ctx *ctx, const uint8_t *y, int yrs, size_t yfs, const uint8_t *cb, const uint8_t *cr,
                            int crs, size_t cfs, int width, int height, int n, int orientation, int mem, b200_edges *edges,
                            b200_corner_points *corners, uint8_t *all_found, b200_line *lines) try {
  if (!ctx || !y || n < 1 || (!cb) != (!cr)) return fail(ctx, B200_EINVAL, "b200_detect_edges_batch: bad arguments");
  if (!strides_ok(yrs, yfs, width, height) || (cb && !strides_ok(crs, cfs, width / 2, height / 2)))
    return fail(ctx, B200_EINVAL, "b200_detect_edges_batch: row_stride < width or frame_stride < row_stride * height");
  CU(cudaSetDevice(ctx->device));
  Lane *l = &ctx->lane[0];
  int rc = ensure_config(ctx, width, height, orientation, cb ? 3 : 1);
  if (rc) return rc;
  rc = ensure_capacity(ctx, l, n, width, height, mem == B200_MEM_HOST, cb != nullptr);
  if (rc) return rc;
  const uint8_t *dy, *dcb = nullptr, *dcr = nullptr;
  int drs, dcrs = 0;
  size_t dfs, dcfs = 0;
  rc = stage_planes(ctx, l->stream, y, yrs, yfs, width, height, n, mem, l->d_frames, &dy, &drs, &dfs);
  if (rc) return rc;
  if (cb) {
    rc = stage_planes(ctx, l->stream, cb, crs, cfs, width / 2, height / 2, n, mem, l->d_cb, &dcb, &dcrs, &dcfs);
    if (rc) return rc;
    rc = stage_planes(ctx, l->stream, cr, crs, cfs, width / 2, height / 2, n, mem, l->d_cr, &dcr, &dcrs, &dcfs);
    if (rc) return rc;
  }
  rc = detect_sequence(ctx, l, dy, drs, dfs, dcb, dcr, dcrs, dcfs, n, false);
  if (rc) return rc;
  // unpack FrameGeom into the caller's structs (host side; geometry records are small)
  std::vector<FrameGeom> hg(n);
  CU(cudaMemcpyAsync(hg.data(), l->d_geom, sizeof(FrameGeom) * n, cudaMemcpyDeviceToHost, l->stream));
  std::vector<b200_line> hl;
  if (lines) {
    hl.resize((size_t)n * 4);
    CU(cudaMemcpyAsync(hl.data(), l->d_lines, sizeof(b200_line) * n * 4, cudaMemcpyDeviceToHost, l->stream));
  }
  CU(cudaStreamSynchronize(l->stream));
  std::vector<b200_edges> he(n);
  std::vector<b200_corner_points> hc(n);
  std::vector<uint8_t> hf(n);
  for (int i = 0; i < n; i++) {
    b200_found_edge *fe[4] = {&he[i].top, &he[i].left, &he[i].bottom, &he[i].right};
    for (int s = 0; s < 4; s++) fe[s]->found = hg[i].found[s], fe[s]->rho = hg[i].rho[s], fe[s]->theta = hg[i].theta[s];
    memcpy(&hc[i], hg[i].corners, sizeof(float) * 8);
    hf[i] = (uint8_t)hg[i].all_found;
  }
  // host callers get plain memcpy (a cudaMemcpy HostToHost costs a driver round trip per call: the drop-in latency path)
  auto give = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
    if (mem == B200_MEM_DEVICE) return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    memcpy(dst, src, bytes);
    return cudaSuccess;
  };
  if (edges) CU(give(edges, he.data(), sizeof(b200_edges) * n));
  if (corners) CU(give(corners, hc.data(), sizeof(b200_corner_points) * n));
  if (all_found) CU(give(all_found, hf.data(), n));
  if (lines) CU(give(lines, hl.data(), sizeof(b200_line) * n * 4));
  return B200_OK;
} B200_GUARD(ctx)

// D1-D4 alone: the four strips of every Y plane through Sobel-7 / adaptive Canny / gated Hough (best_line_for_sample,
// dmz.cpp:224-271).  BASELINE configs[3] times this entry; the parity tests read the same taps through
// b200_detect_edges_batch's `lines`.
int b200_detect_lines_batch(b200_ctx *ctx, const uint8_t *y, int yrs, size_t yfs, int width, int height, int n, int orientation,
                            int mem, b200_line *lines) try {
  if (!ctx || !y || !lines || n < 1) return fail(ctx, B200_EINVAL, "b200_detect_lines_batch: bad arguments");
  if (!strides_ok(yrs, yfs, width, height)) return fail(ctx, B200_EINVAL, "b200_detect_lines_batch: bad strides");
  CU(cudaSetDevice(ctx->device));
  Lane *l = &ctx->lane[0];
  int rc = ensure_config(ctx, width, height, orientation, 1);
  if (rc) return rc;
  const uint8_t *dy = y;
  int drs = yrs;
  size_t dfs = yfs;
  b200_line *dl = lines;
  if (mem == B200_MEM_HOST) {
    rc = ensure_capacity(ctx, l, n, width, height, true);
    if (rc) return rc;
    rc = stage_planes(ctx, l->stream, y, yrs, yfs, width, height, n, mem, l->d_frames, &dy, &drs, &dfs);
    if (rc) return rc;
    dl = l->d_lines;
  }
  rc = ensure_grad(ctx, l, n);
  if (rc) return rc;
  LAUNCH(launch_detect(ctx->dp[0], dy, drs, dfs, n, nullptr, nullptr, dl, l->d_grad, l->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(lines, dl, sizeof(b200_line) * 4 * (size_t)n, cudaMemcpyDeviceToHost, l->stream));
  CU(cudaStreamSynchronize(l->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_transform_card_batch(b200_ctx *ctx, const uint8_t *sample, int row_stride, size_t frame_stride, int width,
                              int height, int n, const b200_corner_points *corners, const uint8_t *valid, int orientation,
                              int upsample, int mem, uint8_t *cards) try {
  if (!ctx || !sample || !corners || !cards || n < 1) return fail(ctx, B200_EINVAL, "b200_transform_card_batch: bad arguments");
  if (!strides_ok(row_stride, frame_stride, width, height))
    return fail(ctx, B200_EINVAL, "b200_transform_card_batch: row_stride < width or frame_stride < row_stride * height");
  CU(cudaSetDevice(ctx->device));
  Lane *l = &ctx->lane[0];
  int rc = ensure_capacity(ctx, l, n, width, height, mem == B200_MEM_HOST);
  if (rc) return rc;
  const uint8_t *ds;
  int drs;
  size_t dfs;
  rc = stage_planes(ctx, l->stream, sample, row_stride, frame_stride, width, height, n, mem, l->d_frames, &ds, &drs, &dfs);
  if (rc) return rc;
  const b200_corner_points *dc = corners;
  const uint8_t *dv = valid;
  if (mem == B200_MEM_HOST) {
    rc = ensure_misc(ctx, (sizeof(b200_corner_points) + 1) * (size_t)n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_misc, corners, sizeof(b200_corner_points) * n, cudaMemcpyHostToDevice, l->stream));
    dc = (const b200_corner_points *)ctx->d_misc;
    if (valid) {
      uint8_t *p = (uint8_t *)ctx->d_misc + sizeof(b200_corner_points) * (size_t)n;
      CU(cudaMemcpyAsync(p, valid, n, cudaMemcpyHostToDevice, l->stream));
      dv = p;
    }
  }
  LAUNCH(launch_corners_to_geom(dc, dv, n, orientation, upsample, l->d_geom, l->stream));
  uint8_t *dcards = mem == B200_MEM_DEVICE ? cards : l->d_cards;
  {
    // the tile heuristics assume a card that fills the guide rectangle of a frame; a half-size chroma plane shows it at
    // half that scale, which frame_h (the plane's own height) already expresses
    const WarpSource S = warp_source(ds, drs, dfs, width, height, n, Crop());
    const int portrait = orientation == B200_ORIENT_PORTRAIT || orientation == B200_ORIENT_PORTRAIT_UPSIDE_DOWN;
    LAUNCH(launch_warp(S, l->d_geom, dcards, nullptr, WARP_FULL, nullptr, nullptr, portrait, l->stream));
  }
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(cards, l->d_cards, kCardBytes * n, cudaMemcpyDeviceToHost, l->stream));
  CU(cudaStreamSynchronize(l->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_scan_cards_batch(b200_ctx *ctx, const uint8_t *cards, int n, const uint8_t *valid, int mem, b200_scan *scans) try {
  if (!ctx || !cards || !scans || n < 1) return fail(ctx, B200_EINVAL, "b200_scan_cards_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Lane *l = &ctx->lane[0];
  int rc = ensure_capacity(ctx, l, n, l->cap_w ? l->cap_w : 32, l->cap_h ? l->cap_h : 32, false);
  if (rc) return rc;
  const uint8_t *dc = cards, *dv = valid;
  if (mem == B200_MEM_HOST) {
    CU(cudaMemcpyAsync(l->d_cards, cards, kCardBytes * n, cudaMemcpyHostToDevice, l->stream));
    dc = l->d_cards;
    if (valid) {
      rc = ensure_misc(ctx, n);
      if (rc) return rc;
      CU(cudaMemcpyAsync(ctx->d_misc, valid, n, cudaMemcpyHostToDevice, l->stream));
      dv = (const uint8_t *)ctx->d_misc;
    }
  }
  b200_scan *ds = mem == B200_MEM_DEVICE ? scans : l->d_scan;
  LAUNCH(launch_scan(ctx->wts, dc, n, nullptr, dv, l->d_vprob, l->d_q8, ds, l->stream, nullptr, nullptr, nullptr, nullptr));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(scans, l->d_scan, sizeof(b200_scan) * n, cudaMemcpyDeviceToHost, l->stream));
  CU(cudaStreamSynchronize(l->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_process_frames_batch(b200_ctx *ctx, const uint8_t *y, int yrs, size_t yfs, int width, int height, int n,
                              int orientation, int mem, b200_frame_record *records, uint8_t *cards_out) try {
  if (!ctx || !y || !records || n < 1) return fail(ctx, B200_EINVAL, "b200_process_frames_batch: bad arguments");
  if (!strides_ok(yrs, yfs, width, height))
    return fail(ctx, B200_EINVAL, "b200_process_frames_batch: row_stride < width or frame_stride < row_stride * height");
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_config(ctx, width, height, orientation, 1);
  if (rc) return rc;
  const bool lazy = cards_out == nullptr && ctx->card_mode == 0;
  if (mem == B200_MEM_DEVICE) {
    Lane *l = &ctx->lane[0];
    rc = ensure_capacity(ctx, l, n, width, height, false);
    if (rc) return rc;
    uint8_t *dcards = cards_out ? cards_out : l->d_cards;
    rc = pipeline_on_lane(ctx, l, y, yrs, yfs, width, height, n, dcards, records, ctx->profiling != 0, lazy);
    if (rc) return rc;
    CU(cudaStreamSynchronize(l->stream));
    if (ctx->profiling) return collect_stage_times(ctx, n);
    return B200_OK;
  }
  // Host buffers: chunked two-lane pipeline.  Lane k%2 takes chunk k: H2D copy, kernels and the D2H of the
  // records are queued on that lane's stream, so chunk k+1's copy overlaps chunk k's kernels.
  //
  // Only the part of each frame the path can touch is uploaded: the bounding rectangle of the four detection
  // strips plus a margin (PCIe, not the kernels, bounds this path).  The warp's source taps lie inside the hull
  // of the detected corners; finalize_records_kernel flags the (rare) frames whose hull leaves the uploaded
  // rectangle and those are redone below from a full-frame upload, so results never depend on the crop.
  Crop crop;
  if (ctx->crop_margin >= 0 && yfs % (size_t)yrs == 0) {
    int x0 = width, y0 = height, x1 = 0, y1 = 0;
    for (int s = 0; s < 4; s++) {
      const StripDesc &d = ctx->dp[0].strip[s];
      x0 = d.x < x0 ? d.x : x0, y0 = d.y < y0 ? d.y : y0;
      x1 = d.x + d.w > x1 ? d.x + d.w : x1, y1 = d.y + d.h > y1 ? d.y + d.h : y1;
    }
    const int m = ctx->crop_margin;
    // Measured on B200 / PCIe Gen5 (16384-frame steps): whole frames 307 KB @ 53.7 GB/s = 175 k frames/s; whole-row
    // band 200 KB @ 52.9 GB/s = 264 k frames/s; row+column rectangle 155 KB @ 46.9 GB/s (pitched 3-D copy) =
    // 302 k frames/s.  The rectangle wins; B200_DMZ_CROP_ROWS=1 selects the whole-row band instead.
    if (yrs == width && getenv("B200_DMZ_CROP_ROWS")) {
      crop.x0 = 0, crop.x1 = width;
    } else {
      crop.x0 = (x0 - m > 0 ? x0 - m : 0) & ~15;
      crop.x1 = (x1 + m + 15) & ~15;
      if (crop.x1 > width) crop.x1 = width;
    }
    crop.y0 = y0 - m > 0 ? y0 - m : 0;
    crop.y1 = y1 + m < height ? y1 + m : height;
    if ((crop.x1 - crop.x0) % 4 != 0 || (size_t)(crop.x1 - crop.x0) * (crop.y1 - crop.y0) * 10 > (size_t)width * height * 9) crop = Crop();
  }
  const int cw = crop.active() ? crop.x1 - crop.x0 : width, chh = crop.active() ? crop.y1 - crop.y0 : height;
  const int chunk = n < ctx->host_chunk ? n : ctx->host_chunk;
  for (int i = 0; i < 2; i++) {
    if (i == 1 && n <= chunk) break;
    rc = ensure_capacity(ctx, &ctx->lane[i], chunk, width, height, true);
    if (rc) return rc;
  }
  // needs-full-frame flags come back into PINNED memory: an async D2H into pageable memory would block the host
  // at every chunk and serialise the two lanes
  uint8_t *flags = nullptr;
  if (crop.active()) {
    if (ctx->pinned_bytes < (size_t)n) {
      if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
      ctx->h_pinned = nullptr, ctx->pinned_bytes = 0;
      CU(cudaHostAlloc(&ctx->h_pinned, (size_t)n, cudaHostAllocDefault));
      ctx->pinned_bytes = (size_t)n;
    }
    flags = (uint8_t *)ctx->h_pinned;
  }
  int k = 0;
  for (int f0 = 0; f0 < n; f0 += chunk, k++) {
    Lane *l = &ctx->lane[k & 1];
    const int cnt = n - f0 < chunk ? n - f0 : chunk;
    const uint8_t *dy;
    int drs;
    size_t dfs;
    if (crop.active()) {
      if (cw == width && yrs == width) {
        // whole rows of dense frames: one contiguous host segment per frame -> plain 2-D copy (width = the segment)
        CU(cudaMemcpy2DAsync(l->d_frames, (size_t)cw * chh, y + (size_t)f0 * yfs + (size_t)crop.y0 * yrs, yfs, (size_t)cw * chh,
                             (size_t)cnt, cudaMemcpyHostToDevice, l->stream));
      } else {
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        p.srcPtr = make_cudaPitchedPtr((void *)(y + (size_t)f0 * yfs), (size_t)yrs, (size_t)width, yfs / (size_t)yrs);
        p.srcPos = make_cudaPos((size_t)crop.x0, (size_t)crop.y0, 0);
        p.dstPtr = make_cudaPitchedPtr(l->d_frames, (size_t)cw, (size_t)cw, (size_t)chh);
        p.extent = make_cudaExtent((size_t)cw, (size_t)chh, (size_t)cnt);
        p.kind = cudaMemcpyHostToDevice;
        CU(cudaMemcpy3DAsync(&p, l->stream));
      }
      ctx->h2d_bytes += (uint64_t)cw * chh * cnt;
      dy = l->d_frames, drs = cw, dfs = (size_t)cw * chh;
    } else {
      rc = stage_planes(ctx, l->stream, y + (size_t)f0 * yfs, yrs, yfs, width, height, cnt, B200_MEM_HOST, l->d_frames, &dy, &drs, &dfs);
      if (rc) return rc;
      ctx->h2d_bytes += (uint64_t)width * height * cnt;
    }
    rc = pipeline_on_lane(ctx, l, dy, drs, dfs, width, height, cnt, l->d_cards, l->d_records, false, lazy, crop);
    if (rc) return rc;
    CU(cudaMemcpyAsync(records + f0, l->d_records, sizeof(b200_frame_record) * cnt, cudaMemcpyDeviceToHost, l->stream));
    ctx->d2h_bytes += sizeof(b200_frame_record) * (uint64_t)cnt + (crop.active() ? cnt : 0) + (cards_out ? kCardBytes * cnt : 0);
    if (crop.active()) CU(cudaMemcpyAsync(flags + f0, l->d_flags, cnt, cudaMemcpyDeviceToHost, l->stream));
    if (cards_out) CU(cudaMemcpyAsync(cards_out + (size_t)f0 * kCardBytes, l->d_cards, kCardBytes * cnt, cudaMemcpyDeviceToHost, l->stream));
  }
  CU(cudaStreamSynchronize(ctx->lane[0].stream));
  if (k > 1) CU(cudaStreamSynchronize(ctx->lane[1].stream));
  // frames whose card quad reaches outside the uploaded rectangle: redo from the whole frame
  for (int i = 0; i < (crop.active() ? n : 0); i++) {
    if (!flags[i]) continue;
    Lane *l = &ctx->lane[0];
    const uint8_t *dy;
    int drs;
    size_t dfs;
    rc = stage_planes(ctx, l->stream, y + (size_t)i * yfs, yrs, yfs, width, height, 1, B200_MEM_HOST, l->d_frames, &dy, &drs, &dfs);
    if (rc) return rc;
    rc = pipeline_on_lane(ctx, l, dy, drs, dfs, width, height, 1, l->d_cards, l->d_records, false, lazy);
    if (rc) return rc;
    CU(cudaMemcpyAsync(records + i, l->d_records, sizeof(b200_frame_record), cudaMemcpyDeviceToHost, l->stream));
    if (cards_out) CU(cudaMemcpyAsync(cards_out + (size_t)i * kCardBytes, l->d_cards, kCardBytes, cudaMemcpyDeviceToHost, l->stream));
    CU(cudaStreamSynchronize(l->stream));
    ctx->full_frame_redos++;
    ctx->h2d_bytes += (uint64_t)width * height;
    ctx->d2h_bytes += sizeof(b200_frame_record);
  }
  return B200_OK;
} B200_GUARD(ctx)

int b200_calc_persp_transform_batch(b200_ctx *ctx, const float *src_pts, const float *dst_pts, int n, float *M) try {
  if (!ctx || !src_pts || !dst_pts || !M || n < 1) return fail(ctx, B200_EINVAL, "b200_calc_persp_transform_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_misc(ctx, sizeof(float) * 25 * (size_t)n);
  if (rc) return rc;
  float *ds = (float *)ctx->d_misc, *dd = ds + (size_t)n * 8, *dm = dd + (size_t)n * 8;
  CU(cudaMemcpyAsync(ds, src_pts, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dd, dst_pts, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(launch_homography_only(ds, dd, n, dm, ctx->stream));
  CU(cudaMemcpyAsync(M, dm, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_categorize_patches_batch(b200_ctx *ctx, const uint8_t *patches, int n, int mem, float *out) try {
  if (!ctx || !patches || !out || n < 1) return fail(ctx, B200_EINVAL, "b200_categorize_patches_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const uint8_t *dp = patches;
  float *dout = out;
  // scratch layout: [prepared patches n * 528][scores n * 160 (host mode)][raw patches n * 513 (host mode)]
  const size_t o_scores = (size_t)n * B200_Q8_STRIDE, o_raw = o_scores + (size_t)n * 160;
  int rc = ensure_misc(ctx, o_raw + (size_t)n * 513 + 64);
  if (rc) return rc;
  uint8_t *q8 = (uint8_t *)ctx->d_misc;
  if (mem == B200_MEM_HOST) {
    dout = (float *)((uint8_t *)ctx->d_misc + o_scores);
    uint8_t *p = (uint8_t *)ctx->d_misc + o_raw;
    CU(cudaMemcpyAsync(p, patches, (size_t)n * 513, cudaMemcpyHostToDevice, ctx->stream));
    dp = p;
  }
  LAUNCH(launch_categorize_patches(ctx->wts, dp, nullptr, n, dout, q8, ctx->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, (size_t)n * 160, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_digit_models_batch(b200_ctx *ctx, const float *patches, int n, int mem, float *out) try {
  if (!ctx || !patches || !out || n < 1) return fail(ctx, B200_EINVAL, "b200_digit_models_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const float *dp = patches;
  float *dout = out;
  if (mem == B200_MEM_HOST) {
    int rc = ensure_misc(ctx, (size_t)n * (513 + 40) * sizeof(float));
    if (rc) return rc;
    float *p = (float *)ctx->d_misc;
    CU(cudaMemcpyAsync(p, patches, (size_t)n * 513 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dp = p;
    dout = p + (size_t)n * 513;
  }
  LAUNCH(launch_categorize_patches(ctx->wts, nullptr, dp, n, dout, nullptr, ctx->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, (size_t)n * 40 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// dmz_deinterleave_uint8_c2 over a batch: n interleaved 2-channel planes (width pixels = 2 * width bytes per row) into
// two dense width x height planes each.
int b200_deinterleave_c2_batch(b200_ctx *ctx, const uint8_t *interleaved, int row_stride, size_t frame_stride, int width, int height,
                               int n, int mem, uint8_t *channel1, uint8_t *channel2) try {
  if (!ctx || !interleaved || !channel1 || !channel2 || n < 1 || width < 1 || height < 1 || row_stride < 2 * width)
    return fail(ctx, B200_EINVAL, "b200_deinterleave_c2_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (mem == B200_MEM_DEVICE) {
    LAUNCH(launch_deinterleave_c2(interleaved, row_stride, frame_stride, width, height, n, channel1, channel2, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  const size_t plane = (size_t)width * height, in_bytes = ((2 * plane * n) + 15) & ~(size_t)15, out_bytes = (plane * n + 15) & ~(size_t)15;
  int rc = ensure_misc(ctx, in_bytes + 2 * out_bytes);
  if (rc) return rc;
  uint8_t *d_in = (uint8_t *)ctx->d_misc, *d_c1 = d_in + in_bytes, *d_c2 = d_c1 + out_bytes;
  for (int i = 0; i < n; i++)  // rows are packed on the way up (2 * width bytes each)
    CU(cudaMemcpy2DAsync(d_in + (size_t)i * 2 * plane, (size_t)2 * width, interleaved + (size_t)i * frame_stride, (size_t)row_stride,
                         (size_t)2 * width, (size_t)height, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(launch_deinterleave_c2(d_in, 2 * width, 2 * plane, width, height, n, d_c1, d_c2, ctx->stream));
  CU(cudaMemcpyAsync(channel1, d_c1, plane * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(channel2, d_c2, plane * n, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h2d_bytes += 2 * plane * n, ctx->d2h_bytes += 2 * plane * n;
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// dmz_YCbCr_to_RGB over a batch (formats.cu).  Host planes are packed on the way up; the RGB image comes back dense.
int b200_ycbcr_to_rgb_batch(b200_ctx *ctx, const uint8_t *y, int yrs, size_t yfs, const uint8_t *cb, const uint8_t *cr, int crs,
                            size_t cfs, int width, int height, int n, int channels, int mem, uint8_t *rgb) try {
  if (!ctx || !y || !cb || !cr || !rgb || n < 1 || (channels != 3 && channels != 4))
    return fail(ctx, B200_EINVAL, "b200_ycbcr_to_rgb_batch: bad arguments");
  if (!strides_ok(yrs, yfs, width, height) || !strides_ok(crs, cfs, width, height))
    return fail(ctx, B200_EINVAL, "b200_ycbcr_to_rgb_batch: row_stride < width or frame_stride < row_stride * height");
  CU(cudaSetDevice(ctx->device));
  if (mem == B200_MEM_DEVICE) {
    LAUNCH(launch_ycbcr_to_rgb(y, yrs, yfs, cb, cr, crs, cfs, width, height, n, channels, rgb, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  const size_t plane = (size_t)width * height, in_bytes = (plane * n + 15) & ~(size_t)15, out_bytes = plane * n * channels;
  int rc = ensure_misc(ctx, 3 * in_bytes + out_bytes);
  if (rc) return rc;
  uint8_t *d_in[3] = {(uint8_t *)ctx->d_misc, (uint8_t *)ctx->d_misc + in_bytes, (uint8_t *)ctx->d_misc + 2 * in_bytes};
  uint8_t *d_out = (uint8_t *)ctx->d_misc + 3 * in_bytes;
  const uint8_t *dp[3];
  int drs = 0;
  size_t dfs = 0;
  if ((rc = stage_planes(ctx, ctx->stream, y, yrs, yfs, width, height, n, mem, d_in[0], &dp[0], &drs, &dfs))) return rc;
  if ((rc = stage_planes(ctx, ctx->stream, cb, crs, cfs, width, height, n, mem, d_in[1], &dp[1], &drs, &dfs))) return rc;
  if ((rc = stage_planes(ctx, ctx->stream, cr, crs, cfs, width, height, n, mem, d_in[2], &dp[2], &drs, &dfs))) return rc;
  LAUNCH(launch_ycbcr_to_rgb(dp[0], drs, dfs, dp[1], dp[2], drs, dfs, width, height, n, channels, d_out, ctx->stream));
  CU(cudaMemcpyAsync(rgb, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h2d_bytes += 3 * plane * n, ctx->d2h_bytes += out_bytes;
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// dmz_deinterleave_RGBA_to_R (formats.cu)
int b200_rgba_to_r_batch(b200_ctx *ctx, const uint8_t *rgba, size_t n_pixels, int mem, uint8_t *r) try {
  if (!ctx || !rgba || !r || n_pixels < 1) return fail(ctx, B200_EINVAL, "b200_rgba_to_r_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (mem == B200_MEM_DEVICE) {
    LAUNCH(launch_rgba_to_r(rgba, n_pixels, r, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  const size_t in_bytes = (4 * n_pixels + 15) & ~(size_t)15;
  int rc = ensure_misc(ctx, in_bytes + n_pixels);
  if (rc) return rc;
  uint8_t *d_in = (uint8_t *)ctx->d_misc, *d_out = d_in + in_bytes;
  CU(cudaMemcpyAsync(d_in, rgba, 4 * n_pixels, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(launch_rgba_to_r(d_in, n_pixels, d_out, ctx->stream));
  CU(cudaMemcpyAsync(r, d_out, n_pixels, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h2d_bytes += 4 * n_pixels, ctx->d2h_bytes += n_pixels;
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy over a batch of planes (formats.cu)
int b200_stencil3_batch(b200_ctx *ctx, const uint8_t *img, int row_stride, size_t frame_stride, int width, int height, int n, int kind,
                        int mem, int16_t *out) try {
  if (!ctx || !img || !out || n < 1 || kind < B200_STENCIL_SCHARR_DX_ABS || kind > B200_STENCIL_SOBEL_DX_DY)
    return fail(ctx, B200_EINVAL, "b200_stencil3_batch: bad arguments");
  if (!strides_ok(row_stride, frame_stride, width, height))
    return fail(ctx, B200_EINVAL, "b200_stencil3_batch: row_stride < width or frame_stride < row_stride * height");
  CU(cudaSetDevice(ctx->device));
  if (mem == B200_MEM_DEVICE) {
    LAUNCH(launch_stencil3(img, row_stride, frame_stride, width, height, n, kind, out, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  const size_t plane = (size_t)width * height, in_bytes = (plane * n + 15) & ~(size_t)15, out_bytes = plane * n * sizeof(int16_t);
  int rc = ensure_misc(ctx, in_bytes + out_bytes);
  if (rc) return rc;
  uint8_t *d_in = (uint8_t *)ctx->d_misc;
  int16_t *d_out = (int16_t *)(d_in + in_bytes);
  const uint8_t *dp = nullptr;
  int drs = 0;
  size_t dfs = 0;
  if ((rc = stage_planes(ctx, ctx->stream, img, row_stride, frame_stride, width, height, n, mem, d_in, &dp, &drs, &dfs))) return rc;
  LAUNCH(launch_stencil3(dp, drs, dfs, width, height, n, kind, d_out, ctx->stream));
  CU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h2d_bytes += plane * n, ctx->d2h_bytes += out_bytes;
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// dmz_focus_score / dmz_brightness_score over a batch.  Host frames: only the scoring rectangle crosses PCIe (the
// reference's ROI clamps the Sobel taps at the rectangle, so nothing outside it is ever read).
int b200_frame_scores_batch(b200_ctx *ctx, const uint8_t *y, int yrs, size_t yfs, int width, int height, int n,
                            int use_full_image, int mem, float *focus, float *brightness) try {
  if (!ctx || !y || n < 1 || width < 1 || height < 1 || yrs < width || (!focus && !brightness))
    return fail(ctx, B200_EINVAL, "b200_frame_scores_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  int rc4[4];
  b200_scoring_rect(width, height, use_full_image, rc4);
  const int rx = rc4[0], ry = rc4[1], rw = rc4[2], rh = rc4[3];
  if (rw < 1 || rh < 1) return fail(ctx, B200_EINVAL, "b200_frame_scores_batch: empty scoring rectangle");
  if (mem == B200_MEM_DEVICE) {
    LAUNCH(launch_frame_scores(y, yrs, yfs, n, rx, ry, rw, rh, focus, brightness, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  const size_t roi_bytes = (size_t)rw * rh;
  const size_t off_scores = ((size_t)n * roi_bytes + 15) & ~(size_t)15;
  int rc = ensure_misc(ctx, off_scores + 2 * sizeof(float) * (size_t)n);
  if (rc) return rc;
  uint8_t *d_roi = (uint8_t *)ctx->d_misc;
  float *d_focus = (float *)((uint8_t *)ctx->d_misc + off_scores), *d_bright = d_focus + n;
  if (yfs % (size_t)yrs == 0) {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr((void *)y, (size_t)yrs, (size_t)width, yfs / (size_t)yrs);
    p.srcPos = make_cudaPos((size_t)rx, (size_t)ry, 0);
    p.dstPtr = make_cudaPitchedPtr(d_roi, (size_t)rw, (size_t)rw, (size_t)rh);
    p.extent = make_cudaExtent((size_t)rw, (size_t)rh, (size_t)n);
    p.kind = cudaMemcpyHostToDevice;
    CU(cudaMemcpy3DAsync(&p, ctx->stream));
  } else {
    for (int i = 0; i < n; i++)
      CU(cudaMemcpy2DAsync(d_roi + (size_t)i * roi_bytes, rw, y + (size_t)i * yfs + (size_t)ry * yrs + rx, yrs, rw, rh,
                           cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->h2d_bytes += (uint64_t)roi_bytes * n;
  LAUNCH(launch_frame_scores(d_roi, rw, roi_bytes, n, 0, 0, rw, rh, d_focus, d_bright, ctx->stream));
  if (focus) CU(cudaMemcpyAsync(focus, d_focus, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (brightness) CU(cudaMemcpyAsync(brightness, d_bright, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->d2h_bytes += (uint64_t)sizeof(float) * n * ((focus != nullptr) + (brightness != nullptr));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// best_expiry_seg over a batch.  The |Scharr| planes are scratch (231 KB per card), so the batch runs in chunks -- deep ones:
// the one-thread-per-card search is latency-bound and its throughput grows with the cards in flight (2048 per chunk: 55 k
// cards/s, 32768: 344 k cards/s; 7.6 GB of scratch, allocated only when a call is that large).
int b200_best_expiry_seg_batch(b200_ctx *ctx, const uint8_t *cards, const uint16_t *y_offsets, int n, int mem,
                               b200_expiry_group *groups, int max_groups, int32_t *n_groups, int32_t *n_dropped, int16_t *sobel_out) try {
  if (!ctx || !cards || !y_offsets || !groups || !n_groups || n < 1 || max_groups < 1)
    return fail(ctx, B200_EINVAL, "b200_best_expiry_seg_batch: bad arguments");
  if (!ctx->d_slash) return fail(ctx, B200_EUNSUPPORTED, "modelm_730c4cbd.bin was not found in the weights directory");
  CU(cudaSetDevice(ctx->device));
  const size_t card_bytes = (size_t)B200_CARD_W * B200_CARD_H, sob_bytes = card_bytes * sizeof(int16_t);
  static const int max_chunk = [] { const char *e = getenv("B200_DMZ_EXPIRY_CHUNK"); return e && atoi(e) > 0 ? atoi(e) : 32768; }();
  const int chunk = n < max_chunk ? n : max_chunk;
  const bool host = mem == B200_MEM_HOST;
  auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t o_sob = 0, o_ls = o_sob + up16(sob_bytes * chunk), o_cards = o_ls + up16(sizeof(int32_t) * B200_EXPIRY_SEG_SCRATCH_INTS * chunk),
               o_yo = o_cards + up16(host ? card_bytes * chunk : 0), o_grp = o_yo + up16(host ? sizeof(uint16_t) * chunk : 0),
               o_cnt = o_grp + up16(host ? sizeof(b200_expiry_group) * (size_t)max_groups * chunk : 0),
               total = o_cnt + up16(host ? 2 * sizeof(int32_t) * chunk : 0);
  int rc = ensure_misc(ctx, total);
  if (rc) return rc;
  uint8_t *base = (uint8_t *)ctx->d_misc;
  int16_t *d_sob = (int16_t *)(base + o_sob);
  int32_t *d_ls = (int32_t *)(base + o_ls);
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int cnt = n - f0 < chunk ? n - f0 : chunk;
    const uint8_t *dc = cards + (size_t)f0 * card_bytes;
    const uint16_t *dy = y_offsets + f0;
    b200_expiry_group *dg = groups + (size_t)f0 * max_groups;
    int32_t *dn = n_groups + f0, *dd = n_dropped ? n_dropped + f0 : nullptr;
    if (host) {
      CU(cudaMemcpyAsync(base + o_cards, dc, card_bytes * cnt, cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaMemcpyAsync(base + o_yo, dy, sizeof(uint16_t) * cnt, cudaMemcpyHostToDevice, ctx->stream));
      dc = base + o_cards, dy = (const uint16_t *)(base + o_yo), dg = (b200_expiry_group *)(base + o_grp);
      dn = (int32_t *)(base + o_cnt), dd = dn + chunk;
    }
    if (host) CU(cudaMemsetAsync(dg, 0, sizeof(b200_expiry_group) * (size_t)max_groups * cnt, ctx->stream));  // unused slots are copied back too
    LAUNCH(launch_expiry_seg(dc, dy, cnt, ctx->d_slash, d_sob, d_ls, dg, max_groups, dn, dd, ctx->stream));
    if (host) {
      CU(cudaMemcpyAsync(groups + (size_t)f0 * max_groups, dg, sizeof(b200_expiry_group) * (size_t)max_groups * cnt, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaMemcpyAsync(n_groups + f0, dn, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
      if (n_dropped) CU(cudaMemcpyAsync(n_dropped + f0, dd, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (sobel_out)
      CU(cudaMemcpyAsync(sobel_out + (size_t)f0 * card_bytes, d_sob, sob_bytes * cnt, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // the scratch planes are reused by the next chunk
  }
  return B200_OK;
} B200_GUARD(ctx)

static int expiry_call(b200_ctx *ctx, const uint8_t *patches, const float *prepared, int n, int mem, float *out) try {
  if (!ctx || (!patches && !prepared) || !out || n < 1) return fail(ctx, B200_EINVAL, "b200_expiry_*: bad arguments");
  if (!ctx->d_expiry) return fail(ctx, B200_EUNSUPPORTED, "modelc_bf4dd6c8.bin was not found in the weights directory");
  CU(cudaSetDevice(ctx->device));
  const uint8_t *dp = patches;
  const float *df = prepared;
  float *dout = out;
  if (mem == B200_MEM_HOST) {
    int rc = ensure_misc(ctx, (size_t)n * (176 * sizeof(float) + 10 * sizeof(float)) + 64);
    if (rc) return rc;
    dout = (float *)ctx->d_misc;
    uint8_t *p = (uint8_t *)ctx->d_misc + (size_t)n * 10 * sizeof(float);
    if (patches) {
      CU(cudaMemcpyAsync(p, patches, (size_t)n * 176, cudaMemcpyHostToDevice, ctx->stream));
      dp = p;
    } else {
      CU(cudaMemcpyAsync(p, prepared, (size_t)n * 176 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      df = (const float *)p;
    }
  }
  LAUNCH(launch_expiry_digits(ctx->d_expiry, dp, df, n, dout, ctx->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, (size_t)n * 10 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

// categorize_expiry_digits' inner step for many characters at once: crop m 16x11 windows out of n_cards warped cards.
int b200_expiry_digits_at_batch(b200_ctx *ctx, const uint8_t *cards, int n_cards, const int32_t *where, int m, int mem, float *out) try {
  if (!ctx || !cards || !where || !out || n_cards < 1 || m < 1) return fail(ctx, B200_EINVAL, "b200_expiry_digits_at_batch: bad arguments");
  if (!ctx->d_expiry) return fail(ctx, B200_EUNSUPPORTED, "modelc_bf4dd6c8.bin was not found in the weights directory");
  CU(cudaSetDevice(ctx->device));
  const uint8_t *dc = cards;
  const int32_t *dw = where;
  float *dout = out;
  if (mem == B200_MEM_HOST) {
    for (int i = 0; i < m; i++)
      if (where[3 * i] < 0 || where[3 * i] >= n_cards) return fail(ctx, B200_EINVAL, "b200_expiry_digits_at_batch: card index out of range");
    const size_t card_bytes = (size_t)B200_CARD_W * B200_CARD_H;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_out = 0, o_where = up16(sizeof(float) * 10 * (size_t)m), o_cards = o_where + up16(sizeof(int32_t) * 3 * (size_t)m);
    int rc = ensure_misc(ctx, o_cards + card_bytes * n_cards);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_misc;
    CU(cudaMemcpyAsync(base + o_cards, cards, card_bytes * n_cards, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(base + o_where, where, sizeof(int32_t) * 3 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    dc = base + o_cards, dw = (const int32_t *)(base + o_where), dout = (float *)(base + o_out);
  }
  LAUNCH(launch_expiry_digits(ctx->d_expiry, dc, nullptr, m, dout, ctx->stream, dw));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, sizeof(float) * 10 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_expiry_digits_batch(b200_ctx *ctx, const uint8_t *patches, int n, int mem, float *out) try {
  return expiry_call(ctx, patches, nullptr, n, mem, out);
} B200_GUARD(ctx)

int b200_expiry_digit_models_batch(b200_ctx *ctx, const float *prepared, int n, int mem, float *out) try {
  return expiry_call(ctx, nullptr, prepared, n, mem, out);
} B200_GUARD(ctx)

int b200_vseg_rows_batch(b200_ctx *ctx, const uint8_t *cards, int n, int mem, float *out) try {
  if (!ctx || !cards || !out || n < 1) return fail(ctx, B200_EINVAL, "b200_vseg_rows_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const uint8_t *dc = cards;
  float *dout = out;
  if (mem == B200_MEM_HOST) {
    const size_t o_cards = ((size_t)n * 540 * sizeof(float) + 15) & ~(size_t)15;
    int rc = ensure_misc(ctx, o_cards + kCardBytes * (size_t)n);
    if (rc) return rc;
    dout = (float *)ctx->d_misc;
    uint8_t *p = (uint8_t *)ctx->d_misc + o_cards;
    CU(cudaMemcpyAsync(p, cards, kCardBytes * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    dc = p;
  }
  LAUNCH(launch_vseg_coarse_rows(ctx->wts, dc, n, dout, ctx->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, (size_t)n * 540 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

int b200_vseg_model_batch(b200_ctx *ctx, const float *rows, int n, int mem, float *out) try {
  if (!ctx || !rows || !out || n < 1) return fail(ctx, B200_EINVAL, "b200_vseg_model_batch: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const float *dr = rows;
  float *dout = out;
  if (mem == B200_MEM_HOST) {
    int rc = ensure_misc(ctx, (size_t)n * (204 + 3) * sizeof(float));
    if (rc) return rc;
    float *p = (float *)ctx->d_misc;
    CU(cudaMemcpyAsync(p, rows, (size_t)n * 204 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dr = p;
    dout = p + (size_t)n * 204;
  }
  LAUNCH(launch_vseg_model(ctx->wts, dr, n, dout, ctx->stream));
  if (mem == B200_MEM_HOST) CU(cudaMemcpyAsync(out, dout, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
} B200_GUARD(ctx)

}  // extern "C"
