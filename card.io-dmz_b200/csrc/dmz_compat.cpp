// card.io-dmz_b200/csrc/dmz_compat.cpp -- the reference's C++ entry points (dmz.h:45-101, scan/scan.h:50-72) on top
// of the C ABI, batch of one.  Host-side only: image bookkeeping, the session EMA and scanner_result's checks
// (scan.cpp:88-194); all image arithmetic happens in the CUDA kernels behind b200_*_batch.
// If CUDA fails the functions report "nothing found / not usable" (the reference has no error channel either).
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <mutex>

#include "b200_dmz.h"
#include "dmz_b200_compat.h"
#include "expiry_session.h"
#include <time.h>

namespace {

std::mutex g_mutex;
b200_ctx *g_default_ctx = nullptr;  // scanner_* carry no dmz_context: they share one lazily created context
// A b200_ctx is not thread-safe.  The reference's functions can be called from any thread (each call is self-contained), so
// calls that go through the SHARED default context are serialised here; a context owned by a dmz_context is the caller's
// (one thread per dmz_context, as in the SDKs).
std::recursive_mutex g_call_mutex;
struct CtxCall {
  std::unique_lock<std::recursive_mutex> lock;
  explicit CtxCall(b200_ctx *ctx) {
    if (ctx != nullptr && ctx == g_default_ctx) lock = std::unique_lock<std::recursive_mutex>(g_call_mutex);
  }
};

b200_ctx *default_ctx() {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g_default_ctx) {
    const char *dev = getenv("B200_DMZ_DEVICE");
    if (b200_ctx_create(&g_default_ctx, dev ? atoi(dev) : 0, nullptr) != B200_OK) {
      fprintf(stderr, "b200dmz: %s\n", b200_last_error(g_default_ctx));
      b200_ctx_destroy(g_default_ctx);
      g_default_ctx = nullptr;
    }
  }
  return g_default_ctx;
}

struct PlaneView {
  const uint8_t *data;
  int w, h, step;
};

// cvGetSize / llcv_get_data_origin semantics: an image with a ROI is the ROI (cv/image_util.cpp:30-37)
PlaneView view_of(const IplImage *img) {
  PlaneView v;
  v.step = img->widthStep;
  if (img->roi) {
    v.data = (const uint8_t *)img->imageData + (size_t)img->roi->yOffset * img->widthStep + img->roi->xOffset;
    v.w = img->roi->width, v.h = img->roi->height;
  } else {
    v.data = (const uint8_t *)img->imageData;
    v.w = img->width, v.h = img->height;
  }
  return v;
}

IplImage *create_image(int width, int height, int depth, int channels) {
  // Prefer the host application's OpenCV allocator so that its cvReleaseImage frees what we hand out.
  typedef struct { int width, height; } Size2;
  typedef IplImage *(*create_fn)(Size2, int, int);
  static create_fn cv_create = (create_fn)dlsym(RTLD_DEFAULT, "cvCreateImage");
  if (cv_create) {
    Size2 s = {width, height};
    return cv_create(s, depth, channels);
  }
  IplImage *img = (IplImage *)calloc(1, sizeof(IplImage));
  img->nSize = sizeof(IplImage);
  img->nChannels = channels;
  img->depth = depth;
  img->width = width, img->height = height;
  img->align = 4;
  img->widthStep = (width * channels * ((depth & 255) / 8) + 3) & ~3;  // cvCreateImage aligns rows to 4 bytes
  img->imageSize = img->widthStep * height;
  img->imageData = img->imageDataOrigin = (char *)malloc((size_t)img->imageSize);
  memcpy(img->colorModel, channels == 1 ? "GRAY" : "RGB\0", 4);
  memcpy(img->channelSeq, channels == 1 ? "GRAY" : (channels == 3 ? "BGR\0" : "BGRA"), 4);
  return img;
}

IplImage *create_gray_image(int width, int height) { return create_image(width, height, IPL_DEPTH_8U, 1); }

// a view of a multi-byte-per-pixel image (ROI honoured): px_bytes per pixel
PlaneView view_of_px(const IplImage *img, int px_bytes) {
  PlaneView v = view_of(img);
  if (img->roi) v.data += (size_t)img->roi->xOffset * (px_bytes - 1);
  return v;
}

// rows of a u8 plane packed into a dense w x h buffer
void pack_rows(const PlaneView &v, uint8_t *dense) {
  for (int r = 0; r < v.h; r++) memcpy(dense + (size_t)r * v.w, v.data + (size_t)r * v.step, (size_t)v.w);
}

IplImage *create_card_image() { return create_gray_image(B200_CARD_W, B200_CARD_H); }


bool luhn(const uint8_t *d, int n) {
  int even = 0, sum = 0;
  for (int i = n - 1; i >= 0; i--) {
    int addend = d[i] * (1 << (even++ & 1));
    sum += addend % 10 + addend / 10;
  }
  return sum % 10 == 0;
}

}  // namespace

// implemented in scanner.cpp's translation unit via the C ABI: reuse its prefix table through a scratch session
extern "C" int b200_card_type_for_number(const uint8_t *digits, int n);

dmz_context *dmz_context_create(void) {
  dmz_context *dmz = (dmz_context *)calloc(1, sizeof(dmz_context));
  b200_ctx *ctx = nullptr;
  const char *dev = getenv("B200_DMZ_DEVICE");
  if (b200_ctx_create(&ctx, dev ? atoi(dev) : 0, nullptr) != B200_OK) {
    fprintf(stderr, "b200dmz: %s\n", b200_last_error(ctx));
    b200_ctx_destroy(ctx);
    ctx = nullptr;
  }
  dmz->mz = ctx;  // dmz.h:17-20: mz is the designated per-platform hook
  return dmz;
}

void dmz_context_destroy(dmz_context *dmz) {
  if (!dmz) return;
  b200_ctx_destroy((b200_ctx *)dmz->mz);
  free(dmz);
}

void dmz_prepare_for_backgrounding(dmz_context *) {}

bool dmz_found_all_edges(dmz_edges e) { return e.top.found && e.bottom.found && e.left.found && e.right.found; }

bool dmz_detect_edges(IplImage *y, IplImage *cb, IplImage *cr, FrameOrientation orientation, dmz_edges *found_edges,
                      dmz_corner_points *corner_points) {
  found_edges->top.found = found_edges->bottom.found = found_edges->left.found = found_edges->right.found = 0;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx || !y || !cb || !cr) return false;
  PlaneView vy = view_of(y), vb = view_of(cb), vr = view_of(cr);
  if (vb.w != vy.w / 2 || vb.h != vy.h / 2 || vr.w != vb.w || vr.h != vb.h || vr.step != vb.step) return false;
  b200_edges e;
  b200_corner_points c;
  uint8_t all = 0;
  int rc = b200_detect_edges_batch(ctx, vy.data, vy.step, (size_t)vy.step * vy.h, vb.data, vr.data, vb.step,
                                   (size_t)vb.step * vb.h, vy.w, vy.h, 1, orientation, B200_MEM_HOST, &e, &c, &all, nullptr);
  if (rc != B200_OK) return false;
  memcpy(found_edges, &e, sizeof(e));  // identical layouts (include/b200_dmz.h)
  if (all) memcpy(corner_points, &c, sizeof(c));
  return all != 0;
}

// dmz_deinterleave_uint8_c2 (dmz.h:64, dmz.cpp:49-56): always allocates both channel images; the caller frees them.
void dmz_deinterleave_uint8_c2(IplImage *interleaved, IplImage **channel1, IplImage **channel2) {
  PlaneView v = view_of(interleaved);  // width in pixels; two bytes per pixel
  if (interleaved->roi) v.data += interleaved->roi->xOffset;  // view_of offsets by xOffset bytes; pixels are 2 bytes wide
  *channel1 = create_gray_image(v.w, v.h);
  *channel2 = create_gray_image(v.w, v.h);
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx) return;
  const size_t plane = (size_t)v.w * v.h;
  uint8_t *tmp = (uint8_t *)malloc(2 * plane);
  if (b200_deinterleave_c2_batch(ctx, v.data, v.step, (size_t)v.step * v.h, v.w, v.h, 1, B200_MEM_HOST, tmp, tmp + plane) == B200_OK) {
    for (int r = 0; r < v.h; r++) {
      memcpy((*channel1)->imageData + (size_t)r * (*channel1)->widthStep, tmp + (size_t)r * v.w, (size_t)v.w);
      memcpy((*channel2)->imageData + (size_t)r * (*channel2)->widthStep, tmp + plane + (size_t)r * v.w, (size_t)v.w);
    }
  }
  free(tmp);
}

// dmz_has_opencv (dmz.h:60, dmz.cpp:42-47): "can a card-sized image be allocated" -- here: is the GPU path usable
int dmz_has_opencv(void) { return default_ctx() != nullptr; }

// dmz_YCbCr_to_RGB (dmz.h:72, dmz.cpp:58-64): allocates a 3-channel image when *rgb is NULL; a caller-provided image may
// have 3 or 4 channels (llcv_YCbCr2RGB_u8_c writes alpha = 255 into the fourth, cv/convert.cpp:453, 496-498).
void dmz_YCbCr_to_RGB(IplImage *y, IplImage *cb, IplImage *cr, IplImage **rgb) {
  if (!y || !cb || !cr || !rgb) return;
  const PlaneView vy = view_of(y), vb = view_of(cb), vr = view_of(cr);
  if (*rgb == NULL) *rgb = create_image(vy.w, vy.h, y->depth, 3);
  IplImage *out = *rgb;
  const int ch = out->nChannels;
  if ((ch != 3 && ch != 4) || vb.w != vy.w || vb.h != vy.h || vr.w != vy.w || vr.h != vy.h) return;
  const PlaneView vo = view_of_px(out, ch);
  if (vo.w != vy.w || vo.h != vy.h) return;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx) return;
  const size_t plane = (size_t)vy.w * vy.h;
  uint8_t *tmp = (uint8_t *)malloc(plane * (size_t)(3 + ch));
  pack_rows(vy, tmp), pack_rows(vb, tmp + plane), pack_rows(vr, tmp + 2 * plane);
  uint8_t *dense = tmp + 3 * plane;
  if (b200_ycbcr_to_rgb_batch(ctx, tmp, vy.w, plane, tmp + plane, tmp + 2 * plane, vy.w, plane, vy.w, vy.h, 1, ch, B200_MEM_HOST, dense) == B200_OK)
    for (int r = 0; r < vy.h; r++) memcpy((uint8_t *)vo.data + (size_t)r * vo.step, dense + (size_t)r * vy.w * ch, (size_t)vy.w * ch);
  free(tmp);
}

// dmz_deinterleave_RGBA_to_R (dmz.h:67, dmz.cpp:66-109)
void dmz_deinterleave_RGBA_to_R(uint8_t *source, uint8_t *dest, int size) {
  if (!source || !dest || size < 1) return;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (ctx) b200_rgba_to_r_batch(ctx, source, (size_t)size, B200_MEM_HOST, dest);
}

namespace {
// the three Cython-only stencil wrappers (dmz.h:105-107, dmz.cpp:519-531): u8 source (ROI honoured), int16 destination
void stencil_into(IplImage *src, IplImage *dst, int kind) {
  if (!src || !dst || dst->depth != (int)IPL_DEPTH_16S) return;
  const PlaneView vs = view_of(src), vd = view_of_px(dst, 2);
  if (vd.w != vs.w || vd.h != vs.h) return;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx) return;
  const size_t plane = (size_t)vs.w * vs.h;
  uint8_t *tmp = (uint8_t *)malloc(plane * 3 + 2);  // the plane, one byte of padding when its size is odd, the int16 result
  pack_rows(vs, tmp);
  int16_t *dense = (int16_t *)(tmp + plane + (plane & 1));
  if (b200_stencil3_batch(ctx, tmp, vs.w, plane, vs.w, vs.h, 1, kind, B200_MEM_HOST, dense) == B200_OK)
    for (int r = 0; r < vs.h; r++) memcpy((uint8_t *)vd.data + (size_t)r * vd.step, dense + (size_t)r * vs.w, (size_t)vs.w * sizeof(int16_t));
  free(tmp);
}
}  // namespace
void dmz_scharr3_dx_abs(IplImage *src, IplImage *dst) { stencil_into(src, dst, B200_STENCIL_SCHARR_DX_ABS); }
void dmz_scharr3_dy_abs(IplImage *src, IplImage *dst) { stencil_into(src, dst, B200_STENCIL_SCHARR_DY_ABS); }
void dmz_sobel3_dx_dy(IplImage *src, IplImage *dst) { stencil_into(src, dst, B200_STENCIL_SOBEL_DX_DY); }

// dmz_best_expiry_seg (dmz.h:111, dmz.cpp:605-620): malloc'ed array of groups, each with a malloc'ed rectangle list; the
// caller frees both.  scores / seen counts are not produced by segmentation (the reference copies uninitialised
// values there); they are zero here.
void dmz_best_expiry_seg(IplImage *card_y, uint16_t starting_y_offset, CythonGroupedRects **expiry_groups, uint16_t *number_of_groups) {
  *expiry_groups = NULL;
  *number_of_groups = 0;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx || !card_y) return;
  PlaneView v = view_of(card_y);
  if (v.w != B200_CARD_W || v.h != B200_CARD_H) return;
  uint8_t *dense = (uint8_t *)malloc((size_t)B200_CARD_W * B200_CARD_H);
  for (int r = 0; r < B200_CARD_H; r++) memcpy(dense + (size_t)r * B200_CARD_W, v.data + (size_t)r * v.step, B200_CARD_W);
  const int max_groups = 64;
  b200_expiry_group *g = (b200_expiry_group *)calloc(max_groups, sizeof(b200_expiry_group));
  int32_t count = 0;
  int rc = b200_best_expiry_seg_batch(ctx, dense, &starting_y_offset, 1, B200_MEM_HOST, g, max_groups, &count, nullptr, nullptr);
  free(dense);
  if (rc == B200_OK && count > 0) {
    CythonGroupedRects *out = (CythonGroupedRects *)calloc((size_t)count, sizeof(CythonGroupedRects));
    for (int i = 0; i < count; i++) {
      out[i].top = g[i].top, out[i].left = g[i].left, out[i].width = g[i].width, out[i].height = g[i].height;
      out[i].character_width = g[i].character_width;
      out[i].pattern = (uint8_t)g[i].pattern;
      out[i].number_of_character_rects = g[i].n_rects;
      out[i].character_rects = (CythonCharacterRect *)malloc(sizeof(CythonCharacterRect) * (size_t)g[i].n_rects);
      for (int k = 0; k < g[i].n_rects; k++) out[i].character_rects[k].top = g[i].rect_top[k], out[i].character_rects[k].left = g[i].rect_left[k];
    }
    *expiry_groups = out;
    *number_of_groups = (uint16_t)count;
  }
  free(g);
}

// dmz_focus_score / dmz_brightness_score (dmz.h:77-80, dmz.cpp:183-195).  The scoring rectangle is derived from
// cvGetSize(image) and applied in whole-image coordinates, as the reference's cvSetImageROI does.  (The reference also
// drops any ROI the caller had set; this layer leaves the caller's IplImage untouched.)
static float frame_score(IplImage *image, bool use_full_image, bool want_focus) {
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx || !image) return 0.0f;
  PlaneView v = view_of(image);
  // b200_frame_scores_batch places the rectangle inside a (w x h) plane; hand it the whole image when sizes agree,
  // otherwise a virtual plane of the ROI's size anchored at the image origin (what cvSetImageROI would address)
  float focus = 0.0f, bright = 0.0f;
  int rc = b200_frame_scores_batch(ctx, (const uint8_t *)image->imageData, image->widthStep, (size_t)image->widthStep * image->height,
                                   v.w, v.h, 1, use_full_image ? 1 : 0, B200_MEM_HOST, want_focus ? &focus : nullptr,
                                   want_focus ? nullptr : &bright);
  if (rc != B200_OK) return 0.0f;
  return want_focus ? focus : bright;
}

float dmz_focus_score(IplImage *image, bool use_full_image) { return frame_score(image, use_full_image, true); }
float dmz_brightness_score(IplImage *image, bool use_full_image) { return frame_score(image, use_full_image, false); }

void dmz_transform_card(dmz_context *dmz, IplImage *sample, dmz_corner_points corner_points, FrameOrientation orientation,
                        bool upsample, IplImage **transformed) {
  b200_ctx *ctx = dmz && dmz->mz ? (b200_ctx *)dmz->mz : default_ctx();
  CtxCall call_guard(ctx);
  if (*transformed == NULL) *transformed = create_card_image();  // dmz.cpp:493-495; the caller frees it
  if (!ctx || !sample) return;
  PlaneView v = view_of(sample);
  IplImage *out = *transformed;
  b200_corner_points c;
  memcpy(&c, &corner_points, sizeof(c));
  if (out->widthStep == B200_CARD_W) {
    b200_transform_card_batch(ctx, v.data, v.step, (size_t)v.step * v.h, v.w, v.h, 1, &c, nullptr, orientation, upsample,
                              B200_MEM_HOST, (uint8_t *)out->imageData);
  } else {
    uint8_t *tmp = (uint8_t *)malloc((size_t)B200_CARD_W * B200_CARD_H);
    if (b200_transform_card_batch(ctx, v.data, v.step, (size_t)v.step * v.h, v.w, v.h, 1, &c, nullptr, orientation, upsample,
                                  B200_MEM_HOST, tmp) == B200_OK)
      for (int r = 0; r < B200_CARD_H; r++) memcpy(out->imageData + (size_t)r * out->widthStep, tmp + (size_t)r * B200_CARD_W, B200_CARD_W);
    free(tmp);
  }
}

void scanner_initialize(ScannerState *state) { scanner_reset(state); }

void scanner_reset(ScannerState *state) {  // scan.cpp:23-35
  state->count15 = 0;
  state->count16 = 0;
  memset(state->aggregated15.v, 0, sizeof(state->aggregated15.v));
  memset(state->aggregated16.v, 0, sizeof(state->aggregated16.v));
  state->session_analytics.num_frames_scanned = 0;
  state->session_analytics.frames_ring_start = 0;
  state->timeOfCardNumberCompletionInMilliseconds = 0;
  state->scan_expiry = false;
  state->expiry_month = 0;
  state->expiry_year = 0;
  state->expiry_groups.clear();
  state->name_groups.clear();
}

void scanner_add_frame(ScannerState *state, IplImage *y, FrameScanResult *result) {
  scanner_add_frame_with_expiry(state, y, false, result);
}

// The reference accepts expired cards only in its DMZ_DEBUG / CYTHON_DMZ builds (expiry_categorize.cpp:378-395); SDK
// builds do not.  Default: SDK behaviour.
static int g_allow_past_expiry = 0;
void b200_compat_set_allow_past_expiry(int allow) { g_allow_past_expiry = allow; }

namespace {

// scan_card_image's expiry step + expiry_extract (frame.cpp:72-80, expiry_categorize.cpp:448-497) for one card:
// segmentation and digit categorization on the GPU, cross-frame aggregation in the caller's ScannerState.
void expiry_step(b200_ctx *ctx, ScannerState *state, const uint8_t *card, FrameScanResult *result, bool usable) {
  result->expiry_groups.clear();
  result->name_groups.clear();
  // the reference's lists are unbounded: start with room for 32 groups and retry with what the kernel reports it had to drop
  std::vector<b200_expiry_group> g(32);
  int32_t count = 0, dropped = 0;
  uint16_t yo = result->vseg.y_offset;
  if (yo < B200_CARD_H - 2 * 15) {  // kCreditCardTargetHeight - 2 * kSmallCharacterHeight, frame.cpp:73
    for (int attempt = 0; attempt < 2; attempt++) {
      if (b200_best_expiry_seg_batch(ctx, card, &yo, 1, B200_MEM_HOST, g.data(), (int)g.size(), &count, &dropped, nullptr) != B200_OK) count = dropped = 0;
      if (dropped <= 0) break;
      g.assign((size_t)count + (size_t)dropped + 8, b200_expiry_group());
    }
    if (count > (int)g.size()) count = (int)g.size();
  }
  for (int i = 0; i < count; i++) {
    GroupedRects gr;
    gr.top = g[i].top, gr.left = g[i].left, gr.width = g[i].width, gr.height = g[i].height;
    gr.grouped_yet = false, gr.sum = 0, gr.character_width = g[i].character_width, gr.pattern = g[i].pattern;
    gr.recently_seen_count = gr.total_seen_count = 0;
    memset(gr.scores, 0, sizeof(gr.scores));
    for (int k = 0; k < g[i].n_rects; k++) {
      CharacterRect cr;
      cr.top = g[i].rect_top[k], cr.left = g[i].rect_left[k], cr.sum = 0;
      gr.character_rects.push_back(cr);
    }
    result->expiry_groups.push_back(gr);
  }
  if (!usable) return;  // scan.cpp:57-59: unusable frames never reach expiry_extract
  state->scan_expiry = true;
  state->name_groups = result->name_groups;
  if (count == 0) return;  // expiry_extract: nothing new, nothing to do
  // categorize characters 0, 1, 3, 4 of every group (categorize_expiry_digits)
  std::vector<int32_t> where_v((size_t)count * 4 * 3);
  std::vector<float> probs_v((size_t)count * 4 * 10);
  int32_t *where = where_v.data();
  float *probs = probs_v.data();
  const int chars[4] = {0, 1, 3, 4};
  for (int i = 0; i < count; i++)
    for (int c = 0; c < 4; c++) {
      int32_t *w = where + (i * 4 + c) * 3;
      w[0] = 0, w[1] = g[i].rect_top[chars[c]], w[2] = g[i].rect_left[chars[c]];
    }
  if (b200_expiry_digits_at_batch(ctx, card, 1, where, count * 4, B200_MEM_HOST, probs) != B200_OK) return;
  const int cap = count + (int)state->expiry_groups.size();
  std::vector<ExpiryAgg> fresh_v((size_t)count), agg_v((size_t)cap);
  ExpiryAgg *fresh = fresh_v.data(), *agg = agg_v.data();
  int n_fresh = 0, n_agg = 0;
  for (int i = 0; i < count; i++) {
    ExpiryAgg &f = fresh[n_fresh++];
    f.top = g[i].top, f.left = g[i].left, f.n_rects = g[i].n_rects, f.recently_seen = f.total_seen = 0, f.tag = -(i + 1);
    memset(f.scores, 0, sizeof(f.scores));
    for (int c = 0; c < 4; c++) memcpy(f.scores[chars[c]], probs + (i * 4 + c) * 10, sizeof(float) * 10);
  }
  for (size_t o = 0; o < state->expiry_groups.size(); o++) {
    const GroupedRects &G = state->expiry_groups[o];
    ExpiryAgg &a = agg[n_agg++];
    a.top = G.top, a.left = G.left, a.n_rects = (int)G.character_rects.size();
    a.recently_seen = G.recently_seen_count, a.total_seen = G.total_seen_count, a.tag = (int)o;
    memcpy(a.scores, G.scores, sizeof(a.scores));  // rows 0..4 of the 11 x 10 row-major matrix
  }
  expiry_aggregate(agg, &n_agg, fresh, n_fresh, cap);
  GroupedRectsList next;
  for (int k = 0; k < n_agg; k++) {
    GroupedRects G = agg[k].tag >= 0 ? state->expiry_groups[(size_t)agg[k].tag] : result->expiry_groups[(size_t)(-agg[k].tag - 1)];
    G.top = agg[k].top, G.left = agg[k].left;
    G.recently_seen_count = agg[k].recently_seen, G.total_seen_count = agg[k].total_seen;
    memcpy(G.scores, agg[k].scores, sizeof(agg[k].scores));
    next.push_back(G);
  }
  // what the reference leaves in result->expiry_groups: the new groups that matched nothing (now also in the aggregate)
  GroupedRectsList unmatched;
  for (int k = 0; k < n_agg; k++)
    if (agg[k].tag < 0) {
      GroupedRects G = result->expiry_groups[(size_t)(-agg[k].tag - 1)];
      memcpy(G.scores, agg[k].scores, sizeof(agg[k].scores));
      unmatched.push_back(G);
    }
  state->expiry_groups.swap(next);
  result->expiry_groups.swap(unmatched);
  time_t now = time(NULL);
  struct tm tmv;
  localtime_r(&now, &tmv);
  for (size_t o = 0; o < state->expiry_groups.size(); o++) {
    const GroupedRects &G = state->expiry_groups[o];
    if (G.total_seen_count < 3) continue;
    stable_month_year(reinterpret_cast<const float(*)[10]>(G.scores), (int)G.character_rects.size(), tmv.tm_year + 1900, tmv.tm_mon + 1,
                      g_allow_past_expiry != 0, &state->expiry_month, &state->expiry_year);
  }
}

}  // namespace

void scanner_add_frame_with_expiry(ScannerState *state, IplImage *y, bool scan_expiry, FrameScanResult *result) {
  // scan_card_image (frame.cpp:24-81) on the GPU, then scan.cpp:41-86
  result->upside_down = false;
  result->usable = false;
  b200_ctx *ctx = default_ctx();
  CtxCall call_guard(ctx);
  if (!ctx || !y || y->width != B200_CARD_W || y->height != B200_CARD_H) return;
  const bool need_number = state->timeOfCardNumberCompletionInMilliseconds == 0;
  const bool need_expiry = scan_expiry && (state->expiry_month == 0 || state->expiry_year == 0);
  b200_scan s;
  int rc;
  uint8_t *tmp = nullptr;
  const uint8_t *card = (const uint8_t *)y->imageData;
  if (y->widthStep != B200_CARD_W) {
    tmp = (uint8_t *)malloc((size_t)B200_CARD_W * B200_CARD_H);
    for (int r = 0; r < B200_CARD_H; r++) memcpy(tmp + (size_t)r * B200_CARD_W, y->imageData + (size_t)r * y->widthStep, B200_CARD_W);
    card = tmp;
  }
  rc = b200_scan_cards_batch(ctx, card, 1, nullptr, B200_MEM_HOST, &s);
  if (rc != B200_OK) {
    free(tmp);
    return;
  }
  memcpy(&result->vseg, &s.vseg, sizeof(s.vseg));
  result->upside_down = s.upside_down;
  if (s.upside_down || !(s.vseg.score > 15)) {  // frame.cpp:38-49: no expiry work on flipped / unusable-vseg frames
    free(tmp);
    return;
  }
  if (need_number) {
    result->usable = s.usable;
    memcpy(&result->hseg, &s.hseg, sizeof(s.hseg));
    memcpy(result->scores.v, s.scores, sizeof(s.scores));
  } else {
    result->usable = true;  // collect_card_number == false: only the vseg gate applies
  }
  if (need_expiry) expiry_step(ctx, state, card, result, result->usable);
  free(tmp);
  if (!result->usable || !need_number) return;
  state->mostRecentUsableHSeg = result->hseg;
  state->mostRecentUsableVSeg = result->vseg;
  NumberScores *agg = nullptr;
  if (result->hseg.n_offsets == 15) agg = &state->aggregated15, state->count15++;
  else if (result->hseg.n_offsets == 16) agg = &state->aggregated16, state->count16++;
  if (!agg) return;
  for (int i = 0; i < 160; i++) agg->v[i] *= 0.8f;                           // kDecayFactor, scan.cpp:75-83
  for (int i = 0; i < 160; i++) agg->v[i] += result->scores.v[i] * (1 - 0.8f);
}

void scanner_result(ScannerState *state, ScannerResult *result) {  // scan.cpp:88-194
  result->complete = false;
  if (state->timeOfCardNumberCompletionInMilliseconds > 0) {
    *result = state->successfulCardNumberResult;
  } else {
    const uint16_t max_count = state->count15 > state->count16 ? state->count15 : state->count16;
    const uint16_t min_count = state->count15 < state->count16 ? state->count15 : state->count16;
    if (max_count - min_count < 3) return;
    if (min_count * 2 > max_count) return;
    result->hseg = state->mostRecentUsableHSeg;
    result->vseg = state->mostRecentUsableVSeg;
    const NumberScores *agg;
    if (state->count15 > state->count16) result->n_numbers = 15, agg = &state->aggregated15;
    else result->n_numbers = 16, agg = &state->aggregated16;
    uint8_t number[16];
    for (uint8_t i = 0; i < result->n_numbers; i++) {
      const float *row = agg->v + i * 10;
      float mx = row[0];
      int arg = 0;
      for (int j = 1; j < 10; j++)
        if (row[j] > mx) mx = row[j], arg = j;
      const float sum = tree_sum(row, 0, 10);
      result->predictions(i) = arg;
      number[i] = (uint8_t)arg;
      if (mx / sum < 0.7f) return;  // kMinStability
    }
    const int type = b200_card_type_for_number(number, result->n_numbers);
    if (type != 0 /* unrecognized */ && type != 1 /* ambiguous */ && luhn(number, result->n_numbers)) {
      struct timeval t;
      gettimeofday(&t, NULL);
      state->timeOfCardNumberCompletionInMilliseconds = (long)((t.tv_sec * 1000) + (t.tv_usec / 1000));
      state->successfulCardNumberResult = *result;
    }
  }
  // once the number is in, allow a little more time for the expiry if it is being collected (scan.cpp:163-193)
  if (state->timeOfCardNumberCompletionInMilliseconds > 0) {
    if (state->scan_expiry) {
      struct timeval t;
      gettimeofday(&t, NULL);
      const long now = (long)((t.tv_sec * 1000) + (t.tv_usec / 1000));
      if ((state->expiry_month > 0 && state->expiry_year > 0) ||
          now - (long)state->timeOfCardNumberCompletionInMilliseconds > 1000 /* EXTRA_TIME_FOR_EXPIRY_IN_MICROSECONDS, compared in ms */) {
        result->expiry_month = state->expiry_month;
        result->expiry_year = state->expiry_year;
        result->complete = true;
      }
    } else {
      result->expiry_month = 0;
      result->expiry_year = 0;
      result->complete = true;
    }
  }
}

void scanner_destroy(ScannerState *) {}
