// tests/compat_main.cpp -- a caller written against the reference's own API (dmz.h / scan/scan.h) and linked with
// libb200dmz.so: the per-frame SDK sequence of SURVEY 3.4.  Built two ways from this one source:
//   (default)                    against include/dmz_b200_compat.h (tests build it on the spot)
//   -DCOMPAT_REFERENCE_HEADERS   against the reference's own UNMODIFIED dmz.h + scan/scan.h with its vendored opencv2 /
//                                Eigen headers (oracle/Makefile target `dropin` -> oracle/_ref/compat_main_refhdr, which
//                                travels to the GPU box like the other oracle/_ref files): the caller's structs, inline
//                                Eigen accessors and mangled call sites are then the reference's, only the library differs
// usage: compat_main frames.bin n width height out.bin
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#ifdef COMPAT_REFERENCE_HEADERS
#include "dmz.h"
#include "scan/scan.h"
#define MATRIX_DATA(m) ((m).data())
#else
#include "dmz_b200_compat.h"
#define MATRIX_DATA(m) ((m).v)
#endif

static void wrap_as(IplImage *img, void *data, int w, int h, int depth, int channels) {
  memset(img, 0, sizeof(*img));
  img->nSize = sizeof(IplImage);
  img->nChannels = channels;
  img->depth = depth;
  img->width = w, img->height = h, img->widthStep = w * channels * ((depth & 255) / 8);
  img->imageSize = img->widthStep * h;
  img->imageData = img->imageDataOrigin = (char *)data;
  img->align = 4;
}
static void wrap(IplImage *img, uint8_t *data, int w, int h) { wrap_as(img, data, w, h, IPL_DEPTH_8U, 1); }

static uint32_t weighted_sum(const uint8_t *p, size_t n) {
  uint32_t c = 0;
  for (size_t i = 0; i < n; i++) c += (uint32_t)(i + 1) * p[i];
  return c;
}
static void release(IplImage *img) {
  free(img->imageDataOrigin);
  free(img);
}

int main(int argc, char **argv) {
  if (argc < 6) return 2;
  const int n = atoi(argv[2]), w = atoi(argv[3]), h = atoi(argv[4]);
  std::vector<uint8_t> frames((size_t)n * w * h), chroma((size_t)(w / 2) * (h / 2), 128);
  std::vector<uint8_t> cb_plane((size_t)(w / 2) * (h / 2)), cr_plane((size_t)(w / 2) * (h / 2));  // structured chroma for the card image
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
  fclose(f);
  FILE *out = fopen(argv[5], "wb");
  dmz_context *dmz = dmz_context_create();
  ScannerState state;
  scanner_initialize(&state);
  for (int k = 0; k < n; k++) {
    IplImage y, cb, cr;
    wrap(&y, frames.data() + (size_t)k * w * h, w, h);
    wrap(&cb, chroma.data(), w / 2, h / 2);
    wrap(&cr, chroma.data(), w / 2, h / 2);
    // the SDK scores focus on the central ninth of the card region before spending time on detection
    float fb[2] = {dmz_focus_score(&y, false), dmz_brightness_score(&y, false)};
    dmz_edges edges;
    dmz_corner_points corners;
    memset(&corners, 0, sizeof(corners));
    bool found = dmz_detect_edges(&y, &cb, &cr, FrameOrientationLandscapeRight, &edges, &corners);
    int32_t rec[8] = {found, dmz_found_all_edges(edges), 0, 0, 0, 0, 0, 0};
    uint8_t digits[16] = {0};
    float scores[160] = {0};
    uint32_t check = 0;
    uint32_t fmt[6] = {0, 0, 0, 0, 0, 0};  // RGB card, RGBA card, its R plane, the three stencils of the card
    if (found) {
      IplImage *card = NULL;
      dmz_transform_card(dmz, &y, corners, FrameOrientationLandscapeRight, false, &card);
      for (int i = 0; i < 428 * 270; i++) check += (uint32_t)(i + 1) * (uint8_t)card->imageData[(i / 428) * card->widthStep + i % 428];
      FrameScanResult fr;
      fr.flipped = false;
      fr.focus_score = 0;
      memset(MATRIX_DATA(fr.scores), 0, sizeof(float) * 160);
      scanner_add_frame_with_expiry(&state, card, false, &fr);
      rec[2] = fr.usable, rec[3] = fr.upside_down, rec[4] = fr.vseg.y_offset;
      memcpy(scores, MATRIX_DATA(fr.scores), sizeof(scores));
      ScannerResult sr;
      memset(MATRIX_DATA(sr.predictions), 0, sizeof(ptrdiff_t) * 16);
      sr.n_numbers = 0;
      scanner_result(&state, &sr);
      rec[5] = sr.complete, rec[6] = sr.n_numbers;
      for (int i = 0; i < 16; i++) digits[i] = (uint8_t)sr.predictions(i);
      // what the SDK does with a finished scan: the colour card image (chroma planes warped with upsample = true), and
      // the Cython layer's taps on the card
      for (int r = 0; r < h / 2; r++)
        for (int c = 0; c < w / 2; c++) {
          const uint8_t v = frames[(size_t)k * w * h + (size_t)(2 * r) * w + 2 * c];
          cb_plane[(size_t)r * (w / 2) + c] = v, cr_plane[(size_t)r * (w / 2) + c] = (uint8_t)(255 - v);
        }
      IplImage cbi, cri, *cb_card = NULL, *cr_card = NULL, *rgb = NULL;
      wrap(&cbi, cb_plane.data(), w / 2, h / 2);
      wrap(&cri, cr_plane.data(), w / 2, h / 2);
      dmz_transform_card(dmz, &cbi, corners, FrameOrientationLandscapeRight, true, &cb_card);
      dmz_transform_card(dmz, &cri, corners, FrameOrientationLandscapeRight, true, &cr_card);
      dmz_YCbCr_to_RGB(card, cb_card, cr_card, &rgb);  // allocates a 3-channel image
      if (rgb && rgb->nChannels == 3 && rgb->width == 428 && rgb->height == 270)
        for (int r = 0; r < 270; r++) fmt[0] += (uint32_t)(r + 1) * weighted_sum((const uint8_t *)rgb->imageData + (size_t)r * rgb->widthStep, 428 * 3);
      std::vector<uint8_t> rgba((size_t)428 * 270 * 4), red((size_t)428 * 270);
      IplImage rgba_img, *rgba_ptr = &rgba_img;
      wrap_as(&rgba_img, rgba.data(), 428, 270, IPL_DEPTH_8U, 4);
      dmz_YCbCr_to_RGB(card, cb_card, cr_card, &rgba_ptr);  // a caller-provided 4-channel image gets alpha = 255
      fmt[1] = weighted_sum(rgba.data(), rgba.size());
      dmz_deinterleave_RGBA_to_R(rgba.data(), red.data(), 428 * 270);
      fmt[2] = weighted_sum(red.data(), red.size());
      std::vector<int16_t> st((size_t)428 * 270);
      IplImage st_img, dense_card;
      std::vector<uint8_t> dense((size_t)428 * 270);
      for (int r = 0; r < 270; r++) memcpy(dense.data() + (size_t)r * 428, card->imageData + (size_t)r * card->widthStep, 428);
      wrap(&dense_card, dense.data(), 428, 270);
      wrap_as(&st_img, st.data(), 428, 270, IPL_DEPTH_16S, 1);
      dmz_scharr3_dx_abs(&dense_card, &st_img);
      fmt[3] = weighted_sum((const uint8_t *)st.data(), st.size() * 2);
      dmz_scharr3_dy_abs(&dense_card, &st_img);
      fmt[4] = weighted_sum((const uint8_t *)st.data(), st.size() * 2);
      dmz_sobel3_dx_dy(&dense_card, &st_img);
      fmt[5] = weighted_sum((const uint8_t *)st.data(), st.size() * 2);
      if (rgb) release(rgb);
      release(cb_card), release(cr_card);
      release(card);
    }
    rec[7] = (int32_t)check;
    fwrite(rec, sizeof(rec), 1, out);
    fwrite(&corners, sizeof(corners), 1, out);
    fwrite(scores, sizeof(scores), 1, out);
    fwrite(digits, 16, 1, out);
    fwrite(fb, sizeof(fb), 1, out);
    fwrite(fmt, sizeof(fmt), 1, out);
  }
  scanner_destroy(&state);
  dmz_context_destroy(dmz);
  fclose(out);
  return 0;
}
