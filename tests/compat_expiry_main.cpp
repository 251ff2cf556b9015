// tests/compat_expiry_main.cpp -- a caller written against the reference's scan.h shape that turns expiry scanning on:
// scanner_add_frame_with_expiry(state, card, true, &frame) per warped card, scanner_result after each.  Linked with
// libb200dmz.so by tests/test_gpu_parity.py::test_cxx_dropin_expiry; the same cards go through the reference's
// SCAN_EXPIRY=1 build in the test.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dmz_b200_compat.h"

static void wrap(IplImage *img, uint8_t *data, int w, int h) {
  memset(img, 0, sizeof(*img));
  img->nSize = sizeof(IplImage);
  img->nChannels = 1;
  img->depth = IPL_DEPTH_8U;
  img->width = w, img->height = h;
  img->widthStep = w;
  img->imageSize = w * h;
  img->imageData = img->imageDataOrigin = (char *)data;
  img->align = 4;
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int n = atoi(argv[2]);
  std::vector<uint8_t> cards((size_t)n * 428 * 270);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(cards.data(), 1, cards.size(), f) != cards.size()) return 3;
  fclose(f);
  FILE *out = fopen(argv[3], "wb");
  b200_compat_set_allow_past_expiry(1);  // the reference build it is compared with is the CYTHON_DMZ one
  ScannerState state;
  scanner_initialize(&state);
  for (int k = 0; k < n; k++) {
    IplImage y;
    wrap(&y, cards.data() + (size_t)k * 428 * 270, 428, 270);
    FrameScanResult fr;
    fr.flipped = false;
    fr.focus_score = 0;
    memset(fr.scores.v, 0, sizeof(fr.scores.v));
    scanner_add_frame_with_expiry(&state, &y, true, &fr);
    ScannerResult sr;
    memset(sr.predictions.v, 0, sizeof(sr.predictions.v));
    sr.n_numbers = 0;
    sr.expiry_month = sr.expiry_year = 0;
    scanner_result(&state, &sr);
    int32_t head[8] = {fr.usable, fr.upside_down, state.expiry_month, state.expiry_year, sr.complete, sr.expiry_month, sr.expiry_year,
                       (int32_t)state.expiry_groups.size()};
    fwrite(head, sizeof(head), 1, out);
    for (size_t g = 0; g < 8; g++) {  // fixed-size records: first eight aggregated groups
      int32_t meta[4] = {0, 0, 0, 0};
      float rows[40] = {0};
      if (g < state.expiry_groups.size()) {
        const GroupedRects &G = state.expiry_groups[g];
        meta[0] = G.top, meta[1] = G.left, meta[2] = G.recently_seen_count, meta[3] = G.total_seen_count;
        const int idx[4] = {0, 1, 3, 4};
        for (int r = 0; r < 4; r++) memcpy(rows + r * 10, G.scores + idx[r] * 10, sizeof(float) * 10);
      }
      fwrite(meta, sizeof(meta), 1, out);
      fwrite(rows, sizeof(rows), 1, out);
    }
  }
  scanner_destroy(&state);
  fclose(out);
  return 0;
}
