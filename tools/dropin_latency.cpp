// tools/dropin_latency.cpp -- latency of the SDK's per-frame call sequence through the C++ drop-in layer:
//   dmz_detect_edges -> dmz_transform_card -> scanner_add_frame, each a synchronous call with host IplImages.
// BASELINE configs[0] (one frame); bench.py runs it and puts the JSON line into `single_frame.dropin_sequence`.
// usage: dropin_latency frames.bin n width height reps
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "dmz_b200_compat.h"

static void wrap(IplImage *img, uint8_t *data, int w, int h) {
  memset(img, 0, sizeof(*img));
  img->nSize = sizeof(IplImage);
  img->nChannels = 1;
  img->depth = IPL_DEPTH_8U;
  img->width = w, img->height = h, img->widthStep = w;
  img->imageSize = w * h;
  img->imageData = img->imageDataOrigin = (char *)data;
  img->align = 4;
}

static double median(std::vector<double> v) {
  if (v.empty()) return 0.0;
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main(int argc, char **argv) {
  if (argc < 6) return 2;
  const int n = atoi(argv[2]), w = atoi(argv[3]), h = atoi(argv[4]), reps = atoi(argv[5]);
  std::vector<uint8_t> frames((size_t)n * w * h), chroma((size_t)(w / 2) * (h / 2), 128);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
  fclose(f);
  dmz_context *dmz = dmz_context_create();
  if (!dmz || !dmz->mz) return 4;
  ScannerState state;
  scanner_initialize(&state);
  IplImage *card = NULL;
  std::vector<double> t_focus, t_detect, t_transform, t_scan, t_total;
  int found_n = 0, usable_n = 0;
  typedef std::chrono::steady_clock clk;
  auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  for (int k = 0; k < reps + 32; k++) {
    IplImage y, cb, cr;
    wrap(&y, frames.data() + (size_t)(k % n) * w * h, w, h);
    wrap(&cb, chroma.data(), w / 2, h / 2);
    wrap(&cr, chroma.data(), w / 2, h / 2);
    if (k % 8 == 0) scanner_reset(&state);  // a session per 8 frames, as in the deck
    const clk::time_point t0 = clk::now();
    volatile float fs = dmz_focus_score(&y, false);
    (void)fs;
    const clk::time_point t1 = clk::now();
    dmz_edges edges;
    dmz_corner_points corners;
    const bool found = dmz_detect_edges(&y, &cb, &cr, FrameOrientationLandscapeRight, &edges, &corners);
    const clk::time_point t2 = clk::now();
    if (!found) continue;
    dmz_transform_card(dmz, &y, corners, FrameOrientationLandscapeRight, false, &card);
    const clk::time_point t3 = clk::now();
    FrameScanResult fr;
    fr.flipped = false;
    scanner_add_frame(&state, card, &fr);
    const clk::time_point t4 = clk::now();
    if (k < 32) continue;  // warm-up
    found_n++, usable_n += fr.usable;
    t_focus.push_back(us(t0, t1)), t_detect.push_back(us(t1, t2)), t_transform.push_back(us(t2, t3)), t_scan.push_back(us(t3, t4));
    t_total.push_back(us(t1, t4));
  }
  printf("{\"frames\": %d, \"usable\": %d, \"dmz_focus_score_us\": %.1f, \"dmz_detect_edges_us\": %.1f, \"dmz_transform_card_us\": %.1f, "
         "\"scanner_add_frame_us\": %.1f, \"detect_transform_scan_us\": %.1f, \"what\": \"median wall time per synchronous call, host IplImages "
         "(Y + Cb + Cr planes to dmz_detect_edges), one frame in flight\"}\n",
         found_n, usable_n, median(t_focus), median(t_detect), median(t_transform), median(t_scan), median(t_total));
  if (card) free(card->imageDataOrigin), free(card);
  scanner_destroy(&state);
  dmz_context_destroy(dmz);
  return 0;
}
