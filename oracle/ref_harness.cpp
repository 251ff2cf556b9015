/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Unity-includes the reference's own translation unit (dmz_all.cpp, README.md:34-36) from where it
 * lies under /root/reference and exposes extern "C" taps on its (static) stage functions, so tests
 * can run the reference itself.  Linked with oracle/cvshim.cpp + oracle/prims.c into
 * oracle/_ref/libdmz_ref.so (git-ignored; recipe: oracle/Makefile target `ref`).  No reference
 * source is copied into this repository.
 *
 * Build flags (pinned, SURVEY section 7 "hard parts"): g++ -std=gnu++03 -O2, x86-64 baseline SSE2, no
 * -march, -ffp-contract=off; -DCYTHON_DMZ=1 (the only off-device client macro, compile.h:11-25:
 * selects the non-NEON, non-GLES code paths) -DSCAN_EXPIRY=0 -DTEST_GENERATED_MODELS=1.
 */
#include "dmz_all.cpp"

#include <malloc.h>
#include <pthread.h>
#include <time.h>

#include "oracle_types.h"

#if !SCAN_EXPIRY
/* CYTHON_DMZ declares these (dmz.h:103-120) but with SCAN_EXPIRY=0 nothing defines them. */
void expiry_extract_group(IplImage *, GroupedRects &, Eigen::Matrix<float, 11, 10, 1, 11, 10> &, int *, int *) { abort(); }
/* declared static in scan/expiry_seg.h:12 and referenced by the Cython-only dmz_best_expiry_seg */
static void best_expiry_seg(IplImage *, uint16_t, GroupedRectsList &, GroupedRectsList &) { abort(); }
static void expiry_extract(IplImage *, GroupedRectsList &, GroupedRectsList &, int *, int *) { abort(); }
#endif

namespace {

struct Hdr {
  IplImage img;
  IplROI roi;
};

/* Wrap caller memory as a single-channel IplImage (no copy). */
void wrap(Hdr *h, const void *data, int w, int hgt, int step, int depth) {
  memset(h, 0, sizeof(*h));
  h->img.nSize = sizeof(IplImage);
  h->img.nChannels = 1;
  h->img.depth = depth;
  h->img.width = w;
  h->img.height = hgt;
  h->img.widthStep = step;
  h->img.imageSize = step * hgt;
  h->img.imageData = h->img.imageDataOrigin = (char *)data;
  h->img.align = 4;
}

uint32_t card_checksum(const uint8_t *p, size_t n) {
  uint32_t c = 0;
  for (size_t i = 0; i < n; i++) c += (uint32_t)(i + 1) * p[i];
  return c;
}

void flatten_scan(const FrameScanResult &r, orc_scan *out) {
  memset(out, 0, sizeof(*out));
  out->usable = r.usable;
  out->upside_down = r.upside_down;
  memcpy(&out->vseg, &r.vseg, sizeof(orc_vseg));
  if (r.usable || (!r.upside_down && r.vseg.score > 15)) {
    /* hseg/scores are only written once the vseg gate passed (frame.cpp:43-66) */
    memcpy(&out->hseg, &r.hseg, sizeof(orc_hseg));
    for (int i = 0; i < 16; i++)
      for (int j = 0; j < 10; j++) out->scores[i * 10 + j] = r.scores(i, j);
  }
}

}  // namespace

extern "C" {

int ref_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(NVerticalSegmentation);
    case 1: return (int)sizeof(NHorizontalSegmentation);
    case 2: return (int)sizeof(NumberScores);
    case 3: return (int)sizeof(FrameScanResult);
    case 4: return (int)sizeof(ScannerResult);
    case 5: return (int)sizeof(ScannerState);
    case 6: return (int)sizeof(dmz_edges);
    case 7: return (int)sizeof(dmz_corner_points);
    case 8: return (int)sizeof(IplImage);
    case 9: return (int)sizeof(GroupedRects);
  }
  return -1;
}

/* The reference's own embedded known-answer tests (models/generated/modelm_befe75da.cpp:1835-1847,
 * modelc_*.cpp:2039-2051).  Bit i of the result = KAT i passed. */
int ref_run_kats(void) {
  int ok = 0;
  ok |= passm_befe75da() ? 1 : 0;
  ok |= passc_5c241121() ? 2 : 0;
  ok |= passc_01266c1b() ? 4 : 0;
  ok |= passc_b00bf70c() ? 8 : 0;
  return ok;
}

/* D0: detection_boxes_for_sample (dmz.cpp:279-341). out = top,bottom,left,right x {x,y,w,h}. */
void ref_detection_boxes(int w, int h, int orientation, int32_t out[16]) {
  Hdr s;
  wrap(&s, NULL, w, h, (w + 3) & ~3, IPL_DEPTH_8U);
  DetectionBoxes b = detection_boxes_for_sample(&s.img, (FrameOrientation)orientation);
  CvRect r[4] = {b.top, b.bottom, b.left, b.right};
  for (int i = 0; i < 4; i++) {
    out[i * 4 + 0] = r[i].x;
    out[i * 4 + 1] = r[i].y;
    out[i * 4 + 2] = r[i].width;
    out[i * 4 + 3] = r[i].height;
  }
}

/* D1: llcv_sobel7 (cv/sobel.cpp:500) on an isolated w x h image; dx, dy are w*h s16, dense. */
void ref_sobel7(const uint8_t *img, int step, int w, int h, int16_t *dx, int16_t *dy) {
  Hdr s, a, b;
  wrap(&s, img, w, h, step, IPL_DEPTH_8U);
  wrap(&a, dx, w, h, w * 2, IPL_DEPTH_16S);
  wrap(&b, dy, w, h, w * 2, IPL_DEPTH_16S);
  llcv_sobel7(&s.img, &a.img, NULL, 1, 0);
  llcv_sobel7(&s.img, &b.img, NULL, 0, 1);
}

/* D2+D3: llcv_adaptive_canny7_precomputed_sobel (cv/canny.cpp:568). edges is w*h u8 dense. */
void ref_adaptive_canny(const uint8_t *img, int step, int w, int h, const int16_t *dx, const int16_t *dy, uint8_t *edges,
                        int32_t *low, int32_t *high) {
  Hdr s, a, b, e;
  wrap(&s, img, w, h, step, IPL_DEPTH_8U);
  wrap(&a, dx, w, h, w * 2, IPL_DEPTH_16S);
  wrap(&b, dy, w, h, w * 2, IPL_DEPTH_16S);
  wrap(&e, edges, w, h, w, IPL_DEPTH_8U);
  llcv_adaptive_canny7_precomputed_sobel(&s.img, &e.img, &a.img, &b.img);
  double mean = (sum_abs_magnitude(&a.img) + sum_abs_magnitude(&b.img)) / (w * h);
  *low = cvFloor(mean);
  *high = cvFloor(3.0f * mean);
}

/* D1..D4 for one strip exactly as best_line_for_sample (dmz.cpp:224-271) does it. */
void ref_best_line(const uint8_t *img, int step, int w, int h, int vertical, orc_line *out) {
  Hdr s;
  wrap(&s, img, w, h, step, IPL_DEPTH_8U);
  ParametricLine l = best_line_for_sample(&s.img, vertical ? LineOrientationVertical : LineOrientationHorizontal);
  memset(out, 0, sizeof(*out));
  out->found = !is_parametric_line_none(l);
  out->rho = l.rho;
  out->theta = l.theta;
  out->max_votes = -1;
  /* integer taps, recomputed through the same static stage functions */
  int16_t *dx = (int16_t *)malloc((size_t)w * h * 2), *dy = (int16_t *)malloc((size_t)w * h * 2);
  uint8_t *e = (uint8_t *)malloc((size_t)w * h);
  ref_sobel7(img, step, w, h, dx, dy);
  ref_adaptive_canny(img, step, w, h, dx, dy, e, &out->low, &out->high);
  for (int i = 0; i < w * h; i++) out->n_edge_px += e[i] != 0;
  if (out->found) {
    int numrho = 2 * (w + h) + 1;
    float base = vertical ? kVerticalAngle : kHorizontalAngle;
    float theta_min = base - kMaxAngleDeviationAllowed;
    out->r = (int)(l.rho + (numrho - 1) * 0.5f);
    out->n = cvRound((l.theta - theta_min) / ((float)CV_PI / 180.0f));
  }
  free(dx);
  free(dy);
  free(e);
}

/* dmz_detect_edges (dmz.cpp:371-439). cb/cr are (w/2)x(h/2). */
int ref_detect_edges(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                     int orientation, orc_detect *out) {
  Hdr a, b, c;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  wrap(&b, cb, w / 2, h / 2, cstep, IPL_DEPTH_8U);
  wrap(&c, cr, w / 2, h / 2, cstep, IPL_DEPTH_8U);
  dmz_edges e;
  dmz_corner_points p;
  memset(&e, 0, sizeof(e));
  memset(&p, 0, sizeof(p));
  bool ok = dmz_detect_edges(&a.img, &b.img, &c.img, (FrameOrientation)orientation, &e, &p);
  const dmz_found_edge *fe[4] = {&e.top, &e.left, &e.bottom, &e.right};
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < 4; i++) {
    out->found[i] = fe[i]->found;
    out->rho[i] = fe[i]->location.rho;
    out->theta[i] = fe[i]->location.theta;
  }
  if (ok) memcpy(out->corners, &p, sizeof(p));
  out->all_found = ok;
  return ok;
}

/* W1: llcv_calc_persp_transform (cv/warp.cpp:34-125), row-major 3x3. pts = x0,y0,...,x3,y3. */
void ref_calc_persp_transform(const float src_pts[8], const float dst_pts[8], float M[9]) {
  dmz_point s[4], d[4];
  for (int i = 0; i < 4; i++) {
    s[i].x = src_pts[2 * i];
    s[i].y = src_pts[2 * i + 1];
    d[i].x = dst_pts[2 * i];
    d[i].y = dst_pts[2 * i + 1];
  }
  llcv_calc_persp_transform(M, 9, true, s, d);
}

/* W0..W2: dmz_transform_card (dmz.cpp:443-497). corners in dmz_corner_points order. card = 428*270 dense. */
void ref_transform_card(const uint8_t *y, int w, int h, int ystep, const float corners[8], int orientation, uint8_t *card) {
  Hdr a, o;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  wrap(&o, card, 428, 270, 428, IPL_DEPTH_8U);
  dmz_corner_points p;
  memcpy(&p, corners, sizeof(p));
  IplImage *outp = &o.img;
  dmz_transform_card(NULL, &a.img, p, (FrameOrientation)orientation, false, &outp);
}

/* the same with the caller's `upsample` flag (sample = a half-size chroma plane, dmz.cpp:473-481) */
void ref_transform_card_up(const uint8_t *y, int w, int h, int ystep, const float corners[8], int orientation, int upsample, uint8_t *card) {
  Hdr a, o;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  wrap(&o, card, 428, 270, 428, IPL_DEPTH_8U);
  dmz_corner_points p;
  memcpy(&p, corners, sizeof(p));
  IplImage *outp = &o.img;
  dmz_transform_card(NULL, &a.img, p, (FrameOrientation)orientation, upsample != 0, &outp);
}

/* V1+V2: the three softmax outputs for one image row of a 428x270 card (scan/n_vseg.cpp:39-47). */
void ref_vseg_row(const uint8_t *card, int row, float probs[3]) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  IplImage *g = cvCreateImage(cvSize(408, 1), IPL_DEPTH_8U, 1);
  IplImage *d = cvCreateImage(cvSize(204, 1), IPL_DEPTH_8U, 1);
  IplImage *f = cvCreateImage(cvSize(204, 1), IPL_DEPTH_32F, 1);
  cvSetImageROI(&c.img, cvRect(10, row, 408, 1));
  VSegProbabilities p = vseg_probabilities_for_hstrip(&c.img, g, d, f);
  cvResetImageROI(&c.img);
  probs[0] = p(0, 0);
  probs[1] = p(0, 1);
  probs[2] = p(0, 2);
  cvReleaseImage(&g);
  cvReleaseImage(&d);
  cvReleaseImage(&f);
}

/* V0: best_n_vseg (scan/n_vseg.cpp:94-168). */
void ref_best_n_vseg(const uint8_t *card, orc_vseg *out) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  NVerticalSegmentation v = best_n_vseg(&c.img);
  memcpy(out, &v, sizeof(v));
}

/* H0: best_n_hseg (scan/n_hseg.cpp:88-152) on ROI (0, vseg.y_offset, 428, 27). */
void ref_best_n_hseg(const uint8_t *card, const orc_vseg *vseg, orc_hseg *out) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  NVerticalSegmentation v;
  memcpy(&v, vseg, sizeof(v));
  cvSetImageROI(&c.img, cvRect(0, v.y_offset, 428, 27));
  NHorizontalSegmentation hs = best_n_hseg(&c.img, v);
  cvResetImageROI(&c.img);
  memset(out, 0, sizeof(*out));
  memcpy(out, &hs, sizeof(hs));
}

/* C0: number_scores (scan/n_categorize.cpp:75-108). scores = 160 floats. */
void ref_number_scores(const uint8_t *card, int y_offset, const orc_hseg *hseg, float *scores) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  NHorizontalSegmentation hs;
  memcpy(&hs, hseg, sizeof(hs));
  cvSetImageROI(&c.img, cvRect(0, y_offset, 428, 27));
  NumberScores s = number_scores(&c.img, hs);
  cvResetImageROI(&c.img);
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 10; j++) scores[i * 10 + j] = s(i, j);
}

/* C0 prep only: 19x27 ROI of a dense strip -> cross gradient -> equalize -> /255 (n_categorize.cpp:94-99). */
void ref_digit_patch_prep(const uint8_t *img, int step, float *patch /*27*19*/) {
  Hdr s;
  wrap(&s, img, 19, 27, step, IPL_DEPTH_8U);
  IplImage *n = cvCreateImage(cvSize(19, 27), IPL_DEPTH_8U, 1);
  IplImage *f = cvCreateImage(cvSize(19, 27), IPL_DEPTH_32F, 1);
  llcv_morph_grad3_2d_cross_u8(&s.img, n);
  llcv_equalize_hist(n, n);
  cvConvertScale(n, f, 1.0f / 255.0f, 0.0f);
  for (int i = 0; i < 27; i++) memcpy(patch + i * 19, f->imageData + i * f->widthStep, 19 * 4);
  cvReleaseImage(&n);
  cvReleaseImage(&f);
}

/* C2: the three digit CNNs + ensemble on one 27x19 float patch (n_categorize.cpp:45-72).
 * out[0..9] ensemble, out[10..39] the three models' probabilities. */
void ref_digit_models(const float *patch, float *out) {
  NumberImage m;
  for (int i = 0; i < 27; i++)
    for (int j = 0; j < 19; j++) m(i, j) = patch[i * 19 + j];
  SingleNumberScores r0 = applyc_5c241121(m), r1 = applyc_01266c1b(m), r2 = applyc_b00bf70c(m);
  SingleNumberScores mx = r0.cwiseMax(r1).cwiseMax(r2);
  SingleNumberScores e = (r0 + r1 + r2 - mx) / 2.0f;
  for (int j = 0; j < 10; j++) {
    out[j] = e(0, j);
    out[10 + j] = r0(0, j);
    out[20 + j] = r1(0, j);
    out[30 + j] = r2(0, j);
  }
}

/* V2 alone on a prepared 204-float row. */
void ref_vseg_model(const float *in204, float probs[3]) {
  Eigen::Map<const VSegModelInput> x(in204);
  VSegProbabilities p = applym_befe75da(x);
  probs[0] = p(0, 0);
  probs[1] = p(0, 1);
  probs[2] = p(0, 2);
}

/* S0: scan_card_image (scan/frame.cpp:24-81). */
void ref_scan_card_image(const uint8_t *card, orc_scan *out) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  FrameScanResult r;
  r.scores = NumberScores::Zero();
  memset(&r.hseg, 0, sizeof(r.hseg));
  memset(&r.vseg, 0, sizeof(r.vseg));
  r.flipped = false;
  r.focus_score = 0;
  scan_card_image(&c.img, true, false, &r);
  flatten_scan(r, out);
}

/* Whole path for one frame: dmz_detect_edges -> dmz_transform_card -> scan_card_image. */
void ref_process_frame(const uint8_t *y, int w, int h, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep,
                       int orientation, orc_frame_record *rec, uint8_t *card_out /* may be NULL */) {
  memset(rec, 0, sizeof(*rec));
  if (!ref_detect_edges(y, w, h, ystep, cb, cr, cstep, orientation, &rec->detect)) return;
  uint8_t *card = card_out ? card_out : (uint8_t *)malloc(428 * 270);
  ref_transform_card(y, w, h, ystep, rec->detect.corners, orientation, card);
  rec->card_check = card_checksum(card, 428 * 270);
  ref_scan_card_image(card, &rec->scan);
  if (!card_out) free(card);
}

/* S1: scanner session (scan/scan.cpp). Opaque handle owns a ScannerState. */
void *ref_scanner_new(void) {
  ScannerState *s = new ScannerState();
  scanner_initialize(s);
  return s;
}
void ref_scanner_free(void *s) { delete (ScannerState *)s; }
void ref_scanner_reset(void *s) { scanner_reset((ScannerState *)s); }

void ref_scanner_add_frame(void *state, const uint8_t *card, orc_scan *out) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  FrameScanResult r;
  r.scores = NumberScores::Zero();
  memset(&r.hseg, 0, sizeof(r.hseg));
  memset(&r.vseg, 0, sizeof(r.vseg));
  r.flipped = false;
  r.focus_score = 0;
  scanner_add_frame((ScannerState *)state, &c.img, &r);
  if (out) flatten_scan(r, out);
}

/* aggregated15 / aggregated16 / counts, for the EMA parity check */
void ref_scanner_peek(void *state, float agg15[160], float agg16[160], int32_t counts[2]) {
  ScannerState *s = (ScannerState *)state;
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 10; j++) {
      agg15[i * 10 + j] = s->aggregated15(i, j);
      agg16[i * 10 + j] = s->aggregated16(i, j);
    }
  counts[0] = s->count15;
  counts[1] = s->count16;
}

/* scanner_result: returns complete; digits[16] predictions; n_numbers. */
int ref_scanner_result(void *state, uint8_t digits[16], int32_t *n_numbers) {
  ScannerResult r;
  memset(&r.hseg, 0, sizeof(r.hseg));
  memset(&r.vseg, 0, sizeof(r.vseg));
  r.n_numbers = 0;
  for (int i = 0; i < 16; i++) r.predictions(i, 0) = 0;
  scanner_result((ScannerState *)state, &r);
  for (int i = 0; i < 16; i++) digits[i] = (uint8_t)r.predictions(i, 0);
  *n_numbers = r.n_numbers;
  return r.complete;
}

/* dmz_focus_score / dmz_brightness_score (dmz.cpp:183-195) */
float ref_focus_score(const uint8_t *y, int ystep, int w, int h, int use_full_image) {
  Hdr a;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  return dmz_focus_score(&a.img, use_full_image != 0);
}
float ref_brightness_score(const uint8_t *y, int ystep, int w, int h, int use_full_image) {
  Hdr a;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  return dmz_brightness_score(&a.img, use_full_image != 0);
}
void ref_scoring_rect(int w, int h, int use_full_image, int rect[4]) {
  CvSize fs = use_full_image ? cvSize(kCreditCardTargetWidth, kCreditCardTargetHeight)
                             : cvSize(kCreditCardTargetWidth / 3, kCreditCardTargetHeight / 3);
  CvRect r = dmz_card_rect_for_screen(fs, cvSize(kLandscapeSampleWidth, kLandscapeSampleHeight), cvSize(w, h));
  rect[0] = r.x, rect[1] = r.y, rect[2] = r.width, rect[3] = r.height;
}

int ref_luhn(const uint8_t *digits, int n) { return dmz_passes_luhn_checksum((uint8_t *)digits, (uint8_t)n); }
int ref_card_type(const uint8_t *digits, int n) {
  return dmz_card_info_for_prefix_and_length((uint8_t *)digits, (uint8_t)n, false).card_type;
}

/* ------------------------------------------------------------------ CPU baseline timing.
 * frames: n dense w*h Y planes; cb/cr shared flat planes (never produce lines). Each thread owns its
 * images (legal: SCAN_EXPIRY=0 has no hidden statics). Returns wall seconds. */
struct BenchJob {
  const uint8_t *frames, *cb, *cr;
  int w, h, lo, hi, orientation;
  orc_frame_record *recs;
};

static void *bench_worker(void *arg) {
  BenchJob *j = (BenchJob *)arg;
  for (int i = j->lo; i < j->hi; i++)
    ref_process_frame(j->frames + (size_t)i * j->w * j->h, j->w, j->h, j->w, j->cb, j->cr, j->w / 2, j->orientation,
                      &j->recs[i], NULL);
  return NULL;
}

double ref_bench_frames(const uint8_t *frames, int n, int w, int h, const uint8_t *cb, const uint8_t *cr, int orientation,
                        int nthreads, orc_frame_record *recs) {
  if (nthreads < 1) nthreads = 1;
  /* The reference allocates and frees ~20 images per frame (cvCreateImage).  With glibc's defaults the larger
   * ones are mmap()ed and unmapped every time, which serialises the threads in the kernel; keep them on the
   * per-thread arenas instead so the many-core baseline measures the reference's arithmetic, not mmap. */
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  mallopt(M_ARENA_MAX, 256);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  BenchJob *jobs = (BenchJob *)malloc(sizeof(BenchJob) * nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    BenchJob j = {frames, cb, cr, w, h, (int)((long)n * t / nthreads), (int)((long)n * (t + 1) / nthreads), orientation, recs};
    jobs[t] = j;
    pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- pixel formats either side of the path ---------------------------------------------------------------------- */

/* dmz_YCbCr_to_RGB (dmz.cpp:58-64 -> llcv_YCbCr2RGB_u8, cv/convert.cpp:449-504) into caller memory (3 or 4 channels) */
void ref_ycbcr_to_rgb(const uint8_t *y, int ystep, const uint8_t *cb, const uint8_t *cr, int cstep, int w, int h, int channels,
                      uint8_t *dst, int dstep) {
  Hdr a, b, c, d;
  wrap(&a, y, w, h, ystep, IPL_DEPTH_8U);
  wrap(&b, cb, w, h, cstep, IPL_DEPTH_8U);
  wrap(&c, cr, w, h, cstep, IPL_DEPTH_8U);
  wrap(&d, dst, w, h, dstep, IPL_DEPTH_8U);
  d.img.nChannels = channels;
  IplImage *out = &d.img;
  dmz_YCbCr_to_RGB(&a.img, &b.img, &c.img, &out);
}

/* dmz_deinterleave_RGBA_to_R (dmz.cpp:66-109) */
void ref_rgba_to_r(const uint8_t *source, uint8_t *dest, size_t size) { dmz_deinterleave_RGBA_to_R((uint8_t *)source, dest, (int)size); }

/* dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy (dmz.cpp:519-531; cv/sobel.cpp:556-900), kind 0 / 1 / 2 */
void ref_stencil3(const uint8_t *img, int step, int w, int h, int kind, int16_t *out) {
  Hdr a, d;
  wrap(&a, img, w, h, step, IPL_DEPTH_8U);
  wrap(&d, out, w, h, w * (int)sizeof(int16_t), IPL_DEPTH_16S);
  if (kind == 0) dmz_scharr3_dx_abs(&a.img, &d.img);
  else if (kind == 1) dmz_scharr3_dy_abs(&a.img, &d.img);
  else dmz_sobel3_dx_dy(&a.img, &d.img);
}

#define BT(x) ref_##x
#include "bench_taps.inc"
#undef BT

#if SCAN_EXPIRY
/* ---- expiry taps: only in the SCAN_EXPIRY=1 build (oracle/_ref/libdmz_ref_expiry.so) -------------------------- */

/* prepare_image_for_cat (scan/expiry_categorize.cpp:37-73): the 16x11 character crop at (left, top) of a u8 image */
void ref_expiry_patch_prep(const uint8_t *img, int step, int w, int h, int left, int top, float *out176) {
  Hdr a;
  wrap(&a, img, w, h, step, IPL_DEPTH_8U);
  IplImage *as_float = cvCreateImage(cvSize(kTrimmedCharacterImageWidth, kTrimmedCharacterImageHeight), IPL_DEPTH_32F, 1);
  CharacterRectList rects;
  rects.push_back(CharacterRect(top, left, 0));
  prepare_image_for_cat(&a.img, as_float, rects.begin());
  for (int r = 0; r < 16; r++) memcpy(out176 + r * 11, as_float->imageData + (size_t)r * as_float->widthStep, 11 * sizeof(float));
  cvReleaseImage(&as_float);
}

/* digit_probabilities (scan/expiry_categorize.cpp:77-108) = applyc_bf4dd6c8 on a prepared 16x11 float image */
void ref_expiry_digit_model(const float *x176, float *out10) {
  IplImage *as_float = cvCreateImage(cvSize(kTrimmedCharacterImageWidth, kTrimmedCharacterImageHeight), IPL_DEPTH_32F, 1);
  for (int r = 0; r < 16; r++) memcpy(as_float->imageData + (size_t)r * as_float->widthStep, x176 + r * 11, 11 * sizeof(float));
  DigitProbabilities *p = digit_probabilities(as_float);
  for (int i = 0; i < 10; i++) out10[i] = p[0](0, i);
  cvReleaseImage(&as_float);
}

/* slash_probabilities (scan/expiry_seg.cpp:41-46) = applym_730c4cbd */
void ref_slash_model(const float *x176, float *out2) {
  IplImage *as_float = cvCreateImage(cvSize(kTrimmedCharacterImageWidth, kTrimmedCharacterImageHeight), IPL_DEPTH_32F, 1);
  for (int r = 0; r < 16; r++) memcpy(as_float->imageData + (size_t)r * as_float->widthStep, x176 + r * 11, 11 * sizeof(float));
  SlashProbabilities p = slash_probabilities(as_float);
  out2[0] = p(0, 0), out2[1] = p(0, 1);
  cvReleaseImage(&as_float);
}

/* llcv_scharr3_dx_abs (cv/sobel.cpp:700-826) */
void ref_scharr3_dx_abs(const uint8_t *img, int step, int w, int h, int16_t *out) {
  Hdr a, d;
  wrap(&a, img, w, h, step, IPL_DEPTH_8U);
  wrap(&d, out, w, h, w * (int)sizeof(int16_t), IPL_DEPTH_16S);
  llcv_scharr3_dx_abs(&a.img, &d.img);
}

/* best_expiry_seg (scan/expiry_seg.cpp:706-903) on a 428x270 card.  Groups are flattened into int32 records:
 * {top, left, width, height, character_width, pattern, n_rects, (rect.top, rect.left) x n_rects}.  Returns the number
 * of expiry groups (or -1 if `cap` ints were not enough); *n_ints = ints written. */
int ref_best_expiry_seg(const uint8_t *card, int y_offset, int32_t *out, int cap, int *n_ints) {
  Hdr a;
  wrap(&a, card, kCreditCardTargetWidth, kCreditCardTargetHeight, kCreditCardTargetWidth, IPL_DEPTH_8U);
  GroupedRectsList expiry_groups, name_groups;
  best_expiry_seg(&a.img, (uint16_t)y_offset, expiry_groups, name_groups);
  int k = 0;
  for (size_t g = 0; g < expiry_groups.size(); g++) {
    const GroupedRects &G = expiry_groups[g];
    if (k + 7 + 2 * (int)G.character_rects.size() > cap) return -1;
    out[k++] = G.top, out[k++] = G.left, out[k++] = G.width, out[k++] = G.height, out[k++] = G.character_width;
    out[k++] = (int)G.pattern, out[k++] = (int)G.character_rects.size();
    for (size_t r = 0; r < G.character_rects.size(); r++) out[k++] = G.character_rects[r].top, out[k++] = G.character_rects[r].left;
  }
  *n_ints = k;
  return (int)expiry_groups.size();
}

/* scanner_add_frame_with_expiry (scan/scan.cpp:41-86) with the expiry branch live */
void ref_scanner_add_frame_with_expiry(void *state, const uint8_t *card, int scan_expiry, orc_scan *out, int32_t *n_frame_groups) {
  Hdr c;
  wrap(&c, card, 428, 270, 428, IPL_DEPTH_8U);
  FrameScanResult r;
  r.scores = NumberScores::Zero();
  memset(&r.hseg, 0, sizeof(r.hseg));
  memset(&r.vseg, 0, sizeof(r.vseg));
  r.flipped = false;
  r.focus_score = 0;
  scanner_add_frame_with_expiry((ScannerState *)state, &c.img, scan_expiry != 0, &r);
  if (out) flatten_scan(r, out);
  if (n_frame_groups) *n_frame_groups = (int32_t)r.expiry_groups.size();
}

/* the session's expiry state: month, year, and per aggregated group {top, left, recently_seen, total_seen} + the four
 * digit rows (characters 0, 1, 3, 4) of its score matrix.  Returns the number of aggregated groups. */
int ref_scanner_expiry_peek(void *state, int32_t *month, int32_t *year, int32_t *meta, float *scores, int cap) {
  ScannerState *s = (ScannerState *)state;
  *month = s->expiry_month, *year = s->expiry_year;
  int n = 0;
  for (size_t g = 0; g < s->expiry_groups.size() && n < cap; g++, n++) {
    const GroupedRects &G = s->expiry_groups[g];
    meta[n * 4 + 0] = G.top, meta[n * 4 + 1] = G.left, meta[n * 4 + 2] = G.recently_seen_count, meta[n * 4 + 3] = G.total_seen_count;
    const int rows[4] = {0, 1, 3, 4};
    for (int r = 0; r < 4; r++)
      for (int d = 0; d < 10; d++) scores[(n * 4 + r) * 10 + d] = G.scores(rows[r], d);
  }
  return n;
}

/* get_stable_expiry_month_and_year (scan/expiry_categorize.cpp:398-441) on a five-character group with the given
 * 5 x 10 score rows; month / year are in-out (the best date so far). */
void ref_expiry_month_year(const float *scores50, int32_t *month, int32_t *year) {
  GroupedRects g;
  g.pattern = ExpiryPatternMMsYY;
  for (int i = 0; i < 5; i++) g.character_rects.push_back(CharacterRect(0, 12 * i, 0));
  g.scores = ExpiryGroupScores::Zero();
  for (int i = 0; i < 5; i++)
    for (int d = 0; d < 10; d++) g.scores(i, d) = scores50[i * 10 + d];
  int m = *month, y = *year;
  get_stable_expiry_month_and_year(g, &m, &y);
  *month = m, *year = y;
}

/* scanner_result with the expiry fields */
int ref_scanner_result_expiry(void *state, uint8_t digits[16], int32_t *n_numbers, int32_t *month, int32_t *year) {
  ScannerResult r;
  memset(&r.hseg, 0, sizeof(r.hseg));
  memset(&r.vseg, 0, sizeof(r.vseg));
  r.n_numbers = 0;
  r.expiry_month = r.expiry_year = 0;
  for (int i = 0; i < 16; i++) r.predictions(i, 0) = 0;
  scanner_result((ScannerState *)state, &r);
  for (int i = 0; i < 16; i++) digits[i] = (uint8_t)r.predictions(i, 0);
  *n_numbers = r.n_numbers;
  *month = r.expiry_month, *year = r.expiry_year;
  return r.complete;
}
#endif

}  // extern "C"
