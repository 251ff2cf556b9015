/*
 * include/dmz_b200_compat.h -- the reference's C++ entry points, re-declared without OpenCV / Eigen.
 *
 * A caller compiled against the reference's dmz.h / scan/scan.h (iOS CardIOVideoFrame.mm, Android
 * nativeRecognizer.cpp, cython_dmz/dmz.pyx:379-484) links against libb200dmz.so unchanged: the functions below
 * have the same C++-mangled names and the types the same byte layout (checked by tests/test_abi.py against the
 * reference build).  Each forwards to the C ABI of include/b200_dmz.h with a batch of one.
 *
 *   reference declaration                                        file:line
 *   dmz_context_create / destroy / prepare_for_backgrounding     dmz.h:48-54
 *   dmz_found_all_edges, dmz_detect_edges                        dmz.h:82-87
 *   dmz_focus_score, dmz_brightness_score                        dmz.h:77-80
 *   dmz_transform_card                                           dmz.h:96
 *   scanner_initialize / reset / add_frame[_with_expiry] /
 *   result / destroy                                             scan/scan.h:51-72
 *
 * Layouts: IplImage is OpenCV's public C struct (opencv2/core/types_c.h:465-506 in the vendored 2.4.5 headers);
 * NumberScores is Eigen::Matrix<float,16,10,RowMajor> == 16-byte aligned float[160]; NumberPredictions is
 * Eigen::Matrix<ptrdiff_t,16,1> == 16-byte aligned ptrdiff_t[16]; build flags as the host SDK's release build
 * (DMZ_DEBUG off, scan/scan.h:27-30).
 */
#ifndef DMZ_B200_COMPAT_H
#define DMZ_B200_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

/* ---- OpenCV C image header (layout only) ---- */
#ifndef __OPENCV_CORE_TYPES_H__
typedef struct _IplROI {
  int coi, xOffset, yOffset, width, height;
} IplROI;

typedef struct _IplImage {
  int nSize, ID, nChannels, alphaChannel, depth;
  char colorModel[4], channelSeq[4];
  int dataOrder, origin, align, width, height;
  struct _IplROI *roi;
  struct _IplImage *maskROI;
  void *imageId;
  struct _IplTileInfo *tileInfo;
  int imageSize;
  char *imageData;
  int widthStep;
  int BorderMode[4], BorderConst[4];
  char *imageDataOrigin;
} IplImage;
#define IPL_DEPTH_8U 8
#define IPL_DEPTH_SIGN 0x80000000
#define IPL_DEPTH_16S (IPL_DEPTH_SIGN | 16)
#endif

/* ---- dmz_olm.h / dmz.h types ---- */
typedef uint8_t FrameOrientation;
enum { FrameOrientationPortrait = 1, FrameOrientationPortraitUpsideDown = 2, FrameOrientationLandscapeRight = 3, FrameOrientationLandscapeLeft = 4 };

typedef struct { float x, y; } dmz_point;
typedef struct { dmz_point top_left, bottom_left, top_right, bottom_right; } dmz_corner_points;
typedef struct { void *mz; } dmz_context;
typedef struct { float rho, theta; } ParametricLine;
typedef struct { int found; ParametricLine location; } dmz_found_edge;
typedef struct { dmz_found_edge top, left, bottom, right; } dmz_edges;

/* ---- scan types (scan/n_vseg.h, n_hseg.h, n_categorize.h, expiry_types.h, frame.h, scan_analytics.h, scan.h) ---- */
typedef struct {
  float score;
  uint16_t y_offset;
  uint8_t pattern_type;
  uint8_t number_pattern[19];
  uint8_t number_pattern_length;
  uint8_t number_length;
} NVerticalSegmentation;

typedef struct {
  uint8_t n_offsets;
  uint16_t offsets[16];
  float score;
  float number_width;
  uint16_t pattern_offset;
} NHorizontalSegmentation;

struct alignas(16) NumberScores {  /* Eigen::Matrix<float, 16, 10, RowMajor> */
  float v[160];
  float &operator()(int r, int c) { return v[r * 10 + c]; }
  float operator()(int r, int c) const { return v[r * 10 + c]; }
};
struct alignas(16) NumberPredictions {  /* Eigen::Matrix<ptrdiff_t, 16, 1> */
  ptrdiff_t v[16];
  ptrdiff_t &operator()(int r, int = 0) { return v[r]; }
};

struct CharacterRect { int top, left; long sum; };
struct GroupedRects {
  int top, left, width, height;
  bool grouped_yet;
  long sum;
  int character_width;
  std::vector<CharacterRect> character_rects;
  int pattern;         /* enum ExpiryPattern */
  float scores[110];   /* Eigen::Matrix<float, 11, 10, RowMajor>: 440 bytes, not a multiple of 16 -> unaligned */
  int recently_seen_count, total_seen_count;
};
typedef std::vector<GroupedRects> GroupedRectsList;

typedef struct {
  float focus_score;
  NumberScores scores;
  NHorizontalSegmentation hseg;
  NVerticalSegmentation vseg;
  GroupedRectsList expiry_groups;
  GroupedRectsList name_groups;
  bool usable;
  bool upside_down;
  bool flipped;
  float brightness_score;
  uint16_t iso_speed;
  float shutter_speed;
  bool torch_is_on;
} FrameScanResult;

typedef struct {
  uint32_t frame_index;
  std::map<std::string, std::string> frame_values;
} ScanFrameAnalytics;
typedef struct {
  uint32_t num_frames_scanned;
  uint8_t frames_ring_start;
  ScanFrameAnalytics frames_ring[20];
} ScanSessionAnalytics;

typedef struct {
  bool complete;
  NumberPredictions predictions;
  NHorizontalSegmentation hseg;
  NVerticalSegmentation vseg;
  uint8_t n_numbers;
  int expiry_month;
  int expiry_year;
} ScannerResult;

typedef struct ScannerState {
  uint16_t count15;
  uint16_t count16;
  NumberScores aggregated15;
  NumberScores aggregated16;
  ScanSessionAnalytics session_analytics;
  ScannerResult successfulCardNumberResult;
  NHorizontalSegmentation mostRecentUsableHSeg;
  NVerticalSegmentation mostRecentUsableVSeg;
  unsigned long timeOfCardNumberCompletionInMilliseconds;
  bool scan_expiry;
  int expiry_month;
  int expiry_year;
  GroupedRectsList expiry_groups;
  GroupedRectsList name_groups;
} ScannerState;

/* sizes measured on the reference build (oracle/_ref, ref_sizeof): x86-64 / libstdc++ */
static_assert(sizeof(IplImage) == 144, "IplImage layout");
static_assert(sizeof(NVerticalSegmentation) == 28 && sizeof(NHorizontalSegmentation) == 48, "segmentation layout");
static_assert(sizeof(NumberScores) == 640 && sizeof(GroupedRects) == 520, "scores / groups layout");
static_assert(sizeof(FrameScanResult) == 816 && sizeof(ScannerResult) == 240 && sizeof(ScannerState) == 2832, "scan.h layout");
static_assert(sizeof(dmz_edges) == 48 && sizeof(dmz_corner_points) == 32, "dmz.h layout");

/* ---- the entry points (C++ linkage, as in the reference) ---- */
dmz_context *dmz_context_create(void);
void dmz_context_destroy(dmz_context *dmz);
void dmz_prepare_for_backgrounding(dmz_context *dmz);
bool dmz_found_all_edges(dmz_edges found_edges);
bool dmz_detect_edges(IplImage *y_sample, IplImage *cb_sample, IplImage *cr_sample, FrameOrientation orientation,
                      dmz_edges *found_edges, dmz_corner_points *corner_points);
void dmz_transform_card(dmz_context *dmz, IplImage *sample, dmz_corner_points corner_points, FrameOrientation orientation,
                        bool upsample, IplImage **transformed);
void dmz_deinterleave_uint8_c2(IplImage *interleaved, IplImage **channel1, IplImage **channel2); /* dmz.h:64 */
float dmz_focus_score(IplImage *image, bool use_full_image);      /* dmz.h:77 */
int dmz_has_opencv(void);                                          /* dmz.h:60 */
void dmz_deinterleave_RGBA_to_R(uint8_t *source, uint8_t *dest, int size);        /* dmz.h:67 */
void dmz_YCbCr_to_RGB(IplImage *y, IplImage *cb, IplImage *cr, IplImage **rgb);   /* dmz.h:72: allocates *rgb (3 channels) when NULL */
void dmz_scharr3_dx_abs(IplImage *src, IplImage *dst);             /* dmz.h:105-107 (CYTHON_DMZ): u8 -> IPL_DEPTH_16S */
void dmz_scharr3_dy_abs(IplImage *src, IplImage *dst);
void dmz_sobel3_dx_dy(IplImage *src, IplImage *dst);
/* dmz.h:103-120 (the CYTHON_DMZ-only block): the expiry segmentation entry point of cython_dmz/dmz.pyx.  Layout of
 * CythonGroupedRects as in scan/expiry_types.h:95-118. */
typedef struct {
  int top;
  int left;
} CythonCharacterRect;
typedef float CythonGroupScores[11][10];
typedef struct {
  int top, left, width, height, character_width;
  uint8_t pattern;
  CythonGroupScores scores;
  int recently_seen_count, total_seen_count;
  int number_of_character_rects;
  CythonCharacterRect *character_rects;
} CythonGroupedRects;
static_assert(sizeof(CythonGroupedRects) == 488, "CythonGroupedRects layout");
void dmz_best_expiry_seg(IplImage *card_y, uint16_t starting_y_offset, CythonGroupedRects **expiry_groups, uint16_t *number_of_groups);
float dmz_brightness_score(IplImage *image, bool use_full_image); /* dmz.h:80 */
void scanner_initialize(ScannerState *state);
void scanner_reset(ScannerState *state);
void scanner_add_frame(ScannerState *state, IplImage *y, FrameScanResult *result);
void scanner_add_frame_with_expiry(ScannerState *state, IplImage *y, bool scan_expiry, FrameScanResult *result);
void scanner_result(ScannerState *state, ScannerResult *result);
void scanner_destroy(ScannerState *state);
/* Not in the reference: its DMZ_DEBUG / CYTHON_DMZ builds also accept expiry dates in the past
 * (scan/expiry_categorize.cpp:378-395); 0 (default) = SDK behaviour, 1 = that variant. */
void b200_compat_set_allow_past_expiry(int allow);

#endif
