// card.io-dmz_b200/csrc/b200_internal.h -- shared declarations of the B200-native hot path.
//
// Layout of the per-batch device state (all in HBM, owned by b200_ctx, sized for `cap` frames):
//   d_frames   cap x (H x W) u8        staging copy of the caller's Y planes (B200_MEM_HOST only)
//   d_lines    3 planes x cap x 4      b200_line   strip taps (Y, Cb, Cr)
//   d_geom     cap                      FrameGeom   edges, corners, float M, double inverse M
//   d_cards    cap x (270 x 428) u8    warped cards
//   d_vprob    cap x 270 x 2 f32       per-row (visa-like, amex-like) probabilities, 0 = not computed
//   d_scan     cap                      b200_scan
//   d_records  cap                      b200_frame_record
#ifndef B200_INTERNAL_H
#define B200_INTERNAL_H

#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "b200_dmz.h"

#define B200_NUMANGLE 10  // cvRound((theta_max - theta_min) / theta) for the +-5 degree window, hough.cpp:98

// One detection strip of one plane (dmz.cpp:279-341 geometry + hough.cpp:98-150 constants).  All float /
// trig derived values are computed on the HOST with the C library (b200_tables.cpp) so that libm
// differences between host and device can never leak into the integer results.
struct StripDesc {
  int x, y, w, h;
  int vertical;    // expected line orientation (LineOrientationVertical == 0 in the reference; here 1 = vertical)
  int numrho;      // 2 (w + h) + 1
  int half;        // (numrho - 1) / 2
  int threshold;   // max(w, h) / 6, dmz.cpp:246
  int ncells;      // compacted accumulator size
  int chunk_rows;  // Sobel work split: rows per work item
  int nchunks;
  int tab_sin[B200_NUMANGLE];
  int tab_cos[B200_NUMANGLE];
  int rlo[B200_NUMANGLE];        // smallest accumulator rho index reachable for angle n
  int rcount[B200_NUMANGLE];     // number of reachable rho indices
  int cell_base[B200_NUMANGLE];  // start of angle n's cells in the compacted accumulator
  float slope_a, slope_b;        // tanf bounds of the gradient gate, hough.cpp:117-124
  float theta[B200_NUMANGLE];    // n * theta + theta_min
};

struct DetectParams {
  StripDesc strip[4];  // detection order: top, bottom, left, right
  int use_global_grad; // dx/dy live in a global scratch slab instead of shared memory (very large strips)
};

// Geometry constants for one (resolution, orientation): everything lineByShiftingOrigin / parametricIntersect
// need that involves libm (geometry.cpp:14-43).
struct GeomParams {
  double delta_rho[3][4][B200_NUMANGLE];  // [plane][strip in detection order][n]
  float cos_t[2][B200_NUMANGLE];          // [0 = horizontal line, 1 = vertical line][n]
  float sin_t[2][B200_NUMANGLE];
  float theta[2][B200_NUMANGLE];
  int orientation;
  int n_planes;
};

struct FrameGeom {
  int32_t found[4];  // top, left, bottom, right
  float rho[4];
  float theta[4];
  int32_t n_idx[4];  // Hough angle index of the accepted line per edge
  float corners[8];  // tl, bl, tr, br
  int32_t all_found;
  float M[9];        // llcv_calc_persp_transform output (src -> dst)
  int32_t pad;
  double Minv[9];    // cv::invert of M promoted to double (dst -> src), used by the warp
};

// The 3 x 8 convolution kernels and biases of the digit CNNs travel BY VALUE inside NetWeights, i.e. in the kernel
// parameter space (constant bank 0): warp-uniform constant-cache reads like a __constant__ array, but per launch and so
// per context -- two contexts with different weight directories on one device never share them.
struct ConvConsts {
  float w[3][8][9];
  float b[3][8];
};

struct NetWeights {  // device pointers (+ the by-value convolution constants)
  const float *vseg;     // modelm_befe75da blob: hidden W 50x204, hidden b 50, logistic W 3x50, logistic b 3
  const float *cnn[3];   // modelc blobs: conv W 8x9, conv b 8, hidden W 32x320, hidden b 32, logistic W 10x32, b 10
  const float *cnn_hwT;  // 3 x [320][32] transposed hidden weights (built on the host)
  const float *vseg_norm;  // [256][256][2]: cvNormalize(MINMAX 0..1) scale / shift of a row whose 8-bit min / max are (mn, mx)
  // tensor-core form of the vseg hidden layer (vseg_mma.cu; built on the host by b200_build_vseg_mma_tables)
  const int8_t *vseg_wq;   // [4 digits][14 K chunks][64 units][16] signed base-128 digits of W1, operand layout of umma.cuh
  const float *vseg_unit;  // [64] VsegUnit
  const float *vseg_sd;    // [256][256][2]: (s, d0) of a row with 8-bit min / max (mn, mx): x_k = (v_k - mn) * s + d0
  // tensor-core form of the digit CNNs (categorize_mma.cu; built on the host by b200_build_cnn_mma_tables)
  const int8_t *cnn_convb;   // [3 models][2 K chunks][3 digits x 80][16]: conv taps as signed base-128 digits, operand layout
  const float *cnn_convf;    // [24] (1/255) / F per (model, kernel), then [24] conv biases
  const uint16_t *cnn_hidb;  // fp16 [3 models][Whi, Wlo][40 cells][32 units][8 kernels]
  ConvConsts conv;
};

struct VsegUnit {  // per hidden unit u (zeros for the padding units 50 .. 63)
  float cu;        // S_u * 2^-27, S_u = max_k |W1[u][k]|: W1[u][k] = cu * Q[u][k], Q a 28-bit integer
  float sumw;      // sum_k W1[u][k]
  float b1;
  float w20, w21, w22;  // logistic weights of the three classes
  float pad[2];
};
void b200_build_vseg_mma_tables(const float *blob /* modelm_befe75da */, int8_t *wq /* 4 * 14 * 64 * 16 */, VsegUnit *units /* 64 */,
                                float *sd /* 256 * 256 * 2 */);  // b200_tables.cpp
void b200_build_cnn_mma_tables(const float *const blobs[3] /* modelc_* */, int8_t *convb /* 23040 */, float *convf /* 48 */,
                               uint16_t *hidb /* 3 * 2 * 40 * 32 * 8 */);  // b200_tables.cpp
int launch_categorize_mma(const NetWeights &wts, const uint8_t *q8, b200_scan *scans, int n, bool raw, float *raw_out, cudaStream_t s);  // categorize_mma.cu
int launch_vseg_rows_mma(const NetWeights &wts, const uint8_t *cards, const uint8_t *gate, const b200_scan *scans, int n, int mode,
                         float *vprob, cudaStream_t s);  // vseg_mma.cu

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting, and one process may hold contexts on several
// devices: remember per kernel which devices have been configured (configuring twice is harmless, so no lock).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <typename F>
  bool ensure(F &&configure) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return true;
    if (!configure()) return false;
    done.fetch_or(bit, std::memory_order_release);
    return true;
  }
};

// launchers (each returns the number of kernels it launched, or -1 after setting a CUDA error)
int launch_detect(const DetectParams &p, const uint8_t *plane, int row_stride, size_t frame_stride, int n,
                  const b200_line *prev_lines, const b200_line *prev_lines2, b200_line *lines, int16_t *grad_scratch,
                  cudaStream_t s, int ox = 0, int oy = 0);
int launch_geometry(const GeomParams &g, const b200_line *lines, size_t plane_stride, int n, FrameGeom *geom, cudaStream_t s);
int launch_homography_only(const float *src_pts, const float *dst_pts, int n, float *M, cudaStream_t s);
int launch_corners_to_geom(const b200_corner_points *corners, const uint8_t *valid, int n, int orientation, int upsample,
                           FrameGeom *geom, cudaStream_t s);
// W2 (warp.cu).  The planes the warp reads: n buffers of bw x bh pixels whose pixel (0, 0) sits at (ox, oy) of a
// frame_w x frame_h frame (the whole plane: ox = oy = 0, bw x bh = the frame; host-buffer path: the uploaded crop).
// Taps outside the buffer read as 0 (BORDER_CONSTANT for whole planes; for crops the caller redoes such frames).
struct WarpSource {
  const uint8_t *base;
  int row_stride;
  size_t frame_stride;
  int bw, bh, ox, oy, frame_w, frame_h, n;
};
enum { WARP_FULL = 0, WARP_COARSE = 1, WARP_FINE = 2, WARP_STRIP = 3 };  // which card rows a launch produces (warp.cu)
int launch_warp(const WarpSource &src, const FrameGeom *geom, uint8_t *cards, unsigned int *card_check, int mode,
                const b200_scan *scans, const uint16_t *coarse_y, int portrait, cudaStream_t s);
// lazy cards: launch_scan warps only the rows scan_card_image reads, at the points of the sequence where they are known
struct LazyWarp {
  WarpSource src;
  int portrait;
};
int launch_scan(const NetWeights &wts, const uint8_t *cards, int n, const FrameGeom *geom_or_null, const uint8_t *valid,
                float *vprob, uint8_t *q8 /* n * 16 * B200_Q8_STRIDE bytes: prepared digit patches */, b200_scan *scans, cudaStream_t s,
                cudaEvent_t ev_vseg, cudaEvent_t ev_hseg, cudaEvent_t ev_cat, cudaEvent_t ev_fin, const LazyWarp *lazy = nullptr,
                uint8_t *lazy_cards = nullptr, cudaEvent_t *lazy_ev = nullptr);
int launch_finalize_records(const FrameGeom *geom, const b200_scan *scans, const unsigned int *card_check, int n,
                            b200_frame_record *recs, cudaStream_t s, uint8_t *needs_full = nullptr, int cx0 = 0, int cy0 = 0,
                            int cx1 = 0, int cy1 = 0);
#define B200_EXPIRY_C2K_OFFSET 74408  // floats: modelc_bf4dd6c8 blob (74406) rounded up to a 16-byte boundary
#define B200_EXPIRY_C2H_OFFSET (B200_EXPIRY_C2K_OFFSET + 50 * 40 * 28)  // floats: start of the fp16 layer-2 weight slices (expiry_mma.cu)
#define B200_EXPIRY_C2H_FLOATS (13 * 2 * 14 * 48 * 8 / 2)                // 139 776 halfs
#define B200_EXPIRY_HWT_OFFSET (B200_EXPIRY_C2H_OFFSET + B200_EXPIRY_C2H_FLOATS)  // floats: hidden weights transposed [120][176]
#define B200_Q8_STRIDE 528  // bytes per prepared digit patch (27 x 19 = 513 padded to 33 x 16)
int launch_categorize_patches(const NetWeights &wts, const uint8_t *patches, const float *float_patches, int n, float *out,
                              uint8_t *q8 /* n * B200_Q8_STRIDE bytes of scratch when `patches` is given */, cudaStream_t s);
int launch_vseg_model(const NetWeights &wts, const float *rows, int n, float *out, cudaStream_t s);
int launch_vseg_coarse_rows(const NetWeights &wts, const uint8_t *cards, int n, float *vprob /* n * 540, zeroed here */, cudaStream_t s);
void fill_conv_constants(const float *cnn_blobs[3], ConvConsts *out);  // host side (nets.cu)
int upload_bilateral_tables(const float *color256, const float *space5);  // E0 prep constants (nets.cu)
#define B200_EXPIRY_SEG_SCRATCH_INTS 1584  // per card: 270 row sums, the picked stripes, 3 x 428 column sums (expiry_seg.cu)
int launch_expiry_seg(const uint8_t *cards, const uint16_t *y_offsets, int n, const float *slash_w, int16_t *sob, int32_t *line_sum,
                      b200_expiry_group *groups, int max_groups, int32_t *n_groups, int32_t *n_dropped, cudaStream_t s);  // expiry_seg.cu
int launch_deinterleave_c2(const uint8_t *src, int row_stride, size_t frame_stride, int w, int h, int n, uint8_t *ch1, uint8_t *ch2,
                           cudaStream_t s);
// formats.cu: pixel formats either side of the path (dmz_YCbCr_to_RGB, dmz_deinterleave_RGBA_to_R, the Cython stencils)
int launch_ycbcr_to_rgb(const uint8_t *y, int yrs, size_t yfs, const uint8_t *cb, const uint8_t *cr, int crs, size_t cfs, int w, int h,
                        int n, int channels, uint8_t *dst /* n x h x w x channels, dense */, cudaStream_t s);
int launch_rgba_to_r(const uint8_t *src, size_t n_px, uint8_t *dst, cudaStream_t s);
int launch_stencil3(const uint8_t *src, int row_stride, size_t frame_stride, int w, int h, int n, int kind /* 0 scharr dx abs, 1 scharr dy abs, 2 sobel dx dy */,
                    int16_t *out /* n x h x w, dense */, cudaStream_t s);
int launch_frame_scores(const uint8_t *frames, int row_stride, size_t frame_stride, int n, int rx, int ry, int rw, int rh,
                        float *focus, float *brightness, cudaStream_t s);
void b200_scoring_rect(int w, int h, int use_full_image, int rect[4]);  // b200_tables.cpp
int launch_expiry_digits(const float *weights, const uint8_t *patches, const float *prepared, int n, float *out, cudaStream_t s,
                         const int32_t *where = nullptr);
void b200_build_expiry_c2_halfs(const float *blob /* modelc_bf4dd6c8 */, uint16_t *out /* 13 * 2 * 14 * 48 * 8 */);  // b200_tables.cpp
int launch_expiry_digits_mma(const float *weights, const uint8_t *patches, int n, float *out, cudaStream_t s, const int32_t *where);  // expiry_mma.cu
int upload_bilateral_tables_mma(const float *color256, const float *space5);
void b200_build_minmax_norm_table(float *table /* 256 * 256 * 2 */);  // b200_tables.cpp
void b200_build_bilateral_tables(float *color256, float *space5);  // b200_tables.cpp  // __constant__ conv kernels / biases (nets.cu)

size_t detect_smem_bytes(const DetectParams &p);

// host tables (b200_tables.cpp, compiled WITHOUT fp contraction)
void b200_build_detect_params(int width, int height, int orientation, int block_threads, DetectParams *out);
void b200_build_geom_params(int width, int height, int orientation, int n_planes, GeomParams *out);

#endif
