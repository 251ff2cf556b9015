/*
 * oracle/cvshim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference links against the OpenCV 2.4.x *libraries* (core, imgproc), which are not under
 * /root/reference (only headers; the bundled .a files are macOS Mach-O -- SURVEY 8c).  This file
 * supplies the ~35 cv* C-API symbols the reference's unity build (dmz_all.cpp) leaves undefined, for
 * the single-channel u8 / s16 / f32 cases the hot path uses, by wrapping oracle/prims.c.  It is
 * compiled against the reference's own vendored headers where they lie (-I/root/reference) and is
 * linked only into oracle/_ref/libdmz_ref.so.  Anything outside the supported cases aborts loudly.
 */
#include "opencv2/core/core_c.h"
#include "opencv2/core/core.hpp"
#include "opencv2/imgproc/imgproc_c.h"
#include "opencv2/imgproc/imgproc.hpp"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "prims.h"

#ifndef CV_IMPL
#define CV_IMPL extern "C"
#endif

#define SHIM_FAIL(msg)                                                          \
  do {                                                                          \
    fprintf(stderr, "cvshim: unsupported call %s: %s\n", __FUNCTION__, (msg)); \
    abort();                                                                    \
  } while (0)

static int depth_bytes(int depth) { return (depth & 255) >> 3; }

static int ipl2cv_depth(int depth) {
  switch (depth) {
    case IPL_DEPTH_8U: return CV_8U;
    case IPL_DEPTH_8S: return CV_8S;
    case IPL_DEPTH_16U: return CV_16U;
    case IPL_DEPTH_16S: return CV_16S;
    case IPL_DEPTH_32S: return CV_32S;
    case IPL_DEPTH_32F: return CV_32F;
    case IPL_DEPTH_64F: return CV_64F;
  }
  SHIM_FAIL("depth");
  return 0;
}

/* A resolved single-plane view: origin of the ROI (or whole image), step in bytes, size, cv depth, channels. */
struct View {
  uchar *p;
  int step, w, h, depth, cn;
};

static View view_of(const CvArr *arr) {
  View v;
  if (CV_IS_IMAGE_HDR(arr)) {
    const IplImage *img = (const IplImage *)arr;
    v.depth = ipl2cv_depth(img->depth);
    v.cn = img->nChannels;
    v.step = img->widthStep;
    if (img->roi) {
      v.p = (uchar *)img->imageData + (size_t)img->roi->yOffset * img->widthStep +
            (size_t)img->roi->xOffset * depth_bytes(img->depth) * img->nChannels;
      v.w = img->roi->width;
      v.h = img->roi->height;
    } else {
      v.p = (uchar *)img->imageData;
      v.w = img->width;
      v.h = img->height;
    }
    return v;
  }
  if (CV_IS_MAT_HDR(arr)) {
    const CvMat *m = (const CvMat *)arr;
    v.p = m->data.ptr;
    v.step = m->step ? m->step : CV_ELEM_SIZE(m->type) * m->cols;
    v.w = m->cols;
    v.h = m->rows;
    v.depth = CV_MAT_DEPTH(m->type);
    v.cn = CV_MAT_CN(m->type);
    return v;
  }
  SHIM_FAIL("array kind");
  return v;
}

/* ---------------------------------------------------------------- images / headers */

CV_IMPL IplImage *cvCreateImageHeader(CvSize size, int depth, int channels) {
  IplImage *img = (IplImage *)calloc(1, sizeof(IplImage));
  img->nSize = sizeof(IplImage);
  img->nChannels = channels;
  img->depth = depth;
  img->width = size.width;
  img->height = size.height;
  img->align = 4;
  img->dataOrder = 0;
  img->origin = 0;
  memcpy(img->colorModel, channels == 1 ? "GRAY" : "RGB\0", 4);
  memcpy(img->channelSeq, channels == 1 ? "GRAY" : "BGR\0", 4);
  /* CV_DEFAULT_IMAGE_ROW_ALIGN = 4 */
  img->widthStep = (((size.width * channels * (depth & ~IPL_DEPTH_SIGN) + 7) / 8) + 3) & ~3;
  img->imageSize = img->widthStep * img->height;
  return img;
}

CV_IMPL IplImage *cvCreateImage(CvSize size, int depth, int channels) {
  IplImage *img = cvCreateImageHeader(size, depth, channels);
  void *mem = NULL;
  /* cvCreateImage -> cv::fastMalloc: a 16-byte aligned malloc; contents uninitialised */
  if (posix_memalign(&mem, 16, (size_t)img->imageSize + 16) != 0) SHIM_FAIL("alloc");
  img->imageData = img->imageDataOrigin = (char *)mem;
  return img;
}

CV_IMPL void cvReleaseImageHeader(IplImage **image) {
  if (image && *image) {
    free((*image)->roi);
    free(*image);
    *image = NULL;
  }
}

CV_IMPL void cvReleaseImage(IplImage **image) {
  if (image && *image) {
    free((*image)->imageDataOrigin);
    cvReleaseImageHeader(image);
  }
}

CV_IMPL void cvSetData(CvArr *arr, void *data, int step) {
  if (!CV_IS_IMAGE_HDR(arr)) SHIM_FAIL("only IplImage");
  IplImage *img = (IplImage *)arr;
  int min_step = depth_bytes(img->depth) * img->width * img->nChannels;
  if (step != CV_AUTOSTEP && img->height > 1) img->widthStep = step;
  else img->widthStep = min_step;
  img->imageSize = img->widthStep * img->height;
  img->imageData = img->imageDataOrigin = (char *)data;
}

CV_IMPL void cvSetImageROI(IplImage *image, CvRect rect) {
  /* cvSetImageROI clips the rectangle to the image (core/array.cpp, 2.4.x) */
  if (rect.width < 0 || rect.height < 0 || rect.x >= image->width || rect.y >= image->height ||
      rect.x + rect.width < (int)(rect.width > 0) || rect.y + rect.height < (int)(rect.height > 0))
    SHIM_FAIL("ROI outside image");
  rect.width += rect.x;
  rect.height += rect.y;
  rect.x = rect.x > 0 ? rect.x : 0;
  rect.y = rect.y > 0 ? rect.y : 0;
  rect.width = rect.width < image->width ? rect.width : image->width;
  rect.height = rect.height < image->height ? rect.height : image->height;
  rect.width -= rect.x;
  rect.height -= rect.y;
  if (!image->roi) image->roi = (IplROI *)calloc(1, sizeof(IplROI));
  image->roi->coi = 0;
  image->roi->xOffset = rect.x;
  image->roi->yOffset = rect.y;
  image->roi->width = rect.width;
  image->roi->height = rect.height;
}

CV_IMPL void cvResetImageROI(IplImage *image) {
  if (image->roi) {
    free(image->roi);
    image->roi = NULL;
  }
}

CV_IMPL CvRect cvGetImageROI(const IplImage *image) {
  if (image->roi) return cvRect(image->roi->xOffset, image->roi->yOffset, image->roi->width, image->roi->height);
  return cvRect(0, 0, image->width, image->height);
}

CV_IMPL CvSize cvGetSize(const CvArr *arr) {
  View v = view_of(arr);
  return cvSize(v.w, v.h);
}

CV_IMPL CvMat *cvGetMat(const CvArr *arr, CvMat *header, int *coi, int allowND) {
  (void)allowND;
  if (coi) *coi = 0;
  if (CV_IS_MAT_HDR(arr)) return (CvMat *)arr;
  View v = view_of(arr);
  int type = CV_MAKETYPE(v.depth, v.cn);
  int min_step = CV_ELEM_SIZE(type) * v.w;
  memset(header, 0, sizeof(*header));
  header->type = CV_MAT_MAGIC_VAL | type | ((v.h == 1 || v.step == min_step) ? CV_MAT_CONT_FLAG : 0);
  header->rows = v.h;
  header->cols = v.w;
  header->step = v.step;
  header->data.ptr = v.p;
  header->refcount = NULL;
  header->hdr_refcount = 0;
  return header;
}

CV_IMPL CvMat *cvCreateMat(int rows, int cols, int type) {
  CvMat *m = (CvMat *)calloc(1, sizeof(CvMat));
  type = CV_MAT_TYPE(type);
  m->step = CV_ELEM_SIZE(type) * cols;
  m->type = CV_MAT_MAGIC_VAL | CV_MAT_CONT_FLAG | type;
  m->rows = rows;
  m->cols = cols;
  m->data.ptr = (uchar *)calloc((size_t)rows * m->step + 64, 1);
  m->hdr_refcount = 1;
  return m;
}

CV_IMPL void cvReleaseMat(CvMat **mat) {
  if (mat && *mat) {
    free((*mat)->data.ptr);
    free(*mat);
    *mat = NULL;
  }
}

/* ---------------------------------------------------------------- core arithmetic */

CV_IMPL void cvAbsDiffS(const CvArr *srcarr, CvArr *dstarr, CvScalar value) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (s.depth != CV_16S || d.depth != CV_16S || s.cn != 1 || value.val[0] != 0) SHIM_FAIL("only cvAbs on s16");
  for (int y = 0; y < s.h; y++) {
    const short *sp = (const short *)(s.p + (size_t)y * s.step);
    short *dp = (short *)(d.p + (size_t)y * d.step);
    for (int x = 0; x < s.w; x++) {
      int v = sp[x] < 0 ? -(int)sp[x] : sp[x];
      dp[x] = (short)(v > 32767 ? 32767 : v);
    }
  }
}

CV_IMPL CvScalar cvSum(const CvArr *arr) {
  View s = view_of(arr);
  double total = 0;
  if (s.cn != 1) SHIM_FAIL("channels");
  for (int y = 0; y < s.h; y++) {
    const uchar *row = s.p + (size_t)y * s.step;
    for (int x = 0; x < s.w; x++) {
      switch (s.depth) {
        case CV_8U: total += row[x]; break;
        case CV_16S: total += ((const short *)row)[x]; break;
        case CV_32F: total += ((const float *)row)[x]; break;
        default: SHIM_FAIL("depth");
      }
    }
  }
  return cvScalar(total);
}

CV_IMPL CvScalar cvAvg(const CvArr *arr, const CvArr *mask) {
  View s = view_of(arr);
  if (mask || s.depth != CV_8U || s.cn != 1) SHIM_FAIL("only unmasked u8");
  return cvScalar(orc_mean_u8(s.p, s.step, s.w, s.h));
}

CV_IMPL void cvAvgSdv(const CvArr *arr, CvScalar *mean, CvScalar *std_dev, const CvArr *mask) {
  View s = view_of(arr);
  if (mask || s.cn != 1 || s.depth != CV_16S) SHIM_FAIL("only unmasked single-channel s16");
  double m, sd;
  orc_mean_stddev_s16((const int16_t *)s.p, s.step, s.w, s.h, &m, &sd);
  if (mean) *mean = cvScalar(m);
  if (std_dev) *std_dev = cvScalar(sd);
}

CV_IMPL void cvConvertScale(const CvArr *srcarr, CvArr *dstarr, double scale, double shift) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (s.cn != 1 || d.cn != 1 || s.w != d.w || s.h != d.h) SHIM_FAIL("shape");
  if (s.depth == CV_8U && d.depth == CV_32F && shift == 0) {
    orc_convert_scale_u8_f32(s.p, s.step, s.w, s.h, (float *)d.p, d.step, (float)scale);
    return;
  }
  if (s.depth == CV_16S && d.depth == CV_32F) {
    float fs = (float)scale, fb = (float)shift;
    for (int y = 0; y < s.h; y++) {
      const short *sp = (const short *)(s.p + (size_t)y * s.step);
      float *dp = (float *)(d.p + (size_t)y * d.step);
      for (int x = 0; x < s.w; x++) {
        volatile float p = (float)sp[x] * fs;
        dp[x] = p + fb;
      }
    }
    return;
  }
  SHIM_FAIL("depth combination");
}

CV_IMPL void cvNormalize(const CvArr *srcarr, CvArr *dstarr, double a, double b, int norm_type, const CvArr *mask) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (!mask && norm_type == CV_C && s.depth == CV_16S && d.depth == CV_16S && s.cn == 1 && s.w == d.w && s.h == d.h) {
    /* cv::normalize(NORM_INF): scale = a / max|v| (0 when max|v| <= DBL_EPSILON), then convertTo(16S -> 16S, scale):
     * cvtScale_<short, short, float> -- float multiply, cvRound, saturate (core/convert.cpp, 2.4.x).
     * (scan/expiry_seg.cpp:277 call site) */
    double nrm = 0;
    for (int y = 0; y < s.h; y++) {
      const short *sp = (const short *)(s.p + (size_t)y * s.step);
      for (int x = 0; x < s.w; x++) {
        double v = sp[x] < 0 ? -(double)sp[x] : (double)sp[x];
        if (v > nrm) nrm = v;
      }
    }
    double scale = nrm > DBL_EPSILON ? a / nrm : 0.;
    float fs = (float)scale;
    for (int y = 0; y < s.h; y++) {
      const short *sp = (const short *)(s.p + (size_t)y * s.step);
      short *dp = (short *)(d.p + (size_t)y * d.step);
      for (int x = 0; x < s.w; x++) {
        volatile float p = (float)sp[x] * fs;
        int r = orc_cv_round((double)p);
        dp[x] = (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
      }
    }
    return;
  }
  if (mask || norm_type != CV_MINMAX || s.depth != CV_32F || d.depth != CV_32F || s.cn != 1 || a != 0.0 || b != 1.0)
    SHIM_FAIL("only f32 MINMAX [0,1]");
  if (s.p != d.p) {
    for (int y = 0; y < s.h; y++) memcpy(d.p + (size_t)y * d.step, s.p + (size_t)y * s.step, (size_t)s.w * 4);
  }
  orc_normalize_minmax_f32((float *)d.p, d.step, d.w, d.h);
}

CV_IMPL void cvReduce(const CvArr *srcarr, CvArr *dstarr, int dim, int op) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (dim != 0 || op != CV_REDUCE_SUM || s.depth != CV_8U || d.depth != CV_32F || s.cn != 1 || d.w != s.w || d.h != 1)
    SHIM_FAIL("only u8->f32 column sums");
  orc_reduce_cols_sum_u8_f32(s.p, s.step, s.w, s.h, (float *)d.p);
}

CV_IMPL void cvSplit(const CvArr *srcarr, CvArr *d0, CvArr *d1, CvArr *d2, CvArr *d3) {
  View s = view_of(srcarr);
  CvArr *dst[4] = {d0, d1, d2, d3};
  if (s.depth != CV_8U) SHIM_FAIL("depth");
  for (int c = 0; c < s.cn && c < 4; c++) {
    if (!dst[c]) continue;
    View d = view_of(dst[c]);
    for (int y = 0; y < s.h; y++)
      for (int x = 0; x < s.w; x++) d.p[(size_t)y * d.step + x] = s.p[(size_t)y * s.step + x * s.cn + c];
  }
}

CV_IMPL void cvCopy(const CvArr *srcarr, CvArr *dstarr, const CvArr *mask) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (mask || s.depth != d.depth || s.cn != d.cn || s.w != d.w || s.h != d.h) SHIM_FAIL("cvCopy shape");
  size_t row = (size_t)s.w * s.cn * (s.depth == CV_8U ? 1 : s.depth == CV_16S ? 2 : 4);
  for (int y = 0; y < s.h; y++) memcpy(d.p + (size_t)y * d.step, s.p + (size_t)y * s.step, row);
}

CV_IMPL void cvSetZero(CvArr *arr) {
  View d = view_of(arr);
  size_t row = (size_t)d.w * d.cn * (d.depth == CV_8U ? 1 : d.depth == CV_16S ? 2 : 4);
  for (int y = 0; y < d.h; y++) memset(d.p + (size_t)y * d.step, 0, row);
}

/* ---------------------------------------------------------------- imgproc */

CV_IMPL void cvSmooth(const CvArr *srcarr, CvArr *dstarr, int smoothtype, int size1, int size2, double sigma1, double sigma2) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (smoothtype != CV_BILATERAL || size1 != size2 || s.depth != CV_8U || d.depth != CV_8U || s.cn != 1 || s.w != d.w || s.h != d.h)
    SHIM_FAIL("only 8-bit single-channel CV_BILATERAL");
  /* cvSmooth(CV_BILATERAL): cv::bilateralFilter(src, dst, d = size1, sigmaColor = sigma1, sigmaSpace = sigma2, BORDER_REPLICATE) */
  orc_bilateral_u8(s.p, s.step, s.w, s.h, d.p, d.step, size1, sigma1, sigma2);
}

CV_IMPL double cvThreshold(const CvArr *srcarr, CvArr *dstarr, double threshold, double max_value, int threshold_type) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (threshold_type != CV_THRESH_TOZERO || s.depth != d.depth || s.cn != 1 || s.w != d.w || s.h != d.h) SHIM_FAIL("only TOZERO");
  for (int y = 0; y < s.h; y++) {
    const uchar *sr = s.p + (size_t)y * s.step;
    uchar *dr = d.p + (size_t)y * d.step;
    for (int x = 0; x < s.w; x++) {
      if (s.depth == CV_16S) {  /* cv::threshold on 16S: integer threshold cvFloor(thresh) */
        short v = ((const short *)sr)[x];
        ((short *)dr)[x] = v > (short)orc_cv_floor(threshold) ? v : 0;
      } else if (s.depth == CV_32F) {
        float v = ((const float *)sr)[x];
        ((float *)dr)[x] = v > (float)threshold ? v : 0.f;
      } else {
        uchar v = sr[x];
        dr[x] = v > (uchar)orc_cv_floor(threshold) ? v : 0;
      }
    }
  }
  return threshold;
}


CV_IMPL void cvSobel(const CvArr *srcarr, CvArr *dstarr, int dx, int dy, int aperture_size) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (s.depth != CV_8U || d.depth != CV_16S || s.cn != 1 || d.cn != 1 || s.w != d.w || s.h != d.h) SHIM_FAIL("types");
  if (aperture_size != 3 && aperture_size != 5 && aperture_size != 7) SHIM_FAIL("aperture (Scharr not shimmed)");
  orc_sobel_u8_s16(s.p, s.step, s.w, s.h, (int16_t *)d.p, d.step, dx, dy, aperture_size);
}

CV_IMPL IplConvKernel *cvCreateStructuringElementEx(int cols, int rows, int anchor_x, int anchor_y, int shape, int *values) {
  if (cols != 3 || rows != 3 || anchor_x != 1 || anchor_y != 1 || shape != CV_SHAPE_CROSS || values) SHIM_FAIL("only 3x3 cross");
  IplConvKernel *k = (IplConvKernel *)calloc(1, sizeof(IplConvKernel));
  k->nCols = cols;
  k->nRows = rows;
  k->anchorX = anchor_x;
  k->anchorY = anchor_y;
  return k;
}

CV_IMPL void cvReleaseStructuringElement(IplConvKernel **element) {
  if (element && *element) {
    free(*element);
    *element = NULL;
  }
}

CV_IMPL void cvMorphologyEx(const CvArr *srcarr, CvArr *dstarr, CvArr *temp, IplConvKernel *element, int operation, int iterations) {
  (void)temp;
  View s = view_of(srcarr), d = view_of(dstarr);
  if (operation != CV_MOP_GRADIENT || iterations != 1 || !element || s.depth != CV_8U || d.depth != CV_8U || s.cn != 1)
    SHIM_FAIL("only u8 cross gradient");
  orc_morph_grad_cross3_u8(s.p, s.step, s.w, s.h, d.p, d.step);
}

CV_IMPL void cvResize(const CvArr *srcarr, CvArr *dstarr, int interpolation) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (interpolation != CV_INTER_LINEAR || s.depth != CV_8U || d.depth != CV_8U || s.cn != 1 || s.h != d.h || s.w != 2 * d.w)
    SHIM_FAIL("only exact 2:1 horizontal u8 INTER_LINEAR");
  orc_resize_half_width_u8(s.p, s.step, s.w, s.h, d.p, d.step);
}

CV_IMPL void cvWarpPerspective(const CvArr *srcarr, CvArr *dstarr, const CvMat *M, int flags, CvScalar fillval) {
  View s = view_of(srcarr), d = view_of(dstarr);
  if (flags != (CV_INTER_LINEAR + CV_WARP_FILL_OUTLIERS) || fillval.val[0] != 0 || s.depth != CV_8U || d.depth != CV_8U ||
      s.cn != 1 || d.cn != 1 || CV_MAT_TYPE(M->type) != CV_32FC1 || M->rows != 3 || M->cols != 3)
    SHIM_FAIL("only u8 1-channel INTER_LINEAR+FILL_OUTLIERS with a 3x3 f32 matrix");
  float m[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) m[r * 3 + c] = ((const float *)(M->data.ptr + (size_t)r * M->step))[c];
  orc_warp_perspective_u8(s.p, s.step, s.w, s.h, d.p, d.step, d.w, d.h, m);
}

/* ---------------------------------------------------------------- C++ bits the headers reference */

namespace cv {
void error(const Exception &exc) {
  fprintf(stderr, "cv::error: %s (%s:%d)\n", exc.err.c_str(), exc.file.c_str(), exc.line);
  abort();
}
Exception::Exception() : code(0), line(0) {}
Exception::Exception(int _code, const string &_err, const string &_func, const string &_file, int _line)
    : code(_code), err(_err), func(_func), file(_file), line(_line) { formatMessage(); }
Exception::~Exception() throw() {}
const char *Exception::what() const throw() { return msg.c_str(); }
void Exception::formatMessage() { msg = file + ": " + func + ": " + err; }
void fastFree(void *ptr) { free(ptr); }
void Mat::deallocate() {}
Mat::Mat(const IplImage *, bool) : size(&rows) { SHIM_FAIL("cv::Mat(IplImage) (dmz_blur_card is out of scope)"); }
void medianBlur(InputArray, OutputArray, int) { SHIM_FAIL("medianBlur (out of scope)"); }
}  // namespace cv

/* cv::_InputArray / _OutputArray are polymorphic; rather than restate their whole vtables for the one
 * out-of-scope caller (dmz_blur_card, dmz.cpp:499-515) the two constructors it references are
 * provided under their mangled names and abort if ever reached. */
extern "C" void cvshim_inputarray_ctor(void *, const void *) asm("_ZN2cv11_InputArrayC1ERKNS_3MatE");
extern "C" void cvshim_inputarray_ctor(void *, const void *) { SHIM_FAIL("cv::_InputArray (out of scope)"); }
extern "C" void cvshim_outputarray_ctor(void *, void *) asm("_ZN2cv12_OutputArrayC1ERNS_3MatE");
extern "C" void cvshim_outputarray_ctor(void *, void *) { SHIM_FAIL("cv::_OutputArray (out of scope)"); }
