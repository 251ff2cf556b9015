// tests/layout_probe.cpp -- prints sizeof / alignof of every type and offsetof + size of every member that crosses the
// drop-in boundary (dmz.h:22-37, dmz_olm.h:29-42, scan/frame.h:14-28, scan/scan.h:19-48, scan/n_vseg.h:14-21,
// scan/n_hseg.h:13-19, scan/expiry_types.h:20-118, opencv2/core/types_c.h IplImage).
//
// Compiled twice from this one source:
//   -DPROBE_REFERENCE_HEADERS  against the reference's own, unmodified dmz.h + scan/scan.h (oracle/Makefile, target
//                              `layout`: output committed as tests/golden/ref_layout.txt, regenerated and compared
//                              whenever /root/reference is present)
//   (default)                  against include/dmz_b200_compat.h
// tests/test_abi.py requires the two outputs to be identical line for line.
#include <stddef.h>
#include <stdio.h>

#ifdef PROBE_REFERENCE_HEADERS
#include "dmz.h"
#include "scan/scan.h"
#else
#include "dmz_b200_compat.h"
#endif

#define TYPE(T) printf("%s sizeof %zu alignof %zu\n", #T, sizeof(T), (size_t)__alignof__(T))
#define MEMBER(T, m) printf("%s.%s offset %zu size %zu\n", #T, #m, (size_t)__builtin_offsetof(T, m), sizeof(((T *)0)->m))

int main() {
  TYPE(IplImage);
  MEMBER(IplImage, nSize); MEMBER(IplImage, nChannels); MEMBER(IplImage, depth); MEMBER(IplImage, dataOrder);
  MEMBER(IplImage, origin); MEMBER(IplImage, align); MEMBER(IplImage, width); MEMBER(IplImage, height);
  MEMBER(IplImage, roi); MEMBER(IplImage, imageSize); MEMBER(IplImage, imageData); MEMBER(IplImage, widthStep);
  MEMBER(IplImage, imageDataOrigin);
  TYPE(IplROI);
  MEMBER(IplROI, coi); MEMBER(IplROI, xOffset); MEMBER(IplROI, yOffset); MEMBER(IplROI, width); MEMBER(IplROI, height);

  TYPE(FrameOrientation);
  TYPE(dmz_point); MEMBER(dmz_point, x); MEMBER(dmz_point, y);
  TYPE(dmz_corner_points);
  MEMBER(dmz_corner_points, top_left); MEMBER(dmz_corner_points, bottom_left); MEMBER(dmz_corner_points, top_right);
  MEMBER(dmz_corner_points, bottom_right);
  TYPE(dmz_context); MEMBER(dmz_context, mz);
  TYPE(ParametricLine); MEMBER(ParametricLine, rho); MEMBER(ParametricLine, theta);
  TYPE(dmz_found_edge); MEMBER(dmz_found_edge, found); MEMBER(dmz_found_edge, location);
  TYPE(dmz_edges); MEMBER(dmz_edges, top); MEMBER(dmz_edges, left); MEMBER(dmz_edges, bottom); MEMBER(dmz_edges, right);

  TYPE(NVerticalSegmentation);
  MEMBER(NVerticalSegmentation, score); MEMBER(NVerticalSegmentation, y_offset); MEMBER(NVerticalSegmentation, pattern_type);
  MEMBER(NVerticalSegmentation, number_pattern); MEMBER(NVerticalSegmentation, number_pattern_length);
  MEMBER(NVerticalSegmentation, number_length);
  TYPE(NHorizontalSegmentation);
  MEMBER(NHorizontalSegmentation, n_offsets); MEMBER(NHorizontalSegmentation, offsets); MEMBER(NHorizontalSegmentation, score);
  MEMBER(NHorizontalSegmentation, number_width); MEMBER(NHorizontalSegmentation, pattern_offset);
  TYPE(NumberScores);
  TYPE(NumberPredictions);

  TYPE(CharacterRect); MEMBER(CharacterRect, top); MEMBER(CharacterRect, left); MEMBER(CharacterRect, sum);
  TYPE(GroupedRects);
  MEMBER(GroupedRects, top); MEMBER(GroupedRects, left); MEMBER(GroupedRects, width); MEMBER(GroupedRects, height);
  MEMBER(GroupedRects, grouped_yet); MEMBER(GroupedRects, sum); MEMBER(GroupedRects, character_width);
  MEMBER(GroupedRects, character_rects); MEMBER(GroupedRects, pattern); MEMBER(GroupedRects, scores);
  MEMBER(GroupedRects, recently_seen_count); MEMBER(GroupedRects, total_seen_count);
  TYPE(GroupedRectsList);

  TYPE(FrameScanResult);
  MEMBER(FrameScanResult, focus_score); MEMBER(FrameScanResult, scores); MEMBER(FrameScanResult, hseg); MEMBER(FrameScanResult, vseg);
  MEMBER(FrameScanResult, expiry_groups); MEMBER(FrameScanResult, name_groups); MEMBER(FrameScanResult, usable);
  MEMBER(FrameScanResult, upside_down); MEMBER(FrameScanResult, flipped); MEMBER(FrameScanResult, brightness_score);
  MEMBER(FrameScanResult, iso_speed); MEMBER(FrameScanResult, shutter_speed); MEMBER(FrameScanResult, torch_is_on);

  TYPE(ScanFrameAnalytics); MEMBER(ScanFrameAnalytics, frame_index); MEMBER(ScanFrameAnalytics, frame_values);
  TYPE(ScanSessionAnalytics);
  MEMBER(ScanSessionAnalytics, num_frames_scanned); MEMBER(ScanSessionAnalytics, frames_ring_start);
  MEMBER(ScanSessionAnalytics, frames_ring);

  TYPE(ScannerResult);
  MEMBER(ScannerResult, complete); MEMBER(ScannerResult, predictions); MEMBER(ScannerResult, hseg); MEMBER(ScannerResult, vseg);
  MEMBER(ScannerResult, n_numbers); MEMBER(ScannerResult, expiry_month); MEMBER(ScannerResult, expiry_year);

  TYPE(ScannerState);
  MEMBER(ScannerState, count15); MEMBER(ScannerState, count16); MEMBER(ScannerState, aggregated15); MEMBER(ScannerState, aggregated16);
  MEMBER(ScannerState, session_analytics); MEMBER(ScannerState, successfulCardNumberResult);
  MEMBER(ScannerState, mostRecentUsableHSeg); MEMBER(ScannerState, mostRecentUsableVSeg);
  MEMBER(ScannerState, timeOfCardNumberCompletionInMilliseconds); MEMBER(ScannerState, scan_expiry);
  MEMBER(ScannerState, expiry_month); MEMBER(ScannerState, expiry_year); MEMBER(ScannerState, expiry_groups);
  MEMBER(ScannerState, name_groups);

  TYPE(CythonCharacterRect); MEMBER(CythonCharacterRect, top); MEMBER(CythonCharacterRect, left);
  TYPE(CythonGroupedRects);
  MEMBER(CythonGroupedRects, top); MEMBER(CythonGroupedRects, left); MEMBER(CythonGroupedRects, width); MEMBER(CythonGroupedRects, height);
  MEMBER(CythonGroupedRects, character_width); MEMBER(CythonGroupedRects, pattern); MEMBER(CythonGroupedRects, scores);
  MEMBER(CythonGroupedRects, recently_seen_count); MEMBER(CythonGroupedRects, total_seen_count);
  MEMBER(CythonGroupedRects, number_of_character_rects); MEMBER(CythonGroupedRects, character_rects);
  return 0;
}
