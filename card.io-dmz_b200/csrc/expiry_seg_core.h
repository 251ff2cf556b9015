// card.io-dmz_b200/csrc/expiry_seg_core.h -- best_expiry_seg's decision logic, written once as
// __host__ __device__ code over plain arrays: the CUDA kernels run it one thread per card, and the CPU-only unit tests
// compile the very same header with g++ to check it against the reference build without a GPU.
//
// What it restates (scan/expiry_seg.cpp):
//   stripe selection                  best_expiry_seg, expiry_seg.cpp:744-857
//   character rectangles of a stripe  find_character_groups_for_stripe, expiry_seg.cpp:379-703
//   grouping / white-space stripping  gather_into_groups, strip_group_white_space, expiry_seg.cpp:98-170
//   grid fitting                      regrid_group, expiry_seg.cpp:172-243
//   rectangle trimming                optimize_character_rects, expiry_seg.cpp:245-336
//   slash test                        is_slash / applym_730c4cbd, expiry_seg.cpp:29-54, modelm_730c4cbd.cpp:2431-2452
// Inputs are the |Scharr-dx| image (zero above the number row) and its per-row sums, produced by expiry_scharr_kernel.
//
// The reference orders candidates with std::sort, whose result for EQUAL keys is whatever libstdc++'s introsort does;
// std_sort_emul below reproduces that algorithm step for step (GCC's bits/stl_algo.h: median-of-three to first,
// unguarded partition, depth limit 2*lg(n) with heap-sort fallback, threshold 16, final insertion sort), so ties fall
// exactly as in the reference build.  All sums are integers; the few float expressions keep the reference's operand
// types and order (long -> float conversions, float accumulation in column order, double constants 0.8).
#ifndef B200_EXPIRY_SEG_CORE_H
#define B200_EXPIRY_SEG_CORE_H

#include <float.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define XHD __host__ __device__ __forceinline__
#define XHDN __host__ __device__
#else
#define XHD inline
#define XHDN inline
#endif

namespace xseg {

constexpr int kW = 428, kH = 270;         // kCreditCardTargetWidth / Height
constexpr int kNumberHeight = 27;         // dmz_constants.h:14
constexpr int kSmallW = 9, kSmallH = 15;  // kSmallCharacterWidth / Height, expiry_types.h:17-18
constexpr int kTrimW = 11, kTrimH = 16;   // kTrimmedCharacterImageWidth / Height, expiry_types.h:20-21
constexpr int kMaxStripes = 3;            // kNumberOfStripesToTry
constexpr int kMaxRects = 160;            // character rectangles alive at once in one stripe (428 / 9 = 47 before regridding)
constexpr int kMaxGroups = 48;

// Sums fit 32 bits with room to spare (the reference's own bounds, expiry_seg.cpp:58-68: a 9 x 17 rectangle of the
// |Scharr| image is at most 624 240, a 15-row stripe of 258-column line sums at most 15 789 600); keeping them in int
// halves the per-thread working set of the one-thread-per-card kernel.
struct CharRect {
  int top, left;
  int sum;
};

struct Group {  // GroupedRects without the std::vector: rects live in a per-stripe pool
  int top, left, width, height, character_width;
  long long sum;
  int first, count;  // pool slice [first, first + count)
};

struct ExpiryGroupOut {  // one accepted MM/YY candidate: five character rectangles, the middle one a slash
  int32_t top, left, width, height, character_width, pattern, n_rects;
  int32_t rect_top[5], rect_left[5];
};

// ---- libstdc++ std::sort, restated -------------------------------------------------------------------------------
template <typename T, typename Less>
XHDN void adjust_heap(T *first, long hole, long len, T value, Less less) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  long parent = (hole - 1) / 2;  // __push_heap
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename T, typename Less>
XHDN void heap_sort_all(T *first, T *last, Less less) {  // __partial_sort(first, last, last)
  const long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    while (true) {
      T v = first[parent];
      adjust_heap(first, parent, len, v, less);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    T v = *last;
    *last = *first;
    adjust_heap(first, 0L, (long)(last - first), v, less);
  }
}

template <typename T, typename Less>
XHDN void unguarded_linear_insert(T *last, Less less) {
  T val = *last;
  T *next = last - 1;
  while (less(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <typename T, typename Less>
XHDN void insertion_sort(T *first, T *last, Less less) {
  if (first == last) return;
  for (T *i = first + 1; i != last; ++i) {
    if (less(*i, *first)) {
      T val = *i;
      for (T *p = i; p != first; --p) *p = *(p - 1);
      *first = val;
    } else {
      unguarded_linear_insert(i, less);
    }
  }
}

template <typename T, typename Less>
XHDN void std_sort_emul(T *first, T *last, Less less) {
  if (first == last) return;
  long n = last - first;
  int depth = 0;
  for (long v = n; v > 1; v >>= 1) depth++;  // __lg(n)
  depth *= 2;
  // __introsort_loop with its recursion on the right part turned into an explicit stack
  struct Frame {
    T *first, *last;
    int depth;
  };
  Frame stack[64];
  int sp = 0;
  stack[sp++] = Frame{first, last, depth};
  while (sp > 0) {
    Frame f = stack[--sp];
    while (f.last - f.first > 16) {
      if (f.depth == 0) {
        heap_sort_all(f.first, f.last, less);
        break;
      }
      --f.depth;
      T *mid = f.first + (f.last - f.first) / 2;
      T *a = f.first + 1, *b = mid, *c = f.last - 1, *pick;
      if (less(*a, *b)) pick = less(*b, *c) ? b : (less(*a, *c) ? c : a);
      else pick = less(*a, *c) ? a : (less(*b, *c) ? c : b);
      T tmp = *f.first;
      *f.first = *pick;
      *pick = tmp;
      T *lo = f.first + 1, *hi = f.last;
      while (true) {
        while (less(*lo, *f.first)) ++lo;
        --hi;
        while (less(*f.first, *hi)) --hi;
        if (!(lo < hi)) break;
        tmp = *lo;
        *lo = *hi;
        *hi = tmp;
        ++lo;
      }
      // the reference recurses into [cut, last) first and then continues with [first, cut): the order in which the
      // two halves are processed does not change the result (they are disjoint), so the right half is deferred
      if (sp < 64) stack[sp++] = Frame{lo, f.last, f.depth};
      f.last = lo;
    }
  }
  if (n > 16) {
    insertion_sort(first, first + 16, less);
    for (T *i = first + 16; i != last; ++i) unguarded_linear_insert(i, less);
  } else {
    insertion_sort(first, last, less);
  }
}

// ---- helpers -----------------------------------------------------------------------------------------------------
struct SumDesc {
  template <typename T>
  XHD bool operator()(const T &a, const T &b) const { return a.sum > b.sum; }
};

struct StripeSum {
  int base_row;
  int sum;
};

XHD int imin(int a, int b) { return a < b ? a : b; }
XHD int imax(int a, int b) { return a > b ? a : b; }

// strip_group_white_space (expiry_seg.cpp:112-140), iterative
XHDN void strip_white_space(Group &g, CharRect *pool) {
  while (g.count > 5) {
    const int i = g.first + (g.count - 4) / 2;
    const long long avg = ((long long)pool[i].sum + pool[i + 1].sum + pool[i + 2].sum + pool[i + 3].sum) / 4;
    const long long threshold = (long long)((double)avg * 0.8);
    if (pool[g.first].sum < threshold) {
      g.first++, g.count--;
      g.left = pool[g.first].left;
    } else if (pool[g.first + g.count - 1].sum < threshold) {
      g.count--;
    } else {
      break;
    }
    g.width = pool[g.first + g.count - 1].left + g.character_width - g.left;
  }
}

// Slash MLP: 176 -> 80 tanh -> 2 softmax (modelm_730c4cbd.cpp:2431-2452).  W = hidden W [80][176], hidden b [80],
// logistic W [2][80], logistic b [2] (the blob card.io-dmz_b200/weights/modelm_730c4cbd.bin).
XHDN float slash_probability(const float *W, const int16_t *sob, int top, int left) {
  float x[kTrimW * kTrimH];
  for (int r = 0; r < kTrimH; r++)
    for (int c = 0; c < kTrimW; c++) x[r * kTrimW + c] = (float)sob[(top + r) * kW + left + c] * (1.0f / 255.0f);  // prepare_image_for_seg
  float z0 = W[14320], z1 = W[14321];
  for (int j = 0; j < 80; j++) {
    const float *w = W + j * 176;
    float a = 0.0f;
    for (int i = 0; i < 176; i++) a += w[i] * x[i];
    const float h = tanhf(a + W[14080 + j]);
    z0 += W[14160 + j] * h;
    z1 += W[14160 + 80 + j] * h;
  }
  const float e0 = expf(z0), e1 = expf(z1);
  return e0 / (e0 + e1);
}

// regrid_group (expiry_seg.cpp:172-243).  Rewrites the group's rectangles into pool[*pool_n ...].
XHDN void regrid(const int16_t *sob, Group &g, CharRect *pool, int *pool_n) {
  const int bl = imax(g.left - 2 * kSmallW, 0), br = imin(g.left + g.width + 2 * kSmallW, kW), bw = br - bl;
  const int min_lines = (int)floorf((float)bw / 11.0f);
  int col_sums[kW];
  long long group_sum = 0;
  for (int c = 0; c < bw; c++) {
    int s = 0;
    for (int r = g.top; r < g.top + g.height; r++) s += sob[r * kW + bl + c];
    col_sums[c] = s;
    group_sum += s;
  }
  int best_spacing = 0, best_start = 0;
  float best_ratio = FLT_MAX;
  for (int spacing = 11; spacing <= 15; spacing++) {
    for (int start = 0; start < spacing; start++) {
      float line_sum = 0.0f;
      int lines = 0;
      for (int o = start; o < bw; o += spacing) {
        lines++;
        line_sum += (float)col_sums[o];
      }
      const float avg = line_sum / (float)lines;
      line_sum = avg * (float)min_lines;
      const float ratio = line_sum / ((float)group_sum - line_sum);
      if (ratio < best_ratio) best_ratio = ratio, best_spacing = spacing, best_start = start;
    }
  }
  const int first = *pool_n;
  int n = first;
  for (int o = best_start; o + 1 < bw && n < kMaxRects; o += best_spacing) {
    int s = 0;
    for (int c = o + 1; c < imin(o + best_spacing, bw); c++) s += col_sums[c];
    pool[n].top = g.top, pool[n].left = bl + o + 1, pool[n].sum = s;
    n++;
  }
  *pool_n = n;
  g.first = first, g.count = n - first;
  g.character_width = best_spacing - 1;
  g.left = pool[first].left;
  g.width = pool[n - 1].left + g.character_width - g.left;
  strip_white_space(g, pool);
}

// optimize_character_rects (expiry_seg.cpp:245-336): trims every rectangle to 11 x 16 around its brightest part.
XHDN void optimize_rects(const int16_t *sob, Group &g, CharRect *pool) {
  const int ciw = g.character_width + 4, cih = g.height + 4;
  int kept = 0;  // survivors are compacted to the front of the slice, in order
  for (int k = 0; k < g.count; k++) {
    CharRect rc = pool[g.first + k];
    const int rl = rc.left - 2, rt = g.top - 2;
    if (rl < 0 || rl + ciw > kW || rt + cih > kH) continue;  // erased
    // cvNormalize(CV_C, 255) then cvThreshold(TOZERO, 100) on the s16 window
    int vmax = 0;
    for (int r = 0; r < cih; r++)
      for (int c = 0; c < ciw; c++) {
        const int v = sob[(rt + r) * kW + rl + c];
        vmax = imax(vmax, v < 0 ? -v : v);
      }
    const double scale = (double)vmax > DBL_EPSILON ? 255.0 / (double)vmax : 0.0;
    const float fs = (float)scale;
    int colsum[32], rowsum[32];
    short win[32][24];
    for (int c = 0; c < ciw; c++) colsum[c] = 0;
    for (int r = 0; r < cih; r++)
      for (int c = 0; c < ciw; c++) {
        const float p = (float)sob[(rt + r) * kW + rl + c] * fs;
        int q = (int)rintf(p);  // cvRound of a float product: round-half-even
        q = q > 32767 ? 32767 : (q < -32768 ? -32768 : q);
        q = q > 100 ? q : 0;
        win[r][c] = (short)q;
        colsum[c] += q;
      }
    int lc = 0, rcol = ciw - 1, w = ciw;
    while (w > kTrimW) {
      if (colsum[lc] <= colsum[rcol]) lc++;
      else rcol--;
      w--;
    }
    for (int r = 0; r < cih; r++) {
      int s = 0;
      for (int c = lc; c <= rcol; c++) s += win[r][c];
      rowsum[r] = s;
    }
    int tr = 0, brow = cih - 1, h = cih;
    while (h > kTrimH) {
      if (rowsum[tr] <= rowsum[brow]) tr++;
      else brow--;
      h--;
    }
    rc.left = rl + lc, rc.top = rt + tr;
    pool[g.first + kept++] = rc;
  }
  g.count = kept;
  if (kept > 0) {
    int hi = kH, lo = 0;
    for (int k = 0; k < kept; k++) hi = imin(hi, pool[g.first + k].top), lo = imax(lo, pool[g.first + k].top);
    g.character_width = kTrimW;
    g.left = pool[g.first].left;
    g.width = pool[g.first + kept - 1].left + kTrimW - g.left;
    g.top = hi;
    g.height = lo + kTrimH - g.top;
  }
}

// find_character_groups_for_stripe (expiry_seg.cpp:379-703).  Appends accepted groups to out[*n_out ...].
// colsum (optional): the 428 column sums of rows base_row .. base_row + exp_h - 1 (stripe_colsums below; the CUDA path
// computes them with one warp per stripe) -- the sliding rectangle sums then cost two loads per column instead of 2 exp_h.
XHDN void stripe_groups(const int16_t *sob, const float *slash_w, int base_row, int stripe_sum, ExpiryGroupOut *out,
                        int *n_out, int max_out, int *overflow, const int32_t *colsum = nullptr) {
  const int exp_top = base_row - 1, exp_h = imin(kSmallH + 2, kH - exp_top);  // == stripe_rows(base_row)
  const long long rect_average = ((long long)stripe_sum * kSmallW) / kW;
  const float too_dim = (float)(rect_average / 5);
  // [1] sliding 9-wide rectangle sums.  NB the reference sums rows base_row .. base_row + exp_h - 1 here (not the
  // expanded stripe's own rows exp_top ..): expiry_seg.cpp:403-405, 426-429
  CharRect cand[kW - kSmallW + 1];
  int n_cand = 0;
  float total = 0.0f;
  int rs = 0;
  if (colsum != nullptr) {
    for (int c = 0; c < kSmallW; c++) rs += colsum[c];
  } else {
    for (int c = 0; c < kSmallW; c++)
      for (int r = 0; r < exp_h; r++) rs += sob[(base_row + r) * kW + c];
  }
  for (int c = 0; c < kW - kSmallW + 1; c++) {
    if ((float)rs > too_dim) {
      cand[n_cand].top = exp_top, cand[n_cand].left = c, cand[n_cand].sum = rs;
      n_cand++;
      total += (float)rs;
    }
    if (c < kW - kSmallW) {
      if (colsum != nullptr) {
        rs += colsum[c + kSmallW] - colsum[c];  // (integer sums: the same value in any order)
      } else {
        for (int r = 0; r < exp_h; r++) rs += sob[(base_row + r) * kW + c + kSmallW] - sob[(base_row + r) * kW + c];
      }
    }
  }
  if (n_cand == 0) return;
  const float average = total / (float)n_cand;
  const float keep_above = (float)(0.8 * (double)average);
  // [2] brightest first (std::sort order, ties included)   [3] greedy non-overlapping pick
  std_sort_emul(cand, cand + n_cand, SumDesc());
  uint32_t taken[(kW + 31) / 32];  // the reference's non_overlapping_rect_mask, one bit per column
  for (int i = 0; i < (kW + 31) / 32; i++) taken[i] = 0u;
  CharRect pool[kMaxRects];
  int pool_n = 0;
  for (int i = 0; i < n_cand; i++) {
    if ((float)cand[i].sum <= keep_above) break;
    const int l = cand[i].left;
    const int e = l + kSmallW - 1;
    if (!((taken[l >> 5] >> (l & 31)) & 1u) && !((taken[e >> 5] >> (e & 31)) & 1u) && pool_n < kMaxRects) {
      pool[pool_n++] = cand[i];
      for (int k = l; k <= e; k++) taken[k >> 5] |= 1u << (k & 31);
    }
  }
  // [4] local groups: runs of rectangles whose gaps are below one character width (gather_into_groups, tolerance 9).
  // Sorting by left edge needs no tie rule (lefts are distinct); an insertion sort gives the same order.
  for (int i = 1; i < pool_n; i++) {
    CharRect v = pool[i];
    int j = i - 1;
    while (j >= 0 && pool[j].left > v.left) pool[j + 1] = pool[j], j--;
    pool[j + 1] = v;
  }
  Group groups[kMaxGroups];
  int n_groups = 0;
  for (int i = 0; i < pool_n && n_groups < kMaxGroups;) {
    Group g;
    g.top = pool[i].top, g.left = pool[i].left, g.width = kSmallW, g.height = exp_h, g.character_width = kSmallW;
    g.sum = pool[i].sum, g.first = i, g.count = 1;
    int j = i + 1;
    while (j < pool_n && pool[j].left - (g.left + g.width) < kSmallW) {
      g.width = pool[j].left + kSmallW - g.left;
      g.sum += pool[j].sum;
      g.count++;
      j++;
    }
    i = j;
    strip_white_space(g, pool);
    if (g.count >= 4) groups[n_groups++] = g;  // kMinimumExpiryStripCharacters - 1: regridding may recover a character
  }
  // regrid, trim, filter
  for (int k = 0; k < n_groups; k++) regrid(sob, groups[k], pool, &pool_n);
  for (int k = n_groups - 1; k >= 0; k--) optimize_rects(sob, groups[k], pool);
  // slash in a plausible position -> MM/YY candidates
  for (int k = 0; k < n_groups; k++) {
    const Group &g = groups[k];
    if (g.count < 5) continue;
    for (int f = 0; f + 4 < g.count; f++) {
      const CharRect &mid = pool[g.first + f + 2];
      if (!(slash_probability(slash_w, sob, mid.top, mid.left) > 0.7f)) continue;
      if (*n_out >= max_out) {
        (*overflow)++;
        continue;
      }
      ExpiryGroupOut &o = out[(*n_out)++];
      int top = pool[g.first + f].top, height = kSmallH, width = kSmallW;
      const int left = pool[g.first + f].left;
      for (int c = 0; c < 5; c++) {
        const CharRect &cr = pool[g.first + f + c];
        const int bottom = top + height;
        top = imin(cr.top, top);
        width = cr.left + kSmallW - left;
        height = imax(cr.top + kSmallH, bottom) - top;
        o.rect_top[c] = cr.top, o.rect_left[c] = cr.left;
      }
      o.top = top, o.left = left, o.width = width, o.height = height;
      o.character_width = kTrimW, o.pattern = 0 /* ExpiryPatternMMsYY */, o.n_rects = 5;
    }
  }
}

// best_expiry_seg's stripe selection (expiry_seg.cpp:744-857).  line_sum[r] = sum of sob[r][27 .. 284] for
// r >= y_offset + 27 (rows above are never read).  Returns the number of stripes picked (<= kMaxStripes).
XHDN int pick_stripes(const int32_t *line_sum, int y_offset, StripeSum *picked) {
  const int first_base = y_offset + kNumberHeight + 1, last_base = kH - (kSmallH + 1);
  StripeSum stripes[kH];
  int n_stripes = 0;
  for (int base = first_base; base < last_base; base++) {
    int sum = 0, peak = 0;
    for (int r = base; r < base + kSmallH; r++) {
      sum += line_sum[r];
      if (line_sum[r] > peak) peak = line_sum[r];
    }
    const int threshold = peak / 2;
    if (line_sum[base] + line_sum[base + 1] < threshold) continue;
    if (line_sum[base + kSmallH - 2] + line_sum[base + kSmallH - 1] < threshold) continue;
    bool good = true;
    for (int r = base; r < base + kSmallH - 3; r++)
      if (line_sum[r + 1] < threshold && line_sum[r + 2] < threshold) {
        good = false;
        break;
      }
    if (good) stripes[n_stripes].base_row = base, stripes[n_stripes].sum = sum, n_stripes++;
  }
  std_sort_emul(stripes, stripes + n_stripes, SumDesc());
  int n_picked = 0;
  for (int i = 0; i < n_stripes && n_picked < kMaxStripes; i++) {
    bool overlap = false;
    for (int p = 0; p < n_picked; p++)
      if (picked[p].base_row - kSmallH < stripes[i].base_row && stripes[i].base_row < picked[p].base_row + kSmallH) overlap = true;
    if (!overlap) picked[n_picked++] = stripes[i];
  }
  return n_picked;
}

// rows a stripe's rectangle sums cover: base_row .. base_row + stripe_rows(base_row) - 1 (find_character_groups_for_stripe's
// expanded height, expiry_seg.cpp:403-405)
XHD int stripe_rows(int base_row) { return imin(kSmallH + 2, kH - (base_row - 1)); }

// the 428 column sums of a stripe (host form; the CUDA path computes them with one warp per stripe)
XHDN void stripe_colsums(const int16_t *sob, int base_row, int32_t *colsum) {
  const int rows = stripe_rows(base_row);
  for (int c = 0; c < kW; c++) {
    int s = 0;
    for (int r = 0; r < rows; r++) s += sob[(base_row + r) * kW + c];
    colsum[c] = s;
  }
}

// best_expiry_seg's per-stripe search over the picked stripes (expiry_seg.cpp:858-903).  colsums (optional):
// kMaxStripes x kW column sums, one row per picked stripe.  Returns the number of groups written.
XHDN int search_stripes(const int16_t *sob, const StripeSum *picked, int n_picked, const float *slash_w, ExpiryGroupOut *out, int max_out,
                        int *overflow, const int32_t *colsums = nullptr) {
  int n_out = 0;
  *overflow = 0;
  for (int p = 0; p < n_picked; p++)
    stripe_groups(sob, slash_w, picked[p].base_row, picked[p].sum, out, &n_out, max_out, overflow, colsums ? colsums + p * kW : nullptr);
  return n_out;
}

// stripe selection + per-stripe search in one call (the CPU unit tests; use_colsums exercises the column-sum form the
// CUDA path takes)
XHDN int best_expiry_groups(const int16_t *sob, const int32_t *line_sum, int y_offset, const float *slash_w, ExpiryGroupOut *out,
                            int max_out, int *overflow, int32_t *colsum_scratch = nullptr /* kMaxStripes * kW, or null */) {
  StripeSum picked[kMaxStripes];
  const int n_picked = pick_stripes(line_sum, y_offset, picked);
  if (colsum_scratch != nullptr)
    for (int p = 0; p < n_picked; p++) stripe_colsums(sob, picked[p].base_row, colsum_scratch + p * kW);
  return search_stripes(sob, picked, n_picked, slash_w, out, max_out, overflow, colsum_scratch);
}

// llcv_scharr3_dx_abs on the rows below the number (cv/sobel.cpp:706-799) for ONE output pixel: |right - left| per
// row (columns clamped at the image), then 3 / 10 / 3 down the column (rows clamped at the ROI [y0, kH)).
XHD int scharr_abs_at(const uint8_t *card, int y0, int x, int y) {
  const int xl = x == 0 ? 0 : x - 1, xr = x == kW - 1 ? kW - 1 : x + 1;
  const int yu = y == y0 ? y0 : y - 1, yd = y == kH - 1 ? kH - 1 : y + 1;
  const int a = card[yu * kW + xr] - card[yu * kW + xl], b = card[y * kW + xr] - card[y * kW + xl], c = card[yd * kW + xr] - card[yd * kW + xl];
  return 3 * ((a < 0 ? -a : a) + (c < 0 ? -c : c)) + 10 * (b < 0 ? -b : b);
}

}  // namespace xseg
#endif
