// card.io-dmz_b200/csrc/expiry_session.h -- the cross-frame half of expiry_extract (scan/expiry_categorize.cpp), host-only,
// shared by the C-ABI session object (scanner.cpp) and the C++ drop-in layer (dmz_compat.cpp, which keeps the state in the
// reference's own ScannerState::expiry_groups).
#ifndef B200_EXPIRY_SESSION_H
#define B200_EXPIRY_SESSION_H

#include <stdint.h>
#include <string.h>

namespace {

// Eigen's completely unrolled non-vectorised redux: balanced binary tree (Core/Redux.h:96-118); used for
// aggregated.row(i).sum() on the 1x10 row block.
float tree_sum(const float *v, int start, int len) {
  if (len == 1) return v[start];
  return tree_sum(v, start, len / 2) + tree_sum(v, start + len / 2, len - len / 2);
}

struct ExpiryAgg {  // what survives of a GroupedRects across frames (expiry_types.h:68-84)
  int top, left, n_rects, recently_seen, total_seen;
  float scores[5][10];  // rows = character index; row 2 (the slash) is never scored
  int tag;              // caller's handle (dmz_compat.cpp: index of the GroupedRects this entry came from)
};
constexpr int kMaxExpiryAgg = 64;  // capacity of the C-ABI session (b200_scanner); the C++ drop-in layer sizes its lists per frame

bool same_place(int top_a, int left_a, int n_a, int top_b, int left_b, int n_b) {
  // GROUPED_RECTS_VERTICAL_ALLOWANCE = 16 / 2, GROUPED_RECTS_HORIZONTAL_ALLOWANCE = 11 / 2 (expiry_categorize.cpp:24-25)
  const int dt = top_a - top_b, dl = left_a - left_b;
  return !((dt < 0 ? -dt : dt) > 8 || (dl < 0 ? -dl : dl) > 5 || n_a != n_b);
}

// expiry_string_to_expiry_month_and_year for ExpiryPatternMMsYY (expiry_categorize.cpp:334-395); ' ' marks an unstable digit
void month_year_from_string(const char *e, int current_year, int current_month, bool allow_past, int *expiry_month, int *expiry_year) {
  int month = -1, year = -1;
  if (e[0] != ' ' && e[1] != ' ' && e[3] != ' ' && e[4] != ' ') {
    month = (uint8_t)(e[0] - '0') * 10 + (uint8_t)(e[1] - '0');
    year = (uint8_t)(e[3] - '0') * 10 + (uint8_t)(e[4] - '0');
  }
  if (month > 12 && year > 0 && year <= 12) {  // YY/MM cards
    const int t = month;
    month = year;
    year = t;
  }
  int full_year = year + 2000;
  if (month > 0 && month <= 12 && (full_year > *expiry_year || (full_year == *expiry_year && month > *expiry_month))) {
    if (full_year < current_year + 5 && (full_year > current_year || (full_year == current_year && month >= current_month))) {
      *expiry_month = month, *expiry_year = full_year;
    } else if (allow_past) {
      if (year > 60) full_year = year + 1900;
      if (full_year < current_year + 5) *expiry_month = month, *expiry_year = full_year;
    }
  }
}

// get_stable_expiry_month_and_year (expiry_categorize.cpp:398-441): a digit counts only if it holds >= 0.7 of its row
void stable_month_year(const float (*scores)[10], int n_chars, int current_year, int current_month, bool allow_past, int *month, int *year) {
  char e[8] = {0};
  for (int i = 0; i < n_chars && i < 5; i++) {
    const float *row = scores[i];
    float mx = row[0];
    int arg = 0;
    for (int j = 1; j < 10; j++)
      if (row[j] > mx) mx = row[j], arg = j;
    const float stability = mx / tree_sum(row, 0, 10);
    e[i] = stability < 0.7f ? ' ' : (char)('0' + arg);  // kExpiryMinStability
  }
  month_year_from_string(e, current_year, current_month, allow_past, month, year);
}

// expiry_aggregate_grouped_rects (expiry_categorize.cpp:258-330): merges the frame's groups `fresh` into the session's `agg`.
inline void expiry_aggregate(ExpiryAgg *agg, int *n_agg, ExpiryAgg *fresh, int n_fresh, int agg_capacity = kMaxExpiryAgg) {
  // (a) equivalent groups inside the new list: running mean into the earlier one
  for (int i = 0; i < n_fresh; i++) {
    const int top1 = fresh[i].top, left1 = fresh[i].left, n1 = fresh[i].n_rects;
    float so_far = 1;
    for (int j = n_fresh - 1; j > i; j--) {
      if (!same_place(fresh[j].top, fresh[j].left, fresh[j].n_rects, top1, left1, n1)) continue;
      for (int r = 0; r < 5; r++)
        for (int d = 0; d < 10; d++) fresh[i].scores[r][d] = ((fresh[i].scores[r][d] * so_far) + fresh[j].scores[r][d]) / (so_far + 1);
      so_far++;
      for (int k = j; k + 1 < n_fresh; k++) fresh[k] = fresh[k + 1];
      n_fresh--;
    }
  }
  // (b) new groups that sit where an old one sat: decay-blend into it
  const float decay = 0.7f, gain = 1 - 0.7f;  // kExpiryDecayFactor
  for (int o = 0; o < (*n_agg); o++) {
    ExpiryAgg &old = agg[o];
    const int old_top = old.top, old_left = old.left, old_n = old.n_rects;
    for (int j = n_fresh - 1; j >= 0; j--) {
      if (!same_place(fresh[j].top, fresh[j].left, fresh[j].n_rects, old_top, old_left, old_n)) continue;
      old.recently_seen++;
      old.total_seen++;
      for (int r = 0; r < 5; r++)
        for (int d = 0; d < 10; d++) old.scores[r][d] = (old.scores[r][d] * decay) + (fresh[j].scores[r][d] * gain);
      old.top = fresh[j].top, old.left = fresh[j].left;
      for (int k = j; k + 1 < n_fresh; k++) fresh[k] = fresh[k + 1];
      n_fresh--;
    }
  }
  // (c) forget groups that have not been seen for a while   (d) adopt the rest as new
  for (int o = (*n_agg) - 1; o >= 0; o--) {
    if (--agg[o].recently_seen <= 0) {
      for (int k = o; k + 1 < (*n_agg); k++) agg[k] = agg[k + 1];
      (*n_agg)--;
    }
  }
  for (int j = 0; j < n_fresh && (*n_agg) < agg_capacity; j++) {
    fresh[j].recently_seen = 3, fresh[j].total_seen = 1;
    agg[(*n_agg)++] = fresh[j];
  }
}

}  // namespace
#endif
