// tests/expiry_host.cpp -- CPU-side unit-test harness for card.io-dmz_b200/csrc/expiry_seg_core.h (the decision logic
// the CUDA kernels run one thread per card).  Built by tests/test_expiry_seg.py with g++; never part of the product.
#include <algorithm>
#include <cstring>
#include <vector>

#include "expiry_seg_core.h"

extern "C" {

// the whole of best_expiry_seg on one 428x270 card, Scharr image and row sums computed here on the host
int xh_best_expiry_seg(const uint8_t *card, int y_offset, const float *slash_w, xseg::ExpiryGroupOut *out, int max_out, int *overflow) {
  std::vector<int16_t> sob((size_t)xseg::kW * xseg::kH, 0);
  std::vector<int32_t> line_sum(xseg::kH, 0);
  const int y0 = y_offset + xseg::kNumberHeight;
  for (int y = y0; y < xseg::kH; y++) {
    int32_t s = 0;
    for (int x = 0; x < xseg::kW; x++) {
      const int v = xseg::scharr_abs_at(card, y0, x, y);
      sob[(size_t)y * xseg::kW + x] = (int16_t)v;
      if (x >= 27 && x < 285) s += v;
    }
    line_sum[y] = s;
  }
  // the column-sum form of the search (what the CUDA path takes) and the direct form must agree on every card
  std::vector<int32_t> colsums((size_t)xseg::kMaxStripes * xseg::kW);
  std::vector<xseg::ExpiryGroupOut> direct((size_t)max_out);
  int overflow_direct = 0;
  const int n_direct = xseg::best_expiry_groups(sob.data(), line_sum.data(), y_offset, slash_w, direct.data(), max_out, &overflow_direct);
  const int n = xseg::best_expiry_groups(sob.data(), line_sum.data(), y_offset, slash_w, out, max_out, overflow, colsums.data());
  if (n != n_direct || *overflow != overflow_direct || memcmp(direct.data(), out, sizeof(xseg::ExpiryGroupOut) * (size_t)(n < max_out ? n : max_out)) != 0) return -1000;
  return n;
}

void xh_scharr(const uint8_t *card, int y_offset, int16_t *out) {
  const int y0 = y_offset + xseg::kNumberHeight;
  memset(out, 0, sizeof(int16_t) * xseg::kW * xseg::kH);
  for (int y = y0; y < xseg::kH; y++)
    for (int x = 0; x < xseg::kW; x++) out[(size_t)y * xseg::kW + x] = (int16_t)xseg::scharr_abs_at(card, y0, x, y);
}

float xh_slash(const float *slash_w, const int16_t *sob, int top, int left) { return xseg::slash_probability(slash_w, sob, top, left); }

// std_sort_emul against the real std::sort on (key, id) records with many equal keys; returns the number of mismatching slots
struct Rec {
  long long sum;
  int id;
};
struct RecDesc {
  bool operator()(const Rec &a, const Rec &b) const { return a.sum > b.sum; }
};
int xh_sort_check(const long long *keys, int n) {
  std::vector<Rec> a(n), b(n);
  for (int i = 0; i < n; i++) a[i].sum = b[i].sum = keys[i], a[i].id = b[i].id = i;
  std::sort(a.begin(), a.end(), RecDesc());
  xseg::std_sort_emul(b.data(), b.data() + n, xseg::SumDesc());
  int bad = 0;
  for (int i = 0; i < n; i++) bad += a[i].id != b[i].id;
  return bad;
}
}
