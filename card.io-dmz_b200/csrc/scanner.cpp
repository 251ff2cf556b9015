// card.io-dmz_b200/csrc/scanner.cpp -- scanner session bookkeeping (scan/scan.cpp:19-194) on the host.
// Pure integer / scalar-float logic that couples the frames of one session sequentially (the 0.8 EMA), so
// it stays on the CPU exactly as in the reference; the per-frame arithmetic it consumes (b200_scan) comes
// from the GPU.  Luhn / prefix table: dmz_olm.cpp (dmz_passes_luhn_checksum,
// dmz_card_info_for_prefix_and_length).  Compiled without fp contraction.
#include <string.h>

#include "b200_dmz.h"
#include "expiry_session.h"

struct b200_scanner {
  uint16_t count15, count16;
  float agg15[160], agg16[160];
  int complete;
  uint8_t digits[16];
  int n_numbers;
  int expiry_month, expiry_year, n_expiry;
  ExpiryAgg expiry[kMaxExpiryAgg];
};

namespace {

bool luhn(const uint8_t *d, int n) {
  int even = 0, sum = 0;
  for (int i = n - 1; i >= 0; i--) {
    int addend = d[i] * (1 << (even++ & 1));
    sum += addend % 10 + addend / 10;
  }
  return sum % 10 == 0;
}

enum { kUnrecognized = 0, kAmbiguous = 1 };

int card_type(const uint8_t *d, int n) {
  static const struct {
    int type, len, plen;
    long lo, hi;
  } t[] = {{5, 16, 4, 2221, 2720}, {6, 14, 3, 300, 305},   {6, 14, 3, 309, 309}, {2, 15, 2, 34, 34}, {3, 16, 4, 3528, 3589},
           {6, 14, 2, 36, 36},     {6, 14, 2, 38, 39},     {2, 15, 2, 37, 37},   {4, 16, 1, 4, 4},   {7, 16, 2, 50, 50},
           {5, 16, 2, 51, 55},     {7, 16, 2, 56, 59},     {6, 16, 4, 6011, 6011}, {7, 16, 2, 61, 61}, {6, 16, 2, 62, 62},
           {7, 16, 2, 63, 63},     {6, 16, 3, 644, 649},   {6, 16, 2, 65, 65},   {7, 16, 2, 66, 69}, {6, 16, 2, 88, 88}};
  int compatible = 0, type = kUnrecognized;
  if (n <= 0) return kUnrecognized;
  for (size_t i = 0; i < sizeof(t) / sizeof(t[0]); i++) {
    if (n != t[i].len) continue;
    int plen = t[i].plen, factor = 1;
    while (plen > n) factor *= 10, plen--;
    long prefix = 0;
    for (int j = 0; j < plen; j++) prefix = prefix * 10 + d[j];
    if (prefix >= t[i].lo / factor && prefix <= t[i].hi / factor) compatible++, type = t[i].type;
  }
  return compatible == 1 ? type : (compatible > 1 ? kAmbiguous : kUnrecognized);
}

}  // namespace

extern "C" {

/* dmz_card_info_for_prefix_and_length(number, n, false).card_type; 0 = unrecognized, 1 = ambiguous (dmz_olm.h:44-53) */
int b200_card_type_for_number(const uint8_t *digits, int n) { return card_type(digits, n); }

b200_scanner *b200_scanner_new(void) {
  b200_scanner *s = new b200_scanner();
  memset(s, 0, sizeof(*s));
  return s;
}
void b200_scanner_free(b200_scanner *s) { delete s; }
void b200_scanner_reset(b200_scanner *s) {
  if (s) memset(s, 0, sizeof(*s));
}

void b200_scanner_add_scan(b200_scanner *s, const b200_scan *r) {
  if (!s || !r || s->complete) return;       // number already collected (scan.cpp:43)
  if (r->upside_down || !r->usable) return;  // scan.cpp:49-59
  float *agg;
  if (r->hseg.n_offsets == 15) agg = s->agg15, s->count15++;
  else if (r->hseg.n_offsets == 16) agg = s->agg16, s->count16++;
  else return;
  for (int i = 0; i < 160; i++) agg[i] *= 0.8f;                       // aggregated *= kDecayFactor
  for (int i = 0; i < 160; i++) agg[i] += r->scores[i] * (1 - 0.8f);  // aggregated += scores * (1 - kDecayFactor)
}

void b200_scanner_peek(const b200_scanner *s, float agg15[160], float agg16[160], int32_t counts[2]) {
  memcpy(agg15, s->agg15, sizeof(s->agg15));
  memcpy(agg16, s->agg16, sizeof(s->agg16));
  counts[0] = s->count15, counts[1] = s->count16;
}

int b200_scanner_result(b200_scanner *s, uint8_t digits[16], int32_t *n_numbers) {
  memset(digits, 0, 16);
  *n_numbers = 0;
  if (!s->complete) {
    const int maxc = s->count15 > s->count16 ? s->count15 : s->count16;
    const int minc = s->count15 < s->count16 ? s->count15 : s->count16;
    if (maxc - minc < 3) return 0;   // three-frame lead
    if (minc * 2 > maxc) return 0;   // significant visa-vs-amex opinion
    const float *agg;
    int n;
    if (s->count15 > s->count16) n = 15, agg = s->agg15;
    else n = 16, agg = s->agg16;
    *n_numbers = n;
    uint8_t num[16] = {0};
    for (int i = 0; i < n; i++) {
      const float *row = agg + i * 10;
      float mx = row[0];
      int arg = 0;
      for (int j = 1; j < 10; j++)
        if (row[j] > mx) mx = row[j], arg = j;
      const float sum = tree_sum(row, 0, 10);
      num[i] = digits[i] = (uint8_t)arg;
      if (mx / sum < 0.7f) return 0;  // kMinStability
    }
    const int type = card_type(num, n);
    if (type != kAmbiguous && type != kUnrecognized && luhn(num, n)) {
      s->complete = 1;
      s->n_numbers = n;
      memcpy(s->digits, num, 16);
    }
  }
  if (s->complete) {
    memcpy(digits, s->digits, 16);
    *n_numbers = s->n_numbers;
    return 1;
  }
  return 0;
}

void b200_scanner_add_expiry(b200_scanner *s, const b200_expiry_group *groups, const float *scores, int n, int current_year,
                             int current_month, int allow_past_dates) {
  if (!s || n <= 0 || !groups || !scores) return;  // expiry_extract returns at once when a frame has no groups
  // the frame's groups with their freshly categorized digits (categorize_expiry_digits, expiry_categorize.cpp:148-256)
  ExpiryAgg fresh[kMaxExpiryAgg];
  int n_fresh = 0;
  for (int g = 0; g < n && n_fresh < kMaxExpiryAgg; g++) {
    ExpiryAgg &f = fresh[n_fresh++];
    f.top = groups[g].top, f.left = groups[g].left, f.n_rects = groups[g].n_rects, f.recently_seen = 0, f.total_seen = 0, f.tag = 0;
    memset(f.scores, 0, sizeof(f.scores));
    const int rows[4] = {0, 1, 3, 4};
    for (int r = 0; r < 4; r++) memcpy(f.scores[rows[r]], scores + ((size_t)g * 4 + r) * 10, sizeof(float) * 10);
  }
  expiry_aggregate(s->expiry, &s->n_expiry, fresh, n_fresh);
  // month / year from groups seen at least three times (get_stable_expiry_month_and_year, expiry_categorize.cpp:398-441)
  for (int o = 0; o < s->n_expiry; o++) {
    const ExpiryAgg &g = s->expiry[o];
    if (g.total_seen < 3) continue;
    stable_month_year(g.scores, g.n_rects, current_year, current_month, allow_past_dates != 0, &s->expiry_month, &s->expiry_year);
  }
}

void b200_expiry_month_year_from_scores(const float *scores, int n_chars, int current_year, int current_month, int allow_past_dates,
                                        int32_t *month, int32_t *year) {
  int m = *month, y = *year;
  stable_month_year(reinterpret_cast<const float(*)[10]>(scores), n_chars, current_year, current_month, allow_past_dates != 0, &m, &y);
  *month = m, *year = y;
}

void b200_scanner_expiry(const b200_scanner *s, int32_t *month, int32_t *year) {
  *month = s->expiry_month, *year = s->expiry_year;
}

int b200_scanner_expiry_peek(const b200_scanner *s, int32_t *meta, float *scores, int cap) {
  int n = 0;
  for (; n < s->n_expiry && n < cap; n++) {
    const ExpiryAgg &g = s->expiry[n];
    meta[n * 4 + 0] = g.top, meta[n * 4 + 1] = g.left, meta[n * 4 + 2] = g.recently_seen, meta[n * 4 + 3] = g.total_seen;
    const int rows[4] = {0, 1, 3, 4};
    for (int r = 0; r < 4; r++) memcpy(scores + ((size_t)n * 4 + r) * 10, g.scores[rows[r]], sizeof(float) * 10);
  }
  return n;
}

}  // extern "C"
