// card.io-dmz_b200/csrc/expiry_seg.cu -- best_expiry_seg (scan/expiry_seg.cpp:706-903) for a batch of
// warped cards.
//
//   expiry_scharr_kernel   one CTA per card: llcv_scharr3_dx_abs on the rows below the number (cv/sobel.cpp:706-799)
//                          -> s16 image in a per-card scratch plane + the 258-column row sums (the cvSum loop of
//                          expiry_seg.cpp:752-755).  The data-parallel part: 428 x <=243 pixels in, 2 bytes out each.
//   expiry_stripes_kernel  one thread per card: stripe selection from the row sums
//   expiry_colsum_kernel   one warp per (card, picked stripe): the stripe's 428 column sums (coalesced)
//   expiry_groups_kernel   one THREAD per card: character rectangles, grouping, grid fitting,
//                          trimming and the slash MLP -- the branchy, list-manipulating part -- through the shared
//                          host/device header expiry_seg_core.h (the same code the CPU unit tests pin on the reference).
// Compiled with -fmad=false: the few float expressions must round like the reference's (and like the host build of
// the header).
#include <stdlib.h>

#include "b200_internal.h"
#include "expiry_seg_core.h"

namespace {

constexpr int kScharrThreads = 256;

__global__ void __launch_bounds__(kScharrThreads)
expiry_scharr_kernel(const uint8_t *__restrict__ cards, const uint16_t *__restrict__ y_offsets, int n, int16_t *__restrict__ sob,
                     int32_t *__restrict__ line_sum) {
  const int card_i = blockIdx.x;
  if (card_i >= n) return;
  const uint8_t *card = cards + (size_t)card_i * (xseg::kW * xseg::kH);
  int16_t *out = sob + (size_t)card_i * (xseg::kW * xseg::kH);
  int32_t *ls = line_sum + (size_t)card_i * B200_EXPIRY_SEG_SCRATCH_INTS;  // row sums at the head of the card's scratch
  const int y0 = min((int)y_offsets[card_i] + xseg::kNumberHeight, xseg::kH);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // rows the trimming step may touch above the ROI (at most three) are zero, as after the reference's cvSetZero
  for (int i = threadIdx.x; i < 3 * xseg::kW; i += kScharrThreads) {
    const int y = y0 - 3 + i / xseg::kW;
    if (y >= 0) out[y * xseg::kW + (i % xseg::kW)] = 0;
  }
  for (int y = y0 + warp; y < xseg::kH; y += kScharrThreads / 32) {  // one warp per row: coalesced loads and stores
    int s = 0;
    for (int x = lane; x < xseg::kW; x += 32) {
      const int v = xseg::scharr_abs_at(card, y0, x, y);
      out[y * xseg::kW + x] = (int16_t)v;
      if (x >= 3 * xseg::kSmallW && x < (xseg::kW * 2) / 3) s += v;  // left_edge = 27, right_edge = 285
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) ls[y] = s;
  }
}

// Per-card scratch of kScratchInts 32-bit words: [0, 270) row sums, [272, 280) picked stripes (count, then base / sum pairs),
// [288, 288 + 3 * 428) the column sums of the picked stripes.
constexpr int kScratchInts = B200_EXPIRY_SEG_SCRATCH_INTS, kPickedAt = 272, kColsumAt = 288;
static_assert(kColsumAt + xseg::kMaxStripes * xseg::kW <= kScratchInts && 1 + 2 * xseg::kMaxStripes <= kColsumAt - kPickedAt, "scratch layout");

// stripe selection: one thread per card (a 270-entry scan and a small sort)
__global__ void __launch_bounds__(128)
expiry_stripes_kernel(int32_t *__restrict__ scratch, const uint16_t *__restrict__ y_offsets, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t *sc = scratch + (size_t)i * kScratchInts;
  xseg::StripeSum picked[xseg::kMaxStripes];
  const int yo = (int)y_offsets[i];
  const int np = yo + xseg::kNumberHeight < xseg::kH ? xseg::pick_stripes(sc, yo, picked) : 0;
  sc[kPickedAt] = np;
  for (int p = 0; p < np; p++) sc[kPickedAt + 1 + 2 * p] = picked[p].base_row, sc[kPickedAt + 2 + 2 * p] = picked[p].sum;
}

// column sums of the picked stripes: one warp per (card, stripe), lanes over the 428 columns (coalesced rows) -- the
// sliding 9-wide rectangle sums of the search then need two loads per column instead of 34
__global__ void __launch_bounds__(128)
expiry_colsum_kernel(const int16_t *__restrict__ sob, int32_t *__restrict__ scratch, int n) {
  const int item = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int i = item / xseg::kMaxStripes, p = item - i * xseg::kMaxStripes;
  if (i >= n) return;
  int32_t *sc = scratch + (size_t)i * kScratchInts;
  if (p >= sc[kPickedAt]) return;
  const int base = sc[kPickedAt + 1 + 2 * p], rows = xseg::stripe_rows(base);
  const int16_t *src = sob + (size_t)i * (xseg::kW * xseg::kH) + (size_t)base * xseg::kW;
  for (int c = lane; c < xseg::kW; c += 32) {
    int s = 0;
    for (int r = 0; r < rows; r++) s += src[r * xseg::kW + c];
    sc[kColsumAt + p * xseg::kW + c] = s;
  }
}

__global__ void __launch_bounds__(32)
expiry_groups_kernel(const int16_t *__restrict__ sob, const int32_t *__restrict__ scratch, int n, const float *__restrict__ slash_w,
                     b200_expiry_group *__restrict__ groups, int max_groups, int32_t *__restrict__ n_groups, int32_t *__restrict__ n_dropped,
                     int cards_per_warp) {
  // The search is branchy and data dependent: the cards of one warp execute one after the other wherever they diverge.
  // cards_per_warp < 32 spreads them over more warps (lanes 0, 32 / cpw, ...), but a lone lane uses 4 bytes of every
  // 128-byte local-memory line: measured on B200 (65 536 deck cards, cards/s) 32 per warp wins once the batch is deep --
  // chunk 2048: 55 k (32) / 75 k (8) / 88 k (2); chunk 8192: 168 k / 172 k / 100 k; chunk 32768: 344 k (32) / 239 k (16) / 180 k (8).
  const int stride = 32 / cards_per_warp;
  if (threadIdx.x % stride != 0) return;
  const int i = blockIdx.x * cards_per_warp + threadIdx.x / stride;
  if (i >= n) return;
  const int32_t *sc = scratch + (size_t)i * kScratchInts;
  int overflow = 0;
  xseg::StripeSum picked[xseg::kMaxStripes];
  const int np = sc[kPickedAt];
  for (int p = 0; p < np; p++) picked[p].base_row = sc[kPickedAt + 1 + 2 * p], picked[p].sum = sc[kPickedAt + 2 + 2 * p];
  const int k = xseg::search_stripes(sob + (size_t)i * (xseg::kW * xseg::kH), picked, np, slash_w,
                                     reinterpret_cast<xseg::ExpiryGroupOut *>(groups + (size_t)i * max_groups), max_groups, &overflow,
                                     sc + kColsumAt);
  n_groups[i] = k;
  if (n_dropped) n_dropped[i] = overflow;
}

}  // namespace

static_assert(sizeof(b200_expiry_group) == sizeof(xseg::ExpiryGroupOut), "b200_expiry_group mirrors xseg::ExpiryGroupOut");

int launch_expiry_seg(const uint8_t *cards, const uint16_t *y_offsets, int n, const float *slash_w, int16_t *sob, int32_t *line_sum,
                      b200_expiry_group *groups, int max_groups, int32_t *n_groups, int32_t *n_dropped, cudaStream_t s) {
  expiry_scharr_kernel<<<n, kScharrThreads, 0, s>>>(cards, y_offsets, n, sob, line_sum);
  if (cudaGetLastError() != cudaSuccess) return -1;
  static const int cpw = [] {  // measured on B200 (tools/gpu_side_bench.py): see DESIGN.md
    const char *e = getenv("B200_DMZ_EXPIRY_CARDS_PER_WARP");
    const int v = e && *e ? atoi(e) : 32;
    return v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32 ? v : 32;
  }();
  expiry_stripes_kernel<<<(n + 127) / 128, 128, 0, s>>>(line_sum, y_offsets, n);
  expiry_colsum_kernel<<<(n * xseg::kMaxStripes + 3) / 4, 128, 0, s>>>(sob, line_sum, n);
  expiry_groups_kernel<<<(n + cpw - 1) / cpw, 32, 0, s>>>(sob, line_sum, n, slash_w, groups, max_groups, n_groups, n_dropped, cpw);
  return cudaGetLastError() == cudaSuccess ? 4 : -1;
}
