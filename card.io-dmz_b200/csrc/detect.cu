// card.io-dmz_b200/csrc/detect.cu -- best_line_for_sample for every (frame, strip): one CTA per strip.
//
// Replaces, for one detection strip (dmz.cpp:224-271):
//   llcv_sobel7 x2                         cv/sobel.cpp:476-530   (7x7 separable, replicate border, s16 saturation)
//   llcv_adaptive_canny7_precomputed_sobel cv/canny.cpp:555-580   (thresholds) + canny.cpp:58-336 (NMS, hysteresis)
//   llcv_hough                             cv/hough.cpp:52-196    (gradient-gated votes, first-max argmax)
//
// Data flow inside the CTA (everything after the first load stays in shared memory):
//   global u8 strip --(32-bit coalesced loads)--> s_src[h][ws] with a 3-pixel replicated halo left and right
//   Sobel: one work item per (column, row chunk) walks down its rows; each row filter is two dp4a over the eight
//          bytes around the pixel (3 aligned LDS.32 + funnel shifts), the last seven row results of both kernels
//          live in a register ring (no intermediate image).  What is stored per pixel is what the later stages read:
//            s_md   u32  low half |dx| + |dy| (65536 is stored as 65535 + a flag bit), high half dx (s16): ONE store, and
//                        one LDS.128 later yields four magnitudes; dy = +-(mag - |dx|) with its sign in the flag byte
//            s_map  u8   bits 0-1 state {0 candidate, 1 no edge, 2 edge}, bit 4 dy < 0, bit 5 mag == 65536
//          both with a zero / "no edge" border and a row pitch that is a multiple of four pixels
//   thresholds: block reduction (warp shuffles) of the saturated |dx| + |dy| sums, 64-bit exact
//   NMS step 1: FOUR magnitudes per LDS.128 against the low threshold; the ones that pass are queued per warp
//   NMS step 2: 32 dense lanes: sector from (|dx|, |dy|, signs) -> neighbour offset, two neighbour magnitudes, state
//          byte; weak candidates go to a work list, strong ones straight to the vote list
//   hysteresis: propagation over the candidate list to the unique fixed point (= the reference's stack walk)
//   Hough: gated edge pixels go to a vote list; shared-memory atomics into a compacted accumulator (only the
//          reachable rho range per angle)
//   argmax: packed (votes, -(r, n)) 64-bit keys reduced with warp shuffles -> reference scan order r outer,
//           n inner, strict '>'
// All integer; the float comparisons of the gradient gate are IEEE divisions (-fmad=false file).  No integer
// division appears inside any per-pixel loop (2-D walks are set up once per thread).
#include <float.h>

#include <type_traits>

#include "b200_internal.h"

namespace {

constexpr int kThreads = 416;  // 13 warps: one Sobel work item per thread for the 389-wide strips
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Work-list entries are padded pixel indices o < (w + 2)(h + 2).  16 bits hold them for every strip that keeps its
// gradients in shared memory (npad <= 65535 is part of that variant's admission test in api.cu); the global-gradient
// variant serves the large strips (1080p portrait: 88 x 899 = 79112) and stores 32-bit indices.
template <typename IdxT>
struct SmemLayout {
  uint8_t *src;          // h x w source strip (dead after Sobel: its storage then holds the work lists)
  unsigned int *md;      // (h + 2) x wpm: |dx| + |dy| (low half, zero border: no bounds checks in the NMS neighbourhood) | dx << 16
  uint8_t *map;          // (h + 2) x wpm: state | (dy < 0) << 4 | (mag == 65536) << 5; border = 1
  IdxT *list;            // candidate list (hysteresis), then vote list (Hough): padded pixel indices; aliases src
  unsigned int *acc;     // compacted Hough accumulator
};

__host__ __device__ __forceinline__ size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
// source row stride: up to 6 bytes before the strip, 3-px halo after it, 4-byte aligned, room for the 12-byte reads
__host__ __device__ __forceinline__ int detect_src_stride(int w) { return ((w + 6 + 3 + 3) & ~3) + 8; }
// row pitch (pixels) of the padded magnitude / dx / map arrays: w + 2 rounded up to four (LDS.64 = four magnitudes)
__host__ __device__ __forceinline__ int detect_pad_pitch(int w) { return (w + 2 + 3) & ~3; }

// Per-thread 2-D walk over a w x h strip without integer division inside the loop:
//   w <= T: tx = tid % w, ty = tid / w (computed once), rows advance by T / w;  w > T: tx = tid, columns advance by T.
struct Walk {
  int tx, ty, xstep, ystep;
  bool active;
};
__device__ __forceinline__ Walk make_walk(int tid, int w, int T) {
  Walk k;
  if (w <= T) {
    k.ystep = T / w;
    k.ty = tid / w;
    k.tx = tid - k.ty * w;
    k.xstep = w;  // a single column per thread
    k.active = k.ty < k.ystep;
  } else {
    k.ystep = 1, k.ty = 0, k.tx = tid, k.xstep = T, k.active = true;
  }
  return k;
}

// kGlobalGrad: dx / dy live in a global scratch (strips too large for shared memory, 1080p).  A template parameter and
// not a run-time switch so that in the usual case every dx / dy access is a plain 32-bit-addressed LDS / STS (a pointer
// that may be shared OR global compiles to generic 64-bit-addressed loads).
template <bool kGlobalGrad>
__global__ void __launch_bounds__(kThreads, 3)
detect_strips_kernel(const __grid_constant__ DetectParams P, const uint8_t *__restrict__ plane, int row_stride,
                     size_t frame_stride, const b200_line *__restrict__ prev_lines, const b200_line *__restrict__ prev_lines2,
                     b200_line *__restrict__ lines, int16_t *__restrict__ grad_scratch, size_t grad_scratch_stride, int ox, int oy) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ unsigned long long s_red[kThreads / 32];
  __shared__ int s_low, s_high, s_nvote, s_nedge, s_overflow, s_sat;
  __shared__ int s_cpre[kWarps + 1];          // per-warp candidate counts, then their exclusive prefix

  const int strip = blockIdx.x;
  const int frame = blockIdx.y;
  const int tid = threadIdx.x;
  const StripDesc &S = P.strip[strip];
  const int w = S.w, h = S.h, npx = w * h;
  const int wp = detect_pad_pitch(w), npad = wp * (h + 2);  // padded arrays: pixel (x, y) at (y + 1) * wp + x + 1
  // Source rows in shared memory: pixel c of the strip sits at byte H + c, H in [3, 6] chosen so that the aligned 32-bit
  // words of the global row land on aligned shared-memory words (whole-word copies); 3-px replicated halo each side.
  const bool word_ok = ((reinterpret_cast<uintptr_t>(plane) | (uintptr_t)row_stride | (uintptr_t)frame_stride) & 3u) == 0;
  const int shift = word_ok ? ((S.x - ox) & 3) : 0;  // bytes of the first aligned word that precede the strip
  const int H = 3 + ((shift + 1) & 3);               // (H - shift) % 4 == 0
  const int ws = detect_src_stride(w);
  const size_t out_idx = (size_t)frame * 4 + strip;

  // Fallback planes: skip strips whose edge was already found on an earlier plane (dmz.cpp:351).
  if ((prev_lines != nullptr && prev_lines[out_idx].found) || (prev_lines2 != nullptr && prev_lines2[out_idx].found)) {
    if (tid == 0) {
      b200_line l;
      l.found = 0, l.r = 0, l.n = 0, l.max_votes = 0, l.low = 0, l.high = 0, l.n_edge_px = 0;
      l.rho = FLT_MAX, l.theta = FLT_MAX;
      lines[out_idx] = l;
    }
    return;
  }

  // ---- carve shared memory
  using IdxT = typename std::conditional<kGlobalGrad, unsigned int, unsigned short>::type;
  constexpr int kIdxShift = kGlobalGrad ? 2 : 1;  // log2(sizeof(IdxT))
  SmemLayout<IdxT> L;
  size_t off = 0;
  L.src = smem_raw + off;
  off = align16(off + (size_t)ws * h);
  if constexpr (kGlobalGrad) {
    L.md = reinterpret_cast<unsigned int *>(grad_scratch + ((size_t)frame * 4 + strip) * grad_scratch_stride);  // stride = 2 npad int16
  } else {
    L.md = reinterpret_cast<unsigned int *>(smem_raw + off);
    off = align16(off + (size_t)npad * 4);
  }
  L.map = smem_raw + off;
  off = align16(off + (size_t)npad);
  // the Hough accumulator (used after the NMS) shares its storage with the per-warp rings of the NMS (128 entries each)
  L.acc = reinterpret_cast<unsigned int *>(smem_raw + off);
  IdxT *q_rings = reinterpret_cast<IdxT *>(smem_raw + off);
  // work lists alias the (by then dead) source strip: one candidate segment per warp (filled without atomics), then
  // the vote list; overflow falls back to full scans
  L.list = reinterpret_cast<IdxT *>(L.src);
  const int cand_cap = (((ws * h) >> kIdxShift) * 3 / 5) / kWarps;  // entries per warp segment
  IdxT *vote_list = L.list + cand_cap * kWarps;
  const int list_cap = ((ws * h) >> kIdxShift) - cand_cap * kWarps;  // entries of the vote list

  // ---- 1. load the strip (32-bit coalesced loads of the covering aligned words); clear borders / accumulator
  {
    const uint8_t *base = plane + (size_t)frame * frame_stride + (size_t)(S.y - oy) * row_stride + (S.x - ox);  // plane origin = (ox, oy)
    if (word_ok) {
      const int words = (shift + w + 3) >> 2;  // words per row; the bytes around the strip they carry are overwritten by the halo
      const int q0 = (H - shift) >> 2;         // shared-memory word of the first global word
      const Walk k = make_walk(tid, words, kThreads);
      if (k.active) {
        if (k.xstep >= words) {
          // one column of words per thread (every strip narrower than 4 * kThreads pixels): pointers advance by
          // constants, four independent loads in flight
          const unsigned int *gp = reinterpret_cast<const unsigned int *>(base + (size_t)k.ty * row_stride - shift) + k.tx;
          unsigned int *sp = reinterpret_cast<unsigned int *>(L.src + k.ty * ws) + q0 + k.tx;
          const size_t gstep = ((size_t)k.ystep * row_stride) >> 2;
          const int sstep = (k.ystep * ws) >> 2;
          int row = k.ty;
          for (; row + 3 * k.ystep < h; row += 4 * k.ystep) {
            const unsigned int a = __ldg(gp), b = __ldg(gp + gstep), c = __ldg(gp + 2 * gstep), d = __ldg(gp + 3 * gstep);
            sp[0] = a, sp[sstep] = b, sp[2 * sstep] = c, sp[3 * sstep] = d;
            gp += 4 * gstep, sp += 4 * sstep;
          }
          for (; row < h; row += k.ystep, gp += gstep, sp += sstep) *sp = __ldg(gp);
        } else {
          for (int row = k.ty; row < h; row += k.ystep)
            for (int q = k.tx; q < words; q += k.xstep)
              reinterpret_cast<unsigned int *>(L.src + row * ws)[q0 + q] =
                  __ldg(reinterpret_cast<const unsigned int *>(base + (size_t)row * row_stride - shift) + q);
        }
      }
    } else {
      const Walk k = make_walk(tid, w, kThreads);
      if (k.active)
        for (int row = k.ty; row < h; row += k.ystep)
          for (int c = k.tx; c < w; c += k.xstep) L.src[row * ws + H + c] = __ldg(base + (size_t)row * row_stride + c);
    }
  }
  // border of the padded arrays: magnitude 0, state "no edge" (top and bottom rows, column 0 and columns w + 1 .. wp - 1)
  for (int i = tid; i < wp; i += kThreads) {
    const int j = (h + 1) * wp + i;
    L.md[i] = 0u, L.map[i] = 1;
    L.md[j] = 0u, L.map[j] = 1;
  }
  {
    const int nb = wp - w;  // border columns per row: 1 on the left, nb - 1 on the right
    for (int i = tid; i < h * nb; i += kThreads) {
      const int y = i / nb, c = i - y * nb;
      const int a = (y + 1) * wp + (c == 0 ? 0 : w + c);
      L.md[a] = 0u, L.map[a] = 1;
    }
  }
  if (tid == 0) s_nvote = 0, s_nedge = 0, s_overflow = 0, s_sat = 0;
  __syncthreads();
  // BORDER_REPLICATE halo: three copies of the first / last pixel of every row
  for (int y = tid; y < h; y += kThreads) {
    uint8_t *r = L.src + y * ws + H;
    const uint8_t a = r[0], b = r[w - 1];
    r[-3] = a, r[-2] = a, r[-1] = a;
    r[w] = b, r[w + 1] = b, r[w + 2] = b;
  }
  __syncthreads();

  // ---- 2. Sobel-7 dx, dy; accumulate the saturated |.| sums on the fly
  unsigned long long abs_sum = 0;
  {
    // the four dp4a tap words live in per-thread registers for the whole pass: as uniform values the compiler would
    // re-materialise them (UMOV) for every output pixel
    unsigned int kd0, kd1, ks0, ks1;
    asm volatile("mov.u32 %0, 0x00FBFCFF;" : "=r"(kd0));  // -1 -4 -5  0
    asm volatile("mov.u32 %0, 0x00010405;" : "=r"(kd1));  //  5  4  1  .
    asm volatile("mov.u32 %0, 0x140F0601;" : "=r"(ks0));  //  1  6 15 20
    asm volatile("mov.u32 %0, 0x0001060F;" : "=r"(ks1));  // 15  6  1  .
    unsigned int sat_any = 0;
    const int items = w * S.nchunks;
    for (int it = tid; it < items; it += kThreads) {
      unsigned int col_sum = 0;  // <= chunk_rows * 65534: 32 bits hold any strip a frame can have
      const int chunk = it / w, x = it - chunk * w;
      const int y0 = chunk * S.chunk_rows;
      const int y1 = min(h, y0 + S.chunk_rows);
      const int xs = x + H - 3;                     // byte of the first of the seven taps (pixel x - 3)
      const int word = xs >> 2, sh = (xs & 3) * 8;
      int hx[7], sx[7];  // ring of row-filter results: derivative taps [-1,-4,-5,0,5,4,1], smoothing taps [1,6,15,20,15,6,1]
      auto row_filter = [&](int row, int &hxo, int &sxo) {
        const unsigned int *r32 = reinterpret_cast<const unsigned int *>(L.src + row * ws) + word;
        const unsigned int w0 = r32[0], w1 = r32[1], w2 = r32[2];
        const unsigned int lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);  // p0..p3, p4..p7
        int a, b;
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(a) : "r"(lo), "r"(kd0), "r"(0));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(a) : "r"(hi), "r"(kd1), "r"(a));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(b) : "r"(lo), "r"(ks0), "r"(0));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(b) : "r"(hi), "r"(ks1), "r"(b));
        hxo = a, sxo = b;
      };
#pragma unroll
      for (int k = 0; k < 6; k++) row_filter(clampi(y0 + k - 3, 0, h - 1), hx[k], sx[k]);  // rows y0-3 .. y0+2
      int o = (y0 + 1) * wp + x + 1;
      // one output row; ring slot k of every tap is a compile-time constant, no register shuffling
      auto out_row = [&](int y, int k) {
        row_filter(min(y + 3, h - 1), hx[(k + 6) % 7], sx[(k + 6) % 7]);
        int gx = (hx[k % 7] + hx[(k + 6) % 7]) + 6 * (hx[(k + 1) % 7] + hx[(k + 5) % 7]) +
                 15 * (hx[(k + 2) % 7] + hx[(k + 4) % 7]) + 20 * hx[(k + 3) % 7];        // smooth down the column
        int gy = (sx[(k + 6) % 7] - sx[k % 7]) + 4 * (sx[(k + 5) % 7] - sx[(k + 1) % 7]) +
                 5 * (sx[(k + 4) % 7] - sx[(k + 2) % 7]);                                  // derivative down the column
        gx = clampi(gx, -32768, 32767);  // saturate_cast<short>
        gy = clampi(gy, -32768, 32767);
        // what the NMS reads (canny.cpp:222-236): the magnitude, dx, and two flag bits from which dy follows
        const unsigned int ax = (unsigned)abs(gx), ay = (unsigned)abs(gy);
        const unsigned int m = ax + ay;      // <= 65536
        const unsigned int hi16 = m >> 16;   // 1 only for dx == dy == -32768
        L.md[o] = (m - hi16) | ((unsigned)gx << 16);  // 65536 -> 65535 + flag bit
        L.map[o] = (uint8_t)(1u | (((unsigned)gy >> 31) << 4) | (hi16 << 5));
        sat_any |= hi16;  // (block-wide flag below: step 2 then rebuilds the 17-bit magnitudes)
        col_sum += min(ax, 32767u) + min(ay, 32767u);  // cvAbs saturates, canny.cpp:355-361
        o += wp;
      };
      int y = y0;
      for (; y + 7 <= y1; y += 7) {  // seven output rows per trip, no per-row bound test
#pragma unroll
        for (int k = 0; k < 7; k++) out_row(y + k, k);
      }
#pragma unroll
      for (int k = 0; k < 6; k++)  // the last y1 - y < 7 rows
        if (y + k < y1) out_row(y + k, k);
      abs_sum += col_sum;
    }
    if (sat_any) s_sat = 1;
  }
  // ---- 3. adaptive thresholds: low = floor(mean), high = floor(3 * mean), canny.cpp:568-580
  abs_sum = warp_sum_u64(abs_sum);
  if ((tid & 31) == 0) s_red[tid >> 5] = abs_sum;
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int i = 0; i < kThreads / 32; i++) tot += s_red[i];
    const double mean = (double)tot / (double)npx;
    s_low = (int)floor(mean);
    s_high = (int)floor(3.0 * mean);  // 3.0f * low_threshold evaluated in double
  }
  __syncthreads();
  const int low = s_low, high = s_high;
  const bool sat = s_sat != 0;  // some magnitude is 65536: compare 17-bit values (mag + flag bit)

  // gradient-direction gate of the Hough stage (hough.cpp:126-150) for the pixel at padded index o
  auto gate = [&](int o) -> bool {
    const unsigned int mdv = L.md[o], b = L.map[o];
    const int del_x = (int)mdv >> 16;
    const int ady = (int)(mdv & 0xFFFFu) + (int)((b >> 5) & 1u) - abs(del_x);
    const int del_y = (b & 16u) ? -ady : ady;
    if (del_x != 0) {
      const float slope = (float)del_y / (float)del_x;
      return S.vertical ? (slope >= S.slope_a && slope <= S.slope_b) : (slope >= S.slope_a || slope <= S.slope_b);
    }
    return !S.vertical;
  };
  auto push_vote = [&](int o) {
    const int slot = atomicAdd(&s_nvote, 1);
    if (slot < list_cap) vote_list[slot] = (IdxT)o;
    else s_overflow = 1;
  };
  int n_edge = 0;

  // ---- 4. non-maxima suppression, canny.cpp:220-285.  The source strip is dead now: weak candidates go to the
  // hysteresis work lists, strong pixels (edges for sure) to the vote list (the direction gate is applied when voting).
  // Two steps per warp, no block barrier between them.  Step 1 walks the padded array four pixels per lane (one 128-bit
  // load; border pixels hold 0 and never pass): magnitude against the low threshold, the pixels that pass are queued in a
  // 128-entry ring of the warp -- two pixels per ballot round (a lane's count is 0..2: two ballots give its slot and the
  // warp total, no atomics).  Step 2, whenever 32 are queued: 32 DENSE lanes rebuild (|dx|, |dy|, signs) from the packed
  // word and the flag byte, turn the sector into a neighbour offset by selects, fetch the two neighbour magnitudes and set
  // the state -- one straight-line body.  Weak candidates are appended to the warp's own list segment, again without
  // atomics.  (The reference's int64 products fit in 32 bits here: |dx|, |dy| <= 32768 after the s16 saturation.)
  {
    const int lane = tid & 31, wid = tid >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    IdxT *q = q_rings + wid * 128;
    IdxT *clist = L.list + wid * cand_cap;
    int qhead = 0, qtail = 0, ccount = 0;
    auto direction_test = [&](int cnt) {  // the first cnt queued pixels, one per lane
      bool weak = false, strong = false;
      int o = 0;
      if (lane < cnt) {
        o = q[(qhead + lane) & 127];
        const unsigned int mdv = L.md[o], b = L.map[o];
        const unsigned int hi = sat ? (b >> 5) & 1u : 0u;
        const int gx = (int)mdv >> 16;
        const int m = (int)((mdv & 0xFFFFu) + hi);
        const unsigned int ax = (unsigned)abs(gx), ay = (unsigned)m - ax;
        const unsigned int tg22x = ax * 13573u;  // TG22 = (int)(0.41421356 * 2^15 + 0.5)
        const unsigned int ys = ay << 15;        // tg22x <= 4.5e8, tg67x = tg22x + ax 2^16 <= 2.6e9 < 2^32, ys <= 2^30
        const bool horiz = ys < tg22x, vert = ys > tg22x + (ax << 16);
        const unsigned int opposite = ((unsigned)gx >> 31) ^ ((b >> 4) & 1u);  // sign(dx) != sign(dy), as (dx ^ dy) < 0
        // horizontal: m > left && m >= right; vertical: m > up && m >= down; diagonal: m > both, along the gradient sign
        int off = wp + 1 - 2 * (int)opposite;
        off = vert ? wp : off;
        off = horiz ? 1 : off;
        const int ge = (horiz || vert) ? 1 : 0;
        int m0 = (int)(L.md[o - off] & 0xFFFFu), m1 = (int)(L.md[o + off] & 0xFFFFu);
        if (sat) m0 += (int)((L.map[o - off] >> 5) & 1u), m1 += (int)((L.map[o + off] >> 5) & 1u);
        const bool is_max = m > m0 && m + ge > m1;
        strong = is_max && m > high;
        weak = is_max && !strong;
        if (is_max) L.map[o] = (uint8_t)((b & ~3u) | (strong ? 2u : 0u));  // (non-maxima keep state 1)
      }
      const unsigned int wm = __ballot_sync(0xffffffffu, weak);
      if (weak) {
        const int slot = ccount + __popc(wm & lt);
        if (slot < cand_cap) clist[slot] = (IdxT)o;
        else s_overflow = 1;
      }
      ccount += __popc(wm);
      if (strong) {
        n_edge++;
        push_vote(o);
      }
    };
    // groups of four padded pixels, rows 1 .. h (the all-zero border rows are skipped)
    const int g0 = wp >> 2, ngroups = (wp >> 2) * h;
    const int iters = (ngroups + kThreads - 1) / kThreads;  // the same trip count for every lane: the loop holds warp votes
    const uint4 *md4 = reinterpret_cast<const uint4 *>(L.md);
    const unsigned int ulow = (unsigned)low;
    for (int it = 0; it < iters; it++) {
      const int g = it * kThreads + tid;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (g < ngroups) v = md4[g0 + g];
      const int o4 = (g0 + g) << 2;
      const unsigned int mg[4] = {v.x & 0xFFFFu, v.y & 0xFFFFu, v.z & 0xFFFFu, v.w & 0xFFFFu};
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        // (a clipped 65536 reads 65535 > low as well: low <= 65534)
        const bool p0 = mg[k] > ulow, p1 = mg[k + 1] > ulow;
        const unsigned int b1 = __ballot_sync(0xffffffffu, p0 != p1), b2 = __ballot_sync(0xffffffffu, p0 && p1);  // count 1, count 2
        int pos = qtail + __popc(b1 & lt) + 2 * __popc(b2 & lt);
        if (p0) q[pos & 127] = (IdxT)(o4 + k), pos++;
        if (p1) q[pos & 127] = (IdxT)(o4 + k + 1);
        qtail += __popc(b1) + 2 * __popc(b2);
        while (qtail - qhead >= 32) {  // (at most 31 + 64 queued: the ring never wraps onto live entries)
          __syncwarp();  // queue entries visible
          direction_test(32);
          qhead += 32;
          __syncwarp();  // entries consumed before the ring wraps onto them
        }
      }
    }
    __syncwarp();
    if (qtail > qhead) direction_test(qtail - qhead);
    if (lane == 0) s_cpre[wid] = min(ccount, cand_cap);
  }
  __syncthreads();
  for (int i = tid; i < S.ncells; i += kThreads) L.acc[i] = 0u;  // the rings are dead: their storage becomes the accumulator
  if (tid == 0) {  // exclusive prefix of the per-warp candidate counts
    int run = 0;
    for (int i = 0; i < kWarps; i++) {
      const int c = s_cpre[i];
      s_cpre[i] = run;
      run += c;
    }
    s_cpre[kWarps] = run;
  }
  __syncthreads();

  // ---- 5. hysteresis: a candidate 8-connected to an edge pixel becomes an edge pixel; iterate over the (short)
  // candidate lists to the fixed point, which is the reference's stack-walk result whatever the visiting order.
  // A promoted candidate is queued for voting on the spot (each candidate is promoted exactly once).
  // Within one sweep a thread may read a neighbour's map byte while its owner promotes it (racecheck reports that
  // read/write pair as a warning): the state only ever goes 0 -> 2, a stale 0 just defers the promotion to the next sweep,
  // and the loop ends only after a sweep without any promotion, so the fixed point does not depend on the interleaving.
  // (state 2 is the only one with bit 1 set: "some neighbour is an edge" is one OR over the eight bytes)
  {
    if (!s_overflow) {
      const int ncand = s_cpre[kWarps];
      while (true) {
        int changed = 0;
        for (int c = tid; c < ncand; c += kThreads) {
          int seg = 0;
          while (c >= s_cpre[seg + 1]) seg++;
          const int o = L.list[seg * cand_cap + (c - s_cpre[seg])];
          const uint8_t *m = L.map + o;
          const unsigned int b = *m;
          if ((b & 3u) != 0u) continue;
          const unsigned int any = m[-wp - 1] | m[-wp] | m[-wp + 1] | m[-1] | m[1] | m[wp - 1] | m[wp] | m[wp + 1];
          if (any & 2u) {
            L.map[o] = (uint8_t)(b | 2u);
            changed = 1;
            n_edge++;
            push_vote(o);
          }
        }
        if (!__syncthreads_or(changed)) break;
      }
    } else {
      // a candidate segment overflowed (pathological texture): sweep the whole map instead
      const Walk k = make_walk(tid, w, kThreads);
      while (true) {
        int changed = 0;
        if (k.active)
          for (int y = k.ty; y < h; y += k.ystep)
            for (int x = k.tx; x < w; x += k.xstep) {
              const int o = (y + 1) * wp + x + 1;
              const uint8_t *m = L.map + o;
              const unsigned int b = *m;
              if ((b & 3u) != 0u) continue;
              const unsigned int any = m[-wp - 1] | m[-wp] | m[-wp + 1] | m[-1] | m[1] | m[wp - 1] | m[wp] | m[wp + 1];
              if (any & 2u) {
                L.map[o] = (uint8_t)(b | 2u);
                changed = 1;
              }
            }
        if (!__syncthreads_or(changed)) break;
      }
    }
  }
  __syncthreads();  // s_overflow / s_nvote final

  // ---- 6/7. votes, hough.cpp:152-158: shared-memory atomics on the compacted accumulator
  if (!s_overflow) {
    const int nvote = s_nvote;
    for (int e = tid; e < nvote; e += kThreads) {
      const int o = vote_list[e];
      if (!gate(o)) continue;  // gradient-direction gate, hough.cpp:126-150: dense here, divergent if applied while queueing
      const int yy = o / wp;
      const int x = o - yy * wp - 1, y = yy - 1;
#pragma unroll
      for (int n = 0; n < B200_NUMANGLE; n++) {
        const int r = ((x * S.tab_cos[n] + y * S.tab_sin[n]) >> 10) + S.half;
        atomicAdd(&L.acc[S.cell_base[n] + (r - S.rlo[n])], 1u);
      }
    }
  } else {
    // a work list overflowed: ignore the lists, recount and vote by scanning the final map
    n_edge = 0;
    const Walk k = make_walk(tid, w, kThreads);
    if (k.active)
      for (int y = k.ty; y < h; y += k.ystep)
        for (int x = k.tx; x < w; x += k.xstep) {
          const int o = (y + 1) * wp + x + 1;
          if ((L.map[o] & 3u) != 2u) continue;
          n_edge++;
          if (gate(o)) {
#pragma unroll
            for (int n = 0; n < B200_NUMANGLE; n++) {
              const int r = ((x * S.tab_cos[n] + y * S.tab_sin[n]) >> 10) + S.half;
              atomicAdd(&L.acc[S.cell_base[n] + (r - S.rlo[n])], 1u);
            }
          }
        }
  }
  n_edge = (int)warp_sum_u64((unsigned long long)n_edge);
  if ((tid & 31) == 0 && n_edge) atomicAdd(&s_nedge, n_edge);
  __syncthreads();

  // ---- 8. argmax in the reference's scan order (r outer, n inner, first strict maximum), hough.cpp:165-176
  unsigned long long best = 0;
  for (int c = tid; c < S.ncells; c += kThreads) {  // one flat pass over the compacted cells; angle n owns [cell_base[n], cell_base[n + 1])
    const unsigned int v = L.acc[c];
    if (v == 0) continue;
    int n = 0;
#pragma unroll
    for (int k = 1; k < B200_NUMANGLE; k++) n += c >= S.cell_base[k];
    const unsigned int r = (unsigned)(S.rlo[n] + (c - S.cell_base[n]));
    const unsigned long long key = ((unsigned long long)v << 32) | (0xFFFFFFFFu - (r * 16u + (unsigned)n));
    best = key > best ? key : best;
  }
  best = warp_max_u64(best);
  if ((tid & 31) == 0) s_red[tid >> 5] = best;  // s_red's threshold use ended several barriers ago
  __syncthreads();
  if (tid == 0) {
    unsigned long long b = 0;
    for (int i = 0; i < kThreads / 32; i++) b = s_red[i] > b ? s_red[i] : b;
    b200_line l;
    l.max_votes = (int)(b >> 32);
    l.low = low, l.high = high, l.n_edge_px = s_nedge;
    l.found = 0, l.r = 0, l.n = 0, l.rho = FLT_MAX, l.theta = FLT_MAX;
    if (l.max_votes > S.threshold) {
      const unsigned int rn = 0xFFFFFFFFu - (unsigned int)(b & 0xFFFFFFFFu);
      l.found = 1;
      l.r = (int)(rn >> 4);
      l.n = (int)(rn & 15u);
      l.rho = ((float)l.r - (float)(S.numrho - 1) * 0.5f) * 1.0f;  // hough.cpp:189
      l.theta = S.theta[l.n];
    }
    lines[out_idx] = l;
  }
}

}  // namespace

size_t detect_smem_bytes(const DetectParams &p) {
  size_t worst = 0;
  for (int s = 0; s < 4; s++) {
    const StripDesc &d = p.strip[s];
    const size_t npad = (size_t)detect_pad_pitch(d.w) * (d.h + 2);
    const size_t ws = (size_t)detect_src_stride(d.w);
    size_t b = align16(ws * d.h) + align16(npad);  // padded src (later the work lists), map
    if (!p.use_global_grad) b += align16(npad * 4);
    const size_t acc = align16((size_t)d.ncells * 4), rings = (size_t)kWarps * 128 * (p.use_global_grad ? 4 : 2);
    b += (acc > rings ? acc : rings) + 64;  // accumulator and per-warp rings share storage
    worst = b > worst ? b : worst;
  }
  return worst;
}

int launch_detect(const DetectParams &p, const uint8_t *plane, int row_stride, size_t frame_stride, int n,
                  const b200_line *prev_lines, const b200_line *prev_lines2, b200_line *lines, int16_t *grad_scratch,
                  cudaStream_t s, int ox, int oy) {
  size_t smem = detect_smem_bytes(p);
  // largest dynamic shared-memory size opted into, per device (contexts on distinct threads may race here: the attribute
  // call is idempotent and the recorded maximum only grows)
  static std::atomic<size_t> configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(detect_strips_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(detect_strips_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    size_t seen = configured[dev & 63].load(std::memory_order_relaxed);
    while (seen < smem && !configured[dev & 63].compare_exchange_weak(seen, smem, std::memory_order_release)) {
    }
  }
  size_t max_npad = 0;
  for (int i = 0; i < 4; i++) {
    size_t v = (size_t)detect_pad_pitch(p.strip[i].w) * (p.strip[i].h + 2);
    max_npad = v > max_npad ? v : max_npad;
  }
  // grid.y is limited to 65535 blocks: split very large batches
  int launches = 0;
  for (int f0 = 0; f0 < n; f0 += 65535) {
    int cnt = n - f0 < 65535 ? n - f0 : 65535;
    const uint8_t *pp = plane + (size_t)f0 * frame_stride;
    const b200_line *p1 = prev_lines ? prev_lines + (size_t)f0 * 4 : nullptr, *p2 = prev_lines2 ? prev_lines2 + (size_t)f0 * 4 : nullptr;
    int16_t *gs = grad_scratch ? grad_scratch + (size_t)f0 * 4 * max_npad * 2 : nullptr;
    if (p.use_global_grad)
      detect_strips_kernel<true><<<dim3(4, cnt), kThreads, smem, s>>>(p, pp, row_stride, frame_stride, p1, p2, lines + (size_t)f0 * 4, gs, max_npad * 2, ox, oy);
    else
      detect_strips_kernel<false><<<dim3(4, cnt), kThreads, smem, s>>>(p, pp, row_stride, frame_stride, p1, p2, lines + (size_t)f0 * 4, gs, max_npad * 2, ox, oy);
    launches++;
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}
