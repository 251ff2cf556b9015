// formats.cu -- the pixel-format steps either side of the detect -> warp -> OCR path (widening rows):
//
//   ycbcr_to_rgb_kernel   dmz_YCbCr_to_RGB            dmz.cpp:58-64 -> llcv_YCbCr2RGB_u8_c, cv/convert.cpp:449-504
//   rgba_to_r_kernel      dmz_deinterleave_RGBA_to_R  dmz.cpp:66-109
//   stencil3_kernel       dmz_scharr3_dx_abs / dmz_scharr3_dy_abs / dmz_sobel3_dx_dy   dmz.cpp:519-531, cv/sobel.cpp:556-900
//
// All three are byte / int16 streaming work bounded by HBM (6 - 7, 5 and 3 bytes per pixel), so the kernels are about
// instructions per byte: 128-bit accesses where the shapes allow, the colour arithmetic as two-way dot products
// (IDP.2A: one instruction per channel per pixel with the luma and the rounding constant riding in the accumulator) and
// saturating packs (I2IP: clamp + pack of two bytes), the stencils on four pixels per thread with byte-SIMD absolute
// differences and 16-bit pairs in one register.  Results are integers: bit-exact against the reference.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200_internal.h"

namespace {

inline int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// Grid-stride kernels are launched with exactly the CTAs that can be resident (SMs x occupancy): a larger grid runs a
// partial second wave on a mostly idle GPU (measured on the stencils: 1184 CTAs over 888 slots cost a third of the time).
template <typename Kernel>
inline size_t resident_ctas(Kernel kernel, int threads) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (size_t)sm_count() * per_sm;
}

// ------------------------------------------------------------------------------------------------
// YCbCr -> RGB(A).  Per pixel (convert.cpp:480-487), with sCb = Cb - 128, sCr = Cr - 128 as int8:
//   B = sat8(Y + ((sCb * 29049                + 8192) >> 14))
//   G = sat8(Y + ((sCb * -5636 + sCr * -11698 + 8192) >> 14))
//   R = sat8(Y + ((sCr * 22987                + 8192) >> 14))
// Y * 16384 + 8192 goes into the accumulator of the dot product: (x + Y 2^14) >> 14 == (x >> 14) + Y for the arithmetic
// shift, so each channel is IDP.2A + SHF, and the saturation happens in the pack.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int pack_sat_u8x4(int b0, int b1, int b2, int b3) {
  unsigned int hi, w;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(b3), "r"(b2), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(b1), "r"(b0), "r"(hi));
  return w;
}

struct Rgb4 {
  int r[4], g[4], b[4];
};

// four pixels from one word of each plane
__device__ __forceinline__ void ycc4(unsigned int yw, unsigned int cbw, unsigned int crw, Rgb4 &o) {
  constexpr int kOne = 16384, kB = 29049, kR = 22987, kGb = -5636, kGr = -11698;
  constexpr unsigned int sel_lo = (unsigned)kOne, sel_hi = (unsigned)kOne << 16;       // (16384, 0) / (0, 16384) as u16 x 2
  constexpr int b_lo = kB, b_hi = (int)((unsigned)kB << 16), r_lo = kR, r_hi = (int)((unsigned)kR << 16);
  constexpr int g_pair = (int)(((unsigned)kGr << 16) | ((unsigned)kGb & 0xFFFFu));     // (kGb, kGr) as s16 x 2
  const int scb = (int)(cbw ^ 0x80808080u), scr = (int)(crw ^ 0x80808080u);            // u8 - 128 as s8, four at once
  const int g01 = (int)__byte_perm((unsigned)scb, (unsigned)scr, 0x5140), g23 = (int)__byte_perm((unsigned)scb, (unsigned)scr, 0x7362);
  int yc[4];
  yc[0] = (int)__dp2a_lo(sel_lo, yw, 8192u), yc[1] = (int)__dp2a_lo(sel_hi, yw, 8192u);
  yc[2] = (int)__dp2a_hi(sel_lo, yw, 8192u), yc[3] = (int)__dp2a_hi(sel_hi, yw, 8192u);
  o.b[0] = __dp2a_lo(b_lo, scb, yc[0]) >> 14, o.b[1] = __dp2a_lo(b_hi, scb, yc[1]) >> 14;
  o.b[2] = __dp2a_hi(b_lo, scb, yc[2]) >> 14, o.b[3] = __dp2a_hi(b_hi, scb, yc[3]) >> 14;
  o.r[0] = __dp2a_lo(r_lo, scr, yc[0]) >> 14, o.r[1] = __dp2a_lo(r_hi, scr, yc[1]) >> 14;
  o.r[2] = __dp2a_hi(r_lo, scr, yc[2]) >> 14, o.r[3] = __dp2a_hi(r_hi, scr, yc[3]) >> 14;
  o.g[0] = __dp2a_lo(g_pair, g01, yc[0]) >> 14, o.g[1] = __dp2a_hi(g_pair, g01, yc[1]) >> 14;
  o.g[2] = __dp2a_lo(g_pair, g23, yc[2]) >> 14, o.g[3] = __dp2a_hi(g_pair, g23, yc[3]) >> 14;
}

// V pixels per work item (1: byte path for any shape; 4: 32-bit accesses; 16: 128-bit accesses), CH = 3 or 4.
// The output is dense, so the 32 items of a warp own one contiguous run of bytes; a lane's own 12 / 16 / 48 / 64 bytes
// would be a strided store (every 32-byte sector written in halves by different instructions).  STAGED: the warp
// transposes through shared memory (conflict-free slots, XOR-swizzled when a lane owns four units) and every store
// instruction writes 128 / 512 contiguous bytes.  (Measured on B200, 8192 frames of 640x480: direct 5.6 TB/s,
// staged -- see DESIGN.md section 11.)
template <int V, int CH, bool STAGED>
__global__ void __launch_bounds__(256)
ycbcr_to_rgb_kernel(const uint8_t *__restrict__ y, int yrs, size_t yfs, const uint8_t *__restrict__ cb, const uint8_t *__restrict__ cr,
                    int crs, size_t cfs, int w, int h, size_t n, uint8_t *__restrict__ dst) {
  constexpr int WPT = V >= 4 ? V / 4 * CH : 1;              // output words per item
  constexpr int UW = V == 16 ? 4 : 1, UPT = WPT / UW;       // words per store unit (uint4 / u32), units per item: 3 or 4
  __shared__ unsigned int s_stage[STAGED ? 8 * 32 * WPT : 1];
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int wv = w / V;
  // (frame, row, column group) of this thread's item, advanced by the grid size without dividing again
  const size_t per_frame = (size_t)h * wv, total = n * per_frame;
  size_t f = tid / per_frame;
  int row = (int)((tid - f * per_frame) / wv), xv = (int)(tid - f * per_frame - (size_t)row * wv);
  const size_t step_f = nthreads / per_frame;
  const int step_row = (int)((nthreads - step_f * per_frame) / wv), step_x = (int)(nthreads - step_f * per_frame - (size_t)step_row * wv);
  for (size_t i = tid; i - lane < total; i += nthreads, f += step_f, row += step_row, xv += step_x) {  // warp-uniform trip count
    if (xv >= wv) xv -= wv, row++;
    if (row >= h) row -= h, f++;
    const bool live = i < total;
    const int x = xv * V;
    const uint8_t *py = y + f * yfs + (size_t)row * yrs + x, *pb = cb + f * cfs + (size_t)row * crs + x, *pr = cr + f * cfs + (size_t)row * crs + x;
    uint8_t *o = dst + i * (size_t)(V * CH);  // dense output: item i starts at byte i V CH
    if constexpr (V == 1) {
      if (live) {
        Rgb4 p;
        ycc4(*py, *pb, *pr, p);
        const unsigned int px = pack_sat_u8x4(p.r[0], p.g[0], p.b[0], 255);
        o[0] = (uint8_t)px, o[1] = (uint8_t)(px >> 8), o[2] = (uint8_t)(px >> 16);
        if (CH == 4) o[3] = 0xff;
      }
    } else {
      unsigned int yw[V / 4], bw[V / 4], rw[V / 4], ow[WPT];
      if (live) {
        if constexpr (V == 16) {
          const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(py)), b = __ldcs(reinterpret_cast<const uint4 *>(pb)),
                      c = __ldcs(reinterpret_cast<const uint4 *>(pr));
          yw[0] = a.x, yw[1] = a.y, yw[2] = a.z, yw[3] = a.w;
          bw[0] = b.x, bw[1] = b.y, bw[2] = b.z, bw[3] = b.w;
          rw[0] = c.x, rw[1] = c.y, rw[2] = c.z, rw[3] = c.w;
        } else {
          yw[0] = __ldcs(reinterpret_cast<const unsigned int *>(py)), bw[0] = __ldcs(reinterpret_cast<const unsigned int *>(pb));
          rw[0] = __ldcs(reinterpret_cast<const unsigned int *>(pr));
        }
#pragma unroll
        for (int q = 0; q < V / 4; q++) {
          Rgb4 p;
          ycc4(yw[q], bw[q], rw[q], p);
          if (CH == 3) {  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            ow[3 * q + 0] = pack_sat_u8x4(p.r[0], p.g[0], p.b[0], p.r[1]);
            ow[3 * q + 1] = pack_sat_u8x4(p.g[1], p.b[1], p.r[2], p.g[2]);
            ow[3 * q + 2] = pack_sat_u8x4(p.b[2], p.r[3], p.g[3], p.b[3]);
          } else {
#pragma unroll
            for (int k = 0; k < 4; k++) ow[4 * q + k] = pack_sat_u8x4(p.r[k], p.g[k], p.b[k], 255);
          }
        }
      }
      if constexpr (!STAGED) {
        if (live) {
          if constexpr (V == 16 || CH == 4) {
#pragma unroll
            for (int k = 0; k < WPT / 4; k++)
              __stcs(reinterpret_cast<uint4 *>(o) + k, make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]));
          } else {
#pragma unroll
            for (int k = 0; k < 3; k++) __stcs(reinterpret_cast<unsigned int *>(o) + k, ow[k]);
          }
        }
      } else {
        // unit k of lane L lives in slot L UPT + (k ^ swizzle(L)); the swizzle only matters for UPT == 4 (strides of 3 are
        // conflict-free as they are): 128-bit units are served a quarter warp at a time, 32-bit units a warp at a time
        unsigned int *stage = s_stage + (threadIdx.x >> 5) * (32 * WPT);
        auto slot = [](int L, int k) { return L * UPT + (UPT == 4 ? (k ^ ((L >> (UW == 4 ? 1 : 3)) & 3)) : k); };
#pragma unroll
        for (int k = 0; k < UPT; k++) {
          if constexpr (UW == 4) reinterpret_cast<uint4 *>(stage)[slot(lane, k)] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
          else stage[slot(lane, k)] = ow[k];
        }
        __syncwarp();
        const size_t ubase = (i - lane) * UPT, utotal = total * UPT;  // first unit of the warp's run, units in the whole output
#pragma unroll
        for (int j = 0; j < UPT; j++) {
          const int g = j * 32 + lane, L = g / UPT, k = g - L * UPT;
          if (ubase + g < utotal) {
            if constexpr (UW == 4) __stcs(reinterpret_cast<uint4 *>(dst) + ubase + g, reinterpret_cast<const uint4 *>(stage)[slot(L, k)]);
            else __stcs(reinterpret_cast<unsigned int *>(dst) + ubase + g, stage[slot(L, k)]);
          }
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// RGBA -> R: dest[i] = source[4 i].  A warp takes 2 KB of source per step: lane l reads 16 bytes at 16 l + 512 j
// (j = 0..3, coalesced 128-bit loads) and writes the four R bytes at 4 l + 128 j (coalesced 32-bit stores).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rgba_to_r_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n_px, int vec_ok) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (vec_ok) {  // src 16-byte aligned, dst 4-byte aligned
    const size_t warps = nthreads >> 5, warp = tid >> 5, n_steps = n_px >> 9;  // 512 pixels per warp step
    const int lane = (int)(tid & 31);
    for (size_t s = warp; s < n_steps; s += warps) {
      const uint4 *p = reinterpret_cast<const uint4 *>(src + (s << 11)) + lane;
      unsigned int *q = reinterpret_cast<unsigned int *>(dst + (s << 9)) + lane;
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = __ldcs(p + 32 * j);
#pragma unroll
      for (int j = 0; j < 4; j++)
        __stcs(q + 32 * j, __byte_perm(__byte_perm(v[j].x, v[j].y, 0x0040), __byte_perm(v[j].z, v[j].w, 0x0040), 0x5410));
    }
    done = n_steps << 9;
  }
  for (size_t i = done + tid; i < n_px; i += nthreads) dst[i] = src[4 * i];
}

// ------------------------------------------------------------------------------------------------
// 3 x 3 stencils, rows and columns clamped at the image.  A CTA stages a (64 + 2) x (128 + 8) byte tile in shared memory
// (clamping happens there, once), then every thread produces four pixels a row for eight rows, sliding a three-row
// window of the words left of / at / right of its four columns.
//   KIND 0  t = |p(x+1) - p(x-1)| in every row;      out = 3 (t(y-1) + t(y+1)) + 10 t(y)
//   KIND 1  t = |p(y+1) - p(y-1)| in every column;   out = 3 (t(x-1) + t(x+1)) + 10 t(x)
//   KIND 2  out = p(x-1,y-1) - p(x+1,y-1) - p(x-1,y+1) + p(x+1,y+1)
// t <= 255 and out <= 4080: two 16-bit lanes per register never carry into each other.  KIND 2 is signed (+-510): the
// difference is taken with a bias of 1024 per lane, removed when the lanes are stored.
// ------------------------------------------------------------------------------------------------
constexpr int kTileW = 128, kTileH = 64, kStThreads = 256, kPitchW = kTileW / 4 + 2;  // words per staged row

__device__ __forceinline__ unsigned int lo_pair(unsigned int t) { return __byte_perm(t, 0u, 0x4140); }  // (t0, t1) as u16 x 2
__device__ __forceinline__ unsigned int hi_pair(unsigned int t) { return __byte_perm(t, 0u, 0x4342); }  // (t2, t3)

// Eight output rows x four pixels of one thread, from a sliding three-row window of the staged tile.  FAST: every row and
// all four pixels exist and the 8-byte store is aligned -- a straight-line body (the generic one carries per-row and per-pixel
// tests; keeping them out of this path halves the instructions of the kernel's hot loop).
template <int KIND, bool FAST>
__device__ __forceinline__ void stencil_band(const unsigned int *__restrict__ tile, int lane, int r0, int16_t *__restrict__ o, int w, int x,
                                             int rows_live) {
  // window of staged rows: index 0 = the row above the output row, 1 = the output row, 2 = the row below
  unsigned int a[3], b[3], c[3];  // KIND 0: t pairs (lo, hi) are kept in a / b;  KIND 1, 2: the words left / at / right
  auto prepare = [&](int r, int slot) {
    const unsigned int *p = tile + r * kPitchW + lane;
    const unsigned int L = p[0], C = p[1], R = p[2];
    if (KIND == 0) {
      const unsigned int tt = __vabsdiffu4(__byte_perm(C, R, 0x4321), __byte_perm(L, C, 0x6543));  // |p(x+1) - p(x-1)| x 4
      a[slot] = lo_pair(tt), b[slot] = hi_pair(tt);
    } else {
      a[slot] = L, b[slot] = C, c[slot] = R;
    }
  };
  const unsigned int row_bytes = 2u * (unsigned int)w;
  prepare(r0, 0);
  prepare(r0 + 1, 1);
#pragma unroll
  for (int rr = 0; rr < 8; rr++) {
    prepare(r0 + rr + 2, 2);
    unsigned int o_lo, o_hi;
    if (KIND == 0) {
      o_lo = (a[0] + a[2]) * 3u + a[1] * 10u, o_hi = (b[0] + b[2]) * 3u + b[1] * 10u;
    } else if (KIND == 1) {
      const unsigned int tL = __vabsdiffu4(a[2], a[0]), tC = __vabsdiffu4(b[2], b[0]), tR = __vabsdiffu4(c[2], c[0]);
      const unsigned int tl = __byte_perm(tL, tC, 0x6543), tr = __byte_perm(tC, tR, 0x4321);  // t(x-1), t(x+1)
      o_lo = (lo_pair(tl) + lo_pair(tr)) * 3u + lo_pair(tC) * 10u, o_hi = (hi_pair(tl) + hi_pair(tr)) * 3u + hi_pair(tC) * 10u;
    } else {
      const unsigned int ul = __byte_perm(a[0], b[0], 0x6543), ur = __byte_perm(b[0], c[0], 0x4321);  // row above: p(x-1), p(x+1)
      const unsigned int dl = __byte_perm(a[2], b[2], 0x6543), dr = __byte_perm(b[2], c[2], 0x4321);  // row below
      o_lo = lo_pair(ul) + lo_pair(dr) + 0x04000400u - lo_pair(ur) - lo_pair(dl);
      o_hi = hi_pair(ul) + hi_pair(dr) + 0x04000400u - hi_pair(ur) - hi_pair(dl);
      const int v0 = (int)(o_lo & 0xFFFFu) - 1024, v1 = (int)(o_lo >> 16) - 1024, v2 = (int)(o_hi & 0xFFFFu) - 1024, v3 = (int)(o_hi >> 16) - 1024;
      o_lo = __byte_perm((unsigned)v0, (unsigned)v1, 0x5410), o_hi = __byte_perm((unsigned)v2, (unsigned)v3, 0x5410);
    }
    int16_t *orow = reinterpret_cast<int16_t *>(reinterpret_cast<char *>(o) + (unsigned int)rr * row_bytes);  // (one IMAD.WIDE per row)
    if (FAST) {
      __stcs(reinterpret_cast<uint2 *>(orow), make_uint2(o_lo, o_hi));
    } else if (rr < rows_live) {
      const unsigned int ww[2] = {o_lo, o_hi};
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (x + j < w) orow[j] = (int16_t)(ww[j >> 1] >> (16 * (j & 1)));
    }
    a[0] = a[1], a[1] = a[2], b[0] = b[1], b[1] = b[2];
    if (KIND != 0) c[0] = c[1], c[1] = c[2];
  }
}

template <int KIND>
__global__ void __launch_bounds__(kStThreads)
stencil3_kernel(const uint8_t *__restrict__ src, int row_stride, size_t frame_stride, int w, int h, int tiles_x, int tiles_y,
                unsigned int n_tiles, int16_t *__restrict__ out, int word_ok) {
  // two tile buffers: staging tile k + 1 may start while other warps still compute tile k, so one barrier per tile is enough
  // (the barrier after staging k + 1 is only passed once every warp has finished computing k - 1, the buffer's last user)
  __shared__ unsigned int s_tile[2][(kTileH + 2) * kPitchW];
  constexpr int kRowsPerWarp = (kTileH + 2 + 7) / 8;  // staging: warp wid takes tile rows wid, wid + 8, ..: lane = interior word
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool store8 = (w % 4 == 0) && ((uintptr_t)out % 8 == 0);
  // tile walk: (frame, tile row, tile column) advanced by the grid size without dividing again
  struct Tile {
    unsigned int f;
    int ty, tx;
  };
  const unsigned int tiles_per_frame = (unsigned)(tiles_x * tiles_y);
  Tile step;
  step.f = gridDim.x / tiles_per_frame;
  step.ty = (int)((gridDim.x - step.f * tiles_per_frame) / (unsigned)tiles_x);
  step.tx = (int)(gridDim.x - step.f * tiles_per_frame) - step.ty * tiles_x;
  auto advance = [&](Tile T) {
    T.f += step.f, T.ty += step.ty, T.tx += step.tx;
    if (T.tx >= tiles_x) T.tx -= tiles_x, T.ty++;
    if (T.ty >= tiles_y) T.ty -= tiles_y, T.f++;
    return T;
  };
  // The next tile's words are fetched into registers while the current one is computed from shared memory.
  // Interior word (r, lane + 1) = image columns x0 + 4 lane .. + 3 of row clamp(y0 - 1 + r); the two halo words of a row
  // only ever contribute one byte each: p(max(x0 - 1, 0)) as byte 3 of word 0 and p(min(x0 + 128, w - 1)) as byte 0 of word 33.
  unsigned int regs[kRowsPerWarp], halo = 0;
  auto fetch = [&](const Tile &T) {
    const uint8_t *img = src + (size_t)T.f * frame_stride;
    const int x0 = T.tx * kTileW, y0 = T.ty * kTileH, gx = x0 + 4 * lane;
    if (word_ok && y0 > 0 && y0 + kTileH < h && x0 + kTileW <= w) {  // no clamping anywhere inside the tile (block-uniform)
      const uint8_t *p = img + (size_t)(y0 - 1 + wid) * row_stride + gx;
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; j++, p += 8 * (size_t)row_stride)
        regs[j] = wid + 8 * j < kTileH + 2 ? __ldg(reinterpret_cast<const unsigned int *>(p)) : 0u;
    } else {
#pragma unroll
      for (int j = 0; j < kRowsPerWarp; j++) {
        const int r = wid + 8 * j;
        unsigned int v = 0;
        if (r < kTileH + 2 && gx <= w) {  // (a word that starts right of column w is never read by a live output)
          int gy = y0 - 1 + r;
          gy = gy < 0 ? 0 : (gy > h - 1 ? h - 1 : gy);
          const uint8_t *row = img + (size_t)gy * row_stride;
          if (word_ok && gx + 3 < w) {
            v = __ldg(reinterpret_cast<const unsigned int *>(row + gx));
          } else {
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (unsigned int)__ldg(row + (gx + b > w - 1 ? w - 1 : gx + b)) << (8 * b);
          }
        }
        regs[j] = v;
      }
    }
    if (threadIdx.x < 2 * (kTileH + 2)) {
      const int r = threadIdx.x >> 1, right = threadIdx.x & 1;
      int gy = y0 - 1 + r;
      gy = gy < 0 ? 0 : (gy > h - 1 ? h - 1 : gy);
      int c = right ? x0 + kTileW : x0 - 1;
      c = c < 0 ? 0 : (c > w - 1 ? w - 1 : c);
      halo = __ldg(img + (size_t)gy * row_stride + c);
    }
  };
  auto stage = [&](unsigned int *tile) {
#pragma unroll
    for (int j = 0; j < kRowsPerWarp; j++)
      if (wid + 8 * j < kTileH + 2) tile[(wid + 8 * j) * kPitchW + lane + 1] = regs[j];
    if (threadIdx.x < 2 * (kTileH + 2)) {
      const int r = threadIdx.x >> 1, right = threadIdx.x & 1;
      tile[r * kPitchW + (right ? kPitchW - 1 : 0)] = right ? halo : halo << 24;
    }
  };
  unsigned int t = blockIdx.x;
  Tile cur;
  cur.f = t / tiles_per_frame;
  cur.ty = (int)((t - cur.f * tiles_per_frame) / (unsigned)tiles_x), cur.tx = (int)(t - cur.f * tiles_per_frame) - cur.ty * tiles_x;
  if (t < n_tiles) fetch(cur);
  for (int buf = 0; t < n_tiles; t += gridDim.x, buf ^= 1) {
    unsigned int *tile = s_tile[buf];
    stage(tile);
    __syncthreads();
    const Tile nxt = advance(cur);
    if (t + gridDim.x < n_tiles) fetch(nxt);
    const int x0 = cur.tx * kTileW, y0 = cur.ty * kTileH, x = x0 + 4 * lane;
    const int r0 = wid * 8;                    // first output row of this warp inside the tile
    const int rows_live = h - (y0 + r0);       // rows of this warp that exist in the image (>= 8: all of them)
    if (x < w && rows_live > 0) {
      int16_t *o = out + ((size_t)cur.f * h + (y0 + r0)) * (size_t)w + x;
      if (store8 && x + 3 < w && rows_live >= 8) stencil_band<KIND, true>(tile, lane, r0, o, w, x, rows_live);
      else stencil_band<KIND, false>(tile, lane, r0, o, w, x, rows_live);
    }
    cur = nxt;
  }
}

}  // namespace

int launch_ycbcr_to_rgb(const uint8_t *y, int yrs, size_t yfs, const uint8_t *cb, const uint8_t *cr, int crs, size_t cfs, int w, int h,
                        int n, int channels, uint8_t *dst, cudaStream_t s) {
  auto aligned = [&](int a) {
    return w % a == 0 && yrs % a == 0 && crs % a == 0 && yfs % a == 0 && cfs % a == 0 && (uintptr_t)y % a == 0 && (uintptr_t)cb % a == 0 &&
           (uintptr_t)cr % a == 0 && (uintptr_t)dst % a == 0;
  };
  static const bool direct = getenv("B200_DMZ_FORMATS_DIRECT") != nullptr;  // experiment switch: per-lane strided stores everywhere
#define YCC(V, CH, ST, Y, CB, CR, YRS, YFS, CRS, CFS, W_, H_, N_, DST)                                                         \
  do {                                                                                                                       \
    const size_t items_ = (size_t)(N_) * (H_) * ((W_) / V);                                                                  \
    size_t blocks_ = (items_ + 255) / 256;                                                                                   \
    const size_t cap_ = resident_ctas(ycbcr_to_rgb_kernel<V, CH, ST>, 256);                                                  \
    blocks_ = blocks_ > cap_ ? cap_ : (blocks_ < 1 ? 1 : blocks_);                                                           \
    ycbcr_to_rgb_kernel<V, CH, ST><<<(unsigned)blocks_, 256, 0, s>>>(Y, YRS, YFS, CB, CR, CRS, CFS, W_, H_, (size_t)(N_), DST); \
    launches++;                                                                                                              \
  } while (0)
#define YCC_CH(V, ST, ...)                                                  \
  do {                                                                      \
    if (channels == 3) YCC(V, 3, ST, __VA_ARGS__); else YCC(V, 4, ST, __VA_ARGS__); \
  } while (0)
  int launches = 0;
  const size_t plane = (size_t)w * h, total = plane * n;
  const bool dense = yrs == w && crs == w && (n == 1 || (yfs == plane && cfs == plane));
  const bool base16 = (uintptr_t)y % 16 == 0 && (uintptr_t)cb % 16 == 0 && (uintptr_t)cr % 16 == 0 && (uintptr_t)dst % 16 == 0;
  if (dense && base16 && total >= 16) {
    // Dense planes are one flat run of pixels whatever the width (the arithmetic does not look at coordinates): 16-pixel
    // items over "rows" of 16 pixels, in slabs whose row count fits an int; the last total % 16 pixels go through the byte path.
    const size_t main_px = total & ~(size_t)15, slab = (size_t)1 << 34;
    for (size_t p0 = 0; p0 < main_px; p0 += slab) {
      const size_t cnt = main_px - p0 < slab ? main_px - p0 : slab;
      if (direct) YCC_CH(16, false, y + p0, cb + p0, cr + p0, 16, cnt, 16, cnt, 16, (int)(cnt / 16), 1, dst + p0 * channels);
      else YCC_CH(16, true, y + p0, cb + p0, cr + p0, 16, cnt, 16, cnt, 16, (int)(cnt / 16), 1, dst + p0 * channels);
    }
    if (total > main_px) {
      const int tail = (int)(total - main_px);
      YCC_CH(1, false, y + main_px, cb + main_px, cr + main_px, tail, (size_t)tail, tail, (size_t)tail, tail, 1, 1, dst + main_px * channels);
    }
  } else if (aligned(16)) {
    if (direct) YCC_CH(16, false, y, cb, cr, yrs, yfs, crs, cfs, w, h, n, dst);
    else YCC_CH(16, true, y, cb, cr, yrs, yfs, crs, cfs, w, h, n, dst);
  } else if (aligned(4)) {
    // four pixels per item: the transposed stores cost more than they save here (measured: 3.4 vs 4.8 TB/s on 428-wide cards)
    if (channels == 3 || (uintptr_t)dst % 16 == 0) YCC_CH(4, false, y, cb, cr, yrs, yfs, crs, cfs, w, h, n, dst);
    else YCC_CH(4, true, y, cb, cr, yrs, yfs, crs, cfs, w, h, n, dst);
  } else {
    YCC_CH(1, false, y, cb, cr, yrs, yfs, crs, cfs, w, h, n, dst);
  }
#undef YCC_CH
#undef YCC
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int launch_rgba_to_r(const uint8_t *src, size_t n_px, uint8_t *dst, cudaStream_t s) {
  const int vec_ok = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 4 == 0);
  size_t blocks = ((n_px >> 4) + 255) / 256;  // a thread moves 16 pixels per step
  const size_t cap = resident_ctas(rgba_to_r_kernel, 256);
  blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
  rgba_to_r_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dst, n_px, vec_ok);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_stencil3(const uint8_t *src, int row_stride, size_t frame_stride, int w, int h, int n, int kind, int16_t *out, cudaStream_t s) {
  const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
  const int word_ok = ((uintptr_t)src % 4 == 0) && (row_stride % 4 == 0) && (frame_stride % 4 == 0);
  const size_t cap = kind == 0 ? resident_ctas(stencil3_kernel<0>, kStThreads) : kind == 1 ? resident_ctas(stencil3_kernel<1>, kStThreads)
                                                                                              : resident_ctas(stencil3_kernel<2>, kStThreads);
  const size_t per_frame = (size_t)tiles_x * tiles_y;
  const size_t max_frames = ((size_t)1 << 30) / per_frame > 0 ? ((size_t)1 << 30) / per_frame : 1;  // tile indices stay 32-bit
  int launches = 0;
  for (size_t f0 = 0; f0 < (size_t)n; f0 += max_frames, launches++) {
    const size_t cnt = (size_t)n - f0 < max_frames ? (size_t)n - f0 : max_frames;
    const unsigned int n_tiles = (unsigned int)(cnt * per_frame);
    const unsigned blocks = (unsigned)(n_tiles > cap ? cap : n_tiles);
    const uint8_t *p = src + f0 * frame_stride;
    int16_t *o = out + f0 * (size_t)w * h;
    if (kind == 0) stencil3_kernel<0><<<blocks, kStThreads, 0, s>>>(p, row_stride, frame_stride, w, h, tiles_x, tiles_y, n_tiles, o, word_ok);
    else if (kind == 1) stencil3_kernel<1><<<blocks, kStThreads, 0, s>>>(p, row_stride, frame_stride, w, h, tiles_x, tiles_y, n_tiles, o, word_ok);
    else stencil3_kernel<2><<<blocks, kStThreads, 0, s>>>(p, row_stride, frame_stride, w, h, tiles_x, tiles_y, n_tiles, o, word_ok);
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}
