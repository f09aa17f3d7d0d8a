// Border following on a bit plane: the arithmetic core of octseg_contour_largest (csrc/contour.cu), written as
// host+device code so tests/contour_core_harness.cpp can run the very same functions on the CPU against cv2.
//
// cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) as used by calculate_thickness_contour
// (src/app/tools/analysis.py:21-57) is Suzuki-Abe border following (OpenCV 4.8.1 contours.cpp, icvFetchContour;
// third-party, restated from its published algorithm):
//   * an outer border starts at the raster-first pixel of an 8-connected component; from there the first non-zero
//     neighbour is searched CLOCKWISE starting after West, then each step searches COUNTER-CLOCKWISE starting
//     after the direction it arrived from; the walk ends when it re-enters the start pixel from the first
//     neighbour found;
//   * CHAIN_APPROX_SIMPLE keeps a point only when the step direction leaving it differs from the one before.
// A walk started at any other "tip" (left and the three upper neighbours empty) meets a raster-earlier pixel and
// is abandoned there, so completed walks are exactly the outer borders, one per component.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define OCTSEG_HD __host__ __device__ __forceinline__
#else
#define OCTSEG_HD inline
#endif

namespace octseg {

// plane: (H + 2) rows x pitch words (+ 1 spare word at the end), pitch = plane_pitch(W); pixel (x, y) at bit (x + 32) of row (y + 1): word 0 of
// a row, rows 0 and H + 1 and the bits from W on are zero, so probes one pixel outside the image need no bounds test
// and a word holds 32 pixels starting at a multiple of 32 (16-byte mask loads pack without crossing words)
OCTSEG_HD int plane_pitch(int W) { return (W + 1 + 31) / 32 + 1; }
OCTSEG_HD uint32_t plane_get(const uint32_t* pl, int pitch, int x, int y) {
  return (pl[(y + 1) * pitch + ((x + 32) >> 5)] >> (x & 31)) & 1u;
}

OCTSEG_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs(static_cast<int>(v)) - 1;
#else
  return __builtin_ctz(v);
#endif
}

// bits x-1, x, x+1 of plane row `row` (bit 0 = x-1); the word after the row's last is the next row's zero word 0
OCTSEG_HD uint32_t row3(const uint32_t* pl, int pitch, int row, int x) {
  const int p = x + 31;  // plane bit of x - 1
  const uint32_t* w = pl + row * pitch + (p >> 5);
  const unsigned long long v = (static_cast<unsigned long long>(w[1]) << 32) | w[0];
  return static_cast<uint32_t>(v >> (p & 31)) & 7u;
}

// bit d = the neighbour of (x, y) in cv2 direction d (0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE) is set
OCTSEG_HD uint32_t neighbours8(const uint32_t* pl, int pitch, int x, int y) {
  const uint32_t up = row3(pl, pitch, y, x), mid = row3(pl, pitch, y + 1, x), dn = row3(pl, pitch, y + 2, x);
  return (mid >> 2) | ((up >> 2) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((mid & 1u) << 4) | ((dn & 1u) << 5) |
         (((dn >> 1) & 1u) << 6) | ((dn >> 2) << 7);
}

struct ContourSums {
  long long a00, a10, a01;  // cv2 contourMoments' integer accumulators (m00 = a00/2, m10 = a10/6, m01 = a01/6 up to sign)
  int nverts;
};

// direction codes of cv2 (y grows downwards): 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE
OCTSEG_HD void code_delta(int s, int& dx, int& dy) {
  // dx: {1, 1, 0, -1, -1, -1, 0, 1}, dy: {0, -1, -1, -1, 0, 1, 1, 1}; packed 2 bits per entry (value + 1)
  dx = static_cast<int>((0x901Au >> (2 * s)) & 3u) - 1;
  dy = static_cast<int>((0xA901u >> (2 * s)) & 3u) - 1;
}

// Walks the border that starts at (x0, y0) (a pixel whose West neighbour is empty).  Returns kWalkDone for a
// completed outer border, kWalkNotFirst when the walk meets a pixel that precedes the start in raster order (the
// start is not an outer border's first pixel), kWalkBudget when max_steps ran out first.
// EMIT: also write the kept points (x, y as int16 pairs, at most cap of them; sums.nverts counts all) and a10/a01.
enum WalkResult { kWalkNotFirst = 0, kWalkDone = 1, kWalkBudget = 2 };
template <bool EMIT>
OCTSEG_HD WalkResult trace_border(const uint32_t* pl, int pitch, int x0, int y0, ContourSums& sums, int16_t* verts, int cap,
                            long long max_steps) {
  sums.a00 = sums.a10 = sums.a01 = 0;
  sums.nverts = 0;
  int s = 4, dx, dy;
  do {
    s = (s - 1) & 7;
    code_delta(s, dx, dy);
  } while (!plane_get(pl, pitch, x0 + dx, y0 + dy) && s != 4);
  if (s == 4) {  // single-pixel component
    sums.nverts = 1;
    if (EMIT && cap > 0) verts[0] = static_cast<int16_t>(x0), verts[1] = static_cast<int16_t>(y0);
    return kWalkDone;
  }
  const int x1 = x0 + dx, y1 = y0 + dy;
  int x3 = x0, y3 = y0, prev_s = s ^ 4;
  int fx = 0, fy = 0, px = 0, py = 0;  // first and previous kept point
  for (long long step = 0; step < max_steps; ++step) {
    // the 8 neighbours of (x3, y3) as one byte (bit d = direction d), then the first one counter-clockwise after s
    const uint32_t nb = neighbours8(pl, pitch, x3, y3);
    const uint32_t rot = ((nb | (nb << 8)) >> ((s + 1) & 7)) & 0xffu;  // never 0: the pixel we came from is set
    s = s + 1 + ctz32(rot);
    code_delta(s & 7, dx, dy);
    const int x4 = x3 + dx, y4 = y3 + dy;
    s &= 7;
    if (s != prev_s) {
      if (sums.nverts == 0) {
        fx = x3, fy = y3;
      } else {
        const long long dxy = static_cast<long long>(px) * y3 - static_cast<long long>(x3) * py;
        sums.a00 += dxy;
        if (EMIT) sums.a10 += dxy * (px + x3), sums.a01 += dxy * (py + y3);
      }
      if (EMIT && sums.nverts < cap) verts[2 * sums.nverts] = static_cast<int16_t>(x3), verts[2 * sums.nverts + 1] = static_cast<int16_t>(y3);
      ++sums.nverts;
      px = x3, py = y3;
      prev_s = s;
    }
    if (y4 < y0 || (y4 == y0 && x4 < x0)) return kWalkNotFirst;
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) {
      const long long dxy = static_cast<long long>(px) * fy - static_cast<long long>(fx) * py;  // closing edge last -> first
      sums.a00 += dxy;
      if (EMIT) sums.a10 += dxy * (px + fx), sums.a01 += dxy * (py + fy);
      return kWalkDone;
    }
    x3 = x4, y3 = y4;
    s = (s + 4) & 7;
  }
  return kWalkBudget;
}

// candidate start bits of one plane word: pixel set, West empty, the three pixels above empty.  `cur` and `touch`
// (bit x = some pixel of the row above in columns x-1..x+1 is set) feed tip_run_touches.
OCTSEG_HD uint32_t tip_bits(const uint32_t* pl, int pitch, int row /* plane row = y + 1 */, int k, uint32_t& cur, uint32_t& touch) {
  const uint32_t* r = pl + row * pitch + k;
  const uint32_t* u = r - pitch;
  cur = r[0];
  const uint32_t west = (cur << 1) | (k > 0 ? r[-1] >> 31 : 0u);
  const uint32_t up = u[0], upw = (up << 1) | (k > 0 ? u[-1] >> 31 : 0u), upe = (up >> 1) | (k < pitch - 1 ? u[1] << 31 : 0u);
  touch = up | upw | upe;
  return cur & ~west & ~touch;
}

// The horizontal run of set pixels that starts at tip bit b (as far as it lies in this word) touches the row above:
// the tip is 8-connected to a raster-earlier pixel, so it is not its component's first pixel and needs no walk.
// (Sound filter, not a complete one: tips joined to earlier pixels only through lower rows still walk and stop at
// the first earlier pixel they meet.)  On a disc it removes every tip of the upper-left arc but the top one.
OCTSEG_HD bool tip_run_touches(uint32_t cur, uint32_t touch, int b) {
  const uint32_t run = ((cur + (1u << b)) ^ cur) & cur;
  return (run & touch) != 0u;
}

}  // namespace octseg
