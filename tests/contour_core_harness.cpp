// CPU harness around the product's border-following core (oct_segmentation_b200/csrc/contour_core.h), built by
// tests/test_contour_core_cpu.py with g++ and driven through ctypes: the same functions the CUDA kernel runs,
// checked against cv2.findContours without a GPU.  Mirrors the kernel's two passes (csrc/contour.cu).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../oct_segmentation_b200/csrc/contour_core.h"

using namespace octseg;

extern "C" int contour_largest_cpu(const uint8_t* mask, int H, int W, long long* sums4, int* nverts, int16_t* verts, int cap,
                                   int* n_outer) {
  const int pitch = plane_pitch(W);
  std::vector<uint32_t> pl(static_cast<size_t>(H + 2) * pitch + 1, 0u);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (mask[static_cast<size_t>(y) * W + x]) pl[(y + 1) * pitch + ((x + 32) >> 5)] |= 1u << (x & 31);
  unsigned long long best = 0;
  bool have = false;
  *n_outer = 0;
  for (int y = 0; y < H; ++y)
    for (int k = 0; k < pitch; ++k) {
      uint32_t cur, touch;
      uint32_t tips = tip_bits(pl.data(), pitch, y + 1, k, cur, touch);
      while (tips) {
        const int b = __builtin_ctz(tips);
        tips &= tips - 1;
        if (tip_run_touches(cur, touch, b)) continue;
        const int x = 32 * k + b - 32;
        ContourSums s;
        if (trace_border<false>(pl.data(), pitch, x, y, s, nullptr, 0, 4LL * H * W + 16) != kWalkDone) continue;
        ++*n_outer;
        const unsigned long long area = static_cast<unsigned long long>(s.a00 < 0 ? -s.a00 : s.a00);
        const unsigned long long key = (area << 32) | static_cast<unsigned>(y * W + x);
        if (area > 0 && (!have || key > best)) best = key, have = true;
      }
    }
  sums4[0] = sums4[1] = sums4[2] = 0;
  sums4[3] = -1;
  *nverts = 0;
  if (!have) return 0;
  const int start = static_cast<int>(best & 0xffffffffu);
  ContourSums s;
  trace_border<true>(pl.data(), pitch, start % W, start / W, s, verts, cap, 4LL * H * W + 16);
  sums4[0] = s.a00, sums4[1] = s.a10, sums4[2] = s.a01, sums4[3] = start;
  *nverts = s.nverts;
  return 0;
}
