// calculate_thickness_contour (src/app/tools/analysis.py:21-57) on the GPU: per (frame, class) the largest outer
// border of cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), bit-exact (same points in the same order),
// with the integer accumulators of cv2's polygon moments; the host finishes centroid / distances / median.
//
// One block per (class, frame).  The class's bit plane (1 bit per pixel, one-pixel zero frame: 132 KB at
// 1000 x 1000) is packed into shared memory with warp ballots, so every probe of the border walk is a
// shared-memory bit test instead of a global byte load.  Pass 1: every "tip" pixel (set, West and the three pixels
// above empty) starts a walk on its own thread; walks that are not an outer border's first pixel stop at the first
// raster-earlier pixel they meet, completed walks compete by (|a00|, start index) through one 64-bit atomicMax
// (ties go to the later start: cv2 returns contours in reverse discovery order and Python's max keeps the first).
// The largest outer border over ALL components is the largest EXTERNAL one (a component nested in a hole is
// strictly inside a larger border), so no hierarchy is needed.  Pass 2: one thread re-walks the winner and writes
// its points.  Latency-bound by design (a 3000-pixel lumen border is a 3000-step dependent chain); the work that
// parallelises -- packing, tip search, all the small walks -- is spread over the block.
#include <cuda_runtime.h>
#include <cstdint>

#include "common.h"
#include "contour_core.h"

namespace octseg {

constexpr int kContourThreads = 512;

__global__ void __launch_bounds__(kContourThreads) contour_largest_kernel(const uint32_t* __restrict__ mask, int H, int W,
                                                                          int pitch, long long* __restrict__ sums,
                                                                          int* __restrict__ nverts,
                                                                          int16_t* __restrict__ verts, int cap) {
  extern __shared__ uint32_t pl[];  // (H + 2) x pitch
  __shared__ unsigned long long best;
  __shared__ int owner;              // start index of the walk that wrote its points during pass 1 (-1: none)
  __shared__ ContourSums owner_sums;
  const int c = blockIdx.x, n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kContourThreads / 32;
  for (int i = tid; i < pitch; i += kContourThreads) pl[i] = 0u, pl[(H + 1) * pitch + i] = 0u;
  if (tid == 0) best = 0ull, owner = -1, pl[(H + 2) * pitch] = 0u;
  const uint32_t* mimg = mask + static_cast<size_t>(n) * H * W;
  // pack: a warp takes whole rows; one step = 128 pixels (4 plane words): a lane loads 4 pixels (16 bytes) and keeps
  // the 4 presence bits of class c, lane pairs join nibbles into bytes and two more xor-shuffles OR the 4 bytes of a
  // word together.  Up to 8 steps of a row (1024 pixels) are loaded before any is combined: the loop is bound by load latency.
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0;
  const int units = (pitch - 1 + 3) / 4;
  const uint32_t sel = 0xffu << (8 * c);
  for (int y = warp; y < H; y += kWarps) {
    const uint32_t* row = mimg + static_cast<size_t>(y) * W;
    uint32_t* prow = pl + (y + 1) * pitch;
    for (int u0 = 0; u0 < units; u0 += 8) {
      uint4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int x = 128 * (u0 + j) + 4 * lane;
        q[j] = make_uint4(0u, 0u, 0u, 0u);
        if (u0 + j < units && x < W) {
          if (vec_ok) {
            q[j] = __ldg(reinterpret_cast<const uint4*>(row + x));
          } else {
            q[j].x = __ldg(row + x);
            if (x + 1 < W) q[j].y = __ldg(row + x + 1);
            if (x + 2 < W) q[j].z = __ldg(row + x + 2);
            if (x + 3 < W) q[j].w = __ldg(row + x + 3);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (u0 + j >= units) break;  // warp-uniform
        const uint32_t nib = ((q[j].x & sel) ? 1u : 0u) | ((q[j].y & sel) ? 2u : 0u) | ((q[j].z & sel) ? 4u : 0u) | ((q[j].w & sel) ? 8u : 0u);
        uint32_t v = (nib | (__shfl_xor_sync(0xffffffffu, nib, 1) << 4)) << (4 * (lane & 6));  // even lane 2b: byte b
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        const int k = 1 + 4 * (u0 + j) + (lane >> 3);
        if ((lane & 7) == 0 && k < pitch) prow[k] = v;
      }
    }
  }
  for (int y = tid; y < H; y += kContourThreads) pl[(y + 1) * pitch] = 0u;
  __syncthreads();

  const long long max_steps = 4ll * H * W + 16;
  constexpr long long kShortWalk = 48;
  const size_t o = static_cast<size_t>(n) * 4 + c;
  int16_t* out_verts = verts + o * cap * 2;
  unsigned long long mine = 0ull;
  for (int i = tid; i < H * pitch; i += kContourThreads) {
    const int y = i / pitch, k = i - y * pitch;
    uint32_t cur, touch;
    uint32_t tips = tip_bits(pl, pitch, y + 1, k, cur, touch);
    while (tips) {
      const int b = __ffs(tips) - 1;
      tips &= tips - 1;
      if (tip_run_touches(cur, touch, b)) continue;
      const int x = 32 * k + b - 32;
      ContourSums s;
      WalkResult res = trace_border<false>(pl, pitch, x, y, s, nullptr, 0, kShortWalk);
      if (res == kWalkBudget) {
        // a long border: the first such walk of the block writes its points straight into the output (in an
        // OCT mask it is the one large object, so the winner rarely has to be walked a second time)
        if (atomicCAS(&owner, -1, y * W + x) == -1) {
          res = trace_border<true>(pl, pitch, x, y, s, out_verts, cap, max_steps);
          owner_sums = s;
        } else {
          res = trace_border<false>(pl, pitch, x, y, s, nullptr, 0, max_steps);
        }
      }
      if (res != kWalkDone) continue;
      const unsigned long long area = static_cast<unsigned long long>(s.a00 < 0 ? -s.a00 : s.a00);
      if (area > 0) mine = max(mine, (area << 32) | static_cast<unsigned>(y * W + x));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, o));
  if (lane == 0 && mine) atomicMax(&best, mine);
  __syncthreads();
  if (tid == 0) {
    ContourSums s;
    s.a00 = s.a10 = s.a01 = 0;
    s.nverts = 0;
    long long start = -1;
    if (best) {
      start = static_cast<long long>(best & 0xffffffffull);
      if (start == owner)
        s = owner_sums;
      else
        trace_border<true>(pl, pitch, static_cast<int>(start % W), static_cast<int>(start / W), s, out_verts, cap, max_steps);
    }
    sums[o * 4] = s.a00;
    sums[o * 4 + 1] = s.a10;
    sums[o * 4 + 2] = s.a01;
    sums[o * 4 + 3] = start;
    nverts[o] = s.nverts;
  }
}

}  // namespace octseg

using namespace octseg;

extern "C" int octseg_contour_largest(const uint8_t* mask, int32_t N, int32_t H, int32_t W, int64_t* sums, int32_t* nverts,
                                      int16_t* verts, int32_t cap, void* stream) {
  if (!mask || !sums || !nverts || !verts) return fail(OCTSEG_EINVAL, "contour_largest: null argument");
  if (reinterpret_cast<uintptr_t>(mask) & 3) return fail(OCTSEG_EINVAL, "contour_largest: mask must be 4-byte aligned");
  if (N <= 0 || H <= 0 || W <= 0) return OCTSEG_OK;
  if (N > 65535 || H > 32766 || W > 32766 || cap < 0) return fail(OCTSEG_EINVAL, "contour_largest: N <= 65535, H, W <= 32766, cap >= 0");
  const int pitch = plane_pitch(W);
  const size_t smem = (static_cast<size_t>(H + 2) * pitch + 1) * 4;
  if (smem > 226 * 1024)
    return fail(OCTSEG_EINVAL, "contour_largest: the %d x %d bit plane (%zu bytes) does not fit shared memory", H, W, smem);
  static unsigned long long configured = 0;  // bit d = attribute set on device d
  int dev = 0;
  OCTSEG_CUDA(cudaGetDevice(&dev));
  if (!((configured >> (dev & 63)) & 1ull)) {
    OCTSEG_CUDA(cudaFuncSetAttribute(contour_largest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured |= 1ull << (dev & 63);
  }
  contour_largest_kernel<<<dim3(4, N), kContourThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint32_t*>(mask), H, W, pitch, reinterpret_cast<long long*>(sums), nverts, verts, cap);
  return check_launch("contour_largest_kernel");
}
