// HBM-bound pre/post-processing kernels of the predict path (integer / byte work, bit-exact).
//   preprocess : preprocessing_img (src/data/utils.py:159-166)   RGB->BGR + cv2 INTER_LINEAR uint8
//   postprocess: predict.py:92-100 + data/utils.py:231-233 + analysis.py:199
//   thickness  : calculate_object_thickness (src/app/tools/analysis.py:60-130)
#include <cuda_runtime.h>
#include <cstdint>

#include "common.h"

namespace octseg {

// cv2's uint8 bilinear: horizontal pass in int32 with 11-bit coefficients, vertical pass
//   dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
// One thread per destination pixel (3 channels, BGR order written).
__global__ void preprocess_resize_bgr_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                             uint8_t* __restrict__ dst, int S, const int* __restrict__ xofs,
                                             const short* __restrict__ xalpha, const int* __restrict__ yofs,
                                             const short* __restrict__ ybeta, int area2x, int N) {
  const size_t total = static_cast<size_t>(N) * S * S;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int dx = idx % S;
    const int dy = (idx / S) % S;
    const int n = idx / (static_cast<size_t>(S) * S);
    const uint8_t* img = src + static_cast<size_t>(n) * Hs * Ws * 3;
    uint8_t o[3];
    if (area2x) {
      // cv2 switches INTER_LINEAR to the fast 2x2 area average when both scales are exactly 2
      const uint8_t* r0 = img + (static_cast<size_t>(2 * dy) * Ws + 2 * dx) * 3;
      const uint8_t* r1 = r0 + static_cast<size_t>(Ws) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2);
    } else {
      const int sx0 = xofs[dx];
      const int sx1 = min(sx0 + 1, Ws - 1);
      const int a0 = xalpha[2 * dx], a1 = xalpha[2 * dx + 1];
      const int sy = yofs[dy];
      const int y0 = min(max(sy, 0), Hs - 1), y1 = min(max(sy + 1, 0), Hs - 1);
      const int b0 = ybeta[2 * dy], b1 = ybeta[2 * dy + 1];
      const uint8_t* r0 = img + static_cast<size_t>(y0) * Ws * 3;
      const uint8_t* r1 = img + static_cast<size_t>(y1) * Ws * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[sx0 * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
        const int h1 = r1[sx0 * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        o[c] = static_cast<uint8_t>((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
      }
    }
    uint8_t* d = dst + idx * 3;
    d[0] = o[2];  // RGB -> BGR
    d[1] = o[1];
    d[2] = o[0];
  }
}

struct PostParams {
  const uint8_t* chan[4];
  const int* lut[4];
  long long img_stride[4];
  int S[4];
  int order[4];
  int n_order;
  int N, Ho, Wo;
  uint8_t* mask;
  uint8_t* label;
  int* counts;
};

// One thread = 4 consecutive output pixels: 16-byte mask store, 4-byte label store.
__global__ void __launch_bounds__(256) postprocess_kernel(const PostParams p) {
  __shared__ int sm_cnt[4];
  if (threadIdx.x < 4) sm_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int Wq = (p.Wo + 3) >> 2;
  const int n = blockIdx.z;
  const int y = blockIdx.y;
  int cnt[4] = {0, 0, 0, 0};
  for (int xq = blockIdx.x * blockDim.x + threadIdx.x; xq < Wq; xq += gridDim.x * blockDim.x) {
    uint32_t m[4] = {0, 0, 0, 0};  // m[px] = 4 class bytes of pixel px
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (!p.chan[c]) continue;
      const int S = p.S[c];
      const int sy = p.lut[c][y];
      const uint8_t* row = p.chan[c] + static_cast<size_t>(n) * p.img_stride[c] + static_cast<size_t>(sy) * S;
      const int* lx = p.lut[c] + p.Ho;
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const int x = xq * 4 + px;
        if (x < p.Wo) {
          const uint32_t v = row[lx[x]] ? 1u : 0u;
          m[px] |= v << (8 * c);
          cnt[c] += v;
        }
      }
    }
    uint32_t lab = 0;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      uint32_t l = 0;
      for (int k = 0; k < p.n_order; ++k) {  // later classes overwrite earlier ones
        const int c = p.order[k];
        if ((m[px] >> (8 * c)) & 1u) l = c + 1;
      }
      lab |= l << (8 * px);
    }
    const size_t pix = (static_cast<size_t>(n) * p.Ho + y) * p.Wo + xq * 4;
    if (xq * 4 + 3 < p.Wo && (p.Wo & 3) == 0) {
      *reinterpret_cast<uint4*>(p.mask + pix * 4) = make_uint4(m[0], m[1], m[2], m[3]);
      if (p.label) *reinterpret_cast<uint32_t*>(p.label + pix) = lab;
    } else {
      for (int px = 0; px < 4 && xq * 4 + px < p.Wo; ++px) {
        reinterpret_cast<uint32_t*>(p.mask)[pix + px] = m[px];
        if (p.label) p.label[pix + px] = (lab >> (8 * px)) & 0xff;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int v = cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm_cnt[c], v);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sm_cnt[threadIdx.x]) atomicAdd(p.counts + n * 4 + threadIdx.x, sm_cnt[threadIdx.x]);
}

// One warp per ray: 32 radii per step, ballots resolve "last radius of the first object run".
__global__ void __launch_bounds__(256) radial_thickness_kernel(const uint8_t* __restrict__ mask, int H, int W,
                                                               const double* __restrict__ cos_sin,
                                                               int* __restrict__ radii, int max_radius) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 8 + warp;
  const int cls = blockIdx.y, n = blockIdx.z;
  if (ray >= 360) return;
  const double cs = cos_sin[ray], sn = cos_sin[360 + ray];
  const double cx = static_cast<double>(W / 2), cy = static_cast<double>(H / 2);
  const uint8_t* img = mask + static_cast<size_t>(n) * H * W * 4 + cls;
  int current = 0;
  bool found = false;
  for (int base = 1; base < max_radius; base += 32) {
    const int r = base + lane;
    bool inb = false, obj = false;
    if (r < max_radius) {
      // int(center + r * cos): separate multiply and add, truncation toward zero (Python semantics)
      const int x = static_cast<int>(__dadd_rn(cx, __dmul_rn(static_cast<double>(r), cs)));
      const int y = static_cast<int>(__dadd_rn(cy, __dmul_rn(static_cast<double>(r), sn)));
      inb = x >= 0 && x < W && y >= 0 && y < H;
      if (inb) obj = img[(static_cast<size_t>(y) * W + x) * 4] != 0;
    }
    const uint32_t objm = __ballot_sync(0xffffffffu, obj);
    const bool found_before = found || (objm & ((1u << lane) - 1u)) != 0;
    const bool exits = !inb || (!obj && found_before);
    const uint32_t exitm = __ballot_sync(0xffffffffu, exits);
    const uint32_t live = exitm ? ((1u << (__ffs(exitm) - 1)) - 1u) : 0xffffffffu;
    const uint32_t hit = objm & live;
    if (hit) {
      found = true;
      current = base + (31 - __clz(hit));
    }
    if (exitm) break;
  }
  if (lane == 0) radii[(static_cast<size_t>(n) * 4 + cls) * 360 + ray] = found ? current : 0;
}

}  // namespace octseg

using namespace octseg;

extern "C" int octseg_preprocess_resize_bgr(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, uint8_t* dst,
                                            int32_t S, const int32_t* xofs, const int16_t* xalpha,
                                            const int32_t* yofs, const int16_t* ybeta, int32_t area_fast_2x,
                                            void* stream) {
  if (!src || !dst) return fail(OCTSEG_EINVAL, "preprocess: null buffer");
  if (!area_fast_2x && (!xofs || !xalpha || !yofs || !ybeta)) return fail(OCTSEG_EINVAL, "preprocess: null LUT");
  const size_t total = static_cast<size_t>(N) * S * S;
  size_t g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  preprocess_resize_bgr_kernel<<<static_cast<int>(g ? g : 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x, N);
  return check_launch("preprocess_resize_bgr_kernel");
}

extern "C" int octseg_postprocess(const uint8_t* const* h_chan, const int32_t* h_S, const int64_t* h_img_stride,
                                  const int32_t* const* h_lut, const int32_t* h_order, int32_t n_order, int32_t N, int32_t Ho, int32_t Wo,
                                  uint8_t* mask, uint8_t* label, int32_t* counts, void* stream) {
  if (!h_chan || !h_S || !h_lut || !mask || !counts) return fail(OCTSEG_EINVAL, "postprocess: null argument");
  if (n_order < 0 || n_order > 4) return fail(OCTSEG_EINVAL, "postprocess: n_order out of range");
  PostParams p;
  for (int c = 0; c < 4; ++c) {
    p.chan[c] = h_chan[c];
    p.lut[c] = h_lut[c];
    p.S[c] = h_S[c];
    p.img_stride[c] = h_img_stride ? h_img_stride[c] : static_cast<long long>(h_S[c]) * h_S[c];
    if (p.chan[c] && !p.lut[c]) return fail(OCTSEG_EINVAL, "postprocess: class %d has no LUT", c);
  }
  for (int k = 0; k < 4; ++k) p.order[k] = k < n_order ? h_order[k] : 0;
  for (int k = 0; k < n_order; ++k)
    if (p.order[k] < 0 || p.order[k] > 3) return fail(OCTSEG_EINVAL, "postprocess: bad class index in order");
  p.n_order = n_order;
  p.N = N;
  p.Ho = Ho;
  p.Wo = Wo;
  p.mask = mask;
  p.label = label;
  p.counts = counts;
  const int Wq = (Wo + 3) / 4;
  dim3 grid(cdiv(Wq, 256), Ho, N);
  postprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("postprocess_kernel");
}

extern "C" int octseg_radial_thickness(const uint8_t* mask, int32_t N, int32_t H, int32_t W, const double* cos_sin,
                                       int32_t* radii, void* stream) {
  if (!mask || !cos_sin || !radii) return fail(OCTSEG_EINVAL, "radial_thickness: null argument");
  // max_radius = int(sqrt(W^2 + H^2)) // 2, evaluated in double like the reference
  const int max_radius = static_cast<int>(sqrt(static_cast<double>(W) * W + static_cast<double>(H) * H)) / 2;
  dim3 grid(45, 4, N);
  radial_thickness_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, H, W, cos_sin, radii, max_radius);
  return check_launch("radial_thickness_kernel");
}
