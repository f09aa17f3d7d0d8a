// Overlay cosmetics of save_results (src/data/utils.py:195-235 + get_img_mask_union_pil,
// src/models/smp/utils.py:203-213) in one pass, bit-exact vs the reference's cv2 + PIL output:
// per class, in cfg.classes order,
//   closed = morphologyEx(mask, CLOSE, ellipse 5x5)        (dilate then erode, cv2 default borders)
//   rim    = dilate(closed, ellipse 7x7) & ~erode(closed, ellipse 7x7)
//   k      = 5x5 binomial sum of closed (GaussianBlur(5,5,0) = k/256, BORDER_REFLECT_101)
//   img    = paste(img, colour, fill_lut[k]);  img = paste(img, colour, rim ? rim_alpha : 0)
// with PIL's paste  out = ((t >> 8) + t) >> 8,  t = dst*(255-a) + src*a + 128.
//
// The four classes of a pixel are the four bytes of one 32-bit mask word, so the morphology runs on all
// classes at once (OR / AND of words) and the binomial sums ride in two registers of 16-bit lanes.
// One block = 32 x 8 output pixels; the mask tile (halo 7), its 5x5 dilation (halo 5) and the closed
// tile (halo 3) live in shared memory.
#include <cuda_runtime.h>
#include <cstdint>

#include "common.h"

namespace octseg {

constexpr int kOvX = 32, kOvY = 8;
constexpr int kMW = kOvX + 14, kMH = kOvY + 14;  // mask tile (halo 2 + 2 + 3)
constexpr int kDW = kOvX + 10, kDH = kOvY + 10;  // dilated tile (halo 2 + 3)
constexpr int kCW = kOvX + 6, kCH = kOvY + 6;    // closed tile (halo 3)
constexpr uint32_t kOnes = 0x01010101u;

struct OverlayParams {
  const uint8_t* img;    // [N][H][W][3] RGB
  const uint32_t* mask;  // [N][H][W]: byte c = class channel c, non-zero = present
  uint8_t* out;          // [N][H][W][3]
  int N, H, W;
  int n_order;
  int order[4];          // class channels in paint order (cfg.classes)
  int color[4][3];       // RGB of class channel c
  int rim_alpha;
  uint8_t fill_lut[260]; // alpha of the blurred fill for k = 0..256
};

// rows of the elliptic structuring elements as bit masks (bit dx+r set = tap present)
__constant__ uint8_t kEl5[5] = {0x04, 0x1f, 0x1f, 0x1f, 0x04};
__constant__ uint8_t kEl7[7] = {0x08, 0x3e, 0x7f, 0x7f, 0x7f, 0x3e, 0x08};

__device__ __forceinline__ int reflect101(int v, int n) {  // cv2 BORDER_REFLECT_101 for |overshoot| < n
  if (v < 0) v = -v;
  if (v >= n) v = 2 * (n - 1) - v;
  return v;
}

__global__ void __launch_bounds__(kOvX* kOvY) overlay_kernel(const OverlayParams p) {
  __shared__ uint32_t sm_m[kMH * kMW];
  __shared__ uint32_t sm_d[kDH * kDW];
  __shared__ uint32_t sm_c[kCH * kCW];
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kOvX, y0 = blockIdx.y * kOvY;
  const int tid = threadIdx.y * kOvX + threadIdx.x;
  const uint32_t* mimg = p.mask + static_cast<size_t>(n) * p.H * p.W;

  // mask tile, bytes normalised to {0,1}; 0 outside the image (dilate's border value)
  for (int i = tid; i < kMH * kMW; i += kOvX * kOvY) {
    const int ty = i / kMW, tx = i - ty * kMW;
    const int y = y0 - 7 + ty, x = x0 - 7 + tx;
    uint32_t v = 0;
    if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
      const uint32_t w = __ldg(mimg + static_cast<size_t>(y) * p.W + x);
      v = __vminu4(w, kOnes);
    }
    sm_m[i] = v;
  }
  __syncthreads();
  // 5x5 dilation; positions outside the image hold all-ones (erode's border value)
  for (int i = tid; i < kDH * kDW; i += kOvX * kOvY) {
    const int ty = i / kDW, tx = i - ty * kDW;
    const int y = y0 - 5 + ty, x = x0 - 5 + tx;
    uint32_t v = kOnes;
    if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
      v = 0;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
          if ((kEl5[dy] >> dx) & 1) v |= sm_m[(ty + dy) * kMW + tx + dx];
    }
    sm_d[i] = v;
  }
  __syncthreads();
  // closed = 5x5 erosion of the dilation (0 outside the image: never read as such, see below)
  for (int i = tid; i < kCH * kCW; i += kOvX * kOvY) {
    const int ty = i / kCW, tx = i - ty * kCW;
    const int y = y0 - 3 + ty, x = x0 - 3 + tx;
    uint32_t v = 0;
    if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
      v = kOnes;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
          if ((kEl5[dy] >> dx) & 1) v &= sm_d[(ty + dy) * kDW + tx + dx];
    }
    sm_c[i] = v;
  }
  __syncthreads();

  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= p.W || y >= p.H) return;
  // 7x7: dilate (outside = 0) and erode (outside = 1) of the closed mask
  uint32_t dil = 0, ero = kOnes;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    const int yy = y + dy - 3;
    const bool yin = yy >= 0 && yy < p.H;
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      if (!((kEl7[dy] >> dx) & 1)) continue;
      const int xx = x + dx - 3;
      const bool in = yin && xx >= 0 && xx < p.W;
      const uint32_t c = sm_c[(threadIdx.y + dy) * kCW + threadIdx.x + dx];
      dil |= in ? c : 0u;
      ero &= in ? c : kOnes;
    }
  }
  const uint32_t rim = dil & ~ero;
  // 5x5 binomial sums (16-bit lanes: classes 0,2 in k02, classes 1,3 in k13), reflected borders
  uint32_t k02 = 0, k13 = 0;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    const int ry = reflect101(y + dy - 2, p.H) - y0 + 3;
    const int wy = dy == 0 || dy == 4 ? 1 : (dy == 2 ? 6 : 4);
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int rx = reflect101(x + dx - 2, p.W) - x0 + 3;
      const int wx = dx == 0 || dx == 4 ? 1 : (dx == 2 ? 6 : 4);
      const uint32_t c = sm_c[ry * kCW + rx];
      k02 += static_cast<uint32_t>(wy * wx) * (c & 0x00ff00ffu);
      k13 += static_cast<uint32_t>(wy * wx) * ((c >> 8) & 0x00ff00ffu);
    }
  }
  const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
  int rgb[3] = {p.img[pix * 3], p.img[pix * 3 + 1], p.img[pix * 3 + 2]};
  for (int i = 0; i < p.n_order; ++i) {
    const int c = p.order[i];
    const uint32_t kk = ((c & 1) ? k13 : k02) >> (16 * (c >> 1)) & 0xffffu;
    const int a1 = p.fill_lut[kk];
    const int a2 = ((rim >> (8 * c)) & 1u) ? p.rim_alpha : 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      int t = rgb[ch] * (255 - a1) + p.color[c][ch] * a1 + 128;
      int v = ((t >> 8) + t) >> 8;
      t = v * (255 - a2) + p.color[c][ch] * a2 + 128;
      rgb[ch] = ((t >> 8) + t) >> 8;
    }
  }
  p.out[pix * 3] = static_cast<uint8_t>(rgb[0]);
  p.out[pix * 3 + 1] = static_cast<uint8_t>(rgb[1]);
  p.out[pix * 3 + 2] = static_cast<uint8_t>(rgb[2]);
}

}  // namespace octseg

using namespace octseg;

extern "C" int octseg_overlay(const uint8_t* img, const uint8_t* mask, uint8_t* out, int32_t N, int32_t H, int32_t W,
                              const int32_t* h_order, int32_t n_order, const uint8_t* h_colors /* [4][3] */,
                              const uint8_t* h_fill_lut /* [257] */, int32_t rim_alpha, void* stream) {
  if (!img || !mask || !out || !h_order || !h_colors || !h_fill_lut) return fail(OCTSEG_EINVAL, "overlay: null argument");
  if (n_order < 0 || n_order > 4) return fail(OCTSEG_EINVAL, "overlay: n_order out of range");
  if (reinterpret_cast<uintptr_t>(mask) & 3) return fail(OCTSEG_EINVAL, "overlay: mask must be 4-byte aligned");
  if (N <= 0 || H <= 0 || W <= 0) return OCTSEG_OK;
  if (N > 65535 || H < 3 || W < 3) return fail(OCTSEG_EINVAL, "overlay: needs N <= 65535 and H, W >= 3");
  OverlayParams p;
  p.img = img;
  p.mask = reinterpret_cast<const uint32_t*>(mask);
  p.out = out;
  p.N = N;
  p.H = H;
  p.W = W;
  p.n_order = n_order;
  for (int i = 0; i < 4; ++i) {
    p.order[i] = i < n_order ? h_order[i] : 0;
    if (i < n_order && (h_order[i] < 0 || h_order[i] > 3)) return fail(OCTSEG_EINVAL, "overlay: bad class index in order");
    for (int c = 0; c < 3; ++c) p.color[i][c] = h_colors[i * 3 + c];
  }
  p.rim_alpha = rim_alpha;
  for (int k = 0; k < 257; ++k) p.fill_lut[k] = h_fill_lut[k];
  dim3 grid(cdiv(W, kOvX), cdiv(H, kOvY), N);
  overlay_kernel<<<grid, dim3(kOvX, kOvY), 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("overlay_kernel");
}
