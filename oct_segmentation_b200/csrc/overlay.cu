// Overlay cosmetics of save_results (src/data/utils.py:195-235 + get_img_mask_union_pil,
// src/models/smp/utils.py:203-213) in one pass, bit-exact vs the reference's cv2 + PIL output:
// per class, in cfg.classes order,
//   closed = morphologyEx(mask, CLOSE, ellipse 5x5)        (dilate then erode, cv2 default borders)
//   rim    = dilate(closed, ellipse 7x7) & ~erode(closed, ellipse 7x7)
//   k      = 5x5 binomial sum of closed (GaussianBlur(5,5,0) = k/256, BORDER_REFLECT_101)
//   img    = paste(img, colour, fill_lut[k]);  img = paste(img, colour, rim ? rim_alpha : 0)
// with PIL's paste  out = ((t >> 8) + t) >> 8,  t = dst*(255-a) + src*a + 128.
//
// overlay_kernel (the product path): the morphology runs on BIT PLANES -- one 32-bit word = 32 pixels of one
// row of one class -- so a dilation / erosion row is a handful of funnel shifts and ORs / ANDs for 32 pixels.
// One block = 128 x TY output pixels (TY = 32); region = tile + 7 halo rows and one halo word (32 px) per side:
//   S0  mask words (4 class bytes per pixel, 16-byte loads) -> bit planes M: the 8 lanes of a plane word exchange
//       their bits with three shuffles + two byte permutes (a 4 x 4 byte transpose yields all four classes at once)
//   S1  D  = dilate5(M), image exterior forced to 1 (erode's border value)
//   S2  E  = erode5(D);  Cd = E inside / 0 outside, Ce = E inside / 1 outside, Cr = Cd + reflected columns
//   S3  RIM = dilate7(Cd) & ~erode7(Ce);  ANY / ALL = OR / AND of the 5x5 (reflected) window of Cr;
//       ACT = per 32-pixel word, the classes with any RIM | ANY bit
//   S4  one warp = 32 pixels of a row: classes absent from the ACT word are skipped warp-uniformly, pixels with a
//       uniform window take k = 256, only object-boundary pixels compute the binomial sum (5 windows, popcounts).
// overlay_bytes_kernel is the first, byte-lane version (4 classes = 4 bytes of a word, 32 x 8 tiles), kept as an
// A/B reference (OCTSEG_OVERLAY_IMPL=bytes).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>

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

__global__ void __launch_bounds__(kOvX* kOvY) overlay_bytes_kernel(const OverlayParams p) {
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

// ---------------------------------------------------------------------------------------------- bit planes
constexpr int kTileWords = 4;            // 128 output pixels per tile row
constexpr int kRW = kTileWords + 2;      // region words per row (one halo word per side)

// OR / AND over the horizontal taps dx = -R..R of a bit-plane row (p, c, n = previous, current, next word)
template <int R>
__device__ __forceinline__ uint32_t hor_or(uint32_t p, uint32_t c, uint32_t n) {
  uint32_t v = c;
#pragma unroll
  for (int s = 1; s <= R; ++s) v |= __funnelshift_l(p, c, s) | __funnelshift_r(c, n, s);
  return v;
}
template <int R>
__device__ __forceinline__ uint32_t hor_and(uint32_t p, uint32_t c, uint32_t n) {
  uint32_t v = c;
#pragma unroll
  for (int s = 1; s <= R; ++s) v &= __funnelshift_l(p, c, s) & __funnelshift_r(c, n, s);
  return v;
}
// bits of the 32 pixels starting at column xs of row y that lie inside the image
__device__ __forceinline__ uint32_t inimg_word(int y, int xs, int H, int W) {
  if (y < 0 || y >= H) return 0u;
  const int lo = max(0, -xs), hi = min(32, W - xs);
  if (hi <= lo) return 0u;
  return (hi == 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u);
}

template <int TY>
__global__ void __launch_bounds__(256) overlay_kernel(const OverlayParams p) {
  constexpr int RH = TY + 14, PL = RH * kRW, TP = TY * kTileWords;
  extern __shared__ uint32_t sm[];
  uint32_t* M = sm;
  uint32_t* D = M + 4 * PL;
  uint32_t* Cd = D + 4 * PL;
  uint32_t* Ce = Cd + 4 * PL;
  uint32_t* Cr = Ce + 4 * PL;
  uint32_t* RIM = Cr + 4 * PL;
  uint32_t* ANY = RIM + 4 * TP;
  uint32_t* ALL = ANY + 4 * TP;
  uint32_t* ACT = ALL + 4 * TP;  // per tile word: bit c = class c has rim or fill pixels among these 32
  uint8_t* lut = reinterpret_cast<uint8_t*>(ACT + TP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, H = p.H, W = p.W;
  const int x0 = blockIdx.x * (32 * kTileWords), y0 = blockIdx.y * TY;
  const int xr0 = x0 - 32, yr0 = y0 - 7;
  const uint32_t* mimg = p.mask + static_cast<size_t>(n) * H * W;
  for (int i = tid; i < 257; i += 256) lut[i] = p.fill_lut[i];
  for (int i = tid; i < TP; i += 256) ACT[i] = 0u;

  // S0: pack.  A lane loads 4 pixels (16 bytes); the 8 lanes of a plane word exchange their bits with three shuffles.
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0;
  for (int it = warp; it < RH * 2; it += 8) {
    const int r = it >> 1, u_idx = it & 1;
    const int y = yr0 + r, x = xr0 + 128 * u_idx + 4 * lane;
    const bool live = !(u_idx == 1 && lane >= 16);  // second unit of a row = words 4, 5 only
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (live && y >= 0 && y < H && x >= 0 && x < W) {
      const uint32_t* src = mimg + static_cast<size_t>(y) * W + x;
      if (vec_ok) {
        q = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
        q.x = __ldg(src);
        if (x + 1 < W) q.y = __ldg(src + 1);
        if (x + 2 < W) q.z = __ldg(src + 2);
        if (x + 3 < W) q.w = __ldg(src + 3);
      }
    }
    const uint32_t t = __vminu4(q.x, kOnes) + 2u * __vminu4(q.y, kOnes) + 4u * __vminu4(q.z, kOnes) + 8u * __vminu4(q.w, kOnes);
    // byte c of t = the 4 presence bits of class c for this lane's 4 pixels.  Pair lanes (nibbles -> bytes: 8 pixels),
    // then transpose the 4 x 4 byte matrix held by the even lanes of each 8-lane group (rows = byte position in the
    // word, columns = classes) with two shuffle + byte-permute rounds: even lane 2k ends up with class k's word.
    uint32_t u = t | (__shfl_xor_sync(0xffffffffu, t, 1) << 4);
    uint32_t v = __shfl_xor_sync(0xffffffffu, u, 2);
    u = __byte_perm(u, v, (lane & 2) ? 0x3715u : 0x6240u);
    v = __shfl_xor_sync(0xffffffffu, u, 4);
    u = __byte_perm(u, v, (lane & 4) ? 0x3276u : 0x5410u);
    if (live && !(lane & 1)) M[((lane & 7) >> 1) * PL + r * kRW + 4 * u_idx + (lane >> 3)] = u;
  }
  __syncthreads();

  // S1: D = dilate(M, ellipse 5x5) (rows +-2: centre tap; rows -1..1: 5 taps); exterior = 1
  for (int i = tid; i < (RH - 4) * kRW; i += 256) {
    const int r = 2 + i / kRW, j = i % kRW;
    const uint32_t ext = ~inimg_word(yr0 + r, xr0 + 32 * j, H, W);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t* m = M + c * PL + r * kRW + j;
      uint32_t acc = m[-2 * kRW] | m[2 * kRW];
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t* row = m + dy * kRW;
        acc |= hor_or<2>(j > 0 ? row[-1] : 0u, row[0], j < kRW - 1 ? row[1] : 0u);
      }
      D[c * PL + r * kRW + j] = acc | ext;
    }
  }
  __syncthreads();

  // S2: E = erode(D, ellipse 5x5) -> the closed mask in its three border conventions
  for (int i = tid; i < (RH - 8) * kRW; i += 256) {
    const int r = 4 + i / kRW, j = i % kRW;
    const uint32_t in = inimg_word(yr0 + r, xr0 + 32 * j, H, W);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t* d = D + c * PL + r * kRW + j;
      uint32_t acc = d[-2 * kRW] & d[2 * kRW];
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t* row = d + dy * kRW;
        acc &= hor_and<2>(j > 0 ? row[-1] : 0xffffffffu, row[0], j < kRW - 1 ? row[1] : 0xffffffffu);
      }
      const int o = c * PL + r * kRW + j;
      Cd[o] = acc & in;
      Ce[o] = acc | ~in;
      Cr[o] = acc & in;
    }
  }
  __syncthreads();
  // BORDER_REFLECT_101 columns of Cr: x = -k <- x = k and x = W-1+k <- x = W-1-k (k = 1, 2); one thread per row
  if (x0 == 0 || xr0 + 32 * kRW > W) {
    for (int i = tid; i < 4 * (RH - 8); i += 256) {
      const int c = i / (RH - 8), r = 4 + i % (RH - 8);
      const int y = yr0 + r;
      if (y < 0 || y >= H) continue;
      uint32_t* row = Cr + c * PL + r * kRW;
#pragma unroll
      for (int k = 1; k <= 2; ++k) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int bs = (side ? W - 1 - k : k) - xr0, bd = (side ? W - 1 + k : -k) - xr0;
          if (bd >= 0 && bd < 32 * kRW && bs >= 0 && bs < 32 * kRW) row[bd >> 5] |= ((row[bs >> 5] >> (bs & 31)) & 1u) << (bd & 31);
        }
      }
    }
    __syncthreads();
  }

  // S3: rim and the uniform-window planes for the tile's words
  for (int i = tid; i < TP * 2; i += 256) {
    const int half = i & 1, j = (i >> 1) & (kTileWords - 1), r = i >> 3;
    const int y = y0 + r;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * half + cc;
      uint32_t rim = 0, any = 0, all = 0;
      if (y < H) {
        const int o = c * PL + (r + 7) * kRW + j + 1;
        const uint32_t* cd = Cd + o;
        const uint32_t* ce = Ce + o;
        uint32_t dil = cd[-3 * kRW] | cd[3 * kRW], ero = ce[-3 * kRW] & ce[3 * kRW];
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
          const uint32_t* a = cd + dy * kRW;
          const uint32_t* b = ce + dy * kRW;
          if (dy == -2 || dy == 2) {
            dil |= hor_or<2>(a[-1], a[0], a[1]);
            ero &= hor_and<2>(b[-1], b[0], b[1]);
          } else {
            dil |= hor_or<3>(a[-1], a[0], a[1]);
            ero &= hor_and<3>(b[-1], b[0], b[1]);
          }
        }
        rim = dil & ~ero;
        all = 0xffffffffu;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
          const uint32_t* a = Cr + c * PL + (reflect101(y + dy, H) - yr0) * kRW + j + 1;
          any |= hor_or<2>(a[-1], a[0], a[1]);
          all &= hor_and<2>(a[-1], a[0], a[1]);
        }
      }
      RIM[c * TP + r * kTileWords + j] = rim;
      ANY[c * TP + r * kTileWords + j] = any;
      ALL[c * TP + r * kTileWords + j] = all;
      if (rim | any) atomicOr(&ACT[r * kTileWords + j], 1u << c);
    }
  }
  __syncthreads();

  // S4: paste.  One warp = one word (32 pixels) of a row.
  const int a_full = lut[256], a_none = lut[0];
  for (int it = warp; it < TP; it += 8) {
    const int r = it / kTileWords, j = it % kTileWords;
    const int y = y0 + r, x = x0 + 32 * j + lane;
    if (y >= H || x0 + 32 * j >= W) continue;
    const bool act = x < W;
    const size_t pix = (static_cast<size_t>(n) * H + y) * W + x;
    int rgb[3] = {0, 0, 0};
    if (act) {
      rgb[0] = p.img[pix * 3];
      rgb[1] = p.img[pix * 3 + 1];
      rgb[2] = p.img[pix * 3 + 2];
    }
    const uint32_t act_classes = a_none == 0 ? ACT[it] : 0xfu;  // warp-uniform
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= p.n_order) break;
      const int c = p.order[i];
      if (!((act_classes >> c) & 1u)) continue;  // nothing of this class near these 32 pixels
      const int o = c * TP + it;
      const uint32_t rimw = RIM[o], anyw = ANY[o];
      const int a2 = ((rimw >> lane) & 1u) ? p.rim_alpha : 0;
      int a1 = a_none;
      if ((anyw >> lane) & 1u) {
        if ((ALL[o] >> lane) & 1u) {
          a1 = a_full;
        } else {  // object boundary: k = 5x5 binomial sum of the closed mask
          int k = 0;
#pragma unroll
          for (int dy = -2; dy <= 2; ++dy) {
            const uint32_t* a = Cr + c * PL + (reflect101(y + dy, H) - yr0) * kRW + j + 1;
            const uint32_t lo = lane >= 2 ? a[0] : a[-1], hi = lane >= 2 ? a[1] : a[0];
            const uint32_t win = __funnelshift_r(lo, hi, (lane - 2) & 31) & 31u;  // bit i = column x - 2 + i
            const int h = __popc(win) + 3 * __popc(win & 14u) + 2 * static_cast<int>((win >> 2) & 1u);
            k += (dy == 0 ? 6 : (dy == -1 || dy == 1 ? 4 : 1)) * h;
          }
          a1 = lut[k];
        }
      }
      if (a1 | a2) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          int t = rgb[ch] * (255 - a1) + p.color[c][ch] * a1 + 128;
          const int v = ((t >> 8) + t) >> 8;
          t = v * (255 - a2) + p.color[c][ch] * a2 + 128;
          rgb[ch] = ((t >> 8) + t) >> 8;
        }
      }
    }
    if (act) {
      p.out[pix * 3] = static_cast<uint8_t>(rgb[0]);
      p.out[pix * 3 + 1] = static_cast<uint8_t>(rgb[1]);
      p.out[pix * 3 + 2] = static_cast<uint8_t>(rgb[2]);
    }
  }
}

template <int TY>
static int launch_overlay(const OverlayParams& p, cudaStream_t stream) {
  constexpr size_t smem = (static_cast<size_t>(5 * 4 * (TY + 14) * kRW + 13 * TY * kTileWords)) * 4 + 272;
  static unsigned long long configured = 0;  // bit d = done on device d (the attribute is per device)
  int dev = 0;
  OCTSEG_CUDA(cudaGetDevice(&dev));
  if (!((configured >> (dev & 63)) & 1ull)) {
    OCTSEG_CUDA(cudaFuncSetAttribute(overlay_kernel<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured |= 1ull << (dev & 63);
  }
  dim3 grid(cdiv(p.W, 32 * kTileWords), cdiv(p.H, TY), p.N);
  overlay_kernel<TY><<<grid, 256, smem, stream>>>(p);
  return check_launch("overlay_kernel");
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
  static const char* impl = getenv("OCTSEG_OVERLAY_IMPL");  // A/B switch for tools/bench_prepost.py, not an API
  if (impl && !strcmp(impl, "bytes")) {
    dim3 grid(cdiv(W, kOvX), cdiv(H, kOvY), N);
    overlay_bytes_kernel<<<grid, dim3(kOvX, kOvY), 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("overlay_bytes_kernel");
  }
  if (impl && !strcmp(impl, "ty64")) return launch_overlay<64>(p, static_cast<cudaStream_t>(stream));
  return launch_overlay<32>(p, static_cast<cudaStream_t>(stream));  // 128 x 32 tiles: 0.40 vs 0.47 ms per 32 frames at 1000 x 1000
}
