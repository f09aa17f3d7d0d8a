// HBM-bound pre/post-processing kernels of the predict path (integer / byte work, bit-exact).
//   preprocess : preprocessing_img (src/data/utils.py:159-166)   RGB->BGR + cv2 INTER_LINEAR uint8
//   postprocess: predict.py:92-100 + data/utils.py:231-233 + analysis.py:199
//   thickness  : calculate_object_thickness (src/app/tools/analysis.py:60-130)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>

#include "common.h"

namespace octseg {

// Stage `bytes` of a global row in shared memory: 16-byte loads when both ends allow it.
__device__ __forceinline__ void stage_row(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
  if (((reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(bytes)) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (bytes >> 4); i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = __ldg(src + i);
  }
}

// Four consecutive destination pixels (4q .. 4q+3, clamped at S-1) of one destination row from its two staged source
// rows: 12 bytes in BGR order.  cv2's uint8 bilinear arithmetic (see the kernel below), or its 2x2 area fast path.
template <int CS>
__device__ __forceinline__ void resize_quad(const uint8_t* r0, const uint8_t* r1, int Ws, int S, const int* __restrict__ xofs,
                                            const short* __restrict__ xalpha, int b0, int b1, int area2x, bool vec, int q,
                                            uint8_t (&o)[12]) {
    int sx[4] = {0, 0, 0, 0}, al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!area2x) {
      if (vec) {
        const int4 s4 = __ldg(reinterpret_cast<const int4*>(xofs) + q);
        const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(xalpha) + q);  // 8 shorts
        sx[0] = s4.x, sx[1] = s4.y, sx[2] = s4.z, sx[3] = s4.w;
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          al[2 * p] = static_cast<short>(aw[p] & 0xffffu);
          al[2 * p + 1] = static_cast<short>(aw[p] >> 16);
        }
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int dx = min(4 * q + p, S - 1);
          sx[p] = xofs[dx];
          al[2 * p] = xalpha[2 * dx];
          al[2 * p + 1] = xalpha[2 * dx + 1];
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int dx = min(4 * q + p, S - 1);
      uint8_t v[3];
      if (area2x) {
        const uint8_t* p0 = r0 + 2 * CS * dx;
        const uint8_t* p1 = r1 + 2 * CS * dx;
#pragma unroll
        for (int c = 0; c < CS; ++c) v[c] = static_cast<uint8_t>((p0[c] + p0[CS + c] + p1[c] + p1[CS + c] + 2) >> 2);
      } else {
        const int x0 = sx[p] * CS, x1 = min(sx[p] + 1, Ws - 1) * CS;
        const int a0 = al[2 * p], a1 = al[2 * p + 1];
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          const int h0 = r0[x0 + c] * a0 + r0[x1 + c] * a1;
          const int h1 = r1[x0 + c] * a0 + r1[x1 + c] * a1;
          v[c] = static_cast<uint8_t>((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
        }
      }
      o[3 * p] = v[CS - 1];  // RGB -> BGR
      o[3 * p + 1] = v[CS / 2];
      o[3 * p + 2] = v[0];
    }
}

// cv2's uint8 bilinear: horizontal pass in int32 with 11-bit coefficients, vertical pass
//   dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
// One block per destination row: its two source rows are staged in shared memory with 16-byte loads,
// a thread produces 4 destination pixels (12 bytes, written as three 32-bit words, BGR order) from
// byte reads of the staged rows and one 16-byte read of each coefficient table.
// CS = source channels: 3 (RGB, swapped to BGR) or 1 (grayscale, replicated into the three output channels: the
// documented extension of SURVEY.md section 8a -- resizing a replicated frame == replicating the resized plane).
constexpr int kPreThreads = 256;
template <int CS>
__global__ void __launch_bounds__(kPreThreads) preprocess_resize_bgr_kernel(
    const uint8_t* __restrict__ src, int Hs, int Ws, uint8_t* __restrict__ dst, int S, const int* __restrict__ xofs,
    const short* __restrict__ xalpha, const int* __restrict__ yofs, const short* __restrict__ ybeta, int area2x) {
  extern __shared__ __align__(16) uint8_t rows[];  // [2][row_pitch]
  const int dy = blockIdx.x, n = blockIdx.y;
  const int row_bytes = Ws * CS, pitch = (row_bytes + 15) & ~15;
  const uint8_t* img = src + static_cast<size_t>(n) * Hs * row_bytes;
  int y0, y1, b0 = 0, b1 = 0;
  if (area2x) {  // cv2 switches INTER_LINEAR to the fast 2x2 area average when both scales are exactly 2
    y0 = 2 * dy;
    y1 = 2 * dy + 1;
  } else {
    const int sy = yofs[dy];
    y0 = min(max(sy, 0), Hs - 1);
    y1 = min(max(sy + 1, 0), Hs - 1);
    b0 = ybeta[2 * dy];
    b1 = ybeta[2 * dy + 1];
  }
  stage_row(rows, img + static_cast<size_t>(y0) * row_bytes, row_bytes);
  stage_row(rows + pitch, img + static_cast<size_t>(y1) * row_bytes, row_bytes);
  __syncthreads();
  const uint8_t* r0 = rows;
  const uint8_t* r1 = rows + pitch;
  uint8_t* drow = dst + (static_cast<size_t>(n) * S + dy) * S * 3;
  const bool vec = (S & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(xofs) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(xalpha) & 15) == 0;
  for (int q = threadIdx.x; q < (S + 3) / 4; q += kPreThreads) {
    uint8_t o[12];
    resize_quad<CS>(r0, r1, Ws, S, xofs, xalpha, b0, b1, area2x, vec, q, o);
    if (vec) {
      uint32_t* d32 = reinterpret_cast<uint32_t*>(drow) + 3 * q;
#pragma unroll
      for (int w = 0; w < 3; ++w)
        d32[w] = o[4 * w] | (o[4 * w + 1] << 8) | (o[4 * w + 2] << 16) | (static_cast<uint32_t>(o[4 * w + 3]) << 24);
    } else {
      for (int p = 0; p < 4 && 4 * q + p < S; ++p)
        for (int c = 0; c < 3; ++c) drow[(4 * q + p) * 3 + c] = o[3 * p + c];
    }
  }
}

// The same resize written straight into the network stem's input format (csrc/direct.cu: stem_pack_s2d_kernel): bf16
// [N][S/2][S/2][16], channel (dy*2 + dx)*3 + c = pixel (2y+dy, 2x+dx) of the BGR frame, channels 12..15 zero.  One
// block per PAIR of destination rows (four staged source rows); a thread turns its two 4-pixel quads into two whole
// 32-byte pixels of the packed tensor (two 16-byte stores each).  uint8 -> bf16 is exact.  Saves the uint8 frame's
// write + read and the separate pack launch of the predict path (no normalisation there, model.py:192).
template <int CS>
__global__ void __launch_bounds__(kPreThreads) preprocess_resize_s2d_kernel(
    const uint8_t* __restrict__ src, int Hs, int Ws, uint4* __restrict__ dst, int S, const int* __restrict__ xofs,
    const short* __restrict__ xalpha, const int* __restrict__ yofs, const short* __restrict__ ybeta, int area2x) {
  extern __shared__ __align__(16) uint8_t rows[];  // [4][row_pitch]
  const int r = blockIdx.x, n = blockIdx.y;
  const int row_bytes = Ws * CS, pitch = (row_bytes + 15) & ~15;
  const uint8_t* img = src + static_cast<size_t>(n) * Hs * row_bytes;
  int b[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int dy = 2 * r + h;
    int y0, y1;
    if (area2x) {
      y0 = 2 * dy;
      y1 = 2 * dy + 1;
    } else {
      const int sy = yofs[dy];
      y0 = min(max(sy, 0), Hs - 1);
      y1 = min(max(sy + 1, 0), Hs - 1);
      b[h][0] = ybeta[2 * dy];
      b[h][1] = ybeta[2 * dy + 1];
    }
    stage_row(rows + (2 * h) * pitch, img + static_cast<size_t>(y0) * row_bytes, row_bytes);
    stage_row(rows + (2 * h + 1) * pitch, img + static_cast<size_t>(y1) * row_bytes, row_bytes);
  }
  __syncthreads();
  const bool vec = (S & 3) == 0 && (reinterpret_cast<uintptr_t>(xofs) & 15) == 0 && (reinterpret_cast<uintptr_t>(xalpha) & 15) == 0;
  uint4* drow = dst + (static_cast<size_t>(n) * (S / 2) + r) * (S / 2) * 2;  // 2 x uint4 per packed pixel
  for (int q = threadIdx.x; q < (S + 3) / 4; q += kPreThreads) {
    uint8_t o0[12], o1[12];
    resize_quad<CS>(rows, rows + pitch, Ws, S, xofs, xalpha, b[0][0], b[0][1], area2x, vec, q, o0);
    resize_quad<CS>(rows + 2 * pitch, rows + 3 * pitch, Ws, S, xofs, xalpha, b[1][0], b[1][1], area2x, vec, q, o1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {          // packed pixel 2q + h = destination pixels 4q + 2h, 4q + 2h + 1 of both rows
      if (2 * q + h >= S / 2) break;
      uint32_t w[8];                       // 16 bf16: row 0 (2 px x 3 ch), row 1 (2 px x 3 ch), 4 zeros
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const uint32_t lo0 = __float_as_uint(static_cast<float>(o0[6 * h + 2 * e])) >> 16;
        const uint32_t hi0 = __float_as_uint(static_cast<float>(o0[6 * h + 2 * e + 1])) & 0xffff0000u;
        const uint32_t lo1 = __float_as_uint(static_cast<float>(o1[6 * h + 2 * e])) >> 16;
        const uint32_t hi1 = __float_as_uint(static_cast<float>(o1[6 * h + 2 * e + 1])) & 0xffff0000u;
        w[e] = lo0 | hi0;
        w[3 + e] = lo1 | hi1;
      }
      w[6] = w[7] = 0u;
      drow[2 * (2 * q + h)] = make_uint4(w[0], w[1], w[2], w[3]);
      drow[2 * (2 * q + h) + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

struct PostParams {
  const uint8_t* chan[4];
  const int* lut[4];
  long long img_stride[4];
  int S[4];
  unsigned long long label_lut;  // 16 x 4 bits: label of every class-presence combination (bit c = class c present)
  int N, Ho, Wo;
  uint8_t* mask;
  uint8_t* label;
  int* counts;
};

// One thread = 4 consecutive output pixels of kPostRows rows: 16-byte mask store, 4-byte label store.
// The kernel is issue-bound (ncu: 76 % issue-active at 125 instructions per pixel in its first form), so the
// per-pixel work is pared down: per class one 16-byte read of the column table per quad and per pixel
// LDG.U8 + min + multiply-add into the pixel's 4-class word; the label is a 16-entry table lookup on the
// class-presence bits ((m * 0x01020408) >> 24); the four per-class counts ride in the four bytes of one
// register (<= 4 px x kPostRows rows per thread) until the end.
constexpr int kPostRows = 8;
__global__ void __launch_bounds__(256) postprocess_kernel(const PostParams p) {
  __shared__ int sm_cnt[4];
  __shared__ uint8_t sm_lab[16];
  if (threadIdx.x < 4) sm_cnt[threadIdx.x] = 0;
  if (threadIdx.x < 16) sm_lab[threadIdx.x] = static_cast<uint8_t>((p.label_lut >> (4 * threadIdx.x)) & 0xF);
  __syncthreads();
  const int Wq = (p.Wo + 3) >> 2;
  const int n = blockIdx.z;
  uint32_t cntp = 0;  // byte c = pixels of class c seen by this thread
  const int y_end = min(p.Ho, (static_cast<int>(blockIdx.y) + 1) * kPostRows);
  for (int y = blockIdx.y * kPostRows; y < y_end; ++y)
    for (int xq = blockIdx.x * blockDim.x + threadIdx.x; xq < Wq; xq += gridDim.x * blockDim.x) {
      uint32_t m[4] = {0, 0, 0, 0};  // m[px] = 4 class bytes of pixel px
      const bool full = xq * 4 + 3 < p.Wo;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (!p.chan[c]) continue;
        const int S = p.S[c];
        const int sy = __ldg(p.lut[c] + y);
        const uint8_t* row = p.chan[c] + static_cast<size_t>(n) * p.img_stride[c] + static_cast<size_t>(sy) * S;
        const int* lx = p.lut[c] + p.Ho + xq * 4;
        int sx[4];
        if (full && (reinterpret_cast<uintptr_t>(lx) & 15) == 0) {  // one 16-byte read of the column table
          const int4 t = __ldg(reinterpret_cast<const int4*>(lx));
          sx[0] = t.x, sx[1] = t.y, sx[2] = t.z, sx[3] = t.w;
        } else {
#pragma unroll
          for (int px = 0; px < 4; ++px) sx[px] = (xq * 4 + px < p.Wo) ? lx[px] : -1;
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          if (full || sx[px] >= 0) {
            const uint32_t v = min(static_cast<uint32_t>(__ldg(row + sx[px])), 1u);
            m[px] = v * (1u << (8 * c)) + m[px];
          }
        }
      }
      cntp += m[0] + m[1] + m[2] + m[3];
      uint32_t lab = 0;
#pragma unroll
      for (int px = 0; px < 4; ++px) lab |= static_cast<uint32_t>(sm_lab[(m[px] * 0x01020408u) >> 24]) << (8 * px);
      const size_t pix = (static_cast<size_t>(n) * p.Ho + y) * p.Wo + xq * 4;
      if (full && (p.Wo & 3) == 0) {
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
    int v = static_cast<int>((cntp >> (8 * c)) & 0xffu);
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

// K-way probability averaging (north-star "ensemble averaging"; opt-in generalisation, SURVEY.md section 8a:
// the reference routes one model per class, src/predict.py:23-28 -- K = 1 is its `sigmoid(y) > 0.5`,
// src/models/smp/model.py:195):  out = (1/K * sum_k sigmoid(logit_k)) > 0.5  over K same-shaped fp32 logit tensors.
// HBM-bound: 4K bytes in + 1 byte out per element; one thread = 4 elements (16-byte loads, 4-byte store),
// the K loads of a thread are independent.
constexpr int kMaxFolds = 8;
struct FoldParams {
  const float* logits[kMaxFolds];
  uint8_t* out;
  long long n;
  int K;
  float inv_k;
};

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) fold_average_threshold_kernel(const FoldParams p) {
  const long long quads = p.n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < quads; q += stride) {
    float4 v[kMaxFolds];
#pragma unroll
    for (int k = 0; k < kMaxFolds; ++k)
      if (k < p.K) v[k] = __ldcs(reinterpret_cast<const float4*>(p.logits[k]) + q);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxFolds; ++k)
      if (k < p.K) {
        s0 += sigmoid_f32(v[k].x);
        s1 += sigmoid_f32(v[k].y);
        s2 += sigmoid_f32(v[k].z);
        s3 += sigmoid_f32(v[k].w);
      }
    const uint32_t w = (s0 * p.inv_k > 0.5f ? 1u : 0u) | (s1 * p.inv_k > 0.5f ? 0x100u : 0u) |
                       (s2 * p.inv_k > 0.5f ? 0x10000u : 0u) | (s3 * p.inv_k > 0.5f ? 0x1000000u : 0u);
    reinterpret_cast<uint32_t*>(p.out)[q] = w;
  }
  // ragged tail (n % 4 elements)
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    const long long i = (quads << 2) + threadIdx.x;
    float s = 0.f;
    for (int k = 0; k < p.K; ++k) s += sigmoid_f32(p.logits[k][i]);
    p.out[i] = s * p.inv_k > 0.5f ? 1 : 0;
  }
}

}  // namespace octseg

using namespace octseg;

template <int CS>
static int launch_preprocess(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, uint8_t* dst, int32_t S,
                             const int32_t* xofs, const int16_t* xalpha, const int32_t* yofs, const int16_t* ybeta,
                             int32_t area_fast_2x, void* stream) {
  if (!src || !dst) return fail(OCTSEG_EINVAL, "preprocess: null buffer");
  if (!area_fast_2x && (!xofs || !xalpha || !yofs || !ybeta)) return fail(OCTSEG_EINVAL, "preprocess: null LUT");
  if (N <= 0 || S <= 0) return OCTSEG_OK;
  if (N > 65535) return fail(OCTSEG_EINVAL, "preprocess: at most 65535 frames per call");
  const size_t smem = 2 * static_cast<size_t>((Ws * CS + 15) & ~15);
  if (smem > 48 * 1024) return fail(OCTSEG_EINVAL, "preprocess: source rows of %d pixels do not fit shared memory", Ws);
  dim3 grid(S, N);
  preprocess_resize_bgr_kernel<CS><<<grid, kPreThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      src, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x);
  return check_launch("preprocess_resize_bgr_kernel");
}

template <int CS>
static int launch_preprocess_s2d(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, void* dst, int32_t S,
                                 const int32_t* xofs, const int16_t* xalpha, const int32_t* yofs, const int16_t* ybeta,
                                 int32_t area_fast_2x, void* stream) {
  if (!src || !dst) return fail(OCTSEG_EINVAL, "preprocess_s2d: null buffer");
  if (!area_fast_2x && (!xofs || !xalpha || !yofs || !ybeta)) return fail(OCTSEG_EINVAL, "preprocess_s2d: null LUT");
  if (S % 2 || (reinterpret_cast<uintptr_t>(dst) & 15)) return fail(OCTSEG_EINVAL, "preprocess_s2d: S must be even and dst 16-byte aligned");
  if (N <= 0 || S <= 0) return OCTSEG_OK;
  if (N > 65535) return fail(OCTSEG_EINVAL, "preprocess_s2d: at most 65535 frames per call");
  const size_t smem = 4 * static_cast<size_t>((Ws * CS + 15) & ~15);
  if (smem > 48 * 1024) return fail(OCTSEG_EINVAL, "preprocess_s2d: source rows of %d pixels do not fit shared memory", Ws);
  dim3 grid(S / 2, N);
  preprocess_resize_s2d_kernel<CS><<<grid, kPreThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      src, Hs, Ws, static_cast<uint4*>(dst), S, xofs, xalpha, yofs, ybeta, area_fast_2x);
  return check_launch("preprocess_resize_s2d_kernel");
}

extern "C" int octseg_preprocess_resize_s2d(const uint8_t* src, int32_t channels, int32_t N, int32_t Hs, int32_t Ws, void* dst,
                                            int32_t S, const int32_t* xofs, const int16_t* xalpha, const int32_t* yofs,
                                            const int16_t* ybeta, int32_t area_fast_2x, void* stream) {
  if (channels == 3) return launch_preprocess_s2d<3>(src, N, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x, stream);
  if (channels == 1) return launch_preprocess_s2d<1>(src, N, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x, stream);
  return fail(OCTSEG_EINVAL, "preprocess_s2d: frames must have 1 or 3 channels");
}

extern "C" int octseg_preprocess_resize_bgr(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, uint8_t* dst,
                                            int32_t S, const int32_t* xofs, const int16_t* xalpha,
                                            const int32_t* yofs, const int16_t* ybeta, int32_t area_fast_2x,
                                            void* stream) {
  return launch_preprocess<3>(src, N, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x, stream);
}

extern "C" int octseg_preprocess_resize_gray(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, uint8_t* dst,
                                             int32_t S, const int32_t* xofs, const int16_t* xalpha,
                                             const int32_t* yofs, const int16_t* ybeta, int32_t area_fast_2x,
                                             void* stream) {
  return launch_preprocess<1>(src, N, Hs, Ws, dst, S, xofs, xalpha, yofs, ybeta, area_fast_2x, stream);
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
  for (int k = 0; k < n_order; ++k)
    if (h_order[k] < 0 || h_order[k] > 3) return fail(OCTSEG_EINVAL, "postprocess: bad class index in order");
  p.label_lut = 0;  // later classes in `order` overwrite earlier ones (src/data/utils.py:231-233)
  for (unsigned idx = 0; idx < 16; ++idx) {
    unsigned long long l = 0;
    for (int k = 0; k < n_order; ++k)
      if ((idx >> h_order[k]) & 1u) l = static_cast<unsigned long long>(h_order[k]) + 1;
    p.label_lut |= l << (4 * idx);
  }
  p.N = N;
  p.Ho = Ho;
  p.Wo = Wo;
  p.mask = mask;
  p.label = label;
  p.counts = counts;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return OCTSEG_OK;
  const int Wq = (Wo + 3) / 4;
  dim3 grid(cdiv(Wq, 256), cdiv(Ho, kPostRows), N);
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

extern "C" int octseg_fold_average_threshold(const float* const* h_logits, int32_t K, int64_t n, uint8_t* out,
                                             void* stream) {
  if (!h_logits || !out) return fail(OCTSEG_EINVAL, "fold_average: null argument");
  if (K < 1 || K > kMaxFolds) return fail(OCTSEG_EINVAL, "fold_average: K must be in 1..%d", kMaxFolds);
  if (n <= 0) return OCTSEG_OK;
  FoldParams p;
  for (int k = 0; k < kMaxFolds; ++k) {
    p.logits[k] = k < K ? h_logits[k] : nullptr;
    if (k < K && (!h_logits[k] || (reinterpret_cast<uintptr_t>(h_logits[k]) & 15)))
      return fail(OCTSEG_EINVAL, "fold_average: logits[%d] must be a 16-byte aligned device pointer", k);
  }
  if (reinterpret_cast<uintptr_t>(out) & 3) return fail(OCTSEG_EINVAL, "fold_average: out must be 4-byte aligned");
  p.out = out;
  p.n = n;
  p.K = K;
  p.inv_k = 1.0f / static_cast<float>(K);
  const long long quads = n >> 2;
  const int blocks = static_cast<int>(std::min<long long>(std::max<long long>((quads + 255) / 256, 1), 148LL * 8));
  fold_average_threshold_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("fold_average_threshold_kernel");
}
