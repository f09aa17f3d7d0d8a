// CUDA-core kernels for the layers that are HBM-bound or have a tiny contraction:
// network stems (Cin = 3), max-pool, depthwise conv (+ squeeze-excite pooling), SE gate,
// per-image SE-scaled projection weights.  Activations are NHWC bf16, 8 channels (16 bytes)
// per thread access.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "common.h"

namespace octseg {

__device__ __forceinline__ float act_f(float x, int act) {
  if (act == OCTSEG_ACT_RELU) return fmaxf(x, 0.f);
  if (act == OCTSEG_ACT_SWISH) {  // x*sigmoid(x) = h*tanh(h) + h, h = x/2: one MUFU
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  if (act == OCTSEG_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return v;
}

// packed fp32x2 math (sm_100 FFMA2, __ffma2_rn): two FMAs per issue slot; float2 lets the compiler
// keep the operands in aligned register pairs without extra moves
// ------------------------------------------------------------------------------------ stem
// One block = 16x16 output pixels; the input patch and the whole filter bank sit in shared
// memory; each thread owns one pixel and CO output channels in registers.
struct StemParams {
  const void* in;
  int in_dtype;
  long long sn, sc, sh, sw;
  int N, H, W;
  const float* weight;
  const float* bias;
  int k, stride, pad_t, pad_l, Ho, Wo, act;
  float mean[3], inv_std[3];
  int normalize;
  __nv_bfloat16* out;
  int out_ldc;
};

template <int CO>
__global__ void __launch_bounds__(256) stem_conv_kernel(const StemParams p) {
  extern __shared__ float smem[];
  const int taps = p.k * p.k * 3;
  float* wsm = smem;               // [taps][CO]
  float* patch = smem + taps * CO; // [ph][pw][3]
  const int pdim = 15 * p.stride + p.k;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * 16, ox0 = blockIdx.x * 16;
  const int iy0 = oy0 * p.stride - p.pad_t, ix0 = ox0 * p.stride - p.pad_l;

  for (int i = threadIdx.x; i < taps * CO; i += 256) wsm[i] = p.weight[i];
  for (int i = threadIdx.x; i < pdim * pdim * 3; i += 256) {
    const int c = i % 3, x = (i / 3) % pdim, y = i / (3 * pdim);
    const int iy = iy0 + y, ix = ix0 + x;
    float v = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
      const long long off = n * p.sn + c * p.sc + iy * p.sh + ix * p.sw;
      v = p.in_dtype == 0 ? reinterpret_cast<const float*>(p.in)[off]
                          : static_cast<float>(reinterpret_cast<const uint8_t*>(p.in)[off]);
      if (p.normalize) v = (v - p.mean[c]) * p.inv_std[c];
    }
    patch[i] = v;
  }
  __syncthreads();

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int oy = oy0 + ty, ox = ox0 + tx;
  float2 acc2[CO / 2];
#pragma unroll
  for (int c = 0; c < CO / 2; ++c) acc2[c] = make_float2(0.f, 0.f);
  for (int ky = 0; ky < p.k; ++ky) {
    for (int kx = 0; kx < p.k; ++kx) {
      const float* pp = patch + ((ty * p.stride + ky) * pdim + tx * p.stride + kx) * 3;
      const float* ww = wsm + (ky * p.k + kx) * 3 * CO;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = pp[c];
        const float2 vv = make_float2(v, v);
        const float4* w4 = reinterpret_cast<const float4*>(ww + c * CO);
#pragma unroll
        for (int g = 0; g < CO / 4; ++g) {
          const float4 w = w4[g];
          acc2[2 * g] = __ffma2_rn(vv, make_float2(w.x, w.y), acc2[2 * g]);
          acc2[2 * g + 1] = __ffma2_rn(vv, make_float2(w.z, w.w), acc2[2 * g + 1]);
        }
      }
    }
  }
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO / 2; ++c) {
    acc[2 * c] = acc2[c].x;
    acc[2 * c + 1] = acc2[c].y;
  }
  if (oy < p.Ho && ox < p.Wo) {
    __nv_bfloat16* o = p.out + ((static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox) * p.out_ldc;
#pragma unroll
    for (int g = 0; g < CO / 8; ++g) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = acc[g * 8 + e] + __ldg(p.bias + g * 8 + e);
      if (p.act == OCTSEG_ACT_RELU) {  // uniform branch: only one activation's code runs
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
      } else if (p.act != OCTSEG_ACT_NONE) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = act_f(f[e], p.act);
      }
      *reinterpret_cast<uint4*>(o + g * 8) = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------ maxpool
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N,
                                    int H, int W, int C8, int Ho, int Wo) {
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C8;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = idx % C8;
    size_t t = idx / C8;
    const int ox = t % Wo;
    t /= Wo;
    const int oy = t % Ho;
    const int n = t / Ho;
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + ((static_cast<size_t>(n) * H + iy) * W + ix) * C8 + c8);
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], f[e]);
      }
    }
    reinterpret_cast<uint4*>(out)[idx] = pack8(m);
  }
}

// ------------------------------------------------------------------------------------ depthwise
// HBM-bound by nature, instruction-bound in practice, so the inner loop is kept lean:
//   * work item = 8 channels (16 B) x 4 consecutive output pixels of one row; a thread keeps ONE
//     channel group for its whole life (bias, SE sums in registers) and strides over pixel groups,
//     consecutive threads take consecutive 16-byte chunks (coalesced whatever C is);
//   * the fp32 filter bank of the block's <= 32 channel groups sits in shared memory;
//   * all input vectors of a filter row are loaded before any is used (one exposed latency per row);
//   * math is packed fp32x2 (fma.rn.f32x2, sm_100): half the FMA issue slots.
// Squeeze-excite sums leave the block as one global atomic per channel.
constexpr int kDwP = 4;          // output pixels per item
constexpr int kDwCgChunk = 32;   // channel groups per block column (256 channels)

// 8 bf16 (uint4) -> four fp32x2 pairs
__device__ __forceinline__ void bf16x8_to_f32x2(const uint4& v, float2 (&f)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}

template <int K, int S, bool CHECK>
__device__ __forceinline__ void dw_rows(float2 (&acc)[kDwP][4], const uint4* __restrict__ in4,
                                        const float* __restrict__ wsm_cg, int C8, int W, int H, int ix0, int iy0) {
  constexpr int WIN = (kDwP - 1) * S + K;
#pragma unroll(K == 3 ? 3 : 1)
  for (int ky = 0; ky < K; ++ky) {
    const int iy = iy0 + ky;
    if (CHECK && (iy < 0 || iy >= H)) continue;
    const uint4* row = in4 + static_cast<size_t>(iy) * W * C8;
    uint4 v[WIN];
#pragma unroll
    for (int dx = 0; dx < WIN; ++dx) {
      const int ix = ix0 + dx;
      v[dx] = (!CHECK || (ix >= 0 && ix < W)) ? __ldg(row + static_cast<size_t>(ix) * C8) : make_uint4(0, 0, 0, 0);
    }
    float2 wr[K][4];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(wsm_cg + (ky * K + kx) * (kDwCgChunk * 8));
      const float4 w1 = *reinterpret_cast<const float4*>(wsm_cg + (ky * K + kx) * (kDwCgChunk * 8) + 4);
      wr[kx][0] = make_float2(w0.x, w0.y);
      wr[kx][1] = make_float2(w0.z, w0.w);
      wr[kx][2] = make_float2(w1.x, w1.y);
      wr[kx][3] = make_float2(w1.z, w1.w);
    }
#pragma unroll
    for (int dx = 0; dx < WIN; ++dx) {
      float2 f[4];
      bf16x8_to_f32x2(v[dx], f);
#pragma unroll
      for (int p = 0; p < kDwP; ++p) {
        const int kx = dx - p * S;  // compile-time after unrolling
        if (kx >= 0 && kx < K) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[p][e] = __ffma2_rn(f[e], wr[kx][e], acc[p][e]);
        }
      }
    }
  }
}

template <int K, int S>
__global__ void __launch_bounds__(256, 2) dwconv_kernel(const __nv_bfloat16* __restrict__ in,
                                                        const __nv_bfloat16* __restrict__ weight,  // [K*K][C] bf16
                                                        const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ out, int H, int W, int C, int pad_t,
                                                        int pad_l, int Ho, int Wo, int pg_per_block, int act,
                                                        float* __restrict__ pool_sum) {
  __shared__ __align__(16) float wsm[K * K][kDwCgChunk * 8];
  __shared__ float sums[8][kDwCgChunk];  // [e][channel group]: conflict-free for consecutive groups
  const int C8 = C >> 3;
  const int cg0 = blockIdx.x * kDwCgChunk;
  const int cgc = min(kDwCgChunk, C8 - cg0);
  const int n = blockIdx.z;
  const int wg = (Wo + kDwP - 1) / kDwP;  // pixel groups per output row
  const int npg = Ho * wg;
  const int pg0 = blockIdx.y * pg_per_block;
  const int pgc = min(pg_per_block, npg - pg0);
  for (int i = threadIdx.x; i < K * K * kDwCgChunk * 8; i += 256) {
    const int tap = i / (kDwCgChunk * 8), cl = i - tap * (kDwCgChunk * 8);
    const int c = cg0 * 8 + cl;
    (&wsm[0][0])[i] = c < C ? __bfloat162float(weight[static_cast<size_t>(tap) * C + c]) : 0.f;
  }
  if (pool_sum)
    for (int i = threadIdx.x; i < 8 * kDwCgChunk; i += 256) (&sums[0][0])[i] = 0.f;
  __syncthreads();
  // thread -> fixed channel group, strided over pixel groups
  const int lanes_pg = 256 / cgc;
  const int cgl = threadIdx.x % cgc;
  const int pgl0 = threadIdx.x / cgc;
  const bool active = pgl0 < lanes_pg;
  const int cg = cg0 + cgl;
  const uint4* in4 = reinterpret_cast<const uint4*>(in) + static_cast<size_t>(n) * H * W * C8 + cg;
  uint4* out4 = reinterpret_cast<uint4*>(out) + static_cast<size_t>(n) * Ho * Wo * C8 + cg;
  const float* wsm_cg = &wsm[0][0] + cgl * 8;
  constexpr int WIN = (kDwP - 1) * S + K;
  float2 bs[4];
  float ps[8];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    bs[e] = active ? make_float2(__ldg(bias + cg * 8 + 2 * e), __ldg(bias + cg * 8 + 2 * e + 1)) : make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 8; ++e) ps[e] = 0.f;
  if (active) {
    for (int pgl = pgl0; pgl < pgc; pgl += lanes_pg) {
      const int pg = pg0 + pgl;
      const int oy = pg / wg;
      const int ox0 = (pg - oy * wg) * kDwP;
      float2 acc[kDwP][4];
#pragma unroll
      for (int p = 0; p < kDwP; ++p)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[p][e] = bs[e];
      const int ix0 = ox0 * S - pad_l;
      const int iy0 = oy * S - pad_t;
      const bool interior = ix0 >= 0 && ix0 + WIN <= W && iy0 >= 0 && iy0 + K <= H;
      if (interior)
        dw_rows<K, S, false>(acc, in4, wsm_cg, C8, W, H, ix0, iy0);
      else
        dw_rows<K, S, true>(acc, in4, wsm_cg, C8, W, H, ix0, iy0);
#pragma unroll
      for (int p = 0; p < kDwP; ++p) {
        if (ox0 + p < Wo) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            y[2 * e] = acc[p][e].x;
            y[2 * e + 1] = acc[p][e].y;
          }
          if (act == OCTSEG_ACT_SWISH) {  // uniform branch: only one activation's code runs
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float h = 0.5f * y[e];
              float t;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
              y[e] = fmaf(h, t, h);
            }
          } else if (act == OCTSEG_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
          } else if (act != OCTSEG_ACT_NONE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = act_f(y[e], act);
          }
          const uint4 o = pack8(y);
          out4[(static_cast<size_t>(oy) * Wo + ox0 + p) * C8] = o;
          if (pool_sum) {  // pool what the next layer actually reads (the bf16-rounded activation)
            float r[8];
            unpack8(o, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) ps[e] += r[e];
          }
        }
      }
    }
    if (pool_sum) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&sums[e][cgl], ps[e]);
    }
  }
  if (pool_sum) {
    __syncthreads();
    for (int i = threadIdx.x; i < cgc * 8; i += 256)
      atomicAdd(pool_sum + static_cast<size_t>(n) * C + cg0 * 8 + i, sums[i & 7][i >> 3]);
  }
}

// ------------------------------------------------------------------------------------ SE gate
// hidden[n][r] = swish(W1[r,:] . mean[n,:] + b1[r]); one warp per (image, hidden unit)
__global__ void __launch_bounds__(256) se_hidden_kernel(const float* __restrict__ pool_sum, float inv_hw,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        float* __restrict__ hidden, int C, int Cr) {
  const int n = blockIdx.y;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= Cr) return;
  const float* m = pool_sum + static_cast<size_t>(n) * C;
  const float* w = w1 + static_cast<size_t>(r) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w + c), __ldg(m + c), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    s = s * inv_hw + b1[r];
    hidden[static_cast<size_t>(n) * Cr + r] = s / (1.f + __expf(-s));
  }
}

// gate[n][k] = sigmoid(W2[k,:] . hidden[n,:] + b2[k]) for this block's 256 input channels k, then
// out[n][row][k] = bf16(w[row][k] * gate[n][k]) for every weight row (0 for the K padding).
__global__ void __launch_bounds__(256) se_scale_weights_kernel(const float* __restrict__ hidden,
                                                               const float* __restrict__ w2t,  // [Cr][C]
                                                               const float* __restrict__ b2,
                                                               const float* __restrict__ w,    // [rows][Ktot]
                                                               __nv_bfloat16* __restrict__ out, int rows, int Ktot,
                                                               int C, int Cr) {
  extern __shared__ float hid[];
  const int n = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  for (int r = threadIdx.x; r < Cr; r += 256) hid[r] = hidden[static_cast<size_t>(n) * Cr + r];
  __syncthreads();
  if (k >= Ktot) return;
  float g = 0.f;
  if (k < C) {
    float s = b2[k];
    for (int r = 0; r < Cr; ++r) s = fmaf(__ldg(w2t + static_cast<size_t>(r) * C + k), hid[r], s);
    g = 1.f / (1.f + __expf(-s));
  }
  __nv_bfloat16* o = out + static_cast<size_t>(n) * rows * Ktot + k;
  const float* wk = w + k;
  const int per = (rows + gridDim.z - 1) / gridDim.z;  // rows are split over blockIdx.z for parallelism
  const int r0 = blockIdx.z * per, r1 = min(rows, r0 + per);
#pragma unroll 4
  for (int row = r0; row < r1; ++row)
    o[static_cast<size_t>(row) * Ktot] = __float2bfloat16_rn(__ldg(wk + static_cast<size_t>(row) * Ktot) * g);
}

static int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 32;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace octseg

using namespace octseg;

extern "C" int octseg_stem_conv(const void* in, int32_t in_dtype, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                int32_t N, int32_t H, int32_t W, const float* weight, const float* bias,
                                int32_t Cout, int32_t k, int32_t stride, int32_t pad_t, int32_t pad_l, int32_t Ho,
                                int32_t Wo, int32_t act, const float* h_mean, const float* h_inv_std, void* out,
                                int32_t out_ldc, void* stream) {
  if (Cout != 32 && Cout != 64) return fail(OCTSEG_EINVAL, "stem conv supports Cout 32 or 64, got %d", Cout);
  if (k > 7 || stride > 2 || out_ldc % 8) return fail(OCTSEG_EINVAL, "stem conv: k<=7, stride<=2, out_ldc%%8==0");
  if (in_dtype != 0 && in_dtype != 1) return fail(OCTSEG_EINVAL, "stem conv: in_dtype must be 0 (f32) or 1 (u8)");
  StemParams p;
  p.in = in;
  p.in_dtype = in_dtype;
  p.sn = sn;
  p.sc = sc;
  p.sh = sh;
  p.sw = sw;
  p.N = N;
  p.H = H;
  p.W = W;
  p.weight = weight;
  p.bias = bias;
  p.k = k;
  p.stride = stride;
  p.pad_t = pad_t;
  p.pad_l = pad_l;
  p.Ho = Ho;
  p.Wo = Wo;
  p.act = act;
  p.normalize = (h_mean && h_inv_std) ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = p.normalize ? h_mean[c] : 0.f;
    p.inv_std[c] = p.normalize ? h_inv_std[c] : 1.f;
  }
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_ldc = out_ldc;
  const int pdim = 15 * stride + k;
  const size_t smem = (static_cast<size_t>(k) * k * 3 * Cout + static_cast<size_t>(pdim) * pdim * 3) * sizeof(float);
  dim3 grid(cdiv(Wo, 16), cdiv(Ho, 16), N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;  // once per process; not a stream operation, keep it out of graph capture
  if (!attr_set) {
    OCTSEG_CUDA(cudaFuncSetAttribute(stem_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    OCTSEG_CUDA(cudaFuncSetAttribute(stem_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  if (Cout == 64)
    stem_conv_kernel<64><<<grid, 256, smem, st>>>(p);
  else
    stem_conv_kernel<32><<<grid, 256, smem, st>>>(p);
  return check_launch("stem_conv_kernel");
}

extern "C" int octseg_maxpool3x3s2(const void* in, void* out, int32_t N, int32_t H, int32_t W, int32_t C, int32_t Ho,
                                   int32_t Wo, void* stream) {
  if (C % 8) return fail(OCTSEG_EINVAL, "maxpool: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(N) * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), N, H, W, C / 8, Ho, Wo);
  return check_launch("maxpool3x3s2_kernel");
}

extern "C" int octseg_dwconv(const void* in, const void* weight, const float* bias, void* out, int32_t N, int32_t H,
                             int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad_t, int32_t pad_l,
                             int32_t Ho, int32_t Wo, int32_t act, float* pool_sum, void* stream) {
  if (C % 8) return fail(OCTSEG_EINVAL, "dwconv: C must be a multiple of 8 (C=%d)", C);
  const int C8 = C / 8;
  const int npg = Ho * cdiv(Wo, kDwP);
  const int cols = cdiv(C8, kDwCgChunk);
  // enough blocks for >= 8 per SM on small feature maps, long strips on large ones
  int pgb = 64;
  while (pgb > 8 && static_cast<long long>(cols) * cdiv(npg, pgb) * N < 148 * 8) pgb >>= 1;
  dim3 grid(cols, cdiv(npg, pgb), N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* i = static_cast<const __nv_bfloat16*>(in);
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(weight);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
#define OCTSEG_DW(KK, SS) \
  dwconv_kernel<KK, SS><<<grid, 256, 0, st>>>(i, w, bias, o, H, W, C, pad_t, pad_l, Ho, Wo, pgb, act, pool_sum)
  if (k == 3 && stride == 1) OCTSEG_DW(3, 1);
  else if (k == 3 && stride == 2) OCTSEG_DW(3, 2);
  else if (k == 5 && stride == 1) OCTSEG_DW(5, 1);
  else if (k == 5 && stride == 2) OCTSEG_DW(5, 2);
  else return fail(OCTSEG_EINVAL, "dwconv: unsupported kernel %d / stride %d", k, stride);
#undef OCTSEG_DW
  return check_launch("dwconv_kernel");
}

extern "C" int octseg_se_hidden(const float* pool_sum, float inv_hw, const float* w1, const float* b1, float* hidden,
                                int32_t N, int32_t C, int32_t Cr, void* stream) {
  dim3 grid(cdiv(Cr, 8), N);
  se_hidden_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pool_sum, inv_hw, w1, b1, hidden, C, Cr);
  return check_launch("se_hidden_kernel");
}

extern "C" int octseg_se_scale_weights(const float* hidden, const float* w2t, const float* b2, const float* w, void* out,
                                       int32_t N, int32_t rows, int32_t Ktot, int32_t C, int32_t Cr, void* stream) {
  if (Cr > 4096) return fail(OCTSEG_EINVAL, "se_scale_weights: Cr too large");
  int zc = 1;  // enough blocks to fill the machine (the gate is recomputed per row chunk, it is tiny)
  while (zc < 16 && cdiv(Ktot, 256) * N * zc < 148 * 4 && rows / (zc * 2) >= 8) zc *= 2;
  dim3 grid(cdiv(Ktot, 256), N, zc);
  se_scale_weights_kernel<<<grid, 256, static_cast<size_t>(Cr) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      hidden, w2t, b2, w, static_cast<__nv_bfloat16*>(out), rows, Ktot, C, Cr);
  return check_launch("se_scale_weights_kernel");
}
