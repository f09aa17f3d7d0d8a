// CUDA-core kernels for the layers that are HBM-bound or have a tiny contraction:
// network stems (Cin = 3), max-pool, SE gate, per-image SE-scaled projection weights
// (the depthwise conv lives in dwconv.cu).  Activations are NHWC bf16, 8 channels (16 bytes)
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
