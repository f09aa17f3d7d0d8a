// CUDA-core kernels for the layers that are HBM-bound or have a tiny contraction:
// network-stem input packing, max-pool, SE gate, per-image SE-scaled projection weights
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

// ------------------------------------------------------------------------------------ stem input
// Network stems are stride-2 convs on the 3-channel input.  They run on the tensor cores as a
// stride-1 conv over the space-to-depth form of the input: one launch of this kernel reads the frame
// in place (uint8 NHWC or the NHWC-strided fp32 tensor of the reference's predict()), applies
// OCTSegmentationModel.forward's normalisation, and writes bf16 [N][H/2][W/2][16] with channel
// (dy*2 + dx)*3 + c = input pixel (2y+dy, 2x+dx), channel c; channels 12..15 are zero.  A 2x2 block of
// pixels is 32 bytes = one TMA row of the kc=16 K-segment of conv_tc_kernel.
struct StemPackParams {
  const void* in;
  int in_dtype;
  long long sn, sc, sh, sw;
  int N, H2, W2;
  float mean[3], inv_std[3];
  uint4* out;
};

__global__ void __launch_bounds__(256) stem_pack_s2d_kernel(const StemPackParams p) {
  const size_t total = static_cast<size_t>(p.N) * p.H2 * p.W2;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % p.W2);
    size_t t = idx / p.W2;
    const int y = static_cast<int>(t % p.H2);
    const int n = static_cast<int>(t / p.H2);
    float v[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const long long off = n * p.sn + c * p.sc + (2 * y + dy) * p.sh + (2 * x + dx) * p.sw;
          const float raw = p.in_dtype == 0 ? reinterpret_cast<const float*>(p.in)[off]
                                            : static_cast<float>(reinterpret_cast<const uint8_t*>(p.in)[off]);
          v[(dy * 2 + dx) * 3 + c] = (raw - p.mean[c]) * p.inv_std[c];
        }
#pragma unroll
    for (int e = 12; e < 16; ++e) v[e] = 0.f;
    float lo[8], hi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      lo[e] = v[e];
      hi[e] = v[8 + e];
    }
    p.out[2 * idx] = pack8(lo);
    p.out[2 * idx + 1] = pack8(hi);
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

extern "C" int octseg_stem_pack(const void* in, int32_t in_dtype, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                int32_t N, int32_t H, int32_t W, const float* h_mean, const float* h_inv_std, void* out,
                                void* stream) {
  if (in_dtype != 0 && in_dtype != 1) return fail(OCTSEG_EINVAL, "stem pack: in_dtype must be 0 (f32) or 1 (u8)");
  if (H % 2 || W % 2 || H < 2 || W < 2) return fail(OCTSEG_EINVAL, "stem pack: H and W must be even (got %dx%d)", H, W);
  if (reinterpret_cast<uintptr_t>(out) & 15) return fail(OCTSEG_EINVAL, "stem pack: out must be 16-byte aligned");
  StemPackParams p;
  p.in = in;
  p.in_dtype = in_dtype;
  p.sn = sn;
  p.sc = sc;
  p.sh = sh;
  p.sw = sw;
  p.N = N;
  p.H2 = H / 2;
  p.W2 = W / 2;
  const bool norm = h_mean && h_inv_std;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = norm ? h_mean[c] : 0.f;
    p.inv_std[c] = norm ? h_inv_std[c] : 1.f;
  }
  p.out = static_cast<uint4*>(out);
  const size_t total = static_cast<size_t>(N) * p.H2 * p.W2;
  stem_pack_s2d_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("stem_pack_s2d_kernel");
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
