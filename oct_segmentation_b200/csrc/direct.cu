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
// Squeeze-excite of one MBConv block = two tiny dense layers on the pooled activation + a per-image
// rescale of the projection weights.  All three steps are latency-bound, so each kernel is laid out for
// many independent loads in flight rather than for FLOPs.
//
// Step 0 (maps with more than one slot): sums[n][c] = sum_s pool[n][s][c] in slot order.  pool holds the depthwise
// kernel's write-once partial sums (one slot per row group of tiles, up to 56 for a 448 x 448 map); adding them in a
// fixed order makes the pooled mean bit-reproducible (the first version accumulated with fp32 atomics).  One thread
// per (image, 4 channels): slots independent 16-byte loads in flight.
__global__ void __launch_bounds__(256) se_pool_reduce_kernel(const float* __restrict__ pool, float* __restrict__ sums,
                                                             int N, int slots, int C4) {
  const size_t total = static_cast<size_t>(N) * C4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = idx / C4, c4 = idx - n * C4;
    const float4* p = reinterpret_cast<const float4*>(pool) + n * slots * C4 + c4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int sl = 0; sl < slots; ++sl) {
      const float4 v = __ldg(p + static_cast<size_t>(sl) * C4);
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      a.w += v.w;
    }
    reinterpret_cast<float4*>(sums)[idx] = a;
  }
}

// hidden[n][r] = swish(W1[r,:] . mean[n,:] + b1[r]): one block per hidden unit r computes it for ALL images,
// so W1's row is read once and every thread has up to 1 + N independent loads per step.
constexpr int kSeMaxN = 8;  // images per block of se_hidden_kernel (blockIdx.y = image group)
__global__ void __launch_bounds__(256) se_hidden_kernel(const float* __restrict__ pool_sum, float inv_hw,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        float* __restrict__ hidden, int N, int C, int Cr) {
  __shared__ float part[8][kSeMaxN];
  const int r = blockIdx.x;
  const int n0 = blockIdx.y * kSeMaxN;
  const int nn = min(kSeMaxN, N - n0);
  const float* w = w1 + static_cast<size_t>(r) * C;
  const float* m = pool_sum + static_cast<size_t>(n0) * C;
  float s[kSeMaxN];
#pragma unroll
  for (int n = 0; n < kSeMaxN; ++n) s[n] = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float wv = __ldg(w + c);
#pragma unroll
    for (int n = 0; n < kSeMaxN; ++n)
      if (n < nn) s[n] = fmaf(wv, __ldg(m + static_cast<size_t>(n) * C + c), s[n]);
  }
#pragma unroll
  for (int n = 0; n < kSeMaxN; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[n] += __shfl_xor_sync(0xffffffffu, s[n], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < kSeMaxN; ++n) part[warp][n] = s[n];
  }
  __syncthreads();
  if (threadIdx.x < nn) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += part[w8][threadIdx.x];
    t = t * inv_hw + b1[r];
    hidden[static_cast<size_t>(n0 + threadIdx.x) * Cr + r] = t / (1.f + __expf(-t));
  }
}

// gate[n][k] = sigmoid(W2[k,:] . hidden[n,:] + b2[k]): 64 channels x 4 slices of the hidden units per block
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ hidden,
                                                      const float* __restrict__ w2t,  // [Cr][C]
                                                      const float* __restrict__ b2, float* __restrict__ gate,
                                                      float* __restrict__ pool_clear, int C, int Cr) {
  extern __shared__ float se_smem[];  // hid[Cr], then part[4][64]
  float* hid = se_smem;
  float* part = se_smem + Cr;
  const int n = blockIdx.y;
  const int kl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + kl;
  for (int r = threadIdx.x; r < Cr; r += 256) hid[r] = hidden[static_cast<size_t>(n) * Cr + r];
  __syncthreads();
  float s = 0.f;
  if (k < C) {
#pragma unroll 8
    for (int r = slice; r < Cr; r += 4) s = fmaf(__ldg(w2t + static_cast<size_t>(r) * C + k), hid[r], s);
  }
  part[slice * 64 + kl] = s;
  __syncthreads();
  if (slice == 0 && k < C) {
    const float t = part[kl] + part[64 + kl] + part[128 + kl] + part[192 + kl] + b2[k];
    gate[static_cast<size_t>(n) * C + k] = 1.f / (1.f + __expf(-t));
    if (pool_clear) pool_clear[static_cast<size_t>(n) * C + k] = 0.f;  // consumed by se_hidden: ready for the next frame batch
  }
}

// out[n][row][k..k+7] = bf16(w[row][k..k+7] * gate[n][k..k+7]) (0 for the K padding k >= C): the per-image
// weights of the projection conv.  Work item = 8 channels of one weight row of one image (32-byte fp32
// read, 16-byte bf16 write), items flattened so consecutive threads write consecutive 16-byte pieces.
__global__ void __launch_bounds__(256) se_scale_weights_kernel(const float* __restrict__ gate,
                                                               const float* __restrict__ w,  // [rows][Ktot]
                                                               __nv_bfloat16* __restrict__ out, int N, int rows,
                                                               int Ktot, int C, int period) {
  const int k8n = Ktot >> 3;
  const size_t per_image = static_cast<size_t>(rows) * k8n;
  const size_t total = per_image * N;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / per_image);
    const size_t rem = idx - static_cast<size_t>(n) * per_image;  // row * k8n + k8
    const int k = static_cast<int>(rem % k8n) * 8;
    float f[8];
    if (k < C) {  // C and Ktot are multiples of 8: a group is entirely real or entirely padding
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(w + rem * 8));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(w + rem * 8 + 4));
      const int kg = k % period;  // pixel-packed problems repeat the channel block `period` along K
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(n) * period + kg));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(n) * period + kg + 4));
      f[0] = a0.x * g0.x;
      f[1] = a0.y * g0.y;
      f[2] = a0.z * g0.z;
      f[3] = a0.w * g0.w;
      f[4] = a1.x * g1.x;
      f[5] = a1.y * g1.y;
      f[6] = a1.z * g1.z;
      f[7] = a1.w * g1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = 0.f;
    }
    reinterpret_cast<uint4*>(out)[idx] = pack8(f);
  }
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

extern "C" int octseg_se_hidden(const float* pool_sum, int32_t slots, float* sums_scratch, float inv_hw, const float* w1,
                                const float* b1, float* hidden, int32_t N, int32_t C, int32_t Cr, void* stream) {
  if (slots < 1) return fail(OCTSEG_EINVAL, "se_hidden: slots must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* sums = pool_sum;
  if (slots > 1) {
    if (!sums_scratch || C % 4 || (reinterpret_cast<uintptr_t>(pool_sum) & 15) || (reinterpret_cast<uintptr_t>(sums_scratch) & 15))
      return fail(OCTSEG_EINVAL, "se_hidden: slots > 1 needs a 16-byte aligned fp32 [N][C] scratch buffer and C %% 4 == 0");
    se_pool_reduce_kernel<<<grid_for(static_cast<size_t>(N) * (C / 4), 256), 256, 0, st>>>(pool_sum, sums_scratch, N, slots, C / 4);
    const int rc = check_launch("se_pool_reduce_kernel");
    if (rc) return rc;
    sums = sums_scratch;
  }
  dim3 grid(Cr, cdiv(N, kSeMaxN));
  se_hidden_kernel<<<grid, 256, 0, st>>>(sums, inv_hw, w1, b1, hidden, N, C, Cr);
  return check_launch("se_hidden_kernel");
}

extern "C" int octseg_se_gate(const float* hidden, const float* w2t, const float* b2, float* gate, float* pool_clear,
                              int32_t N, int32_t C, int32_t Cr, void* stream) {
  if (Cr > 8192) return fail(OCTSEG_EINVAL, "se_gate: Cr too large");
  dim3 grid(cdiv(C, 64), N);
  se_gate_kernel<<<grid, 256, (static_cast<size_t>(Cr) + 256) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      hidden, w2t, b2, gate, pool_clear, C, Cr);
  return check_launch("se_gate_kernel");
}

extern "C" int octseg_se_scale_weights(const float* gate, const float* w, void* out, int32_t N, int32_t rows,
                                       int32_t Ktot, int32_t C, int32_t gate_period, void* stream) {
  if (Ktot % 8 || C % 8 || C > Ktot) return fail(OCTSEG_EINVAL, "se_scale_weights: C and Ktot must be multiples of 8, C <= Ktot");
  if (gate_period <= 0) gate_period = C;
  if (gate_period % 8 || C % gate_period) return fail(OCTSEG_EINVAL, "se_scale_weights: gate_period must be a multiple of 8 dividing C");
  if ((reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(gate) & 15))
    return fail(OCTSEG_EINVAL, "se_scale_weights: pointers must be 16-byte aligned");
  const size_t total = static_cast<size_t>(N) * rows * (Ktot / 8);
  se_scale_weights_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gate, w, static_cast<__nv_bfloat16*>(out), N, rows, Ktot, C, gate_period);
  return check_launch("se_scale_weights_kernel");
}
