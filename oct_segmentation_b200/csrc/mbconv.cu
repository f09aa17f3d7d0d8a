// Fused MBConv front half for sm_100a: expand 1x1 conv (+ folded BN, swish) -> depthwise k x k conv (+ folded BN,
// swish) -> squeeze-excite channel sums, in ONE kernel.  The 6x-wide expanded tensor of efficientnet_pytorch's
// MBConvBlock (_expand_conv/_bn0/swish -> _depthwise_conv/_bn1/swish) never touches HBM: per spatial tile the
// expanded activations are produced by tcgen05 MMAs into TMEM, converted to bf16 into shared memory and consumed
// there by the depthwise tile math (csrc/dw_core.h).  Reads x once per tile, writes the depthwise output once.
//
// Work item: one (image, 8 x 16 output tile); the CTA loops over all 64-channel blocks of the expanded width with
// the x halo tile resident in shared memory.  Per (tile, channel block):
//   GEMM   E[p, c] = sum_k x[p, k] * We[c, k]     p = halo pixel (row of M = 128-row MMA tiles), K = Cin in
//          16-channel chunks (32-byte swizzled rows, one TMA box per chunk), N = 64 expanded channels
//   CONV   TMEM -> registers -> h = (E + b)/2 -> swish = h*tanh(h) + h -> bf16 -> shared tile [pixel][64 ch]
//          (144-byte pixel pitch: conflict-free 16-byte stores by 32 consecutive pixels); halo pixels outside the
//          image are written as ZERO (the depthwise conv pads the EXPANDED tensor, and swish(b) != 0)
//   DW     2 x 4 output pixels x 4 channels per thread, FFMA2, swish, bf16 store, SE sums
// The per-block constants (expand bias, depthwise bias and filters as fp32, pre-halved for the swish form
// h*tanh(h)+h with h = x/2) come as one packed blob per 64-channel block and travel with the We block through a
// two-slot ring filled by bulk copies: the compute warps never load or convert a weight.
// Roles: warps 0-7 convert + run the depthwise math; warp 8 (one lane) issues TMA and the MMAs.  The MMAs of channel
// block i+1 run on the tensor pipe while the CUDA cores do the depthwise math of block i (one TMEM accumulator:
// it is free again as soon as block i has been converted); up to two CTAs per SM overlap the rest.
//
// Precision is that of the unfused path: bf16 weights, fp32 accumulation, the expanded activation rounded to bf16
// before the depthwise conv, fp32 depthwise math, SE sums of the fp32 activation.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "common.h"
#include "dw_core.h"
#include "ptx_sm100.h"

namespace octseg {

constexpr int kMbComputeThreads = 256;              // 8 warps: conversion + depthwise math
constexpr int kMbThreads = kMbComputeThreads + 32;  // + 1 warp: TMA producer / MMA issuer
constexpr int kMbCB = 64;                           // expanded channels per block (MMA N)
constexpr int kMbPixE = 144;                        // bytes per pixel of the expanded tile in shared memory

template <int K, int S>
struct MbCfg {
  static constexpr int SH = 4, SW = 4;
  static constexpr int TH = SH * kDwR, TW = SW * kDwP;                  // 8 x 16 output pixels
  static constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;     // halo tile of the expanded tensor
  static constexpr int NPIX = IH * IW;
  static constexpr int MT = (NPIX + 127) / 128;                          // 128-row MMA tiles
  static constexpr int TMEM_COLS = MT * kMbCB <= 128 ? 128 : (MT * kMbCB <= 256 ? 256 : 512);
  static_assert(MT * kMbCB <= 512, "accumulator does not fit TMEM");
  static constexpr int XCH = MT * 128 * 32;                              // bytes of one 16-channel chunk of the x tile
  static constexpr int E_BYTES = (NPIX * kMbPixE + 1023) / 1024 * 1024;
  static constexpr int BLOB = (2 + K * K) * kMbCB * 4;                   // per-block fp32 constants: b_exp/2, b_dw/2, K*K filters/2
  static constexpr int CTAS = S == 1 ? 2 : 1;
};

struct MbParams {
  const float* blob;            // [ncb][2 + K*K][64] fp32: b_exp/2, b_dw/2, w_dw/2 per 64-channel block (zero padded)
  __nv_bfloat16* out;           // [N][Ho][Wo][Cmid]
  float* pool_sum;              // [N][tiles_h * tiles_w][Cmid] per-tile partial sums (summed by octseg_se_hidden) or null
  int H, W, Cmid, Ho, Wo, pad_t, pad_l;
  int kch;                      // Cin / 16
  int ncb;                      // ceil(Cmid / 64)
  int tiles_w, tiles_h, total_tiles;
  FastDiv fd_tw, fd_th, fd_iw;
};

__device__ __forceinline__ void mb_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mb_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mb_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mb_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mb_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// K-major shared-memory matrix descriptor, 32-byte swizzle: rows are 16 bf16 (32 B), 8-row groups 256 B apart
__device__ __forceinline__ uint64_t mb_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;          // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(256 >> 4) << 32;   // stride byte offset: 8 rows * 32 B
  d |= static_cast<uint64_t>(1) << 46;          // descriptor version (sm_100)
  d |= static_cast<uint64_t>(6) << 61;          // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void mb_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mb_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
// producer-side wait: backs off between polls so the spinning lane does not eat the issue slots of its scheduler
__device__ __forceinline__ void mb_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(64);
    if (++spins > (kSpinLimit >> 4)) __trap();
  }
}
__device__ __forceinline__ void mb_compute_bar() {  // the 256 compute threads only
  asm volatile("bar.sync 1, %0;" ::"n"(kMbComputeThreads) : "memory");
}

template <int K, int S>
__global__ void __launch_bounds__(kMbThreads, MbCfg<K, S>::CTAS)
    mbconv_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const MbParams p) {
  using Cfg = MbCfg<K, S>;
  constexpr int R = kDwR, P = kDwP, CB = kMbCB, LANES = CB / 4, IW = Cfg::IW, MT = Cfg::MT, NPIX = Cfg::NPIX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  // layout: x tile [kch][MT*128 rows][32 B] | 2 slots x { We block [kch][64 rows][32 B], constants blob } | E tile |
  //         SE partials [8 warps][64] | barriers | tmem slot
  const uint32_t sX = smem0;
  const uint32_t b_bytes = static_cast<uint32_t>(p.kch) * (CB * 32);
  const uint32_t slot_bytes = b_bytes + Cfg::BLOB;                    // multiple of 256
  const uint32_t sSlot = sX + static_cast<uint32_t>(p.kch) * Cfg::XCH;
  const uint32_t sE = (sSlot + 2 * slot_bytes + 1023u) & ~1023u;
  const uint32_t sPart = sE + Cfg::E_BYTES;
  const uint32_t sBar = sPart + 8 * CB * 4;
  float* part = reinterpret_cast<float*>(smem_gen + (sPart - smem0));
  const uint32_t bar_x_full = sBar, bar_x_free = sBar + 8, bar_s_full = sBar + 16 /* [2] */, bar_s_free = sBar + 32 /* [2] */;
  const uint32_t bar_acc_full = sBar + 48, bar_acc_empty = sBar + 56, tmem_slot = sBar + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = (p.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (tid == 0) {
    mbar_init(bar_x_full, 1);
    mbar_init(bar_x_free, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_full + 8, 1);
    mbar_init(bar_s_free, kMbComputeThreads / 32);
    mbar_init(bar_s_free + 8, kMbComputeThreads / 32);
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, kMbComputeThreads / 32);
    mbar_fence_init();
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  mb_fence_before();
  __syncthreads();
  mb_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode = [&](int i, int& n, int& th, int& tw) {  // i-th tile of this CTA; tw fastest, then th, then image
    const uint32_t t = blockIdx.x + static_cast<uint32_t>(i) * gridDim.x;
    const uint32_t q = fd_div(t, p.fd_tw);
    tw = static_cast<int>(t - q * p.fd_tw.d);
    const uint32_t nn = fd_div(q, p.fd_th);
    th = static_cast<int>(q - nn * p.fd_th.d);
    n = static_cast<int>(nn);
  };

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer + MMA issuer (one lane)
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(CB >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      const uint32_t x_bytes = static_cast<uint32_t>(p.kch) * (NPIX * 32);
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int n, th, tw;
        decode(i, n, th, tw);
        if (i > 0) mb_wait_sleep(bar_x_free, static_cast<uint32_t>((i - 1) & 1));  // every MMA that read the old x tile is done
        mbar_arrive_expect_tx(bar_x_full, x_bytes);
        for (int j = 0; j < p.kch; ++j)
          tma_load_4d(sX + static_cast<uint32_t>(j) * Cfg::XCH, &tm_x, bar_x_full, j * 16, tw * Cfg::TW * S - p.pad_l,
                      th * Cfg::TH * S - p.pad_t, n);
        for (int cb = 0; cb < p.ncb; ++cb, ++it) {
          const uint32_t slot = it & 1u;
          const uint32_t sb = sSlot + slot * slot_bytes;
          // the slot's previous tenant (iteration it - 2) is free once its depthwise math is done, which implies
          // its conversion and therefore its MMAs are done too
          if (it >= 2) mb_wait_sleep(bar_s_free + 8 * slot, ((it >> 1) - 1) & 1u);
          mbar_arrive_expect_tx(bar_s_full + 8 * slot, slot_bytes);
          for (int j = 0; j < p.kch; ++j)
            mb_tma_load_2d(sb + static_cast<uint32_t>(j) * (CB * 32), &tm_w, bar_s_full + 8 * slot, j * 16, cb * CB);
          mb_bulk_load(sb + b_bytes, reinterpret_cast<const uint8_t*>(p.blob) + static_cast<size_t>(cb) * Cfg::BLOB, Cfg::BLOB,
                       bar_s_full + 8 * slot);
          if (it >= 1) mb_wait_sleep(bar_acc_empty, (it - 1) & 1u);  // the previous block has left the accumulator
          if (cb == 0) mb_wait_sleep(bar_x_full, static_cast<uint32_t>(i & 1));
          mb_wait_sleep(bar_s_full + 8 * slot, (it >> 1) & 1u);
          mb_fence_after();
#pragma unroll 1
          for (int m = 0; m < MT; ++m) {
            for (int j = 0; j < p.kch; ++j) {
              const uint64_t adesc = mb_desc_sw32(sX + static_cast<uint32_t>(j) * Cfg::XCH + static_cast<uint32_t>(m) * (128 * 32));
              const uint64_t bdesc = mb_desc_sw32(sb + static_cast<uint32_t>(j) * (CB * 32));
              mb_mma(tmem_base + static_cast<uint32_t>(m * CB), adesc, bdesc, idesc, j > 0 ? 1u : 0u);
            }
          }
          mb_commit(bar_acc_full);
          if (cb == p.ncb - 1) mb_commit(bar_x_free);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ conversion + depthwise math (8 warps)
    const int lane_c = tid % LANES, slot_t = tid / LANES;
    const int sy = slot_t / Cfg::SW, sx = slot_t - sy * Cfg::SW;
    const uint32_t in_off = static_cast<uint32_t>(((sy * R * S) * IW + sx * P * S) * kMbPixE + lane_c * 8);
    const int q4 = warp & 3, hf = warp >> 2;   // TMEM lane quarter of this warp; which 32 of the 64 columns it converts
    const uint32_t pix_bytes = static_cast<uint32_t>(p.Cmid) * 2u;
    const size_t row_bytes = static_cast<size_t>(p.Wo) * pix_bytes;
    uint32_t it = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int n, th, tw;
      decode(i, n, th, tw);
      const int h0 = th * Cfg::TH * S - p.pad_t, w0 = tw * Cfg::TW * S - p.pad_l;
      for (int cb = 0; cb < p.ncb; ++cb, ++it) {
        const uint32_t slot = it & 1u;
        const uint32_t s_blob = sSlot + slot * slot_bytes + b_bytes;  // [b_exp/2 | b_dw/2 | K*K filters/2] x 64 channels, fp32
        const float* bexp = reinterpret_cast<const float*>(smem_gen + (s_blob - smem0));
        mbar_wait(bar_s_full + 8 * slot, (it >> 1) & 1u);             // this block's constants have landed

        // ---- conversion: TMEM accumulator -> swish -> bf16 -> E tile
        mbar_wait(bar_acc_full, it & 1u);
        mb_fence_after();
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m * 128 + q4 * 32 >= NPIX) continue;  // warp-uniform: no halo pixel in this warp's rows of the tile
          uint32_t v[32];
          mb_tmem_ld32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(m * CB + hf * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int pix = m * 128 + q4 * 32 + lane;
          if (pix < NPIX) {
            const int iy = static_cast<int>(fd_div(static_cast<uint32_t>(pix), p.fd_iw)), ix = pix - iy * IW;
            const int gy = h0 + iy, gx = w0 + ix;
            const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
            const uint32_t dst = sE + static_cast<uint32_t>(pix) * kMbPixE + static_cast<uint32_t>(hf * 64);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 bb = *reinterpret_cast<const float2*>(bexp + hf * 32 + g * 8 + e * 2);
                const float2 a = make_float2(__uint_as_float(v[g * 8 + e * 2]), __uint_as_float(v[g * 8 + e * 2 + 1]));
                const float2 h = __ffma2_rn(a, make_float2(0.5f, 0.5f), bb);              // (acc + b) / 2
                const float2 y = __ffma2_rn(h, make_float2(dw_tanh(h.x), dw_tanh(h.y)), h);  // swish
                o[e] = inside ? dw_cvt_bf16x2(y) : 0u;
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + static_cast<uint32_t>(g * 16)), "r"(o[0]),
                           "r"(o[1]), "r"(o[2]), "r"(o[3])
                           : "memory");
            }
          }
        }
        mb_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty);  // the MMAs of the next block may overwrite the accumulator
        mb_compute_bar();                           // E tile complete

        // ---- depthwise k x k on the E tile
        const int c0 = cb * CB + lane_c * 4;
        const bool cvalid = c0 < p.Cmid;
        float2 bias2[2];
        {
          const float4 b = *reinterpret_cast<const float4*>(bexp + CB + lane_c * 4);
          bias2[0] = make_float2(b.x, b.y);
          bias2[1] = make_float2(b.z, b.w);
        }
        float2 acc[R][P][2];
        dw_patch<K, S, CB, IW, kMbPixE>(sE + in_off, s_blob + 2 * CB * 4 + static_cast<uint32_t>(lane_c * 16), bias2,
                                        OCTSEG_ACT_SWISH, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_s_free + 8 * slot);  // constants (and, by implication, the We block) of this slot are dead
        const int oy0 = th * Cfg::TH + sy * R, ox0 = tw * Cfg::TW + sx * P;
        uint8_t* o0 = reinterpret_cast<uint8_t*>(p.out) +
                      static_cast<size_t>((static_cast<size_t>(n) * p.Ho + oy0) * p.Wo + ox0) * pix_bytes + static_cast<size_t>(c0) * 2;
        float2 ps[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (cvalid && oy0 + R <= p.Ho && ox0 + P <= p.Wo) {  // interior patch: no per-pixel checks
#pragma unroll
          for (int r = 0; r < R; ++r) {
            uint8_t* orow = o0 + r * row_bytes;
#pragma unroll
            for (int q = 0; q < P; ++q) {
              *reinterpret_cast<uint2*>(orow + q * pix_bytes) = make_uint2(dw_cvt_bf16x2(acc[r][q][0]), dw_cvt_bf16x2(acc[r][q][1]));
              ps[0] = __fadd2_rn(ps[0], acc[r][q][0]);
              ps[1] = __fadd2_rn(ps[1], acc[r][q][1]);
            }
          }
        } else if (cvalid) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int q = 0; q < P; ++q) {
              if (oy0 + r < p.Ho && ox0 + q < p.Wo) {
                *reinterpret_cast<uint2*>(o0 + r * row_bytes + q * pix_bytes) =
                    make_uint2(dw_cvt_bf16x2(acc[r][q][0]), dw_cvt_bf16x2(acc[r][q][1]));
                ps[0] = __fadd2_rn(ps[0], acc[r][q][0]);
                ps[1] = __fadd2_rn(ps[1], acc[r][q][1]);
              }
            }
          }
        }
        // SE sums of this (tile, block): registers -> warp shuffle -> shared -> one plain store per channel into the
        // tile's own slot (bit-reproducible: octseg_se_hidden adds the slots in order)
        if (p.pool_sum) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            ps[e].x += __shfl_xor_sync(0xffffffffu, ps[e].x, 16);
            ps[e].y += __shfl_xor_sync(0xffffffffu, ps[e].y, 16);
          }
          if (lane < LANES) {
            float* d = part + warp * CB + lane_c * 4;
            d[0] = ps[0].x;
            d[1] = ps[0].y;
            d[2] = ps[1].x;
            d[3] = ps[1].y;
          }
        }
        mb_compute_bar();  // E tile free for the next conversion; SE partials visible
        if (p.pool_sum && tid < CB && cb * CB + tid < p.Cmid) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += part[w * CB + tid];
          p.pool_sum[(static_cast<size_t>(n) * (p.tiles_h * p.tiles_w) + th * p.tiles_w + tw) * p.Cmid + cb * CB + tid] = sum;
        }
      }
    }
  }

  mb_fence_before();
  __syncthreads();
  mb_fence_after();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

template <int K, int S>
static size_t mb_smem_bytes(int kch) {
  using Cfg = MbCfg<K, S>;
  return static_cast<size_t>(kch) * Cfg::XCH + 2u * (static_cast<size_t>(kch) * (kMbCB * 32) + Cfg::BLOB) + 1024 + Cfg::E_BYTES +
         8 * kMbCB * 4 + 128 + 1024;
}

template <int K, int S>
static int launch_mb(const CUtensorMap& tm_x, const CUtensorMap& tm_w, MbParams p, int N, cudaStream_t st) {
  using Cfg = MbCfg<K, S>;
  p.tiles_w = cdiv(p.Wo, Cfg::TW);
  p.tiles_h = cdiv(p.Ho, Cfg::TH);
  const long long total = static_cast<long long>(N) * p.tiles_h * p.tiles_w;
  if (total >= (1ll << 24)) return fail(OCTSEG_EINVAL, "mbconv: too many tiles (%lld)", total);
  p.total_tiles = static_cast<int>(total);
  p.fd_tw = make_fastdiv(static_cast<uint32_t>(p.tiles_w));
  p.fd_th = make_fastdiv(static_cast<uint32_t>(p.tiles_h));
  p.fd_iw = make_fastdiv(static_cast<uint32_t>(Cfg::IW));
  const size_t smem = mb_smem_bytes<K, S>(p.kch);
  if (smem > 227 * 1024) return fail(OCTSEG_EINVAL, "mbconv: tile does not fit shared memory (Cin too wide: %d chunks)", p.kch);
  const int sms = octseg_sm_count();
  if (sms <= 0) return sms;
  const int per_sm = smem * 2 <= 227 * 1024 ? Cfg::CTAS : 1;
  const int ctas = sms * per_sm;
  const int grid = p.total_tiles < ctas ? p.total_tiles : ctas;
  static bool attr_set = false;
  if (!attr_set) {
    OCTSEG_CUDA(cudaFuncSetAttribute(mbconv_fused_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  mbconv_fused_kernel<K, S><<<grid, kMbThreads, smem, st>>>(tm_x, tm_w, p);
  return check_launch("mbconv_fused_kernel");
}

}  // namespace octseg

using namespace octseg;

extern "C" int octseg_mbconv_smem_bytes(int32_t Cin, int32_t k, int32_t stride) {
  if (Cin % 16 || Cin < 16 || (k != 3 && k != 5) || stride != 1) return -1;
  return static_cast<int>(k == 3 ? mb_smem_bytes<3, 1>(Cin / 16) : mb_smem_bytes<5, 1>(Cin / 16));
}

extern "C" int octseg_mbconv_pool_slots(int32_t k, int32_t Ho, int32_t Wo) {
  return k == 3 ? cdiv(Ho, MbCfg<3, 1>::TH) * cdiv(Wo, MbCfg<3, 1>::TW) : cdiv(Ho, MbCfg<5, 1>::TH) * cdiv(Wo, MbCfg<5, 1>::TW);
}

extern "C" int octseg_mbconv_blob_floats(int32_t Cmid, int32_t k) { return cdiv(Cmid, kMbCB) * (2 + k * k) * kMbCB; }

extern "C" int octseg_mbconv_expand_dw(const void* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t ldc_in,
                                       const void* w_exp, const float* blob, void* out, int32_t Cmid, int32_t k,
                                       int32_t stride, int32_t pad_t, int32_t pad_l, int32_t Ho, int32_t Wo,
                                       float* pool_sum, void* stream) {
  if (Cin % 16 || Cin < 16) return fail(OCTSEG_EINVAL, "mbconv: Cin must be a multiple of 16 (Cin=%d)", Cin);
  if (Cmid % 8 || ldc_in % 8 || ldc_in < Cin) return fail(OCTSEG_EINVAL, "mbconv: Cmid and the input pitch must be multiples of 8");
  if (stride != 1 || (k != 3 && k != 5)) return fail(OCTSEG_EINVAL, "mbconv: unsupported kernel %d / stride %d", k, stride);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(w_exp) & 15) ||
      (reinterpret_cast<uintptr_t>(blob) & 15))
    return fail(OCTSEG_EINVAL, "mbconv: x/out/w_exp/blob must be 16-byte aligned");
  const int IH = k == 3 ? MbCfg<3, 1>::IH : MbCfg<5, 1>::IH, IW = k == 3 ? MbCfg<3, 1>::IW : MbCfg<5, 1>::IW;
  CUtensorMap tm_x, tm_w;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(ldc_in) * 2, static_cast<uint64_t>(W) * ldc_in * 2,
                                 static_cast<uint64_t>(H) * W * ldc_in * 2};
    const uint32_t box[4] = {16u, static_cast<uint32_t>(IW), static_cast<uint32_t>(IH), 1u};
    const uint32_t estr[4] = {1u, 1u, 1u, 1u};
    const int rc = encode_tensor_map_bf16(&tm_x, x, 4, dims, strides, box, estr, 32, 128, "mbconv x");
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cmid)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Cin) * 2};
    const uint32_t box[2] = {16u, static_cast<uint32_t>(kMbCB)};
    const uint32_t estr[2] = {1u, 1u};
    const int rc = encode_tensor_map_bf16(&tm_w, w_exp, 2, dims, strides, box, estr, 32, 128, "mbconv w_exp");
    if (rc) return rc;
  }
  MbParams p;
  p.blob = blob;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.pool_sum = pool_sum;
  p.H = H;
  p.W = W;
  p.Cmid = Cmid;
  p.Ho = Ho;
  p.Wo = Wo;
  p.pad_t = pad_t;
  p.pad_l = pad_l;
  p.kch = Cin / 16;
  p.ncb = cdiv(Cmid, kMbCB);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k == 3) return launch_mb<3, 1>(tm_x, tm_w, p, N, st);
  return launch_mb<5, 1>(tm_x, tm_w, p, N, st);
}
