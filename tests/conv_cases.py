"""Shared conv test cases: torch fp32 reference ops + a CPU emulator of the kernel's addressing.

Test infrastructure only (never imported by the product path).  The emulator walks the packed
weights and K-segments exactly as conv_tc.cu's producer does (segment -> tap row -> tap column
-> 64-channel chunk; zero fill outside the tensor), so the host planner can be validated
without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from oct_segmentation_b200.engine.conv import ConvGeom, pad8


@dataclass
class ConvCase:
    name: str
    srcs: List[Tuple[int, int, int, bool]]     # (C, H, W, upsampled)
    cout: int
    k: int = 3
    stride: int = 1
    pad: Tuple[int, int] = (1, 1)
    pad_br: Optional[Tuple[int, int]] = None    # bottom/right pad when asymmetric
    groups: int = 1
    transposed: bool = False
    act: str = 'relu'
    res_mode: str = 'none'
    out_mode: str = 'bf16_nhwc'
    N: int = 2
    bias: bool = True


def out_hw(case: ConvCase) -> Tuple[int, int]:
    C, H, W, up = case.srcs[0]
    if case.transposed:
        return 2 * H, 2 * W
    if any(s[3] for s in case.srcs):
        s = next(s for s in case.srcs if s[3])
        return 2 * s[1], 2 * s[2]
    pb, pr = case.pad_br if case.pad_br is not None else case.pad
    return ((H + case.pad[0] + pb - case.k) // case.stride + 1, (W + case.pad[1] + pr - case.k) // case.stride + 1)


def make_inputs(case: ConvCase, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(case.N, C, H, W, generator=g).to(torch.bfloat16).float() for C, H, W, _ in case.srcs]
    cin = sum(s[0] for s in case.srcs)
    if case.transposed:
        w = torch.randn(cin, case.cout, 4, 4, generator=g) * (1.0 / (cin * 4) ** 0.5)
    else:
        w = torch.randn(case.cout, cin // case.groups, case.k, case.k, generator=g) * (
            1.0 / (cin // case.groups * case.k * case.k) ** 0.5)
    b = torch.randn(case.cout, generator=g) * 0.1 if case.bias else None
    Ho, Wo = out_hw(case)
    res = torch.randn(case.N, case.cout, Ho, Wo, generator=g).to(torch.bfloat16).float() if case.res_mode != 'none' else None
    return xs, w, b, res


def act_fn(y, act):
    if act == 'relu':
        return F.relu(y)
    if act == 'swish':
        return y * torch.sigmoid(y)
    if act == 'sigmoid':
        return torch.sigmoid(y)
    return y


def reference(case: ConvCase, xs, w, b, res) -> torch.Tensor:
    """What torch computes for the un-fused op sequence (NCHW fp32)."""
    parts = [F.interpolate(x, scale_factor=2, mode='nearest') if s[3] else x for x, s in zip(xs, case.srcs)]
    x = torch.cat(parts, dim=1)
    if case.transposed:
        y = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    else:
        pb, pr = case.pad_br if case.pad_br is not None else case.pad
        x = F.pad(x, (case.pad[1], pr, case.pad[0], pb))
        y = F.conv2d(x, w, b, stride=case.stride, groups=case.groups)
    if case.res_mode == 'before_act':
        y = y + res
    y = act_fn(y, case.act)
    if case.res_mode == 'after_act':
        y = y + res
    if case.out_mode == 'u8_nchw':
        y = (y > 0).to(torch.uint8)
    return y


def emulate(geom: ConvGeom, packed: torch.Tensor, bias_rows: torch.Tensor, xs_nhwc: List[torch.Tensor],
            act: str, res_nhwc: Optional[torch.Tensor], res_mode: str) -> torch.Tensor:
    """CPU walk of the kernel's K loop.  xs_nhwc: float tensors [N,H,W,C]; returns [N,out_H,out_W,Cout]."""
    W = packed.float()
    N, Hq, Wq = geom.N, geom.Hq, geom.Wq
    rows = geom.n_tiles_n * geom.BN
    out = torch.zeros(N, geom.out_H, geom.out_W, geom.Cout)
    for phase in range(geom.phases):
        ph, pw = phase >> 1, phase & 1
        acc = torch.zeros(N, Hq, Wq, rows)
        kofs = 0
        for sg, x in zip(geom.segs, xs_nhwc):
            assert x.shape == (N, sg.H, sg.W, sg.C), (x.shape, sg)
            for ty in range(sg.kh):
                hh = sg.mul * torch.arange(Hq) + sg.off_h[ph] + ty
                for tx in range(sg.kw):
                    ww = sg.mul * torch.arange(Wq) + sg.off_w[pw] + tx
                    hv, wv = (hh >= 0) & (hh < sg.H), (ww >= 0) & (ww < sg.W)
                    g = x[:, hh.clamp(0, sg.H - 1)][:, :, ww.clamp(0, sg.W - 1)]
                    g = g * (hv[:, None] & wv[None, :])[None, :, :, None]
                    for cc in range(sg.cchunks):
                        for t in range(geom.n_tiles_n):
                            c0 = sg.c_per_tile * t + cc * sg.kc
                            a = torch.zeros(N, Hq, Wq, sg.kc)
                            hi = min(c0 + sg.kc, sg.C)
                            if hi > c0:
                                a[..., :hi - c0] = g[..., c0:hi]
                            wt = W[phase, t * geom.BN:(t + 1) * geom.BN, kofs:kofs + sg.kc]
                            acc[..., t * geom.BN:(t + 1) * geom.BN] += a @ wt.t()
                        kofs += sg.kc
        assert kofs == geom.Ktot
        acc = acc + bias_rows[:rows]
        for t in range(geom.n_tiles_n):
            ch0 = t * geom.cout_per_tile
            nvalid = min(geom.cout_per_tile, geom.Cout - ch0)
            if nvalid <= 0:
                continue
            y = acc[..., t * geom.BN:t * geom.BN + nvalid]
            if geom.phases == 4:
                view = out[:, ph::2, pw::2, ch0:ch0 + nvalid]
                r = res_nhwc[:, ph::2, pw::2, ch0:ch0 + nvalid] if res_nhwc is not None else None
            else:
                view = out[:, :, :, ch0:ch0 + nvalid]
                r = res_nhwc[..., ch0:ch0 + nvalid] if res_nhwc is not None else None
            if res_mode == 'before_act':
                y = y + r
            y = act_fn(y, act)
            if res_mode == 'after_act':
                y = y + r
            view.copy_(y)
    return out


CASES = [
    ConvCase('c1x1_64_256', [(64, 16, 16, False)], 256, k=1, pad=(0, 0)),
    ConvCase('c3x3_64_64', [(64, 16, 24, False)], 64),
    ConvCase('c3x3_s2_128', [(128, 18, 18, False)], 128, stride=2),
    ConvCase('c1x1_s2_ds', [(256, 16, 16, False)], 512, k=1, stride=2, pad=(0, 0), act='none'),
    ConvCase('bottleneck_res', [(64, 8, 8, False)], 256, k=1, pad=(0, 0), res_mode='before_act'),
    ConvCase('odd_28', [(80, 28, 28, False)], 48, k=1, pad=(0, 0), act='swish'),
    ConvCase('odd_cin168', [(168, 14, 14, False)], 168, k=1, pad=(0, 0)),
    ConvCase('cout_392', [(168, 14, 14, False)], 392, k=1, pad=(0, 0)),
    ConvCase('grouped_56', [(168, 14, 14, False)], 168, groups=3),
    ConvCase('grouped_56_s2', [(392, 16, 16, False)], 392, groups=7, stride=2),
    ConvCase('concat2', [(64, 16, 16, False), (32, 16, 16, False)], 64),
    ConvCase('up_only', [(32, 8, 8, True)], 16),
    ConvCase('up_concat', [(128, 8, 8, True), (64, 16, 16, False)], 64),
    ConvCase('up_concat3', [(64, 8, 8, True), (64, 16, 16, False), (32, 16, 16, False), (16, 16, 16, False)], 32),
    ConvCase('convT', [(40, 8, 8, False)], 40, k=4, transposed=True),
    ConvCase('convT_small', [(16, 12, 12, False)], 16, k=4, transposed=True),
    ConvCase('linknet_skip', [(40, 16, 16, False)], 80, k=1, pad=(0, 0), res_mode='after_act'),
    ConvCase('head_f32', [(16, 32, 32, False)], 1, out_mode='f32_nchw', act='none'),
    ConvCase('head_u8', [(16, 32, 32, False)], 2, k=1, pad=(0, 0), out_mode='u8_nchw', act='none'),
    ConvCase('cin_small_cout24', [(24, 16, 16, False)], 20, k=1, pad=(0, 0)),
    ConvCase('big_k', [(512, 8, 8, True), (256, 16, 16, False), (256, 16, 16, False)], 256),
    ConvCase('wide_2048', [(1024, 8, 8, False)], 2048, k=1, pad=(0, 0), N=1),
    ConvCase('kc16_3x3', [(16, 32, 32, False)], 16),
    ConvCase('kc32_3x3', [(32, 24, 24, False)], 32),
    ConvCase('kc16x3_1x1_48', [(48, 28, 28, False)], 288, k=1, pad=(0, 0), act='swish'),
    ConvCase('kc_mixed_up', [(32, 16, 16, True), (16, 32, 32, False), (64, 32, 32, False)], 32),
    ConvCase('kc32_s2', [(24, 32, 32, False)], 40, stride=2),
    ConvCase('partial_chunk_168', [(32, 16, 16, False)], 168, k=1, pad=(0, 0)),
    ConvCase('wide_3x3_64', [(64, 6, 128, False)], 64),
    ConvCase('wide_3x3_128_w160', [(128, 5, 160, False)], 128),
    ConvCase('wide_up2x2', [(64, 4, 128, True), (64, 8, 256, False)], 64),
    ConvCase('wide_grouped', [(128, 6, 128, False)], 128, groups=2),
    # halo-tile mode (16x8 tiles, one halo box per chunk, resident weights)
    ConvCase('halo_3x3_64_partial_tiles', [(64, 40, 44, False)], 64),
    ConvCase('halo_grouped_56_28', [(392, 28, 28, False)], 392, groups=7),
    ConvCase('halo_head_u8', [(64, 32, 32, False)], 2, out_mode='u8_nchw', act='none'),
    ConvCase('halo_head_f32', [(64, 32, 48, False)], 1, out_mode='f32_nchw', act='none'),
    ConvCase('halo_5x5_cout16', [(64, 32, 32, False)], 16, k=5, pad=(2, 2)),
    ConvCase('halo_res_before', [(64, 32, 32, False)], 64, res_mode='before_act'),
    ConvCase('halo_two_chunks', [(128, 32, 32, False)], 32, act='swish'),
    ConvCase('halo_asym_pad', [(64, 32, 32, False)], 48, pad=(0, 0), pad_br=(2, 2), act='none'),
    ConvCase('halo_kc16_4x4_stem_like', [(12, 32, 40, False)], 64, k=4, pad=(2, 2), pad_br=(1, 1)),
    ConvCase('halo_kc16_2x2_stem_like', [(12, 48, 32, False)], 32, k=2, pad=(0, 0), pad_br=(1, 1), act='swish'),
    ConvCase('halo_kc32_two_chunks', [(40, 32, 32, False)], 64),
]
