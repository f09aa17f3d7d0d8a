"""CPU: the host planner's packed weights + K-segments reproduce torch's un-fused op sequence
when walked exactly as the kernel walks them (tests/conv_cases.emulate)."""
import pytest
import torch

from oct_segmentation_b200.engine import conv as C
from tests.conv_cases import CASES, emulate, make_inputs, out_hw, reference


def to_nhwc_padded(x):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, C.pad8(c))
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out


def plan_case(case, xs, w):
    srcs = [((case.N, s[1], s[2], s[0], C.pad8(s[0])), s[3]) for s in case.srcs]
    return C.plan_conv(srcs, w, out_hw=None if (case.transposed or any(s[3] for s in case.srcs)) else out_hw(case),
                       stride=case.stride, pad=case.pad, groups=case.groups, transposed=case.transposed,
                       out_bf16=case.out_mode == 'bf16_nhwc')


@pytest.mark.parametrize('case', CASES, ids=lambda c: c.name)
def test_emulated_kernel_matches_torch(case):
    xs, w, b, res = make_inputs(case)
    geom, packed = plan_case(case, xs, w)
    assert geom.TH * geom.TW <= 128 and geom.BN % 16 == 0 and 16 <= geom.BN <= 256
    assert packed.shape == (geom.phases, geom.n_tiles_n * geom.BN, geom.Ktot)
    bias_rows = C.pad_bias(b, geom, case.cout, case.groups)
    xs_nhwc = [x.permute(0, 2, 3, 1).contiguous() for x in xs]
    res_nhwc = to_nhwc_padded(res) if res is not None else None
    act = case.act
    got = emulate(geom, packed, bias_rows, xs_nhwc, act, res_nhwc, case.res_mode)
    want = reference(case, xs, w, b, res)
    if case.out_mode == 'u8_nchw':
        want = reference(case.__class__(**{**case.__dict__, 'out_mode': 'f32_nchw'}), xs, w, b, res)
    got_nchw = got[..., :case.cout].permute(0, 3, 1, 2)
    err = (got_nchw - want).norm() / want.norm().clamp_min(1e-6)
    assert err < 6e-3, f'{case.name}: rel-L2 {err:.3e}'   # only the weights' bf16 rounding differs
    if geom.Cout > case.cout:                              # pad channels stay zero
        assert got[..., case.cout:].abs().max() == 0


def test_tile_and_bn_choice():
    assert C.choose_tile(128, 128) == (1, 128)
    th, tw = C.choose_tile(28, 28)
    assert th * tw <= 128 and 28 % tw == 0
    for cout in (16, 64, 168, 392, 784, 1624, 2048, 24):
        n, bn = C.choose_bn(cout)
        assert bn % 16 == 0 and bn <= 256 and n * bn >= cout


@pytest.mark.parametrize('cs,cout,k,W', [([16], 16, 3, 32), ([16], 1, 3, 32), ([32], 32, 3, 24), ([16, 32], 24, 3, 16),
                                          ([16], 32, 1, 32), ([32], 2, 1, 16), ([24], 20, 3, 16), ([8], 16, 3, 32)])
def test_pixel_packed_conv_is_the_same_conv(cs, cout, k, W):
    """Packing f pixels per GEMM row (engine.conv.pack_conv_weights) is an exact re-indexing."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    cps = [C.pad8(c) for c in cs]
    xs = [torch.randn(2, c, 12, W, generator=g) for c in cs]
    w = torch.randn(cout, sum(cs), k, k, generator=g)
    b = torch.randn(cout, generator=g)
    want = F.conv2d(torch.cat(xs, 1), w, b, padding=k // 2)
    cout_store = C.pad8(cout)
    f = C.pixel_pack_factor(cps, W, k, cout_store)
    assert f >= 2
    wp, bp = C.pack_conv_weights(w, b, cs, cps, f, k // 2, cout_store)
    packed_in = []
    for x, c, cp in zip(xs, cs, cps):
        t = torch.zeros(2, 12, W, cp)
        t[..., :c] = x.permute(0, 2, 3, 1)
        packed_in.append(t.view(2, 12, W // f, f * cp).permute(0, 3, 1, 2))
    y = F.conv2d(torch.cat(packed_in, 1), wp, bp, padding=(k // 2, 1 if k > 1 else 0))       # [2, f*cout_store, 12, W/f]
    y = y.permute(0, 2, 3, 1).reshape(2, 12, W, cout_store)
    assert torch.allclose(y[..., :cout].permute(0, 3, 1, 2), want, atol=1e-4)
    assert y[..., cout:].abs().max() == 0 if cout_store > cout else True


@pytest.mark.parametrize('transposed,cin,cout', [(False, 32, 16), (True, 16, 16), (True, 20, 12), (False, 24, 40)])
def test_depth_to_space_form_is_the_same_op(transposed, cin, cout):
    """engine.conv.d2s_weights: upsample+3x3 conv / ConvTranspose(k4,s2,p1) == one 3x3 conv on the
    low-res grid with (ph, pw, co)-ordered channels followed by a pixel shuffle."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, cin, 6, 10, generator=g)
    b = torch.randn(cout, generator=g)
    if transposed:
        w = torch.randn(cin, cout, 4, 4, generator=g)
        want = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    else:
        w = torch.randn(cout, cin, 3, 3, generator=g)
        want = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, b, padding=1)
    cs = C.pad8(cout)
    wp, bp = C.d2s_weights(w, b, transposed, cs)
    y = F.conv2d(x, wp, bp, padding=1)                                  # [2, 4*cs, 6, 10]
    y = y.view(2, 2, 2, cs, 6, 10).permute(0, 3, 4, 1, 5, 2).reshape(2, cs, 12, 20)
    assert torch.allclose(y[:, :cout], want, atol=1e-4)
    if cs > cout:
        assert y[:, cout:].abs().max() == 0
