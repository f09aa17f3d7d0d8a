"""GPU parity of the tcgen05 implicit-GEMM conv (through the C-ABI) against torch fp32 ops."""
import pytest
import torch

from oct_segmentation_b200.engine import conv as C
from tests.conv_cases import CASES, make_inputs, out_hw, reference
from tests.test_conv_plan import plan_case

pytestmark = pytest.mark.gpu


def dev_nhwc(x):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, C.pad8(c), dtype=torch.bfloat16, device='cuda')
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def run_case(case, seed=0):
    xs, w, b, res = make_inputs(case, seed)
    geom, packed = plan_case(case, xs, w)
    bias_rows = C.pad_bias(b, geom, case.cout, case.groups)
    seg_t = [dev_nhwc(x) for x in xs]
    Ho, Wo = out_hw(case)
    if case.out_mode == 'bf16_nhwc':
        out = torch.full((case.N, Ho, Wo, geom.Cout), float('nan'), dtype=torch.bfloat16, device='cuda')
    elif case.out_mode == 'f32_nchw':
        out = torch.full((case.N, case.cout, Ho, Wo), float('nan'), dtype=torch.float32, device='cuda')
    else:
        out = torch.full((case.N, case.cout, Ho, Wo), 7, dtype=torch.uint8, device='cuda')
    res_t = dev_nhwc(res) if res is not None else None
    plan = C.ConvPlan(geom, packed, bias_rows, seg_t, out, out_mode=case.out_mode, act=case.act, res=res_t,
                      res_mode=case.res_mode, name=case.name)
    plan.run()
    torch.cuda.synchronize()
    want = reference(case, [x.cuda() for x in xs], w.cuda(), b.cuda() if b is not None else None,
                     res.cuda() if res is not None else None)
    return out, want, geom


@pytest.mark.parametrize('case', CASES, ids=lambda c: c.name)
def test_conv_tc_matches_torch(case):
    out, want, geom = run_case(case)
    if case.out_mode == 'bf16_nhwc':
        got = out[..., :case.cout].permute(0, 3, 1, 2).float()
        assert torch.isfinite(out.float()).all(), 'unwritten (NaN) output elements'
        if geom.Cout > case.cout:
            assert out[..., case.cout:].float().abs().max() == 0
        err = ((got - want).norm() / want.norm()).item()
        assert err < 1e-2, f'{case.name}: rel-L2 {err:.3e}'        # bf16 weights + bf16 output rounding
    elif case.out_mode == 'f32_nchw':
        assert torch.isfinite(out).all()
        err = ((out - want).norm() / want.norm()).item()
        assert err < 1e-2, f'{case.name}: rel-L2 {err:.3e}'
    else:
        case_f = case.__class__(**{**case.__dict__, 'out_mode': 'f32_nchw'})
        logits = reference(case_f, *[t if t is None else (
            [x.cuda() for x in t] if isinstance(t, list) else t.cuda()) for t in make_inputs(case)])
        confident = logits.abs() > 0.05
        assert (out[confident] == want[confident]).all()
        assert out.max() <= 1


def test_conv_tc_repeatable():
    """Same plan run twice gives bit-identical output (no stale TMEM / barrier state)."""
    case = CASES[12]
    a, _, _ = run_case(case, seed=3)
    b, _, _ = run_case(case, seed=3)
    assert torch.equal(a, b)


# (cin, cmid, classes, H, W): 16-channel sources pack 4 pixels per GEMM row, 32-channel ones 2, wider ones none
HEAD_CASES = [(16, 32, 2, 40, 48), (16, 32, 1, 33, 36), (32, 48, 3, 24, 40), (64, 64, 4, 20, 24), (24, 16, 2, 16, 64)]


@pytest.mark.parametrize('case', HEAD_CASES, ids=lambda c: 'cin%d_cmid%d_cls%d_%dx%d' % c)
@pytest.mark.parametrize('out_mode', ['f32_nchw', 'u8_nchw'])
@pytest.mark.parametrize('act', ['relu', 'swish'])
def test_conv_with_fused_1x1_head(case, out_mode, act):
    """conv 1x1 + bias + act with a 1x1 segmentation head folded into the epilogue (octseg.h `head_classes`; LinkNet's
    last decoder conv + head) against torch fp32: the logits see the UN-rounded fp32 activations, so they are closer to
    torch than the two-launch path; the u8 planes must equal logit > 0 away from the zero crossing."""
    import torch.nn.functional as F
    from oct_segmentation_b200.engine.builder import Builder
    cin, cmid, classes, H, W = case
    N = 2
    g = torch.Generator().manual_seed(cin * 100 + cmid + classes)
    x = torch.randn(N, cin, H, W, generator=g).to(torch.bfloat16).float()
    w = torch.randn(cmid, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cmid, generator=g) * 0.2
    hw = torch.randn(classes, cmid, 1, 1, generator=g) / cmid ** 0.5
    hb = torch.randn(classes, generator=g) * 0.1
    bld = Builder(torch.device('cuda'), N, reuse=False)
    xa = C.Act(dev_nhwc(x), cin)                   # external input buffer (not an arena activation)
    out = torch.full((N, classes, H, W), 7, dtype=torch.float32 if out_mode == 'f32_nchw' else torch.uint8, device='cuda')
    assert bld.conv([(xa, False)], w, b, name='fused', act=act, out_mode=out_mode, out_tensor=out, head=(hw, hb)) is None
    bld.run()
    torch.cuda.synchronize()
    y = F.conv2d(x.cuda(), w.to(torch.bfloat16).float().cuda(), b.cuda())
    y = F.relu(y) if act == 'relu' else y * torch.sigmoid(y)
    want = F.conv2d(y, hw.cuda(), hb.cuda())
    if out_mode == 'f32_nchw':
        assert torch.isfinite(out).all()
        err = ((out - want).norm() / want.norm()).item()
        assert err < 3e-3, err                   # bf16 conv weights only (tanh.approx swish: ~1e-3)
    else:
        assert out.max() <= 1
        conf = want.abs() > 0.02
        assert (out[conf] == (want[conf] > 0).to(torch.uint8)).all() and conf.float().mean() > 0.9
