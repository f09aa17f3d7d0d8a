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
