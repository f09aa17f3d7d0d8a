"""GPU: depthwise conv kernel (octseg_dwconv, csrc/dwconv.cu) vs torch fp32 on the same bf16 inputs.
Covers both filter sizes / strides of efficientnet-b7, channel counts that are not a multiple of the
CTA's channel block, maps that are not a multiple of the tile, the asymmetric static "same" padding
of efficientnet_pytorch (pad_lo = total // 2, rest at the bottom/right) and the squeeze-excite sums."""
import pytest
import torch
import torch.nn.functional as F

from oct_segmentation_b200 import _lib

pytestmark = pytest.mark.gpu

# (k, stride, C, H, W, N)
CASES = [
    (3, 1, 32, 40, 48, 2), (3, 1, 64, 33, 29, 2), (3, 1, 288, 56, 56, 2), (3, 2, 192, 64, 64, 2),
    (3, 2, 480, 28, 28, 3), (5, 1, 480, 28, 28, 2), (5, 1, 1344, 14, 14, 2), (5, 2, 288, 56, 56, 2),
    (5, 2, 1344, 28, 28, 1), (5, 1, 2304, 28, 28, 1), (3, 1, 3840, 7, 7, 2), (5, 2, 40, 37, 45, 2),
    (3, 1, 8, 16, 16, 1),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'k%d_s%d_C%d_%dx%d_N%d' % c)
@pytest.mark.parametrize('act', ['swish', 'none'])
def test_dwconv_matches_torch(case, act):
    k, s, C, H, W, N = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(k * 1000 + C + H)
    x = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(k, k, C, generator=g) * 0.3).to(torch.bfloat16).cuda()
    b = (torch.randn(C, generator=g) * 0.5).cuda()
    Ho, Wo = -(-H // s), -(-W // s)
    ph, pw = max((Ho - 1) * s + k - H, 0), max((Wo - 1) * s + k - W, 0)
    pt, pl = ph // 2, pw // 2
    out = torch.full((N, Ho, Wo, C), float('nan'), dtype=torch.bfloat16, device='cuda')
    slots = lib.octseg_dwconv_pool_slots(C, Ho, Wo)
    pool3 = torch.full((N, slots, C), float('nan'), device='cuda')       # every slot must be written (no zeroing contract)
    _lib.check(lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H, W, C, k, s, pt, pl,
                                 Ho, Wo, _lib.ACT[act], pool3.data_ptr(), slots, torch.cuda.current_stream().cuda_stream), 'dw')
    pool = pool3.sum(1)
    torch.cuda.synchronize()
    xin = F.pad(x.float().permute(0, 3, 1, 2), (pl, pw - pl, pt, ph - pt))
    ref = F.conv2d(xin, w.float().permute(2, 0, 1).unsqueeze(1), b, stride=s, groups=C)
    if act == 'swish':
        ref = ref * torch.sigmoid(ref)
    got = out.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel <= 4e-3, rel                      # bf16 output rounding (2^-9) + tanh.approx
    assert (got - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)
    want_pool = ref.sum(dim=(2, 3))              # the kernel pools its fp32 activation (before the bf16 rounding)
    assert torch.allclose(pool, want_pool, rtol=2e-3, atol=1e-3 * Ho * Wo)  # tanh.approx: ~5e-4 abs per element


def test_dwconv_without_pool_and_repeatable():
    lib = _lib.load()
    x = torch.randn(2, 24, 24, 96).to(torch.bfloat16).cuda()
    w = (torch.randn(3, 3, 96) * 0.3).to(torch.bfloat16).cuda()
    b = torch.zeros(96).cuda()
    outs = []
    for _ in range(2):
        out = torch.empty(2, 24, 24, 96, dtype=torch.bfloat16, device='cuda')
        _lib.check(lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 2, 24, 24, 96, 3, 1, 1, 1,
                                     24, 24, _lib.ACT['relu'], None, 0, torch.cuda.current_stream().cuda_stream), 'dw')
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(2, 0, 1).unsqueeze(1), b, padding=1, groups=96))
    assert ((outs[0].float().permute(0, 3, 1, 2) - ref).norm() / ref.norm()).item() <= 4e-3


def test_dwconv_pool_sums_are_bit_reproducible():
    """The squeeze-excite sums leave the kernel as write-once slots added in a fixed order (no fp32 atomics): two runs
    on a map with many row groups give identical bits, and a slot count other than the kernel's own is refused."""
    lib = _lib.load()
    N, H, C = 3, 112, 480
    x = torch.randn(N, H, H, C).to(torch.bfloat16).cuda()
    w = (torch.randn(5, 5, C) * 0.2).to(torch.bfloat16).cuda()
    b = torch.randn(C).cuda()
    out = torch.empty(N, H, H, C, dtype=torch.bfloat16, device='cuda')
    slots = lib.octseg_dwconv_pool_slots(C, H, H)
    assert slots == 14
    pools = []
    for _ in range(2):
        pool = torch.full((N, slots, C), float('nan'), device='cuda')
        _lib.check(lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H, H, C, 5, 1, 2, 2, H, H,
                                     _lib.ACT['swish'], pool.data_ptr(), slots, torch.cuda.current_stream().cuda_stream), 'dw')
        pools.append(pool)
    torch.cuda.synchronize()
    assert torch.isfinite(pools[0]).all() and torch.equal(pools[0], pools[1])
    rc = lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H, H, C, 5, 1, 2, 2, H, H,
                           _lib.ACT['swish'], pools[0].data_ptr(), slots + 1, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b'pool_slots' in lib.octseg_last_error()
