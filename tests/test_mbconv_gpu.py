"""GPU: fused MBConv front half (octseg_mbconv_expand_dw, csrc/mbconv.cu: expand 1x1 + swish -> depthwise k x k +
swish -> squeeze-excite sums in one kernel) vs torch fp32 ops on the same bf16 inputs, with the expanded tensor
rounded to bf16 where the unfused path stores it.  Covers the (Cin, Cmid, k) combinations of efficientnet-b7's
stride-1 blocks that fit the kernel, expanded widths that are not a multiple of the 64-channel block, maps that are
not a multiple of the 8 x 16 tile (and smaller than one tile), several images, and the SE sums."""
import pytest
import torch
import torch.nn.functional as F

from oct_segmentation_b200 import _lib

pytestmark = pytest.mark.gpu

# (Cin, Cmid, k, H, W, N)
CASES = [
    (32, 192, 3, 40, 56, 2), (48, 288, 3, 32, 32, 1), (48, 288, 3, 56, 56, 3), (80, 480, 5, 28, 36, 2),
    (160, 960, 3, 14, 18, 1), (160, 960, 5, 28, 28, 2), (48, 288, 5, 17, 23, 1), (16, 64, 3, 9, 9, 1),
    (80, 480, 3, 5, 7, 2), (32, 200, 5, 33, 16, 1),
]


def swish(t):
    return t * torch.sigmoid(t)


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'cin%d_cmid%d_k%d_%dx%d_N%d' % c)
def test_mbconv_expand_dw_matches_torch(case):
    cin, cmid, k, H, W, N = case
    lib = _lib.load()
    assert 0 < lib.octseg_mbconv_smem_bytes(cin, k, 1) <= 227 * 1024
    g = torch.Generator().manual_seed(cin * 7 + cmid + k + H)
    ldc = cin + 8                                                     # a padded channel pitch, like pad8() tensors
    xs = torch.zeros(N, H, W, ldc, dtype=torch.bfloat16)
    xs[..., :cin] = torch.randn(N, H, W, cin, generator=g).to(torch.bfloat16)
    xs[..., cin:] = 7.0                                               # must never be read
    x = xs.cuda()
    we = (torch.randn(cmid, cin, generator=g) / cin ** 0.5).to(torch.bfloat16).cuda()
    be = (torch.randn(cmid, generator=g) * 0.5).cuda()
    wd = (torch.randn(k, k, cmid, generator=g) * 0.3).to(torch.bfloat16).cuda()
    bd = (torch.randn(cmid, generator=g) * 0.5).cuda()
    p = (k - 1) // 2
    out = torch.full((N, H, W, cmid), float('nan'), dtype=torch.bfloat16, device='cuda')
    slots = lib.octseg_mbconv_pool_slots(k, H, W)
    pool3 = torch.full((N, slots, cmid), float('nan'), device='cuda')   # every slot must be written (no zeroing contract)
    blob = _lib.mbconv_blob(be, wd, bd, k).cuda()
    assert blob.numel() == lib.octseg_mbconv_blob_floats(cmid, k)
    _lib.check(lib.octseg_mbconv_expand_dw(x.data_ptr(), N, H, W, cin, ldc, we.data_ptr(), blob.data_ptr(), out.data_ptr(),
                                           cmid, k, 1, p, p, H, W, pool3.data_ptr(), torch.cuda.current_stream().cuda_stream), 'mbconv')
    torch.cuda.synchronize()
    pool = pool3.sum(1)
    xin = x[..., :cin].float().permute(0, 3, 1, 2)
    e = swish(F.conv2d(xin, we.float()[:, :, None, None], be)).to(torch.bfloat16).float()      # stored as bf16 when unfused
    ref = swish(F.conv2d(e, wd.float().permute(2, 0, 1).unsqueeze(1), bd, padding=p, groups=cmid))
    got = out.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel <= 6e-3, rel            # bf16 rounding of the expanded tensor may differ by an ulp (tanh.approx) + output rounding
    assert (got - ref).abs().max().item() <= 3e-2 * max(ref.abs().max().item(), 1.0)
    want_pool = ref.sum(dim=(2, 3))
    assert torch.allclose(pool, want_pool, rtol=3e-3, atol=2e-3 * H * W)


def test_mbconv_rejects_unsupported_shapes():
    lib = _lib.load()
    assert lib.octseg_mbconv_smem_bytes(24, 3, 1) < 0          # Cin not a multiple of 16
    assert lib.octseg_mbconv_smem_bytes(48, 3, 2) < 0          # stride 2 stays on the unfused path
    assert lib.octseg_mbconv_smem_bytes(640, 3, 1) > 227 * 1024
    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device='cuda')
    rc = lib.octseg_mbconv_expand_dw(x.data_ptr(), 1, 8, 8, 24, 24, x.data_ptr(), x.data_ptr(), x.data_ptr(), 64, 3, 1, 1, 1,
                                     8, 8, None, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b'Cin' in lib.octseg_last_error()
