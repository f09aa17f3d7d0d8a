"""A/B: mid-size 1x1 convs with per-tile weight loads (default) vs weights resident in shared memory
(halo-tile mode with a 1x1 window; conv.RESIDENT_1X1).  Checks both against torch and times them.
Usage: python tools/ab/res1x1_test.py [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from oct_segmentation_b200.engine import conv as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SHAPES = [(160, 960, 56), (224, 1344, 56), (384, 2304, 28), (640, 3840, 28), (168, 168, 224), (392, 392, 112),
          (960, 160, 56), (1344, 224, 56), (2304, 384, 28), (256, 256, 128), (512, 512, 64), (1024, 1024, 32)]


def run(cin, cout, H, resident):
    C.RESIDENT_1X1 = resident
    g = torch.Generator().manual_seed(cin + cout)
    w = torch.randn(cout, cin, 1, 1, generator=g) * (1.0 / cin ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    geom, packed = C.plan_conv([((N, H, H, cin, C.pad8(cin)), False)], w)
    x = torch.randn(N, H, H, C.pad8(cin), generator=g).to(torch.bfloat16).cuda()
    out = torch.full((N, H, H, geom.Cout), float('nan'), dtype=torch.bfloat16, device='cuda')
    plan = C.ConvPlan(geom, packed, C.pad_bias(b, geom, cout), [x], out, act='swish', name='x')
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    want = F.silu(F.conv2d(x[:2, ..., :cin].permute(0, 3, 1, 2).float(), w.to(torch.bfloat16).float().cuda(), b.cuda()))
    got = out[:2, ..., :cout].permute(0, 3, 1, 2).float()
    err = ((got - want).norm() / want.norm()).item()
    tail = ((out[-1, ..., :cout].float() - F.silu(F.conv2d(x[-1:, ..., :cin].permute(0, 3, 1, 2).float(),
                                                            w.to(torch.bfloat16).float().cuda(), b.cuda()))[0].permute(1, 2, 0)).norm()
            / want.norm()).item()
    return dict(cin=cin, cout=cout, H=H, resident=resident, halo=geom.halo, BN=geom.BN, ntn=geom.n_tiles_n, tile=(geom.TH, geom.TW),
                ms=round(ms, 4), TFs=round(2 * N * H * H * cin * cout / ms / 1e9, 1), rel=round(err, 5), rel_last=round(tail, 5),
                finite=bool(torch.isfinite(out.float()).all()))


for cin, cout, H in SHAPES:
    for r in (False, True):
        print(json.dumps(run(cin, cout, H, r)), flush=True)
