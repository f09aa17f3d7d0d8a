"""Output-bound 1x1 convs of the MBConv blocks through conv_tc_kernel: activation and chunk-width A/B.
Usage: python tools/bench_expand.py [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200.engine import conv as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SHAPES = [(48, 288, 224), (80, 480, 112), (32, 192, 448), (160, 960, 56), (224, 1344, 56), (288, 48, 224), (64, 256, 128)]


def run(cin, cout, H, act, kc_override=None):
    orig = C.choose_kc
    if kc_override:
        C.choose_kc = lambda c, taps=9: kc_override
    try:
        w = torch.randn(cout, cin, 1, 1) * 0.05
        geom, packed = C.plan_conv([((N, H, H, cin, C.pad8(cin)), False)], w)
    finally:
        C.choose_kc = orig
    x = torch.randn(N, H, H, C.pad8(cin), device='cuda').to(torch.bfloat16)
    out = torch.empty(N, H, H, geom.Cout, dtype=torch.bfloat16, device='cuda')
    plan = C.ConvPlan(geom, packed, C.pad_bias(torch.zeros(cout), geom, cout), [x], out, act=act, name='x')
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
    byt = (x.numel() + out.numel()) * 2
    return dict(cin=cin, cout=cout, H=H, act=act, kc=geom.segs[0].kc, BN=geom.BN, ntn=geom.n_tiles_n, tile=(geom.TH, geom.TW),
                ms=round(ms, 4), GBs=round(byt / ms / 1e6))


for cin, cout, H in SHAPES:
    for act in ('none', 'relu', 'swish'):
        print(json.dumps(run(cin, cout, H, act)), flush=True)
    if cin == 48:
        print(json.dumps(run(cin, cout, H, 'swish', 64)), flush=True)
        print(json.dumps(run(cin, cout, H, 'relu', 64)), flush=True)
