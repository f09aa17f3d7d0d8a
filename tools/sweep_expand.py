"""Sweep of (BN, tile shape, chunk width) for one output-bound 1x1 conv through conv_tc_kernel.
Usage: python tools/sweep_expand.py CIN COUT H [batch] [act]"""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200.engine import conv as C

cin, cout, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
N = int(sys.argv[4]) if len(sys.argv) > 4 else 32
act = sys.argv[5] if len(sys.argv) > 5 else 'swish'
x = torch.randn(N, H, H, C.pad8(cin), device='cuda').to(torch.bfloat16)
w = torch.randn(cout, cin, 1, 1) * 0.05
orig = (C.choose_kc, C.choose_bn, C.choose_tile)
bns = sorted({b for b in (64, 128, 192, 256) if b <= max(64, -(-cout // 64) * 64)})
tiles = [(1, 128), (2, 64), (4, 32), (8, 16)]
for kc, bn, tile in itertools.product((16, 32, 64), bns, tiles):
    if kc > 16 and kc // 2 >= cin:
        continue
    if tile[1] > H or H % tile[1]:
        continue
    C.choose_kc = lambda c, kc=kc: kc
    C.choose_bn = lambda cp, k=0, bn=bn: (-(-cp // bn), bn)
    C.choose_tile = lambda h, w_, tile=tile: tile
    try:
        geom, packed = C.plan_conv([((N, H, H, cin, C.pad8(cin)), False)], w)
        out = torch.empty(N, H, H, geom.Cout, dtype=torch.bfloat16, device='cuda')
        plan = C.ConvPlan(geom, packed, C.pad_bias(torch.zeros(cout), geom, cout), [x], out, act=act, name='x')
    except Exception as e:
        print(json.dumps(dict(kc=kc, BN=bn, tile=tile, error=str(e)[:80])), flush=True)
        continue
    finally:
        C.choose_kc, C.choose_bn, C.choose_tile = orig
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
    print(json.dumps(dict(kc=kc, BN=bn, ntn=geom.n_tiles_n, tile=tile, ms=round(ms, 4), GBs=round((x.numel() + out.numel()) * 2 / ms / 1e6))), flush=True)
