"""Times the fused MBConv kernel (octseg_mbconv_expand_dw) against the two launches it replaces (1x1 expand conv +
depthwise conv) on efficientnet-b7's stride-1 block shapes.  Usage: python tools/bench_mbconv.py [batch] [case ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200 import _lib
from oct_segmentation_b200.engine import conv as C

# name: (Cin, Cmid, k, H)
CASES = {'s2': (48, 288, 3, 224), 's3': (80, 480, 5, 112), 's4': (160, 960, 3, 56), 's5': (224, 1344, 5, 56)}


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    names = [a for a in sys.argv[2:] if not a.startswith('--')] or list(CASES)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for nm in names:
        cin, cmid, k, H = CASES[nm]
        p = (k - 1) // 2
        x = torch.randn(N, H, H, cin, device='cuda').to(torch.bfloat16)
        we = (torch.randn(cmid, cin) / cin ** 0.5)
        wek = we.to(torch.bfloat16).cuda()
        be = torch.randn(cmid, device='cuda') * 0.1
        wd = (torch.randn(k, k, cmid) * 0.3).to(torch.bfloat16).cuda()
        bd = torch.randn(cmid, device='cuda') * 0.1
        out = torch.empty(N, H, H, cmid, dtype=torch.bfloat16, device='cuda')
        pool = torch.zeros(N, max(lib.octseg_mbconv_pool_slots(k, H, H), lib.octseg_dwconv_pool_slots(cmid, H, H)), cmid, device='cuda')
        row = dict(case=nm, batch=N, cin=cin, cmid=cmid, k=k, H=H, out_MB=round(out.numel() * 2 / 1e6, 1))
        if lib.octseg_mbconv_smem_bytes(cin, k, 1) <= 227 * 1024:
            blob = _lib.mbconv_blob(be, wd, bd, k).cuda()

            def fused():
                _lib.check(lib.octseg_mbconv_expand_dw(x.data_ptr(), N, H, H, cin, cin, wek.data_ptr(), blob.data_ptr(), out.data_ptr(),
                                                       cmid, k, 1, p, p, H, H, pool.data_ptr(), st), 'mb')
            row['fused_ms'] = round(timed(fused), 4)
            row['fused_out_GBs'] = round(out.numel() * 2 / row['fused_ms'] / 1e6)
        if '--fused-only' not in sys.argv:
            e = torch.empty(N, H, H, cmid, dtype=torch.bfloat16, device='cuda')
            geom, packed = C.plan_conv([((N, H, H, cin, cin), False)], we[:, :, None, None])
            plan = C.ConvPlan(geom, packed, C.pad_bias(be.cpu(), geom, cmid), [x], e, act='swish', name='expand')

            def unfused():
                plan.run()
                _lib.check(lib.octseg_dwconv(e.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), N, H, H, cmid, k, 1, p, p,
                                             H, H, _lib.ACT['swish'], pool.data_ptr(), lib.octseg_dwconv_pool_slots(cmid, H, H), st), 'dw')
            row['unfused_ms'] = round(timed(unfused), 4)
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
