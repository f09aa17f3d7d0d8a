"""Depthwise-conv micro-benchmark through the C-ABI (timing + ncu target)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200 import _lib

if os.environ.get('OCTSEG_AB_LIB'):     # A/B builds of the library (tools/ab/*.so)
    _lib.LIB_PATH = os.path.abspath(os.environ['OCTSEG_AB_LIB'])

CASES = [(3, 1, 288, 224), (5, 1, 480, 112), (5, 1, 1344, 56), (3, 1, 960, 56), (5, 1, 2304, 28), (3, 1, 32, 448),
         (3, 2, 192, 448), (5, 2, 288, 224), (3, 1, 64, 448), (3, 1, 3840, 28)]


def main():
    lib = _lib.load()
    only = int(sys.argv[1]) if len(sys.argv) > 1 else None
    N = 16
    for ci, (k, s, C, H) in enumerate(CASES):
        if only is not None and ci != only:
            continue
        x = torch.randn(N, H, H, C, device='cuda').to(torch.bfloat16)
        w = (torch.randn(k, k, C, device='cuda') * 0.2).to(torch.bfloat16)
        b = torch.zeros(C, device='cuda')
        Ho = H // s
        out = torch.empty(N, Ho, Ho, C, dtype=torch.bfloat16, device='cuda')
        slots = lib.octseg_dwconv_pool_slots(C, Ho, Ho)
        pool = torch.zeros(N, slots, C, device='cuda')
        pad = max((Ho - 1) * s + k - H, 0) // 2
        st = torch.cuda.current_stream().cuda_stream

        def run():
            _lib.check(lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H, H, C, k, s, pad, pad,
                                         Ho, Ho, 2, pool.data_ptr(), slots, st), 'dw')
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps(dict(k=k, s=s, C=C, H=H, ms=round(ms, 4), TBps=round((x.numel() + out.numel()) * 2 / ms / 1e9, 2))), flush=True)


if __name__ == '__main__':
    main()
