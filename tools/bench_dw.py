"""Depthwise-conv micro-benchmark through the C-ABI (timing + ncu target)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200 import _lib

CASES = [(3, 1, 288, 224), (5, 1, 480, 112), (5, 1, 1344, 56), (3, 1, 960, 56), (5, 1, 2304, 28), (3, 1, 32, 448)]


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
        out = torch.empty_like(x)
        pool = torch.zeros(N, C, device='cuda')
        pad = (k - 1) // 2
        st = torch.cuda.current_stream().cuda_stream

        def run():
            _lib.check(lib.octseg_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H, H, C, k, s, pad, pad,
                                         H, H, 2, pool.data_ptr(), st), 'dw')
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
        print(json.dumps(dict(k=k, s=s, C=C, H=H, ms=round(ms, 4), TBps=round(2 * x.numel() * 2 / ms / 1e9, 2))), flush=True)


if __name__ == '__main__':
    main()
