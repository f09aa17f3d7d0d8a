"""Pre/post-processing kernels in isolation (CUDA events around graph replays, 4 rotating buffer sets > L2): achieved GB/s of
ALGORITHMIC bytes against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200 import prepost


def timeit(fns, reps=5):
    """The calls are captured into ONE CUDA graph (the Python wrappers cost more host time than the
    kernels run), then the graph replay is timed with CUDA events."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns))


def main():
    N, SRC, HO = 32, 512, 1000
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    sets = 4
    frames = [torch.randint(0, 255, (N, SRC, SRC, 3), dtype=torch.uint8, device='cuda') for _ in range(sets)]
    for S in (512, 896):
        outs = [torch.empty(N, S, S, 3, dtype=torch.uint8, device='cuda') for _ in range(sets)]
        ms = timeit([lambda i=i: prepost.preprocess(frames[i], S, outs[i]) for i in range(sets)])
        b = N * (3 * SRC * SRC + 3 * S * S)
        print(json.dumps(dict(kernel='preprocess_resize_bgr', S_src=SRC, S=S, frames=N, ms=round(ms, 4),
                              GBps=round(b / ms / 1e6, 1), frac_of_hbm=round(b / ms / 1e6 / peak, 3))), flush=True)
    planes = [{0: (torch.rand(N, 512, 512, device='cuda') > 0.5).to(torch.uint8),
               1: (torch.rand(N, 896, 896, device='cuda') > 0.5).to(torch.uint8),
               2: (torch.rand(N, 896, 896, device='cuda') > 0.5).to(torch.uint8),
               3: (torch.rand(N, 896, 896, device='cuda') > 0.5).to(torch.uint8)} for _ in range(sets)]
    masks = [torch.empty(N, HO, HO, 4, dtype=torch.uint8, device='cuda') for _ in range(sets)]
    labels = [torch.empty(N, HO, HO, dtype=torch.uint8, device='cuda') for _ in range(sets)]
    counts = [torch.zeros(N, 4, dtype=torch.int32, device='cuda') for _ in range(sets)]
    ms = timeit([lambda i=i: prepost.postprocess(planes[i], [0, 1, 2, 3], HO, HO, N, 'cuda', masks[i], labels[i], counts[i])
                 for i in range(sets)])
    b = N * (512 * 512 + 3 * 896 * 896 + 5 * HO * HO)
    print(json.dumps(dict(kernel='postprocess (+counts.zero_)', Ho=HO, frames=N, ms=round(ms, 4), GBps=round(b / ms / 1e6, 1),
                          frac_of_hbm=round(b / ms / 1e6 / peak, 3))), flush=True)
    ms = timeit([lambda i=i: prepost.radial_thickness(masks[i]) for i in range(sets)])
    print(json.dumps(dict(kernel='radial_thickness', frames=N, ms=round(ms, 4))), flush=True)
    big = [torch.randint(0, 255, (N, HO, HO, 3), dtype=torch.uint8, device='cuda') for _ in range(sets)]
    overs = [torch.empty_like(b_) for b_ in big]
    ms = timeit([lambda i=i: prepost.overlay(big[i], masks[i], [0, 1, 2, 3], overs[i]) for i in range(sets)])
    b = N * HO * HO * (3 + 4 + 3)
    print(json.dumps(dict(kernel='overlay', Ho=HO, frames=N, ms=round(ms, 4), GBps=round(b / ms / 1e6, 1),
                          frac_of_hbm=round(b / ms / 1e6 / peak, 3))), flush=True)


if __name__ == '__main__':
    main()
