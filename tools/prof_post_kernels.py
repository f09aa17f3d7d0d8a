"""Target program for ncu captures of the post-processing kernels added after the first profile pass
(overlay_kernel, contour_largest_kernel, fold_average_threshold_kernel): each runs 3 times on OCT-shaped masks at
1000 x 1000, batch 32.  Usage (B200_PROFILING.md recipe):
  ncu --set full --clock-control none --import-source on -k regex:overlay_kernel -s 1 -c 1 -o gpurun_out/prof_overlay_r1 \
      python tools/prof_post_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oct_segmentation_b200 import prepost


def shaped_masks(N, HO):
    yy, xx = torch.meshgrid(torch.arange(HO, device='cuda'), torch.arange(HO, device='cuda'), indexing='ij')
    rr = ((yy - HO / 2) ** 2 + (xx - HO / 2) ** 2).float().sqrt()
    ang = torch.atan2((yy - HO / 2).float(), (xx - HO / 2).float())
    m = torch.zeros(N, HO, HO, 4, dtype=torch.uint8, device='cuda')
    m[..., 0] = (rr < 0.22 * HO).to(torch.uint8)
    m[..., 1] = ((rr >= 0.22 * HO) & (rr < 0.26 * HO) & (ang.abs() < 1.0)).to(torch.uint8)
    m[..., 2] = ((rr >= 0.26 * HO) & (rr < 0.36 * HO) & (ang.abs() < 0.9)).to(torch.uint8)
    m[..., 3] = (((yy - 0.2 * HO) ** 2 + (xx - 0.3 * HO) ** 2 < 64) | ((yy - 0.75 * HO) ** 2 + (xx - 0.7 * HO) ** 2 < 100)).to(torch.uint8)
    return m


def main():
    N, HO, K = 32, 1000, 5
    m = shaped_masks(N, HO)
    img = torch.randint(0, 255, (N, HO, HO, 3), dtype=torch.uint8, device='cuda')
    out = torch.empty_like(img)
    logits = [torch.randn(N, 1, 896, 896, device='cuda') for _ in range(K)]
    planes = torch.empty(N, 1, 896, 896, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        prepost.overlay(img, m, [0, 1, 2, 3], out)
        prepost.contour_largest(m)
        prepost.fold_average_threshold(logits, planes)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
