"""Deterministic synthetic OCT-like frames and the three shipped model configurations, for
benchmarks and tests (the reference's dataset and trained weights are not available offline).

Frame statistics follow the real demo frames (/root/reference/data/demo/input: sepia RGB, circular
field of view with ~21.5 % zeros outside it); generator per SURVEY.md §8d:
numpy PCG64(seed = 20251018 + frame index), so any rank can generate its own slice.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

FRAME_SEED = 20251018

# eval/training/*/fold_1/config.json of the reference
MODEL_CONFIGS = {
    'LM': {'model_name': 'UnetPlusPlus_resnet101', 'architecture': 'UnetPlusPlus', 'encoder': 'resnet101',
           'input_size': 512, 'classes': ['Lumen']},
    'FC_LC': {'model_name': 'LinkNet_efficientnet-b7', 'architecture': 'LinkNet', 'encoder': 'efficientnet-b7',
              'input_size': 896, 'classes': ['Fibrous cap', 'Lipid core']},
    'VV': {'model_name': 'Unet_timm-regnetx_064', 'architecture': 'Unet', 'encoder': 'timm-regnetx_064',
           'input_size': 896, 'classes': ['Vasa vasorum']},
}
MODEL_SEEDS = {'LM': 1000, 'FC_LC': 1001, 'VV': 1002}


def synthetic_frame(idx: int, size: int = 512) -> np.ndarray:
    """uint8 RGB (size, size, 3) OCT-like frame, deterministic in ``idx``."""
    rng = np.random.Generator(np.random.PCG64(FRAME_SEED + idx))
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    c = (size - 1) / 2.0
    dx, dy = xx - c, yy - c
    r = np.sqrt(dx * dx + dy * dy) / size          # 0 .. ~0.7
    th = np.arctan2(dy, dx)
    # lumen boundary with low-order angular harmonics
    r_l = rng.uniform(0.15, 0.30)
    bound = r_l * np.ones_like(r)
    for k in range(1, 4):
        bound += rng.uniform(0.0, 0.03) * np.cos(k * th + rng.uniform(0, 2 * np.pi))
    inten = np.zeros_like(r)
    wall = r >= bound
    inten[wall] = (200.0 * np.exp(-(r[wall] - bound[wall]) / rng.uniform(0.05, 0.12)))
    inten += 15.0 * (r < bound)                      # blood-free lumen: faint
    ring = np.abs(r - 0.05) < 0.006                   # catheter ring
    inten[ring] = 230.0
    # dark lipid wedge and bright vessel blobs
    if rng.random() < 0.7:
        a0, aw = rng.uniform(-np.pi, np.pi), rng.uniform(0.3, 1.2)
        wedge = (np.abs(np.angle(np.exp(1j * (th - a0)))) < aw / 2) & (r > bound + 0.03) & (r < bound + 0.18)
        inten[wedge] *= 0.25
    for _ in range(int(rng.integers(0, 4))):
        ba, br = rng.uniform(-np.pi, np.pi), rng.uniform(0.32, 0.45)
        bx, by = c + br * size * np.cos(ba), c + br * size * np.sin(ba)
        blob = (xx - bx) ** 2 + (yy - by) ** 2 < (rng.uniform(0.008, 0.02) * size) ** 2
        inten[blob] = 180.0
    speckle = rng.rayleigh(scale=0.8, size=r.shape).astype(np.float32)
    inten = inten * speckle
    inten[r > 0.5] = 0.0                              # circular field of view
    inten = np.clip(inten, 0, 255)
    rgb = np.stack([inten, 0.45 * inten, 0.08 * inten], axis=-1)
    return rgb.astype(np.uint8)


def synthetic_frames(start: int, count: int, size: int = 512) -> np.ndarray:
    return np.stack([synthetic_frame(start + i, size) for i in range(count)])



def random_models(device, keys=('LM', 'FC_LC', 'VV'), input_size=None, arch=None) -> Dict[str, Tuple[object, Dict]]:
    """Seeded random-init models of the shipped architectures (library-default init, unit BatchNorm
    statistics, damped residual gammas) keyed by model_dir, ready for EnsemblePipeline."""
    from .model import OCTSegmentationModel
    out = {}
    for key in keys:
        cfg = dict(MODEL_CONFIGS[key], **((arch or {}).get(key) or {}))      # arch: per-key architecture / encoder override
        if input_size is not None:
            cfg['input_size'] = input_size
        torch.manual_seed(MODEL_SEEDS[key])
        m = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                 in_channels=3, classes=cfg['classes'], encoder_weights=None)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.fill_(0.5)
        out[key] = (m.to(device).eval(), cfg)
    return out
