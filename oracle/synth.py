"""ORACLE-side synthetic data (test / bench infrastructure): deterministic OCT-like frames and
seeded synthetic checkpoints with calibrated BatchNorm statistics.

The reference's trained weights (models/{LM,FC_LC,VV}.dvc) and dataset are not available
(SURVEY.md §0), so parity and throughput are measured on these.  Frame statistics follow the
real demo frames (/root/reference/data/demo/input: sepia RGB, circular field of view with ~21.5 %
zeros outside it); generator per SURVEY.md §8d.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import model_ref, smp_ref

from oct_segmentation_b200.synthetic import FRAME_SEED, synthetic_frame, synthetic_frames  # noqa: F401  (shared generator)

# the three shipped networks (eval/training/*/fold_1/config.json in the reference)
MODEL_CONFIGS = {
    'LM': {'model_name': 'UnetPlusPlus_resnet101', 'architecture': 'UnetPlusPlus', 'encoder': 'resnet101',
           'input_size': 512, 'classes': ['Lumen']},
    'FC_LC': {'model_name': 'LinkNet_efficientnet-b7', 'architecture': 'LinkNet', 'encoder': 'efficientnet-b7',
              'input_size': 896, 'classes': ['Fibrous cap', 'Lipid core']},
    'VV': {'model_name': 'Unet_timm-regnetx_064', 'architecture': 'Unet', 'encoder': 'timm-regnetx_064',
           'input_size': 896, 'classes': ['Vasa vasorum']},
}
MODEL_SEEDS = {'LM': 1000, 'FC_LC': 1001, 'VV': 1002}
# Parity-only cases beside the shipped trio: BASELINE.json configs[0] (plain U-Net on resnet101, the reference's best
# U-Net for the lumen) and two cross pairings of the shipped decoders / encoders (the factory accepts any pairing).
EXTRA_CONFIGS = {
    'U_LM': {'model_name': 'Unet_resnet101', 'architecture': 'Unet', 'encoder': 'resnet101', 'input_size': 512,
             'classes': ['Lumen']},
    'LINK_R101': {'model_name': 'LinkNet_resnet101', 'architecture': 'LinkNet', 'encoder': 'resnet101', 'input_size': 512,
                  'classes': ['Fibrous cap', 'Lipid core']},
    'UPP_REGNET': {'model_name': 'UnetPlusPlus_timm-regnetx_064', 'architecture': 'UnetPlusPlus',
                   'encoder': 'timm-regnetx_064', 'input_size': 512, 'classes': ['Lumen']},
}
EXTRA_SEEDS = {'U_LM': 1010, 'LINK_R101': 1011, 'UPP_REGNET': 1012}


def model_config(key: str) -> dict:
    return MODEL_CONFIGS[key] if key in MODEL_CONFIGS else EXTRA_CONFIGS[key]


def _randomize_bn(model: nn.Module, gen: torch.Generator) -> None:
    """BN affine: gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1); the LAST BatchNorm of every residual
    branch gets gamma ~ U(0.1, 0.3).  A randomly initialised BN network is chaotic (perturbations
    grow exponentially with depth), which no trained checkpoint is; damped residual branches
    (cf. zero-init-residual, which timm applies to RegNet) give a well-conditioned stand-in that
    still exercises every layer."""
    last = set()
    for name, m in model.named_modules():
        cls = type(m).__name__
        if cls == 'Bottleneck' and hasattr(m, 'bn3'):
            last.add(name + '.bn3')
        elif cls == 'BasicBlock':
            last.add(name + '.bn2')
        elif cls == 'RegNetBottleneck':
            last.add(name + '.conv3.bn')
        elif cls == 'MBConvBlock' and m.stride == 1 and m.cin == m.cout:
            last.add(name + '._bn2')
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                g = torch.rand(m.weight.shape, generator=gen)
                m.weight.copy_(g * 0.2 + 0.1 if name in last else g + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


@torch.no_grad()
def calibrate_bn(model: nn.Module, x: torch.Tensor) -> None:
    """Set BN running statistics from a forward pass so activations stay O(1) through the net."""
    saved = {}
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            saved[name] = m.momentum
            m.reset_running_stats()
            m.momentum = None  # cumulative average
    model.train()
    model(x)
    model.eval()
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            m.momentum = saved[name]


def make_model(key: str, calib_size: int = 128, calib_frames: int = 2, logit_gain: float = 1.0) -> model_ref.OCTSegmentationModelRef:
    """Seeded synthetic checkpoint for 'LM' | 'FC_LC' | 'VV': library-default init, randomised BN
    affine, BN statistics calibrated on synthetic frames (BGR, 0..255, un-normalised like predict())."""
    cfg = model_config(key)
    seed = MODEL_SEEDS[key] if key in MODEL_SEEDS else EXTRA_SEEDS[key]
    torch.manual_seed(seed)
    m = model_ref.OCTSegmentationModelRef(arch=cfg['architecture'], encoder_name=cfg['encoder'],
                                          model_name=cfg['model_name'], in_channels=3, classes=cfg['classes'],
                                          encoder_weights=None)
    gen = torch.Generator().manual_seed(seed + 7)
    _randomize_bn(m.model, gen)
    frames = synthetic_frames(0, calib_frames, calib_size)[..., ::-1].copy()       # RGB -> BGR
    x = torch.from_numpy(frames).permute(0, 3, 1, 2).float()
    calibrate_bn(m.model, x)
    if logit_gain != 1.0:
        with torch.no_grad():
            m.model.segmentation_head[0].weight.mul_(logit_gain)
    m.eval()
    return m


def save_checkpoint(model: nn.Module, path: str) -> None:
    """Same container as a pytorch_lightning .ckpt as far as inference reads it."""
    torch.save({'state_dict': model.state_dict(), 'epoch': 0, 'global_step': 0,
                'pytorch-lightning_version': '2.2.1'}, path)


def phantom_targets(start: int, count: int, size: int, n_classes: int) -> np.ndarray:
    """Structured ground truth for a short fit: class 0 = lumen disk of synthetic frame `idx` (same
    random stream as the frame generator), class 1 = the 0.08*size thick wall ring around it."""
    out = np.zeros((count, n_classes, size, size), np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    c = (size - 1) / 2.0
    dx, dy = xx - c, yy - c
    r = np.sqrt(dx * dx + dy * dy) / size
    th = np.arctan2(dy, dx)
    for i in range(count):
        rng = np.random.Generator(np.random.PCG64(FRAME_SEED + start + i))
        bound = rng.uniform(0.15, 0.30) * np.ones_like(r)
        for k in range(1, 4):
            bound += rng.uniform(0.0, 0.03) * np.cos(k * th + rng.uniform(0, 2 * np.pi))
        out[i, 0] = r < bound
        if n_classes > 1:
            out[i, 1] = (r >= bound) & (r < bound + 0.08)
    return out


@contextlib.contextmanager
def deterministic_torch():
    """Bit-reproducible training on one GPU model + software stack: deterministic cuDNN/cuBLAS algorithms only
    (CUBLAS_WORKSPACE_CONFIG must be set before CUDA starts -- tests/conftest.py does).  The same seeds then give
    the same fitted weights on every B200 lease, so a parity gate built on them does not depend on the lease."""
    prev = (torch.are_deterministic_algorithms_enabled(), torch.is_deterministic_algorithms_warn_only_enabled(),
            torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.use_deterministic_algorithms(True, warn_only=True)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        yield
    finally:
        torch.use_deterministic_algorithms(prev[0], warn_only=prev[1])
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = prev[2], prev[3]


def fit_model(model: 'model_ref.OCTSegmentationModelRef', device, steps: int = 300, size: int = 128, batch: int = 8,
              lr: float = 2e-3, seed: int = 0, pool: int = 32, log=None) -> float:
    """Short seeded, DETERMINISTIC fit of the WHOLE oracle network on phantom targets (Adam, BCE-with-logits, cosine
    learning-rate decay, a fixed number of steps) at the resolution the checkpoint will be used at, so that masks are
    structured and |logit| is large away from object boundaries -- the regime a trained checkpoint is in, and the one
    in which Dice between two implementations is meaningful (SURVEY.md S7 'hard parts').  Frames `1000 .. 1000+pool`
    of the synthetic generator are cycled.  Returns the smoothed final loss.  Leaves the model in eval mode."""
    torch.manual_seed(seed)
    net = model.model.to(device)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    n_classes = len(model.classes)
    pool = max(pool, batch)
    frames = torch.from_numpy(synthetic_frames(1000, pool, size)[..., ::-1].copy())
    targets = torch.from_numpy(phantom_targets(1000, pool, size, n_classes))
    ema = None
    with deterministic_torch():
        for step in range(steps):
            for g in opt.param_groups:
                g['lr'] = lr * (0.05 + 0.95 * 0.5 * (1.0 + math.cos(math.pi * step / steps)))
            idx = [(step * batch + j) % pool for j in range(batch)]
            x = frames[idx].to(device).permute(0, 3, 1, 2).float()
            t = targets[idx].to(device)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(net(x), t)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            v = loss.item()
            ema = v if ema is None else 0.9 * ema + 0.1 * v
            if log is not None and (step % 25 == 0 or step == steps - 1):
                log(f'step {step}: loss {v:.4f} (smoothed {ema:.4f})')
    net.eval()
    model.eval()
    return float(ema)
