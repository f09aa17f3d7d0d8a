"""Parameter schemas of the three networks on the hot path, state-dict-key compatible with
``segmentation_models_pytorch==0.3.3`` so the reference's ``weights.ckpt``
files load unchanged (/root/reference/src/predict.py:39-48).

These modules only HOLD parameters (and give the attribute surface the reference's callers use:
``.encoder``, ``.decoder``, ``.segmentation_head``, ``encoder.layer4[-1]``).  They contain no torch
compute: ``SegmentationModel.forward`` hands the tensor to the octseg engine, which runs the
hand-written sm_100a kernels behind include/octseg.h.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn as nn
from torchvision.models.resnet import Bottleneck, ResNet


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f'{type(self).__name__} only holds parameters; run the enclosing model '
                           '(its forward executes on the octseg CUDA engine)')


def _conv(cin, cout, k, stride=1, groups=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=0, groups=groups, bias=bias)


# --------------------------------------------------------------------------------- encoders
class ResNet101Encoder(ResNet):
    """torchvision ResNet-101 parameter tree minus fc (what smp's ResNetEncoder subclasses)."""
    out_channels = (3, 64, 256, 512, 1024, 2048)
    kind = 'resnet'

    def __init__(self):
        super().__init__(block=Bottleneck, layers=[3, 4, 23, 3])
        del self.fc

    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError('encoder parameters are executed by the octseg engine; call the model')


class _ConvBN(_ParamsOnly):
    def __init__(self, cin, cout, k, stride=1, groups=1):
        super().__init__()
        self.conv = _conv(cin, cout, k, stride, groups)
        self.bn = nn.BatchNorm2d(cout)


class _RegNetBlock(_ParamsOnly):
    def __init__(self, cin, cout, stride, group_width):
        super().__init__()
        self.stride = stride
        self.conv1 = _ConvBN(cin, cout, 1)
        self.conv2 = _ConvBN(cout, cout, 3, stride, cout // group_width)
        self.conv3 = _ConvBN(cout, cout, 1)
        if cin != cout or stride != 1:
            self.downsample = _ConvBN(cin, cout, 1, stride)
        else:
            self.downsample = None


class RegNetX064Encoder(_ParamsOnly):
    """timm 0.9.2 regnetx_064 (w0=184, wa=60.83, wm=2.07, group width 56, depth 17)."""
    out_channels = (3, 32, 168, 392, 784, 1624)
    kind = 'regnet'
    widths, depths, group_width = (168, 392, 784, 1624), (2, 4, 10, 1), 56

    def __init__(self):
        super().__init__()
        self.stem = _ConvBN(3, 32, 3, 2)
        cin = 32
        for si, (w, d) in enumerate(zip(self.widths, self.depths), start=1):
            stage = nn.Sequential()
            for bi in range(d):
                stage.add_module(f'b{bi + 1}', _RegNetBlock(cin, w, 2 if bi == 0 else 1, self.group_width))
                cin = w
            setattr(self, f's{si}', stage)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
        for m in self.modules():
            if isinstance(m, _RegNetBlock):
                nn.init.zeros_(m.conv3.bn.weight)


def static_same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    """efficientnet_pytorch static 'same' padding (lo, hi) for a NOMINAL extent ``size``."""
    out = math.ceil(size / s)
    p = max((out - 1) * s + (k - 1) + 1 - size, 0)
    return p // 2, p - p // 2


class _StaticPadConv(nn.Conv2d):
    """Conv2dStaticSamePadding parameter holder; ``pad`` = (top/left, bottom/right)."""

    def __init__(self, cin, cout, k, stride=1, groups=1, bias=False, image_size=1):
        super().__init__(cin, cout, k, stride=stride, groups=groups, bias=bias)
        self.pad = static_same_pad(image_size, k, stride)

    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError('parameter holder; executed by the octseg engine')


class _MBConv(_ParamsOnly):
    def __init__(self, k, stride, expand, cin, cout, image_size):
        super().__init__()
        self.k, self.stride, self.expand, self.cin, self.cout = k, stride, expand, cin, cout
        mid = cin * expand
        if expand != 1:
            self._expand_conv = _StaticPadConv(cin, mid, 1, image_size=image_size)
            self._bn0 = nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3)
        self._depthwise_conv = _StaticPadConv(mid, mid, k, stride, groups=mid, image_size=image_size)
        self._bn1 = nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3)
        sq = max(1, int(cin * 0.25))
        self._se_reduce = _StaticPadConv(mid, sq, 1, bias=True)
        self._se_expand = _StaticPadConv(sq, mid, 1, bias=True)
        self._project_conv = _StaticPadConv(mid, cout, 1)
        self._bn2 = nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3)


class EfficientNetB7Encoder(_ParamsOnly):
    """efficientnet_pytorch 0.7.1 efficientnet-b7 (width 2.0, depth 3.1, nominal image 600)."""
    out_channels = (3, 64, 48, 80, 224, 640)
    kind = 'efficientnet'
    stage_idxs = (11, 18, 38, 55)
    # kernel, stride, expand, in, out, repeats
    stages = ((3, 1, 1, 64, 32, 4), (3, 2, 6, 32, 48, 7), (5, 2, 6, 48, 80, 7), (3, 2, 6, 80, 160, 10),
              (5, 1, 6, 160, 224, 10), (5, 2, 6, 224, 384, 13), (3, 1, 6, 384, 640, 4))

    def __init__(self):
        super().__init__()
        size = 600
        self._conv_stem = _StaticPadConv(3, 64, 3, 2, image_size=size)
        self._bn0 = nn.BatchNorm2d(64, momentum=0.01, eps=1e-3)
        size = math.ceil(size / 2)
        blocks: List[nn.Module] = []
        for k, s, e, cin, cout, reps in self.stages:
            for i in range(reps):
                blocks.append(_MBConv(k, s if i == 0 else 1, e, cin if i == 0 else cout, cout, size))
                if i == 0:
                    size = math.ceil(size / s)
        self._blocks = nn.ModuleList(blocks)
        # present in the checkpoint, never executed by smp's encoder forward
        self._conv_head = _StaticPadConv(640, 2560, 1)
        self._bn1 = nn.BatchNorm2d(2560, momentum=0.01, eps=1e-3)


ENCODERS = {
    'resnet101': ResNet101Encoder,
    'timm-regnetx_064': RegNetX064Encoder,
    'efficientnet-b7': EfficientNetB7Encoder,
}


# --------------------------------------------------------------------------------- decoders
class _Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout, k):
        super().__init__(nn.Conv2d(cin, cout, k, padding=k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _UnetBlock(_ParamsOnly):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout, 3)
        self.attention1 = nn.Identity()
        self.conv2 = _Conv2dReLU(cout, cout, 3)
        self.attention2 = nn.Identity()


class UnetDecoderParams(_ParamsOnly):
    kind = 'unet'

    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList([_UnetBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels)])


class UnetPlusPlusDecoderParams(_ParamsOnly):
    kind = 'unetplusplus'

    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        outs = list(decoder_channels)
        self.depth = len(ins) - 1
        blocks = {}
        for l in range(self.depth):
            for d in range(l + 1):
                if d == 0:
                    spec = (ins[l], skips[l] * (l + 1), outs[l])
                else:
                    spec = (skips[l - 1], skips[l] * (l + 1 - d), skips[l])
                blocks[f'x_{d}_{l}'] = _UnetBlock(*spec)
        blocks[f'x_0_{self.depth}'] = _UnetBlock(ins[-1], 0, outs[-1])
        self.blocks = nn.ModuleDict(blocks)


class _LinknetBlock(_ParamsOnly):
    def __init__(self, cin, cout):
        super().__init__()
        mid = cin // 4
        self.block = nn.Sequential(
            _Conv2dReLU(cin, mid, 1),
            nn.Sequential(nn.ConvTranspose2d(mid, mid, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(mid),
                          nn.ReLU(inplace=True)),
            _Conv2dReLU(mid, cout, 1))


class LinknetDecoderParams(_ParamsOnly):
    kind = 'linknet'

    def __init__(self, encoder_channels, prefinal_channels=32, n_blocks=5):
        super().__init__()
        ch = list(encoder_channels[1:])[::-1] + [prefinal_channels]
        self.blocks = nn.ModuleList([_LinknetBlock(ch[i], ch[i + 1]) for i in range(n_blocks)])


class SegmentationHeadParams(nn.Sequential):
    def __init__(self, cin, classes, k):
        super().__init__(nn.Conv2d(cin, classes, k, padding=k // 2), nn.Identity(), nn.Identity())


def init_decoder(module: nn.Module) -> None:
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_uniform_(m.weight, mode='fan_in', nonlinearity='relu')
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def init_head(module: nn.Module) -> None:
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
