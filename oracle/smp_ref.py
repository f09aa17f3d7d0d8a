"""ORACLE (test infrastructure, never on the product path): pure-PyTorch fp32 restatement of the
third-party arithmetic the reference's hot path runs.

The reference calls ``smp.create_model(arch, encoder_name, in_channels, classes, **kw)``
(/root/reference/src/models/smp/model.py:38-44) and then ``self.model(x)`` (model.py:70,192).
``segmentation_models_pytorch==0.3.3`` and its encoder packages (``timm==0.9.2``,
``efficientnet-pytorch==0.7.1``; /root/reference/environment.yaml:32) are NOT vendored in
/root/reference and are not installable here, so this file restates their published
algorithms with state-dict-key compatible module trees (App. C):

  * Unet / UnetPlusPlus / Linknet decoders + SegmentationHead  (smp 0.3.3)
  * resnet{34,50,101} encoder      = torchvision ResNet minus fc (what smp's ResNetEncoder subclasses)
  * timm-regnetx_064 encoder       = timm 0.9.2 RegNet (Bottleneck, group width 56, no SE)
  * efficientnet-b7 encoder        = efficientnet_pytorch 0.7.1 EfficientNet with *static* same padding

PARITY UNPINNED against the reference's own logits: it ships no golden tensors, tests or weights
for this path.  What pins this file instead:
  * numerically, against an independent implementation (tests/test_oracle_vs_torchvision.py):
    the RegNetX-6.4GF encoder equals torchvision's RegNet under a key-rename map, the
    EfficientNet-B7 encoder equals torchvision's efficientnet_b7 on every block (stride-2 layers
    through a pad-and-crop shim that turns torchvision's symmetric padding into
    efficientnet_pytorch's static "same" padding), resnet101 is torchvision's class itself;
  * structurally: parameter counts that reproduce the DVC checkpoint sizes and the widely quoted
    smp totals (tests/test_oracle_models.py), MAC counts (App. D) and output shapes.
The smp decoders/heads have no independent implementation available offline; they are pinned by the
parameter totals and by the wiring spec of SURVEY.md App. A only.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet


# --------------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------------
class ResNetEncoder(ResNet):
    """smp.encoders.resnet.ResNetEncoder: torchvision ResNet, fc/avgpool removed, 6 feature taps."""

    def __init__(self, out_channels, depth=5, **kwargs):
        super().__init__(**kwargs)
        self._depth = depth
        self.out_channels = out_channels
        self._in_channels = 3
        del self.fc
        del self.avgpool

    def get_stages(self):
        return [
            nn.Identity(),
            nn.Sequential(self.conv1, self.bn1, self.relu),
            nn.Sequential(self.maxpool, self.layer1),
            self.layer2,
            self.layer3,
            self.layer4,
        ]

    def forward(self, x):
        features = []
        for stage in self.get_stages()[: self._depth + 1]:
            x = stage(x)
            features.append(x)
        return features


class ConvNormAct(nn.Module):
    """timm ConvNormAct: keys ``conv.weight`` and ``bn.*``."""

    def __init__(self, cin, cout, k, stride=1, groups=1, apply_act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.apply_act = apply_act

    def forward(self, x):
        x = self.bn(self.conv(x))
        return F.relu(x) if self.apply_act else x


class RegNetBottleneck(nn.Module):
    """timm 0.9.2 regnet.Bottleneck with bottle_ratio=1, se_ratio=0: 1x1 -> grouped 3x3(stride)
    -> 1x1 (no act), 1x1 conv shortcut when shape changes, ReLU(x + shortcut)."""

    def __init__(self, cin, cout, stride, group_size):
        super().__init__()
        groups = cout // group_size
        self.conv1 = ConvNormAct(cin, cout, 1)
        self.conv2 = ConvNormAct(cout, cout, 3, stride=stride, groups=groups)
        self.conv3 = ConvNormAct(cout, cout, 1, apply_act=False)
        self.downsample = ConvNormAct(cin, cout, 1, stride=stride, apply_act=False) if (cin != cout or stride != 1) else None

    def forward(self, x):
        shortcut = x
        x = self.conv3(self.conv2(self.conv1(x)))
        if self.downsample is not None:
            shortcut = self.downsample(shortcut)
        return F.relu(x + shortcut)


def regnet_widths(w0=184, wa=60.83, wm=2.07, depth=17, group_size=56, q=8):
    """timm generate_regnet + adjust_widths_groups_comp -> per-stage (widths, depths)."""
    widths_cont = [w0 + wa * i for i in range(depth)]
    exps = [round(math.log(w / w0) / math.log(wm)) for w in widths_cont]
    widths = [int(round(w0 * wm ** e / q) * q) for e in exps]
    stage_w, stage_d = [], []
    for w in widths:
        if stage_w and stage_w[-1] == w:
            stage_d[-1] += 1
        else:
            stage_w.append(w)
            stage_d.append(1)
    adj = []
    for w in stage_w:
        g = min(group_size, w)
        adj.append(int(round(w / g) * g))
    return adj, stage_d


class RegNetXEncoder(nn.Module):
    """smp TimmRegNetEncoder for timm-regnetx_064: stem + s1..s4, head removed."""

    def __init__(self, depth=5):
        super().__init__()
        widths, depths = regnet_widths()
        assert widths == [168, 392, 784, 1624] and depths == [2, 4, 10, 1], (widths, depths)
        self.out_channels = (3, 32, 168, 392, 784, 1624)
        self._depth = depth
        self.stem = ConvNormAct(3, 32, 3, stride=2)
        cin = 32
        for si, (w, d) in enumerate(zip(widths, depths), start=1):
            stage = nn.Sequential()
            for bi in range(d):
                stage.add_module(f'b{bi + 1}', RegNetBottleneck(cin, w, 2 if bi == 0 else 1, 56))
                cin = w
            setattr(self, f's{si}', stage)

    def forward(self, x):
        features = [x]
        x = self.stem(x)
        features.append(x)
        for s in (self.s1, self.s2, self.s3, self.s4):
            x = s(x)
            features.append(x)
        return features


def _same_pad(size, k, s):
    out = math.ceil(size / s)
    p = max((out - 1) * s + (k - 1) + 1 - size, 0)
    return p // 2, p - p // 2


class Conv2dStaticSamePadding(nn.Conv2d):
    """efficientnet_pytorch.utils.Conv2dStaticSamePadding: TF 'SAME' padding computed once from
    the NOMINAL image size (600 for b7), not from the actual input."""

    def __init__(self, cin, cout, k, stride=1, groups=1, bias=False, image_size=None):
        super().__init__(cin, cout, k, stride=stride, groups=groups, bias=bias)
        ih, iw = (image_size, image_size) if isinstance(image_size, int) else image_size
        pt, pb = _same_pad(ih, k, stride)
        pl, pr = _same_pad(iw, k, stride)
        self.static_pad = (pl, pr, pt, pb)
        self.static_padding = nn.ZeroPad2d(self.static_pad) if any(self.static_pad) else nn.Identity()

    def forward(self, x):
        x = self.static_padding(x)
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


class MBConvBlock(nn.Module):
    """efficientnet_pytorch 0.7.1 MBConvBlock (eval mode: drop-connect is the identity)."""

    def __init__(self, k, stride, expand, cin, cout, image_size, se_ratio=0.25, eps=1e-3):
        super().__init__()
        self.stride, self.cin, self.cout, self.expand = stride, cin, cout, expand
        mid = cin * expand
        if expand != 1:
            self._expand_conv = Conv2dStaticSamePadding(cin, mid, 1, image_size=image_size)
            self._bn0 = nn.BatchNorm2d(mid, momentum=0.01, eps=eps)
        self._depthwise_conv = Conv2dStaticSamePadding(mid, mid, k, stride=stride, groups=mid, image_size=image_size)
        self._bn1 = nn.BatchNorm2d(mid, momentum=0.01, eps=eps)
        sq = max(1, int(cin * se_ratio))
        self._se_reduce = Conv2dStaticSamePadding(mid, sq, 1, bias=True, image_size=(1, 1))
        self._se_expand = Conv2dStaticSamePadding(sq, mid, 1, bias=True, image_size=(1, 1))
        self._project_conv = Conv2dStaticSamePadding(mid, cout, 1, image_size=math.ceil(image_size / stride))
        self._bn2 = nn.BatchNorm2d(cout, momentum=0.01, eps=eps)

    def forward(self, inputs, drop_connect_rate=None):
        x = inputs
        if self.expand != 1:
            x = self._bn0(self._expand_conv(x))
            x = x * torch.sigmoid(x)
        x = self._bn1(self._depthwise_conv(x))
        x = x * torch.sigmoid(x)
        # efficientnet_pytorch: F.adaptive_avg_pool2d(x, 1).  Its CUDA backward is not deterministic, so the
        # (test-only) seeded fits of oracle/synth.py take the same mean through a reduction that is
        s = x.mean((2, 3), keepdim=True) if self.training else F.adaptive_avg_pool2d(x, 1)
        s = self._se_reduce(s)
        s = s * torch.sigmoid(s)
        s = self._se_expand(s)
        x = torch.sigmoid(s) * x
        x = self._bn2(self._project_conv(x))
        if self.stride == 1 and self.cin == self.cout:
            x = x + inputs
        return x


# (kernel, stride, expand, in, out, repeats) after width 2.0 / depth 3.1 scaling
EFFNET_B7_STAGES = [
    (3, 1, 1, 64, 32, 4), (3, 2, 6, 32, 48, 7), (5, 2, 6, 48, 80, 7), (3, 2, 6, 80, 160, 10),
    (5, 1, 6, 160, 224, 10), (5, 2, 6, 224, 384, 13), (3, 1, 6, 384, 640, 4),
]


class EfficientNetB7Encoder(nn.Module):
    """smp EfficientNetEncoder('efficientnet-b7'): stage_idxs (11, 18, 38, 55); `_conv_head` and
    `_bn1` stay in the state dict but are never executed."""

    def __init__(self, depth=5):
        super().__init__()
        self.out_channels = (3, 64, 48, 80, 224, 640)
        self._stage_idxs = (11, 18, 38, 55)
        self._depth = depth
        size = 600
        self._conv_stem = Conv2dStaticSamePadding(3, 64, 3, stride=2, image_size=size)
        self._bn0 = nn.BatchNorm2d(64, momentum=0.01, eps=1e-3)
        size = math.ceil(size / 2)
        blocks = []
        for k, s, e, cin, cout, r in EFFNET_B7_STAGES:
            for i in range(r):
                blocks.append(MBConvBlock(k, s if i == 0 else 1, e, cin if i == 0 else cout, cout, size))
                if i == 0:
                    size = math.ceil(size / s)
        self._blocks = nn.ModuleList(blocks)
        self._conv_head = Conv2dStaticSamePadding(640, 2560, 1, image_size=size)
        self._bn1 = nn.BatchNorm2d(2560, momentum=0.01, eps=1e-3)

    def forward(self, x):
        features = [x]
        x = self._bn0(self._conv_stem(x))
        x = x * torch.sigmoid(x)
        features.append(x)
        lo = 0
        for hi in self._stage_idxs:
            for blk in self._blocks[lo:hi]:
                x = blk(x)
            features.append(x)
            lo = hi
        return features


def _resnet(block, layers, out_channels):
    return lambda: ResNetEncoder(out_channels=out_channels, block=block, layers=layers)


ENCODERS = {
    'resnet34': _resnet(BasicBlock, [3, 4, 6, 3], (3, 64, 64, 128, 256, 512)),
    'resnet50': _resnet(Bottleneck, [3, 4, 6, 3], (3, 64, 256, 512, 1024, 2048)),
    'resnet101': _resnet(Bottleneck, [3, 4, 23, 3], (3, 64, 256, 512, 1024, 2048)),
    'timm-regnetx_064': RegNetXEncoder,
    'efficientnet-b7': EfficientNetB7Encoder,
}


def get_encoder(name: str, in_channels: int = 3, depth: int = 5, weights=None) -> nn.Module:
    if name not in ENCODERS:
        raise KeyError(f'Wrong encoder name `{name}`, supported encoders: {list(ENCODERS)}')
    if in_channels != 3 or depth != 5:
        raise ValueError('oracle restates the reference configuration only: in_channels=3, depth=5')
    if weights is not None:
        raise ValueError('pretrained encoder weights are not available offline (reference passes encoder_weights=None)')
    return ENCODERS[name]()


def get_preprocessing_params(encoder_name: str) -> Dict[str, List[float]]:
    """smp.encoders.get_preprocessing_params: ImageNet statistics for all three encoders."""
    if encoder_name not in ENCODERS:
        raise KeyError(encoder_name)
    return {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225], 'input_space': 'RGB', 'input_range': [0, 1]}


# --------------------------------------------------------------------------------------------
# decoders (smp 0.3.3)
# --------------------------------------------------------------------------------------------
class Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout, k, padding=0):
        super().__init__(nn.Conv2d(cin, cout, k, padding=padding, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class UnetDecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout, 3, padding=1)
        self.attention1 = nn.Identity()
        self.conv2 = Conv2dReLU(cout, cout, 3, padding=1)
        self.attention2 = nn.Identity()

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode='nearest')
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList([UnetDecoderBlock(i, s, o) for i, s, o in zip(in_ch, skip_ch, decoder_channels)])

    def forward(self, *features):
        features = features[1:][::-1]
        x = self.center(features[0])
        skips = features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class UnetPlusPlusDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        self.in_channels = [enc[0]] + list(decoder_channels[:-1])
        self.skip_channels = enc[1:] + [0]
        self.out_channels = list(decoder_channels)
        blocks = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(layer_idx + 1):
                if depth_idx == 0:
                    cin = self.in_channels[layer_idx]
                    cskip = self.skip_channels[layer_idx] * (layer_idx + 1)
                    cout = self.out_channels[layer_idx]
                else:
                    cout = self.skip_channels[layer_idx]
                    cskip = self.skip_channels[layer_idx] * (layer_idx + 1 - depth_idx)
                    cin = self.skip_channels[layer_idx - 1]
                blocks[f'x_{depth_idx}_{layer_idx}'] = UnetDecoderBlock(cin, cskip, cout)
        blocks[f'x_0_{len(self.in_channels) - 1}'] = UnetDecoderBlock(self.in_channels[-1], 0, self.out_channels[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = len(self.in_channels) - 1

    def forward(self, *features):
        features = features[1:][::-1]
        dense = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(self.depth - layer_idx):
                if layer_idx == 0:
                    dense[f'x_{depth_idx}_{depth_idx}'] = self.blocks[f'x_{depth_idx}_{depth_idx}'](
                        features[depth_idx], features[depth_idx + 1])
                else:
                    li = depth_idx + layer_idx
                    cat = [dense[f'x_{idx}_{li}'] for idx in range(depth_idx + 1, li + 1)]
                    cat = torch.cat(cat + [features[li + 1]], dim=1)
                    dense[f'x_{depth_idx}_{li}'] = self.blocks[f'x_{depth_idx}_{li}'](dense[f'x_{depth_idx}_{li - 1}'], cat)
        dense[f'x_0_{self.depth}'] = self.blocks[f'x_0_{self.depth}'](dense[f'x_0_{self.depth - 1}'])
        return dense[f'x_0_{self.depth}']


class TransposeX2(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class LinknetDecoderBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.block = nn.Sequential(Conv2dReLU(cin, cin // 4, 1), TransposeX2(cin // 4, cin // 4), Conv2dReLU(cin // 4, cout, 1))

    def forward(self, x, skip=None):
        x = self.block(x)
        if skip is not None:
            x = x + skip
        return x


class LinknetDecoder(nn.Module):
    def __init__(self, encoder_channels, prefinal_channels=32, n_blocks=5):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        ch = enc + [prefinal_channels]
        self.blocks = nn.ModuleList([LinknetDecoderBlock(ch[i], ch[i + 1]) for i in range(n_blocks)])

    def forward(self, *features):
        features = features[1:][::-1]
        x = features[0]
        skips = features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class SegmentationHead(nn.Sequential):
    """conv(k, padding=k//2, bias) -> Identity upsampling -> Identity activation."""

    def __init__(self, cin, cout, kernel_size=3):
        super().__init__(nn.Conv2d(cin, cout, kernel_size, padding=kernel_size // 2), nn.Identity(), nn.Identity())


def initialize_decoder(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_uniform_(m.weight, mode='fan_in', nonlinearity='relu')
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def initialize_head(module):
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class SegmentationModel(nn.Module):
    def check_input_shape(self, x):
        h, w = x.shape[-2:]
        if h % 32 != 0 or w % 32 != 0:
            nh = (h // 32 + 1) * 32 if h % 32 else h
            nw = (w // 32 + 1) * 32 if w % 32 else w
            raise RuntimeError(
                f'Wrong input shape height={h}, width={w}. Expected image height and width divisible by 32. '
                f'Consider pad your images to shape ({nh}, {nw}).')

    def forward(self, x):
        self.check_input_shape(x)
        features = self.encoder(x)
        return self.segmentation_head(self.decoder(*features))


class Unet(SegmentationModel):
    def __init__(self, encoder_name='resnet34', encoder_weights=None, in_channels=3, classes=1, **_):
        super().__init__()
        self.encoder = get_encoder(encoder_name, in_channels, 5, encoder_weights)
        self.decoder = UnetDecoder(self.encoder.out_channels)
        self.segmentation_head = SegmentationHead(16, classes, 3)
        initialize_decoder(self.decoder)
        initialize_head(self.segmentation_head)


class UnetPlusPlus(SegmentationModel):
    def __init__(self, encoder_name='resnet34', encoder_weights=None, in_channels=3, classes=1, **_):
        super().__init__()
        self.encoder = get_encoder(encoder_name, in_channels, 5, encoder_weights)
        self.decoder = UnetPlusPlusDecoder(self.encoder.out_channels)
        self.segmentation_head = SegmentationHead(16, classes, 3)
        initialize_decoder(self.decoder)
        initialize_head(self.segmentation_head)


class Linknet(SegmentationModel):
    def __init__(self, encoder_name='resnet34', encoder_weights=None, in_channels=3, classes=1, **_):
        super().__init__()
        self.encoder = get_encoder(encoder_name, in_channels, 5, encoder_weights)
        self.decoder = LinknetDecoder(self.encoder.out_channels)
        self.segmentation_head = SegmentationHead(32, classes, 1)
        initialize_decoder(self.decoder)
        initialize_head(self.segmentation_head)


ARCHS = {cls.__name__.lower(): cls for cls in (Unet, UnetPlusPlus, Linknet)}


def create_model(arch: str, encoder_name: str = 'resnet34', encoder_weights: Optional[str] = None,
                 in_channels: int = 3, classes: int = 1, **kwargs) -> nn.Module:
    """smp.create_model: case-insensitive arch lookup, KeyError on unknown arch."""
    try:
        cls = ARCHS[arch.lower()]
    except KeyError:
        raise KeyError(f'Wrong architecture type `{arch}`. Available options are: {list(ARCHS)}')
    return cls(encoder_name=encoder_name, encoder_weights=encoder_weights, in_channels=in_channels, classes=classes, **kwargs)
