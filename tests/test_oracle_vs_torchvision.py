"""CPU: NUMERIC pin of the network oracle (oracle/smp_ref.py) against an independent implementation.

segmentation_models_pytorch / timm / efficientnet_pytorch are not installable here,
but torchvision ships structurally identical encoders written by other people:

  * ``torchvision.models.regnet.RegNet`` with timm's regnetx_064 block parameters
    (depth 17, w_0 184, w_a 60.83, w_m 2.07, group width 56) is the same arithmetic as timm's
    RegNet under a key-rename map -> the five feature taps must be EQUAL in fp32;
  * ``torchvision.models.efficientnet_b7`` has the same channels, repeats, SE widths and BN eps
    as efficientnet_pytorch's b7; only the padding of the stride-2 convs differs (symmetric there,
    TensorFlow-style static "same" padding computed from the nominal 600-pixel image in
    efficientnet_pytorch).  A stride-2 conv with (lo, hi) padding equals
    torchvision's symmetric-p conv applied to the input extended by ``a`` zero rows/columns at the
    top/left (p + a - lo even) and read from output index (p + a - lo)/2 on -- so the torchvision
    modules themselves run every stride-2 layer too, behind a pad-and-crop shim;
  * torchvision ``resnet101`` IS the class smp's ResNetEncoder subclasses -> equality by
    construction, checked for the key set / taps.

Tolerance (written here): RegNet and ResNet taps equal to 1e-6 relative (same ops, possibly a
different oneDNN primitive for an explicit pad); EfficientNet taps 1e-5 relative.
"""
import pytest
import torch
import torch.nn.functional as F
from torchvision.models import efficientnet_b7, resnet101
from torchvision.models.regnet import BlockParams, RegNet

from oracle import smp_ref


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _randomize(model, seed):
    """Non-trivial BN statistics and affine terms, so a swapped or dropped tensor cannot cancel."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.5)
            elif isinstance(m, torch.nn.Conv2d) and m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def _copy_cna(dst_conv, dst_bn, src_conv, src_bn):
    dst_conv.load_state_dict(src_conv.state_dict())
    dst_bn.load_state_dict(src_bn.state_dict())


# ----------------------------------------------------------------------------------- RegNetX-6.4GF
def test_regnetx_064_equals_torchvision_regnet():
    torch.manual_seed(0)
    ours = smp_ref.RegNetXEncoder().eval()
    _randomize(ours, 1)
    tv = RegNet(BlockParams.from_init_params(depth=17, w_0=184, w_a=60.83, w_m=2.07, group_width=56)).eval()
    # key-rename map timm 0.9.2 -> torchvision
    ren = {}
    for k in ours.state_dict():
        p = k.split('.')
        if p[0] == 'stem':
            nk = 'stem.' + ('0' if p[1] == 'conv' else '1') + '.' + '.'.join(p[2:])
        else:
            si, bj = int(p[0][1:]), int(p[1][1:])
            sub = {'conv1': 'f.a', 'conv2': 'f.b', 'conv3': 'f.c', 'downsample': 'proj'}[p[2]]
            nk = f'trunk_output.block{si}.block{si}-{bj - 1}.{sub}.' + ('0' if p[3] == 'conv' else '1') + '.' + '.'.join(p[4:])
        ren[k] = nk
    tv_sd = tv.state_dict()
    enc_keys = {k for k in tv_sd if not k.startswith('fc.')}
    assert set(ren.values()) == enc_keys, 'key map does not cover torchvision RegNet'
    tv.load_state_dict({**{ren[k]: v for k, v in ours.state_dict().items()}, 'fc.weight': tv_sd['fc.weight'],
                        'fc.bias': tv_sd['fc.bias']}, strict=True)
    x = torch.randn(2, 3, 96, 64)
    with torch.no_grad():
        feats = ours(x)
        y = tv.stem(x)
        taps = [y]
        for blk in tv.trunk_output:
            y = blk(y)
            taps.append(y)
    assert [f.shape[1] for f in feats] == [3, 32, 168, 392, 784, 1624]
    for i, (a, b) in enumerate(zip(feats[1:], taps), start=1):
        assert a.shape == b.shape
        assert rel(a, b) <= 1e-6, f'tap f{i}: {rel(a, b):.2e}'


# ----------------------------------------------------------------------------------- ResNet-101
def test_resnet101_encoder_is_torchvision_resnet():
    torch.manual_seed(0)
    ours = smp_ref.get_encoder('resnet101').eval()
    _randomize(ours, 2)
    tv = resnet101(weights=None).eval()
    sd = tv.state_dict()
    sd.update(ours.state_dict())
    tv.load_state_dict(sd, strict=True)
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        feats = ours(x)
        y = tv.relu(tv.bn1(tv.conv1(x)))
        taps = [y]
        y = tv.maxpool(y)
        for layer in (tv.layer1, tv.layer2, tv.layer3, tv.layer4):
            y = layer(y)
            taps.append(y)
    for a, b in zip(feats[1:], taps):
        assert rel(a, b) <= 1e-6


# ----------------------------------------------------------------------------------- EfficientNet-B7
def _shifted_stride2(module, x, k, lo, out_hw):
    """torchvision's symmetric-padding stride-2 conv block run so that it computes (lo, hi) padding."""
    p = (k - 1) // 2
    a = (lo - p) % 2
    shift = (p + a - lo) // 2
    y = module(F.pad(x, (a, k, a, k)))
    return y[:, :, shift:shift + out_hw[0], shift:shift + out_hw[1]]


def test_efficientnet_b7_equals_torchvision_with_static_same_padding():
    torch.manual_seed(0)
    ours = smp_ref.EfficientNetB7Encoder().eval()
    _randomize(ours, 3)
    tv = efficientnet_b7(weights=None).eval()
    _copy_cna(tv.features[0][0], tv.features[0][1], ours._conv_stem, ours._bn0)
    tv_blocks = [blk for stage in list(tv.features)[1:8] for blk in stage]
    assert len(tv_blocks) == len(ours._blocks) == 55
    for ob, tb in zip(ours._blocks, tv_blocks):
        mods = list(tb.block)
        if ob.expand != 1:
            _copy_cna(mods[0][0], mods[0][1], ob._expand_conv, ob._bn0)
            mods = mods[1:]
        assert len(mods) == 3
        _copy_cna(mods[0][0], mods[0][1], ob._depthwise_conv, ob._bn1)
        mods[1].fc1.load_state_dict(ob._se_reduce.state_dict())
        mods[1].fc2.load_state_dict(ob._se_expand.state_dict())
        _copy_cna(mods[2][0], mods[2][1], ob._project_conv, ob._bn2)
        assert mods[0][0].kernel_size == ob._depthwise_conv.kernel_size and mods[0][0].stride == ob._depthwise_conv.stride
        assert mods[1].fc1.out_channels == ob._se_reduce.out_channels
        assert tb.use_res_connect == (ob.stride == 1 and ob.cin == ob.cout)

    x = torch.randn(1, 3, 96, 128) * 50
    with torch.no_grad():
        feats = ours(x)
        pl, pr, pt, pb = ours._conv_stem.static_pad
        assert (pl, pt) == (0, 0)
        y = _shifted_stride2(tv.features[0], x, 3, 0, feats[1].shape[2:])
        assert rel(y, feats[1]) <= 1e-5, f'stem {rel(y, feats[1]):.2e}'
        taps, n_shim = [y], 0
        for i, (ob, tb) in enumerate(zip(ours._blocks, tv_blocks)):
            if ob.stride == 1:
                y = tb(y)                                           # the torchvision block, untouched
            else:
                mods = list(tb.block)
                inp = y
                if ob.expand != 1:
                    inp = mods[0](inp)
                    mods = mods[1:]
                lo = ob._depthwise_conv.static_pad[0]
                assert ob._depthwise_conv.static_pad[0] == ob._depthwise_conv.static_pad[2]
                k = ob._depthwise_conv.kernel_size[0]
                ho = (inp.shape[2] + sum(ob._depthwise_conv.static_pad[2:]) - k) // 2 + 1
                wo = (inp.shape[3] + sum(ob._depthwise_conv.static_pad[:2]) - k) // 2 + 1
                z = _shifted_stride2(mods[0], inp, k, lo, (ho, wo))
                y = mods[2](mods[1](z))
                n_shim += 1
            if (i + 1) in ours._stage_idxs:
                taps.append(y)
        assert n_shim == 4
    assert [f.shape[1] for f in feats] == [3, 64, 48, 80, 224, 640]
    for i, (a, b) in enumerate(zip(feats[1:], taps), start=1):
        assert a.shape == b.shape, (i, a.shape, b.shape)
        assert rel(a, b) <= 1e-5, f'tap f{i}: {rel(a, b):.2e}'


def test_efficientnet_symmetric_padding_would_differ():
    """The shim matters: torchvision's own (symmetric) stride-2 padding gives different features, so the test
    above really pins efficientnet_pytorch's static same padding and not just the channel plan."""
    torch.manual_seed(0)
    ours = smp_ref.EfficientNetB7Encoder().eval()
    tv = efficientnet_b7(weights=None).eval()
    _copy_cna(tv.features[0][0], tv.features[0][1], ours._conv_stem, ours._bn0)
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        a = ours._bn0(ours._conv_stem(x))
        a = a * torch.sigmoid(a)
        b = tv.features[0](x)
    assert a.shape == b.shape and rel(a, b) > 1e-2
