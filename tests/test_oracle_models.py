"""CPU: pins the network oracle (oracle/smp_ref.py) to the only hard evidence available for the
third-party arithmetic: parameter counts that reproduce well-known smp
totals and the reference's DVC checkpoint sizes, MAC counts, padding tables, key sets, shapes."""
import pytest
import torch
import torch.nn as nn

from oct_segmentation_b200.engine import builder as B
from oct_segmentation_b200 import smp as our_smp
from oracle import smp_ref, synth


def count(m):
    return sum(p.numel() for p in m.parameters())


@pytest.mark.parametrize('arch,enc,want', [
    ('Unet', 'resnet34', 24436369), ('UnetPlusPlus', 'resnet34', 26078609), ('Linknet', 'resnet34', 21771937),
    ('Unet', 'resnet50', 32521105), ('UnetPlusPlus', 'resnet50', 48985745), ('Linknet', 'resnet50', 31177777)])
def test_known_smp_parameter_totals(arch, enc, want):
    assert count(smp_ref.create_model(arch, enc, classes=1)) == want


def test_shipped_models_parameter_counts_and_checkpoint_sizes():
    lm = smp_ref.create_model('UnetPlusPlus', 'resnet101', classes=1)
    fc = smp_ref.create_model('LinkNet', 'efficientnet-b7', classes=2)
    vv = smp_ref.create_model('Unet', 'timm-regnetx_064', classes=1)
    assert count(lm.encoder) == 42500160 and count(lm.decoder) == 25477568 and count(lm.segmentation_head) == 145
    assert abs(count(fc.encoder) / 1e6 - 63.79) < 0.01 and abs((count(fc.decoder) + count(fc.segmentation_head)) / 1e6 - 0.636) < 1e-3
    assert abs(count(vv.encoder) / 1e6 - 24.58) < 0.01 and abs((count(vv.decoder) + count(vv.segmentation_head)) / 1e6 - 7.29) < 0.01
    # PL checkpoints = weights + optimizer state: RMSprop (LM, FC_LC) 2x4 B/param, RAdam (VV) 3x4 B/param
    # vs sizes recorded in /root/reference/models/{LM,FC_LC,VV}.dvc (544.66 / 510.81 / 383.06 MB)
    for model, mult, dvc_mb in ((lm, 8, 544.66), (fc, 8, 510.81), (vv, 12, 383.06)):
        assert abs(count(model) * mult / 1e6 - dvc_mb) / dvc_mb < 0.01


def test_regnet_widths_and_efficientnet_static_padding():
    assert smp_ref.regnet_widths() == ([168, 392, 784, 1624], [2, 4, 10, 1])
    enc = smp_ref.EfficientNetB7Encoder()
    assert enc._conv_stem.static_pad == (0, 1, 0, 1)
    s2 = {i: b._depthwise_conv.static_pad for i, b in enumerate(enc._blocks) if b.stride == 2}
    assert s2 == {4: (0, 1, 0, 1), 11: (1, 2, 1, 2), 18: (1, 1, 1, 1), 38: (1, 2, 1, 2)}
    assert len(enc._blocks) == 55
    ours = our_smp.get_encoder('efficientnet-b7')
    assert [b._depthwise_conv.pad for b in ours._blocks] == [b._depthwise_conv.static_pad[:2] for b in enc._blocks]


@pytest.mark.parametrize('arch,enc,cls', [('UnetPlusPlus', 'resnet101', 1), ('LinkNet', 'efficientnet-b7', 2), ('Unet', 'timm-regnetx_064', 1)])
def test_state_dict_keys_match_product_schema(arch, enc, cls):
    a, b = our_smp.create_model(arch, enc, classes=cls).state_dict(), smp_ref.create_model(arch, enc, classes=cls).state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in b)
    assert 'segmentation_head.0.weight' in a and any(k.startswith('encoder.') for k in a)


def conv_macs(model, x):
    total = [0]

    def hook(m, inp, out):
        if isinstance(m, nn.ConvTranspose2d):
            total[0] += inp[0].shape[2] * inp[0].shape[3] * m.in_channels * m.out_channels * 16
        else:
            total[0] += out.shape[2] * out.shape[3] * m.out_channels * (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]
    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    with torch.no_grad():
        y = model(x)
    for h in hs:
        h.remove()
    return total[0], y


@pytest.mark.parametrize('key,size,gflop', [('LM', 512, 498.3), ('VV', 896, 314.5), ('FC_LC', 896, 167.8)])
def test_mac_counts_match_survey(key, size, gflop):
    cfg = synth.MODEL_CONFIGS[key]
    m = smp_ref.create_model(cfg['architecture'], cfg['encoder'], classes=len(cfg['classes'])).eval()
    macs, y = conv_macs(m, torch.zeros(1, 3, size, size))
    assert y.shape == (1, len(cfg['classes']), size, size)
    # SE 1x1 convs on pooled vectors are counted by the hook but are ~1e-4 of the total
    assert abs(2 * macs / 1e9 - gflop) / gflop < 2e-3, 2 * macs / 1e9


def test_input_shape_must_be_divisible_by_32():
    m = smp_ref.create_model('Unet', 'timm-regnetx_064', classes=1).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 100, 128))
    with pytest.raises(KeyError):
        smp_ref.create_model('nope', 'resnet101')


def test_fold_bn_matches_batchnorm():
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(8, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16, eps=1e-3)
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_()
    bn.eval()
    x = torch.randn(2, 8, 10, 10)
    w, b = B.fold_bn(conv.weight, bn)
    assert torch.allclose(nn.functional.conv2d(x, w, b, padding=1), bn(conv(x)), atol=1e-5)
    ct, bn2 = nn.ConvTranspose2d(8, 8, 4, 2, 1), nn.BatchNorm2d(8)
    bn2.running_mean.normal_()
    bn2.running_var.uniform_(0.5, 2)
    bn2.eval()
    w, b = B.fold_bn(ct.weight, bn2, conv_bias=ct.bias, out_dim=1)
    assert torch.allclose(nn.functional.conv_transpose2d(x, w, b, stride=2, padding=1), bn2(ct(x)), atol=1e-5)
