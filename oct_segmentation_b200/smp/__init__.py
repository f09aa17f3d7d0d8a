"""Drop-in for the slice of ``segmentation_models_pytorch`` the reference's inference path uses
(/root/reference/src/models/smp/model.py:38-44,49): ``create_model``, ``Unet``, ``UnetPlusPlus``,
``Linknet`` and ``encoders.get_preprocessing_params``.

``forward(x)`` keeps smp's contract (float NCHW in, float32 logits NCHW out, H and W divisible
by 32 else RuntimeError) but executes on the octseg CUDA engine (tcgen05 implicit-GEMM convs,
fused upsample/concat/residual) instead of torch ops.  No CPU fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import modules as M
from ..engine.network import CompiledNet

_IMAGENET = {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225], 'input_space': 'RGB',
             'input_range': [0, 1]}


def get_preprocessing_params(encoder_name: str, pretrained: str = 'imagenet') -> Dict:
    if encoder_name not in M.ENCODERS:
        raise KeyError(f'Wrong encoder name `{encoder_name}`, supported encoders: {list(M.ENCODERS)}')
    return dict(_IMAGENET)


def get_encoder(name: str, in_channels: int = 3, depth: int = 5, weights: Optional[str] = None) -> nn.Module:
    if name not in M.ENCODERS:
        raise KeyError(f'Wrong encoder name `{name}`, supported encoders: {list(M.ENCODERS)}')
    if in_channels != 3 or depth != 5:
        raise ValueError('the B200 path implements the reference configuration: in_channels=3, encoder_depth=5')
    if weights is not None:
        raise ValueError('pretrained encoder weights are not bundled; the reference passes encoder_weights=None')
    return M.ENCODERS[name]()


encoders = SimpleNamespace(get_preprocessing_params=get_preprocessing_params, get_encoder=get_encoder)


class SegmentationModel(nn.Module):
    """Parameter tree (.encoder/.decoder/.segmentation_head) + engine-backed forward."""

    MAX_COMPILED = 4     # compiled (batch, size, dtype, mode) plans kept per model; each owns an activation arena

    def _post_init(self):
        M.init_decoder(self.decoder)
        M.init_head(self.segmentation_head)
        self._compiled: Dict[Tuple, CompiledNet] = {}

    def check_input_shape(self, x):
        h, w = x.shape[-2:]
        if h % 32 != 0 or w % 32 != 0:
            nh = (h // 32 + 1) * 32 if h % 32 else h
            nw = (w // 32 + 1) * 32 if w % 32 else w
            raise RuntimeError(
                f'Wrong input shape height={h}, width={w}. Expected image height and width divisible by 32. '
                f'Consider pad your images to shape ({nh}, {nw}).')

    def invalidate(self) -> None:
        """Drop compiled plans (call after changing parameters)."""
        self._compiled = {}

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate()
        return r

    def compiled(self, N: int, H: int, W: int, device, in_dtype: str = 'f32', out_mode: str = 'f32_nchw',
                 norm=None, use_graph: bool = True) -> CompiledNet:
        key = (N, H, W, str(device), in_dtype, out_mode, None if norm is None else (tuple(norm[0]), tuple(norm[1])))
        net = self._compiled.pop(key, None)
        if net is None:
            net = CompiledNet(self, N, H, W, device, in_dtype, out_mode, norm, use_graph)
            while len(self._compiled) >= self.MAX_COMPILED:        # least recently used plan (and its arena) goes
                self._compiled.pop(next(iter(self._compiled)))
        self._compiled[key] = net                                  # most recently used last
        return net

    def forward(self, x: torch.Tensor, _norm=None) -> torch.Tensor:
        self.check_input_shape(x)
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f'expected input of shape (N, 3, H, W), got {tuple(x.shape)}')
        if not x.is_cuda:
            raise RuntimeError('octseg models execute on a CUDA (sm_100a) device only; move the input with .to("cuda")')
        in_dtype = 'u8' if x.dtype == torch.uint8 else 'f32'
        if in_dtype == 'f32' and x.dtype != torch.float32:
            x = x.float()
        net = self.compiled(x.shape[0], x.shape[2], x.shape[3], x.device, in_dtype, 'f32_nchw', _norm)
        return net(x).clone()


def _build(self, encoder_name, encoder_weights, in_channels, classes, decoder_cls, head_in, head_k):
    nn.Module.__init__(self)
    self.encoder = get_encoder(encoder_name, in_channels, 5, encoder_weights)
    self.decoder = decoder_cls(self.encoder.out_channels)
    self.segmentation_head = M.SegmentationHeadParams(head_in, classes, head_k)
    self.classification_head = None
    self.name = f'{type(self).__name__.lower()}-{encoder_name}'
    self._post_init()


class Unet(SegmentationModel):
    def __init__(self, encoder_name='resnet101', encoder_weights=None, in_channels=3, classes=1, **kwargs):
        _build(self, encoder_name, encoder_weights, in_channels, classes, M.UnetDecoderParams, 16, 3)


class UnetPlusPlus(SegmentationModel):
    def __init__(self, encoder_name='resnet101', encoder_weights=None, in_channels=3, classes=1, **kwargs):
        _build(self, encoder_name, encoder_weights, in_channels, classes, M.UnetPlusPlusDecoderParams, 16, 3)


class Linknet(SegmentationModel):
    def __init__(self, encoder_name='resnet101', encoder_weights=None, in_channels=3, classes=1, **kwargs):
        _build(self, encoder_name, encoder_weights, in_channels, classes, M.LinknetDecoderParams, 32, 1)


_ARCHS = {cls.__name__.lower(): cls for cls in (Unet, UnetPlusPlus, Linknet)}


def create_model(arch: str, encoder_name: str = 'resnet34', encoder_weights: Optional[str] = None,
                 in_channels: int = 3, classes: int = 1, **kwargs) -> nn.Module:
    """smp.create_model: case-insensitive architecture lookup; KeyError on unknown arch/encoder."""
    try:
        cls = _ARCHS[arch.lower()]
    except KeyError:
        raise KeyError(f'Wrong architecture type `{arch}`. Available options are: {list(_ARCHS)}')
    return cls(encoder_name=encoder_name, encoder_weights=encoder_weights, in_channels=in_channels,
               classes=classes, **kwargs)
