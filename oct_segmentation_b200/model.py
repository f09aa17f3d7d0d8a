"""OCTSegmentationModel — inference surface of /root/reference/src/models/smp/model.py:18-71,
183-200 on the octseg engine: same constructor arguments, ``forward`` (normalise + net),
``predict`` (net + sigmoid + 0.5 threshold, NHWC numpy in/out), ``eval`` and the
``load_from_checkpoint`` classmethod the reference gets from pytorch_lightning
(/root/reference/src/predict.py:39-48).  Training hooks are out of scope.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import smp

CLASS_MAP = {
    'Lumen': {'id': 1, 'color': [228, 30, 199]},
    'Fibrous cap': {'id': 2, 'color': [123, 171, 226]},
    'Lipid core': {'id': 3, 'color': [125, 227, 127]},
    'Vasa vasorum': {'id': 4, 'color': [208, 2, 27]},
}
CLASS_IDS = {name: info['id'] for name, info in CLASS_MAP.items()}
CLASS_IDS_REVERSED = {v: k for k, v in CLASS_IDS.items()}
CLASS_COLORS_RGB = {name: tuple(info['color']) for name, info in CLASS_MAP.items()}


class OCTSegmentationModel(nn.Module):
    """The model dedicated to the segmentation of OCT images (B200 inference build)."""

    def __init__(self, arch: str, encoder_name: str, model_name: str, in_channels: int, classes: List[str],
                 lr: float = 0.0001, data_dir: Optional[str] = None, weight_decay: float = 0.0001,
                 optimizer_name: str = 'Adam', input_size: int = 512, img_save_interval: Optional[int] = 1,
                 save_wandb_media: bool = False, **kwargs):
        super().__init__()
        self.model = smp.create_model(arch=arch, encoder_name=encoder_name, in_channels=in_channels,
                                      classes=len(classes), **kwargs)
        self.classes = classes
        self.data_dir = data_dir
        params = smp.encoders.get_preprocessing_params(encoder_name)
        self.register_buffer('std', torch.tensor(params['std']).view(1, 3, 1, 1))
        self.register_buffer('mean', torch.tensor(params['mean']).view(1, 3, 1, 1))
        self.model_name = model_name
        self.lr, self.weight_decay, self.optimizer = lr, weight_decay, optimizer_name
        self.input_size = input_size
        self.img_save_interval, self.save_wandb_media = img_save_interval, save_wandb_media
        self.class_values = [CLASS_IDS[cl] for cl in self.classes]

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        """(image - mean) / std, then the network (model.py:65-71).  The normalisation is applied
        inside the stem kernel's input load instead of as a separate elementwise pass."""
        norm = (self.mean.flatten().tolist(), self.std.flatten().tolist())
        return self.model(image, _norm=norm)

    @torch.no_grad()
    def predict(self, images: np.ndarray, device: str) -> np.ndarray:
        """images: (N, H, W, C) -> (N, H, W, classes) float32 {0,1}; no normalisation (model.py:192)."""
        dev = torch.device('cuda:0' if device == 'cuda' else device)
        if dev.type != 'cuda':
            raise RuntimeError('OCTSegmentationModel (B200 build) predicts on CUDA only')
        n, h, w, _ = images.shape
        self.model.check_input_shape(torch.empty(0, 3, h, w))
        if images.dtype == np.uint8:
            x = torch.from_numpy(np.ascontiguousarray(images)).to(dev, non_blocking=True)
            in_dtype = 'u8'
        else:
            x = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).to(dev, non_blocking=True)
            in_dtype = 'f32'
        net = self.model.compiled(n, h, w, dev, in_dtype, 'u8_nchw')
        net.x_nhwc.copy_(x)
        masks = net.run()                                   # (N, C, H, W) uint8: logit > 0  ==  sigmoid > 0.5
        return masks.permute(0, 2, 3, 1).cpu().numpy().astype(np.float32)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, map_location=None, **ctor_kwargs) -> 'OCTSegmentationModel':
        """pytorch_lightning semantics for a class without saved hyper-parameters: torch.load ->
        cls(**ctor_kwargs) -> load_state_dict(strict=True) -> .to(map_location)."""
        ckpt = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        model = cls(**ctor_kwargs)
        model.load_state_dict(ckpt['state_dict'], strict=True)
        model.model.invalidate()
        if map_location is not None:
            model = model.to(map_location)
        return model
