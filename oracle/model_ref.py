"""ORACLE (test infrastructure only): restatement of the reference's model wrapper and predict
driver, built on oracle/smp_ref.py instead of the absent smp / pytorch_lightning packages.

Follows /root/reference/src/models/smp/model.py:18-71,183-200 (OCTSegmentationModel ctor,
forward, predict), /root/reference/src/models/smp/utils.py:250-266 (pick_device) and
/root/reference/src/predict.py:23-101 (MODELS_META, load_model, preprocess_images, segment).
PARITY UNPINNED for logits (no golden vectors exist upstream).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import cv2
import numpy as np
import torch
import torch.nn as nn

from . import smp_ref

CLASS_MAP = {
    'Lumen': {'id': 1, 'color': [228, 30, 199]},
    'Fibrous cap': {'id': 2, 'color': [123, 171, 226]},
    'Lipid core': {'id': 3, 'color': [125, 227, 127]},
    'Vasa vasorum': {'id': 4, 'color': [208, 2, 27]},
}
CLASS_IDS = {k: v['id'] for k, v in CLASS_MAP.items()}

# predict.py:23-28 (kept verbatim, including the Lipid core -> channel 0 quirk)
MODELS_META = {
    'Lumen': {'model_dir': 'LM', 'index': 0},
    'Lipid core': {'model_dir': 'FC_LC', 'index': 0},
    'Fibrous cap': {'model_dir': 'FC_LC', 'index': 1},
    'Vasa vasorum': {'model_dir': 'VV', 'index': 0},
}


def pick_device(option: str) -> str:
    if option == 'auto':
        return 'cuda' if torch.cuda.is_available() else 'cpu'
    elif option in ['cpu', 'cuda']:
        return option
    raise ValueError("Invalid device option. Please specify 'cpu', 'cuda', or 'auto'.")


class OCTSegmentationModelRef(nn.Module):
    """model.py:18-71,183-200 without the Lightning/training hooks."""

    def __init__(self, arch: str, encoder_name: str, model_name: str, in_channels: int, classes: List[str], **kwargs):
        super().__init__()
        kwargs = {k: v for k, v in kwargs.items() if k in ('encoder_weights',)}
        self.model = smp_ref.create_model(arch=arch, encoder_name=encoder_name, in_channels=in_channels,
                                          classes=len(classes), **kwargs)
        self.classes = classes
        params = smp_ref.get_preprocessing_params(encoder_name)
        self.register_buffer('std', torch.tensor(params['std']).view(1, 3, 1, 1))
        self.register_buffer('mean', torch.tensor(params['mean']).view(1, 3, 1, 1))
        self.model_name = model_name

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        image = (image - self.mean) / self.std          # model.py:69 (inputs are 0..255; not "fixed")
        return self.model(image)

    def predict(self, images: np.ndarray, device: str) -> np.ndarray:
        images_tensor = torch.Tensor(images.transpose((0, 3, 1, 2))).to(device)   # model.py:189
        y_hat = self.model(images_tensor).cpu().detach()                          # no normalisation (model.py:192)
        masks = (y_hat.sigmoid() > 0.5).float()
        return masks.permute(0, 2, 3, 1).numpy().round()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, map_location=None, **ctor_kwargs):
        """pytorch_lightning 2.2.1 semantics: torch.load -> ctor(**kwargs) -> strict load -> .to()."""
        ckpt = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        model = cls(**ctor_kwargs)
        model.load_state_dict(ckpt['state_dict'], strict=True)
        return model.to(map_location) if map_location is not None else model


def load_model(model_dir: str, device: str) -> Tuple[OCTSegmentationModelRef, Dict]:
    with open(f'{model_dir}/config.json', 'r') as f:
        cfg = json.load(f)
    model = OCTSegmentationModelRef.load_from_checkpoint(
        checkpoint_path=f'{model_dir}/weights.ckpt', encoder_weights=None, arch=cfg['architecture'],
        encoder_name=cfg['encoder'], model_name=cfg['model_name'], in_channels=3, classes=cfg['classes'],
        map_location='cuda:0' if device == 'cuda' else device)
    model.eval()
    return model, cfg


def preprocessing_img(img, input_size: int) -> np.ndarray:
    """src/data/utils.py:159-166."""
    image = np.array(img)
    image = cv2.cvtColor(image, cv2.COLOR_RGB2BGR)
    return cv2.resize(image, (input_size, input_size))


def segment_with_models(images: List, masks: List[np.ndarray], output_size, classes: List[str],
                        models: Dict[str, Tuple[OCTSegmentationModelRef, Dict]], device: str) -> List[np.ndarray]:
    """predict.py:61-101 with the models already in memory (keyed by MODELS_META model_dir)."""
    for class_name in classes:
        meta = MODELS_META[class_name]
        model, cfg = models[meta['model_dir']]
        processed = np.array([preprocessing_img(img, cfg['input_size']) for img in images])
        for img, mask in zip(processed, masks):
            with torch.no_grad():
                pm = model.predict(images=np.array([img]), device=device)[0]
            rm = cv2.resize(pm, tuple(output_size), interpolation=cv2.INTER_NEAREST)
            if rm.ndim > 2:
                rm = rm[:, :, meta['index']]
            mask[:, :, CLASS_IDS[class_name] - 1] = rm
    return masks
