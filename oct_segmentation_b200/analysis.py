"""DICOM front-end of the reference's app, wired to the B200 pipeline.

`get_analysis` of /root/reference/src/app/tools/analysis.py:133-213 reads a DICOM study, normalises every slice
(`cv2.normalize(NORM_MINMAX) -> uint8`, `cv2.cvtColor(BGR2RGB)`, :166-177), and -- behind a `TODO: run inference`
(:138,:166) -- expects 4-channel masks from which it builds the per-class `objects` table (slices, object ids,
area, contour thickness, base64 PNG of the mask; :185-213).  `analyse_volume` is that data path with the TODO filled
in: slices -> normalise (host cv2, the reference's own two calls) -> batched GPU ensemble (`predict.segment`, which
also produces the counts / contour thickness on the GPU) -> the same `data` dict.  The Gradio / plotly widgets that
`get_analysis` returns around it are out of scope.

`file` may be a path (needs `pydicom`, which this image does not ship: a clear ImportError otherwise) or the pixel
array itself, shape (slices, H, W) or (slices, H, W, 3).
"""
from __future__ import annotations

import base64
from io import BytesIO
from typing import Any, Dict, List, Sequence, Union

import cv2
import numpy as np
from PIL import Image

from . import predict as P
from . import prepost
from .model import CLASS_IDS

CLASS_NAMES = list(CLASS_IDS)


def read_dicom(file: str) -> np.ndarray:
    """analysis.py:139-140: `pydicom.dcmread(file).pixel_array`."""
    try:
        import pydicom
    except ImportError as e:          # not installable offline
        raise ImportError('reading a DICOM file needs pydicom; pass the pixel array (slices, H, W[, 3]) instead') from e
    return pydicom.dcmread(file).pixel_array


def normalise_slice(img: np.ndarray) -> np.ndarray:
    """analysis.py:166-177, verbatim: min-max normalise to uint8, then BGR2RGB (a 2-D slice is replicated into three
    channels first -- cv2.cvtColor(BGR2RGB) needs 3 -- the pipeline's documented grayscale extension)."""
    img = cv2.normalize(img, None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_8U)
    if img.ndim == 2:
        img = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def analyse_volume(file: Union[str, np.ndarray], models: Dict, classes: Sequence[str] = tuple(CLASS_NAMES),
                   output_size: Sequence[int] = (1000, 1000), batch_size: int = 16, with_masks: bool = True) -> Dict[str, Any]:
    """The `data` dict of get_analysis (analysis.py:142-160, filled by the loop :185-213) for a DICOM volume:
    {'ratio': int(H_dcm * 150 // 1000), 'objects': {class: {area, thickness_mean, thickness_min, slice, object_id,
    masks (base64 PNG of the {0,255} class mask), img_name}}, 'images': [names]}.
    models: {model_dir: (OCTSegmentationModel, cfg)} as `predict.load_models` returns."""
    dcm = read_dicom(file) if isinstance(file, str) else np.asarray(file)
    if dcm.ndim not in (3, 4):
        raise ValueError(f'expected a (slices, H, W[, 3]) pixel array, got shape {dcm.shape}')
    ratio = int(dcm.shape[1] * 150 // 1000)                                           # analysis.py:155
    if ratio < 1:
        raise ValueError('DICOM frames are too small for the area / thickness scale (ratio = H * 150 // 1000 is 0)')
    names = [f'{i + 1:03d}' for i in range(dcm.shape[0])]
    images = [Image.fromarray(normalise_slice(dcm[i])).resize(tuple(output_size)) for i in range(dcm.shape[0])]   # data_processing, :187
    masks = [np.zeros((output_size[1], output_size[0], 4)) for _ in images]
    rows: List[Dict] = []
    P.segment(images, masks, output_size, list(classes), models_dir='', device='cuda', batch_size=batch_size, models=models,
              quantities=rows, ratio=ratio)
    objects = prepost.objects_table(rows, names)
    for name in CLASS_NAMES:
        objects[name]['masks'] = []
        if with_masks:
            c = CLASS_IDS[name] - 1
            for idx in objects[name]['slice']:
                buff = BytesIO()
                Image.fromarray((masks[idx][:, :, c] != 0).astype(np.uint8) * 255).save(buff, format='png')   # :208-210
                objects[name]['masks'].append(base64.b64encode(buff.getvalue()).decode('utf-8'))
    return {'ratio': ratio, 'objects': objects, 'images': names}
