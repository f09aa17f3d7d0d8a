"""Predict entry point — same functions, arguments and config keys as
/root/reference/src/predict.py (MODELS_META :23-28, load_model :31-50, preprocess_images :53-58,
segment :61-101, main :109-149) and the helpers it imports from src/data/utils.py
(data_processing :169-192, save_results :195-235), executed on the octseg B200 engine.

Differences that do not change results: frames go through the GPU in batches
(`cfg.batch_size`), each model runs once per frame even when it serves two classes, and resize /
threshold / routing / label map / pixel counts are CUDA kernels (bit-exact vs cv2, see tests).
"""
from __future__ import annotations

import json
import logging
import os
import sys
import time
from glob import glob
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from PIL import Image

from . import config as cfgmod
from . import prepost as P
from .model import CLASS_COLORS_RGB, CLASS_IDS, OCTSegmentationModel
from .pipeline import MODELS_META, EnsemblePipeline

log = logging.getLogger(__name__)
log.setLevel(logging.INFO)

PROJECT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKGROUND_RGB = (128, 128, 128)


def pick_device(option: str) -> str:
    """/root/reference/src/models/smp/utils.py:250-266."""
    if option == 'auto':
        return 'cuda' if torch.cuda.is_available() else 'cpu'
    elif option in ['cpu', 'cuda']:
        return option
    else:
        raise ValueError("Invalid device option. Please specify 'cpu', 'cuda', or 'auto'.")


def load_model(model_dir: str, device: str) -> Tuple[OCTSegmentationModel, Dict]:
    """Load a segmentation model from checkpoint (config.json + weights.ckpt)."""
    with open(f'{model_dir}/config.json', 'r') as file:
        model_cfg = json.load(file)
    model = OCTSegmentationModel.load_from_checkpoint(
        checkpoint_path=f'{model_dir}/weights.ckpt',
        encoder_weights=None,
        arch=model_cfg['architecture'],
        encoder_name=model_cfg['encoder'],
        model_name=model_cfg['model_name'],
        in_channels=3,
        classes=model_cfg['classes'],
        map_location='cuda:0' if device == 'cuda' else device,
    )
    model.eval()
    return model, model_cfg


def preprocess_images(images: List[Image.Image], input_size: int, device: str = 'cuda') -> np.ndarray:
    """(N, S, S, 3) uint8 BGR, identical to [preprocessing_img(img, S) for img in images] but resized
    by the CUDA kernel."""
    frames = np.stack([np.array(img.convert('RGB') if img.mode != 'RGB' else img) for img in images])
    dev = torch.device('cuda:0' if device == 'cuda' else device)
    return P.preprocess(torch.from_numpy(frames).to(dev), input_size).cpu().numpy()


def segment(images: List[Image.Image], masks: List[np.ndarray], output_size: Sequence[int], classes: Sequence[str],
            models_dir: str, device: str, batch_size: int = 16, models: Dict = None,
            quantities: List = None, pipe: EnsemblePipeline = None, pool=None, ratio: int = None) -> List[np.ndarray]:
    """Perform segmentation for given images using specified models; fills and returns ``masks``
    (the caller's float64 HxWx4 arrays), channel CLASS_IDS[name]-1 per requested class.

    ``pipe``: an EnsemblePipeline from ``make_pipeline`` to reuse across calls (the streaming main() builds it
    once: networks are compiled for ``batch_size`` frames, staging buffers and streams are allocated once;
    a short last batch is zero-filled, never recompiled).  ``pool``: executor that spreads the uint8 -> float64 fill of
    the caller's masks (32 MB per frame at 1000 x 1000, the reference's format) over host threads."""
    if device != 'cuda':
        raise RuntimeError('the B200 build of segment() runs on CUDA only (device resolved to %r)' % device)
    if not images:
        return masks
    dev = torch.device('cuda:0')
    if models is None:
        models = load_models(models_dir, classes, device)
    frames = np.stack([np.array(img.convert('RGB') if img.mode != 'RGB' else img) for img in images])
    n = frames.shape[0]
    if pipe is None:
        pipe = make_pipeline(models, classes, output_size, min(batch_size, n), frames.shape[1:3], quantities is not None, dev)
    elif (pipe.src_hw != tuple(frames.shape[1:3]) or pipe.classes != list(classes)
          or (pipe.Wo, pipe.Ho) != (int(output_size[0]), int(output_size[1])) or (quantities is not None and not pipe.thickness)):
        raise ValueError('segment(): the pipeline passed in was built for a different frame size / class list / output size')
    batch = pipe.batch
    spans = [(lo, min(lo + batch, n)) for lo in range(0, n, batch)]
    # copies of batch i+1 / i-1 overlap the compute of batch i (EnsemblePipeline.stream_host)
    for (lo, hi), (mask, label, counts, radii, *contours) in zip(spans, pipe.stream_host(frames[lo:hi] for lo, hi in spans)):
        idxs = [CLASS_IDS[class_name] - 1 for class_name in classes]

        def fill(i, lo=lo, mask=mask):
            if len(idxs) == 4:
                masks[i][...] = mask[i - lo]
            else:
                for idx in idxs:
                    masks[i][:, :, idx] = mask[i - lo, :, :, idx]
        if pool is not None:
            list(pool.map(fill, range(lo, hi)))          # numpy copies release the GIL
        else:
            for i in range(lo, hi):
                fill(i)
        if quantities is not None:
            ratio = P.dicom_ratio(pipe.Ho) if ratio is None else ratio      # analysis.py:155 (the DICOM's height when known)
            quantities.extend(P.quantities_from_counts(counts, pipe.Ho, pipe.Wo, ratio, radii,
                                                       contours[0] if contours else None, masks=mask))
    return masks


def make_pipeline(models: Dict, classes: Sequence[str], output_size: Sequence[int], batch: int, src_hw, quantities: bool,
                  dev=None) -> EnsemblePipeline:
    """The batched GPU pipeline segment() drives: one compiled network per model for ``batch`` frames of ``src_hw``."""
    dev = torch.device('cuda:0') if dev is None else dev
    return EnsemblePipeline(models, classes, output_size, dev, batch, src_hw=tuple(src_hw), thickness=quantities,
                            contour=quantities and P.contour_fits(int(output_size[1]), int(output_size[0])))


def list_images(data_path: str) -> List[str]:
    """The file set of data_processing (src/data/utils.py:175-178), in sorted order."""
    if os.path.isfile(data_path):
        return [data_path]
    return sorted(glob(f'{data_path}/*.[pj][np][ge]*'))


def open_images(images_path: Sequence[str], output_size: Sequence[int], pool=None):
    """Loop body of data_processing (src/data/utils.py:180-191) for the given files; decoding + bicubic resize
    run on ``pool`` (a concurrent.futures executor) when given."""
    def one(img_path):
        return Image.open(img_path).resize(tuple(output_size))
    images = list(pool.map(one, images_path)) if pool is not None else [one(q) for q in images_path]
    masks = [np.zeros((output_size[0], output_size[1], 4)) for _ in images_path]
    image_names = [os.path.basename(img_path).split('.')[0] for img_path in images_path]
    return images, masks, image_names


def data_processing(data_path: str, save_dir: str, output_size: Sequence[int]):
    """src/data/utils.py:169-192: open every image, PIL-resize (bicubic) to output_size, allocate the
    float64 HxWx4 mask.  File order is sorted (the reference's glob order is unspecified)."""
    os.makedirs(save_dir, exist_ok=True)
    return open_images(list_images(data_path), output_size)


def color_mask(mask: np.ndarray, classes: Sequence[str]) -> np.ndarray:
    """Colour mask of save_results (src/data/utils.py:208,231-233): grey background, classes painted
    in ``classes`` order so later classes overwrite earlier ones."""
    out = np.empty(mask.shape[:2] + (3,), np.uint8)
    out[:] = BACKGROUND_RGB
    for class_name in classes:
        out[mask[:, :, CLASS_IDS[class_name] - 1] != 0] = CLASS_COLORS_RGB[class_name]
    return out


def save_results(images, masks, images_name, classes, save_dir: str, batch_size: int = 16, device: str = 'cuda',
                 pool=None) -> None:
    """src/data/utils.py:195-235: writes <name>_mask.png (priority colour mask) and <name>_overlay.png
    (closed + blurred fill and dilate/erode rim pasted per class over the frame).  The overlay is computed
    by ``octseg_overlay`` in batches on the GPU, bit-exact vs the reference's cv2 + PIL arithmetic
    (tests/golden/overlay_ref.npz); PNG encoding stays on the host (PIL), like the reference's; ``pool`` spreads it over threads."""
    order = [CLASS_IDS[c] - 1 for c in classes]
    dev = torch.device(device)
    for lo in range(0, len(images), batch_size):
        hi = min(lo + batch_size, len(images))
        frames = np.stack([np.asarray(img.convert('RGB')) for img in images[lo:hi]])
        m8 = np.stack([(np.asarray(m) != 0).astype(np.uint8) for m in masks[lo:hi]])
        over = P.overlay(torch.from_numpy(frames).to(dev), torch.from_numpy(m8).to(dev), order).cpu().numpy()
        def write(k, lo=lo, over=over):
            Image.fromarray(over[k]).save(f'{save_dir}/{images_name[lo + k]}_overlay.png')
            Image.fromarray(color_mask(masks[lo + k], classes)).save(f'{save_dir}/{images_name[lo + k]}_mask.png')
        if pool is not None:
            list(pool.map(write, range(hi - lo)))       # PNG encoding (zlib) releases the GIL
        else:
            for k in range(hi - lo):
                write(k)


def load_models(models_dir: str, classes: Sequence[str], device: str) -> Dict:
    """{model_dir: (model, cfg)} for the requested classes (the loading loop of segment, src/predict.py:70-76);
    a model_dir holding fold_*/ sub-directories instead of config.json loads as a list of folds (opt-in
    probability averaging, see EnsemblePipeline)."""
    models = {}
    for class_name in classes:
        mdir = MODELS_META[class_name]['model_dir']
        if mdir in models:
            continue
        start_load = time.time()
        mpath = os.path.join(models_dir, mdir)
        fold_dirs = sorted(glob(f'{mpath}/fold_*/config.json'))
        if fold_dirs and not os.path.exists(f'{mpath}/config.json'):
            models[mdir] = [load_model(model_dir=os.path.dirname(f), device=device) for f in fold_dirs]
            arch = f"{models[mdir][0][1]['architecture']} x {len(fold_dirs)} folds"
        else:
            models[mdir] = load_model(model_dir=mpath, device=device)
            arch = models[mdir][1]['architecture']
        log.info(f"{arch} loaded successfully. Time taken: {time.time() - start_load:.1f} s")
    return models


def main(cfg) -> None:
    """Main function to perform OCT image segmentation prediction (src/predict.py:109-149).

    ``stream_chunk: K`` (extension, 0 = the reference's hold-everything flow): the file list is walked in chunks
    of K frames -- chunk i+1 is decoded and chunk i-1 is PNG-encoded on ``io_workers`` host threads while chunk i
    is on the GPU, and each chunk's images / float64 masks (35 MB per frame at 1000 x 1000) are released after
    its results are written, so the frame count is not bounded by host memory.
    Results are identical to the non-streaming flow."""
    log.info(f'Config:\n\n{cfgmod.to_yaml(cfg)}')
    device = pick_device(option=cfg.device)
    data_dir = str(os.path.join(PROJECT_DIR, cfg.data_dir))
    models_dir = str(os.path.join(PROJECT_DIR, cfg.models_dir))
    save_dir = str(os.path.join(PROJECT_DIR, cfg.save_dir))
    batch_size = int(cfg.get('batch_size', 16))
    want_quantities = bool(cfg.get('quantities', False))
    chunk = int(cfg.get('stream_chunk', 0) or 0)

    start = time.time()
    if chunk <= 0:
        images, masks, images_name = data_processing(data_path=data_dir, save_dir=save_dir, output_size=cfg.output_size)
        log.info(f'Number of images: {len(images_name)}')
        start_inference = time.time()
        quantities = [] if want_quantities else None
        masks = segment(images=images, masks=masks, output_size=cfg.output_size, classes=cfg.classes,
                        models_dir=models_dir, device=device, batch_size=batch_size, quantities=quantities)
        log.info(f'Prediction time: {time.time() - start_inference:.1f} s')
        save_results(images=images, masks=masks, images_name=images_name, classes=cfg.classes, save_dir=save_dir,
                     batch_size=batch_size)
        table = dict(zip(images_name, quantities)) if want_quantities else None
    else:
        from concurrent.futures import ThreadPoolExecutor
        os.makedirs(save_dir, exist_ok=True)
        paths = list_images(data_dir)
        log.info(f'Number of images: {len(paths)} (streamed in chunks of {chunk})')
        if device != 'cuda':
            raise RuntimeError('the B200 build of segment() runs on CUDA only (device resolved to %r)' % device)
        models = load_models(models_dir, cfg.classes, device) if paths else {}
        table = {} if want_quantities else None
        spans = [(lo, min(lo + chunk, len(paths))) for lo in range(0, len(paths), chunk)]
        # ONE pipeline for the whole run: data_processing resizes every frame to output_size, so all chunks share it
        pipe = make_pipeline(models, cfg.classes, cfg.output_size, min(batch_size, chunk),
                             (int(cfg.output_size[1]), int(cfg.output_size[0])), want_quantities) if paths else None
        with ThreadPoolExecutor(max_workers=int(cfg.get('io_workers', 8))) as io, \
                ThreadPoolExecutor(max_workers=2) as stage:
            def load(span):
                return open_images(paths[span[0]:span[1]], cfg.output_size, pool=io)
            nxt = stage.submit(load, spans[0]) if spans else None
            saving = None
            for k, span in enumerate(spans):
                images, masks, names = nxt.result()
                nxt = stage.submit(load, spans[k + 1]) if k + 1 < len(spans) else None
                quantities = [] if want_quantities else None
                masks = segment(images=images, masks=masks, output_size=cfg.output_size, classes=cfg.classes,
                                models_dir=models_dir, device=device, batch_size=batch_size, models=models,
                                quantities=quantities, pipe=pipe, pool=io)
                if want_quantities:
                    table.update(zip(names, quantities))
                if saving is not None:
                    saving.result()
                saving = stage.submit(save_results, images, masks, names, cfg.classes, save_dir, batch_size, 'cuda', io)
            if saving is not None:
                saving.result()
        log.info(f'Prediction + I/O time: {time.time() - start:.1f} s')
    if table is not None:
        with open(os.path.join(save_dir, 'quantities.json'), 'w') as f:
            json.dump(table, f, indent=1)
        # the per-class object table of get_analysis (src/app/tools/analysis.py:185-213): slices, object ids, areas, thickness
        names = list(table)
        with open(os.path.join(save_dir, 'objects.json'), 'w') as f:
            json.dump({'ratio': P.dicom_ratio(int(cfg.output_size[1])), 'images': names,
                       'objects': P.objects_table([table[k] for k in names], names)}, f, indent=1)
    log.info(f'Overall computation time: {time.time() - start:.1f} s')
    log.info('Complete')


def cli(argv: Sequence[str] = None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    cfg = cfgmod.compose(os.path.join(PROJECT_DIR, 'configs'), 'predict', argv)
    jl = cfg.get('hydra', {}).get('job_logging', {})
    logging.basicConfig(level=getattr(logging, str(jl.get('level', 'INFO'))),
                        format=jl.get('format', '[%(asctime)s][%(levelname)s] - %(message)s'),
                        datefmt=jl.get('datefmt', '%d-%m-%Y %H:%M:%S'), stream=sys.stdout)
    main(cfg)


if __name__ == '__main__':
    cli()
