"""EnsemblePipeline: the batched hot path of the hybrid ensemble on one GPU.

Per batch of frames (uint8 RGB at ``output_size``, as data_processing leaves them):
  H2D -> [per model: fused BGR+bilinear resize kernel -> network (u8 in, thresholded u8 planes
  out, one CUDA-graph replay)] -> routing + nearest resize + label map + per-class counts kernel
  (-> radial thickness kernel) -> D2H of the 4-channel mask / label map / quantities.

It is what /root/reference/src/predict.py:61-101 (`segment`) does frame by frame and class by
class on torch ops; each model runs ONCE per frame here (the reference runs FC_LC twice).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import prepost as P
from .model import CLASS_IDS

# /root/reference/src/predict.py:23-28
MODELS_META = {
    'Lumen': {'model_dir': 'LM', 'index': 0},
    'Lipid core': {'model_dir': 'FC_LC', 'index': 0},
    'Fibrous cap': {'model_dir': 'FC_LC', 'index': 1},
    'Vasa vasorum': {'model_dir': 'VV', 'index': 0},
}


class EnsemblePipeline:
    def __init__(self, models: Dict[str, Tuple[object, Dict]], classes: Sequence[str], output_size: Sequence[int],
                 device, batch: int, src_hw: Optional[Tuple[int, int]] = None, thickness: bool = False):
        """models: {model_dir: (OCTSegmentationModel, cfg with 'input_size')} for every model_dir the
        requested classes need.  output_size: cv2 convention (width, height), as in the reference."""
        self.device = torch.device(device)
        self.classes = list(classes)
        for c in self.classes:
            if c not in MODELS_META:
                raise KeyError(c)
        self.Wo, self.Ho = int(output_size[0]), int(output_size[1])
        self.batch = batch
        self.thickness = thickness
        self.src_hw = tuple(src_hw) if src_hw is not None else (self.Ho, self.Wo)
        self.model_dirs = []
        for c in self.classes:
            d = MODELS_META[c]['model_dir']
            if d not in self.model_dirs:
                self.model_dirs.append(d)
        self.nets, self.sizes = {}, {}
        for d in self.model_dirs:
            model, cfg = models[d]
            S = int(cfg['input_size'])
            self.sizes[d] = S
            self.nets[d] = model.model.compiled(batch, S, S, self.device, 'u8', 'u8_nchw')
        self.order = [CLASS_IDS[c] - 1 for c in self.classes]
        # the networks are independent: each gets its own stream so small launches of one overlap the others
        with torch.cuda.device(self.device):
            self.streams = {d: torch.cuda.Stream() for d in self.model_dirs}
            self.ev_in = torch.cuda.Event()
            self.ev_done = {d: torch.cuda.Event() for d in self.model_dirs}
        Hs, Ws = self.src_hw
        with torch.cuda.device(self.device):
            self.frames_dev = torch.empty(batch, Hs, Ws, 3, dtype=torch.uint8, device=self.device)
            self.mask = torch.empty(batch, self.Ho, self.Wo, 4, dtype=torch.uint8, device=self.device)
            self.label = torch.empty(batch, self.Ho, self.Wo, dtype=torch.uint8, device=self.device)
            self.counts = torch.zeros(batch, 4, dtype=torch.int32, device=self.device)
        self.macs_per_frame = sum(self.nets[d].macs for d in self.model_dirs) / batch
        self.launches_per_batch = sum(self.nets[d].launches for d in self.model_dirs) + len(self.model_dirs) + 1 + int(thickness)

    # ------------------------------------------------------------------ device-resident step
    def run_device(self, frames_dev: torch.Tensor):
        """frames_dev: uint8 CUDA (batch, Hs, Ws, 3) RGB.  Returns device (mask, label, counts[, radii])."""
        class_planes = {}
        cur = torch.cuda.current_stream(self.device)
        self.ev_in.record(cur)
        for d in self.model_dirs:
            net = self.nets[d]
            st = self.streams[d]
            st.wait_event(self.ev_in)
            with torch.cuda.stream(st):
                P.preprocess(frames_dev, self.sizes[d], out=net.x_nhwc)
                out = net.run()                                          # (batch, C, S, S) uint8 {0,1}
                self.ev_done[d].record(st)
            for name in self.classes:
                meta = MODELS_META[name]
                if meta['model_dir'] == d:
                    class_planes[CLASS_IDS[name] - 1] = out[:, meta['index']]
        for d in self.model_dirs:
            cur.wait_event(self.ev_done[d])
        P.postprocess(class_planes, self.order, self.Ho, self.Wo, self.batch, self.device,
                      mask=self.mask, label=self.label, counts=self.counts)
        radii = P.radial_thickness(self.mask) if self.thickness else None
        return self.mask, self.label, self.counts, radii

    # ------------------------------------------------------------------ host-to-host step
    def _host_buffers(self):
        if getattr(self, '_pinned', None) is None:
            Hs, Ws = self.src_hw
            self._pinned = {
                'frames': torch.empty(self.batch, Hs, Ws, 3, dtype=torch.uint8).pin_memory(),
                'mask': torch.empty(self.batch, self.Ho, self.Wo, 4, dtype=torch.uint8).pin_memory(),
                'label': torch.empty(self.batch, self.Ho, self.Wo, dtype=torch.uint8).pin_memory(),
                'counts': torch.empty(self.batch, 4, dtype=torch.int32).pin_memory(),
                'radii': torch.empty(self.batch, 4, 360, dtype=torch.int32).pin_memory(),
            }
        return self._pinned

    def run_host(self, frames: np.ndarray, copy: bool = True):
        """frames: uint8 (n <= batch, Hs, Ws, 3) host array.  Returns host (mask, label, counts[, radii])
        for the n frames.  H2D and D2H copies (through pinned staging buffers, asynchronous, one
        synchronisation at the end) are part of this call.  With copy=False the returned arrays are
        views of the pinned buffers, valid until the next call."""
        n = frames.shape[0]
        assert n <= self.batch and tuple(frames.shape[1:3]) == self.src_hw
        hb = self._host_buffers()
        hb['frames'][:n].copy_(torch.from_numpy(frames))
        with torch.cuda.device(self.device):
            self.frames_dev[:n].copy_(hb['frames'][:n], non_blocking=True)
            if n < self.batch:
                self.frames_dev[n:].zero_()
            mask, label, counts, radii = self.run_device(self.frames_dev)
            hb['mask'][:n].copy_(mask[:n], non_blocking=True)
            hb['label'][:n].copy_(label[:n], non_blocking=True)
            hb['counts'][:n].copy_(counts[:n], non_blocking=True)
            if radii is not None:
                hb['radii'][:n].copy_(radii[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = [hb['mask'][:n].numpy(), hb['label'][:n].numpy(), hb['counts'][:n].numpy(),
               hb['radii'][:n].numpy() if radii is not None else None]
        if copy:
            out = [o.copy() if o is not None else None for o in out]
        return tuple(out)
