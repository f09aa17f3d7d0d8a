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


def _copy(o):
    if o is None:
        return None
    return tuple(a.copy() for a in o) if isinstance(o, tuple) else o.copy()


class EnsemblePipeline:
    def __init__(self, models: Dict[str, Tuple[object, Dict]], classes: Sequence[str], output_size: Sequence[int],
                 device, batch: int, src_hw: Optional[Tuple[int, int]] = None, thickness: bool = False,
                 contour: bool = False):
        """models: {model_dir: (OCTSegmentationModel, cfg with 'input_size')} for every model_dir the
        requested classes need.  output_size: cv2 convention (width, height), as in the reference.
        Opt-in K-way probability averaging (north-star "ensemble averaging"; never the default -- the reference
        routes exactly one model per class): a LIST of K (model, cfg) folds for a model_dir makes every fold run
        to fp32 logits and ``octseg_fold_average_threshold`` produce the thresholded planes."""
        self.device = torch.device(device)
        self.classes = list(classes)
        for c in self.classes:
            if c not in MODELS_META:
                raise KeyError(c)
        self.Wo, self.Ho = int(output_size[0]), int(output_size[1])
        self.batch = batch
        self.thickness = thickness
        self.contour = contour      # also return the largest outer border per class (contour thickness, analysis.py:21-57)
        self.src_hw = tuple(src_hw) if src_hw is not None else (self.Ho, self.Wo)
        self.model_dirs = []
        for c in self.classes:
            d = MODELS_META[c]['model_dir']
            if d not in self.model_dirs:
                self.model_dirs.append(d)
        self.nets, self.fold_nets, self.sizes, self.fold_planes = {}, {}, {}, {}
        for d in self.model_dirs:
            folds = list(models[d]) if isinstance(models[d], list) else [models[d]]
            S = int(folds[0][1]['input_size'])
            self.sizes[d] = S
            if len(folds) == 1:
                self.fold_nets[d] = [folds[0][0].model.compiled(batch, S, S, self.device, 's2d', 'u8_nchw')]
            else:
                if any(int(cfg['input_size']) != S for _, cfg in folds):
                    raise ValueError(f'{d}: all folds must share input_size')
                self.fold_nets[d] = [m.model.compiled(batch, S, S, self.device, 's2d', 'f32_nchw') for m, _ in folds]
                with torch.cuda.device(self.device):
                    self.fold_planes[d] = torch.empty(self.fold_nets[d][0].out.shape, dtype=torch.uint8, device=self.device)
        self.nets = {d: self.fold_nets[d][0] for d in self.model_dirs}
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
        all_nets = [net for d in self.model_dirs for net in self.fold_nets[d]]
        self.macs_per_frame = sum(net.macs for net in all_nets) / batch
        # per net: its graph's launches + the fused resize / stem-pack launch (or the copy, for further folds)
        self.launches_per_batch = (sum(net.launches + 1 for net in all_nets) + len(self.fold_planes) + 1 + int(thickness) + int(contour))

    # ------------------------------------------------------------------ device-resident step
    def run_device(self, frames_dev: torch.Tensor):
        """frames_dev: uint8 CUDA (batch, Hs, Ws, 3) RGB.  Returns device (mask, label, counts, radii | None) and,
        when the pipeline was built with contour=True, a fifth element (sums, nverts, verts) of
        ``prepost.contour_largest``."""
        class_planes = {}
        cur = torch.cuda.current_stream(self.device)
        self.ev_in.record(cur)
        for d in self.model_dirs:
            st = self.streams[d]
            st.wait_event(self.ev_in)
            with torch.cuda.stream(st):
                outs = []
                for k, net in enumerate(self.fold_nets[d]):
                    if k == 0:   # resize + BGR + stem packing in one pass (bit-exact vs cv2, then exact uint8 -> bf16)
                        P.preprocess_s2d(frames_dev, self.sizes[d], out=net.x_s2d)
                    else:        # further folds of the same model dir see the same frames
                        net.x_s2d.copy_(self.fold_nets[d][0].x_s2d, non_blocking=True)
                    outs.append(net.run())                               # (batch, C, S, S) uint8 {0,1} | fp32 logits
                out = outs[0] if d not in self.fold_planes else P.fold_average_threshold(outs, out=self.fold_planes[d])
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
        if self.contour:
            return self.mask, self.label, self.counts, radii, P.contour_largest(self.mask)
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
            if self.contour:
                self._pinned.update(self._contour_pinned())
        return self._pinned

    def _contour_pinned(self):
        return {'c_sums': torch.empty(self.batch, 4, 4, dtype=torch.int64).pin_memory(),
                'c_nverts': torch.empty(self.batch, 4, dtype=torch.int32).pin_memory(),
                'c_verts': torch.empty(self.batch, 4, P.CONTOUR_CAP, 2, dtype=torch.int16).pin_memory()}

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
            mask, label, counts, radii, *extra = self.run_device(self.frames_dev)
            for key, t in zip(('c_sums', 'c_nverts', 'c_verts'), extra[0] if extra else ()):
                hb[key][:n].copy_(t[:n], non_blocking=True)
            hb['mask'][:n].copy_(mask[:n], non_blocking=True)
            hb['label'][:n].copy_(label[:n], non_blocking=True)
            hb['counts'][:n].copy_(counts[:n], non_blocking=True)
            if radii is not None:
                hb['radii'][:n].copy_(radii[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = [hb['mask'][:n].numpy(), hb['label'][:n].numpy(), hb['counts'][:n].numpy(),
               hb['radii'][:n].numpy() if radii is not None else None]
        if self.contour:
            out.append(tuple(hb[key][:n].numpy() for key in ('c_sums', 'c_nverts', 'c_verts')))
        if copy:
            out = [_copy(o) for o in out]
        return tuple(out)

    # ------------------------------------------------------------------ pipelined host-to-host stream
    def _stream_sets(self):
        if getattr(self, '_sets', None) is None:
            Hs, Ws = self.src_hw
            sets = []
            with torch.cuda.device(self.device):
                for _ in range(2):
                    sets.append({
                        'pin_frames': torch.empty(self.batch, Hs, Ws, 3, dtype=torch.uint8).pin_memory(),
                        'pin_mask': torch.empty(self.batch, self.Ho, self.Wo, 4, dtype=torch.uint8).pin_memory(),
                        'pin_label': torch.empty(self.batch, self.Ho, self.Wo, dtype=torch.uint8).pin_memory(),
                        'pin_counts': torch.empty(self.batch, 4, dtype=torch.int32).pin_memory(),
                        'pin_radii': torch.empty(self.batch, 4, 360, dtype=torch.int32).pin_memory(),
                        'dev_frames': torch.empty(self.batch, Hs, Ws, 3, dtype=torch.uint8, device=self.device),
                        'dev_mask': torch.empty_like(self.mask), 'dev_label': torch.empty_like(self.label),
                        'dev_counts': torch.empty_like(self.counts),
                        'dev_radii': torch.empty(self.batch, 4, 360, dtype=torch.int32, device=self.device),
                        **(self._contour_pinned() if self.contour else {}),
                        'ev_h2d': torch.cuda.Event(), 'ev_in_free': torch.cuda.Event(),
                        'ev_out_ready': torch.cuda.Event(), 'ev_d2h': torch.cuda.Event(), 'used': False,
                    })
                self._copy_streams = (torch.cuda.Stream(), torch.cuda.Stream())
            self._sets = sets
        return self._sets

    def stream_host(self, batches, copy: bool = True):
        """Pipelined form of ``run_host`` over an iterable of host batches (uint8 (n <= batch, Hs, Ws, 3)):
        the H2D copy of batch i+1 and the D2H copy of batch i-1 run on their own streams under the compute
        of batch i (two sets of pinned and device staging buffers).  Yields host (mask, label, counts,
        radii) per batch, in order.  With copy=False the arrays are views of pinned buffers that are valid only
        until the generator is advanced again: batch j is yielded during iteration j+1, and the next advance
        (iteration j+2) enqueues new D2H copies into the same staging set."""
        sets = self._stream_sets()
        h2d, d2h = self._copy_streams
        pending = []

        def collect(k, n, with_radii):
            S = sets[k]
            S['ev_d2h'].synchronize()
            out = [S['pin_mask'][:n].numpy(), S['pin_label'][:n].numpy(), S['pin_counts'][:n].numpy(),
                   S['pin_radii'][:n].numpy() if with_radii else None]
            if self.contour:
                out.append(tuple(S[key][:n].numpy() for key in ('c_sums', 'c_nverts', 'c_verts')))
            if copy:
                out = [_copy(o) for o in out]
            return tuple(out)

        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            for i, frames in enumerate(batches):
                k = i % 2
                S = sets[k]
                n = frames.shape[0]
                assert n <= self.batch and tuple(frames.shape[1:3]) == self.src_hw
                if S['used']:
                    S['ev_h2d'].synchronize()               # pinned input of batch i-2 has left the host
                S['pin_frames'][:n].copy_(torch.from_numpy(frames))
                with torch.cuda.stream(h2d):
                    if S['used']:
                        h2d.wait_event(S['ev_in_free'])     # batch i-2 has been moved out of the device staging buffer
                    S['dev_frames'][:n].copy_(S['pin_frames'][:n], non_blocking=True)
                    S['ev_h2d'].record(h2d)
                cur.wait_event(S['ev_h2d'])
                self.frames_dev[:n].copy_(S['dev_frames'][:n], non_blocking=True)
                if n < self.batch:
                    self.frames_dev[n:].zero_()
                S['ev_in_free'].record(cur)
                mask, label, counts, radii, *extra = self.run_device(self.frames_dev)
                if S['used']:
                    cur.wait_event(S['ev_d2h'])             # batch i-2's results have left the device staging buffers
                if extra:                                   # fresh tensors per call: kept alive in S until their D2H is done
                    S['dev_contour'] = extra[0]
                S['dev_mask'][:n].copy_(mask[:n], non_blocking=True)
                S['dev_label'][:n].copy_(label[:n], non_blocking=True)
                S['dev_counts'][:n].copy_(counts[:n], non_blocking=True)
                if radii is not None:
                    S['dev_radii'][:n].copy_(radii[:n], non_blocking=True)
                S['ev_out_ready'].record(cur)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(S['ev_out_ready'])
                    S['pin_mask'][:n].copy_(S['dev_mask'][:n], non_blocking=True)
                    S['pin_label'][:n].copy_(S['dev_label'][:n], non_blocking=True)
                    S['pin_counts'][:n].copy_(S['dev_counts'][:n], non_blocking=True)
                    if radii is not None:
                        S['pin_radii'][:n].copy_(S['dev_radii'][:n], non_blocking=True)
                    for key, t in zip(('c_sums', 'c_nverts', 'c_verts'), S.get('dev_contour', ())):
                        S[key][:n].copy_(t[:n], non_blocking=True)
                        t.record_stream(d2h)
                    S['ev_d2h'].record(d2h)
                S['used'] = True
                pending.append((k, n, radii is not None))
                if len(pending) == 2:
                    yield collect(*pending.pop(0))
            while pending:
                yield collect(*pending.pop(0))

