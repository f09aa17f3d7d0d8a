"""Host side of the pre/post-processing kernels (csrc/prepost.cu): index/coefficient tables built
with the exact arithmetic cv2 uses, and thin wrappers around the C-ABI calls.

  preprocess   <- preprocessing_img, /root/reference/src/data/utils.py:159-166
  postprocess  <- /root/reference/src/predict.py:92-100 (+ MODELS_META :23-28),
                  src/data/utils.py:231-233 (priority merge), src/app/tools/analysis.py:199 (count)
  quantities   <- src/app/tools/analysis.py:155,189,199-200 (ratio, presence, area)
  thickness    <- src/app/tools/analysis.py:60-130 (radial scan)
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CLASS_NAMES = ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']


def linear_tables(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """cv2.resize INTER_LINEAR tables for uint8: source index floor(fx) and 11-bit fixed-point
    weights, fx = (float)((dx + 0.5) * (double)src/dst - 0.5); weights rounded half-to-even like
    saturate_cast<short>.  Horizontal border handling (sx<0 / sx>=src-1 -> weight (2048, 0)) is
    applied by `horizontal_tables`; the vertical pass clamps rows instead (as cv2 does)."""
    scale = float(src) / float(dst)
    fx = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fr = (fx - sx.astype(np.float32)).astype(np.float32)
    a0 = np.rint((np.float32(1.0) - fr) * np.float32(2048.0)).astype(np.int16)
    a1 = np.rint(fr * np.float32(2048.0)).astype(np.int16)
    return sx, np.stack([a0, a1], axis=1)


def horizontal_tables(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    sx, a = linear_tables(src, dst)
    edge = (sx < 0) | (sx >= src - 1)
    a = a.copy()
    a[edge] = (2048, 0)
    return np.clip(sx, 0, src - 1).astype(np.int32), a


def nearest_table(src: int, dst: int) -> np.ndarray:
    """cv2.resize INTER_NEAREST: min(floor(dx * (1 / ((double)dst/src))), src - 1)."""
    inv = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * inv).astype(np.int64), src - 1).astype(np.int32)


@lru_cache(maxsize=64)
def _resize_luts(Hs: int, Ws: int, S: int, device: str):
    xo, xa = horizontal_tables(Ws, S)
    yo, yb = linear_tables(Hs, S)
    dev = torch.device(device)
    return (torch.from_numpy(xo).to(dev), torch.from_numpy(xa.reshape(-1)).to(dev),
            torch.from_numpy(yo).to(dev), torch.from_numpy(yb.reshape(-1)).to(dev))


@lru_cache(maxsize=64)
def _nearest_lut(S: int, Ho: int, Wo: int, device: str) -> torch.Tensor:
    lut = np.concatenate([nearest_table(S, Ho), nearest_table(S, Wo)])
    return torch.from_numpy(lut).to(torch.device(device))


@lru_cache(maxsize=4)
def _cos_sin(device: str) -> torch.Tensor:
    cs = [math.cos(math.radians(a)) for a in range(360)] + [math.sin(math.radians(a)) for a in range(360)]
    return torch.tensor(cs, dtype=torch.float64, device=torch.device(device))


def preprocess(frames_rgb: torch.Tensor, S: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frames_rgb: uint8 CUDA tensor (N, Hs, Ws, 3) -> uint8 (N, S, S, 3) BGR, bit-exact vs
    cv2.resize(cv2.cvtColor(img, COLOR_RGB2BGR), (S, S)).  Grayscale frames, (N, Hs, Ws) or (N, Hs, Ws, 1), are
    replicated into the three channels (== the same call on np.repeat(frame[..., None], 3, -1))."""
    assert frames_rgb.is_cuda and frames_rgb.dtype == torch.uint8 and frames_rgb.is_contiguous()
    gray = frames_rgb.dim() == 3 or frames_rgb.shape[3] == 1
    assert gray or (frames_rgb.dim() == 4 and frames_rgb.shape[3] == 3), 'frames must have 1 or 3 channels'
    N, Hs, Ws = frames_rgb.shape[:3]
    if out is None:
        out = torch.empty(N, S, S, 3, dtype=torch.uint8, device=frames_rgb.device)
    lib = _lib.load()
    area2x = int(Hs == 2 * S and Ws == 2 * S)
    xo, xa, yo, yb = _resize_luts(Hs, Ws, S, str(frames_rgb.device))
    fn = lib.octseg_preprocess_resize_gray if gray else lib.octseg_preprocess_resize_bgr
    with torch.cuda.device(frames_rgb.device):
        _lib.check(fn(frames_rgb.data_ptr(), N, Hs, Ws, out.data_ptr(), S, xo.data_ptr(), xa.data_ptr(), yo.data_ptr(),
                      yb.data_ptr(), area2x, _lib.stream_ptr()), 'preprocess_resize')
    return out


def preprocess_s2d(frames_rgb: torch.Tensor, S: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`preprocess` written straight into the network stem's input format: bf16 CUDA (N, S/2, S/2, 16), channel
    (dy*2 + dx)*3 + c = BGR pixel (2y+dy, 2x+dx), channels 12..15 zero -- what the stem-pack launch would make of
    `preprocess(frames, S)`; one pass over the frame instead of two plus a uint8 round trip.  `unpack_s2d` inverts it."""
    assert frames_rgb.is_cuda and frames_rgb.dtype == torch.uint8 and frames_rgb.is_contiguous() and S % 2 == 0
    gray = frames_rgb.dim() == 3 or frames_rgb.shape[3] == 1
    assert gray or (frames_rgb.dim() == 4 and frames_rgb.shape[3] == 3), 'frames must have 1 or 3 channels'
    N, Hs, Ws = frames_rgb.shape[:3]
    if out is None:
        out = torch.empty(N, S // 2, S // 2, 16, dtype=torch.bfloat16, device=frames_rgb.device)
    assert out.is_contiguous() and tuple(out.shape) == (N, S // 2, S // 2, 16) and out.dtype == torch.bfloat16
    lib = _lib.load()
    area2x = int(Hs == 2 * S and Ws == 2 * S)
    xo, xa, yo, yb = _resize_luts(Hs, Ws, S, str(frames_rgb.device))
    with torch.cuda.device(frames_rgb.device):
        _lib.check(lib.octseg_preprocess_resize_s2d(frames_rgb.data_ptr(), 1 if gray else 3, N, Hs, Ws, out.data_ptr(), S,
                                                    xo.data_ptr(), xa.data_ptr(), yo.data_ptr(), yb.data_ptr(), area2x,
                                                    _lib.stream_ptr()), 'preprocess_resize_s2d')
    return out


def unpack_s2d(x2: torch.Tensor) -> torch.Tensor:
    """(N, H/2, W/2, 16) stem-packed tensor -> (N, H, W, 3) frame (same dtype): inverse of the space-to-depth packing."""
    N, H2, W2, _ = x2.shape
    v = x2[..., :12].reshape(N, H2, W2, 2, 2, 3)                 # (n, y, x, dy, dx, c)
    return v.permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * H2, 2 * W2, 3)


def postprocess(planes: Dict[int, torch.Tensor], order: Sequence[int], Ho: int, Wo: int, N: int, device,
                mask: Optional[torch.Tensor] = None, label: Optional[torch.Tensor] = None,
                counts: Optional[torch.Tensor] = None):
    """planes[c]: uint8 CUDA (N, S_c, S_c) {0,1} plane feeding mask channel c (= class id - 1); may be a
    channel slice of a network's (N, C, S, S) output (only the frame stride may be non-dense).
    order: class channel indices in cfg.classes order (later wins in the label map).
    Returns (mask uint8 [N,Ho,Wo,4], label uint8 [N,Ho,Wo], counts int32 [N,4])."""
    dev = torch.device(device)
    lib = _lib.load()
    if mask is None:
        mask = torch.empty(N, Ho, Wo, 4, dtype=torch.uint8, device=dev)
    if label is None:
        label = torch.empty(N, Ho, Wo, dtype=torch.uint8, device=dev)
    if counts is None:
        counts = torch.zeros(N, 4, dtype=torch.int32, device=dev)
    else:
        counts.zero_()
    chan = (C.c_void_p * 4)()
    luts = (C.c_void_p * 4)()
    sizes = (C.c_int32 * 4)()
    strides = (C.c_int64 * 4)()
    keep = []
    for c in range(4):
        p = planes.get(c)
        if p is None:
            chan[c], luts[c], sizes[c] = None, None, 0
            continue
        S = p.shape[1]
        assert p.is_cuda and p.dtype == torch.uint8 and p.shape[0] == N and p.shape[2] == S
        assert p.stride(2) == 1 and p.stride(1) == S, 'plane rows must be dense'
        strides[c] = p.stride(0) if N > 1 else S * S
        lut = _nearest_lut(S, Ho, Wo, str(dev))
        keep.append(lut)
        chan[c], luts[c], sizes[c] = p.data_ptr(), lut.data_ptr(), S
    ordr = (C.c_int32 * 4)(*(list(order) + [0] * (4 - len(order))))
    with torch.cuda.device(dev):
        _lib.check(lib.octseg_postprocess(chan, sizes, strides, luts, ordr, len(order), N, Ho, Wo, mask.data_ptr(),
                                          label.data_ptr(), counts.data_ptr(), _lib.stream_ptr()), 'postprocess')
    return mask, label, counts


def radial_thickness(mask: torch.Tensor) -> torch.Tensor:
    """mask: uint8 CUDA (N, H, W, 4), non-zero = object -> int32 (N, 4, 360) per-ray radii."""
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
    N, H, W, _ = mask.shape
    radii = torch.empty(N, 4, 360, dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.check(_lib.load().octseg_radial_thickness(mask.data_ptr(), N, H, W, _cos_sin(str(mask.device)).data_ptr(),
                                                       radii.data_ptr(), _lib.stream_ptr()), 'radial_thickness')
    return radii


CLASS_COLORS_RGB = {'Lumen': (228, 30, 199), 'Fibrous cap': (123, 171, 226), 'Lipid core': (125, 227, 127),
                    'Vasa vasorum': (208, 2, 27)}   # src/data/utils.py:15-36 (== model.CLASS_COLORS_RGB, test-checked)


@lru_cache(maxsize=1)
def overlay_alpha_tables() -> Tuple[np.ndarray, int]:
    """Alpha bytes of save_results' two pastes (src/data/utils.py:223-230, src/models/smp/utils.py:209-211):
    ``uint8(blur * 64 * 0.85 * 255)`` for blur = k/256, k = 0..256, and ``uint8(1 * 255 * 0.85 * 255)``; the
    reference's float -> uint8 cast of out-of-range values wraps (truncate, mod 256), kept as is."""
    k = np.arange(257, dtype=np.float64)
    fill = np.trunc(((k / 256.0) * 64) * 0.85 * 255).astype(np.int64) & 0xFF
    rim = int(math.trunc(1.0 * 255 * 0.85 * 255)) & 0xFF
    return fill.astype(np.uint8), rim


def overlay(frames_rgb: torch.Tensor, mask: torch.Tensor, order: Sequence[int],
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """<name>_overlay.png of save_results for a batch.  frames_rgb: uint8 CUDA (N, H, W, 3) at the output size;
    mask: uint8 CUDA (N, H, W, 4) as written by ``postprocess``; order: class channels in cfg.classes order."""
    assert frames_rgb.is_cuda and frames_rgb.dtype == torch.uint8 and frames_rgb.is_contiguous()
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
    N, H, W, _ = frames_rgb.shape
    assert tuple(mask.shape) == (N, H, W, 4)
    if out is None:
        out = torch.empty_like(frames_rgb)
    fill, rim = overlay_alpha_tables()
    colors = np.array([CLASS_COLORS_RGB[n] for n in CLASS_NAMES], np.uint8)
    ordr = (C.c_int32 * 4)(*(list(order) + [0] * (4 - len(order))))
    with torch.cuda.device(frames_rgb.device):
        _lib.check(_lib.load().octseg_overlay(frames_rgb.data_ptr(), mask.data_ptr(), out.data_ptr(), N, H, W, ordr,
                                              len(order), colors.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              fill.ctypes.data_as(C.POINTER(C.c_uint8)), rim, _lib.stream_ptr()),
                   'overlay')
    return out


def fold_average_threshold(logits: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K-way probability averaging (opt-in; the reference routes one model per class): logits = K same-shaped fp32
    CUDA tensors -> uint8 {0,1} tensor of that shape, (mean_k sigmoid(logit_k)) > 0.5.  K = 1 is model.py:195."""
    K = len(logits)
    x0 = logits[0]
    for x in logits:
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape == x0.shape
    if out is None:
        out = torch.empty(x0.shape, dtype=torch.uint8, device=x0.device)
    assert out.is_contiguous() and out.dtype == torch.uint8 and out.shape == x0.shape
    ptrs = (C.c_void_p * K)(*[x.data_ptr() for x in logits])
    with torch.cuda.device(x0.device):
        _lib.check(_lib.load().octseg_fold_average_threshold(ptrs, K, x0.numel(), out.data_ptr(), _lib.stream_ptr()),
                   'fold_average_threshold')
    return out


CONTOUR_CAP = 16384   # kept border points per (frame, class) the device buffer holds


def contour_fits(H: int, W: int) -> bool:
    """octseg_contour_largest keeps one class's bit plane in shared memory (include/octseg.h): 1000 x 1000 needs
    132 KB, the limit of 226 KB is reached near 1330 x 1330."""
    pitch = (W + 1 + 31) // 32 + 1
    return ((H + 2) * pitch + 1) * 4 <= 226 * 1024


def contour_largest(mask: torch.Tensor, cap: int = CONTOUR_CAP):
    """mask: uint8 CUDA (N, H, W, 4), non-zero = object.  Per (frame, class) the largest outer border in
    cv2's CHAIN_APPROX_SIMPLE form (== max(findContours(RETR_EXTERNAL), key=contourArea), analysis.py:25-35):
    returns device (sums int64 (N, 4, 4) = a00, a10, a01, start index | -1; nverts int32 (N, 4); verts int16
    (N, 4, cap, 2) as x, y)."""
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
    N, H, W, _ = mask.shape
    sums = torch.empty(N, 4, 4, dtype=torch.int64, device=mask.device)
    nverts = torch.empty(N, 4, dtype=torch.int32, device=mask.device)
    verts = torch.empty(N, 4, cap, 2, dtype=torch.int16, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.check(_lib.load().octseg_contour_largest(mask.data_ptr(), N, H, W, sums.data_ptr(), nverts.data_ptr(),
                                                      verts.data_ptr(), cap, _lib.stream_ptr()), 'contour_largest')
    return sums, nverts, verts


def thickness_from_contour(sums: np.ndarray, nverts: int, verts: np.ndarray) -> Dict[str, float]:
    """Host finish of calculate_thickness_contour (src/app/tools/analysis.py:37-57) from the device result of one
    (frame, class): centroid from cv2's polygon moments (m00 = a00 * +-0.5, m10 = a10 * +-(1/6), same double
    operations as contourMoments), int-truncated; distances centroid -> kept border points; median / min / max."""
    a00, a10, a01 = int(sums[0]), int(sums[1]), int(sums[2])
    if nverts == 0 or a00 == 0:
        return {'median': 0, 'min': 0, 'max': 0}
    if nverts > verts.shape[0]:
        raise RuntimeError(f'contour of {nverts} points exceeds the device buffer ({verts.shape[0]}); raise `cap`')
    sign = 1.0 if a00 > 0 else -1.0
    m00 = float(a00) * (sign * 0.5)
    m10 = float(a10) * (sign * 0.16666666666666666666666666666667)
    m01 = float(a01) * (sign * 0.16666666666666666666666666666667)
    cx, cy = int(m10 / m00), int(m01 / m00)
    pts = verts[:nverts].astype(np.int64)
    d = np.sqrt((pts[:, 0] - cx) ** 2 + (pts[:, 1] - cy) ** 2)
    return {'median': float(np.median(d)), 'min': float(np.min(d)), 'max': float(np.max(d))}


def dicom_ratio(h: int) -> int:
    return int(h * 150 // 1000)


def quantities_from_counts(counts: np.ndarray, H: int, W: int, ratio: int, radii: Optional[np.ndarray] = None,
                           contours=None, masks: Optional[np.ndarray] = None) -> List[Dict]:
    """Per-frame, per-class table from the device reductions (host side, cheap):
    present <=> 0 < nnz < H*W (== np.unique(ch).shape[0] == 2, analysis.py:189);
    area = sqrt(nnz // ratio) (analysis.py:199-200); radial thickness median/min over rays that hit;
    with ``contours`` (host (sums, nverts, verts) of ``contour_largest``) also the reference table's
    thickness_mean = contour median / ratio and thickness_min = contour min / ratio (analysis.py:202-207), under the
    keys contour_thickness_mean / contour_thickness_min, for present classes.  A border longer than the device
    buffer (CONTOUR_CAP points; only ragged, noise-like masks get there) is walked again on the GPU with a buffer of
    the reported length when the host copy ``masks`` (uint8 (N, H, W, 4)) is given; without it that raises."""
    rows = []
    for n in range(counts.shape[0]):
        row = {}
        for c, name in enumerate(CLASS_NAMES):
            nnz = int(counts[n, c])
            q = {'nnz': nnz, 'present': 0 < nnz < H * W}
            if q['present']:
                q['area'] = pow(nnz // ratio, 0.5)
            if radii is not None:
                r = radii[n, c]
                r = r[r > 0]
                q['thickness_median'] = float(np.median(r)) if r.size else 0
                q['thickness_min'] = int(r.min()) if r.size else 0
                q['thickness_max'] = int(r.max()) if r.size else 0
            if contours is not None and q['present']:
                sums_nc, nv, verts_nc = contours[0][n, c], int(contours[1][n, c]), contours[2][n, c]
                if nv > verts_nc.shape[0] and masks is not None:
                    redo = contour_largest(torch.from_numpy(np.ascontiguousarray(masks[n:n + 1])).cuda(), cap=-(-nv // 1024) * 1024)
                    sums_nc, nv, verts_nc = (redo[0][0, c].cpu().numpy(), int(redo[1][0, c]), redo[2][0, c].cpu().numpy())
                t = thickness_from_contour(sums_nc, nv, verts_nc)
                q['contour_thickness_mean'] = t['median'] / ratio
                q['contour_thickness_min'] = t['min'] / ratio
            row[name] = q
        rows.append(row)
    return rows


def objects_table(rows: Sequence[Dict], image_names: Optional[Sequence[str]] = None) -> Dict[str, Dict[str, list]]:
    """The `objects` dict of get_analysis (src/app/tools/analysis.py:139-154, loop :185-213) from the per-frame rows
    of ``quantities_from_counts`` (frames in slice order, all ranks gathered): per class the slices where it is
    present, their object ids (consecutive slices share an id, a gap starts the next: :191-198), area (:199-200)
    and, when the rows carry contour thickness, thickness_mean / thickness_min (:202-207).  Host side by design:
    a sequential scan over N x 4 booleans after the gather."""
    objects = {name: {'object_id': [], 'slice': [], 'area': [], 'thickness_mean': [], 'thickness_min': [], 'img_name': []}
               for name in CLASS_NAMES}
    for idx, row in enumerate(rows):
        for name in CLASS_NAMES:
            q = row[name]
            if not q['present']:
                continue
            obj = objects[name]
            if len(obj['object_id']) == 0:
                obj['object_id'].append(0)
            elif idx == obj['slice'][-1] + 1:
                obj['object_id'].append(obj['object_id'][-1])
            else:
                obj['object_id'].append(obj['object_id'][-1] + 1)
            obj['slice'].append(idx)
            obj['area'].append(q['area'])
            if 'contour_thickness_mean' in q:
                obj['thickness_mean'].append(q['contour_thickness_mean'])
                obj['thickness_min'].append(q['contour_thickness_min'])
            if image_names is not None:
                obj['img_name'].append(image_names[idx])
    return objects
