"""One small invocation of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke())."""
from __future__ import annotations

import numpy as np
import torch


def run_smoke() -> None:
    from oracle import synth
    from oracle import prepost_ref as R
    from . import prepost as P
    from .model import OCTSegmentationModel

    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda:0')
    key, S, Ho = 'VV', 128, 250
    ref = synth.make_model(key, calib_size=S, calib_frames=2)
    cfg = synth.MODEL_CONFIGS[key]
    ours = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                in_channels=3, classes=cfg['classes'], encoder_weights=None)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.to(dev).eval()
    frames = synth.synthetic_frames(500, 2, Ho)                                   # RGB uint8
    # pre-processing kernel vs cv2 (bit-exact)
    pre = P.preprocess(torch.from_numpy(frames).to(dev), S)
    want_pre = np.stack([R.preprocess_frame(f, S) for f in frames])
    assert np.array_equal(pre.cpu().numpy(), want_pre), 'preprocess kernel differs from cv2'
    # network vs fp32 oracle
    x = pre.permute(0, 3, 1, 2).float()
    with torch.no_grad():
        want = ref.to(dev).model(x)
        got = ours.model(x)
    err = ((got - want).norm() / want.norm()).item()
    assert err < 0.12, f'logits rel-L2 {err:.3e} vs fp32 oracle'
    # post-processing kernel vs numpy oracle (bit-exact)
    plane = (got[:, 0] > 0).to(torch.uint8).contiguous()
    mask, label, counts = P.postprocess({3: plane}, [3], Ho, Ho, 2, dev)
    idx = R.nearest_index(S, Ho)
    want_mask = plane.cpu().numpy()[:, idx][:, :, idx]
    assert np.array_equal(mask[..., 3].cpu().numpy(), want_mask), 'postprocess kernel differs from the oracle'
    assert np.array_equal(counts[:, 3].cpu().numpy(), want_mask.reshape(2, -1).sum(1))
    # overlay + contour kernels vs the numpy / cv2 oracle on that mask (bit-exact)
    over = P.overlay(torch.from_numpy(frames).to(dev), mask, [3])
    m_host = mask.cpu().numpy()
    for n in range(2):
        assert np.array_equal(over[n].cpu().numpy(), R.overlay(frames[n], m_host[n], ['Vasa vasorum'])), 'overlay kernel differs'
    sums, nverts, verts = (t.cpu().numpy() for t in P.contour_largest(mask))
    for n in range(2):
        got_t = P.thickness_from_contour(sums[n, 3], int(nverts[n, 3]), verts[n, 3])
        assert got_t == R.thickness_contour(np.ascontiguousarray(m_host[n, :, :, 3])), 'contour kernel differs from cv2'
    torch.cuda.synchronize()
    print(f'smoke ok: {key} logits rel-L2 vs fp32 oracle {err:.3e}; pre/post/overlay/contour kernels bit-exact; '
          f'{ours.model.compiled(2, S, S, dev, "f32", "f32_nchw").launches} launches')
