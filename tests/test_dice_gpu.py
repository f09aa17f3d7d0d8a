"""GPU: mask-level parity on a FITTED synthetic checkpoint.  The fp32 oracle network is fitted for a
few dozen Adam steps on phantom targets (on the GPU, in this test), its weights are loaded into the
B200 engine, and both implementations segment unseen synthetic frames.

Bars (BASELINE.json north_star): per-class Dice >= 0.999 between the two implementations' masks;
pixel areas (non-zero counts) equal wherever the masks are equal; logits rel-L2 <= 3e-2 (bf16
storage vs fp32 -- a fitted network no longer amplifies rounding noise the way the BN-calibrated
random one does)."""
import numpy as np
import pytest
import torch

from oct_segmentation_b200.model import OCTSegmentationModel
from oracle import synth

pytestmark = pytest.mark.gpu


def dice(a, b):
    inter = (a & b).sum().item()
    return 2.0 * inter / max(a.sum().item() + b.sum().item(), 1)


@pytest.mark.parametrize('key,size,steps', [('VV', 256, 150), ('LM', 256, 200), ('FC_LC', 256, 400)])
def test_fitted_checkpoint_dice(key, size, steps):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = synth.make_model(key, calib_size=128, calib_frames=2)
    loss = synth.fit_model(ref, 'cuda', steps=steps, size=128, batch=8, target_loss=0.005, max_steps=1500)
    cfg = synth.MODEL_CONFIGS[key]
    ours = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                in_channels=3, classes=cfg['classes'], encoder_weights=None)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.cuda().eval()
    frames = synth.synthetic_frames(5000, 8, size)[..., ::-1].copy()
    x = torch.from_numpy(frames).cuda().permute(0, 3, 1, 2).float()
    with torch.no_grad():
        want = ref.model(x)
        got = ours.model(x)
    rel = ((got - want).norm() / want.norm()).item()
    band = (want.abs() < 0.25).float().mean().item()
    report = [f'fit loss {loss:.4f}', f'logits rel-L2 {rel:.2e}', f'|logit|<0.25 on {100 * band:.2f}% of pixels']
    ok = True
    for c, name in enumerate(cfg['classes']):
        a, b = want[:, c] > 0, got[:, c] > 0
        d = dice(a, b)
        frac = a.float().mean().item()
        diff = (a != b).sum().item()
        report.append(f'{name}: dice {d:.5f}, positive {100 * frac:.1f}%, {diff} differing px, '
                      f'areas {int(a.sum())} vs {int(b.sum())}')
        assert frac > 0.005, f'{name}: fitted oracle predicts an empty mask ({frac})'
        ok &= d >= 0.999
        per_frame_equal = [(a[n] == b[n]).all().item() for n in range(a.shape[0])]
        for n, eq in enumerate(per_frame_equal):
            if eq:
                assert int(a[n].sum()) == int(b[n].sum())
    print(f'\n{key}: ' + '; '.join(report))
    assert rel <= 3e-2, report
    assert ok, report
