"""Where does the engine's distance to the fp32 oracle come from?  (round-1 verdict, item 1b)

For a fitted checkpoint of one network, compares on the same unseen frames
  * the fp32 oracle (oracle/smp_ref.py),
  * the B200 engine (C-ABI kernels, bf16 storage),
  * a torch-op emulation of the engine's storage precision (tests/cpu_builder.Bf16Builder semantics: folded
    weights rounded to bf16, every activation rounded to bf16 when written, fp32 accumulation),
  * variants of that emulation with one class of rounding removed (weights / activations of the head,
    of the last decoder block, of the decoder, of the encoder; the `.up` partial of the split decoder convs),
and prints logits rel-L2, Dice, differing pixels by direction and the mean logit shift for each.
engine ~ emulation  => the kernels add nothing beyond bf16 storage; the variants show which rounding matters.

Usage (GPU box): python tools/diag_parity.py KEY [SIZE] [FRAMES] [STEPS]      output also -> gpurun_out/diag_parity_KEY.txt
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('CUBLAS_WORKSPACE_CONFIG', ':4096:8')
import torch

from oct_segmentation_b200.engine.lower import ENCODER_LOWERING, lower_decoder_and_head
from oct_segmentation_b200.model import OCTSegmentationModel
from oracle import synth
from tests.cpu_builder import Bf16Builder, CpuBuilder, _r16


class VarBuilder(Bf16Builder):
    """w_fp32 / a_fp32: regexes of op names whose weights / output activations stay fp32; w_hilo: weights as a
    two-term bf16 sum; split_up: emulate the engine's `.up` split of narrow decoder convs (engine/builder.py)."""

    def __init__(self, N, device, w_fp32=None, a_fp32=None, split_up=True, up_fp32=False, w_hilo=None):
        super().__init__(N, device)
        self.w_fp32, self.a_fp32, self.split_up, self.up_fp32, self.w_hilo = w_fp32, a_fp32, split_up, up_fp32, w_hilo

    def conv(self, srcs, w, b, **k):
        name = k['name']
        wf = w.detach().float()
        if self.w_fp32 and re.search(self.w_fp32, name):
            wq = wf
        elif self.w_hilo and re.search(self.w_hilo, name):
            hi = _r16(wf)
            wq = hi + _r16(wf - hi)
        else:
            wq = _r16(wf)
        bf16_out = k.get('out_mode', 'bf16_nhwc') == 'bf16_nhwc'
        if (self.split_up and bf16_out and not k.get('transposed') and len(srcs) > 1 and srcs[0][1]
                and not any(up for _, up in srcs[1:]) and k.get('groups', 1) == 1 and k.get('res') is None
                and w.shape[0] <= 64 and srcs[1][0].W >= 112):
            cup = srcs[0][0].C
            kk = dict(k)
            kk.update(name=name + '.up', act='none')
            part = CpuBuilder.conv(self, [srcs[0]], wq[:, :cup], None, **kk)
            if not self.up_fp32:
                part.t = _r16(part.t)
            kk = dict(k)
            kk.update(res=part, res_mode='before_act')
            a = CpuBuilder.conv(self, list(srcs[1:]), wq[:, cup:], b, **kk)
        else:
            a = CpuBuilder.conv(self, srcs, wq, b, **k)
        if a is not None and not (self.a_fp32 and re.search(self.a_fp32, name)):
            a.t = _r16(a.t)
        return a


VARIANTS = {
    'emulation: bf16 storage (engine semantics)': {},
    '  .up partial kept fp32': dict(up_fp32=True),
    '  head weights fp32': dict(w_fp32='segmentation_head'),
    '  head + last block weights hi/lo bf16': dict(w_hilo='segmentation_head|x_0_4|decoder.blocks.4'),
    '  decoder weights fp32': dict(w_fp32='decoder|segmentation_head'),
    '  encoder weights fp32': dict(w_fp32='encoder'),
    '  all weights fp32': dict(w_fp32='.'),
    '  decoder activations fp32': dict(a_fp32='decoder', up_fp32=True),
    '  encoder activations fp32': dict(a_fp32='encoder'),
    '  all activations fp32': dict(a_fp32='.', up_fp32=True),
}


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else 'LM'
    cfg = synth.MODEL_CONFIGS[key]
    size = int(sys.argv[2]) if len(sys.argv) > 2 else cfg['input_size']
    nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 300
    batch = 2 if size <= 512 else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda')
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)

    ref = synth.make_model(key, calib_size=128, calib_frames=2)
    loss = synth.fit_model(ref, dev, steps=steps, size=size, batch=batch, log=say)
    ours = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                in_channels=3, classes=cfg['classes'], encoder_weights=None)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref, ours = ref.to(dev).eval(), ours.to(dev).eval()
    frames = synth.synthetic_frames(5000, nfr, size)[..., ::-1].copy()
    x = torch.from_numpy(frames).to(dev).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        want = torch.cat([ref.model(x[i:i + 1]) for i in range(nfr)])
        got = ours.model(x).clone()
    say(f'{key} @ {size}: fit loss {loss:.4f} after {steps} steps; oracle logits mean {want.mean().item():.3f} std '
        f'{want.std().item():.3f}, positive {100 * (want > 0).float().mean().item():.2f}%, |logit|<0.25 on '
        f'{100 * (want.abs() < 0.25).float().mean().item():.3f}% of pixels')

    def report(tag, y, base=want):
        rel = ((y - base).norm() / base.norm()).item()
        a, b = base > 0, y > 0
        d = 2.0 * (a & b).sum().item() / max(a.sum().item() + b.sum().item(), 1)
        say(f'{tag:48s} rel-L2 {rel:.2e}  dice {d:.5f}  differing {int((a != b).sum())} (lost {int((a & ~b).sum())}, '
            f'gained {int((~a & b).sum())})  mean shift {(y - base).mean().item():+.5f}')

    def emulate(**kw):
        b = VarBuilder(nfr, dev, **kw)
        with torch.no_grad():
            feats = ENCODER_LOWERING[ours.model.encoder.kind](b, ours.model.encoder, x, 'f32', None)
            out = torch.zeros_like(want)
            lower_decoder_and_head(b, ours.model, feats, out, 'f32_nchw')
        return out

    say('--- against the fp32 oracle')
    report('B200 engine', got)
    emu = None
    for tag, kw in VARIANTS.items():
        y = emulate(**kw)
        if emu is None:
            emu = y
        report(tag, y)
    say('--- engine against its own storage-precision emulation')
    report('B200 engine vs emulation', got, emu)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    open(os.path.join(ROOT, 'gpurun_out', f'diag_parity_{key}.txt'), 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
