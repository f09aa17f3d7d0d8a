"""Per-network timing through the engine: graph replay time + per-launch breakdown.
Usage: python tools/bench_net.py KEY BATCH [SIZE]   (KEY in LM, VV, FC_LC)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if os.environ.get('OCTSEG_AB_LIB'):     # A/B builds of the library (tools/ab/*.so)
    from oct_segmentation_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ['OCTSEG_AB_LIB'])

from oct_segmentation_b200.model import OCTSegmentationModel
from oct_segmentation_b200.engine.network import CompiledNet
from oracle import synth


def main():
    key, N = sys.argv[1], int(sys.argv[2])
    cfg = synth.MODEL_CONFIGS[key]
    size = int(sys.argv[3]) if len(sys.argv) > 3 else cfg['input_size']
    torch.manual_seed(0)
    m = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                             in_channels=3, classes=cfg['classes'], encoder_weights=None)
    for mod in m.modules():  # non-degenerate BN stats without the oracle
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_var.fill_(1.0)
            mod.weight.data.fill_(0.5)
    net = CompiledNet(m.model, N, size, size, 'cuda', 'u8', 'u8_nchw', use_graph=True)
    net.x_nhwc.copy_(torch.randint(0, 255, net.x_nhwc.shape, dtype=torch.uint8, device='cuda'))
    for _ in range(3):
        net.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        net.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(net=key, batch=N, size=size, ms_per_batch=round(ms, 3), frames_per_s=round(N / ms * 1e3, 1),
                          tflops_algorithmic=round(2 * net.macs / ms / 1e9, 1), launches=net.launches,
                          act_GB=round(net.builder.act_bytes / 1e9, 2), arena_GB=round(net.builder.arena_bytes / 1e9, 2))), flush=True)
    # per-op breakdown (eager, one event pair per op)
    b = net.builder
    times = []
    for name, op in zip(b.op_names, b.ops):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        op()
        e.record()
        times.append((name, s, e))
    torch.cuda.synchronize()
    rows = sorted(((n, s.elapsed_time(e)) for n, s, e in times), key=lambda r: -r[1])
    tot = sum(t for _, t in rows)
    print(f'eager sum {tot:.3f} ms; top ops:')
    for n, t in rows[:25]:
        print(f'  {t:8.4f} ms  {100 * t / tot:5.1f}%  {n}')
    json.dump({n: t for n, t in rows}, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', f'ops_{key}.json'), 'w'))
    groups = {}
    for n, t in rows:
        k = 'stem' if ('conv1' == n.split('.')[-1] and n.count('.') == 1) or 'stem' in n else (
            'maxpool' if 'maxpool' in n else 'dw' if 'depthwise' in n else 'se+project' if '_project' in n else
            'decoder' if n.startswith('decoder') else 'head' if n.startswith('segmentation') else 'encoder-conv')
        groups[k] = groups.get(k, 0) + t
    print('by group:', {k: round(v, 3) for k, v in sorted(groups.items(), key=lambda r: -r[1])})


if __name__ == '__main__':
    main()
