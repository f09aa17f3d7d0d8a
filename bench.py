#!/usr/bin/env python
"""bench.py — frames/s of the hybrid-ensemble inference hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = one pass of the hot path over one batch of B synthetic 512x512 RGB frames per GPU:
bilinear resize+BGR -> LM (U-Net++/resnet101 @512) + FC_LC (LinkNet/efficientnet-b7 @896) +
VV (U-Net/regnetx_064 @896) -> threshold -> nearest resize to 1000x1000 + class routing +
label map + per-class pixel counts (+ radial thickness).  One process per GPU (torchrun for
N > 1), frames sharded by rank, no data-path collective; one final gather of the per-frame
counts table.

Output: ONE JSON line on rank 0 (contract in the task statement): `value` = device-resident
throughput, `e2e` = same metric through EnsemblePipeline.stream_host (host frames in,
host masks/labels/counts out, copies inside the timed region), `roofline` for the dominant
kernel (conv_tc_kernel, tensor-bound), `cpu_baseline` = the CPU oracle port on host cores.

`--impl reference` times the reference's own CPU implementation of the path restated in
oracle/ (the original cannot be installed offline: smp/timm/efficientnet_pytorch/lightning/
hydra are absent) on all host threads, batch 1 per call as src/predict.py does.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

CLASSES = ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']
SRC = 512
OUT_SIZE = [1000, 1000]
WORKLOAD = ('hybrid ensemble LM(UnetPlusPlus/resnet101@512)+FC_LC(LinkNet/efficientnet-b7@896)+'
            'VV(Unet/timm-regnetx_064@896), routing+label map+pixel counts+radial thickness at 1000x1000, '
            'synthetic 512x512 RGB frames')
METRIC = 'frames/s, LM+FC_LC+VV ensemble inference'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1409.7), d.get('hbm_gbs', 6533.8), 'measured (MEASURED_PEAKS.json, sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_oracle_frames_per_s(n_frames: int, warmup: int):
    """The reference path restated in oracle/ (batch 1 per call, FC_LC run once per class exactly like
    src/predict.py:70-100), timed on the host cores.  Returns (frames/s, seconds per frame list)."""
    from oracle import model_ref, synth
    torch.set_num_threads(os.cpu_count() or 1)
    models = {}
    for key in ('LM', 'FC_LC', 'VV'):
        m = synth.make_model(key, calib_size=128, calib_frames=1)
        models[key] = (m, synth.MODEL_CONFIGS[key])
    from PIL import Image
    times = []
    for i in range(warmup + n_frames):
        rgb = synth.synthetic_frame(10_000 + i, SRC)
        img = Image.fromarray(rgb).resize(tuple(OUT_SIZE))          # data_processing (PIL bicubic)
        mask = np.zeros((OUT_SIZE[0], OUT_SIZE[1], 4))
        t0 = time.perf_counter()
        model_ref.segment_with_models([img], [mask], OUT_SIZE, CLASSES, models, 'cpu')
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return len(times) / sum(times), times


def reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fps, times = cpu_oracle_frames_per_s(args.steps, args.warmup)
    ms = 1e3 * float(np.mean(times))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_step': 1, 'note': 'reference-as-written: batch 1 per call, each class '
                   'loops its model (FC_LC runs twice), cv2 pre/post on CPU; runs on rank 0 host cores only'},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} frames (1 frame per step) through the oracle restatement of src/predict.py'},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def ours_arm(args):
    from oct_segmentation_b200 import synthetic
    from oct_segmentation_b200.parallel import gather_table, shard_range
    from oct_segmentation_b200.pipeline import EnsemblePipeline

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    B, K, W = args.batch, args.steps, args.warmup

    models = synthetic.random_models(dev)
    pipe = EnsemblePipeline(models, CLASSES, OUT_SIZE, dev, B, src_hw=(SRC, SRC), thickness=True)

    # this rank's slice of the global synthetic frame list (weak scaling: B*(K+W) frames per rank)
    total = world * B * (K + W)
    lo, hi = shard_range(total, rank, world)
    n_distinct = min(hi - lo, 4 * B)                                     # 4 distinct batches, cycled
    host = torch.from_numpy(synthetic.synthetic_frames(lo, n_distinct, SRC)).pin_memory()
    dev_batches = [host[i:i + B].to(dev) for i in range(0, n_distinct, B)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput
    for i in range(W):
        pipe.run_device(dev_batches[i % len(dev_batches)])
    counts_log = []
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        _, _, counts, _ = pipe.run_device(dev_batches[i % len(dev_batches)])
        counts_log.append(counts.clone())
    table = gather_table(torch.cat(counts_log), world * B * K) if world > 1 else torch.cat(counts_log)
    e1.record()
    barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = t_ms.item()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (t_ms * 1e-3)

    # ---------------------------------------------------------------- end to end (host buffers)
    host_np = host.numpy()
    for _ in pipe.stream_host((host_np[:B] for _ in range(max(min(W, 3), 2))), copy=False):
        pass
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    starts = [(i % (n_distinct // B)) * B for i in range(K)]
    for mask, label, counts, radii in pipe.stream_host((host_np[j:j + B] for j in starts), copy=False):
        pass                                   # every batch's results are on the host when it is yielded
    e3.record()
    barrier()
    e2e_ms = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (e2e_ms.item() * 1e-3)
    h2d = B * SRC * SRC * 3
    d2h = int(mask.nbytes + label.nbytes + counts.nbytes + radii.nbytes)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    roof = None
    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        tc_ms, tc_flops, other_ms, n_tc = 0.0, 0.0, 0.0, 0
        for _ in range(2):                                                # instrumented eager passes (per-launch events)
            for d in pipe.model_dirs:
                b = pipe.nets[d].builder
                evs = []
                for name, op in zip(b.op_names, b.ops):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    op()
                    e.record()
                    evs.append((name, s, e))
                torch.cuda.synchronize()
                for (name, s, e), plan_macs in zip(evs, b.op_macs):
                    if plan_macs is not None:
                        tc_ms += s.elapsed_time(e)
                        tc_flops += 2.0 * plan_macs
                        n_tc += 1
                    else:
                        other_ms += s.elapsed_time(e)
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'launches_r1_traffic.json')
        if os.path.exists(tpath):                       # dram__bytes_read+write per conv_tc_kernel launch (ncu pass)
            tj = json.load(open(tpath))
            if tj.get('batch', 16) == B:
                traffic, traffic_src = tj['dram_bytes_per_launch'], 'profiles/launches_r1_traffic.json'
        roof = {'bound': 'tensor', 'kernel': 'conv_tc_kernel', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved / peak_tf, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'algorithmic_flops_per_launch': tc_flops / max(n_tc, 1),
                'launches_measured': n_tc, 'avg_launch_ms': tc_ms / max(n_tc, 1),
                'share_of_network_time': tc_ms / (tc_ms + other_ms),
                'how': 'algorithmic FLOPs (2 x dense MACs of the smp graph, DESIGN.md) of every conv_tc_kernel launch in one '
                       'ensemble batch divided by the sum of their CUDA-event durations (eager pass after the timed region)'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': t_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu_per_step': B, 'frames_total': world * B * K,
                       'weights': 'seeded random init of the shipped architectures',
                       'l2': 'inputs cycle over 4 distinct batches; per-step activations (>10 GB) exceed the 126 MB L2',
                       'parallelism': f'frame-sharded x{world}, final all_gather of the counts table'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': int(pipe.launches_per_batch * K),
            'roofline': roof,
            'gflop_per_frame_algorithmic': 2 * pipe.macs_per_frame / 1e9,
        }
        if not args.no_cpu_baseline and world == 1:
            fps, times = cpu_oracle_frames_per_s(args.cpu_frames, 1)
            line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                                    'sample': f'{args.cpu_frames} frames through the oracle restatement of src/predict.py '
                                              f'(batch 1 per call, FC_LC run per class, cv2 pre/post), {sum(times):.1f} s'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not add
    to it: everything written to fd 1 from here on goes to stderr; emit() writes the result line to the real one."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-frames', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == '__main__':
    main()
