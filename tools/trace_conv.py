"""Per-tile pipeline timeline of conv_tc_kernel (CTA 0), from a -DOCTSEG_TRACE build of the library:
  nvcc ... -DOCTSEG_TRACE -o tools/ab/liboctseg_trace.so   (see build_trace() below)
Events per tile: 0 producer start | 1 MMA ready, 2 accumulator free, 3 MMAs committed |
4/7 epilogue group 0/1 ready, 5/8 accumulator-full seen, 6/9 accumulator released."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, 'tools', 'ab', 'liboctseg_trace.so')


def build_trace():
    import __graft_entry__ as g
    srcs = [os.path.join(g.CSRC, s) for s in g.SOURCES]
    subprocess.run([os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')] + g.NVCC_FLAGS + ['-DOCTSEG_TRACE', '-o', TRACE_LIB] + srcs,
                   check=True)


if __name__ == '__main__':
    if sys.argv[1:] == ['build']:
        build_trace()
        sys.exit(0)
    import numpy as np
    import torch
    from oct_segmentation_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB
    from oct_segmentation_b200.engine import conv as CV
    lib = _lib.load()
    lib.octseg_debug_trace.argtypes = [C.c_void_p]
    # (name, [(C,H,W)], cout, k, N, act, out_mode)
    LAYERS = {
        'head': ('linknet head 32->2 @896 (packed f=2)', [(64, 896, 448)], 4, 1, 16, 'none'),
        'vvstem': ('regnet stem s2d 16->32 @448 2x2', [(16, 448, 448)], 32, 2, 16, 'relu'),
        'expand': ('effnet expand 48->288 @224', [(48, 224, 224)], 288, 1, 16, 'swish'),
        'expand_relu': ('effnet expand 48->288 @224 relu', [(48, 224, 224)], 288, 1, 16, 'relu'),
        'e32': ('1x1 32->192 @448 relu', [(32, 448, 448)], 192, 1, 16, 'relu'),
        'e80': ('1x1 80->480 @112 relu', [(80, 112, 112)], 480, 1, 16, 'relu'),
        'proj': ('effnet project 32->32 @448', [(32, 448, 448)], 32, 1, 16, 'none'),
        'c3': ('3x3 64->64 @256', [(64, 256, 256)], 64, 3, 32, 'relu'),
    }
    LAYERS['d2s16'] = ('linknet convT 16->16 @448 (depth-to-space)', [(16, 448, 448)], 16, 4, 16, 'relu')
    LAYERS['vvhead'] = ('unet head 16->1 3x3 @896 (packed f=4, u8 planes)', [(64, 896, 224)], 4, 3, 16, 'none')
    LAYERS['linkhead'] = ('linknet 1x1 16->32 relu + fused 1x1 head 32->2 @896 (packed f=4, u8 planes)', [(64, 896, 224)], 2, 1, 16, 'relu')
    LAYERS['linkhead_f32'] = ('the same, fp32 logits', [(64, 896, 224)], 2, 1, 16, 'relu')
    name, srcs, cout, k, n, act = LAYERS[sys.argv[1]]
    spec = [((n, s[1], s[2], s[0], CV.pad8(s[0])), False) for s in srcs]
    seg_t = [torch.randn(n, s[1], s[2], CV.pad8(s[0]), device='cuda').to(torch.bfloat16) for s in srcs]
    if sys.argv[1] == 'd2s16':
        cs = CV.pad8(cout)
        wp, bp = CV.d2s_weights(torch.randn(srcs[0][0], cout, 4, 4) * 0.05, torch.zeros(cout), True, cs)
        geom, packed = CV.plan_conv(spec, wp, pad=(1, 1))
        bias = CV.pad_bias(bp, geom, 4 * cs)
        out = torch.empty(n, 2 * srcs[0][1], 2 * srcs[0][2], cs, dtype=torch.bfloat16, device='cuda')
        plan = CV.ConvPlan(geom, packed, bias, seg_t, out, act=act, name=name, out_ldc=cs, d2s=cs)
    elif sys.argv[1].startswith('linkhead'):
        w16 = torch.randn(32, 16, 1, 1) * 0.2
        wp, bp = CV.pack_conv_weights(w16, torch.zeros(32), [16], [16], 4, 0, 32)
        geom, packed = CV.plan_conv(spec, wp, stride=1, pad=(0, 0), out_bf16=False)
        bias = CV.pad_bias(bp, geom, 128)
        f32 = sys.argv[1].endswith('f32')
        out = torch.empty(n, 2, 896, 896, dtype=torch.float32 if f32 else torch.uint8, device='cuda')
        plan = CV.ConvPlan(geom, packed, bias, seg_t, out, out_mode='f32_nchw' if f32 else 'u8_nchw', act=act, name=name,
                           out_pack=4, out_ldc=2, head=(torch.randn(2, 32) * 0.2, torch.zeros(2), torch.zeros(32)))
    elif sys.argv[1] == 'vvhead':
        w16 = torch.randn(1, 16, 3, 3) * 0.05
        wp, bp = CV.pack_conv_weights(w16, torch.zeros(1), [16], [16], 4, 1, 1)
        geom, packed = CV.plan_conv(spec, wp, stride=1, pad=(1, 1), out_bf16=False)
        bias = CV.pad_bias(bp, geom, 4)
        out = torch.empty(n, 1, 896, 896, dtype=torch.uint8, device='cuda')
        plan = CV.ConvPlan(geom, packed, bias, seg_t, out, out_mode='u8_nchw', act=act, name=name, out_pack=4, out_ldc=1)
    else:
        w = torch.randn(cout, sum(s[0] for s in srcs), k, k) * 0.05
        geom, packed = CV.plan_conv(spec, w, pad=(k // 2, k // 2), out_hw=(srcs[0][1], srcs[0][2]))
        bias = CV.pad_bias(torch.zeros(cout), geom, cout)
        out = torch.empty(n, geom.out_H, geom.out_W, geom.Cout, dtype=torch.bfloat16, device='cuda')
        plan = CV.ConvPlan(geom, packed, bias, seg_t, out, act=act, name=name)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    print('ms per launch (trace build):', e0.elapsed_time(e1) / 5)
    buf = np.zeros((16, 256), dtype=np.uint64)
    _lib.check(lib.octseg_debug_trace(buf.ctypes.data), 'trace')
    t = buf.astype(np.int64)
    t0 = t[0, 0]
    print(name, 'tile', (geom.TH, geom.TW), 'BN', geom.BN, 'ntn', geom.n_tiles_n, 'kc', [s.kc for s in geom.segs], 'halo', geom.halo, 'Ktot', geom.Ktot)
    print('tile  prod  | mma_rdy acc_free commit | g0_rdy g0_full g0_rel | g1_rdy g1_full g1_rel')
    for i in list(range(0, 12)) + list(range(100, 112)):
        r = [int(t[e, i] - t0) for e in range(10)]
        print(f'{i:4d} {r[0]:6d} | {r[1]:6d} {r[2]:6d} {r[3]:6d} | {r[4]:6d} {r[5]:6d} {r[6]:6d} | {r[7]:6d} {r[8]:6d} {r[9]:6d}')
    print('tile  prod: start setup stage_free issued | mma: rdy acc_free k_wait landed issued commit')
    for i in range(20, 28):
        r = [int(t[e, i] - t0) for e in range(16)]
        print(f'{i:4d} {r[0]:7d} {r[10]-r[0]:5d} {r[11]-r[10]:5d} {r[12]-r[11]:5d} | {r[1]:7d} {r[2]-r[1]:5d} {r[13]-r[2]:5d} {r[14]-r[13]:5d} {r[15]-r[14]:5d} {r[3]-r[15]:5d}')
    lib.octseg_debug_trace_chunks.argtypes = [C.c_void_p]
    cb = np.zeros((8, 256), dtype=np.uint64)
    _lib.check(lib.octseg_debug_trace_chunks(cb.ctypes.data), 'trace chunks')
    c = cb.astype(np.int64)
    print('group 0 chunks: start | tmem_ld  math  wait_store  bar1  sts+fence+bar2  issue | gap to next start')
    for k in range(8, 24):
        r = c[:, k]
        if r[0] == 0:
            break
        print(f'{k:4d} {int(r[0] - t0):8d} | {int(r[1]-r[0]):6d} {int(r[2]-r[1]):6d} {int(r[3]-r[2]):6d} {int(r[4]-r[3]):6d} {int(r[5]-r[4]):6d} '
              f'{int(r[6]-r[5]):6d} | {int(c[0, k + 1] - r[6]):6d}')
    d = np.diff(t[3, 50:200])
    print('steady-state cycles per tile (MMA commit to commit):', float(d.mean()))
