"""A/B timing of the conv kernel: current liboctseg.so vs tools/ab/liboctseg_old.so (commit 549e099),
same process, same GPU, same inputs.  Only layers whose segments all use kc=64."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oct_segmentation_b200 import _lib
from oct_segmentation_b200.engine import conv as CV


class OldSeg(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('ldc', C.c_int32), ('kh', C.c_int32), ('kw', C.c_int32), ('mul', C.c_int32),
                ('off_h', C.c_int32 * 2), ('off_w', C.c_int32 * 2), ('c_per_tile', C.c_int32), ('cchunks', C.c_int32)]


class OldDesc(C.Structure):
    _fields_ = [('nseg', C.c_int32), ('seg', OldSeg * 6), ('phases', C.c_int32),
                ('N', C.c_int32), ('Hq', C.c_int32), ('Wq', C.c_int32), ('TH', C.c_int32), ('TW', C.c_int32),
                ('BN', C.c_int32), ('n_tiles_n', C.c_int32), ('cout_per_tile', C.c_int32), ('Cout', C.c_int32),
                ('weight', C.c_void_p), ('Ktot', C.c_int32), ('per_image_weights', C.c_int32),
                ('bias', C.c_void_p), ('act', C.c_int32), ('res_mode', C.c_int32), ('res', C.c_void_p),
                ('res_ldc', C.c_int32), ('out', C.c_void_p), ('out_mode', C.c_int32), ('out_H', C.c_int32),
                ('out_W', C.c_int32), ('out_ldc', C.c_int32), ('out_c_off', C.c_int32)]


old = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'liboctseg_old.so'))
old.octseg_last_error.restype = C.c_char_p
old.octseg_conv_plan_create.argtypes = [C.POINTER(OldDesc), C.POINTER(C.c_void_p)]
old.octseg_conv_run.argtypes = [C.c_void_p, C.c_void_p]

LAYERS = [
    ('x_3_3.conv2 64->64 @256', [(64, 256, 256, False)], 64, 32),
    ('x_1_3.conv1 up256+3x64->64 @256', [(256, 128, 128, True), (64, 256, 256, False), (64, 256, 256, False), (64, 256, 256, False)], 64, 32),
    ('x_1_2.conv2 256->256 @128', [(256, 128, 128, False)], 256, 32),
    ('l1.conv2 64->64 @128', [(64, 128, 128, False)], 64, 32),
]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, srcs, cout, n in LAYERS:
    spec = [((n, s[1], s[2], s[0], s[0]), s[3]) for s in srcs]
    w = torch.randn(cout, sum(s[0] for s in srcs), 3, 3) * 0.02
    geom, packed = CV.plan_conv(spec, w, pad=(1, 1))
    assert all(sg.kc == 64 for sg in geom.segs)
    bias = CV.pad_bias(torch.zeros(cout), geom, cout)
    seg_t = [torch.randn(n, s[1], s[2], s[0], device='cuda').to(torch.bfloat16) for s in srcs]
    out_new = torch.empty(n, geom.out_H, geom.out_W, geom.Cout, dtype=torch.bfloat16, device='cuda')
    out_old = torch.empty_like(out_new)
    plan = CV.ConvPlan(geom, packed, bias, seg_t, out_new, act='relu', name=name)
    d = OldDesc()
    d.nseg = len(geom.segs)
    for i, (sg, t) in enumerate(zip(geom.segs, seg_t)):
        s = d.seg[i]
        s.ptr, s.N, s.H, s.W, s.C, s.ldc = t.data_ptr(), sg.N, sg.H, sg.W, sg.C, sg.ldc
        s.kh, s.kw, s.mul = sg.kh, sg.kw, sg.mul
        s.off_h[0], s.off_h[1] = sg.off_h
        s.off_w[0], s.off_w[1] = sg.off_w
        s.c_per_tile, s.cchunks = sg.c_per_tile, sg.cchunks
    d.phases, d.N, d.Hq, d.Wq, d.TH, d.TW = geom.phases, geom.N, geom.Hq, geom.Wq, geom.TH, geom.TW
    d.BN, d.n_tiles_n, d.cout_per_tile, d.Cout = geom.BN, geom.n_tiles_n, geom.cout_per_tile, geom.Cout
    d.weight, d.Ktot, d.per_image_weights = plan.weight.data_ptr(), geom.Ktot, 0
    d.bias, d.act, d.res_mode, d.res, d.res_ldc = plan.bias.data_ptr(), 1, 0, None, 0
    d.out, d.out_mode, d.out_H, d.out_W, d.out_ldc, d.out_c_off = out_old.data_ptr(), 0, geom.out_H, geom.out_W, out_old.shape[-1], 0
    h = C.c_void_p()
    rc = old.octseg_conv_plan_create(C.byref(d), C.byref(h))
    assert rc == 0, old.octseg_last_error()
    st = torch.cuda.current_stream().cuda_stream
    t_new = timeit(plan.run)
    t_old = timeit(lambda: old.octseg_conv_run(h, st))
    t_new2 = timeit(plan.run)
    same = torch.equal(out_new, out_old)
    print(f'{name:40s} old {t_old:.4f} ms   new {t_new:.4f} / {t_new2:.4f} ms   identical={same}', flush=True)
