"""GPU parity (bit-exact) of the pre/post-processing kernels through the C-ABI against the CPU
oracle, cv2 and the reference-generated goldens."""
import os

import cv2
import numpy as np
import pytest
import torch

from oct_segmentation_b200 import prepost as P
from oracle import prepost_ref as R
from oracle import synth
from tests.test_oracle_prepost import G, unpack

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('src,dst', [(250, 96), (250, 125), (512, 512), (1000, 512), (1000, 896), (1024, 512), (750, 896)])
def test_preprocess_bit_exact_vs_cv2(src, dst):
    rng = np.random.default_rng(src + dst)
    frames = rng.integers(0, 256, (3, src, src, 3), dtype=np.uint8)
    got = P.preprocess(torch.from_numpy(frames).cuda(), dst).cpu().numpy()
    for f, g in zip(frames, got):
        assert np.array_equal(g, cv2.resize(cv2.cvtColor(f, cv2.COLOR_RGB2BGR), (dst, dst)))


def test_preprocess_real_frame_matches_reference_output():
    d = np.load(os.path.join(G, 'demo_frame_small.npz'))
    x = torch.from_numpy(d['rgb'][None].copy()).cuda()
    for S in (96, 160, 125):
        assert np.array_equal(P.preprocess(x, S)[0].cpu().numpy(), d[f'pre{S}'])


def oracle_post(planes_np, order_names, Ho, Wo):
    """predict.py:92-100 per frame with cv2, then label map + counts with numpy."""
    meta = {'Lumen': ('LM', 0), 'Lipid core': ('FC_LC', 0), 'Fibrous cap': ('FC_LC', 1), 'Vasa vasorum': ('VV', 0)}
    N = next(iter(planes_np.values())).shape[0]
    masks, labels, counts = [], [], []
    for n in range(N):
        mask = np.zeros((Ho, Wo, 4))
        for name in order_names:
            mdir, idx = meta[name]
            pm = planes_np[mdir][n].astype(np.float32)              # (S, S, C) like predict() returns
            rm = cv2.resize(pm, (Wo, Ho), interpolation=cv2.INTER_NEAREST)
            if rm.ndim > 2:
                rm = rm[:, :, idx]
            mask[:, :, R.CLASS_NAMES.index(name)] = rm
        masks.append(mask.astype(np.uint8))
        labels.append(R.label_map(mask, order_names))
        counts.append([R.area_count(mask[:, :, c]) for c in range(4)])
    return np.stack(masks), np.stack(labels), np.array(counts)


@pytest.mark.parametrize('Ho,S_lm,S_big,order', [
    (1000, 512, 896, ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']),
    (250, 128, 224, ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']),
    (333, 96, 160, ['Vasa vasorum', 'Lumen']),                      # subset + different paint order
])
def test_postprocess_bit_exact(Ho, S_lm, S_big, order):
    rng = np.random.default_rng(Ho)
    N = 2

    def blobs(S, C):
        yy, xx = np.mgrid[0:S, 0:S]
        out = np.zeros((N, S, S, C), np.uint8)
        for n in range(N):
            for c in range(C):
                cx, cy, r = rng.uniform(0.3, 0.7) * S, rng.uniform(0.3, 0.7) * S, rng.uniform(0.1, 0.3) * S
                out[n, :, :, c] = ((xx - cx) ** 2 + (yy - cy) ** 2 < r * r) ^ (rng.random((S, S)) < 0.02)
        return out
    planes_np = {'LM': blobs(S_lm, 1), 'FC_LC': blobs(S_big, 2), 'VV': blobs(S_big, 1)}
    dev = {k: torch.from_numpy(np.ascontiguousarray(v.transpose(0, 3, 1, 2))).cuda() for k, v in planes_np.items()}
    # routing of predict.py:23-28: class channel -> (model, model channel)
    route = {0: dev['LM'][:, 0], 1: dev['FC_LC'][:, 1], 2: dev['FC_LC'][:, 0], 3: dev['VV'][:, 0]}
    planes = {c: route[c].contiguous() for c in range(4) if R.CLASS_NAMES[c] in order}
    mask, label, counts = P.postprocess(planes, [R.CLASS_NAMES.index(n) for n in order], Ho, Ho, N, 'cuda')
    wm, wl, wc = oracle_post(planes_np, order, Ho, Ho)
    assert np.array_equal(mask.cpu().numpy(), wm)
    assert np.array_equal(label.cpu().numpy(), wl)
    assert np.array_equal(counts.cpu().numpy(), wc)


def test_counts_and_radial_thickness_match_reference_goldens():
    d = np.load(os.path.join(G, 'masks_app_demo.npz'))
    q = np.load(os.path.join(G, 'quantities_ref.npz'))
    masks = np.stack([unpack(p, d['shape']) for p in d['packed']])           # (12, 750, 750, 4) {0,255}
    dev = torch.from_numpy(masks).cuda()
    radii = P.radial_thickness(dev).cpu().numpy()
    counts = (dev != 0).sum(dim=(1, 2)).cpu().numpy()
    assert np.array_equal(counts, d['counts'][d['keep_idx']])
    rows = P.quantities_from_counts(counts, 750, 750, int(q['ratio']), radii)
    for k in range(masks.shape[0]):
        for c, name in enumerate(R.CLASS_NAMES):
            present, nnz, area, _, _, rmed, rmin, rmax = q['q'][k, c]
            hits = q['radii_hits'][k, c]
            r = radii[k, c]
            assert np.array_equal(r[r > 0], hits[hits >= 0]), (k, name)
            row = rows[k][name]
            assert row['present'] == bool(present) and row['nnz'] == int(nnz)
            if present:
                assert row['area'] == area
            assert row['thickness_median'] == rmed and row['thickness_min'] == rmin and row['thickness_max'] == rmax


def test_postprocess_full_size_properties():
    """At BASELINE's full output size: counts equal the mask's own popcount, label obeys the priority
    rule, and upscaling a plane by the nearest LUT preserves its bounding rows/cols."""
    N, Ho = 4, 1000
    g = torch.Generator(device='cuda').manual_seed(5)
    planes = {c: (torch.rand(N, S, S, device='cuda', generator=g) < 0.3).to(torch.uint8)
              for c, S in zip(range(4), (512, 896, 896, 896))}
    mask, label, counts = P.postprocess(planes, [0, 1, 2, 3], Ho, Ho, N, 'cuda')
    assert torch.equal(counts.long(), mask.long().sum(dim=(1, 2)))
    want = torch.zeros_like(label)
    for c in range(4):
        want[mask[..., c] != 0] = c + 1
    assert torch.equal(label, want)
    lut = torch.from_numpy(P.nearest_table(512, Ho)).cuda().long()
    assert torch.equal(mask[..., 0], planes[0][:, lut][:, :, lut])


def test_overlay_matches_reference_save_results_outputs():
    """octseg_overlay vs the <name>_overlay.png the reference's own save_results wrote
    (src/data/utils.py:195-235; tests/golden/make_golden.py), bit for bit."""
    d = np.load(os.path.join(G, 'overlay_ref.npz'))
    for i in range(int(d['n'])):
        order = [R.CLASS_IDS[str(c)] - 1 for c in d[f'classes{i}']]
        got = P.overlay(torch.from_numpy(d[f'frame{i}'][None].copy()).cuda(),
                        torch.from_numpy(d[f'mask{i}'][None].copy()).cuda(), order)[0].cpu().numpy()
        ref = d[f'overlay{i}']
        assert np.array_equal(got, ref), f'case {i}: {(got != ref).any(axis=2).sum()} pixels differ'


@pytest.mark.parametrize('H,W,N', [(3, 3, 1), (8, 5, 2), (33, 70, 3), (257, 130, 2)])
def test_overlay_bit_exact_vs_oracle_ragged_sizes(H, W, N):
    """Tile borders, images smaller than one tile, random speckle + blobs, non-{0,1} mask bytes, class subsets."""
    rng = np.random.default_rng(H * 1000 + W)
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    mask = (rng.random((N, H, W, 4)) > 0.8).astype(np.uint8) * rng.integers(1, 256, (N, H, W, 4), dtype=np.uint8)
    yy, xx = np.mgrid[:H, :W]
    mask[:, :, :, 0] |= ((yy - H / 2) ** 2 + (xx - W / 2) ** 2 < (min(H, W) / 3) ** 2).astype(np.uint8)
    for order in ([0, 1, 2, 3], [3, 0], [2], []):
        got = P.overlay(torch.from_numpy(frames).cuda(), torch.from_numpy(mask).cuda(), order).cpu().numpy()
        names = [R.CLASS_NAMES[c] for c in order]
        for n in range(N):
            assert np.array_equal(got[n], R.overlay(frames[n], mask[n], names)), (order, n)


def test_overlay_full_size_properties():
    """1000 x 1000 (configs/predict.yaml output_size): an empty mask leaves the frame untouched, a full mask
    gives the closed-form double paste everywhere, and painting is local (pixels > 7 away from any object
    keep the frame value)."""
    rng = np.random.default_rng(11)
    H = W = 1000
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    x = torch.from_numpy(frames).cuda()
    empty = torch.zeros(2, H, W, 4, dtype=torch.uint8, device='cuda')
    assert torch.equal(P.overlay(x, empty, [0, 1, 2, 3]), x)
    full = torch.zeros(2, H, W, 4, dtype=torch.uint8, device='cuda')
    full[..., 0] = 1
    fill, rim = P.overlay_alpha_tables()
    want = R.pil_paste(frames, R.CLASS_COLORS_RGB['Lumen'], np.full((2, H, W), fill[256]))   # no rim: erode == 1
    assert np.array_equal(P.overlay(x, full, [0]).cpu().numpy(), want)
    m = np.zeros((2, H, W, 4), np.uint8)
    m[:, 400:600, 300:500, 3] = 1
    m[:, 0:20, 980:1000, 1] = 1
    got = P.overlay(x, torch.from_numpy(m).cuda(), [0, 1, 2, 3]).cpu().numpy()
    far = cv2.dilate(m.any(axis=3)[0].astype(np.uint8), np.ones((15, 15), np.uint8)) == 0
    assert np.array_equal(got[0][far], frames[0][far]) and (got[0][~far] != frames[0][~far]).any()
    for n in range(2):
        assert np.array_equal(got[n], R.overlay(frames[n], m[n], R.CLASS_NAMES))


@pytest.mark.parametrize('src,dst', [(250, 96), (250, 125), (512, 512), (1000, 896), (1024, 512), (37, 64)])
def test_preprocess_grayscale_equals_replicated_frame(src, dst):
    """Grayscale extension: (N, H, W) and (N, H, W, 1) frames == cv2 on the replicated frame."""
    rng = np.random.default_rng(src * 7 + dst)
    g = rng.integers(0, 256, (2, src, src + 6), dtype=np.uint8)
    a = P.preprocess(torch.from_numpy(g).cuda(), dst).cpu().numpy()
    b = P.preprocess(torch.from_numpy(g[..., None].copy()).cuda(), dst).cpu().numpy()
    for n in range(2):
        want = cv2.resize(cv2.cvtColor(np.repeat(g[n][..., None], 3, axis=2), cv2.COLOR_RGB2BGR), (dst, dst))
        assert np.array_equal(a[n], want) and np.array_equal(b[n], want)


@pytest.mark.parametrize('K,shape', [(1, (2, 1, 64, 64)), (3, (2, 2, 96, 96)), (5, (1, 1, 33, 7)), (8, (3, 2, 32, 32))])
def test_fold_average_threshold_vs_oracle(K, shape):
    """K-way probability averaging (opt-in): equal to the float64 oracle wherever |mean - 0.5| exceeds the
    fp32 rounding band (1e-6); K = 1 is exactly `y > 0` (== sigmoid(y) > 0.5, src/models/smp/model.py:195)."""
    rng = np.random.default_rng(K)
    logits = [(rng.standard_normal(shape) * 3).astype(np.float32) for _ in range(K)]
    logits[0].flat[:5] = [0.0, -0.0, 1e-3, -1e-3, 80.0]
    got = P.fold_average_threshold([torch.from_numpy(x).cuda() for x in logits]).cpu().numpy()
    want, margin = R.fold_average_threshold(logits)
    sure = margin > 1e-6
    assert sure.mean() > 0.999 and np.array_equal(got[sure], want[sure])
    if K == 1:
        assert np.array_equal(got, (logits[0] > 0).astype(np.uint8))


def _gpu_contour_thickness(mask4, cap=P.CONTOUR_CAP):
    sums, nverts, verts = (t.cpu().numpy() for t in P.contour_largest(torch.from_numpy(np.ascontiguousarray(mask4)).cuda(), cap))
    return sums, nverts, verts


@pytest.mark.parametrize('H,W', [(1, 1), (5, 3), (31, 33), (64, 64), (70, 129), (200, 255)])
def test_contour_largest_equals_cv2(H, W):
    """octseg_contour_largest vs cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + max(contourArea): same points
    in the same order, same thickness dict as calculate_thickness_contour (analysis.py:21-57); noise at several
    densities, blobs, nested components, arbitrary non-zero bytes, empty and full planes."""
    rng = np.random.default_rng(H * 31 + W)
    N = 6
    m = np.zeros((N, H, W, 4), np.uint8)
    for n in range(N):
        for c in range(4):
            p = (rng.random((H, W)) < rng.choice([0.05, 0.3, 0.5, 0.7, 0.95])).astype(np.uint8)
            if (n + c) % 2:
                p = cv2.dilate(p, np.ones((3, 3), np.uint8))
            if c == 3:
                p = p * rng.integers(1, 256, (H, W)).astype(np.uint8)
            m[n, :, :, c] = p
    m[0, :, :, 0] = 0
    m[0, :, :, 1] = 1
    sums, nverts, verts = _gpu_contour_thickness(m)
    for n in range(N):
        for c in range(4):
            ch = np.ascontiguousarray(m[n, :, :, c])
            contours, _ = cv2.findContours(ch, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            best = max(contours, key=cv2.contourArea) if contours else None
            if best is None or cv2.contourArea(best) == 0:
                assert nverts[n, c] == 0 and sums[n, c, 3] == -1
            else:
                assert np.array_equal(verts[n, c, :nverts[n, c]], best.reshape(-1, 2)), (n, c)
                assert abs(int(sums[n, c, 0])) == 2 * cv2.contourArea(best)
            assert P.thickness_from_contour(sums[n, c], int(nverts[n, c]), verts[n, c]) == R.thickness_contour(ch), (n, c)


def test_contour_thickness_matches_reference_function_goldens():
    """tests/golden/quantities_ref.npz: calculate_thickness_contour's own outputs on the reference's 750 x 750 demo masks."""
    d = np.load(os.path.join(G, 'masks_app_demo.npz'))
    q = np.load(os.path.join(G, 'quantities_ref.npz'))
    masks = np.stack([unpack(p, d['shape']) for p in d['packed']])
    sums, nverts, verts = _gpu_contour_thickness(masks)
    for k in range(masks.shape[0]):
        for c in range(4):
            t = P.thickness_from_contour(sums[k, c], int(nverts[k, c]), verts[k, c])
            assert t['median'] == q['q'][k, c, 3] and t['min'] == q['q'][k, c, 4], (k, c)


def test_contour_thickness_full_size():
    """1000 x 1000: OCT-shaped masks and noise against the cv2 oracle; a disc's border is found whole (every kept
    point within one pixel of the radius); capacity overflow is reported, never truncated silently."""
    H = W = 1000
    yy, xx = np.mgrid[:H, :W]
    rr = np.sqrt((yy - 500.0) ** 2 + (xx - 480.0) ** 2)
    m = np.zeros((2, H, W, 4), np.uint8)
    m[0, :, :, 0] = rr < 220
    m[0, :, :, 1] = (rr >= 220) & (rr < 260) & (np.abs(np.arctan2(yy - 500.0, xx - 480.0)) < 1.0)
    m[0, :, :, 2] = (np.random.default_rng(1).random((H, W)) < 0.3)      # below the 8-connected percolation threshold
    m[0, :, :, 3] = ((yy - 200) ** 2 + (xx - 300) ** 2 < 64) | ((yy - 750) ** 2 + (xx - 700) ** 2 < 100)
    m[1, 0, :, 0] = 1
    m[1, :, 0, 0] = 1
    m[1, -1, :, 0] = 1
    m[1, :, -1, 0] = 1                                    # a frame-sized ring: the border walks all four image edges
    m[1, 3:-3:2, 3:-3, 1] = 1
    m[1, 3:-3, 3, 1] = 1                                  # a comb with ~ 2000 kept points
    sums, nverts, verts = _gpu_contour_thickness(m)
    for n in range(2):
        for c in range(4):
            ch = np.ascontiguousarray(m[n, :, :, c])
            assert P.thickness_from_contour(sums[n, c], int(nverts[n, c]), verts[n, c]) == R.thickness_contour(ch), (n, c)
    v = verts[0, 0, :nverts[0, 0]].astype(np.float64)
    assert np.all(np.abs(np.sqrt((v[:, 1] - 500) ** 2 + (v[:, 0] - 480) ** 2) - 220) < 1.5)
    s2, n2, v2 = _gpu_contour_thickness(m[1:], cap=64)
    assert n2[0, 1] > 64
    with pytest.raises(RuntimeError):
        P.thickness_from_contour(s2[0, 1], int(n2[0, 1]), v2[0, 1])


@pytest.mark.parametrize('Hs,Ws,S', [(1000, 1000, 512), (1000, 1000, 896), (512, 512, 896), (1024, 1024, 512), (250, 300, 128), (64, 64, 32)])
def test_preprocess_s2d_equals_cv2_then_stem_pack(Hs, Ws, S):
    """octseg_preprocess_resize_s2d = preprocessing_img (cv2 bilinear + BGR, bit for bit) followed by the exact uint8 ->
    bf16 space-to-depth packing the network stems read: unpacking it gives cv2's frame, and it equals what the separate
    stem-pack launch makes of the uint8 result (so the fused path feeds the networks identical bits)."""
    import ctypes as C
    from oct_segmentation_b200 import _lib
    rng = np.random.default_rng(Hs + S)
    frames = rng.integers(0, 256, (2, Hs, Ws, 3), dtype=np.uint8)
    dev = torch.device('cuda')
    x2 = P.preprocess_s2d(torch.from_numpy(frames).to(dev), S)
    assert x2.shape == (2, S // 2, S // 2, 16) and x2.dtype == torch.bfloat16 and (x2[..., 12:] == 0).all()
    want = np.stack([R.preprocess_frame(f, S) for f in frames])
    assert np.array_equal(P.unpack_s2d(x2).float().cpu().numpy(), want.astype(np.float32))
    # the two-launch path: uint8 resize, then octseg_stem_pack
    u8 = P.preprocess(torch.from_numpy(frames).to(dev), S)
    ref = torch.empty_like(x2)
    xv = u8.permute(0, 3, 1, 2)
    sn, sc, sh, sw = xv.stride()
    _lib.check(_lib.load().octseg_stem_pack(u8.data_ptr(), 1, sn, sc, sh, sw, 2, S, S, None, None, ref.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), 'stem_pack')
    assert torch.equal(x2, ref)
    # grayscale frames are replicated
    g = P.preprocess_s2d(torch.from_numpy(np.ascontiguousarray(frames[..., 0])).to(dev), S)
    want_g = np.stack([R.preprocess_frame(np.repeat(f[..., :1], 3, -1), S) for f in frames])
    assert np.array_equal(P.unpack_s2d(g).float().cpu().numpy(), want_g.astype(np.float32))


@pytest.mark.parametrize('in_dtype', ['f32', 'u8'])
def test_stem_pack_applies_forward_normalisation(in_dtype):
    """OCTSegmentationModel.forward's (image - mean) / std (model.py:65-71) happens inside octseg_stem_pack: its packed
    bf16 output equals torch's fp32 normalisation rounded to bf16 (one bf16 ulp of slack for x * (1/std) vs x / std),
    and a std that is off by 2 % -- the smallest mix-up between the three ImageNet channels -- does NOT pass."""
    import ctypes as C
    from oct_segmentation_b200 import _lib
    from oct_segmentation_b200.smp import _IMAGENET
    mean, std = _IMAGENET['mean'], _IMAGENET['std']
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 96, generator=g)                          # forward() sees whatever scale the caller uses
    if in_dtype == 'u8':
        x = (x * 255).to(torch.uint8)
    xd = x.cuda()
    sn, sc, sh, sw = xd.stride()
    out = torch.full((2, 32, 48, 16), float('nan'), dtype=torch.bfloat16, device='cuda')
    _lib.check(_lib.load().octseg_stem_pack(xd.data_ptr(), 0 if in_dtype == 'f32' else 1, sn, sc, sh, sw, 2, 64, 96,
                                            (C.c_float * 3)(*mean), (C.c_float * 3)(*[1.0 / v for v in std]),
                                            out.data_ptr(), torch.cuda.current_stream().cuda_stream), 'stem_pack')
    got = P.unpack_s2d(out).float().cpu()                               # (N, H, W, 3)
    assert (out[..., 12:] == 0).all()

    def want(std_):
        m = torch.tensor(mean).view(1, 3, 1, 1)
        sd = torch.tensor(std_).view(1, 3, 1, 1)
        return ((x.float() - m) / sd).to(torch.bfloat16).float().permute(0, 2, 3, 1)
    w = want(std)
    ulp = w.abs().clamp_min(2.0 ** -126) * 2.0 ** -7
    assert ((got - w).abs() <= ulp).all()
    assert (got == w).float().mean() > 0.99
    wrong = want([std[1], std[0], std[2]])                              # 0.229 <-> 0.224
    assert not ((got - wrong).abs() <= wrong.abs().clamp_min(2.0 ** -126) * 2.0 ** -7).all()
