"""CPU: the pre/post oracle (oracle/prepost_ref.py) pinned against (a) cv2 itself, (b) outputs of the
reference's own functions run in the build container (tests/golden/make_golden.py), (c) the
reference's mask -> mask_color fixture pairs; and the product's host-side tables against the oracle."""
import os

import cv2
import numpy as np
import pytest

from oct_segmentation_b200 import prepost as P
from oracle import prepost_ref as R

G = os.path.join(os.path.dirname(__file__), 'golden')


def unpack(packed, shape):
    h, w = int(shape[0]), int(shape[1])
    return (np.unpackbits(packed)[:h * w * 4].reshape(h, w, 4) * 255).astype(np.uint8)


@pytest.mark.parametrize('src,dst', [(250, 96), (250, 160), (250, 125), (1000, 512), (1000, 896), (750, 512), (512, 896), (1024, 512)])
def test_resize_linear_matches_cv2(src, dst):
    rng = np.random.default_rng(src * 7 + dst)
    img = rng.integers(0, 256, (src, src, 3), dtype=np.uint8)
    want = cv2.resize(img, (dst, dst))
    got = R.resize_linear_u8(img, (dst, dst))
    assert np.array_equal(got, want)


def test_preprocess_matches_reference_function_output():
    d = np.load(os.path.join(G, 'demo_frame_small.npz'))
    for S in (96, 160, 125):
        want = d[f'pre{S}']                       # produced by the reference's preprocessing_img
        assert np.array_equal(R.preprocess_frame(d['rgb'], S), want)
        assert np.array_equal(R.resize_linear_u8(d['rgb'][..., ::-1], (S, S)), want)


@pytest.mark.parametrize('src,dst', [(512, 1000), (896, 1000), (1024, 1000), (128, 300), (96, 250)])
def test_nearest_index_matches_cv2(src, dst):
    ramp = np.arange(src, dtype=np.float32)[None, :].repeat(2, 0)
    want = cv2.resize(ramp, (dst, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int32)
    assert np.array_equal(R.nearest_index(src, dst), want)
    assert np.array_equal(P.nearest_table(src, dst), want)
    if dst == 1000:
        assert np.array_equal(want, (np.arange(dst) * src) // dst)      # SURVEY App. E integer rule


@pytest.mark.parametrize('src,dst', [(250, 96), (1000, 512), (1000, 896), (750, 896)])
def test_product_linear_tables_match_oracle(src, dst):
    xo, xa = P.linear_tables(src, dst)
    ro, ra = R.linear_coeffs(src, dst)
    assert np.array_equal(xo, ro) and np.array_equal(xa, ra)


def test_color_mask_matches_reference_fixture_pairs():
    d = np.load(os.path.join(G, 'colorize_pairs.npz'))
    for packed, want in zip(d['packed'], d['colors']):
        m = unpack(packed, d['shape'])
        assert np.array_equal(R.color_mask(m), want)
        lab = R.label_map(m)
        # priority VV > LC > FC > LM where channels overlap
        assert (lab[m[:, :, 3] != 0] == 4).all()
        assert (lab[(m[:, :, 2] != 0) & (m[:, :, 3] == 0)] == 3).all()


def test_quantities_match_reference_function_outputs():
    d = np.load(os.path.join(G, 'masks_app_demo.npz'))
    q = np.load(os.path.join(G, 'quantities_ref.npz'))
    ratio = int(q['ratio'])
    assert ratio == R.dicom_ratio(750)
    for k, packed in enumerate(d['packed']):
        m = unpack(packed, d['shape'])
        assert np.array_equal([R.area_count(m[:, :, c]) for c in range(4)], d['counts'][d['keep_idx'][k]])
        for c in range(4):
            ch = np.ascontiguousarray(m[:, :, c])
            present, nnz, area, cmed, cmin, rmed, rmin, rmax = q['q'][k, c]
            assert R.class_present(ch) == bool(present) and R.area_count(ch) == int(nnz)
            assert R.area_value(ch, ratio) == area
            t = R.thickness_contour(ch)
            assert t['median'] == cmed and t['min'] == cmin
            if k < 3:                                   # the pure-Python radial scan is slow
                radii = R.radial_radii(ch)
                hits = q['radii_hits'][k, c]
                assert np.array_equal(radii[radii > 0], hits[hits >= 0])
                rt = R.radial_thickness(ch)
                assert rt['median'] == rmed and rt['min'] == rmin and rt['max'] == rmax


def test_object_id_tracking():
    assert R.object_ids([True, True, False, True, False, False, True, True]) == [0, 0, 1, 2, 2]
    assert R.object_ids([False, False]) == []


def test_quantities_from_counts_host_logic():
    counts = np.array([[0, 150 * 4 + 7, 1000 * 1000, 12345]])
    rows = P.quantities_from_counts(counts, 1000, 1000, 150)
    r = rows[0]
    assert not r['Lumen']['present'] and not r['Lipid core']['present']       # empty / full channel
    assert r['Fibrous cap']['present'] and r['Fibrous cap']['area'] == 2.0
    assert r['Vasa vasorum']['area'] == pow(12345 // 150, 0.5)


def test_overlay_matches_reference_save_results_outputs():
    """oracle.overlay / color_mask vs the PNGs the reference's own save_results (src/data/utils.py:195-235)
    wrote for the same frames and masks (tests/golden/make_golden.py), bit for bit."""
    d = np.load(os.path.join(G, 'overlay_ref.npz'))
    for i in range(int(d['n'])):
        classes = [str(c) for c in d[f'classes{i}']]
        got = R.overlay(d[f'frame{i}'], d[f'mask{i}'], classes)
        assert np.array_equal(got, d[f'overlay{i}']), f'case {i}: {(got != d[f"overlay{i}"]).any(axis=2).sum()} pixels differ'
        assert np.array_equal(R.color_mask(d[f'mask{i}'], classes), d[f'colormask{i}'])


def test_overlay_building_blocks_match_cv2_and_numpy():
    assert np.array_equal(cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), R.ELLIPSE5)
    assert np.array_equal(cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7)), R.ELLIPSE7)
    assert np.allclose(cv2.getGaussianKernel(5, 0).ravel() * 16, [1, 4, 6, 4, 1])
    rng = np.random.default_rng(3)
    m = (rng.random((40, 52)) > 0.6).astype(np.float64)
    closed = cv2.morphologyEx(m, cv2.MORPH_CLOSE, R.ELLIPSE5)
    assert np.array_equal(closed != 0, R._morph(R._morph(m != 0, R.ELLIPSE5, True), R.ELLIPSE5, False))
    assert np.array_equal(cv2.dilate(m, R.ELLIPSE7) != 0, R._morph(m != 0, R.ELLIPSE7, True))
    assert np.array_equal(cv2.erode(m, R.ELLIPSE7) != 0, R._morph(m != 0, R.ELLIPSE7, False))
    fill, rim = R.overlay_alpha_tables()
    k = np.arange(257, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        assert np.array_equal(fill, (((k / 256.0) * 64) * 0.85 * 255).astype('uint8'))   # numpy's wrapping cast (x86)
    assert rim == 231 and fill[256] == 48
    # the product's host tables are the oracle's
    pf, pr = P.overlay_alpha_tables()
    assert np.array_equal(pf, fill) and pr == rim


def test_fold_average_oracle_reduces_to_the_reference_threshold():
    """K = 1: mean sigmoid > 0.5 == `y.sigmoid() > 0.5` (src/models/smp/model.py:195) == y > 0; K folds: majority
    of confident folds wins, symmetric logits tie at exactly 0.5 -> 0."""
    import torch
    y = (np.random.default_rng(0).standard_normal((2, 1, 16, 16)) * 4).astype(np.float32)
    m, _ = R.fold_average_threshold([y])
    assert np.array_equal(m, (torch.from_numpy(y).sigmoid() > 0.5).numpy().astype(np.uint8))
    assert np.array_equal(m, (y > 0).astype(np.uint8))
    a = np.array([10.0, 10.0, -10.0, 2.0], np.float32)
    b = np.array([10.0, -10.0, -10.0, -2.0], np.float32)
    c = np.array([-10.0, -10.0, 10.0, 0.5], np.float32)
    m, margin = R.fold_average_threshold([a, b, c])
    assert m.tolist() == [1, 0, 0, 1]
    m2, margin2 = R.fold_average_threshold([a[3:], b[3:]])
    assert m2.tolist() == [0] and margin2[0] < 1e-12
