"""CPU: the product's border-following core (csrc/contour_core.h, the functions octseg_contour_largest runs on the
GPU) compiled with g++ and checked against cv2.findContours and against the reference-generated goldens, plus the
host finish `prepost.thickness_from_contour` against calculate_thickness_contour (src/app/tools/analysis.py:21-57)."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

from oct_segmentation_b200 import prepost as P
from oracle import prepost_ref as R
from tests.test_oracle_prepost import G, unpack

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('contour') / 'contour_core_harness.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', so, os.path.join(HERE, 'contour_core_harness.cpp')])
    lib = C.CDLL(so)
    lib.contour_largest_cpu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

    def run(mask_u8, cap=16384):
        m = np.ascontiguousarray(mask_u8)
        H, W = m.shape
        sums = np.zeros(4, np.int64)
        nverts, n_outer = C.c_int(0), C.c_int(0)
        verts = np.zeros((cap, 2), np.int16)
        lib.contour_largest_cpu(m.ctypes.data, H, W, sums.ctypes.data, C.byref(nverts), verts.ctypes.data, cap, C.byref(n_outer))
        return sums, nverts.value, verts, n_outer.value
    return run


def cv2_largest(mask_u8):
    contours, _ = cv2.findContours(mask_u8, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return max(contours, key=cv2.contourArea) if contours else None


def test_core_equals_cv2_on_random_masks(harness):
    """Noise at several densities, dilated blobs, nested components inside holes, objects touching every border."""
    rng = np.random.default_rng(0)
    for trial in range(400):
        H, W = (int(v) for v in rng.integers(1, 70, 2))
        m = (rng.random((H, W)) < rng.choice([0.1, 0.3, 0.5, 0.7, 0.9])).astype(np.uint8)
        if trial % 3 == 0:
            m = cv2.dilate(m, np.ones((3, 3), np.uint8))
        if trial % 5 == 0:
            m = m * rng.integers(1, 256, (H, W)).astype(np.uint8)       # any non-zero byte is foreground
        sums, nv, verts, n_outer = harness(m)
        _, hier = cv2.findContours(m, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
        assert n_outer == (0 if hier is None else int((hier[0][:, 3] == -1).sum()))
        best = cv2_largest(m)
        want = R.thickness_contour(m)
        if best is None or cv2.contourArea(best) == 0:
            assert nv == 0 and sums[0] == 0
        else:
            assert np.array_equal(verts[:nv], best.reshape(-1, 2)), trial
            assert abs(int(sums[0])) == 2 * cv2.contourArea(best)
            M = cv2.moments(best)
            assert abs(int(sums[1])) / 6 == pytest.approx(M['m10'], rel=1e-12) and abs(int(sums[2])) / 6 == pytest.approx(M['m01'], rel=1e-12)
        assert P.thickness_from_contour(sums, nv, verts) == want, trial


def test_core_and_host_finish_match_reference_function_goldens(harness):
    """tests/golden/quantities_ref.npz holds calculate_thickness_contour's own outputs on the reference's demo masks."""
    d = np.load(os.path.join(G, 'masks_app_demo.npz'))
    q = np.load(os.path.join(G, 'quantities_ref.npz'))
    for k, packed in enumerate(d['packed']):
        m = unpack(packed, d['shape'])
        for c in range(4):
            ch = np.ascontiguousarray(m[:, :, c])
            sums, nv, verts, _ = harness(ch)
            t = P.thickness_from_contour(sums, nv, verts)
            assert t['median'] == q['q'][k, c, 3] and t['min'] == q['q'][k, c, 4]


def test_overflowing_contour_is_reported(harness):
    m = np.zeros((40, 40), np.uint8)
    m[5:35:2, 5:35] = 1
    m[5:35, 5] = 1                                      # a comb: many kept points
    sums, nv, verts, _ = harness(m, cap=8)
    assert nv > 8
    with pytest.raises(RuntimeError):
        P.thickness_from_contour(sums, nv, verts)
