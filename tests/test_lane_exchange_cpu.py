"""CPU: the warp-level bit exchange of overlay_kernel's packing stage (csrc/overlay.cu, S0), emulated lane by lane
with the byte-permute selectors READ FROM THE SOURCE: after the pair shuffle and the two transpose rounds, even lane
2k of every 8-lane group must hold the 32-pixel plane word of class k."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oct_segmentation_b200', 'csrc', 'overlay.cu')


def byte_perm(x, y, s):
    b = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(b[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def test_overlay_pack_transpose_selectors():
    src = open(SRC).read()
    m1 = re.search(r'__byte_perm\(u, v, \(lane & 2\) \? (0x[0-9a-fA-F]+)u : (0x[0-9a-fA-F]+)u\)', src)
    m2 = re.search(r'__byte_perm\(u, v, \(lane & 4\) \? (0x[0-9a-fA-F]+)u : (0x[0-9a-fA-F]+)u\)', src)
    assert m1 and m2, 'packing stage of overlay.cu changed: update this emulation'
    r1_odd, r1_even, r2_hi, r2_lo = int(m1.group(1), 16), int(m1.group(2), 16), int(m2.group(1), 16), int(m2.group(2), 16)
    rng = np.random.default_rng(0)
    for _ in range(50):
        present = rng.integers(0, 2, (128, 4))                    # 128 pixels of a row x 4 classes
        # lane L holds pixels 4L..4L+3: byte c of t = their 4 presence bits of class c
        t = [sum(int(present[4 * L + j, c]) << (8 * c + j) for c in range(4) for j in range(4)) for L in range(32)]
        u = [(t[L] | (t[L ^ 1] << 4)) & 0xFFFFFFFF for L in range(32)]
        v = [u[L ^ 2] for L in range(32)]
        u = [byte_perm(u[L], v[L], r1_odd if L & 2 else r1_even) for L in range(32)]
        v = [u[L ^ 4] for L in range(32)]
        u = [byte_perm(u[L], v[L], r2_hi if L & 4 else r2_lo) for L in range(32)]
        for L in range(0, 32, 2):
            k, word = (L & 7) >> 1, L >> 3
            want = sum(int(present[32 * word + i, k]) << i for i in range(32))
            assert u[L] == want, (L, hex(u[L]), hex(want))
