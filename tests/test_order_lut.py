"""The fast kernel's candidate-ordering table (octree_trace_fast.cuh, g_order_lut) against a literal
restatement of the reference's ordering loop (octree_fsh.c L273-330) on concrete points.  CPU only."""
import ctypes
import itertools

import numpy as np

from qubatron_b200 import connector

INF = float("inf")
H = (10.0, 10.0, 10.0)  # hlf


def _reference_order(points):
    """points: list of (x, y, z, w, kind) in the reference's candidate order (entry, z, x, y hits that passed the
    range tests).  Returns [(kind, octant)] as L273-330 produce them when every child exists."""
    hitp = [list(p) for p in points]
    out, pre = [], -1
    for i in range(len(hitp)):
        for j in range(i + 1, len(hitp)):
            if hitp[j][3] < hitp[i][3]:
                hitp[i], hitp[j] = hitp[j], hitp[i]
        x, y, z, _, kind = hitp[i]
        o = (1 if x > H[0] else 0) + (2 if y < H[1] else 0) + (4 if z < H[2] else 0)
        if o == pre:
            if x == H[0]:
                o ^= 1  # horpairs
            elif y == H[1]:
                o ^= 2  # verpairs
            elif z == H[2]:
                o ^= 4  # deppairs
        pre = o
        out.append((kind, o))
    return out


def _side(bit, hi, lo):
    return hi if bit else lo


def test_order_lut_equals_the_reference_ordering_loop():
    lib = connector.load_library()
    lut = np.zeros(4096, dtype=np.uint64)
    lib.octree_cuc_debug_order_lut(lut.ctypes.data_as(ctypes.c_void_p))
    checked = 0
    weights = (1.0, 2.0, 3.0, INF)
    for a, b, c in itertools.product(weights, repeat=3):
        if a == b and a != INF:
            continue  # the z/x tie is the kernel's general case
        # sign bits as the kernel takes them: inf - inf = NaN -> 0
        p1 = 1 if b < a else 0
        p2 = 1 if c < a else 0
        p3 = 1 if c < b else 0
        for bits in range(512):
            o0x, o0y, o0z, zx, zy, xy, xz, yx, yz = [(bits >> (8 - k)) & 1 for k in range(9)]
            idx = (bits << 3) | (p1 << 2) | (p2 << 1) | p3
            entry = (_side(o0x, 11.0, 9.0), _side(o0y, 9.0, 11.0), _side(o0z, 9.0, 11.0), 0.5, 0)
            pts = [entry]
            if a != INF:
                pts.append((_side(zx, 11.0, 9.0), _side(zy, 9.0, 11.0), H[2], a, 1))
            if b != INF:
                pts.append((H[0], _side(xy, 9.0, 11.0), _side(xz, 9.0, 11.0), b, 2))
            if c != INF:
                pts.append((_side(yx, 11.0, 9.0), H[1], _side(yz, 9.0, 11.0), c, 3))
            want = _reference_order(pts)
            e = int(lut[idx])
            lo, hi = e & 0xFFFFFFFF, e >> 32
            got = [((lo >> (8 * i + 3)) & 3, (lo >> (8 * i)) & 7) for i in range(len(pts))]
            assert got == want, (idx, a, b, c, got, want)
            for i, (_, o) in enumerate(want):
                assert (hi >> (8 * i)) & 0xFF == 1 << o
            checked += 1
    assert checked == (64 - 12) * 512


def test_keep_mask_arithmetic():
    """flags (bit 8i+7 = keep candidate i) -> byte offset of the compaction selector, as the kernel computes it."""
    for keep in range(16):
        flags = sum(0x80 << (8 * i) for i in range(4) if (keep >> i) & 1)
        assert ((flags * 0x00204081) & 0xFFFFFFFF) >> 26 == keep * 4
    # a byte of `hits` is zero or ONE bit (one-hot octant AND child mask), so the byte-wise add cannot carry
    one_hot = [0] + [1 << k for k in range(8)]
    for b in itertools.product(one_hot, repeat=4):
        hits = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)
        f = ((hits + 0x7F7F7F7F) & 0xFFFFFFFF) & 0x80808080
        assert f == sum(0x80 << (8 * i) for i in range(4) if b[i])
