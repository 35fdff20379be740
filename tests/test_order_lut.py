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
                pts.append((_side(yx, 11.0, 9.0), H[1], _side(yz, 9.0, 11.0), c, 4))  # kinds are one-hot: z 1, x 2, y 4
            want = _reference_order(pts)
            e = int(lut[idx])
            lo, hi = e & 0xFFFFFFFF, e >> 32
            got = [((lo >> (8 * i + 3)) & 7, (lo >> (8 * i)) & 7) for i in range(len(pts))]
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


def _prmt_compact(bytes_word, keep):
    """what PRMT does with the kernel's selector table: kept bytes move to the front in order, the rest read 0"""
    out, k = 0, 0
    for i in range(4):
        if (keep >> i) & 1:
            out |= ((bytes_word >> (8 * i)) & 0xFF) << (8 * k)
            k += 1
    return out, k


def test_sign_bit_index_and_keep_mask_reproduce_the_reference_on_random_nodes():
    """The kernel's common expansion case end to end on the CPU, in numpy fp32 with the kernel's own expressions
    (GLSL division a * (1/b)): twelve sign bits -> table -> keep mask -> compacted list, against the literal
    ordering loop of octree_fsh.c L251-330 on the same node, ray and child mask.  Both signs of the NaN that
    inf - inf produces (two invalid hits) are tried: the result may not depend on it."""
    lib = connector.load_library()
    lut = np.zeros(4096, dtype=np.uint64)
    lib.octree_cuc_debug_order_lut(lut.ctypes.data_as(ctypes.c_void_p))
    f = np.float32
    rng = np.random.default_rng(20261017)
    u = f(1800.0 / 4096.0)
    done = general = 0
    for case in range(6000):
        lvl = int(rng.integers(0, 12))
        su = 4096 >> lvl
        X, Y, Z = [int(rng.integers(0, 4096 // su)) * su for _ in range(3)]
        x0, y1, z1, sz = f(X) * u, f(Y + su) * u, f(Z + su) * u, f(su) * u
        x1, y0, z0 = f(x0 + sz), f(y1 - sz), f(z1 - sz)
        hs = f(sz * f(0.5))
        hx, hy, hz = f(x0 + hs), f(y1 - hs), f(z1 - hs)
        # a ray through the cube: origin outside, aimed at a point inside (sometimes exactly at the centre / on a
        # mid plane to provoke ties and on-plane hits)
        o = (rng.uniform(-600, 2400, 3)).astype(f)
        tgt = np.array([rng.uniform(x0, x1), rng.uniform(y0, y1), rng.uniform(z0, z1)], dtype=f)
        mode = rng.integers(0, 6)
        if mode == 0:
            tgt = np.array([hx, hy, hz], dtype=f)
        elif mode == 1:
            tgt[int(rng.integers(0, 3))] = (hx, hy, hz)[int(rng.integers(0, 3))]
        d = (tgt - o).astype(f)
        if mode == 2:
            d[int(rng.integers(0, 3))] = f(0.0)  # axis-parallel: 1/0 = inf
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            r = (f(1.0) / d).astype(f)
            ox, oy, oz = o
            dx, dy, dz = d
            # entry point: where the ray meets the cube's near face along its dominant axis (any point with
            # a finite w works for the ordering logic), or the origin with w = 0
            ew = f(rng.uniform(0.0, 1.0)) if rng.random() < 0.8 else f(0.0)
            ex, ey, ez = f(ox + dx * ew), f(oy + dy * ew), f(oz + dz * ew)
            wz, wx, wy = f(f(hz - oz) * r[2]), f(f(hx - ox) * r[0]), f(f(hy - oy) * r[1])
            zx, zy = f(ox + f(dx * wz)), f(oy + f(dy * wz))
            xy, xz = f(oy + f(dy * wx)), f(oz + f(dz * wx))
            yx, yz = f(ox + f(dx * wy)), f(oz + f(dz * wy))
            vz = bool(wz > 0 and x0 < zx and zx <= x1 and y1 > zy and zy >= y0)
            vx = bool(wx > 0 and y1 > xy and xy >= y0 and z1 > xz and xz >= z0)
            vy = bool(wy > 0 and x0 < yx and yx <= x1 and z1 > yz and yz >= z0)
            INFF = f(np.inf)
            mz, mx, my = (wz if vz else INFF), (wx if vx else INFF), (wy if vy else INFF)
            if mz < ew or mx < ew or my < ew or zx == hx or zy == hy or yx == hx or wz == wx:
                general += 1
                continue
            mask = int(rng.integers(0, 256))
            # reference
            pts = [(ex, ey, ez, ew, 0)]
            if vz:
                pts.append((zx, zy, hz, wz, 1))
            if vx:
                pts.append((hx, xy, xz, wx, 2))
            if vy:
                pts.append((yx, hy, yz, wy, 4))  # one-hot kinds: z 1, x 2, y 4
            hitp = [list(p) for p in pts]
            want, pre = [], -1
            for i in range(len(hitp)):
                for j in range(i + 1, len(hitp)):
                    if hitp[j][3] < hitp[i][3]:
                        hitp[i], hitp[j] = hitp[j], hitp[i]
                px, py, pz, _, kind = hitp[i]
                oc = (1 if px > hx else 0) + (2 if py < hy else 0) + (4 if pz < hz else 0)
                if oc == pre:
                    if px == hx:
                        oc ^= 1
                    elif py == hy:
                        oc ^= 2
                    elif pz == hz:
                        oc ^= 4
                pre = oc
                if (mask >> oc) & 1:
                    want.append((kind, oc))
            # kernel
            diffs = [hx - ex, ey - hy, ez - hz, hx - zx, zy - hy, xy - hy, xz - hz, hx - yx, yz - hz, mx - mz,
                     my - mz, my - mx]
            for nan_sign in (0, 1):
                idx = 0
                for v in diffs:
                    v = f(v)
                    bit = nan_sign if np.isnan(v) else int(np.signbit(v))
                    idx = (idx << 1) | bit
                e = int(lut[idx])
                lo, hi = e & 0xFFFFFFFF, e >> 32
                inval = (0 if vz else 8) + (0 if vx else 8) + (0 if vy else 8)
                hits = hi & ((mask * 0x01010101) & 0xFFFFFFFF)
                flags = ((hits + 0x7F7F7F7F) & 0xFFFFFFFF) & (0x80808080 >> inval)
                keep = (((flags * 0x00204081) & 0xFFFFFFFF) >> 26) >> 2
                lst, n = _prmt_compact(lo, keep)
                assert n == bin(flags).count("1")
                got = [((lst >> (8 * i + 3)) & 7, (lst >> (8 * i)) & 7) for i in range(n)]
                assert got == want, (case, nan_sign, got, want)
            done += 1
    assert done > 3000 and general > 100, (done, general)
