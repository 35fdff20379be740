"""Synthetic scenes in the reference's data formats, and the host-side data model.

The reference's assets (media/*.bin, res/*) are not in the tree, so every input
is generated: seeded point clouds on watertight surface grids (spacing <= 0.30 in
a cube of 1800 units, depth 12 -> 0.44-unit leaves), voxelised with the rules of
qmc (qmc.c L62-83, L259, L291-327) and inserted into 12-int octrees with the
rules of octree_insert_point (octree.c L95-147) -- both implemented in
host/qb_host.c and checked against the reference's compiled code by
tests/test_host_model.py.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .build import build_host, lib_paths

BASESIZE = 1800.0  # qubatron.c L573
LEVELS = 12        # qubatron.c L568

_host = None


def host_lib():
    global _host
    if _host is not None:
        return _host
    path = lib_paths()["host"]
    if not os.path.exists(path):
        build_host()
    lib = C.CDLL(path)
    lib.qb_octree_create.restype = C.c_void_p
    lib.qb_octree_create.argtypes = [C.c_float, C.c_int]
    lib.qb_octree_delete.argtypes = [C.c_void_p]
    lib.qb_octree_reset.argtypes = [C.c_void_p]
    lib.qb_octree_len.restype = C.c_int64
    lib.qb_octree_len.argtypes = [C.c_void_p]
    lib.qb_octree_nodes.restype = C.POINTER(C.c_int32)
    lib.qb_octree_nodes.argtypes = [C.c_void_p]
    lib.qb_octree_insert_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.qb_octree_insert_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    lib.qb_octree_insert_paths.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    lib.qb_octree_remove_point.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qb_octree_adopt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.qb_upload_node_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.qb_voxelise_order.restype = C.c_int64
    lib.qb_voxelise_order.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_gather_f3.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    _host = lib
    return lib


class HostOctree:
    """Growable 12-int node array (octree.c L11-23) with the reference's insert / remove rules."""

    def __init__(self, basesize=BASESIZE, levels=LEVELS):
        self.lib = host_lib()
        self.h = self.lib.qb_octree_create(float(basesize), int(levels))
        self.basesize, self.levels = float(basesize), int(levels)

    def __len__(self):
        return int(self.lib.qb_octree_len(self.h))

    def reset(self):
        self.lib.qb_octree_reset(self.h)

    def insert_points(self, pts, first_modind=0):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        if len(pts) and (pts.min() < 0 or pts.max() >= self.basesize):
            raise ValueError("points must lie inside the base cube [0, basesize)")
        self.lib.qb_octree_insert_points(self.h, pts.ctypes.data_as(C.c_void_p), len(pts), int(first_modind))

    def insert_point(self, pnt, modind):
        """Returns the touched-node list (octindarr of octree.c L113-119, 13 entries)."""
        p = np.ascontiguousarray(pnt, dtype=np.float32)
        touched = np.zeros(13, dtype=np.int32)
        self.lib.qb_octree_insert_point(self.h, p.ctypes.data_as(C.c_void_p), int(modind),
                                        touched.ctypes.data_as(C.c_void_p))
        return touched

    def adopt(self, nodes):
        """Take over an existing int32 [n,12] node array (e.g. a level's flat file) instead of re-inserting."""
        nodes = np.ascontiguousarray(nodes, dtype=np.int32).reshape(-1, 12)
        self.lib.qb_octree_adopt(self.h, nodes.ctypes.data_as(C.c_void_p), len(nodes))

    def upload_node_ranges(self, rc, node_index, buftype):
        """The engine's per-node upload loop (modelutil.c L429-437) in C: one octree_glc_upload_texbuffer_data call
        per touched node through the connector `rc` (an OctreeGlc)."""
        idx = np.ascontiguousarray(node_index, dtype=np.int32)
        fn = C.cast(rc.lib.octree_glc_upload_texbuffer_data, C.c_void_p)
        self.lib.qb_upload_node_ranges(fn, C.cast(rc._p, C.c_void_p), self.h, idx.ctypes.data_as(C.c_void_p), len(idx),
                                       int(buftype))

    def insert_paths(self, paths, first_modind=0):
        paths = np.ascontiguousarray(paths, dtype=np.int32).reshape(-1, 12)
        self.lib.qb_octree_insert_paths(self.h, paths.ctypes.data_as(C.c_void_p), len(paths), int(first_modind))

    def remove_point(self, pnt):
        p = np.ascontiguousarray(pnt, dtype=np.float32)
        m, o = C.c_int32(-1), C.c_int32(-1)
        self.lib.qb_octree_remove_point(self.h, p.ctypes.data_as(C.c_void_p), C.byref(m), C.byref(o))
        return m.value, o.value

    def nodes(self, copy=True):
        """int32 [len,12] view (or copy) of the node array -- what gets uploaded."""
        n = len(self)
        arr = np.ctypeslib.as_array(self.lib.qb_octree_nodes(self.h), shape=(n, 12))
        return arr.copy() if copy else arr

    def __del__(self):
        try:
            self.lib.qb_octree_delete(self.h)
        except Exception:
            pass


def voxelise(pos, col_u8, nrm, size=1800, levels=LEVELS):
    """qmc: returns (pnt, col, nrm) float32 [m,3] sorted x-major, one point per 2^(levels+1) grid cell."""
    lib = host_lib()
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    n = len(pos)
    order = np.empty(n, dtype=np.int64)
    dropped = C.c_int64(0)
    m = lib.qb_voxelise_order(pos.ctypes.data_as(C.c_void_p), n, int(size), int(levels),
                              order.ctypes.data_as(C.c_void_p), C.byref(dropped))
    order = order[:m]
    out_p = np.empty((m, 3), dtype=np.float32)
    lib.qb_gather_f3(pos.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p), m,
                     out_p.ctypes.data_as(C.c_void_p))
    nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
    out_n = np.empty((m, 3), dtype=np.float32)
    lib.qb_gather_f3(nrm.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p), m,
                     out_n.ctypes.data_as(C.c_void_p))
    # qmc.c L52-54: colour = uchar / 255.0 evaluated in double, stored as float
    col_u8 = np.ascontiguousarray(col_u8, dtype=np.uint8).reshape(-1, 3)
    out_c = (col_u8[order].astype(np.float64) / 255.0).astype(np.float32)
    return out_p, out_c, out_n


def flat_ranges(pnt, size=1800, levels=LEVELS):
    """The .rng stream of qmc (qmc.c L291-313): for every surviving point whose (x, y) grid column differs from the
    previous point's -- the point before the first one counts as column (0, 0) -- one int32 triple
    (index of the point, column x, column y).  Grid index = floor(p / precision) in fp32 (qmc.c L62-68, L217-218)."""
    pnt = np.ascontiguousarray(pnt, dtype=np.float32).reshape(-1, 3)
    division = 2 << int(levels)
    precision = np.float32(size) / np.float32(division)
    ix = np.floor(pnt[:, 0] / precision).astype(np.int32)
    iy = np.floor(pnt[:, 1] / precision).astype(np.int32)
    px = np.concatenate([[0], ix[:-1]]).astype(np.int32)
    py = np.concatenate([[0], iy[:-1]]).astype(np.int32)
    first = np.nonzero((ix != px) | (iy != py))[0]
    return np.stack([first.astype(np.int32), ix[first], iy[first]], axis=1).astype(np.int32)


def write_flat(prefix, pnt, nrm, col, size=1800, levels=LEVELS):
    """The voxeliser's four output files <prefix>.pnt / .nrm / .col / .rng exactly as qmc writes them (qmc.c
    L266-327: raw float32 x y z per point; .rng see flat_ranges) -- what model_load_flat (model.c L53-111) reads."""
    np.ascontiguousarray(pnt, dtype=np.float32).tofile(prefix + ".pnt")
    np.ascontiguousarray(nrm, dtype=np.float32).tofile(prefix + ".nrm")
    np.ascontiguousarray(col, dtype=np.float32).tofile(prefix + ".col")
    flat_ranges(pnt, size, levels).tofile(prefix + ".rng")


def load_flat(prefix):
    """model_load_flat (model.c L53-111): (pnt, col, nrm float32 [n,3], ranges int32 [r,3]); the point count comes
    from the size of the .pnt file, as in the reference."""
    pnt = np.fromfile(prefix + ".pnt", dtype=np.float32).reshape(-1, 3)
    nrm = np.fromfile(prefix + ".nrm", dtype=np.float32).reshape(-1, 3)
    col = np.fromfile(prefix + ".col", dtype=np.float32).reshape(-1, 3)
    rng = np.fromfile(prefix + ".rng", dtype=np.int32).reshape(-1, 3)
    if not (len(pnt) == len(nrm) == len(col)):
        raise ValueError("flat model %s: .pnt / .nrm / .col hold different point counts" % prefix)
    return pnt, col, nrm, rng


@dataclass
class Scene:
    name: str
    pnt_s: np.ndarray
    col_s: np.ndarray
    nrm_s: np.ndarray
    oct_s: np.ndarray
    pnt_d: np.ndarray
    col_d: np.ndarray
    nrm_d: np.ndarray
    oct_d: np.ndarray
    basesize: float = BASESIZE
    levels: int = LEVELS
    raw_static: int = 0
    raw_dynamic: int = 0
    cameras: list = None  # [(position, angle), ...] poses the configs are measured from

    def describe(self):
        return {"name": self.name, "static_points_raw": int(self.raw_static), "static_points": int(len(self.pnt_s)),
                "static_nodes": int(len(self.oct_s)), "dynamic_points_raw": int(self.raw_dynamic),
                "dynamic_points": int(len(self.pnt_d)), "dynamic_nodes": int(len(self.oct_d)),
                "basesize": self.basesize, "levels": self.levels}


def _empty3():
    return np.zeros((0, 3), dtype=np.float32)


def build_scene(name, static_raw, dynamic_raw=None, basesize=BASESIZE, levels=LEVELS, voxelised=True):
    """static_raw / dynamic_raw = (pos f32[n,3], col u8[n,3], nrm f32[n,3]) or None."""
    def one(raw):
        if raw is None:
            t = HostOctree(basesize, levels)
            return _empty3(), _empty3(), _empty3(), t.nodes(), 0
        pos, col, nrm = raw
        n_raw = len(pos)
        if voxelised:
            p, c, n = voxelise(pos, col, nrm, int(basesize), levels)
        else:
            p = np.ascontiguousarray(pos, dtype=np.float32)
            c = (np.asarray(col, dtype=np.float64) / 255.0).astype(np.float32)
            n = np.ascontiguousarray(nrm, dtype=np.float32)
        t = HostOctree(basesize, levels)
        t.insert_points(p)
        return p, c, n, t.nodes(), n_raw
    ps, cs, ns, os_, rs = one(static_raw)
    pd, cd, nd, od, rd = one(dynamic_raw)
    return Scene(name, ps, cs, ns, os_, pd, cd, nd, od, float(basesize), int(levels), rs, rd)


# ---------------------------------------------------------------------------
# surface samplers: regular parameter grids, so surfaces are watertight at the
# 0.44-unit leaf size (random sampling leaves holes)
# ---------------------------------------------------------------------------

def _colour(pos, base, seed):
    """procedural uchar colours: base tint modulated by a position checker."""
    p = pos.astype(np.float32)
    chk = ((np.floor(p[:, 0] / 8.0) + np.floor(p[:, 1] / 8.0) + np.floor(p[:, 2] / 8.0)).astype(np.int64) & 1)
    k = (0.65 + 0.35 * chk)[:, None] * np.asarray(base, dtype=np.float32)[None, :]
    rng = np.random.default_rng(seed)
    k = k + rng.integers(-6, 7, size=k.shape).astype(np.float32)
    return np.clip(k, 0, 255).astype(np.uint8)


def _rect(origin, eu, ev, nu, nv, normal, spacing):
    """points origin + i*spacing*eu + j*spacing*ev on a nu x nv unit-area patch (sizes in units)."""
    iu = np.arange(0.0, nu, spacing, dtype=np.float32)
    iv = np.arange(0.0, nv, spacing, dtype=np.float32)
    U, Vv = np.meshgrid(iu, iv, indexing="ij")
    pos = (np.asarray(origin, np.float32)[None, :] + U.reshape(-1, 1) * np.asarray(eu, np.float32)[None, :]
           + Vv.reshape(-1, 1) * np.asarray(ev, np.float32)[None, :])
    nrm = np.broadcast_to(np.asarray(normal, np.float32), pos.shape).copy()
    return pos.astype(np.float32), nrm


def _sphere(centre, r, spacing):
    nth = max(int(np.pi * r / spacing), 4)
    th = (np.arange(nth, dtype=np.float32) + 0.5) * np.float32(np.pi / nth)
    parts = []
    for t in th:
        nph = max(int(2 * np.pi * r * np.sin(t) / spacing), 3)
        ph = np.arange(nph, dtype=np.float32) * np.float32(2 * np.pi / nph)
        d = np.stack([np.sin(t) * np.cos(ph), np.full_like(ph, np.cos(t)), np.sin(t) * np.sin(ph)], axis=1)
        parts.append(d)
    d = np.concatenate(parts).astype(np.float32)
    return (np.asarray(centre, np.float32)[None, :] + r * d).astype(np.float32), d


def _box(lo, hi, spacing, with_bottom=False):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    sx, sy, sz = (hi - lo)
    P, N = [], []
    faces = [
        ((lo[0], lo[1], lo[2]), (1, 0, 0), (0, 1, 0), sx, sy, (0, 0, -1)),
        ((lo[0], lo[1], hi[2]), (1, 0, 0), (0, 1, 0), sx, sy, (0, 0, 1)),
        ((lo[0], lo[1], lo[2]), (0, 0, 1), (0, 1, 0), sz, sy, (-1, 0, 0)),
        ((hi[0], lo[1], lo[2]), (0, 0, 1), (0, 1, 0), sz, sy, (1, 0, 0)),
        ((lo[0], hi[1], lo[2]), (1, 0, 0), (0, 0, 1), sx, sz, (0, 1, 0)),
    ]
    if with_bottom:
        faces.append(((lo[0], lo[1], lo[2]), (1, 0, 0), (0, 0, 1), sx, sz, (0, -1, 0)))
    for o, eu, ev, nu, nv, nr in faces:
        p, n = _rect(o, eu, ev, nu, nv, nr, spacing)
        P.append(p)
        N.append(n)
    return np.concatenate(P), np.concatenate(N)


def _capsule(a, b, r, spacing, shells=1, shell_gap=0.25):
    """shells concentric cylinder surfaces between a and b plus end caps (y-up limbs)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    axis = b - a
    h = float(np.linalg.norm(axis))
    w = axis / h
    t = np.array([1, 0, 0], np.float32) if abs(w[0]) < 0.9 else np.array([0, 1, 0], np.float32)
    e1 = np.cross(w, t)
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(w, e1)
    P, N = [], []
    for s in range(shells):
        rr = r - s * shell_gap
        if rr <= spacing:
            break
        nph = max(int(2 * np.pi * rr / spacing), 6)
        ph = np.arange(nph, dtype=np.float32) * np.float32(2 * np.pi / nph)
        hh = np.arange(0.0, h, spacing, dtype=np.float32)
        PH, HH = np.meshgrid(ph, hh, indexing="ij")
        d = np.cos(PH).reshape(-1, 1) * e1[None, :] + np.sin(PH).reshape(-1, 1) * e2[None, :]
        P.append(a[None, :] + HH.reshape(-1, 1) * w[None, :] + rr * d)
        N.append(d)
        for c in (a, b):
            sp, sn = _sphere(c, rr, spacing)
            P.append(sp)
            N.append(sn)
    return np.concatenate(P).astype(np.float32), np.concatenate(N).astype(np.float32)


def _inside(pos, size):
    return np.all((pos >= 0.0) & (pos < size - 0.25), axis=1)


def _pack(parts, size=BASESIZE):
    pos = np.concatenate([p for p, _, _ in parts]).astype(np.float32)
    col = np.concatenate([c for _, c, _ in parts]).astype(np.uint8)
    nrm = np.concatenate([n for _, _, n in parts]).astype(np.float32)
    keep = _inside(pos, size)
    return pos[keep], col[keep], nrm[keep]


# cameras: (position, angle) -- C1 is the reference's start pose (qubatron.c L579)
CAMERA_C1 = ((700.0, 150.0, 350.0), (0.4636, 0.0, 0.0))


def make_c1(spacing=0.30, seed=1234):
    """Config C1: ~1 M raw points, a small room in front of the reference's start pose."""
    parts = []

    def add(p, n, base, k):
        parts.append((p, _colour(p, base, seed + k), n))

    p, n = _rect((650, 100, 120), (1, 0, 0), (0, 0, 1), 200, 200, (0, 1, 0), spacing)  # floor
    add(p, n, (200, 190, 170), 0)
    p, n = _rect((650, 100, 120), (1, 0, 0), (0, 1, 0), 200, 80, (0, 0, 1), spacing)   # back wall
    add(p, n, (170, 120, 100), 1)
    p, n = _rect((850, 100, 120), (0, 0, 1), (0, 1, 0), 200, 80, (-1, 0, 0), spacing)  # side wall
    add(p, n, (110, 140, 180), 2)
    p, n = _sphere((760, 125, 220), 25.0, spacing)
    add(p, n, (220, 80, 60), 3)
    p, n = _box((700, 100, 160), (730, 130, 190), spacing)
    add(p, n, (90, 200, 110), 4)
    return build_scene("c1_room_1M", _pack(parts))


def zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.2, shells=16, seed=4321):
    """A ~163-unit capsule-limb figure (skeleton_glc.c L146-154 proportions)."""
    bx, by, bz = base
    limbs = [
        ((bx - 12, by + 2, bz), (bx - 12, by + 75, bz), 10.0),       # legs
        ((bx + 12, by + 2, bz), (bx + 12, by + 75, bz), 10.0),
        ((bx, by + 78, bz), (bx, by + 132, bz), 21.0),               # torso
        ((bx - 30, by + 70, bz), (bx - 30, by + 126, bz + 6), 7.0),  # arms
        ((bx + 30, by + 70, bz), (bx + 30, by + 126, bz + 6), 7.0),
        ((bx, by + 146, bz), (bx, by + 150, bz), 13.0),              # head
    ]
    parts = []
    for k, (a, b, r) in enumerate(limbs):
        p, n = _capsule(a, b, r, spacing, shells=shells)
        parts.append((p, _colour(p, (120, 160, 110), seed + k), n))
    return _pack(parts)


def zombie_bones(base=(760.0, 100.0, 230.0), pose=1.0, shift=(0.0, 0.0, 0.0)):
    """Twenty joints (ten bone pairs, zombie.c L110-131 order: head-neck, neck-hip, two arms and two legs in two
    segments each) for zombie_raw's figure.  Returns (oribones, newbones), float32 [20, 4]: oribones.w = radius of
    effect of the pair's first joint, newbones.w = twist about the bone (skeleton_vsh.c L113, L137).
    pose = 0 is the rest pose, larger values bend the limbs further; shift moves the whole figure."""
    bx, by, bz = base
    j = dict(
        head=(bx, by + 158, bz), neck=(bx, by + 132, bz), hip=(bx, by + 78, bz),
        rshol=(bx - 30, by + 126, bz + 6), relbo=(bx - 30, by + 98, bz + 3), rhand=(bx - 30, by + 70, bz),
        lshol=(bx + 30, by + 126, bz + 6), lelbo=(bx + 30, by + 98, bz + 3), lhand=(bx + 30, by + 70, bz),
        rleg=(bx - 12, by + 75, bz), rknee=(bx - 12, by + 38, bz), rfoot=(bx - 12, by + 2, bz),
        lleg=(bx + 12, by + 75, bz), lknee=(bx + 12, by + 38, bz), lfoot=(bx + 12, by + 2, bz))
    order = ["head", "neck", "neck", "hip", "rshol", "relbo", "relbo", "rhand", "lshol", "lelbo", "lelbo", "lhand",
             "rleg", "rknee", "rknee", "rfoot", "lleg", "lknee", "lknee", "lfoot"]
    effect = dict(head=20.0, neck=30.0, hip=30.0, rshol=12.0, relbo=12.0, rhand=12.0, lshol=12.0, lelbo=12.0,
                  lhand=12.0, rleg=16.0, rknee=16.0, rfoot=16.0, lleg=16.0, lknee=16.0, lfoot=16.0)
    ori = np.array([j[k] + (effect[k],) for k in order], dtype=np.float32)
    p = np.float32(pose)
    moved = {k: np.array(v, dtype=np.float32) for k, v in j.items()}
    moved["rhand"] += np.float32([0.0, 6.0, 14.0]) * p      # right forearm swings forward
    moved["lelbo"] += np.float32([4.0, 0.0, -8.0]) * p      # left arm swings back
    moved["lhand"] += np.float32([7.0, 3.0, -17.0]) * p
    moved["rknee"] += np.float32([0.0, 2.0, 12.0]) * p      # right leg steps
    moved["rfoot"] += np.float32([0.0, 5.0, 6.0]) * p
    moved["lfoot"] += np.float32([0.0, 3.0, -10.0]) * p
    moved["head"] += np.float32([3.0, 0.0, 2.0]) * p        # head tilts
    twist = dict(head=0.15, neck=0.0, hip=0.0, rshol=0.1, relbo=0.2, lshol=0.0, lelbo=-0.1, rleg=0.0, rknee=0.05,
                 lleg=0.0, lknee=0.0, rhand=0.0, lhand=0.0, rfoot=0.0, lfoot=0.0)
    sh = np.float32(shift)
    new = np.array([tuple(moved[k] + sh) + (np.float32(twist[k]) * p,) for k in order], dtype=np.float32)
    return ori, new


def _terrain_height(x, z):
    return (60.0 + 22.0 * np.sin(x / 140.0) * np.cos(z / 170.0) + 6.0 * np.sin(x / 23.0 + z / 31.0)).astype(np.float32)


def make_c2(scale=1.0, spacing=0.30, seed=1234, zombie=True, progress=None):
    """Config C2: 'abandoned-scale' level.  scale=1.0 -> ~90 M raw static points
    (terrain over the whole cube + building shells) and a ~10 M-point figure;
    scale<1 shrinks the covered area (same spacing, same density) for tests."""
    S = BASESIZE
    ext = float(S) * float(np.sqrt(scale)) if scale < 1.0 else float(S)
    # keep the C1 camera region inside the covered area
    x_lo = max(0.0, min(600.0, S - ext))
    z_lo = max(0.0, min(100.0, S - ext))
    x_hi, z_hi = min(S - 0.5, x_lo + ext), min(S - 0.5, z_lo + ext)
    parts = []
    # terrain in strips to bound temporary memory
    xs = np.arange(x_lo, x_hi, spacing, dtype=np.float32)
    zs = np.arange(z_lo, z_hi, spacing, dtype=np.float32)
    strip = 512
    for i in range(0, len(xs), strip):
        X, Zz = np.meshgrid(xs[i:i + strip], zs, indexing="ij")
        X, Zz = X.reshape(-1), Zz.reshape(-1)
        Yy = _terrain_height(X, Zz)
        pos = np.stack([X, Yy, Zz], axis=1)
        # analytic normal of the height field
        dhx = (22.0 / 140.0) * np.cos(X / 140.0) * np.cos(Zz / 170.0) + (6.0 / 23.0) * np.cos(X / 23.0 + Zz / 31.0)
        dhz = -(22.0 / 170.0) * np.sin(X / 140.0) * np.sin(Zz / 170.0) + (6.0 / 31.0) * np.cos(X / 23.0 + Zz / 31.0)
        nrm = np.stack([-dhx, np.ones_like(dhx), -dhz], axis=1).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        parts.append((pos, _colour(pos, (150, 170, 120), seed), nrm))
        if progress:
            progress("terrain strip %d/%d" % (i // strip + 1, (len(xs) + strip - 1) // strip))
    # building shells on a jittered grid, skipping the camera / figure corridor
    rng = np.random.default_rng(seed)
    pitch = 174.0
    k = 0
    inside = (0.0, 0.0, 0.0)
    for gx in np.arange(x_lo + 40.0, x_hi - 130.0, pitch):
        for gz in np.arange(z_lo + 40.0, z_hi - 130.0, pitch):
            bx = float(gx + rng.uniform(0, 30))
            bz = float(gz + rng.uniform(0, 30))
            w, d, h = float(rng.uniform(90, 125)), float(rng.uniform(90, 125)), float(rng.uniform(70, 130))
            # keep a clear corridor around the C1 camera and where it looks
            if bx < 900.0 and bx + w > 640.0 and bz < 420.0 and bz + d > 150.0:
                continue
            y0 = float(_terrain_height(np.float32(bx), np.float32(bz))) - 8.0
            p, n = _box((bx, y0, bz), (bx + w, y0 + h, bz + d), spacing)
            tint = (int(rng.integers(110, 220)), int(rng.integers(100, 200)), int(rng.integers(90, 190)))
            parts.append((p, _colour(p, tint, seed + 100 + k), n))
            if k == 0 or (abs(bx - 1000.0) + abs(bz - 600.0) < abs(inside[0] - 1000.0) + abs(inside[2] - 600.0)):
                inside = (bx + 0.5 * w, y0 + 0.45 * h, bz + 0.5 * d)
            k += 1
    if progress:
        progress("%d buildings" % k)
    # a few large spheres
    for cx, cy, cz, r in ((820.0, 120.0, 200.0, 28.0), (980.0, 140.0, 520.0, 60.0), (400.0, 150.0, 900.0, 80.0)):
        if x_lo <= cx <= x_hi and z_lo <= cz <= z_hi:
            p, n = _sphere((cx, cy, cz), r, spacing)
            parts.append((p, _colour(p, (210, 90, 70), seed + 7), n))
    static_raw = _pack(parts)
    del parts
    dyn = None
    if zombie:
        by = float(_terrain_height(np.float32(760.0), np.float32(230.0)))
        shells = 18 if scale >= 1.0 else max(2, int(18 * scale) + 1)
        dyn = zombie_raw(base=(760.0, by, 230.0), spacing=0.2, shells=shells)
    name = "c2_abandoned_%dM" % round(len(static_raw[0]) / 1e6)
    sc = build_scene(name, static_raw, dyn)
    gy = float(_terrain_height(np.float32(1100.0), np.float32(700.0)))
    sc.cameras = [
        CAMERA_C1,                                               # reference start pose, figure in view
        ((inside[0], inside[1], inside[2]), (-2.438, -0.115, 0.0)),  # inside a building shell
        ((880.0, 260.0, 640.0), (2.6, 0.55, 0.0)),               # looking up: mostly open sky
        ((1100.0, gy + 2.5, 700.0), (-1.0, 0.0, 0.0)),           # grazing the terrain
    ]
    return sc


def make_test5():
    """The reference's only fixture: the 5-point OCTTEST scene (modelutil.c L89-110),
    inserted without qmc, camera at (900,900,3000) (qubatron.c L136)."""
    pos = np.array([[10, 690, 10], [10, 340, 10], [10, 340, 690], [10, 10, 10], [690, 10, 690]], dtype=np.float32)
    nrm = np.tile(np.array([[0, 0, -1]], dtype=np.float32), (5, 1))
    col = np.full((5, 3), 255, dtype=np.uint8)
    return build_scene("octtest_5pt", (pos, col, nrm), None, voxelised=False)


def make_random(n_static=20000, n_dynamic=3000, seed=7, levels=LEVELS, basesize=BASESIZE, clustered=True):
    """Fuzz scene: random clouds in both trees (overlapping region), arbitrary colours / non-unit normals."""
    rng = np.random.default_rng(seed)

    def cloud(n, centre, spread):
        if clustered:
            p = centre + rng.normal(0, spread, size=(n, 3))
        else:
            p = rng.uniform(0, basesize, size=(n, 3))
        p = np.clip(p, 0.5, basesize - 0.5).astype(np.float32)
        c = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
        nr = rng.normal(0, 1, size=(n, 3)).astype(np.float32)
        nr[np.abs(nr).sum(axis=1) == 0] = 1.0
        return p, c, nr
    s = cloud(n_static, np.array([basesize * 0.45, basesize * 0.1, basesize * 0.15]), basesize * 0.03)
    d = cloud(n_dynamic, np.array([basesize * 0.44, basesize * 0.11, basesize * 0.16]), basesize * 0.012) \
        if n_dynamic else None
    return build_scene("random_%d_%d" % (n_static, n_dynamic), s, d, basesize=basesize, levels=levels)


def octant_paths(points, basesize=BASESIZE, levels=LEVELS):
    """The 12 octant digits of every point (what skeleton_vsh.c L188-226 emits per skinned point and
    octree_insert_path consumes, qubatron.c L439-452): same arithmetic as octree_insert_point (octree.c L102-109)."""
    p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((len(p), 12), dtype=np.int32)
    size = np.float32(basesize)
    for level in range(levels):
        size = np.float32(size / np.float32(2.0))
        q = (p / size).astype(np.int32) % 2
        out[:, level] = q[:, 0] + np.where(q[:, 1] == 0, 2, 0) + np.where(q[:, 2] == 0, 4, 0)
    return out
