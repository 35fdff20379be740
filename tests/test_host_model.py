"""Host data model (qubatron_b200/host/qb_host.c) against the reference's own compiled code.

oracle/_ref/libqubatron_ref.so is octree.c built unmodified from /root/reference
(oracle/Makefile); oracle/_ref/qmc is the reference voxeliser.  When oracle/_ref is
absent (a box without the prebuilt files) the reference comparisons skip and the
self-consistency checks still run."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import qb_oracle as O
from qubatron_b200 import scene as S

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference)")


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(1.0, 1799.0, size=(n, 3)).astype(np.float32)


@needs_ref
def test_insert_points_equals_reference():
    """octree_insert_point (octree.c L95-147): node-for-node identical arrays."""
    pts = _cloud(50000, 1)
    mine = S.HostOctree()
    mine.insert_points(pts)
    ref = O.RefOctree()
    ref.insert_points(pts)
    assert len(mine) == len(ref)
    assert np.array_equal(mine.nodes(), ref.nodes())


@needs_ref
def test_insert_point_touched_list_equals_reference():
    """octindarr (octree.c L113-119, used by modelutil.c L486-501)."""
    pts = _cloud(300, 2)
    mine, ref = S.HostOctree(), O.RefOctree()
    for i, p in enumerate(pts):
        assert np.array_equal(mine.insert_point(p, i), ref.insert_point(p, i))
    assert np.array_equal(mine.nodes(), ref.nodes())


@needs_ref
def test_remove_point_equals_reference():
    """octree_remove_point (octree.c L182-218): zeroes the parent's slot, reports (model, parent)."""
    pts = _cloud(5000, 3)
    mine, ref = S.HostOctree(), O.RefOctree()
    mine.insert_points(pts)
    ref.insert_points(pts)
    rng = np.random.default_rng(4)
    probes = np.concatenate([pts[rng.integers(0, len(pts), 200)], _cloud(200, 5)])
    for p in probes:
        assert mine.remove_point(p) == ref.remove_point(p)
    assert np.array_equal(mine.nodes(), ref.nodes())


@needs_ref
def test_insert_paths_equals_reference():
    """octree_insert_path (octree.c L149-180), the per-frame dynamic rebuild of qubatron.c L439-452."""
    rng = np.random.default_rng(6)
    paths = rng.integers(0, 8, size=(4000, 12)).astype(np.int32)
    mine, ref = S.HostOctree(), O.RefOctree()
    mine.insert_paths(paths)
    ref.insert_paths(paths)
    assert np.array_equal(mine.nodes(), ref.nodes())
    mine.reset()
    ref.reset()
    assert len(mine) == len(ref) == 1


def _write_ply(path, pos, col, nrm):
    n = len(pos)
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty float nx\n"
           "property float ny\nproperty float nz\nend_header\n" % n)
    rec = np.zeros(n, dtype=[("p", "<f4", 3), ("c", "u1", 3), ("n", "<f4", 3)])
    rec["p"], rec["c"], rec["n"] = pos, col, nrm
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(rec.tobytes())


@pytest.mark.skipif(not os.path.exists(O.REF_QMC), reason="oracle/_ref/qmc not built")
def test_voxelise_equals_reference_qmc():
    """qmc -s 1800 -l 12 (qmc.c) on a PLY of the same points: identical .pnt/.nrm/.col streams."""
    rng = np.random.default_rng(8)
    # surface-like cloud with many duplicates per grid cell and some out-of-cube points
    pos = (np.array([700.0, 100.0, 300.0]) + rng.uniform(0, 40, size=(60000, 3)) * np.array([1, 0.02, 1])).astype(
        np.float32)
    pos[:50] = rng.uniform(-50, 1900, size=(50, 3)).astype(np.float32)
    col = rng.integers(0, 256, size=(len(pos), 3)).astype(np.uint8)
    nrm = rng.normal(size=(len(pos), 3)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        _write_ply(os.path.join(d, "t.ply"), pos, col, nrm)
        r = subprocess.run([O.REF_QMC, "-s", "1800", "-l", "12", "-i", "t.ply"], cwd=d, stdout=subprocess.DEVNULL)
        assert r.returncode == 0
        rp = np.fromfile(os.path.join(d, "t.ply.pnt"), dtype=np.float32).reshape(-1, 3)
        rn = np.fromfile(os.path.join(d, "t.ply.nrm"), dtype=np.float32).reshape(-1, 3)
        rc = np.fromfile(os.path.join(d, "t.ply.col"), dtype=np.float32).reshape(-1, 3)
        rr = open(os.path.join(d, "t.ply.rng"), "rb").read()
        p, c, n = S.voxelise(pos, col, nrm)
        assert len(p) == len(rp)
        assert np.array_equal(p, rp) and np.array_equal(n, rn) and np.array_equal(c, rc)
        # the flat-file writer (qmc.c L266-327) and reader (model.c L53-111): all four files byte for byte, incl.
        # the column ranges of the .rng file
        S.write_flat(os.path.join(d, "mine"), p, n, c)
        for ext in ("pnt", "nrm", "col", "rng"):
            assert open(os.path.join(d, "mine." + ext), "rb").read() == open(os.path.join(d, "t.ply." + ext), "rb").read(), ext
        assert len(rr) % 12 == 0 and len(rr) > 12 * 50
        lp, lc, ln, lr = S.load_flat(os.path.join(d, "t.ply"))
        assert np.array_equal(lp, p) and np.array_equal(lc, c) and np.array_equal(ln, n)
        assert lr[:, 0].max() < len(p) and np.all(np.diff(lr[:, 0]) > 0)


def test_octree_invariants():
    """Structure rules of octree.c L11-23: root = node 0, children have larger indices, every node
    created by point p stores oct[8] = p, padding is zero, leaves are at depth `levels`."""
    sc = S.make_random(5000, 0, seed=3)
    nodes = sc.oct_s
    ch = nodes[:, :8]
    idx = np.arange(len(nodes))[:, None]
    assert ((ch == 0) | (ch > idx)).all()
    assert (nodes[:, 9:] == 0).all()
    assert nodes[:, 8].max() < len(sc.pnt_s)
    # every non-root node is referenced exactly once
    ref = np.bincount(ch[ch > 0].ravel(), minlength=len(nodes))
    assert (ref[1:] == 1).all() and ref[0] == 0
    # depth of leaves
    depth = np.zeros(len(nodes), dtype=np.int32)
    for i in range(len(nodes)):
        c = ch[i][ch[i] > 0]
        depth[c] = depth[i] + 1
    leaf = (ch == 0).all(axis=1)
    leaf[0] = False
    assert (depth[leaf] == sc.levels).all()


def test_voxelise_properties():
    """qmc output is x-major sorted, one point per grid cell, all inside the cube (qmc.c L66-88, L259, L291)."""
    rng = np.random.default_rng(9)
    pos = rng.uniform(-100, 1900, size=(40000, 3)).astype(np.float32)
    col = rng.integers(0, 256, size=(len(pos), 3)).astype(np.uint8)
    nrm = rng.normal(size=(len(pos), 3)).astype(np.float32)
    p, c, n = S.voxelise(pos, col, nrm)
    prec = np.float32(1800.0) / np.float32(8192.0)
    g = np.floor(p / prec).astype(np.int64)
    assert (g >= 0).all() and (g < 8192).all()
    key = (g[:, 0] * 8192 + g[:, 1]) * 8192 + g[:, 2]
    assert (np.diff(key) > 0).all()
    assert ((c >= 0) & (c <= 1)).all()
    # empty and single inputs
    e = np.zeros((0, 3), np.float32)
    p0, c0, n0 = S.voxelise(e, np.zeros((0, 3), np.uint8), e)
    assert len(p0) == 0
