"""The oracle restatement against the reference's own compiled CPU twin and the committed golden frames."""
import math
import os
import struct

import numpy as np
import pytest

from oracle import qb_oracle as O
from qubatron_b200 import scene as S

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference)")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@needs_ref
@pytest.mark.parametrize("pos,ang", [(S.CAMERA_C1[0], S.CAMERA_C1[1]), ((760.0, 160.0, 330.0), (-0.7, -0.35, 0.0)),
                                     ((1200.0, 300.0, 900.0), (-0.9, -0.2, 0.0))])
def test_trace_equals_reference_cpu_twin(scene_c1, pos, ang):
    """cube_trace_line restatement (IEEE-division build, like the compiled C of the reference) vs
    octree_trace_line (octree.c L341-537), static tree, every pixel's primary ray: same hit mask, same hit model
    index."""
    ref = O.RefOctree()
    ref.insert_points(scene_c1.pnt_s)
    u = O.uniforms(320, 180, pos, ang)
    rays = O.pixel_rays(u).reshape(-1, 3)
    org = np.tile(np.asarray(pos, np.float32), (len(rays), 1))
    idx, tlf = ref.trace(org, rays)
    res, nodes, models, isp = O.trace_batch(O.OracleScene(scene_c1), u, org, rays, div=O.DIV_IEEE)
    leaf = res == 1
    assert leaf.sum() > 500
    assert np.array_equal(idx[leaf], models[leaf, 0])
    assert (idx[~leaf] == 0).all()


@needs_ref
def test_shadow_rays_equal_reference_cpu_twin(scene_c1):
    """Same for the rays the shadow pass shoots: from the light to every primary hit point."""
    ref = O.RefOctree()
    ref.insert_points(scene_c1.pnt_s)
    osc = O.OracleScene(scene_c1)
    u = O.uniforms(320, 180, *S.CAMERA_C1)
    rays = O.pixel_rays(u).reshape(-1, 3)
    org = np.tile(np.asarray(S.CAMERA_C1[0], np.float32), (len(rays), 1))
    res, _, _, isp = O.trace_batch(osc, u, org, rays, div=O.DIV_IEEE)
    hit = (res == 1) & (isp[:, 3] > 0)
    light = np.array([u.light[0], u.light[1], u.light[2]], np.float32)
    sdir = (isp[hit, :3] - light[None, :]).astype(np.float32)
    sorg = np.tile(light, (len(sdir), 1))
    idx, _ = ref.trace(sorg, sdir)
    res2, _, models2, _ = O.trace_batch(osc, u, sorg, sdir, div=O.DIV_IEEE)
    leaf = res2 == 1
    assert np.array_equal(idx[leaf], models2[leaf, 0]) and (idx[~leaf] == 0).all()


def test_octtest_fixture():
    """The reference's only in-tree fixture: 5 points (modelutil.c L89-110) seen from (900,900,3000)
    (qubatron.c L136).  56 nodes (SURVEY 8c probe); leaf model indices are the five inserted points."""
    sc = S.make_test5()
    assert len(sc.oct_s) == 56
    u = O.uniforms(320, 200, (900.0, 900.0, 3000.0), (0.0, 0.0, 0.0))
    r = O.render(O.OracleScene(sc), u)
    leaf = (r["flags"] & O.FLAG_LEAF) > 0
    assert set(np.unique(r["aux"][leaf][:, 0]).tolist()) <= {0, 1, 2, 3, 4}
    # camera is outside the cube: pixels whose ray misses the cube are discarded (octree_fsh.c L195)
    assert ((r["flags"] & O.FLAG_DISCARD) > 0).any()
    assert (r["rgba"][(r["flags"] & O.FLAG_DISCARD) > 0] == 0).all()


def test_uniform_setup():
    """octree_glc.c L263-284: light from lightc and lighta (double arithmetic), render size width/(6-quality/2)."""
    u = O.uniforms(1200, 800, (1, 2, 3), (0.5, -0.25, 9.0), lighta=0.7, quality=10, maxlevel=11, basesize=1800.0)
    s = np.float32(math.sin(np.float32(0.7)))
    assert u.light[0] == 420.0
    assert u.light[1] == np.float32(200.0 - float(np.sin(np.float32(0.7), dtype=np.float32)) * 20.0)
    assert u.light[2] == np.float32(680.0 - float(np.sin(np.float32(0.7), dtype=np.float32)) * 200.0)
    assert (u.vp_w, u.vp_h) == (1200, 800) and u.angle_in[2] == 0.0 and u.maxlevel == 11
    assert tuple(u.basecube) == (0.0, 1800.0, 1800.0, 1800.0)
    u4 = O.uniforms(1200, 800, (1, 2, 3), (0, 0, 0), quality=4)
    assert (u4.vp_w, u4.vp_h) == (300, 200)
    del s


def test_disc_threshold_is_a_dot_threshold():
    """The CUDA path evaluates `acos(d) < 0.02` (octree_fsh.c L418, L452) as d >= d_min with d_min found by
    bisection on the host's acosf.  That is exact iff acosf is monotone non-increasing around the threshold:
    scan every float in [0.9995, 1]."""
    lo = struct.unpack("<I", struct.pack("<f", 0.9995))[0]
    hi = struct.unpack("<I", struct.pack("<f", 1.0))[0]
    bits = np.arange(lo, hi + 1, dtype=np.uint32)
    d = bits.view(np.float32)
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.acosf.restype = ctypes.c_float
    libm.acosf.argtypes = [ctypes.c_float]
    a = np.array([libm.acosf(float(x)) for x in d], dtype=np.float32)
    assert (np.diff(a) <= 0).all()
    below = a < np.float32(0.02)
    first = int(np.argmax(below))
    assert below[first:].all() and not below[:first].any()


def test_counters_and_algorithmic_bytes(scene_random):
    u = O.uniforms(128, 64, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))
    r = O.render(O.OracleScene(scene_random), u)
    c = r["counters"]
    f = r["flags"]
    assert c["rays_primary"] == 128 * 64
    assert c["rays_shadow"] == c["hits"] == int(((f & O.FLAG_SHADED) > 0).sum())
    assert c["rays_disc"] == int(((f & O.FLAG_DISC_TEST) > 0).sum()) + 0
    assert c["expand_d"] > 0 and c["leaf_d"] > 0
    b = O.algorithmic_bytes(c, 128 * 64)
    assert b == 32 * (c["expand_s"] + c["expand_d"]) + 4 * (c["leaf_s"] + c["leaf_d"]) + 24 * c["hits"] + 4 * 128 * 64


def test_rows_and_threads_are_deterministic(scene_random):
    osc = O.OracleScene(scene_random)
    u = O.uniforms(96, 48, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))
    a = O.render(osc, u, threads=1)
    b = O.render(osc, u, threads=4)
    assert np.array_equal(a["rgba"], b["rgba"]) and np.array_equal(a["aux"], b["aux"])
    assert a["counters"] == b["counters"]
    top = O.render(osc, u, rows=(24, 48), threads=2)
    assert np.array_equal(top["rgba"][24:], a["rgba"][24:]) and (top["rgba"][:24] == 0).all()


def test_coord_override_hook_is_transparent(scene_random):
    """render_with_coords (the hook the llvmpipe sweep uses to hand the oracle a rasteriser's interpolated `coord`)
    with the exact pixel centres is the plain render; a shifted coord plane moves the frame."""
    W, H = 96, 64
    u = O.uniforms(W, H, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))
    osc = O.OracleScene(scene_random)
    plain = O.render(osc, u)
    cx = (np.arange(W, dtype=np.float32) + np.float32(0.5))[None, :].repeat(H, 0)
    cy = (np.arange(H, dtype=np.float32) + np.float32(0.5))[:, None].repeat(W, 1)
    same = O.render_with_coords(osc, u, cx, cy)
    assert np.array_equal(same["rgba"], plain["rgba"]) and np.array_equal(same["aux"], plain["aux"])
    moved = O.render_with_coords(osc, u, cx + np.float32(3.0), cy)
    assert np.array_equal(moved["rgba"][:, :-3], plain["rgba"][:, 3:])
    again = O.render(osc, u)          # the hook is cleared afterwards
    assert np.array_equal(again["rgba"], plain["rgba"])
