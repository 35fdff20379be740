"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bar: flags (discard / leaf / shaded / shadow-visibility / disc) and all six hit-index planes bit-exact;
RGBA within +-1/255 (parity.RGB_TOL).  Every test runs BOTH kernels (generic stack, fast register/shared
stack) unless the base cube is not exactly representable, where only the generic kernel is legal."""
import numpy as np
import pytest

import parity
from oracle import qb_oracle as O
from qubatron_b200 import connector as K
from qubatron_b200 import scene as S

pytestmark = pytest.mark.gpu

KERNELS = [(K.KERNEL_GENERIC, "generic"), (K.KERNEL_FAST, "fast")]
# both reproducible executions of the reference's `/` (include/octree_cuc.h, octree_cuc_set_division)
DIVS = [(K.DIV_GLSL, O.DIV_GLSL, "glsl"), (K.DIV_IEEE, O.DIV_IEEE, "ieee")]


def _render_and_compare(sc, W, H, pos, ang, kernels=KERNELS, rc=None, divs=DIVS, **kw):
    own = rc is None
    if own:
        rc = K.OctreeGlc(b"", device=0)
        rc.upload_scene(sc)
    rc.enable_aux(True)
    rc.enable_counters(True)
    osc = O.OracleScene(sc)
    outs = {}
    ref = None
    for kdiv, odiv, dname in divs:
        r = O.render(osc, O.uniforms(W, H, pos, ang, **kw), div=odiv)
        ref = ref or r
        rc.set_division(kdiv)
        for kern, name in kernels:
            rc.set_kernel(kern)
            rc.update(W, H, pos, ang, **kw)
            rgba = rc.read_frame()
            flags, aux = rc.read_aux()
            outs[name + "/" + dname] = parity.compare(rgba, flags, aux, r, what=name + "/" + dname)
            assert rc.read_counters() == r["counters"], name + "/" + dname
            assert rc.last_kernel() == kern
    rc.set_division(K.DIV_GLSL)
    if own:
        rc.destroy()
    return ref, outs


def test_c1_reference_start_pose(scene_c1):
    """BASELINE config 1: 1 M-point cloud, 640x360, camera (700,150,350) angle (0.4636,0), light angle 0."""
    ref, _ = _render_and_compare(scene_c1, 640, 360, *S.CAMERA_C1)
    f = ref["flags"]
    assert ((f & O.FLAG_SHADED) > 0).sum() > 100000 and ((f & O.FLAG_LIT) == 0).sum() > 1000


@pytest.mark.parametrize("pos,ang,kw", [
    ((700.0, 150.0, 350.0), (0.4636, 0.0, 0.0), dict(shoot=1, lighta=1.0)),
    ((760.0, 125.0, 225.0), (2.0, 0.3, 0.0), {}),                 # camera inside the sphere
    ((1200.0, 300.0, 900.0), (-0.9, -0.2, 0.0), {}),              # light disc visible
    ((2600.0, 2300.0, 900.0), (-1.3, -0.5, 0.0), {}),             # outside the cube: discards
    ((700.0, 100.3, 350.0), (0.4636, 0.02, 0.0), dict(lighta=4.0)),  # grazing the floor
    ((760.0, 160.0, 330.0), (-0.7, -0.35, 0.0), dict(quality=8)),  # render size width/2
])
def test_c1_cameras(scene_c1, pos, ang, kw):
    _render_and_compare(scene_c1, 640, 360, pos, ang, **kw)


def test_axis_parallel_rays(scene_c1):
    """angle (0,0): the centre columns have direction components that are exactly 0 or tiny
    (octree_fsh.c L65/L78/L91 parallel-plane sentinel)."""
    _render_and_compare(scene_c1, 320, 180, (750.0, 140.0, 330.0), (0.0, 0.0, 0.0))
    _render_and_compare(scene_c1, 321, 181, (750.5, 140.25, 330.0), (np.pi / 2, 0.0, 0.0))


def test_dynamic_tree_and_sparse_cloud(scene_random):
    """Both trees populated and overlapping: dual-tree candidate merge (L313-317), dynamic override (L233-241),
    heavy backtracking (80 expansions per ray)."""
    ref, _ = _render_and_compare(scene_random, 256, 128, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))
    assert (ref["aux"][..., K.AUX_MODEL_D] > 0).sum() > 100


def test_octtest_fixture():
    """The reference's 5-point scene (modelutil.c L89-110) from (900,900,3000) (qubatron.c L136)."""
    _render_and_compare(S.make_test5(), 320, 200, (900.0, 900.0, 3000.0), (0.0, 0.0, 0.0))


def test_empty_scene():
    """Nothing uploaded at all: every trace misses, frame is clear colour, no crash."""
    rc = K.OctreeGlc(b"", device=0)
    for kern, _ in KERNELS:
        rc.set_kernel(kern)
        rc.update(160, 90, (700.0, 150.0, 350.0), (0.3, 0.0, 0.0))
        assert (rc.read_frame() == 0).all()
    rc.destroy()


def test_ragged_frame_sizes(scene_random):
    """Frame sizes that are not multiples of the 16x8 CTA block or the 64x64 shard tile."""
    for W, H in ((1, 1), (17, 9), (63, 65), (130, 70)):
        _render_and_compare(scene_random, W, H, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))


@pytest.mark.parametrize("levels", [4, 9, 12])
def test_other_depths(levels):
    """maxlevel / -l option (qubatron.c L598-614): shallower trees on the same cube."""
    sc = S.make_random(20000, 3000, seed=5, levels=levels)
    _render_and_compare(sc, 200, 120, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0), maxlevel=levels)


def test_inexact_base_cube_uses_generic_kernel():
    """A base size whose grid is not exactly representable in fp32 must take the generic kernel and still match."""
    sc = S.make_random(20000, 2000, seed=6, basesize=1000.1)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_aux(True)
    pos, ang = (422.0, 111.0, 233.0), (-0.05, -0.12, 0.0)
    ref = O.render(O.OracleScene(sc), O.uniforms(200, 120, pos, ang, basesize=1000.1))
    rc.set_kernel(K.KERNEL_AUTO)
    rc.update(200, 120, pos, ang, basesize=1000.1)
    assert rc.last_kernel() == K.KERNEL_GENERIC
    flags, aux = rc.read_aux()
    parity.compare(rc.read_frame(), flags, aux, ref, what="inexact")
    assert ((ref["flags"] & O.FLAG_LEAF) > 0).sum() > 100
    # and the auto choice for the reference cube is the fast kernel
    rc.update(200, 120, pos, ang, basesize=1800.0)
    assert rc.last_kernel() == K.KERNEL_FAST
    rc.destroy()


def test_range_updates_zero_and_append(scene_c1):
    """modelutil_punch_hole's upload pattern (modelutil.c L429-437, L486-501, L528-546): zero child slots,
    append paths, upload each touched 48-byte node, then a colour/normal sub-range.  The device copy must equal
    a fresh full upload of the final host arrays, and both must match the oracle."""
    tree = S.HostOctree()
    tree.insert_points(scene_c1.pnt_s)
    col = scene_c1.col_s.copy()
    nrm = scene_c1.nrm_s.copy()
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_points(col, K.STATIC_COLOR)
    rc.upload_points(nrm, K.STATIC_NORMAL)
    rc.upload_octree(tree.nodes(copy=False))
    rc.upload_octree(np.zeros((1, 12), np.int32), dynamic=True)

    rng = np.random.default_rng(3)
    centre = np.array([760.0, 125.0, 245.0], np.float32)  # front of the sphere
    cand = np.nonzero(np.linalg.norm(scene_c1.pnt_s - centre[None, :], axis=1) < 9.0)[0]
    victims = rng.choice(cand, size=min(400, len(cand)), replace=False)
    touched_models = []
    for v in victims:
        m, o = tree.remove_point(scene_c1.pnt_s[v])
        if o >= 0:
            nodes = tree.nodes(copy=False)
            rc.upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, o * 48, (o + 1) * 48, K.STATIC_OCTREE)
            touched_models.append(m)
    for m in touched_models:
        newp = scene_c1.pnt_s[m] + rng.normal(0, 3.0, size=3).astype(np.float32)
        t = tree.insert_point(newp, m)
        nodes = tree.nodes(copy=False)
        for j in t:
            if j > 0:
                rc.upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, int(j) * 48, (int(j) + 1) * 48,
                                         K.STATIC_OCTREE)
        col[m] += 0.2
        nrm[m] = centre - newp
    lo, hi = min(touched_models), max(touched_models) + 1
    rc.upload_points(col, K.STATIC_COLOR, lo, hi)
    rc.upload_points(nrm, K.STATIC_NORMAL, lo, hi)

    final = S.Scene("punched", scene_c1.pnt_s, col, nrm, tree.nodes(), scene_c1.pnt_d, scene_c1.col_d,
                    scene_c1.nrm_d, scene_c1.oct_d)
    pos, ang = (745.0, 135.0, 300.0), (0.15, -0.1, 0.0)
    ref, _ = _render_and_compare(final, 320, 180, pos, ang, rc=rc, shoot=1)
    rc.destroy()
    # the hole changed the picture
    before = O.render(O.OracleScene(scene_c1), O.uniforms(320, 180, pos, ang, shoot=1))
    assert (before["aux"] != ref["aux"]).any()


def test_texel_granularity_of_uploads(scene_random):
    """start/end are rounded DOWN to itemsize (octree_glc.c L432-433): a range that covers part of an item
    must not upload that item."""
    sc = scene_random
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    wrong = sc.oct_s.copy()
    wrong[100:200, :8] = 0
    # bytes [100*48 + 5, 200*48 - 3) -> items [300, 599): node 199's last texel (model index + pad) is skipped,
    # node 100's first texel IS included because start rounds down
    rc.upload_texbuffer_data(wrong, K.GL_INT, wrong.size * 4, 16, 100 * 48 + 5, 200 * 48 - 3, K.STATIC_OCTREE)
    expect = sc.oct_s.copy()
    expect[100:200, :8] = 0
    final = S.Scene("granular", sc.pnt_s, sc.col_s, sc.nrm_s, expect, sc.pnt_d, sc.col_d, sc.nrm_d, sc.oct_d)
    _render_and_compare(final, 160, 90, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0), rc=rc)
    rc.destroy()


def test_growth_keeps_contents(scene_random):
    """Appending beyond the device capacity (size grows) re-uploads the whole array like the reference
    (octree_glc.c L412-428) and keeps rendering correct."""
    sc = scene_random
    rc = K.OctreeGlc(b"", device=0)
    half = len(sc.oct_s) // 2
    # first a truncated tree (children pointing past the uploaded range read as absent)
    rc.upload_points(sc.col_s, K.STATIC_COLOR)
    rc.upload_points(sc.nrm_s, K.STATIC_NORMAL)
    rc.upload_texbuffer_data(sc.oct_s[:half].copy(), K.GL_INT, half * 48, 16, 0, half * 48, K.STATIC_OCTREE)
    rc.update(64, 32, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))
    rc.sync()
    # then the whole array with a range that only names the tail
    rc.upload_texbuffer_data(sc.oct_s, K.GL_INT, len(sc.oct_s) * 48, 16, half * 48, len(sc.oct_s) * 48,
                             K.STATIC_OCTREE)
    rc.upload_points(sc.col_d, K.DYNAMIC_COLOR)
    rc.upload_points(sc.nrm_d, K.DYNAMIC_NORMAL)
    rc.upload_octree(sc.oct_d, dynamic=True)
    _render_and_compare(sc, 160, 90, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0), rc=rc)
    assert rc.memsize > 0
    rc.destroy()


def test_multi_view_batch(scene_random):
    """octree_cuc_update_views: n views in one launch equal n single frames."""
    rng = np.random.default_rng(777)
    n = 5
    pos = np.stack([rng.uniform(700, 900, n), rng.uniform(150, 260, n), rng.uniform(250, 500, n)], axis=1)
    ang = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.4, 0.2, n), np.zeros(n)], axis=1)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    rc.enable_aux(True)
    osc = O.OracleScene(scene_random)
    for kern, name in KERNELS:
        rc.set_kernel(kern)
        rc.update_views(128, 64, pos, ang)
        rgba = rc.read_frame(views=n).reshape(n, 64, 128, 4)
        flags, aux = rc.read_aux(views=n)
        flags, aux = flags.reshape(n, 64, 128), aux.reshape(n, 64, 128, 6)
        for v in range(n):
            ref = O.render(osc, O.uniforms(128, 64, pos[v], ang[v]))
            parity.compare(rgba[v], flags[v], aux[v], ref, what="%s view %d" % (name, v))
    rc.destroy()


def test_tile_sharding_on_one_gpu(scene_random):
    """Image-tile sharding: world ranks render disjoint interleaved tiles; their union is the full frame."""
    W, H, world = 200, 150, 3
    pos, ang = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)
    ref = O.render(O.OracleScene(scene_random), O.uniforms(W, H, pos, ang))
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    for kern, name in KERNELS:
        rc.set_kernel(kern)
        acc = np.zeros((H, W, 4), np.uint8)
        owner = np.full((H, W), -1)
        for r in range(world):
            rc.set_shard(r, world, 32, 16)
            # poison, then render: only this rank's tiles may change
            rc.set_shard(0, 1, 32, 16)
            rc.update(W, H, (5000.0, 5000.0, 5000.0), (0.0, 0.0, 0.0))  # all-discard frame = zeros
            rc.set_shard(r, world, 32, 16)
            rc.update(W, H, pos, ang)
            f = rc.read_frame()
            ty, tx = np.arange(H)[:, None] // 16, np.arange(W)[None, :] // 32
            mine = ((ty * ((W + 31) // 32) + tx) % world) == r
            assert (f[~mine] == 0).all(), name
            acc[mine] = f[mine]
            owner[mine] = r
        assert (owner >= 0).all()
        assert np.abs(acc.astype(np.int16) - ref["rgba"].astype(np.int16)).max() <= parity.RGB_TOL, name
    rc.destroy()


def test_external_frame_target(scene_random):
    """Rendering straight into caller-owned device memory (a torch tensor) with a row pitch."""
    import torch
    W, H, pitch = 120, 50, 128
    pos, ang = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)
    ref = O.render(O.OracleScene(scene_random), O.uniforms(W, H, pos, ang))
    buf = torch.zeros((H, pitch, 4), dtype=torch.uint8, device="cuda:0")
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    rc.set_frame_target(buf.data_ptr(), pitch, keepalive=buf)
    rc.update(W, H, pos, ang)
    rc.sync()
    got = buf.cpu().numpy()
    assert np.abs(got[:, :W].astype(np.int16) - ref["rgba"].astype(np.int16)).max() <= parity.RGB_TOL
    assert (got[:, W:] == 0).all()
    rc.destroy()


def test_full_size_properties(scene_c1):
    """1920x1080 (BASELINE full frame size) through size-independent properties: both kernels agree bit for
    bit with each other, counters agree, alpha is 255 exactly on leaf pixels, shadowed pixels are darker than
    their lit version would be (LIT bit off => no 0.7 light term)."""
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_c1)
    rc.enable_aux(True)
    rc.enable_counters(True)
    out = {}
    for kern, name in KERNELS:
        rc.set_kernel(kern)
        rc.update(1920, 1080, *S.CAMERA_C1)
        out[name] = (rc.read_frame(), rc.read_aux(), rc.read_counters())
    g, f = out["generic"], out["fast"]
    assert np.array_equal(g[0], f[0]) and np.array_equal(g[1][0], f[1][0]) and np.array_equal(g[1][1], f[1][1])
    assert g[2] == f[2]
    flags = f[1][0]
    leaf = (flags & K.FLAG_LEAF) > 0
    assert ((f[0][..., 3] == 255) == leaf).all()
    assert f[2]["rays_primary"] == 1920 * 1080 and f[2]["rays_shadow"] == int(((flags & K.FLAG_SHADED) > 0).sum())
    # a sub-window rendered by the oracle pins the full-size frame to the reference semantics
    u = O.uniforms(1920, 1080, *S.CAMERA_C1)
    ref = O.render(O.OracleScene(scene_c1), u, rows=(500, 560))
    assert np.array_equal(ref["flags"][500:560], flags[500:560])
    assert np.array_equal(ref["aux"][500:560], f[1][1][500:560])
    assert np.abs(ref["rgba"][500:560].astype(np.int16) - f[0][500:560].astype(np.int16)).max() <= parity.RGB_TOL
    rc.destroy()


def test_hoisted_division_equals_ieee_division():
    """The fast kernel replaces `(c - o) / d` by nvcc's own fast-path FFMA sequence with the reciprocal
    hoisted out of the loop; on 2 x 200 M operand pairs drawn like the traversal's it must give the IEEE
    quotient every time (+-0 aside, which no comparison in the traversal can tell apart)."""
    rc = K.OctreeGlc(b"", device=0)
    for seed in (1, 2):
        assert rc.selftest_div(seed, 100_000_000) == 0
    rc.destroy()


@pytest.mark.parametrize("name", __import__("golden_util").CASES)
def test_golden_frames_of_the_reference_shader(name):
    """The CUDA path against what the reference's own shader produced on llvmpipe (tests/golden/): in the default
    (GLSL) division mode both kernels must give the shader's hit voxel indices and shadow bits exactly and its RGBA
    within +-1/255 -- they give it exactly."""
    import golden_util
    sc, args, g = golden_util.load(name)
    if sc is None:
        pytest.skip("scene generator gives different bits on this CPU (hash mismatch); embedded cases still run")
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_aux(True)
    for kern, kname in KERNELS:
        rc.set_kernel(kern)
        rc.update(args["width"], args["height"], args["position"], args["angle"], args.get("lighta", 0.0),
                  args.get("quality", 10), args.get("maxlevel", 12), args.get("basesize", 1800.0), args.get("shoot", 0))
        rgba = rc.read_frame()
        flags, aux = rc.read_aux()
        assert np.abs(rgba.astype(np.int16) - g["rgba"].astype(np.int16)).max() <= parity.RGB_TOL, kname
        assert np.array_equal(rgba, g["rgba"]), kname
        leaf = (flags & K.FLAG_LEAF) > 0
        assert np.array_equal(g["model_s"][leaf], aux[leaf][:, K.AUX_MODEL_S]), kname
        assert np.array_equal(g["model_d"][leaf], aux[leaf][:, K.AUX_MODEL_D]), kname
        shaded = (flags & K.FLAG_SHADED) > 0
        assert np.array_equal(g["shadow"][shaded] != 0, (flags[shaded] & K.FLAG_LIT) > 0), kname
    rc.destroy()


def test_c_host_demo():
    """examples/host_demo.c: a plain-C host making the reference engine's own call sequence against the library."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ex = os.path.join(root, "examples")
    r = subprocess.run(["make", "-C", ex], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([os.path.join(ex, "host_demo")], cwd=ex, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert "leaf pixels" in r.stdout and "dynamic pipeline" in r.stdout


def test_dynamic_tree_rebuilt_every_frame(scene_c1):
    """BASELINE config 4 at test size: per frame the dynamic tree is reset and rebuilt from re-pathed points
    (octree_reset + octree_insert_path, qubatron.c L439-452), uploaded whole (L508-516) together with the normals
    [0, n) from a DIFFERENT host buffer (L521-529), then the frame is rendered."""
    rng = np.random.default_rng(12)
    fig_p, fig_c, fig_n = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.6, shells=2)
    fig_p, fig_cf, fig_n = S.voxelise(fig_p, fig_c, fig_n)
    n = len(fig_p)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_points(scene_c1.col_s, K.STATIC_COLOR)
    rc.upload_points(scene_c1.nrm_s, K.STATIC_NORMAL)
    rc.upload_octree(scene_c1.oct_s)
    rc.upload_points(fig_cf, K.DYNAMIC_COLOR)
    rc.enable_aux(True)
    tree = S.HostOctree()
    W, H = 320, 180
    seen_dynamic = 0
    for frame in range(3):
        moved = (fig_p + np.array([3.0 * frame, 0.0, -2.0 * frame], np.float32)).astype(np.float32)
        nrm_out = (fig_n + rng.normal(0, 0.05, size=fig_n.shape)).astype(np.float32)  # skelglc.nrm_out analogue
        tree.reset()
        tree.insert_paths(S.octant_paths(moved))
        nodes = tree.nodes()
        rc.upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, 0, len(nodes) * 48, K.DYNAMIC_OCTREE)
        rc.upload_texbuffer_data(nrm_out, K.GL_FLOAT, n * 12, 12, 0, n * 12, K.DYNAMIC_NORMAL)
        sc = S.Scene("dyn", scene_c1.pnt_s, scene_c1.col_s, scene_c1.nrm_s, scene_c1.oct_s, moved, fig_cf, nrm_out,
                     nodes)
        ref, _ = _render_and_compare(sc, W, H, *S.CAMERA_C1, rc=rc, lighta=0.3 * frame)
        seen_dynamic += int((ref["aux"][..., K.AUX_MODEL_D] > 0).sum())
    assert seen_dynamic > 3000
    rc.destroy()


@pytest.fixture(scope="module")
def c2_level():
    """BASELINE configs[1..4] share one level: the 95 M-point / 62 M-node static tree + the 10 M-point figure, at
    FULL size (bench.py's cached scene; QB_TEST_SCALE shrinks it with the same generator), uploaded once."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    scale = float(os.environ.get("QB_TEST_SCALE", "1.0"))
    sc, meta = bench.get_scene(scale, 0, lambda: None)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_aux(True)
    rc.enable_counters(True)
    yield sc, rc, O.OracleScene(sc)
    rc.destroy()


def test_c2_full_scale_frames_equal_the_oracle(c2_level):
    """BASELINE configs[1] at FULL size: 1920x1080, all four bench poses -- every pixel's flags (hit / shadow
    visibility / disc), hit indices and RGBA against the oracle (~1 s per frame on the box's host cores)."""
    sc, rc, osc = c2_level
    for pose, (pos, ang) in enumerate(sc.cameras):
        ref = O.render(osc, O.uniforms(1920, 1080, pos, ang))
        rc.set_kernel(K.KERNEL_AUTO)
        rc.update(1920, 1080, pos, ang)
        assert rc.last_kernel() == K.KERNEL_FAST
        flags, aux = rc.read_aux()
        out = parity.compare(rc.read_frame(), flags, aux, ref, what="C2 pose %d" % pose)
        assert out["rgba_maxdiff"] == 0
        assert rc.read_counters() == ref["counters"]


def _compare_bands(rgba, flags, aux, osc, u, bands, div, what):
    """the oracle renders only `bands` (row ranges) of the frame: flags and indices exact, RGBA identical there"""
    checked = 0
    for r0, r1 in bands:
        ref = O.render(osc, u, rows=(r0, r1), div=div)
        assert np.array_equal(ref["flags"][r0:r1], flags[r0:r1]), "%s rows %d-%d: flags" % (what, r0, r1)
        assert np.array_equal(ref["aux"][r0:r1], aux[r0:r1]), "%s rows %d-%d: hit indices" % (what, r0, r1)
        assert np.array_equal(ref["rgba"][r0:r1], rgba[r0:r1]), "%s rows %d-%d: rgba" % (what, r0, r1)
        checked += int(((flags[r0:r1] & K.FLAG_LEAF) > 0).sum())
    return checked


def test_c3_2160p_frames_equal_the_oracle(c2_level):
    """BASELINE configs[2]: the same level at 3840x2160 (a frame the reference cannot render: its FBO is 2048^2,
    octree_glc.c L237).  Four poses, both division modes; the oracle renders three bands of 48 rows per frame (bottom
    edge, middle, top edge), and the two kernels must agree on EVERY pixel of the frame."""
    sc, rc, osc = c2_level
    W, H = 3840, 2160
    bands = [(0, 48), (1056, 1104), (H - 48, H)]
    leafs = 0
    for kdiv, odiv, dname in DIVS:
        rc.set_division(kdiv)
        for pose, (pos, ang) in enumerate(sc.cameras):
            u = O.uniforms(W, H, pos, ang)
            got = {}
            for kern, name in KERNELS:
                rc.set_kernel(kern)
                rc.update(W, H, pos, ang)
                assert rc.last_kernel() == kern
                flags, aux = rc.read_aux()
                got[name] = (rc.read_frame().copy(), flags, aux, rc.read_counters())
            f, g = got["fast"], got["generic"]
            for a, b, nm in zip(f[:3], g[:3], ("rgba", "flags", "aux")):
                assert np.array_equal(a, b), "4K pose %d %s: kernels differ on %s" % (pose, dname, nm)
            assert f[3] == g[3] and f[3]["rays_primary"] == W * H
            leafs += _compare_bands(f[0], f[1], f[2], osc, u, bands, odiv, "4K pose %d %s" % (pose, dname))
    rc.set_division(K.DIV_GLSL)
    assert leafs > 200000


def test_c5_random_views_batch_equals_the_oracle(c2_level):
    """BASELINE configs[4]: bench.py's 64 incoherent cameras (default_rng(777)) on the full level; 8 of them rendered
    as ONE octree_cuc_update_views batch at 1080p.  GLSL division: every pixel of every view against the oracle;
    IEEE division: three bands per view."""
    sc, rc, osc = c2_level
    n = 64
    rng = np.random.default_rng(777)
    pos = np.stack([rng.uniform(150, 1650, n), rng.uniform(90, 330, n), rng.uniform(150, 1650, n)], axis=1)
    ang = np.stack([rng.uniform(0, 2 * np.pi, n), rng.uniform(-0.6, 0.6, n), np.zeros(n)], axis=1)
    pick = list(range(0, n, 8))
    W, H = 1920, 1080
    rc.set_kernel(K.KERNEL_AUTO)
    leafs = 0
    for kdiv, odiv, dname in DIVS:
        rc.set_division(kdiv)
        rc.update_views(W, H, pos[pick], ang[pick])
        assert rc.last_kernel() == K.KERNEL_FAST
        rgba = rc.read_frame(views=len(pick))
        flags, aux = rc.read_aux(views=len(pick))
        total = dict.fromkeys(rc.read_counters(), 0)
        for k, v in enumerate(pick):
            u = O.uniforms(W, H, tuple(pos[v].astype(np.float32)), tuple(ang[v].astype(np.float32)))
            sl = slice(k * H, (k + 1) * H)
            if odiv == O.DIV_GLSL:
                ref = O.render(osc, u, div=odiv)
                out = parity.compare(rgba[sl], flags[sl], aux[sl], ref, what="C5 view %d" % v)
                assert out["rgba_maxdiff"] == 0
                for key in total:
                    total[key] += ref["counters"][key]
                leafs += int(((ref["flags"] & O.FLAG_LEAF) > 0).sum())
            else:
                _compare_bands(rgba[sl], flags[sl], aux[sl], osc, u, [(0, 40), (520, 560), (H - 40, H)], odiv,
                               "C5 view %d ieee" % v)
        if odiv == O.DIV_GLSL:
            assert rc.read_counters() == total
    rc.set_division(K.DIV_GLSL)
    assert leafs > 1000000


def test_pipelined_readback_returns_the_right_frames(scene_random):
    """octree_cuc_read_frame_async: frame i is copied to the host while frame i+1 renders into the second
    framebuffer; every host buffer must hold exactly its own frame."""
    import torch
    cams = [((760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)), ((800.0, 230.0, 380.0), (-0.6, -0.3, 0.0)),
            ((700.0, 260.0, 500.0), (0.3, -0.4, 0.0)), ((900.0, 150.0, 300.0), (-1.2, 0.1, 0.0)),
            ((760.0, 200.0, 420.0), (-0.05, -0.12, 0.0))]
    W, H = 200, 120
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    sync = []
    for pos, ang in cams:
        rc.update(W, H, pos, ang)
        sync.append(rc.read_frame().copy())
    bufs = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in cams]
    for (pos, ang), b in zip(cams, bufs):
        rc.update(W, H, pos, ang, shoot=0)
        rc.read_frame_async(b.numpy())
    rc.wait_reads()
    for a, b in zip(sync, bufs):
        assert np.array_equal(a, b.numpy())
    # the synchronous read still returns the latest frame once the ring is active
    rc.update(W, H, *cams[1])
    assert np.array_equal(rc.read_frame(), sync[1])
    rc.destroy()


def test_gpu_tree_build_from_paths_equals_sequential_insert():
    """octree_cuc_build_octree_from_paths ("next" row 8f #1): node-for-node the array that octree_reset +
    octree_insert_path (octree.c L149-180; qubatron.c L439-452) produce -- which tests/test_host_model.py pins
    against the reference's compiled code -- for random digits, for a surface-like figure, for duplicates and
    for the empty and single-point cases."""
    rng = np.random.default_rng(31)
    fig_p, fig_c, fig_n = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.5, shells=3)
    cases = {
        "random": rng.integers(0, 8, size=(60000, 12)).astype(np.int32),
        "figure": S.octant_paths(fig_p),
        "duplicates": np.repeat(rng.integers(0, 8, size=(500, 12)).astype(np.int32), 7, axis=0),
        "single": rng.integers(0, 8, size=(1, 12)).astype(np.int32),
        "empty": np.zeros((0, 12), np.int32),
    }
    rc = K.OctreeGlc(b"", device=0)
    for name, paths in cases.items():
        host = S.HostOctree()
        host.insert_paths(paths, first_modind=5)
        n = rc.build_octree_from_paths(paths, first_modind=5, dynamic=True)
        got = rc.download_octree(dynamic=True)
        want = host.nodes()
        assert n == len(want) == len(got), name
        assert np.array_equal(got, want), name
    rc.destroy()


def test_render_with_gpu_built_dynamic_tree(scene_c1):
    """The frame rendered from a dynamic tree built on the GPU equals the oracle's frame for the host-built tree."""
    fig_p, fig_c, fig_n = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.6, shells=2)
    fig_p, fig_cf, fig_n = S.voxelise(fig_p, fig_c, fig_n)
    paths = S.octant_paths(fig_p)
    host = S.HostOctree()
    host.insert_paths(paths)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_points(scene_c1.col_s, K.STATIC_COLOR)
    rc.upload_points(scene_c1.nrm_s, K.STATIC_NORMAL)
    rc.upload_octree(scene_c1.oct_s)
    rc.upload_points(fig_cf, K.DYNAMIC_COLOR)
    rc.upload_points(fig_n, K.DYNAMIC_NORMAL)
    rc.build_octree_from_paths(paths)
    sc = S.Scene("gpu-built", scene_c1.pnt_s, scene_c1.col_s, scene_c1.nrm_s, scene_c1.oct_s, fig_p, fig_cf, fig_n,
                 host.nodes())
    ref, _ = _render_and_compare(sc, 320, 180, *S.CAMERA_C1, rc=rc)
    assert (ref["aux"][..., K.AUX_MODEL_D] > 0).sum() > 1000
    rc.destroy()


@pytest.mark.parametrize("div", [K.DIV_GLSL, K.DIV_IEEE])
def test_skinning_pass_equals_the_oracle(div):
    """octree_cuc_skeleton_update ("next" row 8f #1, first half): the twelve octant digits, the blended normal and
    the skinned position of every point equal oracle/skeleton_vsh_oracle.c (skeleton_vsh.c L74-226) bit for bit,
    in both division modes, for the rest pose, two bent poses and a partial model_count."""
    pos, col, nrm = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.5, shells=3)
    rc = K.OctreeGlc(b"", device=0)
    rc.set_division(div)
    rc.skeleton_alloc_in(pos, nrm)
    for pose, count in ((0.0, None), (1.0, None), (2.5, None), (1.0, len(pos) // 3), (1.0, 0)):
        ob, nb = S.zombie_bones(pose=pose, shift=(3.0 * pose, 0.0, -2.0 * pose))
        n = len(pos) if count is None else count
        rc.skeleton_update(ob, nb, model_count=count, build_tree=False)
        digits, nrm_out, pnt_out = rc.skeleton_read_out(n)
        want_d, want_n, want_p = O.skin(ob, nb, pos[:n], nrm[:n], div=div)
        assert np.array_equal(digits, want_d), pose
        assert np.array_equal(nrm_out.view(np.uint32), want_n.view(np.uint32)), pose
        assert np.array_equal(pnt_out.view(np.uint32), want_p.view(np.uint32)), pose
        if pose > 0 and n:
            assert (np.abs(want_p - pos[:n]).max(axis=1) > 1.0).mean() > 0.2   # the pose really moves the figure
    rc.destroy()


@pytest.mark.parametrize("name", __import__("golden_util").SKIN_CASES)
def test_skinning_pass_equals_the_reference_vertex_program_on_llvmpipe(name):
    """The CUDA skinning pass against skeleton_vsh.c itself (tests/golden/skin_*.npz, transform feedback on
    llvmpipe): given the driver's ten bone rotations, every digit, normal and skinned position is bit-identical."""
    import golden_util
    pos, nrm, g = golden_util.load_skin(name)
    rc = K.OctreeGlc(b"", device=0)
    rc.skeleton_alloc_in(pos, nrm)
    rc.skeleton_set_rotations(g["rotations"])
    rc.skeleton_update(g["oldbones"], g["newbones"], build_tree=False)
    digits, nrm_out, pnt_out = rc.skeleton_read_out(len(pos))
    assert np.array_equal(digits, g["digits"])
    assert np.array_equal(nrm_out.view(np.uint32), g["normal_out"].view(np.uint32))
    assert np.array_equal(pnt_out.view(np.uint32), g["pnt"].view(np.uint32))
    rc.skeleton_set_rotations(None)   # back to libm: equals the oracle with libm rotations
    rc.skeleton_update(g["oldbones"], g["newbones"], build_tree=False)
    digits, nrm_out, pnt_out = rc.skeleton_read_out(len(pos))
    want_d, want_n, want_p = O.skin(g["oldbones"], g["newbones"], pos, nrm)
    assert np.array_equal(digits, want_d) and np.array_equal(pnt_out.view(np.uint32), want_p.view(np.uint32))
    rc.destroy()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_skinning_pass_random_poses_equal_the_oracle(seed):
    """Random skeleton poses (every joint jittered, random twists, shrunken radii so that some points are out of
    range of every bone pair and keep their position with a zero normal), both division modes: bit-exact."""
    rng = np.random.default_rng(900 + seed)
    fig_p, _, fig_n = S.zombie_raw(spacing=0.9, shells=3)
    sel = rng.choice(len(fig_p), 20000, replace=False)
    pos, nrm = fig_p[sel], fig_n[sel]
    ob, nb = S.zombie_bones(pose=float(rng.uniform(0, 3)), shift=tuple(rng.normal(0, 8, 3)))
    nb = nb.copy()
    nb[:, :3] += rng.normal(0, 2.0, (20, 3)).astype(np.float32)
    nb[::2, 3] = rng.normal(0, 0.4, 10).astype(np.float32)
    ob = ob.copy()
    ob[::2, 3] *= np.float32(0.6)
    rc = K.OctreeGlc(b"", device=0)
    rc.skeleton_alloc_in(pos, nrm)
    loose = 0
    for div in (K.DIV_GLSL, K.DIV_IEEE):
        rc.set_division(div)
        rc.skeleton_update(ob, nb, build_tree=False)
        digits, nrm_out, pnt_out = rc.skeleton_read_out(len(pos))
        want_d, want_n, want_p = O.skin(ob, nb, pos, nrm, div=div)
        assert np.array_equal(digits, want_d)
        assert np.array_equal(nrm_out.view(np.uint32), want_n.view(np.uint32))
        assert np.array_equal(pnt_out.view(np.uint32), want_p.view(np.uint32))
        loose = int(((want_n == 0).all(axis=1) & (want_p == pos).all(axis=1)).sum())
    assert loose > 0
    rc.destroy()


def test_skin_build_render_on_the_device_equals_the_host_pipeline(scene_c1):
    """The reference's per-frame dynamic-model pipeline (qubatron.c L425-452 + L508-548): skin -> read back ->
    octree_reset + octree_insert_path -> upload tree and normals -> render.  Here all three stages stay on the
    device; the frame must equal the oracle's frame of the host-side pipeline (oracle skinning, host octree)."""
    pos, col, nrm = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.6, shells=2)
    pos, colf, nrm = S.voxelise(pos, col, nrm)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_points(scene_c1.col_s, K.STATIC_COLOR)
    rc.upload_points(scene_c1.nrm_s, K.STATIC_NORMAL)
    rc.upload_octree(scene_c1.oct_s)
    rc.upload_points(colf, K.DYNAMIC_COLOR)
    rc.skeleton_alloc_in(pos, nrm)
    for pose in (0.5, 1.5):
        ob, nb = S.zombie_bones(pose=pose)
        nodes = rc.skeleton_update(ob, nb, build_tree=True)
        digits, nrm_out, pnt_out = O.skin(ob, nb, pos, nrm)
        host = S.HostOctree()
        host.insert_paths(digits)
        want = host.nodes()
        assert nodes == len(want)
        assert np.array_equal(rc.download_octree(dynamic=True), want)
        sc = S.Scene("skinned", scene_c1.pnt_s, scene_c1.col_s, scene_c1.nrm_s, scene_c1.oct_s, pnt_out, colf, nrm_out,
                     want)
        ref, _ = _render_and_compare(sc, 320, 180, *S.CAMERA_C1, rc=rc)
        assert (ref["aux"][..., K.AUX_MODEL_D] > 0).sum() > 1000
    rc.destroy()


@pytest.mark.parametrize("kern", [K.KERNEL_AUTO, K.KERNEL_GENERIC])
@pytest.mark.parametrize("name", __import__("golden_util").PARTICLE_CASES)
def test_particle_step_equals_the_reference_vertex_program_on_llvmpipe(name, kern):
    """octree_cuc_particles_update ("next" row 8f #2) against particle_vsh.c itself (tests/golden, transform feedback
    on llvmpipe): after one step and after `steps` steps kept on the device every position and speed is
    bit-identical; the parked count equals the host's end-of-simulation test.  Both traversals: the fast one (the
    1800-unit cube's grid is exact; axis-parallel particles fall back per particle) and the generic one."""
    import golden_util
    g = golden_util.load_particles(name)
    steps = int(g["steps"])
    rc = K.OctreeGlc(b"", device=0)
    rc.set_kernel(kern)
    rc.upload_octree(g["oct_s"])
    rc.particles_alloc_in(g["pos"], g["spd"])
    rc.particles_update(steps=1)
    pos, spd, parked = rc.particles_read_out()
    assert np.array_equal(pos.view(np.uint32), g["pos_1"].view(np.uint32))
    assert np.array_equal(spd.view(np.uint32), g["spd_1"].view(np.uint32))
    assert parked == int((g["spd_1"][:, 0] < -900).sum())
    rc.particles_update(steps=steps - 1)
    pos, spd, parked = rc.particles_read_out()
    assert np.array_equal(pos.view(np.uint32), g["pos_%d" % steps].view(np.uint32))
    assert np.array_equal(spd.view(np.uint32), g["spd_%d" % steps].view(np.uint32))
    assert parked == int((g["spd_%d" % steps][:, 0] < -900).sum())
    rc.destroy()


@pytest.mark.parametrize("kern", [K.KERNEL_AUTO, K.KERNEL_GENERIC])
@pytest.mark.parametrize("div", [K.DIV_GLSL, K.DIV_IEEE])
def test_particle_step_equals_the_oracle(scene_c1, div, kern):
    """Debris over the C1 room (1.3 M-node static tree), both division modes, both traversals, 8 chained steps,
    partial counts."""
    rng = np.random.default_rng(17)
    n = 50000
    idx = rng.integers(0, len(scene_c1.pnt_s), n)
    pos = (scene_c1.pnt_s[idx] + rng.normal(0, 5, (n, 3))).astype(np.float32)
    spd = rng.normal(0, 2.5, (n, 3)).astype(np.float32)
    spd[:1000, 0] = 0
    spd[1000:2000, 1] = 0.4       # gravity makes it exactly 0: a ray parallel to the y planes
    spd[2000:3000, 2] = 0
    rc = K.OctreeGlc(b"", device=0)
    rc.set_division(div)
    rc.set_kernel(kern)
    rc.upload_octree(scene_c1.oct_s)
    rc.particles_alloc_in(pos, spd)
    rc.particles_update(steps=8, count=n - 777)
    got_p, got_s, parked = rc.particles_read_out(count=n - 777)
    p, s = pos[:n - 777], spd[:n - 777]
    for _ in range(8):
        p, s, hit = O.particles(scene_c1.oct_s, p, s, div=div)
    assert np.array_equal(got_p.view(np.uint32), p.view(np.uint32))
    assert np.array_equal(got_s.view(np.uint32), s.view(np.uint32))
    assert parked == int((s[:, 0] < -900).sum()) and parked > 1000
    rc.destroy()


def test_dust_step_equals_the_reference_vertex_program_on_llvmpipe():
    import golden_util
    g = golden_util.load_particles("dust_box")
    steps = int(g["steps"])
    rc = K.OctreeGlc(b"", device=0)
    rc.particles_alloc_in(g["pos"], g["spd"], kind=K.DUST)
    rc.particles_update(kind=K.DUST, campos=tuple(g["campos"]), steps=1)
    pos, spd, _ = rc.particles_read_out(kind=K.DUST)
    assert np.array_equal(pos.view(np.uint32), g["pos_1"].view(np.uint32))
    assert np.array_equal(spd.view(np.uint32), g["spd_1"].view(np.uint32))
    rc.particles_update(kind=K.DUST, campos=tuple(g["campos"]), steps=steps - 1)
    pos, spd, _ = rc.particles_read_out(kind=K.DUST)
    assert np.array_equal(pos.view(np.uint32), g["pos_%d" % steps].view(np.uint32))
    assert np.array_equal(spd.view(np.uint32), g["spd_%d" % steps].view(np.uint32))
    rc.destroy()


@pytest.mark.parametrize("q", __import__("golden_util").PRESENT_QUALITIES)
def test_presentation_pass_equals_the_reference_window_on_llvmpipe(q):
    """octree_glc_update with presentation enabled ("next" row 8f #4b) against the window image the reference's own
    octree_glc_update leaves behind on llvmpipe (render target -> LINEAR-filtered quad -> crosshair), at every render
    scale, for both kernels: the pass is bit-exact given the frame; the frame is within the renderer's 1/255."""
    import golden_util
    sc, args, g = golden_util.load_present(q)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_present(True)
    for kern, name in KERNELS:
        rc.set_kernel(kern)
        rc.update(args["width"], args["height"], args["position"], args["angle"], quality=args["quality"],
                  shoot=args["shoot"])
        frame = rc.read_frame()
        assert np.abs(frame.astype(int) - g["frame"].astype(int)).max() <= parity.RGB_TOL, name
        win = rc.read_window()
        assert win.shape == g["window"].shape
        # the pass itself is exact: presenting the connector's own frame with the pinned oracle gives the same bytes
        assert np.array_equal(win, O.present(frame, O.uniforms(**args), args["width"], args["height"])), name
        if np.array_equal(frame, g["frame"]):
            assert np.array_equal(win, g["window"]), name
        assert np.abs(win.astype(int) - g["window"].astype(int)).max() <= parity.RGB_TOL, name
    rc.enable_present(False)
    rc.destroy()


def test_presentation_random_windows_equal_the_oracle(scene_random):
    """Random window sizes and render scales: the connector's window image equals the (llvmpipe-pinned) oracle's
    presentation of the connector's own frame, byte for byte."""
    rng = np.random.default_rng(77)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    rc.enable_present(True)
    for _ in range(12):
        q = int(rng.integers(4, 11))
        ww, wh = int(rng.integers(120, 900)), int(rng.integers(90, 600))
        pos, ang = (760.0 + float(rng.normal(0, 20)), 200.0, 420.0), (-0.05 + float(rng.normal(0, 0.2)), -0.12, 0.0)
        rc.update(ww, wh, pos, ang, quality=q)
        u = O.uniforms(ww, wh, pos, ang, quality=q)
        frame = rc.read_frame()
        assert frame.shape[:2] == (u.vp_h, u.vp_w)
        assert np.array_equal(rc.read_window(), O.present(frame, u, ww, wh)), (q, ww, wh)
    rc.destroy()


def test_presentation_at_full_size_equals_the_oracle(scene_c1):
    """1080p window at quality 10 (a copy + crosshair) and at quality 7 (2.5x upscale of a 768 x 432 frame)."""
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_c1)
    rc.enable_present(True)
    osc = O.OracleScene(scene_c1)
    for quality in (10, 7):
        rc.update(1920, 1080, *S.CAMERA_C1, quality=quality)
        u = O.uniforms(1920, 1080, *S.CAMERA_C1, quality=quality)
        frame = rc.read_frame()
        assert np.array_equal(rc.read_window(), O.present(frame, u, 1920, 1080))
        r = O.render(osc, u)
        assert np.abs(frame.astype(int) - r["rgba"].astype(int)).max() <= parity.RGB_TOL
    rc.destroy()


def test_slot_records_follow_tree_replacement_removals_and_appends():
    """The fast traversal reads slot records derived from the child array (octree_types.cuh, TreeDev::slot).  They
    must stay current through everything the reference API can do to a tree: a big tree replaced by a smaller,
    different one in the same buffers (stale nodes and stale parent links behind the new extent), leaves removed
    (a parent's slot zeroed, the child's own node never uploaded again), points appended one by one with only the
    touched 48-byte nodes uploaded (new nodes land on the old tree's stale nodes) -- every frame against the oracle,
    both kernels (the generic one reads the child array itself)."""
    big = S.make_random(40000, 6000, seed=21)
    small = S.make_random(9000, 1500, seed=22)
    pos, ang = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(big)
    _render_and_compare(big, 192, 108, pos, ang, rc=rc, divs=DIVS[:1])
    # a smaller, different level and figure over the old ones (whole-array uploads, like a level change)
    rc.upload_scene(small)
    _render_and_compare(small, 192, 108, pos, ang, rc=rc, divs=DIVS[:1])
    # zero-and-append on both trees, per-node uploads only
    rng = np.random.default_rng(23)
    trees, pts, cols, nrms = [], [], [], []
    for dyn, (p0, c0, n0) in enumerate(((small.pnt_s, small.col_s, small.nrm_s), (small.pnt_d, small.col_d, small.nrm_d))):
        t = S.HostOctree()
        t.insert_points(p0)
        victims = rng.permutation(len(p0))[:400]
        touched = []
        for v in victims:
            m, o = t.remove_point(p0[v])
            if o >= 0:
                touched.append(o)
        extra = np.clip((p0[rng.integers(0, len(p0), 700)] + rng.normal(0, 4.0, (700, 3))).astype(np.float32), 1.0, 1798.0)
        for k, q in enumerate(extra):
            touched.extend(int(j) for j in t.insert_point(q, len(p0) + k) if j > 0)
        p1 = np.concatenate([p0, extra])
        c1 = np.concatenate([c0, np.tile(np.array([[0.9, 0.4, 0.1]], np.float32), (len(extra), 1))])
        n1 = np.concatenate([n0, np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (len(extra), 1))])
        bt = K.DYNAMIC_OCTREE if dyn else K.STATIC_OCTREE
        t.upload_node_ranges(rc, sorted(set(touched)), bt)
        rc.upload_points(c1, K.DYNAMIC_COLOR if dyn else K.STATIC_COLOR, len(p0), len(p1))
        rc.upload_points(n1, K.DYNAMIC_NORMAL if dyn else K.STATIC_NORMAL, len(p0), len(p1))
        trees.append(t.nodes()), pts.append(p1), cols.append(c1), nrms.append(n1)
        # (the device extent is still the bigger tree's: its nodes behind the new tree are stale, like stale texels)
        assert np.array_equal(rc.download_octree(dynamic=bool(dyn))[:len(trees[-1])], trees[-1])
    final = S.Scene("edited", pts[0], cols[0], nrms[0], trees[0], pts[1], cols[1], nrms[1], trees[1])
    ref, _ = _render_and_compare(final, 192, 108, pos, ang, rc=rc, divs=DIVS[:1])
    assert ((ref["flags"] & O.FLAG_LEAF) > 0).sum() > 500 and (ref["aux"][..., K.AUX_MODEL_D] > 0).sum() > 20
    rc.destroy()


def test_tuning_switches_do_not_change_the_frame(scene_random):
    """octree_cuc_set_occupancy (resident CTAs per SM capped by shared-memory padding, incl. the > 48 KB opt-in) and
    octree_cuc_set_persisting_window (L2 access-policy window): frames, flags and hit indices stay identical."""
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_random)
    rc.enable_aux(True)
    pos, ang = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)

    def frame():
        rc.update(400, 230, pos, ang)
        f, a = rc.read_aux()
        return rc.read_frame().copy(), f, a

    want = frame()
    assert (want[1] & K.FLAG_LEAF).any()
    for cap in (5, 3, 1, 0):
        rc.set_occupancy(cap)
        for g, e in zip(frame(), want):
            assert np.array_equal(g, e), cap
    rc.enable_aux(False)
    want_rgba = want[0]
    for cap in (2, 1, 0):                       # the plain instantiation takes the large-shared-memory opt-in
        rc.set_occupancy(cap)
        rc.update(400, 230, pos, ang)
        assert np.array_equal(rc.read_frame(), want_rgba), cap
    for mb in (32, 0):
        rc.set_persisting_window(mb << 20)
        rc.update(400, 230, pos, ang)
        assert np.array_equal(rc.read_frame(), want_rgba), mb
    rc.destroy()


def test_tile_feedback_changes_the_order_not_the_frame(scene_c1):
    """Tile scheduling by measured cost (octree_cuc_set_tile_feedback): whatever order the tiles are launched in --
    image order, the order learned from the previous rendering of the view, an order inherited from another view,
    another tiling after a resize, more remembered views than slots, a sharded frame -- every byte of the frame and
    of the parity planes equals the frame rendered in image order."""
    views = [S.CAMERA_C1, ((760.0, 125.0, 225.0), (2.0, 0.3, 0.0)), ((700.0, 180.0, 300.0), (0.9, -0.4, 0.0))]
    views += [((700.0 + 3.0 * k, 150.0, 350.0), (0.4636, 0.0, 0.0)) for k in range(9)]   # > 8 remembered views
    sizes = [(640, 360), (333, 207), (640, 360)]
    want = {}
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(scene_c1)
    rc.enable_aux(True)
    rc.set_tile_feedback(False)
    for (W, H) in set(sizes):
        for i, (pos, ang) in enumerate(views):
            rc.update(W, H, pos, ang)
            want[(W, H, i)] = (rc.read_frame(), rc.read_aux())
    rc.set_tile_feedback(True)
    for rep in range(3):
        for (W, H) in sizes:
            for i, (pos, ang) in enumerate(views):
                rc.update(W, H, pos, ang)
                frame, (flags, aux) = rc.read_frame(), rc.read_aux()
                ref_frame, (ref_flags, ref_aux) = want[(W, H, i)]
                assert np.array_equal(frame, ref_frame), (rep, W, H, i)
                assert np.array_equal(flags, ref_flags) and np.array_equal(aux, ref_aux), (rep, W, H, i)
    # a shard of the frame: the untouched tiles keep the marker, the rendered ones equal the full frame
    rc.set_shard(1, 3, 64, 64)
    for rep in range(3):
        rc.update(640, 360, *views[0])
        got = rc.read_frame()
        own = S_tile_mask(640, 360, 64, 1, 3)
        assert np.array_equal(got[own], want[(640, 360, 0)][0][own]), rep
    rc.destroy()


def S_tile_mask(W, H, tile, rank, world):
    ty, tx = np.meshgrid(np.arange(H) // tile, np.arange(W) // tile, indexing="ij")
    tiles_x = (W + tile - 1) // tile
    return ((ty * tiles_x + tx) % world) == rank


def test_gpu_voxelise_and_bulk_build_equal_the_host_model():
    """octree_cuc_voxelise_and_build ("next" row 8f #3): the same survivors in the same order as the qmc rules
    (host voxeliser, itself byte-identical to the reference qmc binary in tests/test_host_model.py), the same
    colours / normals, and node-for-node the tree octree_insert_point builds from them; then the frame."""
    rng = np.random.default_rng(41)
    pos = (np.array([700.0, 100.0, 300.0]) + rng.uniform(0, 60, size=(150000, 3)) * np.array([1, 0.05, 1])).astype(
        np.float32)
    pos[:300] = rng.uniform(-80, 1900, size=(300, 3)).astype(np.float32)  # some outside the cube
    pos[300:600] = pos[600:900]                                            # exact duplicates
    col = rng.integers(0, 256, size=(len(pos), 3)).astype(np.uint8)
    nrm = rng.normal(size=(len(pos), 3)).astype(np.float32)
    hp, hc, hn = S.voxelise(pos, col, nrm)
    tree = S.HostOctree()
    tree.insert_points(hp)

    rc = K.OctreeGlc(b"", device=0)
    m, order, gp = rc.voxelise_and_build(pos, col, nrm, 1800, 12, dynamic=False)
    assert m == len(hp)
    assert np.array_equal(gp, hp)
    assert np.array_equal(pos[order], hp)
    gc, gn = rc.download_points(dynamic=False)
    assert np.array_equal(gc, hc) and np.array_equal(gn, hn)
    assert np.array_equal(rc.download_octree(dynamic=False), tree.nodes())
    # the voxeliser's flat files (qmc.c L266-327: .pnt / .nrm / .col / .rng) written from the GPU result are the
    # bytes the host voxeliser's are (which test_host_model.py compares with the reference qmc binary's files)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        S.write_flat(d + "/gpu", gp, gn, gc)
        S.write_flat(d + "/host", hp, hn, hc)
        for ext in ("pnt", "nrm", "col", "rng"):
            assert open(d + "/gpu." + ext, "rb").read() == open(d + "/host." + ext, "rb").read(), ext
        lp, lc, ln, lr = S.load_flat(d + "/gpu")
        assert np.array_equal(lp, hp) and len(lr) > 100
    # and it renders like the host-built scene
    e3 = np.zeros((0, 3), np.float32)
    sc = S.Scene("gpu-qmc", hp, hc, hn, tree.nodes(), e3, e3, e3, np.zeros((1, 12), np.int32))
    rc.upload_octree(np.zeros((1, 12), np.int32), dynamic=True)
    _render_and_compare(sc, 240, 135, (735.0, 140.0, 380.0), (0.05, -0.35, 0.0), rc=rc)
    # other depths and the dynamic model
    for levels in (5, 9):
        hp2, hc2, hn2 = S.voxelise(pos, col, nrm, 1800, levels)
        t2 = S.HostOctree(1800.0, levels)
        t2.insert_points(hp2)
        m2, order2, gp2 = rc.voxelise_and_build(pos, col, nrm, 1800, levels, dynamic=True)
        assert m2 == len(hp2) and np.array_equal(gp2, hp2)
        assert np.array_equal(rc.download_octree(dynamic=True), t2.nodes())
    rc.destroy()


needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference)")


@needs_ref
def test_batched_trace_lines_equal_the_reference_cpu_function(scene_c1, scene_random):
    """octree_cuc_trace_lines ("next" row 8f #4) against the reference's own compiled octree_trace_line
    (oracle/_ref, octree.c unmodified) on the same rays: camera rays, shadow-like rays from the light, rays with
    zero direction components (the CPU twin's (0,0,0,FLT_MAX) sentinel), rays from outside the cube; static and
    dynamic tree.  Hit index and leaf cube must be identical."""
    rng = np.random.default_rng(51)
    n = 40000
    org = np.stack([rng.uniform(600, 900, n), rng.uniform(100, 250, n), rng.uniform(100, 450, n)], axis=1).astype(
        np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:2000, 0] = 0.0                       # parallel to the x planes
    d[2000:4000, 1] = 0.0
    d[4000:5000, :2] = 0.0                  # along z only
    org[5000:7000] += np.float32(2500.0)    # outside the cube
    cases = [(scene_c1.pnt_s, scene_c1.oct_s, False), (scene_random.pnt_d, scene_random.oct_d, True)]
    rc = K.OctreeGlc(b"", device=0)
    for pts, nodes, dyn in cases:
        ref = O.RefOctree()
        ref.insert_points(pts)
        assert np.array_equal(ref.nodes(), nodes)
        rc.upload_octree(nodes, dynamic=dyn)
        dd = d.copy()
        aim = rng.integers(0, len(pts), n // 2)   # half of the rays aim at points of the model
        dd[n // 2:] = (np.asarray(pts)[aim] - org[n // 2:]).astype(np.float32)
        want_idx, want_tlf = ref.trace(org, dd)
        assert (want_idx != 0).sum() > 5000
        hit = want_idx != 0
        for kern in (K.KERNEL_AUTO, K.KERNEL_GENERIC):   # the fast traversal (exact grid) and the generic one
            rc.set_kernel(kern)
            got_idx, got_tlf = rc.trace_lines(org, dd, dynamic=dyn)
            assert np.array_equal(got_idx, want_idx), kern
            assert np.array_equal(got_tlf[hit], want_tlf[hit]), kern
    rc.set_kernel(K.KERNEL_AUTO)
    rc.destroy()
