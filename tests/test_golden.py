"""Golden vectors: frames the reference's UNMODIFIED shader produced on Mesa llvmpipe (tests/golden/).

They pin the oracle: the GLSL-division build must reproduce the shader's RGBA, hit model indices and shadow bits on
EVERY pixel (it does: Mesa lowers `/` to `* rcp`, which is the only arithmetic difference to the reference's compiled
C), and the IEEE-division build -- pinned separately against the reference's CPU twin in test_oracle.py -- may
differ from llvmpipe only on the handful of pixels where one rounding of a quotient flips a comparison."""
import numpy as np
import pytest

import golden_util
from oracle import qb_oracle as O


@pytest.mark.parametrize("name", golden_util.CASES)
def test_oracle_reproduces_the_reference_shader(name):
    sc, args, g = golden_util.load(name)
    if sc is None:
        pytest.skip("scene generator gives different bits on this CPU (hash mismatch); embedded cases still run")
    r = O.render(O.OracleScene(sc), O.uniforms(**args), div=O.DIV_GLSL)
    assert np.array_equal(r["rgba"], g["rgba"]), "RGBA differs from the shader on %d pixels" % (
        (r["rgba"] != g["rgba"]).any(axis=-1).sum())
    leaf = (r["flags"] & O.FLAG_LEAF) > 0
    disc = (r["flags"] & O.FLAG_DISC_ON) > 0
    assert np.array_equal(g["rgba"][..., 3] == 255, leaf | disc)          # alpha 255 <=> a leaf was returned
    assert np.array_equal(g["model_s"][leaf], r["aux"][leaf][:, 0])        # hit voxel index, static tree
    assert np.array_equal(g["model_d"][leaf], r["aux"][leaf][:, 1])        # hit voxel index, dynamic tree
    shaded = (r["flags"] & O.FLAG_SHADED) > 0
    assert np.array_equal(g["shadow"][shaded] != 0, (r["flags"][shaded] & O.FLAG_LIT) > 0)  # shadow visibility


@pytest.mark.parametrize("name", golden_util.CASES)
def test_ieee_oracle_is_within_a_few_pixels_of_the_shader(name):
    sc, args, g = golden_util.load(name)
    if sc is None:
        pytest.skip("scene generator gives different bits on this CPU (hash mismatch)")
    r = O.render(O.OracleScene(sc), O.uniforms(**args), div=O.DIV_IEEE)
    differing = int((r["rgba"] != g["rgba"]).any(axis=-1).sum())
    assert differing <= 4, differing


@pytest.mark.parametrize("name", golden_util.SKIN_CASES)
def test_skin_oracle_equals_the_reference_vertex_program(name):
    """oracle/skeleton_vsh_oracle.c against skeleton_vsh.c as llvmpipe ran it through transform feedback: with the
    driver's own ten bone rotations (sin / cos / acos are implementation-defined) every digit, normal and skinned
    position is bit-identical; with libm rotations the residual is the driver's acos polynomial (< 0.01 units)."""
    pos, nrm, g = golden_util.load_skin(name)
    d, no, pnt = O.skin(g["oldbones"], g["newbones"], pos, nrm, rotations=g["rotations"])
    assert np.array_equal(d, g["digits"])
    assert np.array_equal(no.view(np.uint32), g["normal_out"].view(np.uint32))
    assert np.array_equal(pnt.view(np.uint32), g["pnt"].view(np.uint32))
    libm = O.bone_rotations(g["oldbones"], g["newbones"])
    assert np.abs(libm - g["rotations"]).max() < 1e-4
    d2, no2, pnt2 = O.skin(g["oldbones"], g["newbones"], pos, nrm)
    assert np.abs(pnt2 - g["pnt"]).max() < 1e-2 and np.abs(no2 - g["normal_out"]).max() < 1e-3
    assert (d2 != g["digits"]).any(axis=1).mean() < 0.01
    if name == "skin_rest":
        assert np.array_equal(d2, g["digits"]) and np.array_equal(pnt2.view(np.uint32), g["pnt"].view(np.uint32))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", golden_util.PARTICLE_CASES)
def test_particle_oracle_equals_the_reference_vertex_program(name):
    """oracle/particle_vsh_oracle.c against particle_vsh.c as llvmpipe ran it through transform feedback: positions
    and speeds after one step and after `steps` chained steps are bit-identical, including the program's vec4(0.0)
    parallel-ray result and its missing texture-row wrap (particles_rowend: a level-3 node at a row end hides the
    half of a sheet that lies in its child slots 4..7)."""
    g = golden_util.load_particles(name)
    steps = int(g["steps"])
    p, s = g["pos"], g["spd"]
    for k in range(1, steps + 1):
        p, s, hit = O.particles(g["oct_s"], p, s)
        if k in (1, steps):
            assert np.array_equal(_bits(p), _bits(g["pos_%d" % k])), k
            assert np.array_equal(_bits(s), _bits(g["spd_%d" % k])), k
    if name == "particles_rowend":
        low = g["pos"][:, 2] < 1012.5
        stuck = g["spd_1"][:, 0] < -900
        assert stuck[low].sum() == 0 and stuck[~low].sum() > 1000


def test_dust_oracle_equals_the_reference_vertex_program():
    g = golden_util.load_particles("dust_box")
    steps = int(g["steps"])
    p, s = g["pos"], g["spd"]
    for k in range(1, steps + 1):
        p, s = O.dust(g["campos"], p, s)
        if k in (1, steps):
            assert np.array_equal(_bits(p), _bits(g["pos_%d" % k])), k
            assert np.array_equal(_bits(s), _bits(g["spd_%d" % k])), k


@pytest.mark.parametrize("q", golden_util.PRESENT_QUALITIES)
def test_present_oracle_equals_the_reference_window(q):
    """oracle/present_oracle.c against the window image of the reference's full octree_glc_update on llvmpipe
    (octree_glc.c L308-351: LINEAR-filtered textured quad + crosshair): every pixel, all four channels, at every
    render scale the engine offers; the oracle's own frame (RGB within 1/255 of the shader's) presents to within 1."""
    sc, args, g = golden_util.load_present(q)
    u = O.uniforms(**args)
    assert g["frame"].shape[:2] == (u.vp_h, u.vp_w)
    win = O.present(g["frame"], u, args["width"], args["height"])
    assert np.array_equal(win, g["window"])
    r = O.render(O.OracleScene(sc), u)
    own = O.present(r["rgba"], u, args["width"], args["height"])   # the oracle's frame is within +-1 of the shader's
    assert np.abs(own.astype(int) - g["window"].astype(int)).max() <= 1
    cy, cx = args["height"] // 2, args["width"] // 2
    assert (g["window"][cy - 1:cy + 1, cx - 1:cx + 1] == 255).all()
