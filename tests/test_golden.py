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
