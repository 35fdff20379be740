"""Shared parity check: CUDA connector output vs the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): per-pixel hit voxel index and shadow visibility
bit-exact (we compare ALL flag bits and all six aux planes exactly), RGB within
+-1/255 per channel (tolerance stated here: RGB_TOL)."""
import numpy as np

RGB_TOL = 1  # +-1/255 per channel


def compare(rgba, flags, aux, ref, what=""):
    """Returns a dict of mismatch counts; raises AssertionError with a diagnosis on failure."""
    out = {
        "flags_mismatch": int((flags != ref["flags"]).sum()),
        "aux_mismatch": int((aux != ref["aux"]).any(axis=-1).sum()),
        "rgba_maxdiff": int(np.abs(rgba.astype(np.int16) - ref["rgba"].astype(np.int16)).max()),
        "rgba_offby1": int((rgba != ref["rgba"]).any(axis=-1).sum()),
        "pixels": int(flags.size),
    }
    msg = []
    if out["flags_mismatch"]:
        ys, xs = np.nonzero(flags != ref["flags"])
        msg.append("flags differ on %d pixels, first (x=%d,y=%d) got %d want %d" % (
            out["flags_mismatch"], xs[0], ys[0], flags[ys[0], xs[0]], ref["flags"][ys[0], xs[0]]))
    if out["aux_mismatch"]:
        ys, xs = np.nonzero((aux != ref["aux"]).any(axis=-1))
        msg.append("aux differ on %d pixels, first (x=%d,y=%d) got %s want %s" % (
            out["aux_mismatch"], xs[0], ys[0], aux[ys[0], xs[0]].tolist(), ref["aux"][ys[0], xs[0]].tolist()))
    if out["rgba_maxdiff"] > RGB_TOL:
        msg.append("rgba max abs diff %d > %d" % (out["rgba_maxdiff"], RGB_TOL))
    assert not msg, "%s: %s" % (what, "; ".join(msg))
    return out
