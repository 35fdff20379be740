"""Loading of tests/golden/*.npz (frames of the reference shader on llvmpipe; see tests/golden/make_golden.py)."""
import ast
import hashlib
import os

import numpy as np

from qubatron_b200 import scene as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["octtest5", "cloud_a", "cloud_b", "cloud_disc_l9", "c1_start_pose", "c1_inside_sphere"]


def scene_hash(sc):
    h = hashlib.sha256()
    for a in (sc.oct_s, sc.oct_d, sc.col_s, sc.nrm_s, sc.col_d, sc.nrm_d):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


_c1 = None


def load(name):
    """Returns (scene, args dict, golden dict) or (None, ...) when a hashed scene cannot be regenerated here."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    args = ast.literal_eval(str(g["args"]))
    if "oct_s" in g.files:
        e3 = np.zeros((0, 3), np.float32)
        sc = S.Scene(name, e3, g["col_s"], g["nrm_s"], g["oct_s"], e3, g["col_d"], g["nrm_d"], g["oct_d"])
    else:
        global _c1
        if _c1 is None:
            _c1 = S.make_c1()
        sc = _c1
        if scene_hash(sc) != str(g["scene_sha256"]):
            return None, args, g
    return sc, args, g


SKIN_CASES = ["skin_rest", "skin_walk", "skin_bent"]


def load_skin(name):
    """(positions, normals, golden dict) of a skinning fixture (tests/golden/make_golden_skin.py)."""
    i = np.load(os.path.join(GOLDEN, "skin_inputs.npz"))
    return i["positions"], i["normals"], np.load(os.path.join(GOLDEN, name + ".npz"))


PARTICLE_CASES = ["particles_cloud", "particles_rowend"]


def load_particles(name):
    """golden dict of a particle / dust fixture (tests/golden/make_golden_particles.py)."""
    return np.load(os.path.join(GOLDEN, name + ".npz"))


PRESENT_QUALITIES = [10, 9, 8, 7, 6, 5, 4]


def load_present(q):
    """(scene, args, golden) of a presentation fixture (tests/golden/make_golden_present.py); scene = cloud_a's."""
    sc, _, _ = load("cloud_a")
    g = np.load(os.path.join(GOLDEN, "present_q%d.npz" % q))
    return sc, ast.literal_eval(str(g["args"])), g
