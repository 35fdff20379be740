"""Randomised sweep of the HOST build of the fast kernel (tests/host_emu) against the oracle -- the CPU twin of
scripts/parity_fuzz.py: random clouds, cameras (inside, outside, axis-parallel central rays), light angles, depths,
render scales, exact base cubes; both division modes; flags, hit indices and node ids bit-exact, RGBA identical.
It checks the LOGIC of the kernel source without a GPU (code generation is the GPU parity tests' business).
usage: python scripts/emu_fuzz.py [cases] [first_seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
import host_emu as E
from oracle import qb_oracle as O
from qubatron_b200 import scene as S


def fuzz_case(seed):
    """(scene, W, H, pos, angle, kwargs) of one random case; the cube is one the fast kernel takes"""
    rng = np.random.default_rng(seed)
    levels = int(rng.choice([6, 9, 11, 12]))
    basesize = float(rng.choice([1800.0, 1800.0, 2048.0, 1000.0]))
    sc = S.make_random(int(rng.integers(500, 6000)), int(rng.integers(0, 1500)), seed=seed, levels=levels,
                       basesize=basesize, clustered=bool(rng.integers(0, 2)))
    pts = sc.pnt_s
    target = pts[rng.integers(0, len(pts))]
    inside = rng.random() < 0.7
    pos = (target + rng.normal(0, 120, 3)).astype(np.float32) if inside else rng.uniform(-600, basesize + 600, 3).astype(np.float32)
    d = target - pos
    yaw = float(np.arctan2(d[0], -d[2])) + float(rng.normal(0, 0.2))
    pitch = float(np.arctan2(d[1], np.hypot(d[0], d[2]))) + float(rng.normal(0, 0.2))
    if rng.random() < 0.15:
        yaw, pitch = float(rng.choice([0.0, np.pi / 2, np.pi])), 0.0
        pos = np.round(pos)
    if rng.random() < 0.1:  # a camera ON a face / edge / corner of the base cube, or on a grid plane inside it
        k = rng.integers(1, 4)
        ax = rng.choice(3, size=k, replace=False)
        pos = pos.copy()
        pos[ax] = rng.choice([0.0, basesize, basesize / 2, basesize / 4], size=k).astype(np.float32)
    W, H = int(rng.choice([96, 160, 201, 256])), int(rng.choice([64, 90, 113, 144]))
    kw = dict(lighta=float(rng.uniform(0, 6.28)), quality=int(rng.choice([10, 10, 8, 7, 5])), maxlevel=levels,
              basesize=basesize, shoot=int(rng.integers(0, 2)))
    return sc, W, H, tuple(float(v) for v in pos), (yaw, pitch, 0.0), kw


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0 = time.time()
    frames = 0
    pixels = 0
    for c in range(cases):
        sc, W, H, pos, ang, kw = fuzz_case(seed0 + c)
        osc, esc = O.OracleScene(sc), E.EmuScene(sc)
        for ediv, odiv, dname in ((0, O.DIV_GLSL, "glsl"), (1, O.DIV_IEEE, "ieee")):
            ref = O.render(osc, O.uniforms(W, H, pos, ang, **kw), div=odiv)
            r = E.render(esc, W, H, pos, ang, div=ediv, **kw)
            what = "case %d seed %d host-emu/%s" % (c, seed0 + c, dname)
            parity.compare(r["rgba"], r["flags"], r["aux"], ref, what=what)
            assert np.array_equal(r["rgba"], ref["rgba"]), what + ": rgba differs"
            frames += 1
            pixels += r["flags"].size
    print(json.dumps({"cases": cases, "first_seed": seed0, "frames_compared": frames, "pixels": pixels,
                      "seconds": round(time.time() - t0, 1)}))
