"""Randomised parity sweep on the GPU box: random clouds, cameras, light angles, depths, render scales and base cubes;
every frame of both kernels in both division modes against the oracle (tests/parity.py rules: hit indices, shadow
bits and node ids bit-exact, RGBA within 1/255), plus the work counters.
usage: python scripts/parity_fuzz.py [cases] [first_seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from oracle import qb_oracle as O
from qubatron_b200 import connector as K, scene as S

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
KERNELS = [(K.KERNEL_GENERIC, "generic"), (K.KERNEL_FAST, "fast")]
DIVS = [(K.DIV_GLSL, O.DIV_GLSL, "glsl"), (K.DIV_IEEE, O.DIV_IEEE, "ieee")]


def grid_is_exact(basesize, maxlevel):
    """basesize = m * 2^e with bits(m) + maxlevel <= 23 (the fast kernel's condition)"""
    m = int(np.float32(basesize).view(np.uint32)) & 0x7FFFFF | 0x800000
    while m % 2 == 0:
        m //= 2
    return m.bit_length() + maxlevel <= 23


t0 = time.time()
frames = 0
worst = 0
for c in range(cases):
    rng = np.random.default_rng(seed0 + c)
    levels = int(rng.choice([6, 9, 11, 12]))
    basesize = float(rng.choice([1800.0, 1800.0, 2048.0, 1000.0, 1234.567]))
    sc = S.make_random(int(rng.integers(500, 6000)), int(rng.integers(0, 1500)), seed=seed0 + c, levels=levels,
                       basesize=basesize, clustered=bool(rng.integers(0, 2)))
    pts = sc.pnt_s
    target = pts[rng.integers(0, len(pts))]
    inside = rng.random() < 0.7
    pos = (target + rng.normal(0, 120, 3)).astype(np.float32) if inside else rng.uniform(-600, basesize + 600, 3).astype(np.float32)
    d = target - pos
    yaw = float(np.arctan2(d[0], -d[2])) + float(rng.normal(0, 0.2))           # angle 0 looks along -z, +yaw turns to +x
    pitch = float(np.arctan2(d[1], np.hypot(d[0], d[2]))) + float(rng.normal(0, 0.2))
    if rng.random() < 0.15:
        yaw, pitch = float(rng.choice([0.0, np.pi / 2, np.pi])), 0.0             # axis-parallel central rays
        pos = np.round(pos)
    W, H = int(rng.choice([96, 160, 201, 256])), int(rng.choice([64, 90, 113, 144]))
    kw = dict(lighta=float(rng.uniform(0, 6.28)), quality=int(rng.choice([10, 10, 8, 7, 5])), maxlevel=levels,
              basesize=basesize, shoot=int(rng.integers(0, 2)))
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_aux(True)
    rc.enable_counters(True)
    osc = O.OracleScene(sc)
    exact = grid_is_exact(basesize, levels)
    for kdiv, odiv, dname in DIVS:
        ref = O.render(osc, O.uniforms(W, H, tuple(pos), (yaw, pitch, 0.0), **kw), div=odiv)
        rc.set_division(kdiv)
        for kern, name in KERNELS:
            if kern == K.KERNEL_FAST and not exact:
                continue
            rc.set_kernel(kern)
            rc.update(W, H, tuple(pos), (yaw, pitch, 0.0), **kw)
            rgba = rc.read_frame()
            flags, aux = rc.read_aux()
            what = "case %d seed %d %s/%s" % (c, seed0 + c, name, dname)
            parity.compare(rgba, flags, aux, ref, what=what)
            assert rc.read_counters() == ref["counters"], what
            worst = max(worst, int(np.abs(rgba.astype(int) - ref["rgba"].astype(int)).max()))
            frames += 1
    rc.destroy()
print(json.dumps({"cases": cases, "first_seed": seed0, "frames_compared": frames, "max_rgba_difference": worst,
                  "seconds": round(time.time() - t0, 1)}))
