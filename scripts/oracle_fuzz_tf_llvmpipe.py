"""Randomised pin of the transform-feedback oracles against the reference programs on Mesa llvmpipe
(oracle/_ref/glsl_ref modes 20-22, 30, 31): skinning with random skeleton poses (the driver's own bone rotations fed to
the oracle, see oracle/skeleton_vsh_oracle.c), particle steps over random trees, dust steps.
Needs oracle/_ref (build container).  usage: python scripts/oracle_fuzz_tf_llvmpipe.py [cases] [first_seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import qb_oracle as O
from qubatron_b200 import scene as S

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
assert O.have_glsl()
bits = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
t0 = time.time()
out = {"cases": cases, "first_seed": seed0, "skin_points": 0, "skin_mismatches": 0, "particles": 0,
       "particle_mismatches": 0, "particles_stuck": 0, "dust": 0, "dust_mismatches": 0}
fig_p, _, fig_n = S.zombie_raw(spacing=0.9, shells=3)
for c in range(cases):
    rng = np.random.default_rng(seed0 + c)
    # ---- skinning: random pose of the capsule figure's skeleton, random subset of its points
    sel = rng.choice(len(fig_p), 6000, replace=False)
    pos, nrm = fig_p[sel], fig_n[sel]
    ob, nb = S.zombie_bones(pose=float(rng.uniform(0, 3)), shift=tuple(rng.normal(0, 8, 3)))
    nb = nb.copy()
    nb[:, :3] += rng.normal(0, 2.0, (20, 3)).astype(np.float32)        # every joint jittered
    nb[::2, 3] = rng.normal(0, 0.4, 10).astype(np.float32)             # random twists
    ob = ob.copy()
    ob[::2, 3] *= np.float32(rng.uniform(0.6, 1.6))                     # radius of effect
    gd, gn, _ = O.glsl_skin(ob, nb, pos, nrm)
    _, gp, _ = O.glsl_skin(ob, nb, pos, nrm, points=True)
    rot, seen = O.glsl_bone_rotations(ob, nb, pos, nrm)
    d, no, po = O.skin(ob, nb, pos, nrm, rotations=rot)
    # A point out of range of every bone pair keeps its position and takes `corner_normals[0]`, which the shader never
    # wrote (skeleton_vsh.c L172): undefined in GLSL, garbage on llvmpipe, 0 in the oracle and the connector.  Its
    # normal is not compared.
    loose = (bits(po) == bits(pos)).all(axis=1) & (no == 0).all(axis=1)
    bad = int(((d != gd).any(axis=1) | ((bits(no) != bits(gn)).any(axis=1) & ~loose) | (bits(po) != bits(gp)).any(axis=1)).sum())
    out["skin_points"] += len(pos)
    out["skin_mismatches"] += bad
    out["skin_points_out_of_every_bones_range"] = out.get("skin_points_out_of_every_bones_range", 0) + int(loose.sum())
    # ---- particles over a random tree
    levels = int(rng.choice([8, 10, 12]))
    sc = S.make_random(int(rng.integers(2000, 12000)), 0, seed=seed0 + c, levels=levels, clustered=True)
    n = 3000
    idx = rng.integers(0, len(sc.pnt_s), n)
    pp = (sc.pnt_s[idx] + rng.normal(0, 7, (n, 3))).astype(np.float32)
    ps = rng.normal(0, 2.5, (n, 3)).astype(np.float32)
    ps[:100, 0] = 0
    ps[100:200, 1] = 0.4
    ps[200:300, 2] = 0
    ps[300:330] = 0
    pp[330:400] = np.round(pp[330:400])
    p, s = pp, ps
    for step in range(3):
        gpp, gps, _ = O.glsl_particles(sc.oct_s, p, s, maxlevel=levels)
        op, os_, hit = O.particles(sc.oct_s, p, s, maxlevel=levels)
        out["particle_mismatches"] += int(((bits(op) != bits(gpp)).any(axis=1) | (bits(os_) != bits(gps)).any(axis=1)).sum())
        out["particles"] += n
        out["particles_stuck"] += int(hit.sum())
        p, s = gpp, gps
    # ---- dust
    cam = tuple(rng.uniform(300, 800, 3))
    dp = np.stack([rng.uniform(380, 820, n), rng.uniform(-10, 310, n), rng.uniform(-10, 410, n)], axis=1).astype(np.float32)
    ds = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
    gdp, gds, _ = O.glsl_particles(sc.oct_s, dp, ds, dust_campos=cam)
    odp, ods = O.dust(cam, dp, ds)
    out["dust"] += n
    out["dust_mismatches"] += int(((bits(odp) != bits(gdp)).any(axis=1) | (bits(ods) != bits(gds)).any(axis=1)).sum())
    if bad:
        print("case", c, "skin mismatches", bad, flush=True)
out["seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
