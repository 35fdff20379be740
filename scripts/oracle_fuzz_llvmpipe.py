"""Randomised pin of the ORACLE against the reference shader itself: random scenes and views rendered by the unmodified
octree_fsh.c on Mesa llvmpipe (oracle/_ref/glsl_ref) and by oracle/octree_fsh_oracle.c (GLSL division mode).
Needs oracle/_ref (build container).  Viewport sizes are ones where llvmpipe interpolates `coord` exactly (DESIGN
section 9, parity note).  usage: python scripts/oracle_fuzz_llvmpipe.py [cases] [first_seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import qb_oracle as O
from qubatron_b200 import scene as S

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 50
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
SIZES = [(192, 108), (160, 100), (160, 120), (200, 112), (256, 144), (128, 72)]
assert O.have_glsl()
t0 = time.time()
out = {"cases": cases, "first_seed": seed0, "pixels": 0, "rgba_differing_pixels": 0, "rgba_max_difference": 0,
       "aux_cases": 0, "model_or_shadow_mismatches": 0, "leaf_pixels": 0}
for c in range(cases):
    rng = np.random.default_rng(seed0 + c)
    levels = int(rng.choice([6, 9, 11, 12]))
    basesize = float(rng.choice([1800.0, 1800.0, 2048.0, 1000.0, 1234.567]))
    sc = S.make_random(int(rng.integers(500, 5000)), int(rng.integers(0, 1200)), seed=seed0 + c, levels=levels,
                       basesize=basesize, clustered=bool(rng.integers(0, 2)))
    target = sc.pnt_s[rng.integers(0, len(sc.pnt_s))]
    pos = (target + rng.normal(0, 120, 3)).astype(np.float32) if rng.random() < 0.75 else \
        rng.uniform(-600, basesize + 600, 3).astype(np.float32)
    d = target - pos
    yaw = float(np.arctan2(d[0], -d[2])) + float(rng.normal(0, 0.2))
    pitch = float(np.arctan2(d[1], np.hypot(d[0], d[2]))) + float(rng.normal(0, 0.2))
    if rng.random() < 0.15:
        yaw, pitch, pos = float(rng.choice([0.0, np.pi / 2, np.pi])), 0.0, np.round(pos)
    W, H = SIZES[int(rng.integers(0, len(SIZES)))]
    u = O.uniforms(W, H, tuple(pos), (yaw, pitch, 0.0), lighta=float(rng.uniform(0, 6.28)), maxlevel=levels,
                   basesize=basesize, shoot=int(rng.integers(0, 2)))
    ref = O.render(O.OracleScene(sc), u)
    rgba, _ = O.glsl_render(sc, u, mode=0)
    diff = np.abs(rgba.astype(int) - ref["rgba"].astype(int)).max(axis=2)
    out["pixels"] += W * H
    out["leaf_pixels"] += int(((ref["flags"] & 2) > 0).sum())
    out["rgba_differing_pixels"] += int((diff > 0).sum())
    out["rgba_max_difference"] = max(out["rgba_max_difference"], int(diff.max()))
    if c % 4 == 0:   # the aux dumps cost three more shader runs
        leaf = (ref["flags"] & 2) > 0
        shaded = (ref["flags"] & 4) > 0
        ms, _ = O.glsl_render(sc, u, mode=1)
        md, _ = O.glsl_render(sc, u, mode=2)
        sh, _ = O.glsl_render(sc, u, mode=3)
        bad = int((ms[leaf] != ref["aux"][..., 0][leaf]).sum() + (md[leaf] != ref["aux"][..., 1][leaf]).sum()
                  + (sh[shaded] != ((ref["flags"][shaded] & 8) > 0)).sum())
        out["aux_cases"] += 1
        out["model_or_shadow_mismatches"] += bad
    if diff.max() > 0:
        # not a traversal difference if the oracle, given the coord llvmpipe's rasteriser produced, agrees
        cx, cy = O.glsl_coords(sc, u)
        quats = O.glsl_quats(sc, u)   # ... and the view quaternions its sin / cos produced
        ref2 = O.render_with_coords(O.OracleScene(sc), u, cx, cy, quats=quats)
        diff2 = np.abs(rgba.astype(int) - ref2["rgba"].astype(int)).max(axis=2)
        off = float(max(np.abs(cx - (np.arange(W, dtype=np.float32) + 0.5)[None, :]).max(),
                        np.abs(cy - (np.arange(H, dtype=np.float32) + 0.5)[:, None]).max()))
        out.setdefault("cases_with_inexact_coord", 0)
        out.setdefault("differing_pixels_left_with_llvmpipe_coord", 0)
        out["cases_with_inexact_coord"] += 1
        out["differing_pixels_left_with_llvmpipe_coord"] += int((diff2 > 0).sum())
        print("case", c, "seed", seed0 + c, "differs on", int((diff > 0).sum()), "pixels (max %d);" % int(diff.max()),
              "llvmpipe's coord is off by up to %.2e;" % off, "with llvmpipe's coord and view quaternions:",
              int((diff2 > 0).sum()), "differ", flush=True)
out["seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
