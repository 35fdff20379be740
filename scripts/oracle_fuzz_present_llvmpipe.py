"""Randomised pin of the presentation oracle against the reference's full octree_glc_update on Mesa llvmpipe
(glsl_ref mode 40: render target -> LINEAR-filtered textured quad -> crosshair): random window sizes and render
scales.  Needs oracle/_ref (build container).  usage: python scripts/oracle_fuzz_present_llvmpipe.py [cases] [seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import qb_oracle as O
from qubatron_b200 import scene as S

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
assert O.have_glsl()
sc = S.make_random(3500, 700, seed=21)
t0 = time.time()
out = {"cases": cases, "first_seed": seed0, "window_pixels": 0, "differing_pixels": 0, "max_difference": 0}
for c in range(cases):
    rng = np.random.default_rng(seed0 + c)
    q = int(rng.integers(4, 11))
    ww, wh = int(rng.integers(120, 900)), int(rng.integers(90, 600))
    pos = (800.0 + float(rng.normal(0, 30)), 230.0 + float(rng.normal(0, 20)), 380.0 + float(rng.normal(0, 30)))
    u = O.uniforms(width=ww, height=wh, position=pos, angle=(-0.6 + float(rng.normal(0, 0.3)), -0.3, 0.0), quality=q,
                   shoot=int(rng.integers(0, 2)))
    frame, _ = O.glsl_render(sc, u, mode=0)
    window, _ = O.glsl_render(sc, u, mode=40, window=(ww, wh))
    mine = O.present(frame, u, ww, wh)
    d = np.abs(mine.astype(int) - window.astype(int)).max(axis=2)
    out["window_pixels"] += ww * wh
    out["differing_pixels"] += int((d > 0).sum())
    out["max_difference"] = max(out["max_difference"], int(d.max()))
    if d.max():
        print("case", c, "quality", q, "window", (ww, wh), "render", (u.vp_w, u.vp_h), list(u.dimensions), "differs on",
              int((d > 0).sum()), "pixels, max", int(d.max()), flush=True)
out["seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
