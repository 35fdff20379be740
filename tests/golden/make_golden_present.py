"""Generate tests/golden/present_q*.npz: the WINDOW image the reference's full octree_glc_update leaves behind --
octree_vsh.c / octree_fsh.c into the 2048 x 2048 render target, then texquad_vsh.c / texquad_fsh.c + the crosshair
(octree_glc.c L308-351) -- run headless on Mesa llvmpipe by oracle/_ref/glsl_ref mode 40.

Run in the build container only (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_golden_present.py

The scene is the one embedded in cloud_a.npz.  Each fixture: the octree_glc_update arguments, `frame` (mode 0, the
render target's (int)ow x (int)oh texels) and `window` (mode 40, (int)width x (int)height), both uint8 RGBA with
row 0 = bottom.  One fixture per render scale the engine offers (qubatron.c L314-320: quality 4..10 -> 4, 3.5, 3,
2.5, 2, 1.5, 1), with window sizes that do and do not divide evenly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_util  # noqa: E402
from oracle import qb_oracle as O  # noqa: E402

CASES = {10: (192, 108), 9: (300, 170), 8: (400, 226), 7: (480, 270), 6: (501, 333), 5: (700, 394), 4: (512, 288)}

if __name__ == "__main__":
    O.build(ref=True)
    assert O.have_glsl(), "oracle/_ref/glsl_ref or the Mesa libGL is missing"
    sc, _, _ = golden_util.load("cloud_a")
    for q, (ww, wh) in CASES.items():
        args = dict(width=ww, height=wh, position=(800.0, 230.0, 380.0), angle=(-0.6, -0.3, 0.0), quality=q, shoot=1)
        u = O.uniforms(**args)
        frame, _ = O.glsl_render(sc, u, mode=0)
        window, info = O.glsl_render(sc, u, mode=40, window=(ww, wh))
        path = os.path.join(HERE, "present_q%d.npz" % q)
        np.savez_compressed(path, args=np.array(repr(args)), frame=frame, window=window,
                            renderer=np.array(info["renderer"] + " / " + info["version"]))
        print("present_q%-2d frame %s window %s  %d KB" % (q, frame.shape, window.shape, os.path.getsize(path) // 1024))
