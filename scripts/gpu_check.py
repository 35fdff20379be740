"""Developer script for the GPU box: parity of both kernels on several scenes + quick timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from qubatron_b200 import connector as K, scene as S
from oracle import qb_oracle as O
import parity

def check(sc, W, H, pos, ang, name, **kw):
    rc = K.OctreeGlc(b"", device=0)
    rc.upload_scene(sc)
    rc.enable_aux(True)
    rc.enable_counters(True)
    u = O.uniforms(W, H, pos, ang, **kw)
    t = time.time(); ref = O.render(O.OracleScene(sc), u); tor = time.time() - t
    res = {}
    for kern, kn in ((K.KERNEL_GENERIC, "generic"), (K.KERNEL_FAST, "fast")):
        rc.set_kernel(kern)
        rc.update(W, H, pos, ang, **kw)
        rgba = rc.read_frame(); flags, aux = rc.read_aux(); cnt = rc.read_counters()
        try:
            out = parity.compare(rgba, flags, aux, ref, what="%s/%s" % (name, kn))
            out["counters_equal"] = cnt == ref["counters"]
            if not out["counters_equal"]:
                out["counters"] = cnt; out["counters_ref"] = ref["counters"]
        except AssertionError as e:
            out = {"FAIL": str(e)}
        # timing without aux/counters
        rc.enable_aux(False); rc.enable_counters(False)
        for _ in range(3): rc.update(W, H, pos, ang, **kw)
        ms = []
        for _ in range(10):
            rc.update(W, H, pos, ang, **kw); ms.append(rc.last_frame_ms())
        rc.enable_aux(True); rc.enable_counters(True)
        out["ms"] = float(np.median(ms))
        rays = ref["counters"]["rays_primary"] + ref["counters"]["rays_shadow"] + ref["counters"]["rays_disc"]
        out["mrays_s"] = rays / out["ms"] / 1e3
        res[kn] = out
    res["oracle_s"] = tor
    res["ref_counters"] = ref["counters"]
    rc.destroy()
    print(name, json.dumps(res), flush=True)
    return res

if __name__ == "__main__":
    print(K.load_library().octree_cuc_version())
    check(S.make_test5(), 320, 200, (900.0, 900.0, 3000.0), (0.0, 0.0, 0.0), "test5")
    check(S.make_random(30000, 4000, seed=11), 256, 128, (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0), "random")
    c1 = S.make_c1()
    check(c1, 640, 360, *S.CAMERA_C1, name="c1")
    check(c1, 640, 360, (700.0, 150.0, 350.0), (0.4636, 0.0, 0.0), name="c1_shoot", shoot=1, lighta=1.0)
    check(c1, 1920, 1080, *S.CAMERA_C1, name="c1_1080p")
    check(c1, 640, 360, (760.0, 125.0, 225.0), (2.0, 0.3, 0.0), name="c1_inside_sphere")
    check(c1, 640, 360, (1200.0, 300.0, 900.0), (-0.9, -0.2, 0.0), name="c1_far_lightdisc")
