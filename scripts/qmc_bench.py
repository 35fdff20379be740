"""GPU qmc + bulk static-tree build on the bench level's raw points vs the host path."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from qubatron_b200 import connector as K, scene as S
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
# raw points of the level, before voxelisation (regenerated: the cache only keeps the voxelised arrays)
t = time.time()
S_build = S.build_scene
raw = {}
def capture(name, static_raw, dynamic_raw=None, **kw):
    raw["s"] = static_raw
    raise StopIteration
S.build_scene = capture
try:
    S.make_c2(scale=scale, zombie=False)
except StopIteration:
    pass
S.build_scene = S_build
pos, col, nrm = raw["s"]
t_gen = time.time() - t
n = len(pos)
t = time.time(); hp, hc, hn = S.voxelise(pos, col, nrm); t_vox = time.time() - t
t = time.time(); tree = S.HostOctree(); tree.insert_points(hp); t_tree = time.time() - t
rc = K.OctreeGlc(b"", device=0)
rc.voxelise_and_build(pos[:1000], col[:1000], nrm[:1000], want_order=False)   # warm the pool / cub
t = time.time(); m, order, gp = rc.voxelise_and_build(pos, col, nrm, 1800, 12, dynamic=False, want_order=True); t_gpu = time.time() - t
same = bool(m == len(hp) and np.array_equal(gp, hp) and np.array_equal(rc.download_octree(dynamic=False), tree.nodes()))
t = time.time(); rc.voxelise_and_build(pos, col, nrm, 1800, 12, dynamic=False, want_order=False); t_gpu2 = time.time() - t
print(json.dumps({"raw_points": n, "survivors": int(m), "nodes": len(tree), "identical_to_host": same,
                  "host_voxelise_s": t_vox, "host_tree_build_s": t_tree,
                  "gpu_total_s_incl_h2d_and_order_readback": t_gpu, "gpu_total_s_incl_h2d": t_gpu2,
                  "h2d_bytes": int(pos.nbytes + col.nbytes + nrm.nbytes), "generate_raw_s": t_gen}))
rc.destroy()
