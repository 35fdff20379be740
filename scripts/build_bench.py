"""GPU tree build from paths vs the host's sequential insert on the bench level's ~10 M-point figure."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K, scene as S
sc, meta = bench.get_scene(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, 0, lambda: None)
pts = np.asarray(sc.pnt_d)
n = len(pts)
t = time.time(); paths = S.octant_paths(pts); t_paths = time.time() - t
host = S.HostOctree()
t = time.time(); host.insert_paths(paths); t_host = time.time() - t
rc = K.OctreeGlc(b"", device=0)
t = time.time(); nodes = rc.build_octree_from_paths(paths); t_gpu_host = time.time() - t
ok = np.array_equal(rc.download_octree(), host.nodes())
# paths already on the device (what a GPU skinning pass would leave behind)
d = [torch.from_numpy(np.ascontiguousarray(paths[:, a:a + 4])).cuda() for a in (0, 4, 8)]
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t = time.time()
    rc.build_octree_from_device_paths(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n)
    ts.append(time.time() - t)
ok2 = np.array_equal(rc.download_octree(), host.nodes())
print(json.dumps({"points": n, "nodes": nodes, "identical_to_host_insert": bool(ok and ok2),
                  "host_insert_paths_s": t_host, "gpu_build_from_host_paths_s": t_gpu_host,
                  "gpu_build_from_device_paths_ms": 1e3 * float(np.median(ts)),
                  "h2d_bytes_paths": int(paths.nbytes), "numpy_path_digits_s": t_paths}))
rc.destroy()
