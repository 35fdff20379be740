"""Per-tile kernel times (each 64x64 tile rendered alone = latency-bound) for one pose of the bench workload."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from qubatron_b200 import connector as K
pose = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc, meta = bench.get_scene(1.0, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0); rc.upload_scene(sc)
pos, ang = sc.cameras[pose]
W, H = bench.WIDTH, bench.HEIGHT
tx, ty = (W + tile - 1) // tile, (H + tile - 1) // tile
n = tx * ty
def frame(): rc.update(W, H, pos, ang, 0.0, 10, bench.MAXLEVEL, bench.BASESIZE, 0); return rc.last_frame_ms()
rc.set_shard(0, 1, tile, tile)
for _ in range(3): full = frame()
times = np.zeros(n)
for t in range(n):
    rc.set_shard(t, n, tile, tile)
    frame(); times[t] = frame()
print(json.dumps({"pose": pose, "tile": tile, "full_frame_ms": full, "tiles": n, "tile_ms_sum": float(times.sum()),
                  "tile_ms_max": float(times.max()), "tile_ms_mean": float(times.mean()),
                  "tile_ms_p90": float(np.percentile(times, 90)), "argmax": int(times.argmax()),
                  "top5": [round(float(v), 4) for v in np.sort(times)[-5:]]}))
np.save(os.path.join(ROOT, "gpurun_out", "tile_times_pose%d.npy" % pose), times)
