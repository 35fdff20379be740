"""The lone-warp floor of a frame split over many GPUs: the heaviest 64x64 tile of a bench pose rendered ALONE
(set_shard(t, n_tiles)), L2 flushed before every launch like the bench does.

    python scripts/lone_tile.py <pose> [tile|-1] [reps]      tile -1 = find the heaviest one first

Prints one JSON line; QB_CUC_LIB selects the build, QB_PREFETCH=1 the prefetching variant's switch."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K

pose = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tile = int(sys.argv[2]) if len(sys.argv) > 2 else -1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
T = 64
sc, meta = bench.get_scene(1.0, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
pos, ang = sc.cameras[pose]
W, H = bench.WIDTH, bench.HEIGHT
n = ((W + T - 1) // T) * ((H + T - 1) // T)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def frame(cold=True):
    if cold:
        flush.fill_(1)
        torch.cuda.synchronize()
    rc.update(W, H, pos, ang, 0.0, 10, bench.MAXLEVEL, bench.BASESIZE, 0)
    return rc.last_frame_ms()


if tile < 0:
    times = np.zeros(n)
    for t in range(n):
        rc.set_shard(t, n, T, T)
        frame(False)
        times[t] = frame(False)
    tile = int(times.argmax())
rc.set_shard(tile, n, T, T)
frame()
cold = [frame(True) for _ in range(reps)]
warm = [frame(False) for _ in range(reps)]
rc.set_shard(0, 1, T, T)
frame()
full = [frame(True) for _ in range(4)]
print(json.dumps({"lib": os.environ.get("QB_CUC_LIB", "default"), "prefetch": os.environ.get("QB_PREFETCH"),
                  "pose": pose, "tile": tile, "lone_tile_ms_cold_l2": float(np.median(cold)),
                  "lone_tile_ms_warm_l2": float(np.median(warm)), "full_frame_ms": float(np.median(full))}))
rc.destroy()
