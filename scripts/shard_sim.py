"""What ONE rank of an N-GPU split renders, on one GPU: for world = 2, 4, 8 every rank's share of the tiles
(set_shard(r, world)) of each bench pose, cold L2, tile order learned -- the slowest rank's kernel is the kernel time
of the N-GPU frame (no fence, no NVLink).  Sweeps the resident-CTA cap (octree_cuc_set_occupancy).

    python scripts/shard_sim.py [caps, e.g. 0,5,4,3,2] [worlds, e.g. 4,8]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K

caps = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,5,4,3,2").split(",")]
worlds = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
sc, meta = bench.get_scene(1.0, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
W, H = bench.WIDTH, bench.HEIGHT
out = {"what": "kernel ms of the slowest rank's share per pose (one GPU rendering each rank's tiles in turn)", "rows": []}
for world in worlds:
    for cap in caps:
        rc.set_occupancy(cap)
        per_pose = []
        for pos, ang in sc.cameras:
            worst = 0.0
            for r in range(world):
                rc.set_shard(r, world, 64, 64)
                t = []
                for k in range(4):
                    flush.fill_(1)
                    torch.cuda.synchronize()
                    rc.update(W, H, pos, ang)
                    t.append(rc.last_frame_ms())
                worst = max(worst, float(np.median(t[1:])))
            per_pose.append(worst)
        row = {"world": world, "cta_cap": cap, "kernel_ms_by_pose": [round(v, 4) for v in per_pose],
               "mean_ms": float(np.mean(per_pose))}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
rc.destroy()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "shard_sim.json"), "w"), indent=1)
