"""One process, one thread, N GPUs behind the C ABI (octree_cuc_set_gpus): the bench poses at 1080p and 2160p.

    python scripts/group_bench.py [max_gpus] [steps]

For n = 1, 2, 4, 8 (up to the GPUs of the box): octree replicated by the connector, tiles `mod n`, peer stores into
GPU 0's framebuffer, device-side completion flags.  Time = octree_cuc_last_step_ms (events on the primary's stream:
start of its kernel .. every device's tiles have arrived), L2 of every device flushed before each frame.  Prints one
JSON line per n; frames are CRC-checked against n = 1."""
import json
import os
import sys
import zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K

max_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
sc, meta = bench.get_scene(1.0, 0, lambda: None)
poses = sc.cameras
ndev = min(max_gpus, torch.cuda.device_count())
flush = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % k) for k in range(ndev)]
crc_ref = {}
n = 1
while n <= ndev:
    torch.cuda.set_device(0)
    rc = K.OctreeGlc(b"", device=0)
    if n > 1:
        rc.set_gpus(n, list(range(n)))
    rc.upload_scene(sc)
    rc.sync()
    out = {"gpus": n, "api": "octree_cuc_set_gpus (one process, one thread)", "steps": steps}
    for (W, H, tag) in ((1920, 1080, "1080p"), (3840, 2160, "2160p")):
        rc.enable_counters(True)
        rays = []
        for pos, ang in poses:
            rc.update(W, H, pos, ang)
            c = rc.read_counters()
            rays.append(c["rays_primary"] + c["rays_shadow"] + c["rays_disc"])
        rc.enable_counters(False)
        crcs = []
        for pos, ang in poses:
            rc.update(W, H, pos, ang)
            crcs.append(zlib.crc32(rc.read_frame().tobytes()))
        if n == 1:
            crc_ref[tag] = crcs
        ms, kms = [], []
        for i in range(steps + 4):
            for k in range(n):
                flush[k].fill_(1)
            torch.cuda.synchronize()
            for k in range(1, n):
                torch.cuda.synchronize(k)
            pos, ang = poses[i % len(poses)]
            rc.update(W, H, pos, ang)
            if i >= 4:
                ms.append(rc.last_step_ms())
                kms.append(rc.last_frame_ms())
        tot_rays = sum(rays[(i + 4) % len(poses)] for i in range(steps))
        out[tag] = {"ms_per_step": float(np.mean(ms)), "kernel_ms_slowest_device": float(np.mean(kms)),
                    "mrays_s": tot_rays / float(np.sum(ms)) / 1e3, "frame_crc32": crcs,
                    "frames_equal_one_gpu": crcs == crc_ref[tag],
                    "ms_by_pose": {str(p): float(np.mean([m for i, m in enumerate(ms) if (i + 4) % len(poses) == p]))
                                   for p in range(len(poses))}}
    print(json.dumps(out), flush=True)
    rc.destroy()
    n *= 2
