"""Particle simulation steps (particle_vsh.c) over the bench level's static tree, state kept on the device."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K

sc, meta = bench.get_scene(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, 0, lambda: None)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
rng = np.random.default_rng(5)
pnt = np.asarray(sc.pnt_s)
idx = rng.integers(0, len(pnt), n)
pos = (pnt[idx] + rng.normal(0, 6, (n, 3))).astype(np.float32)
pos[:, 1] += 20.0
spd = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
rc = K.OctreeGlc(b"", device=0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rc.set_stream(stream.cuda_stream)
rc.upload_octree(sc.oct_s)
out = {"particles": n, "static_nodes": int(len(sc.oct_s)), "steps": []}
rc.particles_alloc_in(pos, spd)
rc.particles_update(steps=1)          # warm-up (one step of the real trajectory)
for k in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    rc.particles_update(steps=1)
    b.record(stream)
    b.synchronize()
    out["steps"].append(round(a.elapsed_time(b), 4))
_, spd_out, parked = rc.particles_read_out()
out["parked_after_31_steps"] = parked
out["first_step_ms"] = out["steps"][0]
out["mean_step_ms"] = float(np.mean(out["steps"]))
out["Mparticles_per_s_first_step"] = n / out["steps"][0] / 1e3
out["host_round_trip_bytes_per_step_in_the_reference"] = 48 * n
print(json.dumps(out))
rc.destroy()
