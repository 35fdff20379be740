"""Particle simulation steps (particle_vsh.c) and batched CPU ray queries (octree_trace_line) over the bench level's
static tree:  python scripts/particle_bench.py [scale] [particles] [kernel: 0 auto = fast traversal, 1 generic]"""
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
kern = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rc.set_kernel(kern)
out = {"particles": n, "static_nodes": int(len(sc.oct_s)), "traversal": {0: "fast (exact grid)", 1: "generic"}[kern],
       "steps": []}
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
# batched octree_trace_line queries: the same origins, random directions; the call includes the copies of the rays
# to the device and of the results back (44 bytes per ray)
import time
d = rng.normal(size=(n, 3)).astype(np.float32)
rc.trace_lines(pos, d)
t = []
for _ in range(5):
    t0 = time.time()
    idx_out, _tlf = rc.trace_lines(pos, d)
    t.append(time.time() - t0)
out["trace_lines_call_ms"] = round(1e3 * float(np.median(t)), 3)
out["trace_lines_hits"] = int((idx_out != 0).sum())
out["trace_lines_Mrays_per_s_incl_copies"] = n / float(np.median(t)) / 1e6
print(json.dumps(out))
rc.destroy()
