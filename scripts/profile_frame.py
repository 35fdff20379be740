"""Render the bench workload's poses a few times (for ncu): python scripts/profile_frame.py [scale] [frames] [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qubatron_b200 import connector as K

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sc, meta = bench.get_scene(scale, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
rc.set_kernel(kernel)
for i in range(frames):
    pos, ang = sc.cameras[i % len(sc.cameras)]
    rc.update(bench.WIDTH, bench.HEIGHT, pos, ang, 0.0, 10, bench.MAXLEVEL, bench.BASESIZE, 0)
    print(i, rc.last_frame_ms(), flush=True)
rc.destroy()
