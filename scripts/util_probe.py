"""Experiment (needs a library built with EXTRA=-DQB_UTIL_PROBE, QB_CUC_LIB pointing at it): how much of the fast
kernel's lane idleness comes from lanes that finished their pixel before the warp's longest lane."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from qubatron_b200 import connector as K
sc, meta = bench.get_scene(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
rc.enable_counters(True)
out = []
for i, (pos, ang) in enumerate(sc.cameras):
    rc.update(1920, 1080, pos, ang)
    c = rc.read_counters()
    out.append({"pose": i, "lane_iterations": c["hits"], "warp_slots": c["discards"], "utilisation": c["hits"] / max(c["discards"], 1)})
print(json.dumps(out))
rc.destroy()
