"""Two frames of skin -> build -> render on the bench figure, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from qubatron_b200 import connector as K, scene as S
sc, meta = bench.get_scene(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, 0, lambda: None)
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
rc.skeleton_alloc_in(np.asarray(sc.pnt_d), np.asarray(sc.nrm_d))
by = float(S._terrain_height(np.float32(760.0), np.float32(230.0)))
for p in (0.5, 1.0, 1.5):
    rc.skeleton_update(*S.zombie_bones(base=(760.0, by, 230.0), pose=p), build_tree=True)
    rc.update(1920, 1080, *sc.cameras[0])
rc.sync()
rc.destroy()
