"""The per-frame dynamic-model pipeline (skin -> tree build -> render) kept on the device, on the bench level's
~10 M-point figure, next to the host-side stages the reference runs between its two GL passes
(qubatron.c L425-452: read back digits and normals, octree_reset + octree_insert_path, upload tree and normals)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from qubatron_b200 import connector as K, scene as S

sc, meta = bench.get_scene(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, 0, lambda: None)
pos, nrm = np.asarray(sc.pnt_d), np.asarray(sc.nrm_d)
n = len(pos)
rc = K.OctreeGlc(b"", device=0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rc.set_stream(stream.cuda_stream)
rc.upload_scene(sc)
rc.skeleton_alloc_in(pos, nrm)
W, H = 1920, 1080
cam = sc.cameras[0] if hasattr(sc, "cameras") else S.CAMERA_C1


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


by = float(S._terrain_height(np.float32(760.0), np.float32(230.0)))   # where make_c2 stands the figure
poses = [S.zombie_bones(base=(760.0, by, 230.0), pose=p, shift=(2.0 * p, 0.0, -1.0 * p)) for p in (0.3, 0.8, 1.3, 1.8)]
state = {"i": 0}


def bones():
    state["i"] += 1
    return poses[state["i"] % len(poses)]


out = {"points": n}
out["skin_ms"] = timed(lambda: rc.skeleton_update(*bones(), build_tree=False))
out["skin_and_build_ms"] = timed(lambda: rc.skeleton_update(*bones(), build_tree=True))


def frame():
    rc.skeleton_update(*bones(), build_tree=True)
    rc.update(W, H, cam[0], cam[1])


out["skin_build_render_1080p_ms"] = timed(frame)
out["render_only_1080p_ms"] = timed(lambda: rc.update(W, H, cam[0], cam[1]))
# algorithmic bytes of the skin kernel: 24 B in (position, normal), 48 B digits + 16 B normal + 12 B point out
out["skin_bytes_per_point"] = 24 + 48 + 16 + 12
out["skin_GBps"] = out["skin_bytes_per_point"] * n / (out["skin_ms"] * 1e-3) / 1e9

# the reference's host-side stages for the same frame
ob, nb = poses[1]
nodes = rc.skeleton_update(ob, nb, build_tree=True)
t = time.time(); digits, nrm_out, pnt_out = rc.skeleton_read_out(n); out["readback_digits_normals_s"] = time.time() - t
host = S.HostOctree()
t = time.time(); host.insert_paths(digits); out["host_insert_paths_s"] = time.time() - t
want = host.nodes()
out["nodes"] = int(nodes)
out["device_tree_identical_to_host_insert"] = bool(nodes == len(want) and np.array_equal(rc.download_octree(dynamic=True), want))
rc2 = K.OctreeGlc(b"", device=0)
t = time.time()
rc2.upload_octree(want, dynamic=True)
rc2.upload_points(nrm_out, K.DYNAMIC_NORMAL)
rc2.sync()
out["host_upload_tree_normals_s"] = time.time() - t
out["host_pipeline_bytes_d2h"] = int(digits.nbytes + nrm_out.nbytes)
out["host_pipeline_bytes_h2d"] = int(want.nbytes + nrm_out.nbytes)
print(json.dumps(out))
rc.destroy()
rc2.destroy()
