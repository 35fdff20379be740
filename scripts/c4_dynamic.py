"""BASELINE configs[3] (C4): dynamic scene.  Per frame: the ~10 M-point figure is re-pathed and the dynamic tree
rebuilt on the host (octree_reset + octree_insert_path, qubatron.c L439-452 -- a "next" row, not timed), then the
connector receives what the engine sends it every frame and renders:
    full DYNAMIC_OCTREE (qubatron.c L508-516), DYNAMIC_NORMAL [0, n) from the skinning output buffer (L521-529),
    a punch-hole batch on the static tree (~500 zeroed slots + ~500 appended paths -> 48-byte node uploads +
    one colour / normal sub-range, modelutil.c L429-546),
    octree_glc_update at 1080p.
Reports the time split upload vs render, with pageable host buffers (what the engine has) and with the two
per-frame buffers page-locked (octree_cuc_pin_host_buffer).

    python scripts/c4_dynamic.py [scale] [frames]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from qubatron_b200 import connector as K, scene as S  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, H = 1920, 1080

sc, meta = bench.get_scene(scale, 0, lambda: None)
pos, ang = sc.cameras[0]
rc = K.OctreeGlc(b"", device=0)
rc.upload_scene(sc)
rc.sync()
rc.take_upload_ms()

# host state the engine would own
stat = S.HostOctree()
t0 = time.time()
stat.insert_points(np.asarray(sc.pnt_s))
print("static host tree rebuilt in %.1f s (%d nodes)" % (time.time() - t0, len(stat)), file=sys.stderr)
xs_sorted = np.maximum.accumulate(np.asarray(sc.pnt_s[:, 0]))  # monotone envelope of the x-major order
col_s = np.array(sc.col_s)
nrm_s = np.array(sc.nrm_s)
fig = np.array(sc.pnt_d)
fig_n = np.array(sc.nrm_d)
n = len(fig)
dyn = S.HostOctree()
rng = np.random.default_rng(4)
# fixed-size per-frame buffers (the engine reuses skelglc.nrm_out and dynaoctr.octs)
nrm_out = np.empty_like(fig_n)
node_cap = int(len(sc.oct_d) * 1.3)
dyn_nodes = np.zeros((node_cap, 12), np.int32)

results = {}
for mode in ("pageable", "pinned"):
    if mode == "pinned":
        rc.pin_host_buffer(nrm_out)
        rc.pin_host_buffer(dyn_nodes)
    rows = []
    for f in range(frames):
        # ---- host side, untimed: skin + re-path + rebuild (next-row work)
        moved = (fig + np.array([2.0 * f, 0.0, -1.5 * f], np.float32)).astype(np.float32)
        dyn.reset()
        dyn.insert_paths(S.octant_paths(moved))
        ln = len(dyn)
        assert ln <= node_cap
        dyn_nodes[:ln] = dyn.nodes(copy=False)
        nrm_out[...] = fig_n
        # punch-hole batch on the static tree near the figure
        centre = np.array([760.0 + 3 * f, 62.0, 300.0], np.float32)
        # points are x-major sorted (qmc): bracket the x slab first
        lo_i = int(np.searchsorted(xs_sorted, centre[0] - 30.0))
        hi_i = int(np.searchsorted(xs_sorted, centre[0] + 30.0))
        slab = np.asarray(sc.pnt_s[lo_i:hi_i])
        cand = lo_i + np.nonzero(np.linalg.norm(slab - centre[None, :], axis=1) < 30.0)[0]
        victims = cand[rng.permutation(len(cand))[:500]] if len(cand) else []
        edits = []
        touched = []
        for v in victims:
            m, o = stat.remove_point(sc.pnt_s[v])
            if o >= 0:
                edits.append(o)
                touched.append(m)
        for m in touched:
            newp = (np.asarray(sc.pnt_s[m]) + rng.normal(0, 2.0, 3)).astype(np.float32)
            newp = np.clip(newp, 1.0, 1798.0)
            edits.extend(int(j) for j in stat.insert_point(newp, m) if j > 0)
            col_s[m] += 0.2
        snodes = stat.nodes(copy=False)
        rc.sync()
        rc.take_upload_ms()
        # ---- what the connector sees every frame (timed)
        t0 = time.time()
        rc.upload_texbuffer_data(dyn_nodes, K.GL_INT, ln * 48, 16, 0, ln * 48, K.DYNAMIC_OCTREE)
        rc.upload_texbuffer_data(nrm_out, K.GL_FLOAT, n * 12, 12, 0, n * 12, K.DYNAMIC_NORMAL)
        t_bulk = time.time() - t0
        t0 = time.time()
        for o in edits:
            rc.upload_texbuffer_data(snodes, K.GL_INT, len(snodes) * 48, 16, o * 48, (o + 1) * 48, K.STATIC_OCTREE)
        if touched:
            lo, hi = min(touched), max(touched) + 1
            rc.upload_points(col_s, K.STATIC_COLOR, lo, hi)
        t_edit = time.time() - t0
        t0 = time.time()
        rc.update(W, H, pos, ang, 0.0, 10, 12, 1800.0, 1)
        render_ms = rc.last_frame_ms()   # waits for the frame
        t_frame = time.time() - t0
        rows.append({"dyn_nodes": ln, "bulk_bytes": ln * 48 + n * 12, "bulk_upload_ms": 1e3 * t_bulk,
                     "bulk_GBps": (ln * 48 + n * 12) / t_bulk / 1e9, "node_ranges": len(edits),
                     "range_upload_calls_ms": 1e3 * t_edit, "update_call_to_frame_done_ms": 1e3 * t_frame,
                     "render_kernel_ms": render_ms})
    results[mode] = rows
    if mode == "pinned":
        rc.unpin_host_buffer(nrm_out)
        rc.unpin_host_buffer(dyn_nodes)

summary = {"config": "C4 dynamic scene: %d dynamic points, %d static points, 1080p" % (n, len(sc.pnt_s)),
           "frames": results,
           "median": {m: {k: float(np.median([r[k] for r in rows])) for k in rows[0]} for m, rows in results.items()}}
print(json.dumps(summary))
rc.destroy()
