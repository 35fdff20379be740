"""Multi-GPU behind the C ABI (octree_cuc_set_gpus, octree_cuc_set_fence), through the connector exactly as a C host
would drive it: one thread, one octree_glc_update per frame.

Every test runs on a ONE-GPU box as well: a group may name the same device several times, so that N connectors
share the GPU and the whole mechanism -- replicated uploads, tiles `mod N`, every connector's kernel storing into the
first one's framebuffer, the device-side completion fence, the "previous frame consumed" gate -- is exercised there
too.  With >= 2 GPUs visible the same tests also run across real devices (peer stores over NVLink).

Bar: frames of a group are IDENTICAL to the single-connector frames (same kernels, disjoint tiles), which in turn
match the oracle bit for bit on flags / hit indices and within +-1/255 on RGB (tests/parity.py)."""
import os
import subprocess

import numpy as np
import pytest

import parity
from oracle import qb_oracle as O
from qubatron_b200 import connector as K
from qubatron_b200 import scene as S

pytestmark = pytest.mark.gpu

POS, ANG = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]                       # shards sharing GPU 0
    if n >= 2:
        sets.append(list(range(min(n, 8))))          # one shard per GPU
    return sets


def _group(devices):
    rc = K.OctreeGlc(b"", device=devices[0])
    rc.set_gpus(len(devices), devices)
    assert rc.gpu_count() == len(devices)
    return rc


def _frame(rc, W, H, pos=POS, ang=ANG, **kw):
    rc.update(W, H, pos, ang, **kw)
    rgba = rc.read_frame().copy()
    flags, aux = rc.read_aux()
    return rgba, flags, aux


def test_group_frames_equal_single_connector_and_oracle(scene_random):
    W, H = 300, 170                                   # ragged against the 32-pixel shard tiles used below
    single = K.OctreeGlc(b"", device=0)
    single.upload_scene(scene_random)
    single.enable_aux(True)
    single.enable_counters(True)
    want = _frame(single, W, H)
    want_c = single.read_counters()
    ref = O.render(O.OracleScene(scene_random), O.uniforms(W, H, POS, ANG))
    parity.compare(*want, ref, what="single")
    for devices in _device_sets():
        rc = _group(devices)
        rc.upload_scene(scene_random)
        rc.enable_aux(True)
        rc.enable_counters(True)
        for kern in (K.KERNEL_FAST, K.KERNEL_GENERIC):
            rc.set_kernel(kern)
            single.set_kernel(kern)
            for it in range(4):                       # consecutive frames: the consumed gate and the flags cycle
                kw = dict(shoot=it & 1)
                got = _frame(rc, W, H, **kw)
                exp = _frame(single, W, H, **kw)
                for g, e, name in zip(got, exp, ("rgba", "flags", "aux")):
                    assert np.array_equal(g, e), "%s differs, devices %s kernel %d frame %d" % (name, devices, kern, it)
            assert rc.read_counters() == single.read_counters()
        assert rc.last_step_ms() >= 0.0 and rc.last_frame_ms() > 0.0
        # a batch of views (configs[4] shape): tiles of every view are spread over the group
        pos = np.array([POS, (700.0, 150.0, 350.0), (820.0, 260.0, 500.0)], np.float32)
        ang = np.array([ANG, (0.4636, 0.0, 0.0), (2.0, -0.3, 0.0)], np.float32)
        rc.set_kernel(K.KERNEL_AUTO)
        single.set_kernel(K.KERNEL_AUTO)
        rc.update_views(160, 96, pos, ang)
        single.update_views(160, 96, pos, ang)
        assert np.array_equal(rc.read_frame(views=3), single.read_frame(views=3))
        assert np.array_equal(rc.read_aux(views=3)[1], single.read_aux(views=3)[1])
        rc.destroy()
    assert want_c == ref["counters"]
    single.destroy()


def test_group_range_updates_bulk_and_growth(scene_random):
    """The zero-and-append flow of modelutil_punch_hole (modelutil.c L429-546) on a group: hundreds of 48-byte node
    ranges, a colour range larger than the batching threshold (bulk path), and appended points / nodes that force
    the device arrays to grow -- every device must end up with the same model."""
    sc = scene_random
    W, H = 256, 144
    for devices in _device_sets()[:1] + _device_sets()[2:]:
        rc = _group(devices)
        rc.upload_scene(sc)
        rc.enable_aux(True)
        tree = S.HostOctree()
        tree.insert_points(sc.pnt_s)
        col = sc.col_s.copy()
        centre = sc.pnt_s[np.argmin(np.linalg.norm(sc.pnt_s - np.array([810.0, 180.0, 270.0], np.float32), axis=1))]
        near = np.nonzero(np.linalg.norm(sc.pnt_s - centre[None, :], axis=1) < 25.0)[0][:300]
        rc.update(W, H, POS, ANG)                     # a frame before the edit: batched ranges wait for the next one
        for v in near:
            m, o = tree.remove_point(sc.pnt_s[v])
            if o >= 0:
                nodes = tree.nodes(copy=False)
                rc.upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, o * 48, (o + 1) * 48, K.STATIC_OCTREE)
                col[m] = (1.0, 0.0, 1.0)
        rc.upload_points(col, K.STATIC_COLOR, 0, len(col))      # 360 KB: the bulk path
        # append: new points, new nodes -> both arrays outgrow their capacity (+25 %)
        rng = np.random.default_rng(5)
        extra = (sc.pnt_s[rng.integers(0, len(sc.pnt_s), 20000)] + rng.normal(0, 6.0, (20000, 3))).astype(np.float32)
        extra = np.clip(extra, 1.0, 1799.0)
        n0, nodes0 = len(sc.pnt_s), len(tree)
        tree.insert_points(extra, first_modind=n0)
        pnt2 = np.concatenate([sc.pnt_s, extra])
        col2 = np.concatenate([col, np.tile(np.array([[0.2, 1.0, 0.3]], np.float32), (len(extra), 1))])
        nrm2 = np.concatenate([sc.nrm_s, np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (len(extra), 1))])
        nodes = tree.nodes()
        assert len(nodes) > nodes0 * 1.3
        rc.upload_points(col2, K.STATIC_COLOR, n0, len(col2))
        rc.upload_points(nrm2, K.STATIC_NORMAL, n0, len(nrm2))
        rc.upload_octree(nodes, dynamic=False, start_node=0, end_node=len(nodes))
        final = S.Scene("upd", pnt2, col2, nrm2, nodes, sc.pnt_d, sc.col_d, sc.nrm_d, sc.oct_d)
        ref = O.render(O.OracleScene(final), O.uniforms(W, H, POS, ANG))
        for it in range(2):
            parity.compare(*_frame(rc, W, H), ref, what="group %s after updates" % devices)
        rc.destroy()


def test_group_skinning_and_tree_build_on_every_device(scene_c1):
    """octree_cuc_skeleton_update on a group: every device skins and builds for itself (nothing is sent around);
    the frame equals the single-connector frame of the same pipeline (which test_parity_gpu checks against the
    oracle's frame of the host pipeline)."""
    pos, col, nrm = S.zombie_raw(base=(760.0, 100.0, 230.0), spacing=0.6, shells=2)
    pos, colf, nrm = S.voxelise(pos, col, nrm)
    ob, nb = S.zombie_bones(pose=1.5)
    frames = []
    for devices in [[0], [0, 0, 0]] + _device_sets()[2:]:
        rc = K.OctreeGlc(b"", device=0)
        if len(devices) > 1:
            rc.set_gpus(len(devices), devices)
        rc.upload_points(scene_c1.col_s, K.STATIC_COLOR)
        rc.upload_points(scene_c1.nrm_s, K.STATIC_NORMAL)
        rc.upload_octree(scene_c1.oct_s)
        rc.upload_points(colf, K.DYNAMIC_COLOR)
        rc.skeleton_alloc_in(pos, nrm)
        nodes = rc.skeleton_update(ob, nb, build_tree=True)
        assert nodes > 1000
        rc.enable_aux(True)
        frames.append(_frame(rc, 320, 180, *S.CAMERA_C1))
        rc.destroy()
    for f in frames[1:]:
        for g, e in zip(f, frames[0]):
            assert np.array_equal(g, e)
    assert (frames[0][2][..., K.AUX_MODEL_D] > 0).sum() > 1000    # the figure is in view


def test_two_connectors_one_gpu_fence_blob_and_staged_readback(scene_random):
    """What the torchrun ranks do, in ONE process on ONE GPU with raw device pointers in place of IPC handles: two
    connectors, rank 1 renders into rank 0's framebuffer, device-side fence (octree_cuc_set_fence), range updates
    through the replication log (export_pending -> apply_blob), pipelined host readback (read_frame_staged)."""
    import torch
    sc = scene_random
    W, H = 300, 170
    rcs = [K.OctreeGlc(b"", device=0) for _ in range(2)]
    for r in rcs:
        r.upload_scene(sc)
    for k, r in enumerate(rcs):
        r.set_shard(k, 2, 32, 32)
    rcs[0].reserve_frame(W, H, 1)
    rcs[0].enable_replication_log(True)
    ptrs = [r.fence_device() for r in rcs]
    for k, r in enumerate(rcs):
        r.set_fence(k, 2, ptrs)
    rcs[1].set_frame_target(rcs[0].frame_device(), W)

    def both(**kw):
        for r in rcs:                                  # rank 0 first (ranks share the device)
            r.update(W, H, POS, ANG, **kw)

    osc = O.OracleScene(sc)
    refs = [O.render(osc, O.uniforms(W, H, POS, ANG, shoot=s))["rgba"] for s in (0, 1)]
    assert not np.array_equal(refs[0], refs[1])
    bufs = [np.zeros((H, W, 4), np.uint8) for _ in range(2)]
    for b in bufs:
        torch.cuda.cudart().cudaHostRegister(b.ctypes.data, b.nbytes, 0)
    for it in range(8):                                # frames alternate while their host copies are in flight
        both(shoot=it & 1)
        rcs[0].read_frame_staged(bufs[it & 1])
    rcs[0].wait_reads()
    for k in (0, 1):
        assert np.abs(bufs[k].astype(np.int16) - refs[k].astype(np.int16)).max() <= parity.RGB_TOL
    for b in bufs:
        torch.cuda.cudart().cudaHostUnregister(b.ctypes.data)

    # zero-and-append on rank 0 only, with a bulk colour range; rank 1 gets it as a blob
    tree = S.HostOctree()
    tree.insert_points(sc.pnt_s)
    col = sc.col_s.copy()
    centre = sc.pnt_s[np.argmin(np.linalg.norm(sc.pnt_s - np.array([810.0, 180.0, 270.0], np.float32), axis=1))]
    near = np.nonzero(np.linalg.norm(sc.pnt_s - centre[None, :], axis=1) < 25.0)[0][:300]
    for v in near:
        m, o = tree.remove_point(sc.pnt_s[v])
        if o >= 0:
            nodes = tree.nodes(copy=False)
            rcs[0].upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, o * 48, (o + 1) * 48, K.STATIC_OCTREE)
            col[m] = (1.0, 0.0, 1.0)
    rcs[0].upload_points(col, K.STATIC_COLOR, 0, len(col))       # > 256 KB: flushes the queued node ranges on rank 0
    blob = rcs[0].export_pending()
    assert len(blob) > len(col) * 12
    # applied from DEVICE memory, as the receive buffer of the NCCL broadcast is (octree_cuc_apply_blob_device);
    # a third connector takes the same blob through the host path and must end up identical
    blob_dev = torch.from_numpy(blob.copy()).cuda()
    rcs[1].apply_blob_device(blob_dev.data_ptr(), len(blob))
    rcs[1].sync()
    third = K.OctreeGlc(b"", device=0)
    third.upload_scene(sc)
    third.apply_blob(blob)
    for a, b in zip(third.download_points(False), rcs[1].download_points(False)):
        assert np.array_equal(a, b)
    assert np.array_equal(third.download_octree(False), rcs[1].download_octree(False))
    third.destroy()
    assert len(rcs[0].export_pending()) == 16                    # drained
    both()
    final = S.Scene("upd", sc.pnt_s, col, sc.nrm_s, tree.nodes(), sc.pnt_d, sc.col_d, sc.nrm_d, sc.oct_d)
    ref2 = O.render(O.OracleScene(final), O.uniforms(W, H, POS, ANG))["rgba"]
    got2 = rcs[0].read_frame()
    assert np.abs(got2.astype(np.int16) - ref2.astype(np.int16)).max() <= parity.RGB_TOL
    assert (ref2 != refs[0]).any(axis=-1).sum() > 0
    for a, b in zip(rcs[0].download_points(False), rcs[1].download_points(False)):
        assert np.array_equal(a, b)
    assert np.array_equal(rcs[0].download_octree(False), rcs[1].download_octree(False))
    rcs[1].set_frame_target(0, 0)
    for r in rcs:
        r.sync()
    for r in rcs:
        r.set_fence(0, 1, None)
        r.destroy()


def test_malformed_blob_is_rejected():
    """apply_blob validates every descriptor before touching anything (it aborts the process: run it in a child)."""
    code = r'''
import numpy as np, sys
from qubatron_b200 import connector as K, multigpu as M
rc = K.OctreeGlc(b"", device=0)
blob = M.pack_ranges([(K.STATIC_OCTREE, 0, np.zeros(48, np.uint8))])
d = blob[16:16 + M.DESC_DTYPE.itemsize].view(M.DESC_DTYPE)
d["%s"] = %d
rc.apply_blob(blob)
print("APPLIED")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for field, value in (("buftype", 9), ("nwords", 1 << 20), ("src_word", 77), ("dst_word", 1)):
        r = subprocess.run([os.sys.executable, "-c", code % (field, value)], cwd=root, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode != 0 and "APPLIED" not in r.stdout and "apply_blob" in r.stdout, r.stdout[-500:]


def test_c_host_drives_a_group(tmp_path):
    """examples/host_demo --gpus N: the reference's call sequence from C on N shards; the PPM equals N = 1."""
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "examples")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    assert r.returncode == 0, r.stdout
    outs = []
    ns = [1, 3] + ([torch.cuda.device_count()] if torch.cuda.device_count() >= 2 else [])
    for n in ns:
        d = tmp_path / ("n%d" % n)
        d.mkdir()
        args = [os.path.join(root, "examples", "host_demo"), "--gpus", str(n)]
        if n == 3:
            args += ["--same-device"]
        r = subprocess.run(args, cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        print(r.stdout[-1500:])
        assert r.returncode == 0, r.stdout[-1500:]
        outs.append((d / "frame.ppm").read_bytes())
    for o in outs[1:]:
        assert o == outs[0]
