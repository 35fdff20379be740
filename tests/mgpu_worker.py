"""Worker of the multi-GPU parity test (launched by torchrun, one rank per GPU, NCCL).

Checks, against the CPU oracle on rank 0:
  * tile-sharded frame assembled by peer stores over NVLink (gather="p2p") and by NCCL reduce (gather="nccl");
  * range updates received by rank 0 through the reference API, broadcast as a blob and applied on every rank.
"""
import os
import sys
import zlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402
from oracle import qb_oracle as O  # noqa: E402
from qubatron_b200 import connector as K, multigpu, scene as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    sc = S.make_random(30000, 4000, seed=11)
    W, H = 300, 170
    pos, ang = (760.0, 200.0, 420.0), (-0.05, -0.12, 0.0)
    ok = True
    frame_crc = None
    for gather in ("p2p", "p2p_nccl", "nccl"):
        rc = K.OctreeGlc(b"", device=local)
        rc.set_stream(stream.cuda_stream)
        rc.upload_scene(sc)
        sh = multigpu.ShardedFrame(rc, W, H, rank, world, dev, gather=gather, tile=32)
        for it in range(3):
            rc.update(W, H, pos, ang, shoot=it & 1)
            sh.assemble()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            got = sh.read_frame()
            ref = O.render(O.OracleScene(sc), O.uniforms(W, H, pos, ang, shoot=0))
            crc = zlib.crc32(got.tobytes())
            if frame_crc is None:
                frame_crc = crc
            ok = ok and crc == frame_crc                 # every gather assembles the same bytes
            d = int(np.abs(got.astype(np.int16) - ref["rgba"].astype(np.int16)).max())
            print("gather=%s world=%d max rgba diff %d" % (gather, world, d), flush=True)
            ok = ok and d <= parity.RGB_TOL

        # pipelined readback (collective): frames alternate between shoot = 0 / 1 while their host copies are in
        # flight; every host buffer must end up holding ITS frame
        sync_frames = []
        for shoot in (0, 1):
            rc.update(W, H, pos, ang, shoot=shoot)
            sh.assemble()
            torch.cuda.synchronize()
            dist.barrier()
            sync_frames.append(sh.read_frame().copy() if rank == 0 else None)
            dist.barrier()
        bufs = [np.zeros((H, W, 4), np.uint8) for _ in range(2)]
        for b in bufs:
            torch.cuda.cudart().cudaHostRegister(b.ctypes.data, b.nbytes, 0)
        for it in range(8):
            rc.update(W, H, pos, ang, shoot=it & 1)
            sh.assemble()
            sh.read_frame_async(bufs[it & 1] if rank == 0 else None)
        if rank == 0:
            rc.wait_reads()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            same = all(np.array_equal(bufs[k], sync_frames[k]) for k in (0, 1))
            differ = not np.array_equal(sync_frames[0], sync_frames[1])
            print("gather=%s pipelined readback: buffers hold their frames %s (frames differ %s)" % (gather, same, differ),
                  flush=True)
            ok = ok and same and differ
        for b in bufs:
            torch.cuda.cudart().cudaHostUnregister(b.ctypes.data)

        # zero-and-append on rank 0 only, then broadcast
        tree = S.HostOctree()
        tree.insert_points(sc.pnt_s)
        col = sc.col_s.copy()
        if rank == 0:
            centre = sc.pnt_s[np.argmin(np.linalg.norm(sc.pnt_s - np.array([810.0, 180.0, 270.0], np.float32), axis=1))]
            near = np.nonzero(np.linalg.norm(sc.pnt_s - centre[None, :], axis=1) < 25.0)[0][:300]
            for v in near:
                m, o = tree.remove_point(sc.pnt_s[v])
                if o >= 0:
                    nodes = tree.nodes(copy=False)
                    rc.upload_texbuffer_data(nodes, K.GL_INT, len(nodes) * 48, 16, o * 48, (o + 1) * 48,
                                             K.STATIC_OCTREE)
                    col[m] = (1.0, 0.0, 1.0)
            # the whole colour array (360 KB): a BULK range, which flushes the queued node ranges on rank 0 --
            # the replication log must carry both
            rc.upload_points(col, K.STATIC_COLOR, 0, len(col))
        nbytes = sh.broadcast_updates(dev)
        rc.update(W, H, pos, ang)
        sh.assemble()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            final = S.Scene("upd", sc.pnt_s, col, sc.nrm_s, tree.nodes(), sc.pnt_d, sc.col_d, sc.nrm_d, sc.oct_d)
            ref2 = O.render(O.OracleScene(final), O.uniforms(W, H, pos, ang))
            got2 = sh.read_frame()
            d2 = int(np.abs(got2.astype(np.int16) - ref2["rgba"].astype(np.int16)).max())
            changed = int((ref2["rgba"] != ref["rgba"]).any(axis=-1).sum())
            print("gather=%s after broadcast of %d blob bytes: max rgba diff %d, %d pixels changed by the edit"
                  % (gather, nbytes, d2, changed), flush=True)
            ok = ok and d2 <= parity.RGB_TOL and changed > 0
        sh.close()
        dist.barrier()
        rc.destroy()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
