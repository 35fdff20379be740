"""Image-tile sharding of one frame over the GPUs of a box (one process per GPU, torch.distributed).

The path shards by pixels: the octree and the point arrays are replicated, every rank renders the tiles whose
row-major index is congruent to its rank modulo the world size (interleaving averages out sky-vs-geometry
imbalance), and the only per-frame exchange is the assembly of the tiles:

  gather="p2p"   rank 0 exports its framebuffer as a CUDA IPC handle; the other ranks map it and their render
                 kernels store their pixels straight into it over NVLink (peer stores), so the transfer overlaps
                 the traversal; completion is a device-side flag per rank (octree_cuc_set_fence), no collective.
  gather="nccl"  every rank renders into a zero-initialised full-size buffer of its own and the frame is the
                 NCCL reduce(SUM) of the buffers viewed as int32 (tiles are disjoint, so the sum is the union).

Range updates ("zero-and-append", modelutil.c L429-501) are received by rank 0 through the reference API, exported
as one packed blob (octree_cuc_export_pending), broadcast, and applied by every rank (octree_cuc_apply_blob).

The tile ownership rule and the blob format are plain host logic, kept here so that CPU (gloo) tests cover them.
"""
import numpy as np

BLOCK_W, BLOCK_H = 16, 8  # CTA footprint of the render kernels (octree_types.cuh)


def tile_grid(width, height, tile_w, tile_h):
    return (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h


def tile_owner_map(width, height, world, tile_w=64, tile_h=64):
    """int32 [H,W]: rank that renders each pixel (same rule as render kernels: tile index % world)."""
    if tile_w % BLOCK_W or tile_h % BLOCK_H:
        raise ValueError("tile size must be a multiple of %dx%d" % (BLOCK_W, BLOCK_H))
    tx, _ = tile_grid(width, height, tile_w, tile_h)
    ys = np.arange(height)[:, None] // tile_h
    xs = np.arange(width)[None, :] // tile_w
    return ((ys * tx + xs) % world).astype(np.int32)


def tiles_of_rank(width, height, rank, world, tile_w=64, tile_h=64):
    tx, ty = tile_grid(width, height, tile_w, tile_h)
    return list(range(rank, tx * ty, world))


# ---- range-update blob: { uint64 ndesc, uint64 payload_bytes, desc[ndesc], payload } ---------------------------
DESC_DTYPE = np.dtype([("dst_word", "<u8"), ("src_word", "<u4"), ("nwords", "<u4"), ("buftype", "<i4"), ("pad", "<i4")])


def pack_ranges(ranges):
    """ranges: list of (buftype, start_byte, data_bytes ndarray uint8 with len % 4 == 0) -> blob (uint8 array).
    Same layout octree_cuc_export_pending produces, so host code can also build blobs directly."""
    descs = np.zeros(len(ranges), dtype=DESC_DTYPE)
    payload = []
    used = 0
    for i, (buftype, start, data) in enumerate(ranges):
        data = np.ascontiguousarray(data).view(np.uint8).ravel()
        if start % 4 or data.size % 4:
            raise ValueError("ranges are 4-byte granular")
        descs[i] = (start // 4, used // 4, data.size // 4, buftype, 0)
        payload.append(data)
        used += data.size
    hdr = np.array([len(ranges), used], dtype="<u8")
    parts = [hdr.view(np.uint8), descs.view(np.uint8)] + payload
    return np.concatenate(parts) if parts else np.zeros(16, np.uint8)


def unpack_ranges(blob):
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    nd, nbytes = [int(v) for v in blob[:16].view("<u8")]
    descs = blob[16:16 + nd * DESC_DTYPE.itemsize].view(DESC_DTYPE)
    payload = blob[16 + nd * DESC_DTYPE.itemsize:]
    assert payload.size == nbytes
    out = []
    for d in descs:
        s, n = int(d["src_word"]) * 4, int(d["nwords"]) * 4
        out.append((int(d["buftype"]), int(d["dst_word"]) * 4, payload[s:s + n].copy()))
    return out


def broadcast_blob(blob, src=0, device=None):
    """Broadcast a variable-size blob from `src` with torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    size = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    buf = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)))
    dist.broadcast(buf, src)
    return buf.cpu().numpy()


class ShardedFrame:
    """Frame assembly for N ranks (one process per GPU, torchrun) around one OctreeGlc per rank; used by bench.py and
    the multi-GPU tests.  A C host that owns all the GPUs of a box in one process uses octree_cuc_set_gpus instead,
    which wires the same mechanism inside the connector.  This class only exchanges CUDA IPC handles:

      gather="p2p"         peer stores into rank 0's framebuffer, completed by the connector's device-side fence
                           (octree_cuc_set_fence): no collective in the frame path at all
      gather="p2p_nccl"    the same stores fenced by one 4-byte NCCL all-reduce per frame (round 1, kept for A/B)
      gather="nccl"        every rank renders into its own zeroed buffer, frame = NCCL reduce(SUM)
    """

    def __init__(self, rc, width, height, rank, world, device, gather="p2p", tile=64):
        import torch
        import torch.distributed as dist
        self.rc, self.rank, self.world, self.gather = rc, rank, world, gather
        self.width, self.height = width, height
        self.dist, self.torch = dist, torch
        self.peer_ptr = 0
        self.opened = []
        self.frame_t = None
        self.out_t = None
        self.fence = torch.zeros(1, device=device)
        self._blob_dev = None
        self._blob_applied = False
        rc.set_shard(rank, world, tile, tile)
        if world == 1:
            return
        rc.reserve_frame(width, height, 1)
        if rank == 0:
            rc.enable_replication_log(True)   # range uploads from here on are exported by broadcast_updates
        if gather in ("p2p", "p2p_nccl"):
            h = torch.from_numpy(rc.ipc_export_frame().copy()).to(device) if rank == 0 else torch.zeros(
                64, dtype=torch.uint8, device=device)
            dist.broadcast(h, 0)
            if rank != 0:
                self.peer_ptr = rc.ipc_open(h.cpu().numpy())
                self.opened.append(self.peer_ptr)
                rc.set_frame_target(self.peer_ptr, width)
            if gather == "p2p":
                own = rc.fence_device()
                mine = torch.from_numpy(rc.ipc_export_ptr(own).copy()).to(device)
                every = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
                dist.all_gather(every, mine)
                ptrs = []
                for k in range(world):
                    if k == rank:
                        ptrs.append(own)
                    else:
                        ptrs.append(rc.ipc_open(every[k].cpu().numpy()))
                        self.opened.append(ptrs[-1])
                dist.barrier()
                rc.set_fence(rank, world, ptrs)
                dist.barrier()            # nobody renders before every rank's words are zeroed and wired
        elif gather == "nccl":
            self.frame_t = torch.zeros((height, width), dtype=torch.int32, device=device)
            # rank 0 reduces into a second buffer: an in-place reduce would leave the other ranks' tiles in
            # its render target and add them again on the next frame
            self.out_t = torch.zeros_like(self.frame_t) if rank == 0 else None
            rc.set_frame_target(self.frame_t.data_ptr(), width, keepalive=self.frame_t)
        else:
            raise ValueError(gather)

    def assemble(self):
        """Queue the per-frame exchange on the current stream (after the rank's render)."""
        if self.world == 1 or self.gather == "p2p":
            return                            # the fence is part of octree_glc_update
        if self.gather == "p2p_nccl":
            self.dist.all_reduce(self.fence)  # every rank's peer stores for this frame are complete after this
        else:
            if self.rank == 0:
                self.out_t.copy_(self.frame_t)
                self.dist.reduce(self.out_t, 0, op=self.dist.ReduceOp.SUM)  # disjoint tiles, zeros elsewhere
            else:
                self.dist.reduce(self.frame_t, 0, op=self.dist.ReduceOp.SUM)

    def read_frame(self, out=None):
        """Rank 0: the assembled RGBA8 frame as uint8 [H,W,4] (synchronises rank 0's stream).  With gather="p2p" the
        other ranks cannot overwrite the frame before rank 0's next update; with "p2p_nccl" the caller must keep them
        from rendering the next frame until this returns (a barrier)."""
        if self.world == 1 or self.gather in ("p2p", "p2p_nccl"):
            return self.rc.read_frame(out)
        host = self.out_t.cpu().numpy().view(np.uint8).reshape(self.height, self.width, 4)
        if out is not None:
            out[...] = host
            return out
        return host

    def read_frame_async(self, out=None):
        """Call on every rank after assemble(): rank 0 queues the copy of the assembled frame into `out` (page-locked
        uint8 [H,W,4]) so that it overlaps the next frame; finish with rc.wait_reads() on rank 0.
        Rank 0 snapshots its framebuffer into a staging buffer on the render stream (a few us) and the host copy runs
        from there on the copy stream.  gather="p2p": the other ranks' next stores wait for rank 0's next kernel,
        which is queued behind the snapshot -- nothing else to do.  "p2p_nccl": a second all-reduce."""
        if self.world == 1:
            return self.rc.read_frame_async(out)
        if self.gather == "nccl":
            return self.read_frame(out) if self.rank == 0 else None
        if self.rank == 0:
            self.rc.read_frame_staged(out)
        if self.gather == "p2p_nccl":
            self.dist.all_reduce(self.fence)
        return out

    def broadcast_updates(self, device):
        """Rank 0's range uploads since the last call -> every rank (rank 0 has applied its own already).  The blob
        goes from the connector's page-locked log to rank 0's GPU (octree_cuc_export_pending_device), through one NCCL
        broadcast, and is applied from the receive buffer on the device (octree_cuc_apply_blob_device): no pageable
        copy, nothing through the other ranks' hosts."""
        if self.world == 1:
            return 0
        torch, dist = self.torch, self.dist
        if device is None or torch.device(device).type != "cuda":      # CPU (gloo) tests: the host path
            blob = self.rc.export_pending() if self.rank == 0 else None
            blob = broadcast_blob(blob, 0, device)
            if self.rank != 0:
                self.rc.apply_blob(blob)
            return len(blob)
        if self._blob_applied:
            self.rc.sync()
            self._blob_applied = False
        need = self.rc.export_pending_device(0, 0) if self.rank == 0 else 0
        size = torch.tensor([need], dtype=torch.int64, device=device)
        dist.broadcast(size, 0)
        need = int(size.item())
        if self._blob_dev is None or self._blob_dev.numel() < need:
            self._blob_dev = torch.empty(need + need // 4 + 4096, dtype=torch.uint8, device=device)
        buf = self._blob_dev[:need]
        if self.rank == 0:
            self.rc.export_pending_device(buf.data_ptr(), need)   # page-locked log -> send buffer, stream-ordered
        dist.broadcast(buf, 0)
        if self.rank != 0 and need > 16:
            # the connector may run on a stream of its own: the blob is complete before it reads it, and (the
            # buffer is reused) applied before the next broadcast can overwrite it -- rc.sync in the next call
            torch.cuda.current_stream().synchronize()
            self.rc.apply_blob_device(buf.data_ptr(), need)
            self._blob_applied = True
        return need

    def close(self):
        if self.world > 1 and self.gather == "p2p":
            self.rc.sync()
            self.dist.barrier()
            self.rc.set_fence(0, 1, None)
        if self.peer_ptr:
            self.rc.set_frame_target(0, 0)
        for p in self.opened:
            self.rc.ipc_close(p)
        self.opened = []
        self.peer_ptr = 0
