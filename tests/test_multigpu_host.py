"""Host logic of the N > 1 path on CPU: two processes over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qubatron_b200 import multigpu as M


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. tile sharding: every rank paints its own tiles of a zeroed frame; reduce(SUM) over int32 = union
        owner = M.tile_owner_map(W, H, world, tile, tile)
        rng = np.random.default_rng(5)
        full = rng.integers(1, 2 ** 31 - 1, size=(H, W), dtype=np.int32)  # "the frame", identical on all ranks
        mine = np.where(owner == rank, full, 0).astype(np.int32)
        t = torch.from_numpy(mine.copy())
        dist.reduce(t, 0, op=dist.ReduceOp.SUM)
        ok_union = bool(rank != 0 or np.array_equal(t.numpy(), full))
        # the rule the kernels use: tile index % world
        tx, ty = M.tile_grid(W, H, tile, tile)
        tiles = M.tiles_of_rank(W, H, rank, world, tile, tile)
        ok_tiles = all(owner[(i // tx) * tile, (i % tx) * tile] == rank for i in tiles)
        counts = torch.tensor([len(tiles)])
        dist.all_reduce(counts)
        ok_tiles = ok_tiles and int(counts.item()) == tx * ty

        # 2. range updates: rank 0 packs zero-and-append node ranges + a colour sub-range, everyone receives them
        blob = None
        ranges = None
        if rank == 0:
            r2 = np.random.default_rng(9)
            ranges = [(2, 48 * 17, r2.integers(0, 1000, 12, dtype=np.int32).view(np.uint8)),
                      (2, 48 * 90000, np.zeros(48, np.uint8)),
                      (0, 12 * 5, r2.random(9, dtype=np.float32).view(np.uint8)),
                      (5, 0, r2.integers(0, 50, 24, dtype=np.int32).view(np.uint8))]
            blob = M.pack_ranges(ranges)
        got = M.broadcast_blob(blob, 0)
        un = M.unpack_ranges(got)
        sig = [(b, s, bytes(d)) for b, s, d in un]
        # compare with rank 0's original through a checksum exchange
        import hashlib
        h = hashlib.sha256(repr(sig).encode()).digest()
        ht = torch.tensor(list(h), dtype=torch.int64)
        h0 = ht.clone()
        dist.broadcast(h0, 0)
        ok_blob = bool((ht == h0).all()) and len(un) == 4 and un[1][1] == 48 * 90000
        if rank == 0:
            ok_blob = ok_blob and all(a[0] == b[0] and a[1] == b[1] and bytes(a[2]) == bytes(b[2])
                                      for a, b in zip(ranges, un))
        q.put((rank, ok_union, ok_tiles, ok_blob))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H,tile", [(1920, 1080, 64), (200, 150, 32)])
def test_two_ranks_over_gloo(W, H, tile):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, W, H, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_union, ok_tiles, ok_blob in res:
        assert ok_union and ok_tiles and ok_blob, (rank, ok_union, ok_tiles, ok_blob)


def test_owner_map_balances_tiles():
    for world in (1, 2, 4, 8):
        owner = M.tile_owner_map(3840, 2160, world)
        counts = np.bincount(owner.ravel(), minlength=world)
        assert counts.min() > 0 and counts.max() / counts.min() < 1.05
    with pytest.raises(ValueError):
        M.tile_owner_map(64, 64, 2, 20, 8)


def test_blob_round_trip_and_layout():
    """The blob layout is the C struct RangeDesc {u64 dst_word; u32 src_word; u32 nwords; i32 buftype; i32 pad}."""
    assert M.DESC_DTYPE.itemsize == 24
    rng = np.random.default_rng(1)
    ranges = [(2, 48 * i, rng.integers(0, 99, 12, dtype=np.int32).view(np.uint8)) for i in (3, 5, 1000)]
    blob = M.pack_ranges(ranges)
    assert blob.size == 16 + 3 * 24 + 3 * 48
    un = M.unpack_ranges(blob)
    for a, b in zip(ranges, un):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
    assert M.unpack_ranges(M.pack_ranges([])) == []
