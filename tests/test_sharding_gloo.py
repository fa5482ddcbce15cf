"""world_size-2 gloo test (CPU) of the multi-GPU host logic: tile ranges partition the triangle, replicate
striding covers every replicate once, and the packed-tile all-gather + unpack reassembles a symmetric matrix.
The pack/unpack used here are numpy stand-ins for gj_epilogue_kernel(packed) / unpack_tiles_kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gravity2_b200 import sharding

TILE = sharding.TILE


def _pack(M, t0, t1, per):
    n = M.shape[0]
    nt = (n + TILE - 1) // TILE
    Mp = np.zeros((nt * TILE, nt * TILE))
    Mp[:n, :n] = M
    out = np.zeros((per, TILE * TILE))
    for k, t in enumerate(range(t0, t1)):
        bi, bj = sharding.tile_coords(t, nt)
        out[k] = Mp[bi * TILE:(bi + 1) * TILE, bj * TILE:(bj + 1) * TILE].reshape(-1)
    return out


def _unpack(gathered, n, world):
    nt = (n + TILE - 1) // TILE
    ntiles = sharding.num_tiles(n)
    per = sharding.tiles_per_rank(ntiles, world)
    Mp = np.zeros((nt * TILE, nt * TILE))
    for r in range(world):
        a, b = sharding.tile_range(ntiles, r, world)
        for k, t in enumerate(range(a, b)):
            bi, bj = sharding.tile_coords(t, nt)
            blk = gathered[r * per + k].reshape(TILE, TILE)
            Mp[bi * TILE:(bi + 1) * TILE, bj * TILE:(bj + 1) * TILE] = blk
            Mp[bj * TILE:(bj + 1) * TILE, bi * TILE:(bi + 1) * TILE] = blk.T
    return Mp[:n, :n]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    M = rng.random((n, n))
    M = (M + M.T) / 2
    ntiles = sharding.num_tiles(n)
    per = sharding.tiles_per_rank(ntiles, world)
    t0, t1 = sharding.tile_range(ntiles, rank, world)
    packed = torch.from_numpy(_pack(M, t0, t1, per))
    gathered = sharding.all_gather_tiles(packed, world).numpy()
    ok = np.array_equal(_unpack(gathered, n, world), M)
    reps = sharding.replicate_indices(11, rank, world)
    allreps = [None] * world
    dist.all_gather_object(allreps, reps)
    ok = ok and sorted(sum(allreps, [])) == list(range(11))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [300, 129])
def test_tile_sharded_all_gather_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_tile_ranges_partition_and_coords():
    for n, world in [(1, 1), (1000, 3), (10000, 8), (129, 2)]:
        ntiles = sharding.num_tiles(n)
        nt = (n + TILE - 1) // TILE
        seen = []
        for r in range(world):
            a, b = sharding.tile_range(ntiles, r, world)
            seen += list(range(a, b))
        assert seen == list(range(ntiles))
        coords = [sharding.tile_coords(t, nt) for t in range(min(ntiles, 200))]
        assert coords == [(i, j) for i in range(nt) for j in range(i, nt)][:len(coords)]
