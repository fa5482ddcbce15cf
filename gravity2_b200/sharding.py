"""Multi-GPU sharding of the path (one process per GPU, torch.distributed over NCCL / NVLink).

The tables are small (cfg3: 1.2 GB as 32-bit fixed point) and are replicated on every GPU; the WORK is sharded:
  * the base (un-resampled) matrix by contiguous ranges of the upper-triangular 128x128 pair-tile list
    (every tile costs the same; diagonal tiles are computed in full), gathered with one all-gather of
    fixed-size packed tile blocks;
  * bootstrap replicates by index (rank r takes replicates r, r + world, ...): no exchange at all, each rank
    streams its own replicate matrices to its consumer.
There is no other collective on the data path.
"""
from __future__ import annotations

from typing import List, Tuple

TILE = 128


def num_tiles(n: int) -> int:
    nt = (n + TILE - 1) // TILE
    return nt * (nt + 1) // 2


def tile_coords(t: int, nt: int) -> Tuple[int, int]:
    """linear upper-triangular tile index -> (bi <= bj); the host mirror of gv2_tile_coords (gj_kernels.cu)"""
    bi = 0
    while (bi + 1) * nt - (bi + 1) * bi // 2 <= t:
        bi += 1
    return bi, bi + t - (bi * nt - bi * (bi - 1) // 2)


def tiles_per_rank(ntiles: int, world: int) -> int:
    return (ntiles + world - 1) // world


def tile_range(ntiles: int, rank: int, world: int) -> Tuple[int, int]:
    per = tiles_per_rank(ntiles, world)
    return min(ntiles, rank * per), min(ntiles, (rank + 1) * per)


def replicate_indices(n_boot: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_boot, world))


def all_gather_tiles(packed, world: int):
    """packed: [tiles_per_rank, TILE*TILE] float64 tensor of this rank's tile blocks (zero padded to the
    common length).  Returns [world * tiles_per_rank, TILE*TILE]; rank r's blocks start at r * tiles_per_rank."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out
