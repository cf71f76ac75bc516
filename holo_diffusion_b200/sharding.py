"""Multi-GPU work split: independent units (grids / views) are dealt round-robin to ranks; the only collective
is one gather of the finished images to rank 0.  The reference has no multi-GPU inference
path (generate_samples.py:34 uses a single device); one grid per process matches its training layout
(holo_diffusion_model.py:326).  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Unit i runs on rank i % world."""
    return list(range(rank, n_units, world))


def query_shard(T: int, world: int, rank: int, tile: int = 128):
    """Query-sharded attention (cfg #5): the T // tile query tiles are dealt in contiguous blocks of
    ceil(tiles / world).  Returns (q_begin, q_count, chunk): this rank computes queries [q_begin, q_begin + q_count)
    (q_count may be 0 for trailing ranks) and contributes rows [rank * chunk, (rank + 1) * chunk) to the all-gather;
    chunk * world >= T."""
    assert T % tile == 0
    tiles = T // tile
    per = (tiles + world - 1) // world
    t0 = min(tiles, rank * per)
    t1 = min(tiles, (rank + 1) * per)
    return rank * per * tile, (t1 - t0) * tile, per * tile


def row_shard(H: int, world: int, rank: int):
    """Image rows of ONE view dealt in contiguous blocks of ceil(H / world): returns (h0, h1, per)."""
    per = (H + world - 1) // world
    return min(H, rank * per), min(H, (rank + 1) * per), per


def gather_rows(local: torch.Tensor, H: int, rank: int, world: int, group=None) -> torch.Tensor:
    """local (B, h, W, C) = this rank's row block of an image-shaped render output -> (B, H, W, C) on EVERY rank
    (one all-gather of equal-sized, zero-padded blocks)."""
    if world == 1:
        return local
    B, h, W, C = local.shape
    per = (H + world - 1) // world
    pad = local.new_zeros(per, B, W, C)
    pad[:h] = local.permute(1, 0, 2, 3)
    buf = local.new_empty(world * per, B, W, C)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return buf[:H].permute(1, 0, 2, 3).contiguous()


def gather_images(local: torch.Tensor, n_units: int, rank: int, world: int) -> torch.Tensor:
    """local (n_local, ...) images of this rank's units (in shard_units order) -> on rank 0: (n_units, ...) in unit
    order; other ranks get an empty tensor.  One all_gather of equal-sized (padded) blocks."""
    if world == 1:
        return local
    per = (n_units + world - 1) // world
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty(world * per, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad)
    if rank != 0:
        return local.new_empty(0, *local.shape[1:])
    buf = buf.view(world, per, *local.shape[1:])
    out = torch.empty(n_units, *local.shape[1:], dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_units(n_units, r, world)
        if idx:
            out[idx] = buf[r, : len(idx)]
    return out
