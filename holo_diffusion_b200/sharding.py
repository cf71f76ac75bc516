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
