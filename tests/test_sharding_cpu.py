"""World-size-2 gloo test of the multi-GPU work split + gather logic (host side; no CUDA)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from holo_diffusion_b200.sharding import gather_images, shard_units
    mine = shard_units(n_units, rank, world)
    imgs = torch.stack([torch.full((5, 4, 4), float(u)) for u in mine]) if mine else torch.empty(0, 5, 4, 4)
    out = gather_images(imgs, n_units, rank, world)
    if rank == 0:
        q.put(out[:, 0, 0, 0].tolist())
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_units = 5  # ragged: 3 + 2
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    [p.start() for p in ps]
    res = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert res == [0.0, 1.0, 2.0, 3.0, 4.0]  # rank 0 holds every unit, in unit order


def test_shard_units_partition():
    from holo_diffusion_b200.sharding import shard_units
    for n in (0, 1, 7, 32):
        for w in (1, 2, 8):
            got = sorted(u for r in range(w) for u in shard_units(n, r, w))
            assert got == list(range(n))
