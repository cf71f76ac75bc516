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


def test_query_shard_partition():
    """Query tiles of a sharded attention block: every tile on exactly one rank, equal all-gather chunks."""
    from holo_diffusion_b200.sharding import query_shard
    for T in (128, 512, 4096, 128 * 13):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                q0, qn, chunk = query_shard(T, w, r)
                assert q0 == r * chunk and q0 % 128 == 0 and qn % 128 == 0 and 0 <= qn <= chunk
                seen += list(range(q0, q0 + qn, 128))
            assert seen == list(range(0, T, 128)) and chunk * w >= T


def _attn_worker(rank, world, port, T, C, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from holo_diffusion_b200 import ops
    from holo_diffusion_b200.unet import UNetExecutor

    def fake_flash(q_hi, q_lo, vt_hi, vt_lo, T_, heads, ch, out, out_hi, out_lo, scale=0.0, q_begin=0, q_count=0):
        rows = torch.arange(q_begin, q_begin + q_count, dtype=torch.float32)[:, None]   # stand-in for the kernel:
        out_hi[q_begin:q_begin + q_count] = (rows + 0.25).to(out_hi.dtype)                # row index as the payload
        out_lo[q_begin:q_begin + q_count] = (-rows).to(out_lo.dtype)
        return 0

    ops.attention_flash = fake_flash
    ex = UNetExecutor.__new__(UNetExecutor)   # host logic only: no parameters, no CUDA
    ex.attn_group, ex.attn_shard_min_tokens, ex.pair_dtype = dist.group.WORLD, 256, torch.float32
    assert ex._attn_world(128) == (1, 0) and ex._attn_world(T) == (world, rank)
    dummy = torch.empty(1)
    a_hi, a_lo = ex._attn_sharded(dummy, dummy, dummy, dummy, T, 1, C, 0.125, world, rank)
    q.put((rank, a_hi[:, 0].tolist(), a_lo[:, 0].tolist(), tuple(a_hi.shape)))
    dist.destroy_process_group()


def test_query_sharded_attention_gather_world2():
    """Host logic of the query-sharded attention (kernel replaced by a stand-in): after the all-gather every rank
    holds all T rows of the operand pair, in query order; T = 5 tiles is ragged over 2 ranks (3 + 2)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    T, C = 5 * 128, 8
    ps = [ctx.Process(target=_attn_worker, args=(r, 2, port, T, C, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in ps]
    for rank, hi, lo, shape in res:
        assert shape == (T, C)
        assert hi == [t + 0.25 for t in range(T)] and lo == [-float(t) for t in range(T)]


def _rows_worker(rank, world, port, H, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from holo_diffusion_b200.sharding import gather_rows, row_shard
    h0, h1, per = row_shard(H, world, rank)
    full = torch.arange(H * 3 * 2, dtype=torch.float32).reshape(1, H, 3, 2)
    out = gather_rows(full[:, h0:h1].contiguous(), H, rank, world)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


def test_row_sharded_render_gather_world2():
    """One view split by image rows over 2 ranks (cfg #5): every rank ends with the full image, rows in order
    (H = 7 is ragged: 4 + 3)."""
    from holo_diffusion_b200.sharding import row_shard
    for H in (7, 8, 512):
        for w in (1, 2, 8):
            rows = [r for k in range(w) for r in range(*row_shard(H, w, k)[:2])]
            assert rows == list(range(H))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_rows_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res)
