"""Query-sharded attention across 2 GPUs (NCCL): one sample, every rank evaluates the UNet on the same input and the
attention blocks split their query tiles; the result must agree with the un-sharded evaluation to summation-order
noise (the attention rows themselves are bit-identical, but the small-grid convolutions split K with fp32 atomics, so
two evaluations differ at the 1e-7 level anyway), be identical on both ranks' inputs to the projection, and match the
oracle.  Needs 2 GPUs: skipped on the 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import rel_err
    from holo_diffusion_b200.unet import SimpleUnet3D
    from oracle import unet_oracle as uo
    kw = dict(model_channels=64, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), num_heads=2)
    sd = uo.make_unet_state_dict(16, 16, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), seed=2)
    net = SimpleUnet3D(image_size=16, in_channels=16, out_channels=16, **kw)
    net._net.load_state_dict(sd, strict=True)
    net.cuda()
    x = torch.tanh(torch.randn(1, 16, 16, 16, 16, generator=torch.Generator().manual_seed(0)))
    tt = torch.zeros(1, dtype=torch.long)
    single = net(x.cuda(), tt.cuda())
    net.shard_attention(min_tokens=512)          # T = 4096 (16^3) and 512 (8^3) are both sharded
    sharded = net(x.cuda(), tt.cuda())
    torch.cuda.synchronize()
    ref = uo.unet_forward(sd, x, tt)
    q.put((rank, rel_err(sharded, single), rel_err(sharded, ref)))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_query_sharded_attention_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=300) for _ in range(2)]
    [p.join(60) for p in ps]
    for rank, d_single, err in res:
        print(f"rank {rank}: sharded vs un-sharded {d_single:.2e}, vs oracle {err:.2e}")
        # the single-GPU attention shares the keys of a query tile between CTAs (split-KV, the default) and merges the
        # partial softmaxes; the query-sharded one keeps one CTA per tile: same arithmetic, another rounding order
        assert d_single < 2e-5, f"rank {rank}: sharded attention differs from the single-GPU result ({d_single})"
        assert err < 2e-5, (rank, err)
