"""Host logic of the UNet executor on CPU: every C-ABI entry point it calls is replaced by a torch stand-in
(tests/fake_unet_ops.py), so the orchestration -- skip concat consumed in place, FiLM slices, epilogue statistics,
residual wiring, stride-2 / upsampling convolutions, the three attention dispatches, zero-padded heads, the
query-sharded attention under gloo -- is compared with the oracle without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fake_unet_ops
from conftest import rel_err
from oracle import unet_oracle as uo

BASE = dict(model_channels=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), num_heads=2)


def _net(in_ch, R, sd, monkeypatch=None, **kw):
    from holo_diffusion_b200 import ops
    from holo_diffusion_b200.unet import SimpleUnet3D
    fake_unet_ops.install(ops, monkeypatch.setattr if monkeypatch is not None else setattr)
    net = SimpleUnet3D(image_size=R, in_channels=in_ch, out_channels=in_ch, **kw)
    net._net.load_state_dict(sd, strict=True)
    net._exec.pair_dtype = torch.float32   # exact operand "pairs": hi = value, lo = 0
    return net


def _forward(net, x, t):
    _, C, D, H, W = x.shape
    y = net._exec.forward_cl(x[0].reshape(C, -1).t().contiguous(), (D, H, W), t)
    return y.t().reshape(1, -1, D, H, W)


@pytest.mark.parametrize("R,tc", [(16, True), (16, False), (32, True)])
def test_executor_orchestration_matches_oracle(R, tc, monkeypatch):
    """Base-args UNet (5 levels): every level >= 4^3 takes the tensor-core dispatch (conv / GN-epilogue statistics /
    raw pairs for the skip convolutions), below that and with tc=False the CUDA-core dispatch; at 32^3 the 8^3 level
    (T = 512) is the fused-attention dispatch."""
    sd = uo.make_unet_state_dict(16, 16, seed=2)
    net = _net(16, R, sd, monkeypatch, use_tensor_cores=tc, **BASE)
    x = torch.tanh(torch.randn(1, 16, R, R, R, generator=torch.Generator().manual_seed(0)))
    for t in (0, 500):
        tt = torch.full((1,), t, dtype=torch.long)
        out = _forward(net, x, tt)
        assert rel_err(out, uo.unet_forward(sd, x, tt)) < 2e-5
    assert (net._exec.tc_calls > 0) == tc


def test_executor_fused_skip_tail(monkeypatch):
    """HOLO_FUSE_SKIP=1: ResBlocks with a 1x1 skip connection run conv3^3 + skip conv + add as one launch
    (holo_conv3d_tc_skip); weight concatenation [taps | skip columns], summed bias and the statistics hand-over are
    host logic -- checked here against the oracle with the torch stand-in of the kernel."""
    monkeypatch.setenv("HOLO_FUSE_SKIP", "1")
    sd = uo.make_unet_state_dict(16, 16, seed=2)
    net = _net(16, 16, sd, monkeypatch, **BASE)
    from holo_diffusion_b200 import ops
    n = []
    orig = ops.conv3d_tc_skip
    monkeypatch.setattr(ops, "conv3d_tc_skip", lambda *a, **k: (n.append(a[5]), orig(*a, **k))[1])
    x = torch.tanh(torch.randn(1, 16, 16, 16, 16, generator=torch.Generator().manual_seed(0)))
    tt = torch.zeros(1, dtype=torch.long)
    assert rel_err(_forward(net, x, tt), uo.unet_forward(sd, x, tt)) < 2e-5
    # fused at the levels that run on tensor cores (>= 4^3): the channel-changing input blocks and the output blocks
    assert len(n) >= 8 and set(n) <= {64, 128, 192, 256, 384, 512, 768, 1024}, n


@pytest.mark.parametrize("flash", ["1", "0"])
def test_executor_attention_dispatches(flash, monkeypatch):
    """16^3 x (64, 128)-channel model with attention at both levels: T = 4096 (ch 32: zero-padded heads on the fused
    dispatch, or the CUDA-core kernel with HOLO_ATTN_FLASH=0) and T = 512 (ch 64: fused, or the three-launch pipeline)."""
    monkeypatch.setenv("HOLO_ATTN_FLASH", flash)
    kw = dict(model_channels=64, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), num_heads=2)
    sd = uo.make_unet_state_dict(16, 16, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), seed=3)
    net = _net(16, 16, sd, monkeypatch, **kw)
    from holo_diffusion_b200 import ops
    seen = []
    for name in ("attention_flash", "attention_simt", "gemm_tc"):
        orig = getattr(ops, name)
        monkeypatch.setattr(ops, name, (lambda o, n: lambda *a, **k: (seen.append(n), o(*a, **k))[1])(orig, name))
    x = torch.tanh(torch.randn(1, 16, 16, 16, 16, generator=torch.Generator().manual_seed(1)))
    tt = torch.zeros(1, dtype=torch.long)
    out = _forward(net, x, tt)
    assert rel_err(out, uo.unet_forward(sd, x, tt)) < 2e-5
    if flash == "1":
        assert set(seen) == {"attention_flash"}
    else:
        assert "attention_flash" not in seen and "attention_simt" in seen and "gemm_tc" in seen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    kw = dict(model_channels=64, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), num_heads=2)
    sd = uo.make_unet_state_dict(16, 16, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), seed=3)
    net = _net(16, 8, sd, None, **kw)
    x = torch.tanh(torch.randn(1, 16, 8, 8, 8, generator=torch.Generator().manual_seed(1)))
    tt = torch.zeros(1, dtype=torch.long)
    single = _forward(net, x, tt)
    net.shard_attention(min_tokens=512)      # T = 512 at 8^3: 4 query tiles over 2 ranks; T = 64 at 4^3 stays local
    calls = []
    from holo_diffusion_b200 import ops
    orig = ops.attention_flash
    ops.attention_flash = lambda *a, **k: (calls.append((a[4], a[11] if len(a) > 11 else k.get("q_begin", 0),
                                                         a[12] if len(a) > 12 else k.get("q_count", 0))), orig(*a, **k))[1]
    sharded = _forward(net, x, tt)
    q.put((rank, rel_err(sharded, single), rel_err(sharded, uo.unet_forward(sd, x, tt)), calls))
    dist.destroy_process_group()


def test_query_sharded_unet_forward_world2():
    """The whole denoiser under gloo on 2 ranks: the T = 512 attention block computes half of the query tiles per rank
    and all-gathers the projection operand; both ranks reproduce the un-sharded result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in ps]
    for rank, d_single, d_oracle, calls in res:
        assert d_single < 1e-6 and d_oracle < 2e-5
        assert (512, rank * 256, 256) in calls, calls          # this rank's half of the 512 queries
