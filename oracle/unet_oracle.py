"""CPU restatement of the guided-diffusion 3-D UNet (TEST INFRASTRUCTURE ONLY).

A pure function of a reference-named state dict, so that the same weights drive the reference
``UNetModel`` (imported from /root/reference by ``tests/golden/make_golden.py`` to pin this file),
this oracle and the CUDA path.  Follows ``holo_diffusion/guided_diffusion/unet.py``:
``UNetModel.forward`` :800-837, ``ResBlock._forward`` :236-256, ``AttentionBlock._forward`` :397-406,
``QKVAttentionLegacy.forward`` :438-455, ``Upsample`` :91-106, ``Downsample`` :136-138 and
``nn.py`` ``timestep_embedding`` :109-127, ``GroupNorm32`` :23-25,99-106.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

GN_GROUPS = 32
GN_EPS = 1e-5


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)  # built on the CPU, then moved (nn.py:119-122)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1)


def _f32(x):
    """``x.float()`` as the reference writes it (nn.py:25, unet.py:452); an fp64 tensor stays fp64 so that the same
    code evaluates the noise-free twin from a ``.double()`` state dict."""
    return x if x.dtype == torch.float64 else x.float()


def _gn(sd, pre, x):
    return F.group_norm(_f32(x), GN_GROUPS, sd[pre + ".weight"], sd[pre + ".bias"], GN_EPS)


def _conv(sd, pre, x, stride=1, padding=1):
    return F.conv3d(x, sd[pre + ".weight"], sd[pre + ".bias"], stride=stride, padding=padding)


def _resblock(sd, pre, x, emb):
    h = _conv(sd, pre + ".in_layers.2", F.silu(_gn(sd, pre + ".in_layers.0", x)))
    e = F.linear(F.silu(emb), sd[pre + ".emb_layers.1.weight"], sd[pre + ".emb_layers.1.bias"])
    scale, shift = e[:, :, None, None, None].chunk(2, 1)
    h = _gn(sd, pre + ".out_layers.0", h) * (1 + scale) + shift
    h = _conv(sd, pre + ".out_layers.3", F.silu(h))
    if pre + ".skip_connection.weight" in sd:
        x = _conv(sd, pre + ".skip_connection", x, padding=0)
    return x + h


def _attention(sd, pre, x, n_heads):
    b, c = x.shape[:2]
    sp = x.shape[2:]
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, pre + ".norm", xf), sd[pre + ".qkv.weight"], sd[pre + ".qkv.bias"])
    T = qkv.shape[-1]
    ch = c // n_heads
    q, k, v = qkv.reshape(b * n_heads, 3 * ch, T).split(ch, 1)  # head-major [q_h | k_h | v_h]
    s = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(_f32(torch.einsum("bct,bcs->bts", q * s, k * s)), -1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, T)
    h = F.conv1d(a, sd[pre + ".proj_out.weight"], sd[pre + ".proj_out.bias"])
    return (xf + h).reshape(b, c, *sp)


def _run_block(sd, pre, h, emb, n_heads):
    j = 0
    while True:
        p = f"{pre}.{j}"
        if p + ".in_layers.0.weight" in sd:
            h = _resblock(sd, p, h, emb)
        elif p + ".qkv.weight" in sd:
            h = _attention(sd, p, h, n_heads)
        elif p + ".op.weight" in sd:
            h = _conv(sd, p + ".op", h, stride=2)
        elif p + ".conv.weight" in sd:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv(sd, p + ".conv", h)
        elif p + ".weight" in sd:
            h = _conv(sd, p, h)
        else:
            return h
        j += 1


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, n_heads: int = 2,
                 prefix: str = "") -> torch.Tensor:
    """x (N,C,D,H,W) fp32, t (N,) int64 -> (N,Cout,D,H,W)."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    mc = sd["time_embed.0.weight"].shape[1]
    dt = sd["time_embed.0.weight"].dtype  # fp32 = the reference's arithmetic; fp64 = noise-free twin
    emb = F.linear(timestep_embedding(t, mc).to(dt), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    n_in = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("input_blocks."))
    n_out = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("output_blocks."))
    hs = []
    h = x.to(dt)
    for i in range(n_in):
        h = _run_block(sd, f"input_blocks.{i}", h, emb, n_heads)
        hs.append(h)
    h = _run_block(sd, "middle_block", h, emb, n_heads)
    for i in range(n_out):
        h = torch.cat([h, hs.pop()], 1)
        h = _run_block(sd, f"output_blocks.{i}", h, emb, n_heads)
    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h)))


# ----------------------------------------------------------------------------------------
# state-dict fixture with the reference's parameter names and SimpleUnet3D's initialisation
# (holo_diffusion/utils/diffusion_utils.py:56-80), plus the randomisation SURVEY.md section 4 asks for
# ----------------------------------------------------------------------------------------
def make_unet_state_dict(in_ch: int, out_ch: int, model_ch: int = 64, num_res_blocks: int = 2,
                         channel_mult: Sequence[int] = (1, 1, 2, 4, 8), attention_resolutions: Sequence[int] = (4, 8),
                         seed: int = 2, randomize: bool = True) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def xavier(*shape):
        rf = 1
        for s in shape[2:]:
            rf *= s
        bound = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
        return ((torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).float()

    def rnd(n, std, mean=0.0):
        return (torch.randn(n, generator=g, dtype=torch.float64) * std + mean).float()

    def conv3(pre, ci, co, k=3):
        sd[pre + ".weight"] = xavier(co, ci, k, k, k)
        sd[pre + ".bias"] = rnd(co, 0.05) if randomize else torch.zeros(co)

    def lin(pre, ci, co):
        sd[pre + ".weight"] = xavier(co, ci)
        sd[pre + ".bias"] = rnd(co, 0.05) if randomize else torch.zeros(co)

    def gn(pre, c):
        sd[pre + ".weight"] = rnd(c, 0.2, 1.0) if randomize else torch.ones(c)
        sd[pre + ".bias"] = rnd(c, 0.2) if randomize else torch.zeros(c)

    def conv1(pre, ci, co, zero):
        if zero and not randomize:
            sd[pre + ".weight"] = torch.zeros(co, ci, 1)
            sd[pre + ".bias"] = torch.zeros(co)
        else:
            sd[pre + ".weight"] = xavier(co, ci, 1)
            sd[pre + ".bias"] = rnd(co, 0.1)

    def res(pre, ci, co, te):
        gn(pre + ".in_layers.0", ci)
        conv3(pre + ".in_layers.2", ci, co)
        lin(pre + ".emb_layers.1", te, 2 * co)
        gn(pre + ".out_layers.0", co)
        conv3(pre + ".out_layers.3", co, co)
        if ci != co:
            conv3(pre + ".skip_connection", ci, co, k=1)

    def attn(pre, c):
        gn(pre + ".norm", c)
        conv1(pre + ".qkv", c, 3 * c, False)
        conv1(pre + ".proj_out", c, c, True)

    te = 4 * model_ch
    lin("time_embed.0", model_ch, te)
    lin("time_embed.2", te, te)
    ch = int(channel_mult[0] * model_ch)
    conv3("input_blocks.0.0", in_ch, ch)
    chans = [ch]
    ds, bi = 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            res(f"input_blocks.{bi}.0", ch, int(mult * model_ch), te)
            ch = int(mult * model_ch)
            if ds in attention_resolutions:
                attn(f"input_blocks.{bi}.1", ch)
            chans.append(ch)
            bi += 1
        if level != len(channel_mult) - 1:
            conv3(f"input_blocks.{bi}.0.op", ch, ch)
            chans.append(ch)
            bi += 1
            ds *= 2
    res("middle_block.0", ch, ch, te)
    attn("middle_block.1", ch)
    res("middle_block.2", ch, ch, te)
    bo = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            res(f"output_blocks.{bo}.0", ch + ich, int(model_ch * mult), te)
            ch = int(model_ch * mult)
            j = 1
            if ds in attention_resolutions:
                attn(f"output_blocks.{bo}.{j}", ch)
                j += 1
            if level and i == num_res_blocks:
                conv3(f"output_blocks.{bo}.{j}.conv", ch, ch)
                ds //= 2
            bo += 1
    gn("out.0", ch)
    conv3("out.2", int(channel_mult[0] * model_ch), out_ch)
    return sd
