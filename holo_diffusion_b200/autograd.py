"""Differentiable renderer stages (renderer half): torch.autograd Functions whose forward
and backward are the CUDA kernels (holo_if_fwd / holo_if_bwd, holo_ea_raymarch / holo_ea_raymarch_bwd), so that
``objective.backward()`` (/root/reference/trainer/training_loop.py:518-556) reaches the voxel grid and the RenderMLP
parameters through the ray marcher (holo_multipass_ea.py:96-100, density noise :87-91) and the implicit function
(holo_voxel_grid_implicit_function.py:182-269).  The refiner stays under no_grad, as in the reference.

The density net is evaluated in its collapsed form ``LeakyReLU(W_eff x + b_eff)``; ``compose_density_net`` is the
differentiable (fp64, torch) twin of ``holo_affine_compose_f64``: autograd carries dL/d(W_eff, b_eff) back to the four
layers -- O(parameters) work per step, independent of the number of points.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ops
from ._lib import lib


def compose_density_net(layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], skips: Sequence[int]):
    """(W_i, b_i) of MLPWithInputSkips' Linear layers (activation only after the last one, custom_modules.py:108-112;
    a skip layer consumes cat(hidden, input), :157) -> A (out, C), c (out) in fp64 with y = A x + c before that
    activation.  Differentiable."""
    A = c = None
    for li, (W, b) in enumerate(layers):
        W64, b64 = W.double(), b.double()
        if A is None:
            A, c = W64, b64
            continue
        rows = A.shape[0]
        if li in skips:
            A, c = W64[:, :rows] @ A + W64[:, rows:], W64[:, :rows] @ c + b64
        else:
            A, c = W64 @ A, W64 @ c + b64
    return A, c


class ImplicitFunctionRays(torch.autograd.Function):
    """(grid (1,C,D,H,W), A (H+1,C) f64, c (H+1) f64, Wr (3,H+E), br (3); origins, dirs, lengths) ->
    densities (n,S), rgb (n,S,3)."""

    @staticmethod
    def forward(ctx, grid, A, c, Wr, br, origins, dirs, lengths, extent: float, n_harm: int):
        _, C, D, H, W = grid.shape
        V = D * H * W
        dev = grid.device
        grid_cl = ops.transpose2d(grid.detach().contiguous().float().reshape(-1), C, V).view(D, H, W, C)
        Hd = A.shape[0] - 1
        E = Wr.shape[1] - Hd
        packed = torch.empty(int(lib().cdll.holo_render_mlp_packed_floats(Hd, C, E)), device=dev)
        lib().call("holo_pack_render_mlp", ops._ptr(A.detach().contiguous(), torch.float64),
                   ops._ptr(c.detach().contiguous(), torch.float64), ops._ptr(Wr.detach().contiguous()),
                   ops._ptr(br.detach().contiguous()), Hd, C, E, ops._ptr(packed), ops._stream())
        n, S = lengths.shape
        o, d, l = origins.contiguous().float(), dirs.contiguous().float(), lengths.contiguous().float()
        dens, feats, _ = ops.if_fwd(grid_cl, extent, packed, Hd, n_harm, origins=o, dirs=d, lengths=l)
        ctx.save_for_backward(grid_cl, packed, o, d, l)
        ctx.meta = (extent, n_harm, Hd, E, C, (D, H, W))
        return dens.view(n, S), feats.view(n, S, 3)

    @staticmethod
    def backward(ctx, g_dens, g_rgb):
        grid_cl, packed, o, d, l = ctx.saved_tensors
        extent, n_harm, Hd, E, C, (D, H, W) = ctx.meta
        dev = grid_cl.device
        n, S = l.shape
        d_grid = torch.zeros(D, H, W, C, device=dev)
        d_W = torch.zeros(Hd + 1, C, device=dev)
        d_b = torch.zeros(Hd + 1, device=dev)
        d_Wr = torch.zeros(3, Hd + E, device=dev)
        d_br = torch.zeros(3, device=dev)
        lib().call("holo_if_bwd", ops._ptr(grid_cl), D, H, W, C, float(extent), ops._ptr(packed), Hd, n_harm, ops._ptr(o),
                   ops._ptr(d), ops._ptr(l), n * S, S, ops._ptr(g_dens.contiguous().float()),
                   ops._ptr(g_rgb.contiguous().float()), ops._ptr(d_grid), ops._ptr(d_W), ops._ptr(d_b), ops._ptr(d_Wr),
                   ops._ptr(d_br), ops._stream())
        V = D * H * W
        g_grid = ops.transpose2d(d_grid.reshape(-1), V, C).view(1, C, D, H, W)
        return g_grid, d_W.double(), d_b.double(), d_Wr, d_br, None, None, None, None, None


class EARaymarch(torch.autograd.Function):
    """(densities (n,S), features (n,S,F), lengths (n,S), noise (n,S) | None) -> features (n,F), depths (n,1), masks (n,1),
    weights (n,S) (pytorch3d EmissionAbsorptionRaymarcher with the shipped settings, configs/base.yaml:149-159)."""

    @staticmethod
    def forward(ctx, dens, feats, lengths, noise, bg, bg_opacity: float):
        dens, feats, lengths = dens.contiguous().float(), feats.contiguous().float(), lengths.contiguous().float()
        o = ops.ea_raymarch(dens, feats, lengths, bg, bg_opacity, noise=noise)
        Fd = feats.shape[-1]
        bg_dev = torch.tensor([float(x) for x in (bg if len(bg) == Fd else tuple(bg) * Fd)], device=dens.device)
        ctx.save_for_backward(dens, feats, lengths, noise if noise is not None else torch.empty(0, device=dens.device), bg_dev)
        ctx.meta = (float(bg_opacity), noise is not None)
        return o["features"], o["depths"], o["masks"], o["weights"]

    @staticmethod
    def backward(ctx, g_feat, g_depth, g_mask, g_w):
        dens, feats, lengths, noise, bg_dev = ctx.saved_tensors
        bg_opacity, has_noise = ctx.meta
        n, S = dens.shape
        Fd = feats.shape[-1]
        d_dens = torch.empty(n, S, device=dens.device)
        d_feats = torch.empty(n, S, Fd, device=dens.device)

        def p(t):
            return None if t is None else ops._ptr(t.contiguous().float())

        lib().call("holo_ea_raymarch_bwd", ops._ptr(dens), ops._ptr(feats), ops._ptr(lengths), ops._ptr(noise) if has_noise else None,
                   n, S, Fd, ops._ptr(bg_dev), bg_opacity, p(g_feat if g_feat is not None else torch.zeros(n, Fd, device=dens.device)),
                   p(g_depth), p(g_mask), p(g_w), ops._ptr(d_dens), ops._ptr(d_feats), ops._stream())
        return d_dens, d_feats, None, None, None, None
