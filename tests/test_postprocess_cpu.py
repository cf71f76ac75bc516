"""Anchors for oracle/postprocess_oracle.py (the pytorch3d leaves it restates are absent: parity unpinned), against
closed forms: order statistics by sorting, a fronto-parallel plane whose Gouraud shading is analytic."""
import math

import torch

from oracle import postprocess_oracle as po


def test_make_depth_image_order_statistics():
    g = torch.Generator().manual_seed(0)
    d = 5 + torch.rand(1, 1, 40, 50, generator=g) * 3
    m = (torch.rand(1, 1, 40, 50, generator=g) > 0.3).float()
    v, nf = po.make_depth_image(d, m)
    ok = d[m > 0.5].sort().values
    n = ok.numel()
    assert nf[0, 0] == ok[max(int(round(0.02 * n)), 1) - 1]          # k-th smallest
    assert nf[0, 1] == ok[n - max(int(round((1 - 0.98) * n)), 1)]    # k-th largest
    assert float(v.min()) >= 0.0 and float(v.max()) <= 1.0 and torch.all(v[m == 0] == 0)
    mid = (d - nf[0, 0]) / (nf[0, 1] - nf[0, 0]) * 0.8 + 0.1
    sel = (m > 0.5) & (mid > 0) & (mid < 1)
    assert torch.allclose(v[sel], mid[sel], atol=1e-6)


def test_shade_depth_plane_is_analytic():
    H = W = 48
    z0 = 7.0
    d = torch.full((H, W), z0)
    m = torch.ones(H, W)
    f, pp = (3.2, 3.2), (0.0, 0.0)
    out, used, k = po.shade_depth(d, m, f, pp)
    assert k == int(math.ceil(0.005 * math.sqrt(2) * H))
    # interior pixels (the box filter divides by the zero-padded area at the border, which bends the plane there)
    s = slice(k + 1, H - k - 1)
    xs = 1.0 - 2.0 * (torch.arange(W, dtype=torch.float64) + 0.5) / W
    yy, xx = torch.meshgrid(xs, xs, indexing="ij")
    p = torch.stack([xx / 3.2 * z0, yy / 3.2 * z0, torch.full_like(xx, z0)], -1)
    ndl = (z0 / p.norm(dim=-1))                      # normal (0, 0, -1), light direction -p / |p|
    spec = (2 * ndl ** 2 - 1).clamp(0) ** 128.0
    ref = torch.stack([0.5 + 0.3 * ndl + 0.2 * spec, 0.5 + 0.3 * ndl + 0.2 * spec, 0.5 + 0.3 * ndl + 0.2 * 0.9 * spec]).clamp(0, 1)
    assert used.all()
    assert (out[:, s, s].double() - ref[:, s, s]).abs().max() < 1e-5


def test_frame_u8_clip_and_round():
    x = torch.tensor([[[-0.5, 0.0, 0.5, 1.0, 2.0]]])
    assert po.frame_u8(x)[0, :, 0].tolist() == [0, 0, 128, 255, 255]
    assert po.frame_u8(x).shape == (1, 5, 3)
