"""Shared seeded fixtures.  Oracle-side only helpers live in oracle/."""
import torch

from oracle import render_oracle as ro


def make_grid(C, R, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.tanh(torch.randn(1, C, R, R, R, generator=g))


def make_mlp(C, seed=1, density_scale=8.0, empty_space_logit=-0.25):
    """Xavier RenderMLP with the density head scaled so that compositing is exercised, and its bias shifted so
    that a zero feature (a point outside the grid) decodes to an empty-space density (<= 0)."""
    p = ro.make_render_mlp_params(C, seed=seed, density_scale=density_scale, density_bias=0.0)
    z = torch.zeros(1, C)
    with torch.no_grad():
        y = z
        for li in range(4):
            if li == 2:
                y = torch.cat((y, z), -1)
            y = torch.nn.functional.linear(y, p[f"_density_net.mlp.{li}.0.weight"], p[f"_density_net.mlp.{li}.0.bias"])
        p["_density_net.mlp.3.0.bias"][-1] += empty_space_logit - y[0, -1]
    return p
