"""CPU restatement of the DDPM schedule and ancestral step (TEST INFRASTRUCTURE ONLY).

Follows ``holo_diffusion/guided_diffusion/gaussian_diffusion.py``: ``get_named_beta_schedule`` :25-51
(linear), ``GaussianDiffusion.__init__`` :150-187 (fp64 tables), ``q_sample`` :209-227,
``q_posterior_mean_variance`` :229-251, ``p_mean_variance`` :253-355 (START_X / FIXED_SMALL, clip),
``p_sample`` :459-508, ``ddim_sample`` / ``ddim_reverse_sample`` :645-731, ``_extract_into_tensor`` :1046-1059 (table fp64 -> indexed -> .float()).
Pinned against the imported reference by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np
import torch


def schedule_tables(num_steps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02) -> Dict[str, np.ndarray]:
    scale = 1000 / num_steps
    betas = np.linspace(scale * beta_start, scale * beta_end, num_steps, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "alphas_cumprod_next": np.append(ac[1:], 0.0),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


def _ext(arr: np.ndarray, t: torch.Tensor, ndim: int) -> torch.Tensor:
    r = torch.from_numpy(arr)[t].float()
    return r.reshape(-1, *([1] * (ndim - 1)))


def q_sample(tab, x0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    return _ext(tab["sqrt_alphas_cumprod"], t, x0.ndim) * x0 + _ext(tab["sqrt_one_minus_alphas_cumprod"], t, x0.ndim) * noise


def p_sample(tab, model: Callable, x: torch.Tensor, t: torch.Tensor, noise: torch.Tensor, clip: bool = True):
    """One ancestral step; returns dict(sample, pred_xstart)."""
    x0 = model(x, t)
    if clip:
        x0 = x0.clamp(-1, 1)
    mean = _ext(tab["posterior_mean_coef1"], t, x.ndim) * x0 + _ext(tab["posterior_mean_coef2"], t, x.ndim) * x
    logvar = _ext(tab["posterior_log_variance_clipped"], t, x.ndim)
    nz = (t != 0).float().view(-1, *([1] * (x.ndim - 1)))
    return {"sample": mean + nz * torch.exp(0.5 * logvar) * noise, "pred_xstart": x0}


def ddim_sample(tab, model: Callable, x: torch.Tensor, t: torch.Tensor, noise: torch.Tensor, eta: float = 0.0,
                clip: bool = True, reverse: bool = False):
    """ddim_sample :645-693 (reverse=False) / ddim_reverse_sample :695-731 (reverse=True, eta must be 0)."""
    x0 = model(x, t)
    if clip:
        x0 = x0.clamp(-1, 1)
    eps = (_ext(tab["sqrt_recip_alphas_cumprod"], t, x.ndim) * x - x0) / _ext(tab["sqrt_recipm1_alphas_cumprod"], t, x.ndim)
    ab = _ext(tab["alphas_cumprod"], t, x.ndim)
    if reverse:
        abn = _ext(tab["alphas_cumprod_next"], t, x.ndim)
        return {"sample": x0 * torch.sqrt(abn) + torch.sqrt(1 - abn) * eps, "pred_xstart": x0}
    abp = _ext(tab["alphas_cumprod_prev"], t, x.ndim)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    nz = (t != 0).float().view(-1, *([1] * (x.ndim - 1)))
    return {"sample": mean + nz * sigma * noise, "pred_xstart": x0}
