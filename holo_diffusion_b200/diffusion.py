"""DDPM schedule + ancestral sampler on the device, behind the reference's ``ImplicitronGaussianDiffusion`` surface.

Mirrors /root/reference/holo_diffusion/utils/diffusion_utils.py:89-140 and the parts of
/root/reference/holo_diffusion/guided_diffusion/gaussian_diffusion.py it reaches (``__init__`` :129-187,
``q_sample`` :209-227, ``p_mean_variance`` :253-355, ``p_sample`` :459-508, ``p_sample_loop[_progressive]`` :510-643)
for the shipped setting START_X / FIXED_SMALL.  The fp64 schedule is built once on the host (numpy, as the
reference does) and kept resident on the device as fp32 tables; each step is one fused kernel
(``holo_ddpm_step``) indexed by the device-side timestep -- no per-step H2D copies, no host syncs.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np
import torch

from . import ops


class ImplicitronGaussianDiffusion:
    def __init__(self, beta_schedule_type: str = "linear", num_steps: int = 1000, beta_start_unscaled: float = 0.0001,
                 beta_end_unscaled: float = 0.02, model_mean_type: str = "START_X", model_var_type: str = "FIXED_SMALL",
                 schedule_sampler_type: str = "uniform"):
        if beta_schedule_type != "linear":
            raise NotImplementedError("only the linear beta schedule of the shipped configs")
        if str(model_mean_type).split(".")[-1] != "START_X" or str(model_var_type).split(".")[-1] != "FIXED_SMALL":
            raise NotImplementedError("only START_X / FIXED_SMALL (configs/base.yaml:103-104)")
        if schedule_sampler_type != "uniform":
            raise NotImplementedError("only the uniform timestep sampler")
        self.num_timesteps = int(num_steps)
        scale = 1000 / num_steps
        betas = np.linspace(scale * beta_start_unscaled, scale * beta_end_unscaled, num_steps, dtype=np.float64)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.tables64 = {
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
        self._dev: Dict[str, Dict[str, torch.Tensor]] = {}

    def _tables(self, device) -> Dict[str, torch.Tensor]:
        key = str(device)
        if key not in self._dev:
            # _extract_into_tensor: fp64 table -> indexed -> .float(); the fp32 table holds the same values
            self._dev[key] = {k: torch.from_numpy(v).float().to(device) for k, v in self.tables64.items()}
        return self._dev[key]

    # ------------------------------------------------------------------ schedule sampler (uniform)
    def sample_timesteps(self, batch_size: int, device):
        w = np.ones([self.num_timesteps])
        p = w / np.sum(w)
        idx = np.random.choice(len(p), size=(batch_size,), p=p)
        t = torch.from_numpy(idx).long().to(device)
        return t, torch.from_numpy(1 / (len(p) * p[idx])).float().to(device)

    # ------------------------------------------------------------------ diffusion
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        tab = self._tables(x_start.device)
        out = torch.empty_like(x_start)
        ops.q_sample(x_start.contiguous(), noise.contiguous(), t.contiguous(), tab["sqrt_alphas_cumprod"],
                     tab["sqrt_one_minus_alphas_cumprod"], out)
        return out

    def p_mean_variance(self, model: Callable, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn")
        model_out = model(x, t, **(model_kwargs or {}))
        tab = self._tables(x.device)
        mean = torch.empty_like(x)
        x0 = torch.empty_like(x)
        ops.ddpm_step(model_out.contiguous(), x.contiguous(), None, t.contiguous(), tab["posterior_mean_coef1"],
                      tab["posterior_mean_coef2"], tab["posterior_log_variance_clipped"], clip_denoised, mean, x0)
        shape = [-1] + [1] * (x.ndim - 1)
        return {"mean": mean, "variance": tab["posterior_variance"][t].view(shape).expand_as(x),
                "log_variance": tab["posterior_log_variance_clipped"][t].view(shape).expand_as(x), "pred_xstart": x0}

    def p_sample(self, model: Callable, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 noise_sampler=None):
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn")
        model_out = model(x, t, **(model_kwargs or {}))
        noise = noise_sampler(int(t[0].item()), x.shape, x.device) if noise_sampler is not None else torch.randn_like(x)
        tab = self._tables(x.device)
        sample = torch.empty_like(x)
        x0 = torch.empty_like(x)
        ops.ddpm_step(model_out.contiguous(), x.contiguous(), noise.contiguous(), t.contiguous(),
                      tab["posterior_mean_coef1"], tab["posterior_mean_coef2"], tab["posterior_log_variance_clipped"],
                      clip_denoised, sample, x0)
        return {"sample": sample, "pred_xstart": x0, "noise": noise}

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, max_iter=None, noise_sampler=None):
        if device is None:
            device = next(model.parameters()).device
        if noise is not None:
            img = noise
        elif noise_sampler is not None:
            img = noise_sampler(self.num_timesteps, shape, device)
        else:
            img = torch.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if max_iter is not None and len(indices) > max_iter:
            if max_iter == 1:
                indices = [indices[0]]
            else:
                indices = [indices[int(i)] for i in torch.round(torch.linspace(0, len(indices) - 1, max_iter)).long()]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        # all timesteps live on the device up front: no per-step H2D
        for i in indices:
            t = torch.full((shape[0],), i, device=device, dtype=torch.int64)
            with torch.no_grad():
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                    model_kwargs=model_kwargs, noise_sampler=noise_sampler)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ DDIM (gaussian_diffusion.py:645-815)
    def ddim_sample(self, model: Callable, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                    eta=0.0):
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn")
        model_out = model(x, t, **(model_kwargs or {}))
        noise = torch.randn_like(x)  # drawn whatever eta is, as the reference does (:682)
        tab = self._tables(x.device)
        sample, x0 = torch.empty_like(x), torch.empty_like(x)
        ops.ddim_step(model_out.contiguous(), x.contiguous(), noise, t.contiguous(), tab["alphas_cumprod"],
                      tab["alphas_cumprod_prev"], tab["sqrt_recip_alphas_cumprod"], tab["sqrt_recipm1_alphas_cumprod"],
                      eta, clip_denoised, sample, x0)
        return {"sample": sample, "pred_xstart": x0}

    def ddim_reverse_sample(self, model: Callable, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        assert eta == 0.0, "Reverse ODE only for deterministic path"
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn")
        model_out = model(x, t, **(model_kwargs or {}))
        tab = self._tables(x.device)
        sample, x0 = torch.empty_like(x), torch.empty_like(x)
        ops.ddim_step(model_out.contiguous(), x.contiguous(), None, t.contiguous(), tab["alphas_cumprod"],
                      tab["alphas_cumprod_next"], tab["sqrt_recip_alphas_cumprod"], tab["sqrt_recipm1_alphas_cumprod"],
                      0.0, clip_denoised, sample, x0)
        return {"sample": sample, "pred_xstart": x0}

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else torch.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = torch.full((shape[0],), i, device=device, dtype=torch.int64)
            with torch.no_grad():
                out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                       cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)
                yield out
                img = out["sample"]

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0):
        final = None
        for sample in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                        denoised_fn=denoised_fn, cond_fn=cond_fn,
                                                        model_kwargs=model_kwargs, device=device, progress=progress,
                                                        eta=eta):
            final = sample
        return final["sample"]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, return_all_samples=False, max_iter=None,
                      noise_sampler=None):
        samples = []
        final = None
        for s in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                                                device=device, progress=progress, max_iter=max_iter,
                                                noise_sampler=noise_sampler):
            if return_all_samples:
                samples.append(s)
            final = s["sample"]
        return (final, samples) if return_all_samples else final
