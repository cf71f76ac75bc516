"""Generate the committed golden vectors.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

* ``unet_ref.npz``      -- outputs of the UNMODIFIED reference ``UNetModel`` (imported from /root/reference) for the
                           seeded fixtures of ``oracle.unet_oracle.make_unet_state_dict``; pins oracle + CUDA path.
* ``unet_keys.json``    -- the reference module's state-dict keys and shapes (checkpoint compatibility).
* ``diffusion_ref.npz`` -- the reference ``GaussianDiffusion`` tables and p_sample / q_sample outputs.
* ``ddim_ref.npz``      -- the reference ``GaussianDiffusion.ddim_sample`` (eta 0 and 0.5) / ``ddim_reverse_sample``
                           outputs (``python tests/golden/make_golden.py --only-ddim`` writes just this file).
* ``render_golden.npz`` -- outputs of ``oracle.render_oracle`` on a tiny scene (regression pin of the restatement;
                           the renderer has no reference-side vectors: parity unpinned, see oracle/__init__.py).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from holo_diffusion.guided_diffusion.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType,  # noqa: E402
                                                                ModelVarType, get_named_beta_schedule)
from holo_diffusion.guided_diffusion.unet import UNetModel  # noqa: E402

from fixtures import make_grid, make_mlp  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

CASES = {
    "base16": dict(in_ch=16, R=16, model_ch=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8),
                   heads=2, seed=2),
    "small8": dict(in_ch=8, R=8, model_ch=32, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(2,), heads=1, seed=5),
}


def ref_unet(c):
    net = UNetModel(image_size=c["R"], in_channels=c["in_ch"], model_channels=c["model_ch"], out_channels=c["in_ch"],
                    num_res_blocks=c["num_res_blocks"], attention_resolutions=c["attention_resolutions"], dropout=0.0,
                    channel_mult=c["channel_mult"], dims=3, num_heads=c["heads"], use_scale_shift_norm=True,
                    resblock_updown=False, zero_last_conv=False, homogeneous_resample=True)
    sd = uo.make_unet_state_dict(c["in_ch"], c["in_ch"], c["model_ch"], c["num_res_blocks"], c["channel_mult"],
                                 c["attention_resolutions"], seed=c["seed"])
    net.load_state_dict(sd, strict=True)
    return net.eval(), sd


def _ref_diffusion():
    return GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000, 0.0001, 0.02), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE, rescale_timesteps=False)


def make_ddim():
    gd = _ref_diffusion()
    g = torch.Generator().manual_seed(17)
    x = torch.randn(3, 4, 4, 4, 4, generator=g)
    model = lambda z, t: torch.tanh(1.7 * z) * 1.3  # noqa: E731  (exercises the clamp)
    t = torch.tensor([0, 412, 999])
    d = dict(x=x.numpy(), t=t.numpy())
    for k in ("alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod"):
        d[k] = getattr(gd, k)
    for eta in (0.0, 0.5):
        torch.manual_seed(123)  # ddim_sample draws th.randn_like(x) from the global generator
        r = gd.ddim_sample(model, x, t, clip_denoised=True, eta=eta)
        d[f"ddim_eta{eta}"] = r["sample"].numpy()
    torch.manual_seed(123)
    d["noise"] = torch.randn_like(x).numpy()
    r = gd.ddim_reverse_sample(model, x, t, clip_denoised=True)
    d["ddim_reverse"] = r["sample"].numpy()
    d["pred_xstart"] = r["pred_xstart"].numpy()
    np.savez_compressed(os.path.join(HERE, "ddim_ref.npz"), **d)


def main():
    make_ddim()
    if "--only-ddim" in sys.argv:
        return
    out = {}
    keys = {}
    for name, c in CASES.items():
        net, sd = ref_unet(c)
        keys[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
        x = make_grid(c["in_ch"], c["R"], seed=0)
        for t in (0, 500):
            with torch.no_grad():
                y = net(x, torch.full((1,), t, dtype=torch.long))
            out[f"{name}_t{t}"] = y.numpy().reshape(-1)[::4].copy()
            out[f"{name}_t{t}_stats"] = np.array([y.mean().item(), y.std().item(), y.abs().max().item()])
    np.savez_compressed(os.path.join(HERE, "unet_ref.npz"), **out)
    json.dump(keys, open(os.path.join(HERE, "unet_keys.json"), "w"), indent=0)

    gd = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000, 0.0001, 0.02), model_mean_type=ModelMeanType.START_X,
                           model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE, rescale_timesteps=False)
    d = {k: getattr(gd, k) for k in ("betas", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance",
                                     "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 4, 4, 4, 4, generator=g)
    noise = torch.randn(2, 4, 4, 4, 4, generator=g)
    model = lambda z, t: torch.tanh(1.7 * z) * 1.3  # noqa: E731  (exercises the clamp)
    t = torch.tensor([0, 637])
    ps = gd.p_sample(model, x, t, clip_denoised=True, noise_sampler=lambda *_: noise)
    d.update(x=x.numpy(), noise=noise.numpy(), t=t.numpy(), p_sample=ps["sample"].numpy(), pred_xstart=ps["pred_xstart"].numpy(),
             q_sample=gd.q_sample(x, t, noise).numpy())
    np.savez_compressed(os.path.join(HERE, "diffusion_ref.npz"), **d)

    C, R, HW, S = 8, 8, 12, 8
    grid, p = make_grid(C, R), make_mlp(C)
    cams = ro.simple_360_cameras(8)
    b = ro.sample_rays(cams[3], HW, HW, S)
    o = ro.render_chunked(p, grid, b, R, 8.0, 2, 4, chunk_size_grid=0)
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"), features=o.features.numpy(), depths=o.depths.numpy(),
                        masks=o.masks.numpy(), lengths=o.lengths.numpy(), prev_features=o.prev_stage.features.numpy(),
                        prev_weights=o.prev_stage.weights.numpy(), origins=b.origins.numpy(), directions=b.directions.numpy(),
                        coarse_lengths=b.lengths.numpy(), R=cams.R.numpy(), T=cams.T.numpy())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
