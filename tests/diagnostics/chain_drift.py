"""Drift of the DDPM sampling chain (BASELINE cfg #3: 1000 ancestral steps) between our kernels and the oracle
restatement evaluated with torch ops on the same GPU (fp64 with --f64, else fp32 with TF32 off), both fed the SAME
injected noise through `noise_sampler` (gaussian_diffusion.py:495-498,597-602).  Prints one JSON object with the
relative distance of x_t every `--every` steps and of the final sample.  Test-side diagnostic (executes oracle/).

  python tests/diagnostics/chain_drift.py --resol 32 --steps 1000 --f64
"""
import argparse
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import diffusion_oracle as do  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resol", type=int, default=32)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--steps", type=int, default=1000, help="< 1000: max_iter sub-sampling of the 1000-step schedule")
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--with-eager32", action="store_true", help="also run the fp32 eager chain (TF32 off): its distance "
                    "to the fp64 chain is the yardstick for ours")
    a = ap.parse_args()
    import holo_diffusion_b200 as hd
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    C, R = a.channels, a.resol
    kw = dict(model_channels=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), num_heads=2)
    sd = uo.make_unet_state_dict(C, C, seed=2)
    net = hd.SimpleUnet3D(image_size=R, in_channels=C, out_channels=C, use_cuda_graph=True, **kw)
    net._net.load_state_dict(sd, strict=True)
    net.cuda()
    diff = hd.ImplicitronGaussianDiffusion()
    dt = torch.float64 if a.f64 else torch.float32
    sd_ref = {k: v.cuda().to(dt) for k, v in sd.items()}
    tab = do.schedule_tables()
    shape = (1, C, R, R, R)
    gen = torch.Generator().manual_seed(3)
    noises = {}

    def noise_sampler(t, shp, device):   # one draw per timestep, shared by both chains
        if t not in noises:
            noises[t] = torch.randn(*shp, generator=gen)
        return noises[t].to(device)

    indices = list(range(1000))[::-1]
    if a.steps < 1000:
        indices = [indices[int(i)] for i in torch.round(torch.linspace(0, 999, a.steps)).long()]
    x_ours = noise_sampler(1000, shape, "cuda")
    x_ref = x_ours.to(dt)
    x_e32 = x_ours.clone() if a.with_eager32 else None
    sd32 = {k: v.cuda() for k, v in sd.items()} if a.with_eager32 else None
    trace = []
    t0 = time.perf_counter()
    with torch.no_grad():
        for n, i in enumerate(indices):
            t = torch.full((1,), i, device="cuda", dtype=torch.int64)
            nz = noise_sampler(i, shape, "cuda")
            x_ours = diff.p_sample(net, x_ours, t, noise_sampler=lambda *_: nz)["sample"]
            # the oracle step on the device (diffusion_oracle.p_sample indexes host tables: restated inline for cuda)
            x0 = uo.unet_forward(sd_ref, x_ref, t).clamp(-1, 1)
            c1, c2 = float(tab["posterior_mean_coef1"][i]), float(tab["posterior_mean_coef2"][i])
            lv = float(tab["posterior_log_variance_clipped"][i])
            if not a.f64:   # _extract_into_tensor: fp64 table -> .float()
                c1, c2, lv = (float(torch.tensor(v, dtype=torch.float64).float()) for v in (c1, c2, lv))
            x_ref = c1 * x0 + c2 * x_ref + (0.0 if i == 0 else 1.0) * torch.exp(torch.tensor(0.5 * lv, dtype=dt, device="cuda")) * nz.to(dt)
            if x_e32 is not None:
                c32 = [float(torch.tensor(float(tab[k][i]), dtype=torch.float64).float()) for k in
                       ("posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped")]
                x0e = uo.unet_forward(sd32, x_e32, t).clamp(-1, 1)
                x_e32 = c32[0] * x0e + c32[1] * x_e32 + (0.0 if i == 0 else 1.0) * math.exp(0.5 * c32[2]) * nz
            if (n + 1) % a.every == 0 or n + 1 == len(indices):
                den = x_ref.double().abs().max()
                e = float((x_ours.double() - x_ref.double()).abs().max() / den)
                row = {"step": n + 1, "t": i, "rel_err": e}
                if x_e32 is not None:
                    row["eager32_rel_err"] = float((x_e32.double() - x_ref.double()).abs().max() / den)
                trace.append(row)
    torch.cuda.synchronize()
    out = {"workload": f"DDPM chain, {len(indices)} steps, {R}^3 x {C}ch, base UNet args, identical injected noise",
           "reference": "oracle restatement, torch ops on the same GPU, " + ("fp64" if a.f64 else "fp32 (TF32 off)"),
           "final_rel_err": trace[-1]["rel_err"], "final_eager32_rel_err": trace[-1].get("eager32_rel_err"), "max_rel_err": max(x["rel_err"] for x in trace), "trace": trace,
           "final_clipped_rel_err": float((x_ours.clamp(-1, 1).double() - x_ref.clamp(-1, 1).double()).abs().max()),
           "seconds": time.perf_counter() - t0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
