#!/usr/bin/env python
"""BASELINE.json configs[2] (cfg #3): one full DDPM sample (1000 ancestral steps of the base-args UNet on a
64^3 x 32ch grid) followed by an 8-view turntable render -- what generate_samples.py does for one sample
(/root/reference/holo_diffusion/utils/render_utils/flyaround.py:215-260).  Synthetic (random-init) weights.

    python tools/sample_turntable.py [--steps 1000] [--views 8] [--ddim]

Prints one JSON line with the wall-clock (CUDA-event) time of the sampling loop and of the renders.  Not a bench.py
line: the headline metric is measured on configs[1] (see bench.py); this reports the whole-pipeline latency.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--resol", type=int, default=64)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--image", type=int, default=256)
    ap.add_argument("--ddim", action="store_true", help="deterministic DDIM (eta 0) instead of ancestral sampling")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    import holo_diffusion_b200 as hd
    from bench import UNET_ARGS

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(2)  # random-init weights from the package's own (reference-faithful) initialisers
    model = hd.HoloDiffusionModel(
        resol=a.resol, feature_size=a.channels, num_passes=2, render_image_width=a.image, render_image_height=a.image,
        net_3d_SimpleUnet3D_args=dict(UNET_ARGS), diffusion_args=dict(num_steps=1000),
        raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=64),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
            n_pts_per_ray_fine_evaluation=16, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))),
        use_cuda_graph=not a.no_graph)
    with torch.no_grad():
        model._implicit_functions[0]._fn.render_mlp._density_net.mlp[-1][0].weight[-1] *= 8.0
    model.to(dev)
    shape = (1, a.channels, a.resol, a.resol, a.resol)
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, a.views, -math.pi / 6, 10.0, (-0.0396, -0.8306, -0.5554), 3.2)
    torch.manual_seed(0)
    # warm-up: builds both CUDA graphs, packs the weights
    d = model.diffusion
    d.p_sample_loop(model.net_3d, shape, device=dev, max_iter=2)
    model(camera=cams[[0]].to(dev), voxel_features=torch.zeros(shape, device=dev))
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    if a.ddim:
        grid = torch.randn(*shape, device=dev)
        idx = [int(i) for i in torch.round(torch.linspace(999, 0, a.steps)).long()] if a.steps < 1000 else list(range(999, -1, -1))
        for i in idx:
            grid = d.ddim_sample(model.net_3d, grid, torch.full((1,), i, device=dev, dtype=torch.int64))["sample"]
    else:
        grid = d.p_sample_loop(model.net_3d, shape, device=dev, max_iter=a.steps if a.steps < 1000 else None)
    grid = torch.clip(grid, -1.0, 1.0)
    e[1].record()
    imgs = []
    for v in range(a.views):
        preds = model(camera=cams[[v]].to(dev), voxel_features=grid)
        imgs.append(preds["images_render"].clone())
    e[2].record()
    torch.cuda.synchronize()
    t_s, t_r = e[0].elapsed_time(e[1]) / 1e3, e[1].elapsed_time(e[2]) / 1e3
    img = torch.cat(imgs, 0)
    print(json.dumps({"workload": f"cfg#3: {a.steps}-step {'DDIM' if a.ddim else 'DDPM'} sample + {a.views}-view turntable, "
                                  f"{a.resol}^3 x {a.channels}ch, {a.image}^2, 64+16 pts/ray",
                      "sampling_s": t_s, "ms_per_denoise_step": 1e3 * t_s / a.steps, "render_views_s": t_r,
                      "ms_per_view": 1e3 * t_r / a.views, "total_s": t_s + t_r, "finite": bool(torch.isfinite(img).all()),
                      "image_mean": float(img.mean()), "cuda_graph": not a.no_graph}))


if __name__ == "__main__":
    main()
