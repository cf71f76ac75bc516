#!/usr/bin/env python
"""BASELINE.json configs[3] (cfg #4): a batch of independent 64^3 x 32ch grids, one denoise step (DDPM ancestral step
at t = 500) + one 256^2 render each, sharded over the GPUs of one box with ONE NCCL gather of the finished
(n, 5, H, W) images to rank 0 (sharding.shard_units / gather_images).  Synthetic grids,
random-init weights (the package's own initialisers, same seed on every rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \\
        tools/batch_sharded.py [--batch 32 --resol 64 --channels 32 --image 256 --pts 64 --fine 16 --t 500]

Prints one JSON line from rank 0 (CUDA-event time of the whole batch incl. the gather, max over ranks).  Not a
bench.py line: the headline metric is measured on configs[1].
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--resol", type=int, default=64)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--image", type=int, default=256)
    ap.add_argument("--pts", type=int, default=64)
    ap.add_argument("--fine", type=int, default=16)
    ap.add_argument("--t", type=int, default=500)
    ap.add_argument("--repeats", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import holo_diffusion_b200 as hd
    from bench import UNET_ARGS
    from holo_diffusion_b200.sharding import gather_images, shard_units

    torch.manual_seed(2)   # identical random-init weights on every rank
    model = hd.HoloDiffusionModel(
        resol=a.resol, feature_size=a.channels, num_passes=2, render_image_width=a.image, render_image_height=a.image,
        net_3d_SimpleUnet3D_args=dict(UNET_ARGS), diffusion_args=dict(num_steps=1000),
        raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=a.pts),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
            n_pts_per_ray_fine_evaluation=a.fine, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))))
    with torch.no_grad():
        model._implicit_functions[0]._fn.render_mlp._density_net.mlp[-1][0].weight[-1] *= 8.0
    model.to(dev)
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, 8, -math.pi / 6, 10.0, (-0.0396, -0.8306, -0.5554), 3.2)
    mine = shard_units(a.batch, rank, world)
    shape = (1, a.channels, a.resol, a.resol, a.resol)
    grids = [torch.tanh(torch.randn(*shape, generator=torch.Generator().manual_seed(100 + u))).to(dev) for u in mine]
    t = torch.full((1,), a.t, device=dev, dtype=torch.int64)

    def one_batch():
        imgs = []
        for u, g in zip(mine, grids):
            x = model.diffusion.p_sample(model.net_3d, g, t)["sample"]          # one ancestral step at t
            preds = model(camera=cams[[u % 8]].to(dev), voxel_features=torch.clip(x, -1.0, 1.0))
            imgs.append(torch.cat([preds["images_render"][0], preds["depths_render"][0], preds["masks_render"][0]], 0))
        local = torch.stack(imgs) if imgs else torch.empty(0, 5, a.image, a.image, device=dev)
        return gather_images(local, a.batch, rank, world)

    out = one_batch()   # warm-up: CUDA graphs, packed weights, NCCL channels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.repeats):
        out = one_batch()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.repeats], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"cfg#4: {a.batch} independent {a.resol}^3 x {a.channels}ch grids, DDPM step at t={a.t} "
                                      f"+ {a.image}^2 render ({a.pts}+{a.fine} pts/ray) each, gathered to rank 0",
                          "n_gpus": world, "ms_per_batch": float(ms), "units_per_s": a.batch / (float(ms) / 1e3),
                          "gathered_shape": list(out.shape), "finite": bool(torch.isfinite(out).all())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
