#!/usr/bin/env python
"""BASELINE.json configs[4] (cfg #5): ONE grid, ONE view, several GPUs cooperating -- 128^3 x 32ch grid, UNet with
attention at every level, 512^2 render at 128 (+16) pts/ray.  Every rank evaluates the denoiser on the same input with
the queries of the large attention blocks split over the ranks (one all-gather per block), renders its block of image
rows and all-gathers the images (HoloDiffusionModel.shard_one_sample).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/one_sample_multi_gpu.py [--resol 128 --channels 32 --image 512 --pts 128 --fine 16] [--check]

--check (small sizes only) also runs the un-sharded model on rank 0's GPU and the CPU oracle and reports the
distances.  Prints one JSON line from rank 0 (CUDA-event time, max over ranks).  Not a bench.py line.
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
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resol", type=int, default=128)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--image", type=int, default=512)
    ap.add_argument("--pts", type=int, default=128)
    ap.add_argument("--fine", type=int, default=16)
    ap.add_argument("--attn-min-tokens", type=int, default=1 << 14)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import holo_diffusion_b200 as hd
    from fixtures import make_grid, make_mlp
    from oracle import render_oracle as ro
    from oracle import unet_oracle as uo   # fixture weights only (make_unet_state_dict); the oracle runs under --check

    C, R, HW = a.channels, a.resol, a.image
    unet = dict(model_channels=64, num_res_blocks=2, num_heads=2, channel_mult=[1, 1, 2, 4, 8],
                attention_resolutions=[1, 2, 4, 8, 16])

    def build():
        m = hd.HoloDiffusionModel(
            resol=R, feature_size=C, num_passes=2, render_image_width=HW, render_image_height=HW,
            net_3d_SimpleUnet3D_args=dict(unet), raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=a.pts),
            renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
                n_pts_per_ray_fine_evaluation=a.fine, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))),
            use_cuda_graph=False)
        m.net_3d._net.load_state_dict(sd, strict=True)
        m._implicit_functions[0]._fn.render_mlp.load_state_dict(mlp, strict=True)
        return m.to(dev)

    sd = uo.make_unet_state_dict(C, C, attention_resolutions=(1, 2, 4, 8, 16), seed=2)   # identical on every rank
    mlp = make_mlp(C)
    model = build()
    if world > 1:
        model.shard_one_sample(attn_min_tokens=a.attn_min_tokens)
    grid = make_grid(C, R, seed=0).to(dev)
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, 8, -math.pi / 6, 10.0, ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    cam = cams[[3]].to(dev)
    preds = model(camera=cam, voxel_features=grid)   # warm-up: packs weights, NCCL channels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        preds = model(camera=cam, voxel_features=grid)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = {"workload": f"cfg#5-style: one {R}^3 x {C}ch grid, UNet with attention at every level, {HW}^2 view, "
                       f"{a.pts}+{a.fine} pts/ray", "n_gpus": world, "ms_per_view": float(ms), "steps": a.steps,
           "sharding": "attention queries + image rows, all-gather" if world > 1 else "none"}
    if a.check:
        def rel(p, q):
            return float((p.double().cpu() - q.double().cpu()).abs().max() / q.double().cpu().abs().max())
        img = preds["images_render"].contiguous()
        if world > 1:   # every rank must hold the same full image
            ref_img = img.clone()
            dist.broadcast(ref_img, 0)
            same = torch.tensor([float(torch.equal(ref_img, img))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            rec["all_ranks_identical"] = bool(same.item())
        if rank == 0:
            single = build()(camera=cam, voxel_features=grid)
            rec["sharded_vs_single_gpu_image"] = rel(img, single["images_render"])
            rec["sharded_vs_single_gpu_grid"] = rel(preds["voxel_features"], single["voxel_features"])
            g = torch.tanh(uo.unet_forward(sd, grid.cpu(), torch.zeros(1, dtype=torch.long)))
            rec["grid_vs_oracle"] = rel(preds["voxel_features"], g)
            b = ro.sample_rays(ro.simple_360_cameras(8)[3], HW, HW, a.pts)
            ref = ro.render_chunked(mlp, preds["voxel_features"].cpu(), b, R, 8.0, 1, 0, chunk_size_grid=0)
            rec["coarse_image_vs_oracle_on_our_grid"] = None   # the 2-pass image is ill-conditioned
            one = hd.HoloDiffusionModel(
                resol=R, feature_size=C, num_passes=1, render_image_width=HW, render_image_height=HW, net_3d_enabled=False,
                raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=a.pts),
                renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
                    n_pts_per_ray_fine_evaluation=a.fine, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))),
                use_cuda_graph=False)
            one._implicit_functions[0]._fn.render_mlp.load_state_dict(mlp, strict=True)
            one.to(dev)
            p1 = one(camera=cam, voxel_features=preds["voxel_features"])
            rec["coarse_image_vs_oracle_on_our_grid"] = rel(p1["images_render"], ref.features.permute(0, 3, 1, 2))
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
