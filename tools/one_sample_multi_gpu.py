#!/usr/bin/env python
"""BASELINE.json configs[4] (cfg #5): ONE grid, ONE view, several GPUs cooperating -- 128^3 x 32ch grid, UNet with
attention at every level, 512^2 render at 128 (+16) pts/ray.  Every rank evaluates the denoiser on the same input with
the queries of the large attention blocks split over the ranks (one all-gather per block), renders its block of image
rows and all-gathers the images (HoloDiffusionModel.shard_one_sample).  Synthetic grid,
random-init weights (the package's own reference-faithful initialisers, same seed on every rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/one_sample_multi_gpu.py [--resol 128 --channels 32 --image 512 --pts 128 --fine 16]

Prints one JSON line from rank 0 (CUDA-event time, max over ranks).  Not a bench.py line.  The parity check of this
path against the oracle is tests/diagnostics/check_one_sample_multi_gpu.py (same launcher, small sizes).
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

UNET_ALL_LEVELS = dict(model_channels=64, num_res_blocks=2, num_heads=2, channel_mult=[1, 1, 2, 4, 8],
                       attention_resolutions=[1, 2, 4, 8, 16])
UP_AXIS = (-0.0396, -0.8306, -0.5554)   # visualize_reconstruction.py:35


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--resol", type=int, default=128)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--image", type=int, default=512)
    ap.add_argument("--pts", type=int, default=128)
    ap.add_argument("--fine", type=int, default=16)
    ap.add_argument("--attn-min-tokens", type=int, default=1 << 14)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--profile-attention", action="store_true",
                    help="CUDA-event time of every sharded attention block: fused kernel on this rank's queries, all-gather")
    return ap.parse_args(argv)


def init_dist():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    return rank, world, dev


def build_model(a, dev, unet_sd=None, mlp_sd=None, num_passes=2, net_3d_enabled=True):
    import holo_diffusion_b200 as hd
    torch.manual_seed(2)   # identical random-init weights on every rank
    m = hd.HoloDiffusionModel(
        resol=a.resol, feature_size=a.channels, num_passes=num_passes, render_image_width=a.image,
        render_image_height=a.image, net_3d_enabled=net_3d_enabled, net_3d_SimpleUnet3D_args=dict(UNET_ALL_LEVELS),
        raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=a.pts),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
            n_pts_per_ray_fine_evaluation=a.fine, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))),
        use_cuda_graph=False)
    if unet_sd is not None:
        m.net_3d._net.load_state_dict(unet_sd, strict=True)
    if mlp_sd is not None:
        m._implicit_functions[0]._fn.render_mlp.load_state_dict(mlp_sd, strict=True)
    else:
        with torch.no_grad():   # a density head that exercises compositing
            m._implicit_functions[0]._fn.render_mlp._density_net.mlp[-1][0].weight[-1] *= 8.0
    return m.to(dev)


def run(a, rank, world, dev, unet_sd=None, mlp_sd=None, grid=None):
    """Returns (preds of the last step, model, camera, grid, record)."""
    import holo_diffusion_b200 as hd
    model = build_model(a, dev, unet_sd, mlp_sd)
    if world > 1:
        model.shard_one_sample(attn_min_tokens=a.attn_min_tokens)
    if grid is None:
        g = torch.Generator().manual_seed(0)
        grid = torch.tanh(torch.randn(1, a.channels, a.resol, a.resol, a.resol, generator=g))
    grid = grid.to(dev)
    cam = hd.get_simple_360_camera_trajectory(2 * math.pi, 8, -math.pi / 6, 10.0, UP_AXIS, 3.2)[[3]].to(dev)
    preds = model(camera=cam, voxel_features=grid)   # warm-up: packs weights, NCCL channels
    torch.cuda.synchronize()
    prof = []
    if a.profile_attention and world > 1:
        from holo_diffusion_b200 import ops
        ex = model.net_3d._exec
        orig_flash, orig_gather = ops.attention_flash, dist.all_gather_into_tensor

        def flash(*args, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_flash(*args, **kw)
            e.record()
            prof.append(["attn", args[4], s, e])
            return r

        def gather(out, inp, group=None, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_gather(out, inp, group=group, **kw)
            e.record()
            prof.append(["gather", out.numel() * out.element_size(), s, e])
            return r

        ops.attention_flash, dist.all_gather_into_tensor = flash, gather
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
    rec = {"workload": f"cfg#5-style: one {a.resol}^3 x {a.channels}ch grid, UNet with attention at every level, "
                       f"{a.image}^2 view, {a.pts}+{a.fine} pts/ray", "n_gpus": world, "ms_per_view": float(ms),
           "steps": a.steps, "sharding": "attention queries + image rows, all-gather" if world > 1 else "none"}
    if prof:
        ops.attention_flash, dist.all_gather_into_tensor = orig_flash, orig_gather
        n = max(1, a.steps)
        att = [(T, s.elapsed_time(e)) for k, T, s, e in prof if k == "attn"]
        gat = [(b, s.elapsed_time(e)) for k, b, s, e in prof if k == "gather"]
        by_T = {}
        for T, t in att:
            d = by_T.setdefault(int(T), [0, 0.0])
            d[0] += 1
            d[1] += t
        rec["rank0_attention_blocks"] = {str(T): {"launches_per_view": c // n, "ms_per_launch": t / c} for T, (c, t) in sorted(by_T.items())}
        rec["rank0_attention_ms_per_view"] = sum(t for _, t in att) / n
        rec["rank0_all_gather_ms_per_view"] = sum(t for _, t in gat) / n
        rec["rank0_all_gather_bytes_per_view"] = sum(b for b, _ in gat) // n
        rec["rank0_all_gather_calls_per_view"] = len(gat) // n
    return preds, model, cam, grid, rec


def main():
    a = parse()
    rank, world, dev = init_dist()
    _, _, _, _, rec = run(a, rank, world, dev)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
