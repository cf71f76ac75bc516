#!/usr/bin/env python
"""Parity check of the one-sample multi-GPU path (tools/one_sample_multi_gpu.py) at small sizes: all ranks hold the
same image, sharded vs un-sharded evaluation, denoiser and coarse render vs the oracle.  Test-side diagnostic (it
executes oracle/, so it lives under tests/).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/diagnostics/check_one_sample_multi_gpu.py --resol 16 --channels 16 --image 64 --pts 16 --attn-min-tokens 512
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import one_sample_multi_gpu as launcher  # noqa: E402
from fixtures import make_grid, make_mlp  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402


def rel(p, q):
    return float((p.double().cpu() - q.double().cpu()).abs().max() / q.double().cpu().abs().max())


def main():
    a = launcher.parse()
    rank, world, dev = launcher.init_dist()
    C, R, HW = a.channels, a.resol, a.image
    sd = uo.make_unet_state_dict(C, C, attention_resolutions=(1, 2, 4, 8, 16), seed=2)   # identical on every rank
    mlp = make_mlp(C)
    grid = make_grid(C, R, seed=0)
    preds, model, cam, grid_dev, rec = launcher.run(a, rank, world, dev, sd, mlp, grid)
    img = preds["images_render"].contiguous()
    if world > 1:   # every rank must hold the same full image
        ref_img = img.clone()
        dist.broadcast(ref_img, 0)
        same = torch.tensor([float(torch.equal(ref_img, img))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rec["all_ranks_identical"] = bool(same.item())
    if rank == 0:
        single = launcher.build_model(a, dev, sd, mlp)(camera=cam, voxel_features=grid_dev)
        rec["sharded_vs_single_gpu_image"] = rel(img, single["images_render"])
        rec["sharded_vs_single_gpu_grid"] = rel(preds["voxel_features"], single["voxel_features"])
        g = torch.tanh(uo.unet_forward(sd, grid, torch.zeros(1, dtype=torch.long)))
        rec["grid_vs_oracle"] = rel(preds["voxel_features"], g)
        # the 2-pass image is ill-conditioned: check the coarse pass on OUR grid
        b = ro.sample_rays(ro.simple_360_cameras(8)[3], HW, HW, a.pts)
        ref = ro.render_chunked(mlp, preds["voxel_features"].cpu(), b, R, 8.0, 1, 0, chunk_size_grid=0)
        one = launcher.build_model(a, dev, None, mlp, num_passes=1, net_3d_enabled=False)
        p1 = one(camera=cam, voxel_features=preds["voxel_features"])
        rec["coarse_image_vs_oracle_on_our_grid"] = rel(p1["images_render"], ref.features.permute(0, 3, 1, 2))
        # where the sharded-vs-single image gap comes from: the two grids differ at the 1e-6 level (another summation
        # order: the single-GPU attention shares its keys between CTAs, the query-sharded one does not); the COARSE
        # images of the two grids agree at that level, and the importance re-sampling of the second pass amplifies
        # it -- by as much as it amplifies the fp32 oracle's own rounding (its distance to the fp64 twin, below)
        p_s = one(camera=cam, voxel_features=single["voxel_features"])
        rec["coarse_image_sharded_grid_vs_single_grid"] = rel(p1["images_render"], p_s["images_render"])
        g_our = preds["voxel_features"].cpu()
        r32 = ro.render_chunked(mlp, g_our, b, R, 8.0, 2, a.fine, chunk_size_grid=0)
        b64 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), b.xys.double())
        r64 = ro.render_chunked({k: v.double() for k, v in mlp.items()}, g_our.double(), b64, R, 8.0, 2, a.fine, chunk_size_grid=0)
        rec["two_pass_image_fp32_oracle_vs_fp64_twin"] = rel(r32.features, r64.features)
        rec["two_pass_image_vs_fp64_twin"] = rel(img, r64.features.permute(0, 3, 1, 2))
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
