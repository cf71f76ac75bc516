"""The view-pooling encoder as PyTorch eager ops ON THE SAME GPU (the oracle's restatement moved to cuda: grid_sample +
Linear layers, chunked over the grid points so that the (views x points x 128) activations fit) next to the kernels,
at the reference's size.  Test infrastructure (uses oracle/): prints one JSON line.

    python tests/diagnostics/encoder_eager_compare.py [--views 10] [--chunk 32768]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=32768)
    a = ap.parse_args()
    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import encoder as en, ops
    from oracle import encoder_oracle as eo
    from oracle import render_oracle as ro
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    n, R, C = a.views, 64, 32
    ext = hd.ResNetFeatureExtractor(proj_dim=16, image_rescale=0.32).to(dev).eval()
    imgs = torch.nn.functional.interpolate(torch.rand(n, 3, 50, 50), size=(800, 800), mode="bilinear").to(dev)
    fg = torch.nn.functional.interpolate(torch.rand(n, 1, 50, 50), size=(800, 800), mode="bilinear").to(dev)
    with torch.no_grad():
        feats = ext(imgs, fg)
    oc = ro.simple_360_cameras(n, focal_length=3.2)
    ocd = ro.OracleCameras(oc.R.to(dev), oc.T.to(dev), oc.focal.to(dev), oc.pp.to(dev))
    cams = hd.PerspectiveCameras(oc.focal.clone(), oc.pp.clone(), oc.R.clone(), oc.T.clone()).to(dev)
    sd = {k: v.to(dev) for k, v in eo.make_aggregator_params(68 + 21, seed=31).items()}
    pooler = hd.ViewPooler(feature_aggregator_class_type="MLPMeanFeatureAggregator").to(dev)
    pooler.feature_aggregator.load_state_dict(sd)
    pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
    mapper = en.LazyLinearWithXavierInit(C).to(dev)
    pts = en.coord_grid(R, 8.0, dev)
    grid_cf = torch.empty(C * R ** 3, device=dev)

    def ours():
        rows = en.pool_views(pooler, pts, cams, feats, None, None, mapper=mapper)
        ops.act_range(rows, R ** 3, C, 1, None, grid_cf, None)
        return grid_cf.view(1, C, R, R, R)

    def eager():
        rows = []
        for i in range(0, pts.shape[0], a.chunk):
            pc = pts[i:i + a.chunk]
            fs, ms = eo.sample_views(ocd, pc, feats, None, False)
            rows.append(eo.mlp_mean_aggregate(sd, fs, ms, ocd, pc))
        pooled = torch.cat(rows, 2)
        v = torch.nn.functional.linear(pooled, mapper.weight.detach(), mapper.bias.detach()).permute(0, 3, 1, 2)
        return torch.tanh(v.reshape(1, -1, R, R, R))

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, out

    with torch.no_grad():
        t_ours, g_ours = timed(ours, 5)
        g_ours = g_ours.clone()
        t_eager, g_eager = timed(eager, 2)
    err = float((g_ours - g_eager).abs().max() / g_eager.abs().max())
    print(json.dumps({"workload": f"view pooling, 64^3 x 32ch grid from {n} views (ResNet34 features, MLPMean aggregator)",
                      "ours_ms": round(t_ours, 3), "eager_gpu_ms": round(t_eager, 3), "speedup": round(t_eager / t_ours, 2),
                      "eager_chunk_points": a.chunk, "ours_vs_eager_rel": err, "device": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
