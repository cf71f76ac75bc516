"""Device time of the view-pooling encoder (views -> voxel grid) at the reference's size: 64^3 x 32ch grid from N source
views of 256^2, ResNet34 features (4 stages x 16 ch + mask + image = 68 columns), MLPMean or angle-weighted aggregator.

    python tools/encoder_bench.py [--views 10] [--aggregator mlp_mean|angle] [--chunk POINTS] [--iters 5]

Prints one JSON line: total ms per grid (feature extractor excluded and timed separately), the per-entry-point split
(CUDA events around every call, queued behind a device-side sleep so that launch latency stays out), algorithmic
FLOP/s of the Linear layers and the HBM bytes the intermediates would cost un-chunked.  No oracle, no CPU path."""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--aggregator", default="mlp_mean", choices=["mlp_mean", "angle"])
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--resol", type=int, default=64)
    a = ap.parse_args()
    if a.chunk:
        os.environ["HOLO_VIEWPOOL_CHUNK"] = str(a.chunk)
    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import encoder as en, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    n, R, C = a.views, a.resol, 32
    ext = hd.ResNetFeatureExtractor(proj_dim=16, image_rescale=0.32).to(dev).eval()    # configs/base.yaml:162-164
    imgs = torch.nn.functional.interpolate(torch.rand(n, 3, 50, 50), size=(800, 800), mode="bilinear").to(dev)
    fg = torch.nn.functional.interpolate(torch.rand(n, 1, 50, 50), size=(800, 800), mode="bilinear").to(dev)
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, n, -math.pi / 6, 10.0, (0.0, 1.0, 0.0), 3.2).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    with torch.no_grad():
        feats = ext(imgs, fg)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(3):
            feats = ext(imgs, fg)
        e1.record()
        torch.cuda.synchronize()
    ext_ms = e0.elapsed_time(e1) / 3
    if a.aggregator == "mlp_mean":
        pooler = hd.ViewPooler(feature_aggregator_class_type="MLPMeanFeatureAggregator").to(dev)
    else:
        pooler = hd.ViewPooler().to(dev)
    pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
    mapper = en.LazyLinearWithXavierInit(C).to(dev)
    pts = en.coord_grid(R, 8.0, dev)
    grid_cf = torch.empty(C * R ** 3, device=dev)

    def run():
        rows = en.pool_views(pooler, pts, cams, feats, None, None, mapper=mapper)
        ops.act_range(rows, R ** 3, C, 1, None, grid_cf, None)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / a.iters
    # ---- per entry point
    names = ["viewpool_sample", "viewpool_angle_reduce", "gemm_tc", "gemm_tc_act", "viewpool_act_split", "viewpool_reduce",
             "act_range"]
    spans = {k: [] for k in names}
    orig = {k: getattr(ops, k) for k in names}

    def wrap(k):
        def f(*args, **kw):
            s, t = ev(), ev()
            s.record()
            r = orig[k](*args, **kw)
            t.record()
            spans[k].append((s, t))
            return r
        return f

    for k in names:
        setattr(ops, k, wrap(k))
    torch.cuda._sleep(200_000_000)   # ~0.1 s: the host queues the whole grid behind it
    run()
    torch.cuda.synchronize()
    for k in names:
        setattr(ops, k, orig[k])
    split = {k: {"launches": len(v), "ms": round(sum(s.elapsed_time(t) for s, t in v), 4)} for k, v in spans.items() if v}
    rows_total = n * R ** 3
    lin_flops = (rows_total * 2.0 * (128 * 128 + 128 * 128) + R ** 3 * 2.0 * (128 * 128 + 128 * C)) if a.aggregator == "mlp_mean" \
        else R ** 3 * 2.0 * 192 * C
    gemm_ms = split.get("gemm_tc", {}).get("ms", 0.0) + split.get("gemm_tc_act", {}).get("ms", 0.0)
    print(json.dumps({
        "workload": f"view pooling: {R}^3 grid x {C}ch from {n} views, feature maps "
                    + ", ".join(f"{k} {tuple(v.shape[1:])}" for k, v in feats.items()),
        "aggregator": a.aggregator, "ms_per_grid": round(total_ms, 4), "grids_per_s": round(1e3 / total_ms, 2),
        "feature_extractor_ms": round(ext_ms, 3), "chunk_points": en._chunk_points(n), "per_entry_point": split,
        "linear_layers_algorithmic_tflops": round(lin_flops / max(gemm_ms, 1e-9) / 1e9, 1) if gemm_ms else None,
        "intermediate_bytes_unchunked": int(rows_total * 128 * (4 + 4 + 4 + 4)) if a.aggregator == "mlp_mean" else 0,
        "device": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
