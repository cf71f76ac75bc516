"""View-pooling encoder (views -> voxel grid) on the GPU: every stage of the fused pooling against the oracle, the
model's encoder branch against the reference-forward golden vectors, and the full-size grid (64^3, 10 source views)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import encoder_oracle as eo
from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "encoder_intree_ref.npz")


def _cams_gpu(c):
    import holo_diffusion_b200 as hd
    return hd.PerspectiveCameras(c.focal.clone(), c.pp.clone(), c.R.clone(), c.T.clone()).to("cuda")


def _gpu(d):
    return {k: v.cuda() for k, v in d.items()}


def _mlp_pooler(sd, n_hidden, dim_out, n_layers, masked):
    import holo_diffusion_b200 as hd
    p = hd.ViewPooler(view_sampler_args=dict(masked_sampling=masked), feature_aggregator_class_type="MLPMeanFeatureAggregator",
                      feature_aggregator_MLPMeanFeatureAggregator_args=dict(n_hidden=n_hidden, dim_out=dim_out, n_layers=n_layers))
    p.feature_aggregator.load_state_dict(sd)
    p.feature_aggregator.exclude_target_view = p.feature_aggregator.exclude_target_view_mask_features = False
    return p.cuda()


def _points(n, seed, spread=5.0):
    """Points around the grid volume; the first few sit behind / beside the cameras (clamped divide, zeros padding)."""
    pts = (torch.rand(n, 3, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * spread
    pts[0] = torch.tensor([0.0, 0.0, 0.0])
    pts[1] = torch.tensor([30.0, -20.0, 25.0])
    pts[2] = torch.tensor([-12.0, 9.0, -11.0])
    return pts


@pytest.mark.parametrize("hw,masked,n_layers", [((40, 40), False, 1), ((24, 56), True, 3), ((56, 24), True, 2)])
def test_pooling_stages_match_oracle(hw, masked, n_layers):
    """Every intermediate of pool_views (sampled rows X, their weighted mean, the first Linear pair, the logits Z, the
    pooled feature, the mapped rows) against the oracle: square and both non-square aspect ratios, masked sampling,
    a camera from another sequence (view weight 0), 1 / 2 / 3 MLP layers, a ragged last chunk."""
    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import encoder as en
    os.environ["HOLO_VIEWPOOL_CHUNK"] = "256"
    try:
        n_src, H, D, Cg = 5, 128, 128, 32
        cams, feats, mask_crop = eo.make_views(n_src, hw, seed=11)
        Kx = 64 + 1 + 3 + 21
        sd = eo.make_aggregator_params(Kx, H, D, n_layers, seed=12)
        pooler = _mlp_pooler(sd, H, D, n_layers, masked)
        mapper = en.LazyLinearWithXavierInit(Cg).cuda()
        pts = _points(600, 13)
        vw = torch.tensor([1.0, 1.0, 1.0, 0.0, 1.0])
        dbg = {}
        rows = en.pool_views(pooler, pts.cuda(), _cams_gpu(cams), _gpu(feats), mask_crop.cuda(), vw.cuda(), mapper=mapper, debug=dbg)
        torch.cuda.synchronize()
        assert rows.shape == (600, Cg)
        # ---- oracle, stage by stage (first chunk = 256 points for the debug tensors)
        n0 = 256
        fs, ms = eo.sample_views(cams, pts[:n0], feats, mask_crop, masked, view_weight=vw)
        w = ms[0, ..., 0]
        ray = ro.harmonic_embedding(eo.point_to_camera_ray_dirs(cams, pts[:n0]), 3)[0]
        x = torch.cat([*[f[0] for f in fs.values()], ray], -1) * w[..., None]
        mean = eo.wmean(x[None], w[None])[0, 0]
        e = {"x": rel_err(dbg["x"], x), "mean": rel_err(dbg["mean"], mean)}
        pooled = eo.mlp_mean_aggregate(sd, fs, ms, cams, pts[:n0])[0, 0]
        e["pooled"] = rel_err(dbg["pooled"], pooled)
        ref = eo.encode(cams, feats, mask_crop, sd, mapper.weight.detach().cpu(), mapper.bias.detach().cpu(), 0, 0.0,
                        masked_sampling=masked, view_weight=vw, pts=pts)[0, :, 0].t()
        e["grid"] = rel_err(torch.tanh(rows), ref)
        print(f"view pooling {hw} masked={masked} layers={n_layers}:", {k: f"{v:.2e}" for k, v in e.items()})
        assert e["x"] < 1e-5 and e["mean"] < 1e-5, e   # noise images: 1e-7 on the projected position is 1e-6 on a tap
        assert e["pooled"] < 1e-4 and e["grid"] < 1e-4, e
        assert w.min() == 0 and (x[:, :3, :64].abs().sum(-1) == 0).any()   # the test does reach the zero-padding cases
    finally:
        os.environ.pop("HOLO_VIEWPOOL_CHUNK", None)


def test_angle_weighted_pooling_matches_oracle():
    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import encoder as en
    cams, feats, mask_crop = eo.make_views(6, (48, 32), seed=21)
    for red in (("AVG", "STD"), ("AVG",)):
        pooler = hd.ViewPooler(feature_aggregator_AngleWeightedReductionFeatureAggregator_args=dict(reduction_functions=red))
        pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
        mapper = en.LazyLinearWithXavierInit(32).cuda()
        pts = _points(500, 22)
        dbg = {}
        rows = en.pool_views(pooler, pts.cuda(), _cams_gpu(cams), _gpu(feats), mask_crop.cuda(), None, mapper=mapper, debug=dbg)
        fs, ms = eo.sample_views(cams, pts, feats, mask_crop, False)
        pooled = eo.angle_weighted_aggregate(fs, ms, cams, pts, with_std=len(red) == 2)[0, 0]
        ref = torch.tanh(torch.nn.functional.linear(pooled, mapper.weight.detach().cpu(), mapper.bias.detach().cpu()))
        e = (rel_err(dbg["pooled"], pooled[: dbg["pooled"].shape[0]]), rel_err(torch.tanh(rows), ref))
        print(f"angle-weighted pooling {red}: pooled {e[0]:.2e}, grid {e[1]:.2e}")
        assert e[0] < 1e-5 and e[1] < 1e-4, e


def test_model_encoder_branch_matches_reference_forward():
    """HoloDiffusionModel.forward(image_rgb=...) on the kernels reproduces the grid the reference's own forward binds
    (tests/golden/encoder_intree_ref.npz: the reference's source executed over the stub)."""
    import holo_diffusion_b200 as hd
    gold = np.load(GOLD, allow_pickle=False)
    t = lambda a: torch.from_numpy(np.asarray(a))   # noqa: E731
    sel = [1, 3, 4]
    img, fg = t(gold["fwd/image_rgb"]), t(gold["fwd/fg"])
    feats = {k[len("fwd/feats/"):]: t(gold[k])[sel].cuda() for k in gold.files if k.startswith("fwd/feats/")}
    m = hd.HoloDiffusionModel(
        resol=4, feature_size=8, num_passes=1, render_image_width=8, render_image_height=8, net_3d_enabled=False,
        diffusion_enabled=False, use_cuda_graph=False, view_pooler_enabled=True,
        view_pooler_args=dict(feature_aggregator_class_type="MLPMeanFeatureAggregator",
                              feature_aggregator_MLPMeanFeatureAggregator_args=dict(n_hidden=16, dim_out=8)),
        raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=4),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(n_pts_per_ray_fine_evaluation=2))
    m.view_pooler.feature_aggregator.load_state_dict({str(k): t(gold["fwd/agg/" + str(k)]) for k in gold["fwd/agg_keys"]})
    m.pooled_feature_mapper.load_state_dict({"weight": t(gold["fwd/mapper_w"]), "bias": t(gold["fwd/mapper_b"])})

    class Extractor(torch.nn.Module):
        def forward(self, imgs, masks):
            return dict(feats, mask=masks, image=imgs)

    m.image_feature_extractor = Extractor()
    m = m.cuda()
    cams = hd.PerspectiveCameras(t(gold["fwd/cam_focal"]), t(gold["fwd/cam_pp"]), t(gold["fwd/cam_R"]), t(gold["fwd/cam_T"])).to("cuda")
    preds = m(image_rgb=img.cuda(), camera=cams, fg_probability=fg.cuda(), mask_crop=t(gold["fwd/mask_crop"]).cuda(),
              sequence_name=[str(s) for s in gold["fwd/names"]])
    e = rel_err(preds["voxel_features"], t(gold["fwd/grid"]))
    print(f"encoder branch of forward vs the reference's: {e:.2e}")
    assert e < 1e-4
    assert preds["images_render"].shape == (1, 3, 8, 8) and torch.isfinite(preds["images_render"]).all()


def test_full_size_grid_with_resnet_features():
    """cfg-sized encoder: 64^3 grid x 32 channels from 10 source views of 256^2 (base.yaml: ResNet34 stages 1-4 at 16
    channels + mask + image = 68 feature columns), MLPMean aggregator.  Parity on 4096 of the 262144 grid points
    (the CPU oracle takes the extractor's own feature maps); the device time of the pooling is printed."""
    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import encoder as en
    torch.manual_seed(0)
    n_src, R, C = 10, 64, 32
    ext = hd.ResNetFeatureExtractor(proj_dim=16, image_rescale=1.0).cuda().eval()
    cams = ro.simple_360_cameras(n_src, focal_length=3.2)
    up = lambda t: torch.nn.functional.interpolate(t, size=(256, 256), mode="bilinear")   # noqa: E731  smooth, image-like
    imgs, fg = up(torch.rand(n_src, 3, 32, 32)).cuda(), up(torch.rand(n_src, 1, 32, 32)).cuda()
    with torch.no_grad():
        feats = ext(imgs, fg)
    assert list(feats) == ["res_layer_1", "res_layer_2", "res_layer_3", "res_layer_4", "mask", "image"]
    assert [f.shape[1] for f in feats.values()] == [16, 16, 16, 16, 1, 3] and feats["res_layer_1"].shape[-1] == 64
    sd = eo.make_aggregator_params(68 + 21, seed=31)
    pooler = _mlp_pooler(sd, 128, 128, 1, False)
    mapper = en.LazyLinearWithXavierInit(C).cuda()
    pts = en.coord_grid(R, 8.0, "cuda")
    assert torch.allclose(pts.cpu(), eo.coord_grid(R, 8.0)[0])
    g = _cams_gpu(cams)
    rows = en.pool_views(pooler, pts, g, feats, None, None, mapper=mapper)   # warm-up: packs the weights
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        rows = en.pool_views(pooler, pts, g, feats, None, None, mapper=mapper)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    idx = torch.randperm(R ** 3, generator=torch.Generator().manual_seed(1))[:4096]
    cpu_feats = {k: v.float().cpu() for k, v in feats.items()}
    ref = eo.encode(cams, cpu_feats, None, sd, mapper.weight.detach().cpu(), mapper.bias.detach().cpu(), 0, 0.0,
                    pts=pts.cpu()[idx])[0, :, 0].t()
    e = rel_err(torch.tanh(rows[idx.cuda()]), ref)
    rows_total = n_src * R ** 3
    flops = rows_total * 2.0 * (128 * 128 + 128 * 128) + R ** 3 * 2.0 * (128 * 128 + 128 * C)
    print(f"full-size view pooling (64^3 points x {n_src} views): {ms:.2f} ms, {flops / ms / 1e9:.1f} TFLOP/s algorithmic, "
          f"rel err {e:.2e}")
    assert e < 1e-4


@pytest.mark.parametrize("N,act,rows_mod", [(128, "leakyrelu", 256), (64, "softplus", 0), (128, "relu", 128), (64, "identity", 384)])
def test_gemm_with_activation_epilogue(N, act, rows_mod):
    """holo_gemm_tc_act: act(a b^T / s + bias + row_term[m % rows]) against fp64, fp32 output and operand-pair output."""
    from holo_diffusion_b200 import ops
    g = torch.Generator().manual_seed(N + rows_mod)
    M, K = 768, 192
    a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    term = torch.randn(rows_mod, N, generator=g) if rows_mod else None
    scale = 1024.0

    def pair(x):
        hi = x.cuda().half()
        return hi, (x.cuda() - hi.float()).half()

    a_hi, a_lo = pair(a)
    w_hi, w_lo = pair(w * scale)
    out = torch.empty(M, N, device="cuda")
    o_hi, o_lo = torch.empty(M, N, device="cuda", dtype=torch.half), torch.empty(M, N, device="cuda", dtype=torch.half)
    rc = ops.gemm_tc_act(a_hi, a_lo, K, M, K, w_hi, w_lo, K, N, bias.cuda(), None if term is None else term.cuda(), rows_mod,
                         act, N, out, o_hi, o_lo, acc_scale=1.0 / scale)
    assert rc == 0
    y = a.double() @ w.double().t() + bias.double()
    if term is not None:
        y = y + term.double()[torch.arange(M) % rows_mod]
    ref = {"identity": lambda t: t, "relu": torch.relu, "leakyrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2),
           "softplus": torch.nn.functional.softplus}[act](y)
    e1, e2 = rel_err(out, ref), rel_err(o_hi.float() + o_lo.float(), ref)
    print(f"gemm + {act} epilogue N={N} rows_mod={rows_mod}: fp32 out {e1:.2e}, pair out {e2:.2e}")
    assert e1 < 5e-6 and e2 < 5e-6


def test_fused_activation_epilogue_matches_separate_pass():
    """pool_views with the activations folded into the GEMM epilogues (HOLO_VIEWPOOL_FUSE_ACT=1) == GEMM +
    holo_viewpool_act_split (the default: faster, see encoder.py)."""
    from holo_diffusion_b200 import encoder as en
    cams, feats, mask_crop = eo.make_views(4, (40, 48), seed=41)
    sd = eo.make_aggregator_params(64 + 1 + 3 + 21, 128, 128, 3, seed=42)
    pts = _points(700, 43).cuda()
    outs = []
    for flag in ("1", "0"):
        os.environ["HOLO_VIEWPOOL_FUSE_ACT"], os.environ["HOLO_VIEWPOOL_CHUNK"] = flag, "384"
        try:
            pooler = _mlp_pooler(sd, 128, 128, 3, True)
            rows = en.pool_views(pooler, pts, _cams_gpu(cams), _gpu(feats), mask_crop.cuda(), None, mapper=None)
            outs.append(rows.clone())
        finally:
            os.environ.pop("HOLO_VIEWPOOL_FUSE_ACT", None), os.environ.pop("HOLO_VIEWPOOL_CHUNK", None)
    e = rel_err(outs[0], outs[1])
    fs, ms = eo.sample_views(cams, pts.cpu(), feats, mask_crop, True)
    ref = eo.mlp_mean_aggregate(sd, fs, ms, cams, pts.cpu())[0, 0]
    print(f"fused vs separate activation pass: {e:.2e}; fused vs oracle {rel_err(outs[0], ref):.2e}")
    assert e < 2e-6 and rel_err(outs[0], ref) < 1e-4
