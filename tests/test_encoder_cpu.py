"""View-pooling encoder on a CPU box: the oracle against the reference's in-tree code (tests/golden/encoder_intree_ref.npz,
made by tests/golden/make_encoder_intree_golden.py), independent anchors for the pytorch3d leaves the oracle restates
from memory, and the HOST logic of holo_diffusion_b200/encoder.py (weight folding / packing, chunking, padding, the
model's encoder branch) over stand-ins of the kernels (tests/fake_encoder_ops.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import encoder_oracle as eo
from oracle import render_oracle as ro

GOLD = os.path.join(os.path.dirname(__file__), "golden", "encoder_intree_ref.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _cams(g, prefix):
    return ro.OracleCameras(_t(g[prefix + "R"]), _t(g[prefix + "T"]), _t(g[prefix + "focal"]), _t(g[prefix + "pp"]))


def _sd(g, tag, keys):
    return {k: _t(g[f"{tag}{k}"]) for k in keys}


# ------------------------------------------------------------------------------------------------ oracle vs reference
@pytest.mark.parametrize("n_layers", [1, 2])
def test_oracle_matches_reference_aggregator(gold, n_layers):
    """MLPMeanFeatureAggregator.forward of the reference (executed over the stub) == oracle restatement, bit for bit
    up to fp32 summation order."""
    cams, pts = _cams(gold, "agg/cam_"), _t(gold["agg/pts"])
    fs = {str(k): _t(gold["agg/feats/" + str(k)]) for k in gold["agg/feat_keys"]}
    sd = _sd(gold, f"agg{n_layers}/sd/", [str(k) for k in gold[f"agg{n_layers}/sd_keys"]])
    y = eo.mlp_mean_aggregate(sd, fs, _t(gold["agg/masks"]), cams, pts)
    ref = _t(gold[f"agg{n_layers}/out"])
    assert y.shape == ref.shape == (1, 1, 24, 8)
    assert (y - ref).abs().max() <= 1e-6 * ref.abs().max()
    # the activation placement matters: swapping the two activations is far outside that bound
    if n_layers == 2:
        wrong = eo.mlp_mean_aggregate(sd, fs, _t(gold["agg/masks"]), cams, pts, hidden_activation="softplus",
                                      last_activation="leakyrelu")
        assert (wrong - ref).abs().max() > 1e-3 * ref.abs().max()


def test_oracle_ray_dirs_match_reference(gold):
    d = eo.point_to_camera_ray_dirs(_cams(gold, "agg/cam_"), _t(gold["agg/pts"]))
    assert torch.allclose(d, _t(gold["agg/ray_dirs"]), atol=1e-7)


def _forward_inputs(gold):
    sel = [1, 3, 4]   # the log of the reference run: views of sequence "seq_a" after the target view 0
    log = [str(x) for x in gold["fwd/log"]]
    assert "('extractor_views', (1, 3, 4))" in log and any(s.startswith("('pooler', (1, 64, 3), ['seq_a'], ['seq_a', 'seq_a', 'seq_a'], 3")
                                                           for s in log)
    assert "('locator', 1, (4, 4, 4), 2.0)" in log        # voxel_size = volume_extent / resol
    img, fg = _t(gold["fwd/image_rgb"]), _t(gold["fwd/fg"])
    painted = (fg > 0.5).float() * img                      # preprocess_input, bg_color = 0
    feats = {k[len("fwd/feats/"):]: _t(gold[k])[sel] for k in gold.files if k.startswith("fwd/feats/")}
    feats["mask"], feats["image"] = fg[sel], painted[sel]
    agg = _sd(gold, "fwd/agg/", [str(k) for k in gold["fwd/agg_keys"]])
    return sel, feats, agg


def test_oracle_encode_matches_reference_forward(gold):
    """The encoder branch of the reference's forward (source executed on a stand-in self) == oracle encode()."""
    assert bool(gold["fwd/exclusion_assert"])               # forward refuses an aggregator that excludes the target view
    sel, feats, agg = _forward_inputs(gold)
    cams = _cams(gold, "fwd/cam_")[sel]
    grid = eo.encode(cams, feats, _t(gold["fwd/mask_crop"])[sel], agg, _t(gold["fwd/mapper_w"]), _t(gold["fwd/mapper_b"]), 4, 8.0)
    ref = _t(gold["fwd/grid"])
    assert grid.shape == ref.shape == (1, 8, 4, 4, 4)
    assert (grid - ref).abs().max() <= 2e-6


# ------------------------------------------------------------------------------------------------ leaf anchors
def test_leaf_projection_inverts_unprojection():
    """project_ndc against the renderer oracle's unproject (a separate restatement, anchored on the ray geometry tests)."""
    cams = ro.simple_360_cameras(6)
    g = torch.Generator().manual_seed(0)
    xy = torch.rand(6, 50, 2, generator=g) * 2 - 1
    depth = torch.rand(6, 50, 1, generator=g) * 8 + 4
    world = cams.unproject(torch.cat([xy, depth], -1))
    for i in range(6):
        back = eo.project_ndc(cams[i], world[i])[0]
        assert torch.allclose(back, xy[i], atol=2e-5)


def test_leaf_projection_clamps_the_divide():
    cams = ro.simple_360_cameras(1)
    centre = cams.centre()[0]
    p = centre[None] + torch.tensor([[0.3, 0.2, 0.0]]) @ cams.R[0].t()     # in the camera plane: z_cam = 0
    xy = eo.project_ndc(cams, p, eps=1e-2)
    assert torch.isfinite(xy).all() and xy.abs().max() > 10               # divided by +eps, not by zero


def test_leaf_ndc_grid_sample_linear_image():
    """On an image that is linear in the pixel index, bilinear sampling returns the linear function of the NDC
    coordinates: +x left / +y up, the shorter side spans [-1, 1], align_corners=False."""
    H, W = 6, 12
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.stack([xs, ys])[None]                                      # (1, 2, H, W)
    s = W / H
    xy = torch.tensor([[[0.0, 0.0], [1.2, 0.5], [-1.5, -0.7], [0.3, 0.8]]])
    got = eo.ndc_grid_sample(img, xy)[0].t()                               # (P, 2)
    px = (-xy[0, :, 0] / s + 1) * W / 2 - 0.5                               # pixel centre convention
    py = (-xy[0, :, 1] + 1) * H / 2 - 0.5
    assert torch.allclose(got[:, 0], px, atol=1e-5) and torch.allclose(got[:, 1], py, atol=1e-5)
    out = eo.ndc_grid_sample(img + 1, torch.tensor([[[2.5, 0.0], [0.0, 1.5]]]))   # outside: zeros padding
    assert out.abs().max() == 0


def test_leaf_coord_grid():
    g = eo.coord_grid(4, 8.0)[0].reshape(4, 4, 4, 3)
    assert torch.allclose(g[0, 0, :, 0], torch.tensor([-3.0, -1.0, 1.0, 3.0]))   # voxel centres, x fastest
    assert torch.allclose(g[:, 0, 0, 2], torch.tensor([-3.0, -1.0, 1.0, 3.0])) and torch.allclose(g[0, :, 0, 1], g[0, 0, :, 0])
    assert g.sum().abs() < 1e-5


def test_leaf_angle_weighted_statistics():
    """With all weights equal the angle-weighted mean / std are the plain ones."""
    cams, feats, _ = eo.make_views(3, (16, 16), stage_channels=(4,), seed=1)
    same = ro.OracleCameras(cams.R[:1].expand(3, 3, 3), cams.T[:1].expand(3, 3), cams.focal, cams.pp)
    pts = torch.zeros(5, 3)
    fs = {"a": torch.randn(1, 3, 5, 4, generator=torch.Generator().manual_seed(2))}
    out = eo.angle_weighted_aggregate(fs, torch.ones(1, 3, 5, 1), same, pts)
    assert torch.allclose(out[..., :4], fs["a"].mean(1, keepdim=True), atol=1e-6)
    var = ((fs["a"] - fs["a"].mean(1, keepdim=True)) ** 2).mean(1, keepdim=True)
    assert torch.allclose(out[..., 4:], var.clamp(1e-4).sqrt(), atol=1e-6)


# ------------------------------------------------------------------------------------------------ host logic
@pytest.fixture()
def enc(monkeypatch):
    import holo_diffusion_b200  # noqa: F401  (loads the library; no kernel is launched on this box)
    from holo_diffusion_b200 import encoder, ops

    import fake_encoder_ops
    import fake_model_ops
    fake_model_ops.install(ops, monkeypatch.setattr)
    fake_encoder_ops.install(ops, monkeypatch.setattr)
    monkeypatch.setattr(encoder, "PAIR_DTYPE", torch.float32)   # the stand-in GEMM keeps everything in the hi half
    return encoder


def _b200_cams(c):
    import holo_diffusion_b200 as hd
    return hd.PerspectiveCameras(c.focal, c.pp, c.R, c.T)


@pytest.mark.parametrize("n_layers,chunk", [(1, None), (3, "128")])
def test_host_pool_views_mlp_mean(enc, monkeypatch, n_layers, chunk):
    """pool_views: folded first layers, padded packing, activation placement, chunked points, mapper -- against the
    oracle, through the stand-in kernels; 300 points in chunks of 128 leave a ragged last chunk."""
    if chunk:
        monkeypatch.setenv("HOLO_VIEWPOOL_CHUNK", chunk)
    cams, feats, mask_crop = eo.make_views(4, (24, 40), stage_channels=(8, 8), seed=5)
    Kx = 8 + 8 + 1 + 3 + 21
    sd = eo.make_aggregator_params(Kx, n_hidden=32, dim_out=24, n_layers=n_layers, seed=4)
    pooler = enc.ViewPooler(view_sampler_args=dict(masked_sampling=True), feature_aggregator_class_type="MLPMeanFeatureAggregator",
                            feature_aggregator_MLPMeanFeatureAggregator_args=dict(n_hidden=32, dim_out=24, n_layers=n_layers))
    pooler.feature_aggregator.load_state_dict(sd)
    pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
    mapper = enc.LazyLinearWithXavierInit(16)
    pts = (torch.rand(300, 3, generator=torch.Generator().manual_seed(6)) * 2 - 1) * 4
    vw = torch.tensor([1.0, 1.0, 0.0, 1.0])
    rows = enc.pool_views(pooler, pts, _b200_cams(cams), feats, mask_crop, vw, mapper=mapper)
    assert rows.shape == (300, 16)
    ref = eo.encode(cams, feats, mask_crop, sd, mapper.weight.detach(), mapper.bias.detach(), 0, 0.0, masked_sampling=True,
                    view_weight=vw, pts=pts)
    assert (torch.tanh(rows).t() - ref[0, :, 0]).abs().max() < 2e-5
    # the view pooler's own output (no mapper)
    pooled = pooler(pts=pts[None], seq_id_pts=["a"], camera=_b200_cams(cams), seq_id_camera=["a", "a", "b", "a"], feats=feats,
                    masks=mask_crop)
    fs, ms = eo.sample_views(cams, pts, feats, mask_crop, True, view_weight=vw)
    assert pooled.shape == (1, 1, 300, 24)
    assert (pooled - eo.mlp_mean_aggregate(sd, fs, ms, cams, pts)).abs().max() < 2e-5
    # a changed parameter re-packs
    with torch.no_grad():
        pooler.feature_aggregator._last.bias.add_(1.0)
    again = pooler(pts=pts[None], seq_id_pts=None, camera=_b200_cams(cams), seq_id_camera=None, feats=feats, masks=mask_crop)
    assert (again - pooled).abs().max() > 0.1


def test_host_pool_views_angle_weighted(enc):
    cams, feats, mask_crop = eo.make_views(5, (24, 24), stage_channels=(8,), seed=8)
    pooler = enc.ViewPooler()                                              # pytorch3d's default aggregator
    pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
    mapper = enc.LazyLinearWithXavierInit(16)
    pts = (torch.rand(200, 3, generator=torch.Generator().manual_seed(9)) * 2 - 1) * 4
    rows = enc.pool_views(pooler, pts, _b200_cams(cams), feats, mask_crop, None, mapper=mapper)
    ref = eo.encode(cams, feats, mask_crop, None, mapper.weight.detach(), mapper.bias.detach(), 0, 0.0, pts=pts)
    assert mapper.weight.shape == (16, 2 * (8 + 1 + 3))
    assert (torch.tanh(rows).t() - ref[0, :, 0]).abs().max() < 2e-5


def test_host_model_encoder_branch_matches_reference_forward(enc, gold):
    """HoloDiffusionModel.forward(image_rgb=...) -- source selection, background painting, extractor / pooler
    arguments, mapper, tanh, layout -- reproduces the grid the reference's forward binds (golden)."""
    import holo_diffusion_b200 as hd
    sel, feats, agg = _forward_inputs(gold)
    m = hd.HoloDiffusionModel(
        resol=4, feature_size=8, num_passes=1, render_image_width=4, render_image_height=4, net_3d_enabled=False,
        diffusion_enabled=False, use_cuda_graph=False, view_pooler_enabled=True,
        view_pooler_args=dict(feature_aggregator_class_type="MLPMeanFeatureAggregator",
                              feature_aggregator_MLPMeanFeatureAggregator_args=dict(n_hidden=16, dim_out=8)),
        raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=4),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(n_pts_per_ray_fine_evaluation=2))
    assert not m.view_pooler.feature_aggregator.exclude_target_view
    m.view_pooler.feature_aggregator.load_state_dict(agg)
    seen = {}

    class Extractor(torch.nn.Module):
        def forward(self, imgs, masks):
            seen["imgs"], seen["masks"] = imgs, masks
            return feats

    m.image_feature_extractor = Extractor()
    m.pooled_feature_mapper.load_state_dict({"weight": _t(gold["fwd/mapper_w"]), "bias": _t(gold["fwd/mapper_b"])})
    cams = _b200_cams(_cams(gold, "fwd/cam_"))
    kw = dict(image_rgb=_t(gold["fwd/image_rgb"]), camera=cams, fg_probability=_t(gold["fwd/fg"]),
              mask_crop=_t(gold["fwd/mask_crop"]), sequence_name=[str(s) for s in gold["fwd/names"]])
    grid = m.encode_views(**kw)
    assert torch.equal(seen["imgs"], feats["image"]) and torch.equal(seen["masks"], feats["mask"])
    assert (grid - _t(gold["fwd/grid"])).abs().max() < 2e-5
    preds = m(**kw)                                                        # and through forward (render on stand-ins)
    assert (preds["voxel_features"] - _t(gold["fwd/grid"])).abs().max() < 2e-5
    assert preds["images_render"].shape == (1, 3, 4, 4)
    with pytest.raises(AssertionError, match="Cannot provide both"):
        m(voxel_features=grid, **kw)
    m.view_pooler_enabled = False
    with pytest.raises(AssertionError, match="view_pooler must be enabled"):
        m(**kw)


def test_host_source_selection():
    from holo_diffusion_b200.encoder import select_sources, view_weights
    assert select_sources(["a", "a", "b", "a"], 4, 1) == [1, 3]
    assert select_sources(["a", "b", "b"], 3, 1) == [0, 1, 2]              # nothing left: everything (:301-303)
    assert select_sources(None, 3, 1) == [1, 2]
    assert view_weights(["a"], ["a", "a"], "cpu") is None
    assert view_weights(["a"], ["a", "b"], "cpu").tolist() == [1.0, 0.0]


def test_host_resnet_feature_extractor_contract():
    """ResNetFeatureExtractor (library network, torch ops): the feature dict the pooling consumes -- names and order,
    stage resolutions (1/4 ... 1/32 of the rescaled image), proj_dim channels L2-normalised to 1 / sqrt(n_stages),
    the un-touched masks, the normalised + rescaled image, and the reference's parameter names."""
    import holo_diffusion_b200  # noqa: F401
    from holo_diffusion_b200.encoder import ResNetFeatureExtractor
    torch.manual_seed(0)
    ext = ResNetFeatureExtractor(proj_dim=16, image_rescale=0.32).eval()    # configs/base.yaml:162-164
    imgs, fg = torch.rand(2, 3, 400, 400), torch.rand(2, 1, 400, 400)
    feats = ext(imgs, fg)
    assert list(feats) == ["res_layer_1", "res_layer_2", "res_layer_3", "res_layer_4", "mask", "image"]
    assert [tuple(f.shape[1:]) for f in feats.values()] == [(16, 32, 32), (16, 16, 16), (16, 8, 8), (16, 4, 4), (1, 400, 400),
                                                             (3, 128, 128)]
    for k in ("res_layer_1", "res_layer_4"):
        assert torch.allclose(feats[k].norm(dim=1), torch.full_like(feats[k][:, 0], 0.5), atol=1e-5)
    assert feats["mask"] is fg
    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    assert torch.allclose(feats["image"], F.interpolate((imgs - mean) / std, scale_factor=0.32, mode="bilinear"), atol=1e-6)
    keys = list(ext.state_dict())
    assert "stem.0.weight" in keys and "layers.3.2.bn2.running_var" in keys and "proj_layers.0.weight" in keys
    assert ext.get_feat_dims() == 4 * 16 + 1 + 3
    # stages can be dropped, the projection switched off
    few = ResNetFeatureExtractor(stages=(1, 3), proj_dim=0, add_images=False, image_rescale=1.0).eval()
    f2 = few(torch.rand(1, 3, 64, 64), torch.rand(1, 1, 64, 64))
    assert list(f2) == ["res_layer_1", "res_layer_3", "mask"] and f2["res_layer_3"].shape == (1, 256, 4, 4)


def test_host_pool_views_rejects_what_is_not_built(enc):
    cams, feats, mask_crop = eo.make_views(2, (16, 16), stage_channels=(8,), seed=1)
    pts = torch.zeros(10, 3)
    pooler = enc.ViewPooler()
    with pytest.raises(NotImplementedError, match="target-view exclusion"):      # pytorch3d's default flags (:115-116 clear them)
        enc.pool_views(pooler, pts, _b200_cams(cams), feats, mask_crop, None, mapper=None)
    pooler.feature_aggregator.exclude_target_view = pooler.feature_aggregator.exclude_target_view_mask_features = False
    pooler.view_sampler.sampling_mode = "nearest"
    with pytest.raises(NotImplementedError, match="bilinear"):
        enc.pool_views(pooler, pts, _b200_cams(cams), feats, mask_crop, None, mapper=None)
    with pytest.raises(NotImplementedError, match="not built"):
        enc.ViewPooler(feature_aggregator_class_type="ReductionFeatureAggregator")
    wide = {f"m{i}": torch.randn(2, 64, 4, 4) for i in range(5)}                  # 320 columns > the kernels' 256
    pooler.view_sampler.sampling_mode = "bilinear"
    with pytest.raises(NotImplementedError, match="rows of up to 256"):
        enc.pool_views(pooler, pts, _b200_cams(cams), wide, None, None, mapper=None)
