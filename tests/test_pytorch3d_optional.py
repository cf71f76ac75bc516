"""OPTIONAL tier: the pytorch3d LEAVES of the renderer and encoder oracles against the real pytorch3d (0.7.x), wherever
it is installed.
pytorch3d is not installable in the build container or on the GPU box (no network), so every test here SKIPS there and
this file has never been executed -- it is the hook that turns "parity unpinned" into "pinned" on a machine that has
the dependency (conda env of the reference, environment.yaml:139).  CPU only."""
import math
import os

import pytest
import torch

from oracle import render_oracle as ro

pytorch3d = pytest.importorskip("pytorch3d")
if "pt3d_stub" in os.path.abspath(getattr(pytorch3d, "__file__", "") or ""):
    pytest.skip("the test-side stand-in is on sys.path, not the real pytorch3d", allow_module_level=True)


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


def test_harmonic_embedding():
    from pytorch3d.renderer import HarmonicEmbedding
    x = torch.randn(11, 3, generator=_g())
    for n in (0, 4, 6):
        assert torch.allclose(HarmonicEmbedding(n_harmonic_functions=n)(x), ro.harmonic_embedding(x, n), atol=1e-6)


def test_so3_exp_map_and_look_at():
    from pytorch3d.renderer import look_at_view_transform
    from pytorch3d.transforms import so3_exp_map
    v = torch.randn(9, 3, generator=_g(1))
    assert torch.allclose(so3_exp_map(v), ro.so3_exp_map(v), atol=1e-6)
    az = torch.linspace(0, 360, 9)[:8]
    R, T = look_at_view_transform(dist=torch.full((8,), 10.0), elev=torch.full((8,), -30.0), azim=az, up=((0.0, -1.0, 0.0),))
    Ro, To = ro.look_at_rotation_translation(torch.full((8,), 10.0), torch.full((8,), -30.0), az, up=(0.0, -1.0, 0.0))
    assert torch.allclose(R, Ro, atol=1e-6) and torch.allclose(T, To, atol=1e-5)


def test_cameras_unproject_and_ray_bundle():
    from pytorch3d.renderer import PerspectiveCameras
    oc = ro.simple_360_cameras(4)
    cams = PerspectiveCameras(focal_length=oc.focal, principal_point=oc.pp, R=oc.R, T=oc.T)
    assert torch.allclose(cams.get_camera_center(), oc.centre(), atol=1e-5)
    xyd = torch.cat([torch.rand(4, 7, 2, generator=_g(2)) * 2 - 1, torch.rand(4, 7, 1, generator=_g(3)) * 5 + 1], -1)
    assert torch.allclose(cams.unproject_points(xyd, world_coordinates=True), oc.unproject(xyd), atol=1e-4)
    # full-grid NDC rays of the evaluation ray sampler
    from pytorch3d.renderer import NDCMultinomialRaysampler
    H = W = 6
    S = 5
    b = ro.sample_rays(oc[1], H, W, S)
    rs = NDCMultinomialRaysampler(image_width=W, image_height=H, n_pts_per_ray=S, min_depth=float(b.lengths[0, 0, 0, 0]),
                                  max_depth=float(b.lengths[0, 0, 0, -1]))
    rb = rs(cams[1])
    assert torch.allclose(rb.xys, b.xys, atol=1e-6)                       # pixel order and NDC centres
    assert torch.allclose(rb.lengths, b.lengths, atol=1e-5)
    # the implicitron sampler normalises the directions; origins + t * dirs must describe the same rays
    d = torch.nn.functional.normalize(rb.directions, dim=-1)
    assert torch.allclose(d, b.directions, atol=1e-5)


def test_voxel_grid_sampling():
    from pytorch3d.implicitron.models.implicit_function.voxel_grid import FullResolutionVoxelGrid, FullResolutionVoxelGridValues
    from pytorch3d.implicitron.tools.config import expand_args_fields
    from pytorch3d.structures.volumes import VolumeLocator
    expand_args_fields(FullResolutionVoxelGrid)
    R, C, ext = 8, 5, 8.0
    grid = torch.randn(1, C, R, R, R, generator=_g(4))
    pts = (torch.rand(1, 50, 3, generator=_g(5)) - 0.5) * 9.0             # some outside the volume
    loc = VolumeLocator(batch_size=1, grid_sizes=(R, R, R), device=pts.device, voxel_size=ext / R)
    got = FullResolutionVoxelGrid(n_features=C).evaluate_world(pts, FullResolutionVoxelGridValues(grid), loc)
    ref = ro.sample_grid(grid, ro.world_to_local(pts.reshape(-1, 3), R, ext))
    assert torch.allclose(got.reshape(-1, C), ref, atol=1e-5)


def test_emission_absorption_raymarcher():
    from pytorch3d.implicitron.models.renderer.raymarcher import EmissionAbsorptionRaymarcher
    from pytorch3d.implicitron.tools.config import expand_args_fields
    expand_args_fields(EmissionAbsorptionRaymarcher)
    rm = EmissionAbsorptionRaymarcher(bg_color=(1.0, 1.0, 1.0))
    n, S = 13, 9
    dens = torch.randn(1, n, S, 1, generator=_g(6)) * 2
    feats = torch.rand(1, n, S, 3, generator=_g(7))
    z = torch.sort(torch.rand(1, n, S, generator=_g(8)) * 8 + 6, -1)[0]
    out = rm(rays_densities=dens, rays_features=feats, aux={}, ray_lengths=z)
    o = ro.ea_raymarch(dens, feats, z, bg=(1.0, 1.0, 1.0))
    for a, b in ((out.features, o.features), (out.depths, o.depths), (out.masks, o.masks), (out.weights, o.weights)):
        assert torch.allclose(a, b, atol=1e-6)


def test_ray_point_refiner():
    from pytorch3d.implicitron.models.renderer.base import ImplicitronRayBundle
    from pytorch3d.implicitron.models.renderer.ray_point_refiner import RayPointRefiner
    from pytorch3d.implicitron.tools.config import expand_args_fields
    expand_args_fields(RayPointRefiner)
    n, S, N = 10, 12, 16
    z = torch.sort(torch.rand(1, n, S, generator=_g(9)) * 8 + 6, -1)[0]
    w = torch.rand(1, n, S, generator=_g(10))
    o = torch.zeros(1, n, 3)
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=_g(11)), dim=-1)
    bundle = ImplicitronRayBundle(origins=o, directions=d, lengths=z, xys=torch.zeros(1, n, 2))
    fine = RayPointRefiner(n_pts_per_ray=N, random_sampling=False, add_input_samples=True)(bundle, w)
    assert torch.allclose(fine.lengths, ro.refine_lengths(z, w, N), atol=1e-5)
    assert math.isclose(float(fine.lengths.shape[-1]), S + N)


# ------------------------------------------------------------------------------------------------ encoder leaves
def _pt3d_views(n=4, hw=(24, 40)):
    from pytorch3d.renderer import PerspectiveCameras
    from oracle import encoder_oracle as eo
    oc, feats, mask_crop = eo.make_views(n, hw, stage_channels=(8, 8), seed=3)
    cams = PerspectiveCameras(focal_length=oc.focal, principal_point=oc.pp, R=oc.R, T=oc.T)
    return oc, cams, feats, mask_crop


def test_volume_locator_coord_grid():
    from pytorch3d.structures.volumes import VolumeLocator
    from oracle import encoder_oracle as eo
    loc = VolumeLocator(batch_size=1, grid_sizes=(6, 6, 6), device=torch.device("cpu"), voxel_size=8.0 / 6)
    assert torch.allclose(loc.get_coord_grid().reshape(1, -1, 3), eo.coord_grid(6, 8.0), atol=1e-6)


def test_view_sampler_projection_and_grid_sample():
    """ViewSampler.forward (project_points_and_sample + ndc_grid_sample + the sequence mask), bilinear features and
    nearest masks, non-square feature maps, points behind the cameras."""
    from pytorch3d.implicitron.models.view_pooler.view_sampler import ViewSampler
    from pytorch3d.implicitron.tools.config import expand_args_fields
    from oracle import encoder_oracle as eo
    expand_args_fields(ViewSampler)
    oc, cams, feats, mask_crop = _pt3d_views()
    pts = (torch.rand(1, 60, 3, generator=_g(11)) * 2 - 1) * 6.0
    pts[0, 0] = oc.centre()[0] + 1e-3                                      # (almost) at a camera centre: the eps clamp
    for masked in (False, True):
        vs = ViewSampler(masked_sampling=masked, sampling_mode="bilinear")
        f_ref, m_ref = vs(pts=pts, seq_id_pts=["a"], camera=cams, seq_id_camera=["a", "a", "b", "a"], feats=feats,
                          masks=mask_crop)
        f_o, m_o = eo.sample_views(oc, pts[0], feats, mask_crop, masked, view_weight=torch.tensor([1.0, 1.0, 0.0, 1.0]))
        assert torch.allclose(m_ref, m_o, atol=1e-6)
        for k in feats:
            assert torch.allclose(f_ref[k], f_o[k], atol=1e-5), k


def test_angle_weighted_aggregator():
    from pytorch3d.implicitron.models.view_pooler.feature_aggregator import AngleWeightedReductionFeatureAggregator
    from pytorch3d.implicitron.tools.config import expand_args_fields
    from oracle import encoder_oracle as eo
    expand_args_fields(AngleWeightedReductionFeatureAggregator)
    oc, cams, feats, mask_crop = _pt3d_views()
    pts = (torch.rand(70, 3, generator=_g(12)) * 2 - 1) * 4.0
    fs, ms = eo.sample_views(oc, pts, feats, mask_crop, True)
    agg = AngleWeightedReductionFeatureAggregator(exclude_target_view=False, exclude_target_view_mask_features=False)
    ref = agg(fs, ms, camera=cams, pts=pts[None])
    assert torch.allclose(ref, eo.angle_weighted_aggregate(fs, ms, oc, pts), atol=1e-5)


def test_resnet_feature_extractor():
    """Our restatement of pytorch3d's ResNetFeatureExtractor with the library's own weights loaded into it."""
    from pytorch3d.implicitron.models.feature_extractor.resnet_feature_extractor import ResNetFeatureExtractor as Ref
    from pytorch3d.implicitron.tools.config import expand_args_fields
    import holo_diffusion_b200  # noqa: F401  (needs the built library: skips nothing, fails loudly without it)
    from holo_diffusion_b200.encoder import ResNetFeatureExtractor
    expand_args_fields(Ref)
    ref = Ref(pretrained=False, proj_dim=16, image_rescale=0.32).eval()
    ours = ResNetFeatureExtractor(proj_dim=16, image_rescale=0.32).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    imgs, fg = torch.rand(2, 3, 200, 200, generator=_g(13)), torch.rand(2, 1, 200, 200, generator=_g(14))
    with torch.no_grad():
        a, b = ref(imgs, fg), ours(imgs, fg)
    assert list(a) == list(b)
    for k in a:
        assert torch.allclose(a[k], b[k], atol=1e-5), k


def test_mlp_mean_aggregator_over_the_real_leaves():
    """The reference's in-tree aggregator on the REAL pytorch3d helpers (wmean, the cartesian product, the harmonic
    embedding) against the oracle: what tests/golden/make_encoder_intree_golden.py does over the stand-in."""
    import sys
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("the reference checkout is not here")
    sys.path.insert(0, ref_root)
    sys.modules.pop("holo_diffusion", None)
    try:
        from holo_diffusion.custom_modules import MLPMeanFeatureAggregator
    finally:
        sys.path.remove(ref_root)
    from oracle import encoder_oracle as eo
    oc, cams, feats, mask_crop = _pt3d_views()
    pts = (torch.rand(40, 3, generator=_g(15)) * 2 - 1) * 4.0
    fs, ms = eo.sample_views(oc, pts, feats, mask_crop, True)
    torch.manual_seed(3)
    agg = MLPMeanFeatureAggregator(n_hidden=32, dim_out=16, n_layers=2, checkpointed_mlp=False)
    agg.exclude_target_view = agg.exclude_target_view_mask_features = False
    with torch.no_grad():
        ref = agg(fs, ms, camera=cams, pts=pts[None])
    sd = {k: v.detach() for k, v in agg.state_dict().items()}
    assert torch.allclose(ref, eo.mlp_mean_aggregate(sd, fs, ms, oc, pts), atol=1e-5)
