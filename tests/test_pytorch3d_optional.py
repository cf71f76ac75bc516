"""OPTIONAL tier: the renderer oracle's pytorch3d LEAVES against the real pytorch3d (0.7.x), wherever it is installed.
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
