"""GPU parity of the per-stage renderer plug-ins (implicit function, RenderMLP, ray marcher, refiner, normals,
chunked rendering) against the CPU oracle, through the C-ABI.  The first tests read like the reference's own
(/root/reference/holo_diffusion/tests/test_voxel_grid_implicit_function.py) with a parity check added."""
import pytest
import torch

from conftest import rel_err
from fixtures import make_grid, make_mlp
from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _load_mlp(render_mlp, p):
    render_mlp.load_state_dict(p, strict=True)
    return render_mlp.cuda()


def _head_params(C_hidden, F, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"_feature_net.mlp.0.0.weight": (torch.rand(F, C_hidden, generator=g) * 2 - 1) * 0.15,
            "_feature_net.mlp.0.0.bias": (torch.rand(F, generator=g) * 2 - 1) * 0.1}


def test_RenderMLP_forward():
    """reference test_RenderMLP_forward: default RenderMLP (128 -> 1 + 3 + 64), 16 random features / view dirs."""
    import holo_diffusion_b200 as hd
    p = make_mlp(128)
    p.update(_head_params(256, 64))
    render_mlp = _load_mlp(hd.RenderMLP(), p)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(16, render_mlp.input_dims, generator=g)
    dirs = torch.randn(16, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    densities, rad_feats, vp_idp_feats = render_mlp(feats.cuda(), dirs.cuda())
    for t in (densities, rad_feats, vp_idp_feats):
        assert torch.isnan(t).sum().item() == 0
    d, rgb, head = ro.render_mlp(p, feats, dirs, return_head=True)
    assert densities.shape == (16, 1) and rad_feats.shape == (16, 3) and vp_idp_feats.shape == (16, 64)
    assert rel_err(densities, d) < TOL and rel_err(rad_feats, rgb) < TOL and rel_err(vp_idp_feats, head) < TOL


@pytest.mark.parametrize("C,R,F,dims", [(128, 32, 64, (2, 12, 12, 16)), (16, 8, 0, (1, 7, 5, 9)), (32, 16, 0, (3, 40))])
def test_VoxelGridImplicitFunction_forward_pts3d(C, R, F, dims):
    """reference test_VoxelGridImplicitFunction_forward: explicit pts_3d inside the volume, dummy directions."""
    import holo_diffusion_b200 as hd
    p = make_mlp(C)
    if F:
        p.update(_head_params(256, F))
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=F)
    _load_mlp(fn.render_mlp, p)
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(*dims, 3, generator=g) * 2.0 - 1.0) * (fn.volume_extent / 2.0) * 1.1  # some fall outside
    grid = make_grid(C, R)
    densities, features, _ = fn(pts_3d=pts.cuda(), voxel_grid_features=grid.cuda())
    for t in (densities, features):
        assert torch.isnan(t).sum().item() == 0
    d, f = ro.implicit_function(p, grid, None, R, 8.0, pts_3d=pts)
    assert densities.shape == d.shape and features.shape == f.shape
    assert rel_err(densities, d) < TOL and rel_err(features, f) < TOL


@pytest.mark.parametrize("C,R,HW,S", [(16, 8, 12, 9), (32, 16, 20, 16), (64, 8, 8, 5)])
def test_implicit_function_bundle_and_normals(C, R, HW, S):
    import holo_diffusion_b200 as hd
    grid, p = make_grid(C, R), make_mlp(C)
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=0, render_normals=True)
    _load_mlp(fn.render_mlp, p)
    cams = ro.simple_360_cameras(8)
    b = ro.sample_rays(cams[3], HW, HW, S)
    bundle = hd.ImplicitronRayBundle(b.origins.cuda(), b.directions.cuda(), b.lengths.cuda(), b.xys.cuda())
    dens, feats, aux = fn(ray_bundle=bundle, voxel_grid_features=grid.cuda())
    d, f, n = ro.implicit_function(p, grid, b, R, 8.0, render_normals=True)
    assert rel_err(dens, d) < TOL and rel_err(feats, f) < TOL
    # normals are unit vectors of an fp32 gradient: bound the comparison by the fp32 oracle's own distance to fp64
    p64 = {k: v.double() for k, v in p.items()}
    b64 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), b.xys.double())
    n64 = ro.implicit_function(p64, grid.double(), b64, R, 8.0, render_normals=True)[2]
    own = rel_err(n, n64)
    assert aux["normals"].shape == n.shape
    assert rel_err(aux["normals"], n64) < 3 * own + TOL
    assert float((n64.norm(dim=-1) > 0.5).float().mean()) > 0.3, "fixture must have non-trivial normals"
    # points outside the grid have a zero gradient -> zero normal (F.normalize eps)
    zero = n64.norm(dim=-1) == 0
    assert bool((aux["normals"].cpu()[zero] == 0).all())


@pytest.mark.parametrize("S,Fd,with_noise,with_normals", [(16, 3, False, False), (9, 3, True, True), (33, 67, False, False)])
def test_raymarcher_matches_oracle(S, Fd, with_noise, with_normals):
    import holo_diffusion_b200 as hd
    g = torch.Generator().manual_seed(11)
    n = 300
    dens = torch.randn(1, n, S, 1, generator=g) * 2.0
    feats = torch.rand(1, n, S, Fd, generator=g)
    z = torch.sort(torch.rand(1, n, S, generator=g) * 8 + 6, -1)[0]
    noise = torch.randn(1, n, S, generator=g) if with_noise else None
    normals = torch.nn.functional.normalize(torch.randn(1, n, S, 3, generator=g), dim=-1) if with_normals else None
    bg = tuple(float(x) for x in torch.rand(Fd, generator=g))
    ref = ro.ea_raymarch(dens, feats, z, bg=bg, noise=noise)
    from holo_diffusion_b200 import ops
    o = ops.ea_raymarch(dens.reshape(n, S).cuda(), feats.reshape(n, S, Fd).cuda(), z.reshape(n, S).cuda(), bg, 1e10,
                        noise=None if noise is None else noise.reshape(n, S).cuda(),
                        normals=None if normals is None else normals.reshape(n, S, 3).cuda())
    for k in ("features", "depths", "masks", "weights"):
        assert rel_err(o[k], getattr(ref, k).reshape(n, -1)) < TOL, k
    if with_normals:
        assert rel_err(o["normals"], (normals * ref.weights[..., None]).sum(-2).reshape(n, 3)) < TOL
    # the plug-in class: same numbers through EmissionAbsorptionRaymarcher.forward (noise drawn inside -> only eval)
    if not with_noise:
        rm = hd.EmissionAbsorptionRaymarcher(bg_color=bg)
        out = rm(dens.cuda(), feats.cuda(), {} if normals is None else {"normals": normals.cuda()}, ray_lengths=z.cuda())
        assert out.features.shape == (1, n, Fd) and out.depths.shape == (1, n, 1) and out.weights.shape == (1, n, S)
        assert torch.equal(out.features.reshape(n, Fd), o["features"])


@pytest.mark.parametrize("S,n_fine,add,random", [(16, 16, True, False), (64, 16, True, True), (9, 33, False, True),
                                                 (3, 5, True, False)])
def test_refiner_matches_oracle(S, n_fine, add, random):
    from holo_diffusion_b200 import ops
    g = torch.Generator().manual_seed(13)
    n = 257
    z = torch.sort(torch.rand(n, S, generator=g) * 8 + 6, -1)[0]
    w = torch.rand(n, S, generator=g) ** 4
    w[: n // 3] = 0.0  # empty rays: uniform pdf from the 1e-5 floor
    u = torch.rand(n, n_fine, generator=g) if random else None
    ref32 = ro.refine_lengths(z, w, n_fine, add_input=add, u=u)
    ref64 = ro.refine_lengths(z.double(), w.double(), n_fine, add_input=add, u=None if u is None else u.double())
    out = ops.ray_refine(z.cuda(), w.cuda(), n_fine, add, None if u is None else u.cuda())
    assert out.shape == ref32.shape
    assert bool((out[:, 1:] >= out[:, :-1]).all())
    assert rel_err(out, ref64) < 3 * rel_err(ref32, ref64) + TOL  # cdf summation order differs from torch.cumsum


def _model(C, R, HW, S, n_passes, n_fine, **kw):
    import holo_diffusion_b200 as hd
    ra = dict(n_pts_per_ray_fine_evaluation=n_fine, n_pts_per_ray_fine_training=n_fine, return_weights=True,
              raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0)))
    ra.update(kw.pop("renderer", {}))
    m = hd.HoloDiffusionModel(resol=R, feature_size=C, num_passes=n_passes, render_image_width=HW, render_image_height=HW,
                              net_3d_enabled=False, diffusion_enabled=False,
                              raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=S),
                              renderer_HoloMultiPassEmissionAbsorptionRenderer_args=ra, use_cuda_graph=False, **kw)
    m._implicit_functions[0]._fn.render_mlp.load_state_dict(make_mlp(C), strict=True)
    return m.cuda()


def test_staged_and_chunked_rendering_match_fused():
    """a16: GenericModel._render chunk loop over the staged plug-ins == the one-launch fused kernel; the
    re-assembly (ray order, prev_stage chain) is bit-exact."""
    import holo_diffusion_b200 as hd
    C, R, HW, S, nf = 16, 8, 20, 8, 8
    grid = make_grid(C, R).cuda()
    cams = hd.get_simple_360_camera_trajectory(2 * 3.141592653589793, 8, -3.141592653589793 / 6, 10.0,
                                               ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    cam = cams[[5]].to("cuda")
    fused = _model(C, R, HW, S, 2, nf)(camera=cam, voxel_features=grid)["rendered"]
    whole = _model(C, R, HW, S, 2, nf, chunk_size_grid=0, renderer=dict(fused=False))(camera=cam, voxel_features=grid)["rendered"]
    chunked = _model(C, R, HW, S, 2, nf, chunk_size_grid=S * 24, renderer=dict(fused=False))(camera=cam, voxel_features=grid)["rendered"]
    for a, b in ((whole, chunked), (whole.prev_stage, chunked.prev_stage)):
        for k in ("features", "depths", "masks", "weights"):
            assert getattr(a, k).shape == getattr(b, k).shape
            assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(whole.aux["lengths"], chunked.aux["lengths"])
    assert whole.features.shape == (1, HW, HW, 3) and whole.prev_stage.weights.shape == (1, HW, HW, S)
    # staged vs fused: same arithmetic up to summation order in the coarse pass ...
    for k in ("features", "depths", "masks", "weights"):
        assert rel_err(getattr(whole.prev_stage, k), getattr(fused.prev_stage, k)) < TOL, k
    # ... and the fine pass differs only through the (ill-conditioned) refiner: bounded like test_render_gpu
    assert rel_err(whole.features, fused.features) < 5e-3
    with pytest.raises(ValueError):
        _model(C, R, HW, S, 2, nf, chunk_size_grid=S * 24 + 1, renderer=dict(fused=False))(camera=cam, voxel_features=grid)


@pytest.mark.parametrize("n_passes", [1, 2])
def test_model_render_normals(n_passes):
    """a12: render_normals=True (teddybear.yaml:203) -> RendererOutput.normals = sum_s w n."""
    import holo_diffusion_b200 as hd
    C, R, HW, S, nf = 16, 8, 16, 12, 6
    grid, p = make_grid(C, R), make_mlp(C)
    m = _model(C, R, HW, S, n_passes, nf, chunk_size_grid=S * 50,
               implicit_function_HoloVoxelGridImplicitFunction_args=dict(render_normals=True))
    cams = hd.get_simple_360_camera_trajectory(2 * 3.141592653589793, 8, -3.141592653589793 / 6, 10.0,
                                               ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    out = m(camera=cams[[1]].to("cuda"), voxel_features=grid.cuda())["rendered"]
    b = ro.sample_rays(ro.simple_360_cameras(8)[1], HW, HW, S)
    p64 = {k: v.double() for k, v in p.items()}
    b64 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), b.xys.double())
    ref = ro.render_multipass(p, grid, b, R, 8.0, n_passes, nf, render_normals=True)
    ref64 = ro.render_multipass(p64, grid.double(), b64, R, 8.0, n_passes, nf, render_normals=True)
    first, first32, first64 = out, ref, ref64
    while first.prev_stage is not None:
        first, first32, first64 = first.prev_stage, first32.prev_stage, first64.prev_stage
    assert first.normals.shape == (1, HW, HW, 3)
    assert float(first64.normals.abs().max()) > 0.1
    assert rel_err(first.normals, first64.normals) < 3 * rel_err(first32.normals, first64.normals) + TOL
    assert rel_err(first.features, first32.features) < TOL
    if n_passes == 2:
        assert rel_err(out.normals, ref64.normals) < 3 * rel_err(ref.normals, ref64.normals) + 5e-3
        assert rel_err(out.features, ref64.features) < 3 * rel_err(ref.features, ref64.features) + 5e-3


def test_training_mode_renderer_runs():
    """Training-mode forward of the renderer plug-in: density noise (std 1) + stratified refinement drawn with
    torch's generator -- seeded runs repeat, differ from evaluation, stay finite and sorted."""
    import holo_diffusion_b200 as hd
    C, R, HW, S, nf = 16, 8, 12, 8, 8
    m = _model(C, R, HW, S, 2, nf)
    grid = make_grid(C, R).cuda()
    w = m._implicit_functions[0]
    w.bind_args(voxel_grid_features=grid)
    cams = hd.get_simple_360_camera_trajectory(2 * 3.141592653589793, 8, -3.141592653589793 / 6, 10.0,
                                               ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    bundle = m.raysampler(cams[[0]].to("cuda"), hd.EvaluationMode.EVALUATION)
    fns = list(m._implicit_functions)
    torch.manual_seed(0)
    a = m.renderer(bundle, fns, hd.EvaluationMode.TRAINING)
    torch.manual_seed(0)
    b = m.renderer(bundle, fns, hd.EvaluationMode.TRAINING)
    e = m.renderer(bundle, fns, hd.EvaluationMode.EVALUATION)
    assert torch.equal(a.features, b.features)
    assert torch.isfinite(a.features).all() and torch.isfinite(a.depths).all()
    assert not torch.equal(a.features, e.features)
    l = a.aux["lengths"]
    assert l.shape[-1] == S + nf and bool((l[..., 1:] >= l[..., :-1]).all())
