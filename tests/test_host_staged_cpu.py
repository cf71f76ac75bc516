"""Host logic of the staged renderer plug-ins on CPU: the CUDA entry points are replaced by the oracle so that the
Python glue (shapes, plug-in chaining, chunk re-assembly, fused/staged dispatch) is exercised without a GPU.
The kernels themselves are checked in tests/test_render_stages_gpu.py."""
import pytest
import torch

from fixtures import make_grid, make_mlp
from oracle import render_oracle as ro


@pytest.fixture
def fake_ops(monkeypatch):
    from holo_diffusion_b200 import ops

    def collapse(layers, skips, rw, rb, C):
        p = {}
        for i, (w, b) in enumerate(layers):
            p[f"_density_net.mlp.{i}.0.weight"], p[f"_density_net.mlp.{i}.0.bias"] = w, b
        p["_radiance_net.mlp.0.0.weight"], p["_radiance_net.mlp.0.0.bias"] = rw, rb
        H = layers[-1][0].shape[0] - 1
        return p, H, rw.shape[1] - H, None

    def with_head(p, head):
        p = dict(p)
        if head is not None:
            p["_feature_net.mlp.0.0.weight"], p["_feature_net.mlp.0.0.bias"] = head
        return p

    def if_fwd(grid_dhwc, extent, packed, hidden, n_harm, *, origins=None, dirs=None, lengths=None, pts_3d=None, S=1,
               head=None, normals=False):
        grid = grid_dhwc.permute(3, 0, 1, 2)[None]
        R = grid.shape[-1]
        p = with_head(packed, head)
        if pts_3d is None:
            b = ro.OracleRayBundle(origins[None], dirs[None], lengths[None], None)
            o = ro.implicit_function(p, grid, b, R, extent, render_normals=normals)
        else:
            n = pts_3d.shape[0] // S
            b = None if dirs is None else ro.OracleRayBundle(None, dirs.reshape(n, 3), None, None)
            o = ro.implicit_function(p, grid, b, R, extent, render_normals=normals, pts_3d=pts_3d.reshape(n, S, 3))
        return o[0].reshape(-1), o[1].reshape(o[0].numel(), -1), (o[2].reshape(-1, 3) if normals else None)

    def render_mlp_fwd(feats, dirs, packed, hidden, n_harm, head=None):
        d, rgb, h = ro.render_mlp(with_head(packed, head), feats, dirs, return_head=True)
        return d.reshape(-1), rgb if h is None else torch.cat([rgb, h], -1)

    def ea(dens, feats, lengths, bg, bg_op=1e10, noise=None, normals=None):
        bgv = bg if len(bg) > 1 else bg * feats.shape[-1]
        r = ro.ea_raymarch(dens[..., None], feats, lengths, bg=bgv, background_opacity=bg_op, noise=noise)
        return {"features": r.features, "depths": r.depths, "masks": r.masks, "weights": r.weights,
                "normals": None if normals is None else (normals * r.weights[..., None]).sum(-2)}

    def refine(lengths, weights, n_fine, add=True, u=None):
        return ro.refine_lengths(lengths, weights, n_fine, add_input=add, u=u)

    def transpose2d(src, rows, cols, out=None):
        return src.view(rows, cols).t().contiguous().view(-1)

    for name, fn in dict(collapse_and_pack_render_mlp=collapse, if_fwd=if_fwd, render_mlp_fwd=render_mlp_fwd,
                         ea_raymarch=ea, ray_refine=refine, transpose2d=transpose2d).items():
        monkeypatch.setattr(ops, name, fn)
    return ops


def _bundle(HW, S, pose=1):
    import holo_diffusion_b200 as hd
    b = ro.sample_rays(ro.simple_360_cameras(8)[pose], HW, HW, S)
    return b, hd.ImplicitronRayBundle(b.origins, b.directions, b.lengths, b.xys)


def test_plugins_chain_like_the_reference(fake_ops):
    import holo_diffusion_b200 as hd
    C, R, HW, S, nf = 16, 8, 6, 8, 4
    grid, p = make_grid(C, R), make_mlp(C)
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=0, render_normals=True)
    fn.render_mlp.load_state_dict(p, strict=True)
    b, bundle = _bundle(HW, S)
    dens, feats, aux = fn(ray_bundle=bundle, voxel_grid_features=grid)
    assert dens.shape == (1, HW, HW, S, 1) and feats.shape == (1, HW, HW, S, 3) and aux["normals"].shape == (1, HW, HW, S, 3)
    # explicit points, dummy directions (reference test_VoxelGridImplicitFunction_forward)
    pts = torch.rand(2, 5, 4, S, 3) * 8 - 4
    d2, f2, a2 = fn(pts_3d=pts, voxel_grid_features=grid)
    assert d2.shape == (2, 5, 4, S, 1) and f2.shape == (2, 5, 4, S, 3) and a2["normals"].shape == (2, 5, 4, S, 3)
    # RenderMLP with the view-independent head (reference test_RenderMLP_forward)
    mlp = hd.RenderMLP()
    d3, r3, v3 = mlp(torch.randn(16, 128), torch.nn.functional.normalize(torch.randn(16, 3), dim=-1))
    assert d3.shape == (16, 1) and r3.shape == (16, 3) and v3.shape == (16, 64)
    # renderer: staged recursion == oracle multipass, incl. prev_stage chain and rendered normals
    w = hd.ImplicitFunctionWrapper(fn)
    w.bind_args(voxel_grid_features=grid)
    rend = hd.HoloMultiPassEmissionAbsorptionRenderer(
        n_pts_per_ray_fine_evaluation=nf, return_weights=True,
        raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0)))
    assert not rend.is_fused([w, w], hd.EvaluationMode.EVALUATION)  # normals force the staged path
    out = rend(bundle, [w, w], hd.EvaluationMode.EVALUATION)
    ref = ro.render_multipass(p, grid, b, R, 8.0, 2, nf, render_normals=True)
    for k in ("features", "depths", "masks", "weights", "normals"):
        assert torch.allclose(getattr(out, k), getattr(ref, k), atol=1e-6), k
        assert torch.allclose(getattr(out.prev_stage, k), getattr(ref.prev_stage, k), atol=1e-6), k
    assert out.prev_stage.prev_stage is None and out.weights.shape == (1, HW, HW, S + nf)
    # return_weights=False drops them from the outputs but the refiner still sees them
    rend.return_weights = False
    out2 = rend(bundle, [w, w], hd.EvaluationMode.EVALUATION)
    assert out2.weights is None and out2.prev_stage.weights is None and torch.equal(out2.features, out.features)
    # training mode: noise + stratified refinement change the result, seeded runs repeat
    torch.manual_seed(0)
    t1 = rend(bundle, [w, w], hd.EvaluationMode.TRAINING)
    torch.manual_seed(0)
    t2 = rend(bundle, [w, w], hd.EvaluationMode.TRAINING)
    assert torch.equal(t1.features, t2.features) and not torch.equal(t1.features, out.features)


def test_fused_dispatch_rules(fake_ops):
    import holo_diffusion_b200 as hd
    C, R = 16, 8
    grid = make_grid(C, R)
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=0)
    w = hd.ImplicitFunctionWrapper(fn)
    rend = hd.HoloMultiPassEmissionAbsorptionRenderer()
    E, T = hd.EvaluationMode.EVALUATION, hd.EvaluationMode.TRAINING
    assert not rend.is_fused([w], E)  # nothing bound yet
    w.bind_args(voxel_grid_features=grid)
    assert rend.is_fused([w], E) and rend.is_fused([w, w], E)
    assert not rend.is_fused([w, w], T) and not rend.is_fused([w, w, w], E)
    assert not rend.is_fused([w, hd.ImplicitFunctionWrapper(fn)], E)
    w_head = hd.ImplicitFunctionWrapper(hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=8))
    w_head.bind_args(voxel_grid_features=grid)
    assert not rend.is_fused([w_head], E)
    assert not hd.HoloMultiPassEmissionAbsorptionRenderer(fused=False).is_fused([w], E)
    assert not hd.HoloMultiPassEmissionAbsorptionRenderer(stratified_sampling_coarse_evaluation=True).is_fused([w], E)
    with pytest.raises(NotImplementedError):
        hd.EmissionAbsorptionRaymarcher(surface_thickness=2)


def test_chunked_render_reassembly(fake_ops):
    """GenericModel._render: chunk sizes, ray order and the prev_stage chain (a16, bit-exact bookkeeping)."""
    import holo_diffusion_b200 as hd
    C, R, HW, S, nf = 16, 8, 7, 8, 4
    grid, p = make_grid(C, R), make_mlp(C)
    m = hd.HoloDiffusionModel(resol=R, feature_size=C, num_passes=2, render_image_width=HW, render_image_height=HW,
                              net_3d_enabled=False, diffusion_enabled=False,
                              renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
                                  n_pts_per_ray_fine_evaluation=nf, return_weights=True, fused=False,
                                  raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))))
    m._implicit_functions[0]._fn.render_mlp.load_state_dict(p, strict=True)
    for f in m._implicit_functions:
        f.bind_args(voxel_grid_features=grid)
    b, bundle = _bundle(HW, S)
    fns = list(m._implicit_functions)
    kw = dict(implicit_functions=fns, evaluation_mode=hd.EvaluationMode.EVALUATION)
    whole = m._render(ray_bundle=bundle, chunksize=0, **kw)
    calls = []
    orig = m.renderer.forward

    def spy(ray_bundle, **k):
        calls.append(ray_bundle.lengths.shape[1])
        return orig(ray_bundle, **k)

    m.renderer.forward = spy
    chunked = m._render(ray_bundle=bundle, chunksize=S * 10, **kw)
    # n_rays 49, S 8: n_chunks = ceil(392/80) = 5, rays per chunk = ceil(49/5) = 10
    assert calls == [10, 10, 10, 10, 9]
    ref = ro.render_chunked(p, grid, b, R, 8.0, 2, nf, chunk_size_grid=S * 10)
    for a, c, r in ((whole, chunked, ref), (whole.prev_stage, chunked.prev_stage, ref.prev_stage)):
        for k in ("features", "depths", "masks", "weights"):
            assert getattr(c, k).shape == getattr(r, k).shape
            assert torch.equal(getattr(a, k), getattr(c, k)), k
            assert torch.allclose(getattr(c, k), getattr(r, k), atol=1e-6), k
    assert chunked.aux["lengths"].shape == (1, HW, HW, S + nf)
    with pytest.raises(ValueError):
        m._render(ray_bundle=bundle, chunksize=S * 10 + 3, **kw)
