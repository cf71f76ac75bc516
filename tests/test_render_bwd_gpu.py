"""Backward of the staged renderer (renderer half): the CUDA backward kernels behind
holo_diffusion_b200/autograd.py against torch autograd through the oracle restatement (fp64 on the CPU) -- gradients of
the voxel grid, of all five RenderMLP layers, and through the emission-absorption ray marcher with density noise."""
import pytest
import torch

from conftest import rel_err
from fixtures import make_grid, make_mlp
from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu


def _rays(HW, S, pose=2):
    b = ro.sample_rays(ro.simple_360_cameras(8)[pose], HW, HW, S)
    return b


@pytest.mark.parametrize("n,S,Fd,with_noise", [(300, 24, 3, True), (64, 80, 3, False), (17, 5, 7, True)])
def test_ea_raymarch_backward_matches_autograd(n, S, Fd, with_noise):
    from holo_diffusion_b200.autograd import EARaymarch
    g = torch.Generator().manual_seed(1)
    dens = torch.randn(n, S, generator=g) * 2.0            # about half negative: the relu gate is exercised
    feats = torch.rand(n, S, Fd, generator=g)
    lengths = (6.0 + torch.rand(n, S, generator=g).cumsum(-1) * 0.1)
    noise = torch.randn(n, S, generator=g) if with_noise else None
    bg = tuple(float(x) for x in torch.rand(Fd, generator=g))
    cot = [torch.randn(n, Fd, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, S, generator=g)]
    # oracle, fp64
    d64, f64 = dens.double().requires_grad_(), feats.double().requires_grad_()
    o = ro.ea_raymarch(d64[..., None], f64, lengths.double(), bg, 1e10, None if noise is None else noise.double())
    loss = (o.features * cot[0].double()).sum() + (o.depths * cot[1].double()).sum() + (o.masks * cot[2].double()).sum() + \
        (o.weights * cot[3].double()).sum()
    loss.backward()
    # ours
    dc, fc = dens.cuda().requires_grad_(), feats.cuda().requires_grad_()
    f, d, m, w = EARaymarch.apply(dc, fc, lengths.cuda(), None if noise is None else noise.cuda(), bg, 1e10)
    assert rel_err(f, o.features) < 1e-5 and rel_err(w, o.weights) < 1e-5
    ((f * cot[0].cuda()).sum() + (d * cot[1].cuda()).sum() + (m * cot[2].cuda()).sum() + (w * cot[3].cuda()).sum()).backward()
    e_d, e_f = rel_err(dc.grad, d64.grad), rel_err(fc.grad, f64.grad)
    print(f"EA backward n={n} S={S}: d_dens {e_d:.2e}, d_feats {e_f:.2e}")
    assert e_d < 1e-4 and e_f < 1e-5


@pytest.mark.parametrize("C,R,HW,S", [(16, 8, 8, 8), (32, 16, 12, 16)])
def test_implicit_function_backward_matches_autograd(C, R, HW, S):
    import holo_diffusion_b200 as hd
    grid, p = make_grid(C, R), make_mlp(C)
    b = _rays(HW, S)
    g = torch.Generator().manual_seed(2)
    cot_d, cot_f = torch.randn(1, HW, HW, S, 1, generator=g), torch.randn(1, HW, HW, S, 3, generator=g)
    # oracle, fp64 autograd
    p64 = {k: v.double().requires_grad_() for k, v in p.items()}
    g64 = grid.double().requires_grad_()
    b64 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), b.xys.double())
    o = ro.implicit_function(p64, g64, b64, R, 8.0)
    ((o[0] * cot_d.double()).sum() + (o[1] * cot_f.double()).sum()).backward()
    # ours
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=0).cuda()
    fn.render_mlp.load_state_dict(p, strict=True)
    gc = grid.cuda().requires_grad_()
    bundle = hd.ImplicitronRayBundle(b.origins.cuda(), b.directions.cuda(), b.lengths.cuda(), b.xys.cuda())
    dens, feats, _ = fn(ray_bundle=bundle, voxel_grid_features=gc)
    assert rel_err(dens, o[0]) < 2e-5 and rel_err(feats, o[1]) < 2e-5
    ((dens * cot_d.cuda()).sum() + (feats * cot_f.cuda()).sum()).backward()
    e_grid = rel_err(gc.grad, g64.grad)
    errs = {k: rel_err(dict(fn.render_mlp.named_parameters())[k].grad, p64[k].grad) for k in p}
    print(f"IF backward C={C} R={R}: d_grid {e_grid:.2e}, params max {max(errs.values()):.2e}")
    assert e_grid < 1e-4, e_grid
    assert max(errs.values()) < 2e-4, errs
    assert set(errs) == {n for n, _ in fn.render_mlp.named_parameters()}   # every layer received a gradient


def test_two_pass_render_backward_matches_autograd():
    """The reference's recursion (holo_multipass_ea.py:79-125) with autograd on: coarse pass -> refiner (no_grad) ->
    fine pass; loss on both stages' images, depths and masks (loss_prev_stage_* in configs/base.yaml)."""
    import holo_diffusion_b200 as hd
    C, R, HW, S, NF = 16, 8, 8, 8, 4
    grid, p = make_grid(C, R), make_mlp(C)
    b = _rays(HW, S)
    g = torch.Generator().manual_seed(3)
    cots = [torch.randn(1, HW, HW, 3, generator=g), torch.randn(1, HW, HW, 1, generator=g), torch.randn(1, HW, HW, 1, generator=g)]
    fn = hd.HoloVoxelGridImplicitFunction(resol=R, n_hidden=C, feature_dim=0).cuda()
    fn.render_mlp.load_state_dict(p, strict=True)
    w = hd.ImplicitFunctionWrapper(fn)
    gc = grid.cuda().requires_grad_()
    w.bind_args(voxel_grid_features=gc)
    rend = hd.HoloMultiPassEmissionAbsorptionRenderer(
        n_pts_per_ray_fine_evaluation=NF, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0)))
    bundle = hd.ImplicitronRayBundle(b.origins.cuda(), b.directions.cuda(), b.lengths.cuda(), b.xys.cuda())
    out = rend(bundle, [w, w], hd.EvaluationMode.EVALUATION)
    assert out.features.requires_grad and out.prev_stage is not None

    def loss_of(o, dev):
        t = 0
        for st, sc in ((o, 1.0), (o.prev_stage, 0.5)):
            t = t + sc * ((st.features * cots[0].to(dev)).sum() + (st.depths * cots[1].to(dev)).sum() + (st.masks * cots[2].to(dev)).sum())
        return t

    loss_of(out, "cuda").backward()
    # oracle on OUR fine depths (the refiner is under no_grad and ill-conditioned in fp32: matched inputs)
    p64 = {k: v.double().requires_grad_() for k, v in p.items()}
    g64 = grid.double().requires_grad_()
    b1 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), None)
    d1, f1 = ro.implicit_function(p64, g64, b1, R, 8.0)
    o1 = ro.ea_raymarch(d1, f1, b1.lengths)
    b2 = ro.OracleRayBundle(b1.origins, b1.directions, out.aux["lengths"].detach().cpu().double().reshape(1, HW, HW, S + NF), None)
    d2, f2 = ro.implicit_function(p64, g64, b2, R, 8.0)
    o2 = ro.ea_raymarch(d2, f2, b2.lengths)
    o2.prev_stage = o1
    cots = [c.double() for c in cots]
    loss_of(o2, "cpu").backward()
    e_grid = rel_err(gc.grad, g64.grad)
    errs = {k: rel_err(dict(fn.render_mlp.named_parameters())[k].grad, p64[k].grad) for k in p}
    print(f"two-pass backward: d_grid {e_grid:.2e}, params max {max(errs.values()):.2e}")
    assert e_grid < 2e-4 and max(errs.values()) < 3e-4
