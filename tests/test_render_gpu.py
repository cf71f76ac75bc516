"""GPU parity of the fused renderer against the CPU oracle (through the C-ABI)."""
import pytest
import torch

from conftest import rel_err
from fixtures import make_grid, make_mlp
from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-4  # BASELINE.json north_star: fp32 render tensors within 1e-4 relative (max|a-b| <= 1e-4 max|b|)


def _pack(p, C, dev):
    from holo_diffusion_b200 import ops
    layers = [(p[f"_density_net.mlp.{i}.0.weight"].to(dev), p[f"_density_net.mlp.{i}.0.bias"].to(dev)) for i in range(4)]
    return ops.collapse_and_pack_render_mlp(layers, (2,), p["_radiance_net.mlp.0.0.weight"].to(dev),
                                            p["_radiance_net.mlp.0.0.bias"].to(dev), C)


def _cuda_render(grid, p, bundle, extent, n_passes, n_fine, weights=True, tc=False):
    from holo_diffusion_b200 import ops
    dev = "cuda"
    C = grid.shape[1]
    packed, H, E, tc_image = _pack(p, C, dev)
    if tc:
        assert tc_image is not None
    g = grid[0].permute(1, 2, 3, 0).contiguous().to(dev)
    n = bundle.origins[0].reshape(-1, 3).shape[0]
    S = bundle.lengths.shape[-1]
    out = ops.render_fwd(g, extent, packed, H, (E // 3 - 1) // 2, bundle.origins[0].reshape(n, 3).contiguous().to(dev),
                         bundle.directions[0].reshape(n, 3).contiguous().to(dev),
                         bundle.lengths[0].reshape(n, S).contiguous().to(dev), n_passes=n_passes, n_fine=n_fine,
                         return_weights=weights, tc_image=tc_image if tc else None)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("C,R,HW,S,n_passes,n_fine,tc", [
    (16, 32, 64, 16, 1, 0, False),    # BASELINE cfg #1, single pass
    (16, 32, 64, 16, 2, 16, False),   # cfg #1, reference-faithful two-pass
    (32, 16, 24, 64, 2, 16, False),   # base.yaml sampling (64 + 16) on a small grid
    (8, 8, 9, 5, 2, 3, False),        # odd sizes: ragged tile, odd S, tiny n_fine
    (64, 16, 16, 32, 2, 64, False),   # shipped configs: 64 channels, teddybear.yaml n_fine 64
    (16, 32, 64, 16, 1, 0, True),     # the same through the tcgen05 kernel
    (16, 32, 64, 16, 2, 16, True),
    (32, 16, 24, 64, 2, 16, True),
    (32, 8, 9, 5, 2, 3, True),        # ragged: 81 rays in one 256-ray CTA, odd S
    (32, 16, 40, 8, 2, 64, True),     # several CTAs, more fine than coarse samples
])
def test_render_matches_oracle(C, R, HW, S, n_passes, n_fine, tc):
    grid, p = make_grid(C, R), make_mlp(C)
    cams = ro.simple_360_cameras(8)
    b = ro.sample_rays(cams[1], HW, HW, S)
    ref = ro.render_chunked(p, grid, b, R, 8.0, n_passes, n_fine, chunk_size_grid=0)
    out = _cuda_render(grid, p, b, 8.0, n_passes, n_fine, tc=tc)
    n = HW * HW
    assert ref.masks.min() < 0.9 and ref.masks.max() > 0.9, "fixture must exercise compositing"
    if n_passes == 1:
        for k in ("features", "depths", "masks", "weights"):
            assert rel_err(out[k], getattr(ref, k).reshape(n, -1)) < TOL, k
        return
    # ---- two passes.  The importance re-sampling (RayPointRefiner/sample_pdf) is ill-conditioned in fp32: a
    # 1-ulp change of the cdf normaliser moves a sample inside a near-empty bin by ~1e-3 (the fp32 oracle itself
    # is ~1e-3 away from its fp64 twin, see DESIGN.md "Parity method").  So parity is asserted stage by stage
    # on MATCHED inputs at the 1e-4 bar, and end to end against the fp64 twin within the oracle's own noise.
    prev = ref.prev_stage
    for k in ("features", "depths", "masks", "weights"):                      # stage 0: coarse pass
        assert rel_err(out["prev"][k], getattr(prev, k).reshape(n, -1)) < TOL, "prev." + k
    l = out["lengths"]
    assert bool((l[:, 1:] >= l[:, :-1]).all())                                 # torch.sort in the refiner
    w0 = out["prev"]["weights"].cpu()                                          # stage 1: refiner on the SAME weights
    z0 = b.lengths.reshape(n, S)
    l_ref32 = ro.refine_lengths(z0, w0, n_fine)
    l_ref64 = ro.refine_lengths(z0.double(), w0.double(), n_fine)
    noise = rel_err(l_ref32, l_ref64)
    assert rel_err(l, l_ref64) < 3 * noise + 1e-5
    b2 = ro.OracleRayBundle(b.origins.reshape(1, n, 3), b.directions.reshape(1, n, 3), l.cpu()[None], None)
    dens, feats = ro.implicit_function(p, grid, b2, R, 8.0)                   # stage 2: fine pass on the SAME depths
    fine = ro.ea_raymarch(dens, feats, b2.lengths)
    for k in ("features", "depths", "masks", "weights"):
        assert rel_err(out[k], getattr(fine, k).reshape(n, -1)) < TOL, k
    # end to end against the fp64 twin, bounded by the fp32 oracle's own distance to it
    p64 = {k: v.double() for k, v in p.items()}
    b64 = ro.OracleRayBundle(b.origins.double(), b.directions.double(), b.lengths.double(), b.xys.double())
    ref64 = ro.render_chunked(p64, grid.double(), b64, R, 8.0, n_passes, n_fine, chunk_size_grid=0)
    for k in ("features", "depths", "masks"):
        own = rel_err(getattr(ref, k), getattr(ref64, k))
        assert rel_err(out[k], getattr(ref64, k).reshape(n, -1)) < 3 * own + TOL, "e2e." + k


def test_raygen_matches_oracle():
    from holo_diffusion_b200 import ops
    cams = ro.simple_360_cameras(8)
    H = W = 32
    S = 16
    b = ro.sample_rays(cams, H, W, S)
    xy = ro.ndc_xy_grid(H, W).reshape(-1, 2).contiguous().cuda()
    o, d, l = ops.raygen(cams.R.cuda().contiguous(), cams.T.cuda().contiguous(), cams.focal.cuda().contiguous(),
                         cams.pp.cuda().contiguous(), xy, S, 4.0)
    assert rel_err(o, b.origins.reshape(8, -1, 3)) < 1e-5
    assert rel_err(d, b.directions.reshape(8, -1, 3)) < 1e-5
    assert rel_err(l, b.lengths.reshape(8, -1, S)) < 1e-6


def test_render_empty_and_outside():
    """rays that miss the volume composite to pure background; n_rays == 0 is a no-op."""
    from holo_diffusion_b200 import ops
    C, R = 16, 8
    grid, p = make_grid(C, R), make_mlp(C)
    packed, H, E, _ = _pack(p, C, "cuda")
    g = grid[0].permute(1, 2, 3, 0).contiguous().cuda()
    o = torch.tensor([[100.0, 100.0, 100.0]] * 4).cuda()
    d = torch.tensor([[0.0, 0.0, 1.0]] * 4).cuda()
    l = torch.linspace(1, 9, 8)[None].repeat(4, 1).contiguous().cuda()
    out = ops.render_fwd(g, 8.0, packed, H, 4, o, d, l, n_passes=2, n_fine=4)
    torch.cuda.synchronize()
    assert torch.allclose(out["features"].cpu(), torch.ones(4, 3))
    assert float(out["masks"].abs().max()) == 0.0
    out0 = ops.render_fwd(g, 8.0, packed, H, 4, o[:0].contiguous(), d[:0].contiguous(), l[:0].contiguous())
    assert out0["features"].shape == (0, 3)
