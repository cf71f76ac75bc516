"""Full-size parity at BASELINE.json configs[1] (64^3 x 32ch grid, base-args UNet at t = 0, 256^2 view, 64 + 16
pts/ray): our kernels against the oracle restatement executed with plain torch ops on the SAME GPU in fp32 (TF32 off)
-- the CPU oracle needs minutes at this size.  Also records, as an informational line in
gpurun_out/eager_gpu_comparator.json, how long that reference-equivalent eager PyTorch path takes with PyTorch's
default TF32 settings (SURVEY.md section 8d "reference GPU path"): a comparator, not a bench value."""
import json
import math
import os
import time

import pytest
import torch

from conftest import rel_err
from fixtures import make_grid, make_mlp
from oracle import render_oracle as ro
from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu
TOL = 1e-4
UNET = dict(model_channels=64, num_res_blocks=2, num_heads=2, channel_mult=[1, 1, 2, 4, 8], attention_resolutions=[4, 8])


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _bundle_cuda(b, dtype=torch.float32):
    return ro.OracleRayBundle(b.origins.to("cuda", dtype), b.directions.to("cuda", dtype), b.lengths.to("cuda", dtype),
                              b.xys.to("cuda", dtype))


def test_cfg2_full_size_matches_oracle_on_gpu():
    import holo_diffusion_b200 as hd
    C, R, HW, S, NF = 32, 64, 256, 64, 16
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = hd.HoloDiffusionModel(
        resol=R, feature_size=C, num_passes=2, render_image_width=HW, render_image_height=HW,
        net_3d_SimpleUnet3D_args=UNET, raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=S),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
            n_pts_per_ray_fine_evaluation=NF, return_weights=True,
            raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0))),
        use_cuda_graph=False)
    sd, mlp = uo.make_unet_state_dict(C, C, seed=2), make_mlp(C)
    m.net_3d._net.load_state_dict(sd, strict=True)
    m._implicit_functions[0]._fn.render_mlp.load_state_dict(mlp, strict=True)
    m.cuda()
    grid = make_grid(C, R, seed=0)
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, 8, -math.pi / 6, 10.0, ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    preds = m(camera=cams[[3]].to("cuda"), voxel_features=grid.cuda())
    torch.cuda.synchronize()
    out = preds["rendered"]
    n = HW * HW
    # ---- denoiser at full size (165 M parameters, 1.18 TFLOP): tanh(UNet(g, 0))
    sd_c = _cuda(sd)
    with torch.no_grad():
        g_ref = torch.tanh(uo.unet_forward(sd_c, grid.cuda(), torch.zeros(1, dtype=torch.long, device="cuda")))
        sd64 = {k: v.double() for k, v in sd_c.items()}
        g64 = torch.tanh(uo.unet_forward(sd64, grid.cuda().double(), torch.zeros(1, dtype=torch.long, device="cuda")))
    e_unet = rel_err(preds["voxel_features"], g64)        # ours vs the exact (fp64) answer
    e_unet32 = rel_err(preds["voxel_features"], g_ref)    # ours vs the fp32 eager path (cuDNN picks the algorithms)
    own = rel_err(g_ref, g64)                             # the fp32 eager path's own distance to the exact answer
    print(f"tanh(UNet) at 64^3: ours vs fp64 {e_unet:.2e}, ours vs fp32 eager {e_unet32:.2e}, fp32 eager vs fp64 {own:.2e}")
    assert e_unet < TOL, e_unet
    assert e_unet32 < TOL + own, (e_unet32, own)
    # ---- renderer at full size on the MATCHED grid (our tanh(UNet) output), stage by stage as in test_render_gpu
    g_ours = preds["voxel_features"].detach().clone()
    b = ro.sample_rays(ro.simple_360_cameras(8)[3], HW, HW, S)
    assert torch.equal(preds["ray_bundle"].xys.cpu()[0], ro.ndc_xy_grid(HW, HW))  # ray order r = h W + w
    bc = _bundle_cuda(b)
    mlp_c = _cuda(mlp)
    with torch.no_grad():
        ref = ro.render_chunked(mlp_c, g_ours, bc, R, 8.0, 2, NF, chunk_size_grid=163840 * 8)
        for k in ("features", "depths", "masks", "weights"):                     # coarse pass
            assert rel_err(getattr(out.prev_stage, k), getattr(ref.prev_stage, k)) < TOL, "prev." + k
        l = out.aux["lengths"].reshape(n, S + NF)
        assert bool((l[:, 1:] >= l[:, :-1]).all())
        z0, w0 = bc.lengths.reshape(n, S), out.prev_stage.weights.reshape(n, S)
        l32 = ro.refine_lengths(z0, w0, NF)
        l64 = ro.refine_lengths(z0.double(), w0.double(), NF)
        assert rel_err(l, l64) < 3 * rel_err(l32, l64) + 1e-5                     # refiner on the SAME weights
        b2 = ro.OracleRayBundle(bc.origins.reshape(1, n, 3), bc.directions.reshape(1, n, 3), l[None], None)
        dens, feats = ro.implicit_function(mlp_c, g_ours, b2, R, 8.0)             # fine pass on the SAME depths
        fine = ro.ea_raymarch(dens, feats, b2.lengths)
        for k in ("features", "depths", "masks", "weights"):
            assert rel_err(getattr(out, k).reshape(n, -1), getattr(fine, k).reshape(n, -1)) < TOL, k
    w = ref.prev_stage.weights.reshape(n, S)
    eff = float((w.sum(-1) ** 2 / (w * w).sum(-1).clamp_min(1e-30)).median())   # samples that carry a ray's weight
    assert eff > 1.5, "fixture must exercise compositing (several samples per ray contribute)"
    # ---- informational: the same eager torch path with PyTorch's defaults (TF32 convolutions on), timed
    torch.backends.cudnn.allow_tf32 = True
    res = {}
    with torch.no_grad():
        for chunk in (4096, 163840):
            ts = []
            for it in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g = torch.tanh(uo.unet_forward(sd_c, grid.cuda(), torch.zeros(1, dtype=torch.long, device="cuda")))
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                ro.render_chunked(mlp_c, g, bc, R, 8.0, 2, NF, chunk_size_grid=chunk)
                torch.cuda.synchronize()
                ts.append((t1 - t0, time.perf_counter() - t1))
            res[f"chunk_size_grid={chunk}"] = {"unet_s": ts[-1][0], "render_s": ts[-1][1],
                                               "views_per_s": 1.0 / (ts[-1][0] + ts[-1][1])}
    torch.backends.cudnn.allow_tf32 = False
    rec = {"effective_samples_per_ray_median": eff, "mask_min": float(ref.masks.min()), "mask_max": float(ref.masks.max()),
           "what": "oracle restatement of the reference (torch eager ops, cuDNN TF32 default) on the same B200, cfg #2",
           "unet_rel_err_ours_vs_fp64": e_unet, "unet_rel_err_ours_vs_fp32_eager": e_unet32,
           "unet_rel_err_fp32_eager_vs_fp64": own, **res}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rec, open(os.path.join("gpurun_out", "eager_gpu_comparator.json"), "w"))
    print(json.dumps(rec))
