"""Per-view post-processing kernels against the CPU restatement in oracle/postprocess_oracle.py:
depth visualisation (exact order statistics), 8-bit frame packing with resize, screen-space shaded depth; and the
fly-around loop of generate_samples.py through the drop-in package on the GPU."""
import math
import os

import pytest
import torch

from conftest import rel_err
from oracle import postprocess_oracle as po

pytestmark = pytest.mark.gpu


def _depth_mask(H, W, seed, frac=0.6):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    r2 = xx ** 2 + yy ** 2
    mask = (r2 < frac).float() * (0.9 + 0.1 * torch.rand(H, W, generator=g))
    depth = (8.0 + 1.5 * torch.sqrt((frac - r2).clamp(0)) * -1 + 0.05 * torch.randn(H, W, generator=g)) * (mask > 0).float()
    return depth.contiguous(), mask.contiguous()


@pytest.mark.parametrize("H,W,seed", [(256, 256, 0), (64, 96, 1), (33, 17, 2)])
def test_depth_image_matches_oracle(H, W, seed):
    from holo_diffusion_b200 import ops
    d, m = _depth_mask(H, W, seed)
    ref, nf_ref = po.depth_frame(d[None, None], m[None, None])
    out, nf = ops.depth_image(d.cuda(), m.cuda())
    torch.cuda.synchronize()
    assert torch.equal(nf.cpu(), nf_ref[0]), (nf, nf_ref)   # EXACT order statistics (topk semantics)
    assert rel_err(out, ref[0]) < 1e-6


def test_depth_image_degenerate_masks():
    """<= 1 valid pixel -> normalisers (0, 0) (make_depth_image's early exit); everything masked -> white."""
    from holo_diffusion_b200 import ops
    d = torch.full((16, 16), 5.0)
    m = torch.zeros(16, 16)
    out, nf = ops.depth_image(d.cuda(), m.cuda())
    assert nf.tolist() == [0.0, 0.0] and torch.equal(out.cpu(), torch.ones(3, 16, 16))
    m[3, 4] = 1.0
    ref, nf_ref = po.depth_frame(d[None, None], m[None, None])
    out, nf = ops.depth_image(d.cuda(), m.cuda())
    assert nf.tolist() == [0.0, 0.0] and torch.equal(out.cpu(), ref[0])
    # ties: many equal depths
    d = torch.full((32, 32), 7.0)
    d[:4] = 6.0
    m = torch.ones(32, 32)
    ref, nf_ref = po.depth_frame(d[None, None], m[None, None])
    out, nf = ops.depth_image(d.cuda(), m.cuda())
    assert torch.equal(nf.cpu(), nf_ref[0]) and rel_err(out, ref[0]) < 1e-6


@pytest.mark.parametrize("C,size,out_hw", [(3, (64, 64), (256, 256)), (1, (48, 80), (24, 40)), (3, (32, 32), None)])
def test_frame_u8_matches_oracle(C, size, out_hw):
    from holo_diffusion_b200 import ops
    x = torch.rand(C, *size, generator=torch.Generator().manual_seed(3)) * 1.4 - 0.2   # exercises the clip
    ref = po.frame_u8(x, out_hw)
    out = ops.frame_u8(x.cuda(), out_hw).cpu()
    assert out.shape == ref.shape and out.dtype == torch.uint8
    diff = (out.int() - ref.int()).abs()
    assert diff.max() <= 1 and (diff > 0).float().mean() < 0.02   # rounding of values that sit on x.5 only


@pytest.mark.parametrize("H,W", [(128, 128), (64, 96)])
def test_shade_depth_matches_oracle(H, W):
    from holo_diffusion_b200 import ops
    d, m = _depth_mask(H, W, 5)
    focal, pp = (3.2, 3.2), (0.0, 0.0)
    ref, ref_mask, k = po.shade_depth(d, m, focal, pp)
    out, om = ops.shade_depth(d.cuda(), m.cuda(), focal, pp, k)
    torch.cuda.synchronize()
    assert torch.equal(om.cpu(), ref_mask)
    # fp32 kernel vs fp64 restatement; specular exponent 128 amplifies normal errors, hence 2e-3 absolute on [0, 1]
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-3, err
    assert ref_mask.sum() > 0.2 * H * W and float(out.cpu()[:, ref_mask > 0].std()) > 0.01   # a real shaded surface


def test_flyaround_through_the_drop_in_package(tmp_path):
    """generate_samples.py's loop on the GPU: sample (a few DDPM steps), render 3 poses, post-process on the device,
    one D2H per key."""
    import numpy as np
    from holo_diffusion.holo_diffusion_model import HoloDiffusionModel
    from holo_diffusion.utils.render_utils.flyaround import render_flyaround
    torch.manual_seed(0)
    model = HoloDiffusionModel(
        resol=16, feature_size=16, num_passes=2, render_image_width=32, render_image_height=32,
        net_3d_SimpleUnet3D_args=dict(model_channels=64, num_res_blocks=1, channel_mult=[1, 2], attention_resolutions=[2],
                                      num_heads=2),
        diffusion_args=dict(num_steps=50), raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=16),
        renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
            n_pts_per_ray_fine_evaluation=8, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=(1.0, 1.0, 1.0)))).cuda()
    with torch.no_grad():
        model._implicit_functions[0]._fn.render_mlp._density_net.mlp[-1][0].weight[-1] *= 8.0
    written = render_flyaround(dataset=None, sequence_name="s", model=model, output_video_path=str(tmp_path / "video"),
                               n_flyaround_poses=3, trajectory_type="simple_360", video_resize=(48, 48), device="cuda",
                               up=(-0.0396, -0.8306, -0.5554), sample_mode=True,
                               visualize_preds_keys=("images_render", "masks_render", "depths_render", "_shaded_depth_render"),
                               save_voxel_features=True)
    assert set(written) == {"images_render", "masks_render", "depths_render", "_shaded_depth_render"}
    for k, path in written.items():
        fr = np.load(path if path.endswith(".npy") else path[:-4] + ".npy")
        assert fr.shape == (3, 48, 48, 3) and fr.dtype == np.uint8
    imgs = np.load(str(tmp_path / "video_s_images_render.npy"))
    assert imgs.std() > 0 and not np.array_equal(imgs[0], imgs[1])   # different poses, different frames
    assert os.path.exists(str(tmp_path / "s_voxel_features.pth"))
