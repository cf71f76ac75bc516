"""Run by tests/test_shim_cpu.py in a fresh interpreter: the ``holo_diffusion`` drop-in package over the Implicitron
config stand-in (oracle/pt3d_stub on sys.path => the registered facades) or without it (plain re-exports), with the
C-ABI entry points replaced by torch / oracle stand-ins.  Prints one JSON object."""
import json
import math
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
if mode == "stub":
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pt3d_stub"))

import torch  # noqa: E402
import yaml  # noqa: E402

import holo_diffusion_b200 as hd  # noqa: E402
from holo_diffusion_b200 import ops  # noqa: E402

import fake_encoder_ops  # noqa: E402
import fake_model_ops  # noqa: E402

fake_model_ops.install(ops)
fake_encoder_ops.install(ops)
hd.encoder.PAIR_DTYPE = torch.float32   # the GEMM stand-in keeps exact operand "pairs"
_fused_calls = []
_rf = ops.render_fwd
ops.render_fwd = lambda *a, **k: (_fused_calls.append(k.get("n_passes", 1)), _rf(*a, **k))[1]

import holo_diffusion  # noqa: E402
from holo_diffusion.holo_diffusion_model import HoloDiffusionModel  # noqa: E402  (experiment.py:73)
from holo_diffusion.utils.checkpoint_utils import load_experiment  # noqa: E402  (generate_samples.py:28)
from holo_diffusion.utils.render_utils.flyaround import render_flyaround  # noqa: E402  (generate_samples.py:29)

res = {"have_config": holo_diffusion.HAVE_CONFIG}
C, R, HW, S = 16, 8, 8, 8
MODEL_ARGS = dict(   # the keys configs/base.yaml sets under model_HoloDiffusionModel_args
    resol=R, feature_size=C, num_passes=2, render_image_width=HW, render_image_height=HW, chunk_size_grid=4096,
    net_3d_enabled=True, net_3d_class_type="SimpleUnet3D",
    net_3d_SimpleUnet3D_args=dict(model_channels=64, num_res_blocks=1, channel_mult=[1, 2], attention_resolutions=[2],
                                  num_heads=2),
    diffusion_enabled=True, diffusion_args=dict(num_steps=40),
    raysampler_class_type="AdaptiveRaySampler", raysampler_AdaptiveRaySampler_args=dict(n_pts_per_ray_evaluation=S),
    renderer_class_type="HoloMultiPassEmissionAbsorptionRenderer",
    renderer_HoloMultiPassEmissionAbsorptionRenderer_args=dict(
        n_pts_per_ray_fine_evaluation=4, raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=[1.0, 1.0, 1.0])),
    implicit_function_class_type="HoloVoxelGridImplicitFunction",
    implicit_function_HoloVoxelGridImplicitFunction_args=dict(render_mlp_args=dict(dnet_hidden_dim=32)))

if mode == "stub":
    from pytorch3d.implicitron.tools.config import registry
    res["registered"] = sorted(registry.classes)
    from holo_diffusion.utils.diffusion_utils import SimpleUnet3D, Unet3DBase
    res["registry_get"] = registry.get(Unet3DBase, "SimpleUnet3D") is SimpleUnet3D

def exact_pairs(m):   # the conv / attention stand-ins take exact operand "pairs": hi = value, lo = 0
    ex = hd.renderer.impl_of(m.net_3d)._exec
    ex.pair_dtype, ex.use_cuda_graph = torch.float32, False   # (no CUDA graphs on a CPU box)
    return m


torch.manual_seed(0)
model = exact_pairs(HoloDiffusionModel(use_cuda_graph=False, **MODEL_ARGS))
res["model_class"] = type(model).__module__ + "." + type(model).__name__
keys = list(model.state_dict().keys())
res["n_keys"] = len(keys)
res["has_unet_key"] = "net_3d._net.input_blocks.0.0.weight" in keys
res["has_mlp_key"] = "_implicit_functions.0._fn.render_mlp._density_net.mlp.0.0.weight" in keys
res["foreign_keys"] = [k for k in keys if not k.startswith(("net_3d._net.", "_implicit_functions."))][:5]
# the plain B200 model with the same weights gives the same view
plain = exact_pairs(hd.HoloDiffusionModel(use_cuda_graph=False, **MODEL_ARGS))
missing, unexpected = plain.load_state_dict(model.state_dict(), strict=True)
grid = torch.tanh(torch.randn(1, C, R, R, R, generator=torch.Generator().manual_seed(1)))
cams = hd.get_simple_360_camera_trajectory(2 * math.pi, 4, -math.pi / 6, 10.0, (-0.0396, -0.8306, -0.5554), 3.2)
a = model(camera=cams[[1]], voxel_features=grid, evaluation_mode=hd.EvaluationMode.EVALUATION)
b = plain(camera=cams[[1]], voxel_features=grid)
res["fused_render_calls"] = list(_fused_calls)   # facade plug-ins must still take the one-launch renderer
res["preds_keys"] = sorted(a.keys())
res["facade_vs_plain"] = float((a["images_render"] - b["images_render"]).abs().max())
res["image_shape"] = list(a["images_render"].shape)
res["prev_stage"] = a["rendered"].prev_stage is not None

# ---- the view-pooling encoder through the same constructor keys (configs/base.yaml:160-168, hydrant.yaml:184-193)
ENC_ARGS = dict(MODEL_ARGS, view_pooler_enabled=True, image_feature_extractor_class_type="ResNetFeatureExtractor",
                image_feature_extractor_ResNetFeatureExtractor_args=dict(proj_dim=4, image_rescale=0.5, stages=[1, 2]),
                view_pooler_args=dict(view_sampler_args=dict(masked_sampling=False, sampling_mode="bilinear"),
                                      feature_aggregator_class_type="MLPMeanFeatureAggregator",
                                      feature_aggregator_MLPMeanFeatureAggregator_args=dict(n_hidden=16, dim_out=16)))
enc_model = exact_pairs(HoloDiffusionModel(use_cuda_graph=False, **ENC_ARGS)).eval()
g2 = torch.Generator().manual_seed(2)
views = dict(image_rgb=torch.rand(4, 3, 32, 32, generator=g2), fg_probability=torch.rand(4, 1, 32, 32, generator=g2),
             mask_crop=torch.ones(4, 1, 32, 32), sequence_name=["seq"] * 4, camera=cams)
pe = enc_model(**views)
res["encoder_modules"] = sorted({k.split(".")[0] for k in enc_model.state_dict()})
res["encoder_grid"] = [list(pe["voxel_features"].shape), float(pe["voxel_features"].abs().max())]
plain_enc = exact_pairs(hd.HoloDiffusionModel(use_cuda_graph=False, **ENC_ARGS)).eval()
plain_enc.load_state_dict(enc_model.state_dict(), strict=True)
res["encoder_facade_vs_plain"] = float((plain_enc(**views)["images_render"] - pe["images_render"]).abs().max())

with tempfile.TemporaryDirectory() as tmp:
    # a checkpoint of the encoder model loads with its encoder keys (image_feature_extractor.*, view_pooler.*, pooled_feature_mapper.*)
    cfg_e = {"model_factory_ImplicitronModelFactory_args": {"model_class_type": "HoloDiffusionModel",
                                                            "model_HoloDiffusionModel_args": ENC_ARGS}}
    yaml.safe_dump(cfg_e, open(os.path.join(tmp, "expconfig.yaml"), "w"))
    torch.save(dict(enc_model.state_dict()), os.path.join(tmp, "model_epoch_00000001.pth"))
    _, m3, _ = load_experiment(None, tmp, None, (HW, HW), 3, torch.device("cpu"))
    m3.use_cuda_graph = False
    exact_pairs(m3).eval()
    if hasattr(m3, "_impl"):
        m3._impl.use_cuda_graph = False
    res["encoder_ckpt_roundtrip"] = float((m3(**views)["images_render"] - pe["images_render"]).abs().max())

# ---- reconstruction fly-around (flyaround.py:148-171): frames from a dataset, seeded source views, encoder, renders
class _Frame:
    def __init__(self, i):
        g = torch.Generator().manual_seed(100 + i)
        self.image_rgb, self.fg_probability = torch.rand(3, 32, 32, generator=g), torch.rand(1, 32, 32, generator=g)
        self.mask_crop, self.depth_map = torch.ones(1, 32, 32), None
        self.camera, self.sequence_name, self.frame_number = cams[[i % 4]], "seq", i
        self.sequence_category = self.frame_timestamp = None


class _Dataset:
    def sequence_indices_in_order(self, name):
        return iter(range(6))

    def __getitem__(self, i):
        return _Frame(i)


with tempfile.TemporaryDirectory() as tmp:
    rec = render_flyaround(dataset=_Dataset(), sequence_name="seq", model=enc_model, output_video_path=os.path.join(tmp, "v"),
                           n_flyaround_poses=2, trajectory_type="simple_360", device="cpu", sample_mode=False,
                           n_source_views=3, seed=5, up=(-0.0396, -0.8306, -0.5554),
                           visualize_preds_keys=("images_render", "_all_source_images"))
    import numpy as np
    res["reconstruction"] = {k: list(np.load(v).shape) for k, v in rec.items() if v.endswith(".npy")}
    # the reference's way: every pose calls forward with the images (re-encoding the same sources) -- same frames
    with torch.random.fork_rng():
        torch.manual_seed(5)
        pick = torch.nn.functional.pad(torch.randperm(6)[:3], [1, 0]).tolist()
    fr = [_Frame(i) for i in pick]
    traj = hd.get_simple_360_camera_trajectory(2 * math.pi, 2, -30.0 * (2 * math.pi / 360), 10, (-0.0396, -0.8306, -0.5554), 3.2)
    worst = 0
    for n in range(2):
        cam = hd.PerspectiveCameras(torch.cat([traj[[n]].focal_length.expand(1, 2)] + [f.camera.focal_length.expand(1, 2) for f in fr[1:]]),
                                    torch.cat([traj[[n]].principal_point] + [f.camera.principal_point for f in fr[1:]]),
                                    torch.cat([traj[[n]].R] + [f.camera.R for f in fr[1:]]),
                                    torch.cat([traj[[n]].T] + [f.camera.T for f in fr[1:]]))
        direct = enc_model(image_rgb=torch.stack([f.image_rgb for f in fr]), camera=cam,
                           fg_probability=torch.stack([f.fg_probability for f in fr]),
                           mask_crop=torch.stack([f.mask_crop for f in fr]), sequence_name=["seq"] * 4)
        frame = ops.frame_u8(direct["images_render"][0].contiguous(), (8, 8))
        worst = max(worst, int((frame.int() - torch.from_numpy(np.load(rec["images_render"]))[n].int()).abs().max()))
    res["reconstruction_vs_per_pose_forward"] = worst

with tempfile.TemporaryDirectory() as tmp:
    # generate_samples.py:87-138: exp_dir with expconfig.yaml + checkpoint -> load_experiment -> render_flyaround
    cfg = {"model_factory_ImplicitronModelFactory_args": {"model_class_type": "HoloDiffusionModel",
                                                          "model_HoloDiffusionModel_args": MODEL_ARGS},
           "data_source_ImplicitronDataSource_args": {
               "data_loader_map_provider_SequenceDataLoaderMapProvider_args": {"batch_size": 10}}}
    yaml.safe_dump(cfg, open(os.path.join(tmp, "expconfig.yaml"), "w"))
    sd = dict(model.state_dict())
    sd["image_feature_extractor.stem.0.weight"] = torch.zeros(1)   # encoder-side keys of a real checkpoint are skipped
    torch.save(sd, os.path.join(tmp, "model_epoch_00000007.pth"))
    _, m2, data_source = load_experiment(None, tmp, None, (HW, HW), 3, torch.device("cpu"))
    m2.use_cuda_graph = False
    exact_pairs(m2)
    if hasattr(m2, "_impl"):
        m2._impl.use_cuda_graph = False
    assert m2.net_3d_enabled and m2.diffusion_enabled
    res["n_source_views"] = data_source.data_loader_map_provider.batch_size - m2.n_train_target_views
    res["ckpt_loaded"] = bool(torch.equal(m2.state_dict()[keys[0]], model.state_dict()[keys[0]]))
    out_dir = os.path.join(tmp, "generated_samples")
    os.makedirs(out_dir)
    written = render_flyaround(
        dataset=None, sequence_name="sample_00000", model=m2, output_video_path=os.path.join(out_dir, "video"),
        output_video_name="sample_00000", n_source_views=res["n_source_views"], n_flyaround_poses=3,
        trajectory_type="simple_360", video_resize=(16, 16), device="cpu", up=(-0.0396, -0.8306, -0.5554),
        trajectory_scale=1.3, camera_elevation=-30.0 * (2 * math.pi / 360), sample_mode=True,
        progressive_sampling_steps_per_render=-1,
        visualize_preds_keys=("images_render", "masks_render", "depths_render", "noise_render",
                              "images_prev_stage_render", "features_prev_stage_render", "_shaded_depth_render",
                              "_all_source_images"),
        save_voxel_features=True)
    import numpy as np
    res["videos"] = {k: list(np.load(v).shape) for k, v in written.items() if v.endswith(".npy")}
    res["video_dtype"] = str(np.load(next(iter(written.values()))).dtype) if written else None
    res["saved_voxels"] = os.path.exists(os.path.join(out_dir, "sample_00000_voxel_features.pth"))
    # progressive sampling: a render every k denoising steps (generate_samples.py:49)
    written2 = render_flyaround(dataset=None, sequence_name="p", model=m2, output_video_path=os.path.join(out_dir, "pv"),
                                n_flyaround_poses=2, trajectory_type="simple_360", device="cpu", sample_mode=True,
                                progressive_sampling_steps_per_render=1, visualize_preds_keys=("images_render",))
    res["progressive"] = {k: list(np.load(v).shape) for k, v in written2.items() if v.endswith(".npy")}
print("RESULT " + json.dumps(res))
