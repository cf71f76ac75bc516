"""The ``holo_diffusion`` drop-in package (repo root): the reference's module paths / class names over the B200
implementation, (a) as plain re-exports and (b) as Implicitron-registered facades when the config system imports
(here: the stand-in under oracle/pt3d_stub -- pytorch3d itself is not installable in this image).  Each mode runs
in a fresh interpreter (tests/shim_stub_driver.py) so that the stand-in ``pytorch3d`` never leaks into this process;
the driver follows generate_samples.py:87-138: load_experiment(exp_dir) -> render_flyaround(sample_mode=True)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(mode):
    r = subprocess.run([sys.executable, os.path.join(HERE, "shim_stub_driver.py"), mode], capture_output=True, text=True,
                       timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stderr[-3000:]
    return json.loads(lines[-1][len("RESULT "):])


@pytest.mark.parametrize("mode", ["plain", "stub"])
def test_drop_in_package(mode):
    res = _run(mode)
    assert res["have_config"] == (mode == "stub")
    if mode == "stub":   # the five @registry.register names of the reference's holo_diffusion package (custom_modules.py:162,
        # holo_multipass_ea.py:15, holo_voxel_grid_implicit_function.py:148, holo_diffusion_model.py:44, diffusion_utils.py:41)
        assert res["registered"] == ["HoloDiffusionModel", "HoloMultiPassEmissionAbsorptionRenderer",
                                     "HoloVoxelGridImplicitFunction", "MLPMeanFeatureAggregator", "SimpleUnet3D"]
        assert res["registry_get"]
        assert res["model_class"] == "holo_diffusion.holo_diffusion_model.HoloDiffusionModel"
    else:
        assert res["model_class"] == "holo_diffusion_b200.model.HoloDiffusionModel"
    # reference checkpoints load: same parameter names, nothing else in the state dict
    assert res["has_unet_key"] and res["has_mlp_key"] and res["foreign_keys"] == [] and res["ckpt_loaded"]
    # forward(): reference preds keys; facade == plain implementation bit for bit; fused renderer dispatch kept
    assert {"images_render", "depths_render", "masks_render", "implicitron_render", "rendered"} <= set(res["preds_keys"])
    assert res["facade_vs_plain"] == 0.0 and res["image_shape"] == [1, 3, 8, 8] and res["prev_stage"]
    assert res["fused_render_calls"] == [2, 2]
    # generate_samples.py call pattern: 10 - 1 source views, 3 poses, 4 visualised keys (the others are absent in
    # sampling mode and skipped like in the reference), resized 8-bit frames, saved voxel features, progressive mode
    assert res["n_source_views"] == 9
    assert res["videos"] == {k: [3, 16, 16, 3] for k in ("images_render", "masks_render", "depths_render",
                                                         "_shaded_depth_render")}
    assert res["video_dtype"] == "uint8" and res["saved_voxels"]
    assert res["progressive"] == {"images_render": [2, 8, 8, 3]}
    # the view-pooling encoder under the reference's constructor keys and state-dict names: forward(image_rgb=...) runs,
    # facade == plain implementation, and a checkpoint with its encoder keys round-trips through load_experiment
    assert {"image_feature_extractor", "view_pooler", "pooled_feature_mapper"} <= set(res["encoder_modules"])
    assert res["encoder_grid"][0] == [1, 16, 8, 8, 8] and 0 < res["encoder_grid"][1] <= 1.0
    assert res["encoder_facade_vs_plain"] == 0.0 and res["encoder_ckpt_roundtrip"] == 0.0
    # reconstruction fly-around: dataset frames -> seeded source views -> encoder (once) -> renders; the mosaic of the 3
    # source images is 2 x 2 tiles; frames identical to calling forward with the images for every pose (the reference's loop)
    assert res["reconstruction"] == {"images_render": [2, 8, 8, 3], "_all_source_images": [2, 64, 64, 3]}
    assert res["reconstruction_vs_per_pose_forward"] == 0
