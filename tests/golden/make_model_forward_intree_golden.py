"""Golden vectors from the reference's ``HoloDiffusionModel.forward`` (holo_diffusion_model.py:201-540): the method's
SOURCE is read from /root/reference and executed UNMODIFIED (the module itself needs pytorch3d's GenericModel), on a
stand-in ``self`` whose collaborators are the pinned pieces: ``net_3d`` = the UNet oracle, ``_implicit_functions`` +
``_render`` = the renderer oracle (bound ``voxel_grid_features``), ``raysampler`` = the oracle's ray sampler.

What this pins is the ORCHESTRATION of rows a8 / a16 of SURVEY.md section 8: only ``camera[0]`` is rendered in
evaluation mode, the extra UNet pass at t = 0 followed by tanh, the range asserts, bind -> rays -> render -> unbind,
the (B, H, W, C) -> (B, C, H, W) permutes and the ``preds`` keys.

    python tests/golden/make_model_forward_intree_golden.py    # writes tests/golden/model_forward_intree_ref.npz
"""
import ast
import enum
import logging
import os
import sys
import types
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "model_forward_intree_ref.npz")
C, R, HW, S, NF = 8, 8, 6, 8, 4
UNET = dict(model_ch=32, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(2,), heads=1, seed=5)


def generate():
    for p in (ROOT, REF, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "pt3d_stub")):   # REF ahead of ROOT: holo_diffusion = the reference, not the shim
        if p not in sys.path:
            sys.path.insert(0, p)
    from fixtures import make_grid, make_mlp
    from pytorch3d.implicitron.models.renderer.base import EvaluationMode, ImplicitronRayBundle

    from oracle import render_oracle as ro
    from oracle import unet_oracle as uo

    src = open(os.path.join(REF, "holo_diffusion/holo_diffusion_model.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "HoloDiffusionModel")
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")

    class RenderSamplingMode(enum.Enum):
        MASK_SAMPLE = "mask_sample"
        FULL_GRID = "full_grid"

    class ImplicitronRender:
        def __init__(self, image_render=None, depth_render=None, mask_render=None):
            self.image_render, self.depth_render, self.mask_render = image_render, depth_render, mask_render

    ns = {"torch": torch, "np": np, "logger": logging.getLogger("ref"), "Optional": Optional, "Union": Union, "List": List,
          "Dict": Dict, "Any": Any, "CamerasBase": object, "EvaluationMode": EvaluationMode,
          "RenderSamplingMode": RenderSamplingMode, "ImplicitronRayBundle": ImplicitronRayBundle,
          "ImplicitronRender": ImplicitronRender,
          "preprocess_input": lambda image_rgb, fg, depth, *a: (image_rgb, fg, depth),   # nothing to mask: all None
          "VolumeLocator": None, "rasterize_sparse_ray_bundle": None}
    exec(compile(ast.Module(body=[fwd], type_ignores=[]), "holo_diffusion_model.py", "exec"), ns)
    forward = ns["forward"]

    sd = uo.make_unet_state_dict(C, C, UNET["model_ch"], UNET["num_res_blocks"], UNET["channel_mult"],
                                 UNET["attention_resolutions"], seed=UNET["seed"])
    mlp = make_mlp(C)
    log = []

    class Bound:   # ImplicitFunctionWrapper.bind_args / unbind_args
        def __init__(self):
            self.bound = None

        def bind_args(self, **kw):
            log.append("bind")
            self.bound = kw

        def unbind_args(self):
            log.append("unbind")
            self.bound = None

    ifs = [Bound(), Bound()]

    def net_3d(x, timesteps):
        log.append(("net_3d", timesteps.tolist()))
        return uo.unet_forward(sd, x, timesteps, n_heads=UNET["heads"])

    def raysampler(cameras, evaluation_mode, mask=None):
        log.append(("rays", int(cameras.R.shape[0]), evaluation_mode.name, mask is None))
        b = ro.sample_rays(ro.OracleCameras(cameras.R, cameras.T, cameras.focal, cameras.pp), HW, HW, S)
        return ImplicitronRayBundle(b.origins, b.directions, b.lengths, b.xys)

    def _render(*, ray_bundle, sampling_mode, evaluation_mode, implicit_functions, inputs_to_be_chunked):
        log.append(("render", sampling_mode.name, evaluation_mode.name, len(implicit_functions)))
        grid = implicit_functions[0].bound["voxel_grid_features"]
        ob = ro.OracleRayBundle(ray_bundle.origins, ray_bundle.directions, ray_bundle.lengths, ray_bundle.xys)
        return ro.render_chunked(mlp, grid, ob, R, 8.0, 2, NF, chunk_size_grid=0)

    self = types.SimpleNamespace(
        mask_images=False, mask_depths=False, mask_threshold=0.5, bg_color=(1.0, 1.0, 1.0), n_train_target_views=1,
        sampling_mode_training="mask_sample", sampling_mode_evaluation="full_grid", view_pooler_enabled=False,
        net_3d_enabled=True, diffusion_enabled=True, net_3d=net_3d, diffusion=None, feature_size=C,
        _implicit_functions=ifs, raysampler=raysampler, _render=_render, render_image_height=HW, render_image_width=HW,
        output_rasterized_mc=False, view_metrics=lambda **kw: {}, regularization_metrics=lambda **kw: {},
        _get_objective=lambda preds: None, parameters=lambda: [])
    cams = ro.simple_360_cameras(8)[[5, 1, 2]]            # a batch of 3 cameras: only the first is a target
    grid = make_grid(C, R, seed=3)
    with torch.no_grad():
        preds = forward(self, camera=cams, voxel_features=grid, evaluation_mode=EvaluationMode.EVALUATION, image_rgb=None,
                        fg_probability=None, mask_crop=None, depth_map=None, sequence_name=None, frame_timestamp=None,
                        frame_number=None, sequence_category=None)
    out = {"grid": grid.numpy(), "cam_R": cams.R.numpy(), "cam_T": cams.T.numpy(), "cam_focal": cams.focal.numpy(),
           "cam_pp": cams.pp.numpy(), "images_render": preds["images_render"].numpy(),
           "depths_render": preds["depths_render"].numpy(), "masks_render": preds["masks_render"].numpy(),
           "preds_keys": np.array(sorted(preds.keys())), "log": np.array([repr(x) for x in log])}
    # the range assert fires on an out-of-range grid (holo_diffusion_model.py:381)
    try:
        forward(self, camera=cams, voxel_features=grid * 3.0, evaluation_mode=EvaluationMode.EVALUATION)
        out["range_assert"] = np.array(False)
    except AssertionError:
        out["range_assert"] = np.array(True)
    return out


if __name__ == "__main__":
    o = generate()
    np.savez_compressed(OUT, **o)
    print(f"wrote {OUT}: {os.path.getsize(OUT)} bytes; keys {list(o['preds_keys'])}; log {list(o['log'])}")
