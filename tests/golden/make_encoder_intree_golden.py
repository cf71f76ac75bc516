"""Golden vectors for the view-pooling encoder from the reference's IN-TREE code, executed UNMODIFIED from
/root/reference over the pytorch3d stand-in under ``oracle/pt3d_stub`` (pytorch3d itself is not installable here):

    python tests/golden/make_encoder_intree_golden.py     # writes tests/golden/encoder_intree_ref.npz

What runs is the reference's own
  * ``MLPMeanFeatureAggregator`` (custom_modules.py:162-281) -- lazy first layers, the in-tree ``MLPWithInputSkips`` with
    its activation placement, ``_last``, the softmax-weighted sum over the views -- for n_layers = 1 (checkpointed) and 2;
  * ``_get_point_to_source_camera_ray_dirs`` (custom_modules.py:283-334);
  * the encoder branch of ``HoloDiffusionModel.forward`` (holo_diffusion_model.py:248-373: source-view selection,
    feature extractor / view pooler call arguments, ``pooled_feature_mapper``, permute + reshape, tanh), its SOURCE
    executed on a stand-in ``self`` like tests/golden/make_model_forward_intree_golden.py does.
The vectors pin the oracle's restatement of THAT logic (oracle/encoder_oracle.py).  The pytorch3d leaves underneath
(view sampling, wmean, the cartesian product, the harmonic embedding, VolumeLocator, preprocess_input) are the stub's
restatements from memory, so the leaf arithmetic stays unpinned.
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
OUT = os.path.join(HERE, "encoder_intree_ref.npz")
R, C_GRID, EXTENT = 4, 8, 8.0


def generate():
    for p in (ROOT, REF, os.path.join(ROOT, "oracle", "pt3d_stub")):   # REF ahead of ROOT: holo_diffusion = the reference
        if p not in sys.path:
            sys.path.insert(0, p)
    from holo_diffusion.custom_modules import (LazyLinearWithXavierInit, MLPMeanFeatureAggregator,
                                               _get_point_to_source_camera_ray_dirs)
    from pytorch3d.implicitron.models.renderer.base import EvaluationMode, ImplicitronRayBundle

    from oracle import encoder_oracle as eo
    from oracle import render_oracle as ro

    out = {}
    g = torch.Generator().manual_seed(23)

    def cam_arrays(prefix, cams):
        out.update({prefix + "R": cams.R.numpy(), prefix + "T": cams.T.numpy(), prefix + "focal": cams.focal.numpy(),
                    prefix + "pp": cams.pp.numpy()})

    # ---- the aggregator on given sampled features
    cams, feats, mask_crop = eo.make_views(4, (24, 32), stage_channels=(4, 4), seed=3)
    pts = (torch.rand(24, 3, generator=g) * 2 - 1) * 4.0
    fs, ms = eo.sample_views(cams, pts, feats, mask_crop, masked_sampling=True)
    cam_arrays("agg/cam_", cams)
    out["agg/pts"] = pts.numpy()
    out["agg/feat_keys"] = np.array(list(fs.keys()))
    for k, v in fs.items():
        out["agg/feats/" + k] = v.numpy()
    out["agg/masks"] = ms.numpy()
    out["agg/ray_dirs"] = _get_point_to_source_camera_ray_dirs(cams, pts[None]).numpy()
    for n_layers, ckpt in ((1, True), (2, False)):
        torch.manual_seed(100 + n_layers)
        agg = MLPMeanFeatureAggregator(n_hidden=16, dim_out=8, n_layers=n_layers, checkpointed_mlp=ckpt)
        agg.exclude_target_view = False                      # holo_diffusion_model.py:115-116
        agg.exclude_target_view_mask_features = False
        with torch.no_grad():
            agg(fs, ms, camera=cams, pts=pts[None])          # materialises the lazy layers (zero biases)
            for k, v in agg.state_dict().items():
                if k.endswith("bias"):
                    v.uniform_(-0.2, 0.2)
            y = agg(fs, ms, camera=cams, pts=pts[None])
        tag = f"agg{n_layers}/"
        out[tag + "sd_keys"] = np.array(list(agg.state_dict().keys()))
        for k, v in agg.state_dict().items():
            out[tag + "sd/" + k] = v.numpy()
        out[tag + "out"] = y.numpy()

    # ---- the encoder branch of HoloDiffusionModel.forward
    src = open(os.path.join(REF, "holo_diffusion/holo_diffusion_model.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "HoloDiffusionModel")
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")

    class RenderSamplingMode(enum.Enum):
        MASK_SAMPLE = "mask_sample"
        FULL_GRID = "full_grid"

    class ImplicitronRender:
        def __init__(self, image_render=None, depth_render=None, mask_render=None):
            self.image_render, self.depth_render, self.mask_render = image_render, depth_render, mask_render

    log = []

    def preprocess_input(image_rgb, fg_probability, depth_map, mask_images, mask_depths, mask_threshold, bg_color):
        """pytorch3d implicitron.models.utils.preprocess_input [pt3d-recalled]: threshold the foreground probability and
        paint the background."""
        log.append(("preprocess", bool(mask_images), float(mask_threshold), tuple(bg_color)))
        if mask_images and fg_probability is not None and image_rgb is not None:
            fg = (fg_probability > mask_threshold).to(image_rgb)
            image_rgb = fg * image_rgb + (1 - fg) * image_rgb.new_tensor(bg_color).view(1, 3, 1, 1)
        return image_rgb, fg_probability, depth_map

    class VolumeLocator:   # pytorch3d.structures.volumes.VolumeLocator [pt3d-recalled]
        def __init__(self, batch_size, grid_sizes, device=None, voxel_size=1.0):
            log.append(("locator", int(batch_size), tuple(grid_sizes), float(voxel_size)))
            self.n, self.voxel = grid_sizes[0], voxel_size

        def get_coord_grid(self):
            return eo.coord_grid(self.n, self.voxel * self.n).reshape(1, self.n, self.n, self.n, 3)

    ns = {"torch": torch, "np": np, "logger": logging.getLogger("ref"), "Optional": Optional, "Union": Union, "List": List,
          "Dict": Dict, "Any": Any, "CamerasBase": object, "EvaluationMode": EvaluationMode,
          "RenderSamplingMode": RenderSamplingMode, "ImplicitronRayBundle": ImplicitronRayBundle,
          "ImplicitronRender": ImplicitronRender, "preprocess_input": preprocess_input, "VolumeLocator": VolumeLocator,
          "rasterize_sparse_ray_bundle": None}
    exec(compile(ast.Module(body=[fwd], type_ignores=[]), "holo_diffusion_model.py", "exec"), ns)
    forward = ns["forward"]

    B, HW = 5, (24, 32)
    cams5, feats5, mask5 = eo.make_views(B, HW, stage_channels=(4, 4), seed=9)
    image_rgb = torch.rand(B, 3, *HW, generator=g)
    fg = torch.rand(B, 1, *HW, generator=g)
    names = ["seq_a", "seq_a", "seq_b", "seq_a", "seq_a"]    # view 2 belongs to another sequence: not a source
    torch.manual_seed(77)
    agg = MLPMeanFeatureAggregator(n_hidden=16, dim_out=8, n_layers=1, checkpointed_mlp=False)
    mapper = LazyLinearWithXavierInit(C_GRID)

    def image_feature_extractor(imgs, masks):
        """Stand-in extractor: fixed feature maps of the SELECTED views + (as pytorch3d's does) the masks and images."""
        log.append(("extractor", tuple(imgs.shape), tuple(masks.shape)))
        sel = [i for i in range(B) if any(torch.equal(imgs[j], painted[i]) for j in range(imgs.shape[0]))]
        log.append(("extractor_views", tuple(sel)))
        f = {k: v[sel] for k, v in feats5.items() if k.startswith("res_layer")}
        f["mask"], f["image"] = masks, imgs
        return f

    class Pooler:
        feature_aggregator = agg
        calls = []

        def __call__(self, *, pts, seq_id_pts, camera, seq_id_camera, feats, masks):
            log.append(("pooler", tuple(pts.shape), list(seq_id_pts), list(seq_id_camera), int(camera.R.shape[0]),
                        tuple(masks.shape), list(feats.keys())))
            Pooler.calls.append({"cam": camera, "feats": feats, "masks": masks})
            vw = torch.tensor([1.0 if s == seq_id_pts[0] else 0.0 for s in seq_id_camera])
            fs_, ms_ = eo.sample_views(camera, pts[0], feats, masks, masked_sampling=False, view_weight=vw)
            return agg(fs_, ms_, camera=camera, pts=pts)

    class Bound:
        def __init__(self):
            self.bound = None

        def bind_args(self, **kw):
            self.bound = kw

        def unbind_args(self):
            pass

    ifs = [Bound(), Bound()]
    painted = preprocess_input(image_rgb, fg, None, True, True, 0.5, (0.0, 0.0, 0.0))[0]
    log.clear()

    def raysampler(cameras, evaluation_mode, mask=None):
        log.append(("rays", int(cameras.R.shape[0])))
        return None

    def _render(**kw):
        z = torch.zeros(1, 2, 2, 3)
        return types.SimpleNamespace(features=z, depths=z[..., :1], masks=z[..., :1])

    agg.exclude_target_view = True   # forward's asserts require what __post_init__ sets (:115-116)
    self = types.SimpleNamespace(
        mask_images=True, mask_depths=True, mask_threshold=0.5, bg_color=(0.0, 0.0, 0.0), n_train_target_views=1,
        sampling_mode_training="mask_sample", sampling_mode_evaluation="full_grid", view_pooler_enabled=True,
        image_feature_extractor=image_feature_extractor, view_pooler=Pooler(), pooled_feature_mapper=mapper,
        resol=R, volume_extent=EXTENT, net_3d_enabled=False, diffusion_enabled=False, net_3d=None, diffusion=None,
        feature_size=C_GRID, _implicit_functions=ifs, raysampler=raysampler, _render=_render, render_image_height=2,
        render_image_width=2, output_rasterized_mc=False, view_metrics=lambda **kw: {},
        regularization_metrics=lambda **kw: {}, _get_objective=lambda preds: None, parameters=lambda: [])
    kw = dict(camera=cams5, image_rgb=image_rgb, fg_probability=fg, mask_crop=mask5, depth_map=None, sequence_name=names,
              frame_timestamp=None, evaluation_mode=EvaluationMode.EVALUATION)
    try:
        forward(self, **kw)
        out["fwd/exclusion_assert"] = np.array(False)
    except AssertionError:
        out["fwd/exclusion_assert"] = np.array(True)
    agg.exclude_target_view = False
    agg.exclude_target_view_mask_features = False
    log.clear()
    with torch.no_grad():
        forward(self, **kw)
    out["fwd/grid"] = ifs[0].bound["voxel_grid_features"].numpy()
    out["fwd/log"] = np.array([repr(x) for x in log])
    out["fwd/image_rgb"], out["fwd/fg"], out["fwd/mask_crop"] = image_rgb.numpy(), fg.numpy(), mask5.numpy()
    out["fwd/names"] = np.array(names)
    cam_arrays("fwd/cam_", cams5)
    for k, v in feats5.items():
        if k.startswith("res_layer"):
            out["fwd/feats/" + k] = v.numpy()
    out["fwd/agg_keys"] = np.array(list(agg.state_dict().keys()))
    for k, v in agg.state_dict().items():
        out["fwd/agg/" + k] = v.detach().numpy()
    out["fwd/mapper_w"], out["fwd/mapper_b"] = mapper.weight.detach().numpy(), mapper.bias.detach().numpy()
    return out


if __name__ == "__main__":
    o = generate()
    np.savez_compressed(OUT, **o)
    print(f"wrote {OUT}: {os.path.getsize(OUT)} bytes; forward log:")
    for line in o["fwd/log"]:
        print("   ", line)
