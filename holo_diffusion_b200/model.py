"""``HoloDiffusionModel`` -- the sampling / evaluation branch of the reference model on the CUDA kernels.

Mirrors /root/reference/holo_diffusion/holo_diffusion_model.py: constructor fields (:47-75 + GenericModel's
render_image_width/height, raysampler/renderer argument groups), ``sample_random_voxel_features[_progressive]``
:173-199 and the keyword-only ``forward`` :201-214 for the path ``generate_samples.py`` drives
(image_rgb=None, voxel_features given, evaluation_mode=EVALUATION):
    asserts on the grid range :381 -> ``voxel_features = tanh(net_3d(voxel_features, t=0))`` :420-428 ->
    bind ``voxel_grid_features`` :431-438 -> ray sampler :442-448 -> ``_render`` :451-457 -> preds :469-523.
The view-pooling encoder (image_rgb given: :327-373) runs on the kernels of encoder.py and feeds the same path.  The
training branch (losses; it needs the UNet kernels' backward) is not built and raises; the renderer is differentiable
on its own (autograd.py).
Parameter names follow the reference so that checkpoints load: ``net_3d._net.*``,
``_implicit_functions.{i}._fn.render_mlp.*``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .cameras import AdaptiveRaySampler, ImplicitronRayBundle, PerspectiveCameras
from .diffusion import ImplicitronGaussianDiffusion
from . import encoder as _enc
from .renderer import (EvaluationMode, HoloMultiPassEmissionAbsorptionRenderer, HoloVoxelGridImplicitFunction,
                       ImplicitFunctionWrapper, RendererOutput, coerce_mode, impl_of)
from .unet import SimpleUnet3D


@dataclass
class ImplicitronRender:
    image_render: Optional[torch.Tensor] = None
    depth_render: Optional[torch.Tensor] = None
    mask_render: Optional[torch.Tensor] = None


def _cat_render_outputs(outs: List[Optional[RendererOutput]], B: int, spatial) -> Optional[RendererOutput]:
    """apply_chunked's collation: cat along the ray dimension, restore (B, *spatial, -1), recurse into prev_stage."""
    if outs[0] is None:
        return None

    def cat(ts):
        return None if ts[0] is None else torch.cat(ts, dim=1).reshape(B, *spatial, -1)

    aux = {k: (cat([o.aux[k] for o in outs]) if torch.is_tensor(v) else v) for k, v in outs[0].aux.items()}
    return RendererOutput(features=cat([o.features for o in outs]), depths=cat([o.depths for o in outs]),
                          masks=cat([o.masks for o in outs]), normals=cat([o.normals for o in outs]),
                          points=cat([o.points for o in outs]), weights=cat([o.weights for o in outs]), aux=aux,
                          prev_stage=_cat_render_outputs([o.prev_stage for o in outs], B, spatial))


def _clone_render_output(o: Optional[RendererOutput]) -> Optional[RendererOutput]:
    if o is None:
        return None

    def c(t):
        return t.clone() if torch.is_tensor(t) else t

    return RendererOutput(features=c(o.features), depths=c(o.depths), masks=c(o.masks), normals=c(o.normals),
                          points=c(o.points), weights=c(o.weights), aux=dict(o.aux),
                          prev_stage=_clone_render_output(o.prev_stage))


class HoloDiffusionModel(nn.Module):
    def __init__(self, resol: int = 16, volume_extent: float = 8.0, feature_size: int = 64, num_passes: int = 2,
                 render_image_width: int = 256, render_image_height: int = 256, net_3d_enabled: bool = True,
                 net_3d_class_type: str = "SimpleUnet3D", net_3d_SimpleUnet3D_args: Optional[dict] = None,
                 diffusion_enabled: bool = True, diffusion_args: Optional[dict] = None,
                 raysampler_class_type: str = "AdaptiveRaySampler", raysampler_AdaptiveRaySampler_args: Optional[dict] = None,
                 renderer_class_type: str = "HoloMultiPassEmissionAbsorptionRenderer",
                 renderer_HoloMultiPassEmissionAbsorptionRenderer_args: Optional[dict] = None,
                 implicit_function_class_type: str = "HoloVoxelGridImplicitFunction",
                 implicit_function_HoloVoxelGridImplicitFunction_args: Optional[dict] = None,
                 chunk_size_grid: int = 4096, use_cuda_graph: bool = True, view_pooler_enabled: bool = False,
                 image_feature_extractor_class_type: Optional[str] = None,
                 image_feature_extractor_ResNetFeatureExtractor_args: Optional[dict] = None,
                 view_pooler_args: Optional[dict] = None, mask_images: bool = True, mask_threshold: float = 0.5,
                 bg_color=(0.0, 0.0, 0.0), **unused):
        super().__init__()
        if implicit_function_class_type != "HoloVoxelGridImplicitFunction":
            raise ValueError(f"{type(self)} supports only HoloVoxelGridImplicitFunction!")
        if net_3d_class_type != "SimpleUnet3D" or raysampler_class_type != "AdaptiveRaySampler" or \
                renderer_class_type != "HoloMultiPassEmissionAbsorptionRenderer":
            raise NotImplementedError("plug-in type not built")
        self.resol, self.volume_extent, self.feature_size, self.num_passes = resol, volume_extent, feature_size, num_passes
        self.render_image_width, self.render_image_height = render_image_width, render_image_height
        self.chunk_size_grid = chunk_size_grid
        self.net_3d_enabled, self.diffusion_enabled = net_3d_enabled, diffusion_enabled
        self.net_3d = None
        if net_3d_enabled:
            a = dict(net_3d_SimpleUnet3D_args or {})
            a.update(in_channels=feature_size, out_channels=feature_size, image_size=resol)
            a.setdefault("use_cuda_graph", use_cuda_graph)  # the sampling loop replays the denoiser as a graph too
            self.net_3d = SimpleUnet3D(**a)
        self.diffusion = ImplicitronGaussianDiffusion(**(diffusion_args or {})) if diffusion_enabled else None
        rs = dict(raysampler_AdaptiveRaySampler_args or {})
        self.raysampler = AdaptiveRaySampler(image_width=render_image_width, image_height=render_image_height, **rs)
        self.renderer = HoloMultiPassEmissionAbsorptionRenderer(**(renderer_HoloMultiPassEmissionAbsorptionRenderer_args or {}))
        ia = dict(implicit_function_HoloVoxelGridImplicitFunction_args or {})
        ia.update(resol=resol, volume_extent=volume_extent, n_hidden=feature_size, feature_dim=0)
        wrapper = ImplicitFunctionWrapper(HoloVoxelGridImplicitFunction(**ia))
        self._implicit_functions = nn.ModuleList([wrapper for _ in range(num_passes)])
        self._range_stats: Optional[torch.Tensor] = None
        self._t0: Optional[torch.Tensor] = None
        # one CUDA graph per (device, grid shape, weights version); its static output buffers are cloned on return
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_key = None
        self._sample_group = None   # set by shard_one_sample()
        # view-pooling encoder (holo_diffusion_model.py:111-116 + GenericModel's image_feature_extractor / view_pooler)
        self.view_pooler_enabled = view_pooler_enabled
        self.mask_images, self.mask_threshold, self.bg_color = mask_images, mask_threshold, tuple(bg_color)
        self.image_feature_extractor = self.view_pooler = self.pooled_feature_mapper = None
        if view_pooler_enabled:
            if image_feature_extractor_class_type not in (None, "ResNetFeatureExtractor"):
                raise NotImplementedError(f"image feature extractor {image_feature_extractor_class_type}")
            if image_feature_extractor_class_type is not None:
                self.image_feature_extractor = _enc.ResNetFeatureExtractor(
                    **(image_feature_extractor_ResNetFeatureExtractor_args or {}))
            self.view_pooler = _enc.ViewPooler(**(view_pooler_args or {}))
            self._init_encoder()

    def _init_encoder(self):
        self.pooled_feature_mapper = _enc.LazyLinearWithXavierInit(self.feature_size)       # :113
        self.view_pooler.feature_aggregator.exclude_target_view = False                     # :115 ("by hard")
        self.view_pooler.feature_aggregator.exclude_target_view_mask_features = False       # :116

    @classmethod
    def from_parts(cls, *, resol: int, volume_extent: float, feature_size: int, net_3d, diffusion, raysampler, renderer,
                   implicit_functions, render_image_width: int, render_image_height: int, chunk_size_grid: int = 4096,
                   use_cuda_graph: bool = True, image_feature_extractor=None, view_pooler=None, mask_images: bool = True,
                   mask_threshold: float = 0.5, bg_color=(0.0, 0.0, 0.0)) -> "HoloDiffusionModel":
        """Assemble the model from plug-ins that something else constructed (the Implicitron config system through the
        ``holo_diffusion`` shim: registry facades of SimpleUnet3D / the renderer / the implicit function, wrapped in
        pytorch3d's own ImplicitFunctionWrapper).  Same forward path as the plain constructor."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.resol, self.volume_extent, self.feature_size = resol, volume_extent, feature_size
        self.num_passes = len(implicit_functions)
        self.render_image_width, self.render_image_height = render_image_width, render_image_height
        self.chunk_size_grid = chunk_size_grid
        self.net_3d_enabled, self.diffusion_enabled = net_3d is not None, diffusion is not None
        self.net_3d, self.diffusion, self.raysampler, self.renderer = net_3d, diffusion, raysampler, renderer
        self._implicit_functions = implicit_functions if isinstance(implicit_functions, nn.ModuleList) else \
            nn.ModuleList(list(implicit_functions))
        self._range_stats = self._t0 = None
        self.use_cuda_graph = use_cuda_graph
        self._graph = self._graph_key = self._sample_group = None
        self.view_pooler_enabled = view_pooler is not None
        self.mask_images, self.mask_threshold, self.bg_color = mask_images, mask_threshold, tuple(bg_color)
        self.image_feature_extractor, self.view_pooler, self.pooled_feature_mapper = image_feature_extractor, view_pooler, None
        if view_pooler is not None:
            self._init_encoder()
        return self

    def shard_one_sample(self, group=None, attn_min_tokens: int = 1 << 14):
        """Several GPUs cooperate on ONE grid and ONE view (BASELINE cfg #5): every rank calls
        forward() with the same inputs; the denoiser's large attention blocks split their queries over the ranks
        (one all-gather per block, SimpleUnet3D.shard_attention), each rank renders a contiguous block of image rows
        of the same grid, and the row blocks of features / depths / masks (/ weights / normals) are all-gathered, so
        every rank returns the full images.  prev_stage is not gathered (None).  Replays no CUDA graph."""
        import torch.distributed as dist
        self._sample_group = group if group is not None else dist.group.WORLD
        if self.net_3d is not None:
            self.net_3d.shard_attention(self._sample_group, attn_min_tokens)
        self.use_cuda_graph = False

    # ------------------------------------------------------------------ sampling
    def sample_random_voxel_features_progressive(self):
        assert self.net_3d_enabled and self.diffusion_enabled
        for sample in self.diffusion.p_sample_loop_progressive(
                model=self.net_3d, shape=(1, self.feature_size, self.resol, self.resol, self.resol), clip_denoised=True,
                progress=False):
            yield torch.clip(sample["sample"], -1.0, 1.0)

    def sample_random_voxel_features(self) -> torch.Tensor:
        assert self.net_3d_enabled and self.diffusion_enabled
        return self.diffusion.p_sample_loop(model=self.net_3d,
                                            shape=(1, self.feature_size, self.resol, self.resol, self.resol),
                                            clip_denoised=True, progress=False)

    # ------------------------------------------------------------------ range asserts without a sync per assert
    def _check_range(self, where: str):
        mn, mx, nan = ops.decode_range(self._range_stats.cpu().tolist())
        assert nan == 0 and mn >= -1.0 and mx <= 1.0, f"voxel features out of [-1, 1] ({where}): min {mn} max {mx} nan {nan}"

    # ------------------------------------------------------------------ device work of one view (no host sync)
    def _device_forward(self, cam: PerspectiveCameras, voxel_features: torch.Tensor):
        dev = voxel_features.device
        C, R = self.feature_size, self.resol
        V = R ** 3
        x_cl = ops.transpose2d(voxel_features.reshape(-1), C, V).view(V, C)
        ops.range_init(self._range_stats)
        if self.net_3d_enabled:
            # voxel_features = tanh(net_3d(voxel_features, t=0)) (:420-425); the input-range assert (:381) and the
            # output-range asserts (:426,:428) come from fused min/max passes and are checked once, after the
            # whole view has been queued
            ops.act_range(x_cl, V, C, 0, None, None, self._range_stats)
            y_cl = impl_of(self.net_3d)._exec.forward_cl(x_cl, (R, R, R), self._t0)
            grid_cl = torch.empty(V, C, device=dev)
            grid_cf = torch.empty(C * V, device=dev)
            ops.act_range(y_cl, V, C, 1, grid_cl, grid_cf, self._range_stats)
            voxel_features = grid_cf.view(1, C, R, R, R)
        else:
            grid_cl = torch.empty(V, C, device=dev)
            ops.act_range(x_cl, V, C, 0, grid_cl, None, self._range_stats)
        for func in self._implicit_functions:
            func.bind_args(voxel_grid_features=voxel_features, voxel_grid_features_channels_last=grid_cl.view(R, R, R, C))
        ray_bundle = self.raysampler(cam, EvaluationMode.EVALUATION)
        if self._sample_group is not None:
            rendered = self._render_row_sharded(ray_bundle)
        else:
            rendered = self._render(ray_bundle=ray_bundle, chunksize=self.chunk_size_grid,
                                    implicit_functions=list(self._implicit_functions),
                                    evaluation_mode=EvaluationMode.EVALUATION)
        for func in self._implicit_functions:
            func.unbind_args()
        return rendered, ray_bundle, voxel_features

    def _render_row_sharded(self, ray_bundle: ImplicitronRayBundle) -> RendererOutput:
        """Render this rank's block of image rows, then all-gather the image-shaped outputs (shard_one_sample)."""
        import torch.distributed as dist
        from .sharding import gather_rows, row_shard
        world, rank = dist.get_world_size(self._sample_group), dist.get_rank(self._sample_group)
        H = ray_bundle.lengths.shape[1]
        h0, h1, _ = row_shard(H, world, rank)
        sub = ImplicitronRayBundle(ray_bundle.origins[:, h0:h1].contiguous(), ray_bundle.directions[:, h0:h1].contiguous(),
                                   ray_bundle.lengths[:, h0:h1].contiguous(), ray_bundle.xys[:, h0:h1].contiguous())
        out = self._render(ray_bundle=sub, chunksize=self.chunk_size_grid,
                           implicit_functions=list(self._implicit_functions), evaluation_mode=EvaluationMode.EVALUATION)

        def g(t):
            return None if t is None else gather_rows(t.contiguous(), H, rank, world, self._sample_group)

        return RendererOutput(features=g(out.features), depths=g(out.depths), masks=g(out.masks), prev_stage=None,
                              normals=g(out.normals), points=None, weights=g(out.weights), aux={})

    # ------------------------------------------------------------------ GenericModel._render (chunked rendering)
    def _render(self, *, ray_bundle: ImplicitronRayBundle, chunksize: int, **kwargs) -> RendererOutput:
        """pytorch3d GenericModel._render + chunk_generator / apply_chunked (entered holo_diffusion_model.py:451-457):
        rays flattened to (B, n_rays), n_chunks = ceil(n_rays * S / chunksize), rays_per_chunk = ceil(n_rays /
        n_chunks), every RendererOutput field (and the prev_stage chain) concatenated along the ray dimension and
        reshaped to (B, *spatial, -1).  The fused kernel materialises nothing per point, so it takes all rays in
        one launch whatever the chunk size; chunking applies to the staged (per-plug-in) path."""
        if chunksize <= 0 or self.renderer.is_fused(kwargs["implicit_functions"], kwargs["evaluation_mode"]):
            return self.renderer(ray_bundle=ray_bundle, **kwargs)
        B, *spatial, S = ray_bundle.lengths.shape
        if S > 0 and chunksize % S != 0:
            raise ValueError(f"chunk_size_grid ({chunksize}) should be divisible by n_pts_per_ray ({S})")
        n = 1
        for d in spatial:
            n *= d
        n_chunks = -(-n * max(S, 1) // chunksize)
        per = -(-n // n_chunks)
        o, d, l, xy = (ray_bundle.origins.reshape(B, n, 3), ray_bundle.directions.reshape(B, n, 3),
                       ray_bundle.lengths.reshape(B, n, S), ray_bundle.xys.reshape(B, n, 2))
        outs = [self.renderer(ray_bundle=ImplicitronRayBundle(o[:, a:a + per], d[:, a:a + per], l[:, a:a + per],
                                                              xy[:, a:a + per]), **kwargs)
                for a in range(0, n, per)]
        return _cat_render_outputs(outs, B, spatial)

    def _weights_signature(self):
        ps = [p for m in (self.net_3d, self._implicit_functions) if m is not None for p in m.parameters()]
        return sum(p._version for p in ps), ps[0].data_ptr()

    def _graph_forward(self, cam: PerspectiveCameras, voxel_features: torch.Tensor):
        """Replay the whole view (UNet + tanh + rays + render, ~600 launches) as one CUDA graph with static I/O."""
        dev = voxel_features.device
        key = (str(dev), tuple(voxel_features.shape), self._weights_signature())
        if self._graph is None or self._graph_key != key:
            self._g_vox = torch.empty_like(voxel_features)
            self._g_cam = PerspectiveCameras(torch.ones(1, 2), torch.zeros(1, 2), torch.eye(3)[None], torch.zeros(1, 3)).to(dev)
            self._copy_inputs(cam, voxel_features)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._device_forward(self._g_cam, self._g_vox)  # warm-up: packs weights, sets kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._g_out = self._device_forward(self._g_cam, self._g_vox)
            self._graph, self._graph_key = g, key
        self._copy_inputs(cam, voxel_features)
        self._graph.replay()
        # The graph writes into static buffers that the next replay overwrites; the reference returns fresh tensors on
        # every call, and a caller that collects preds across views must not end up with N copies of the last frame:
        # the API outputs (images / depths / masks of every stage, the denoised grid) are cloned.  The ray bundle and
        # the per-ray sample depths (aux["lengths"], 40 MB together) stay views of the static buffers.
        rendered, bundle, vox = self._g_out
        if os.environ.get("HOLO_GRAPH_NO_CLONE") == "1":   # A/B measurements only
            return self._g_out
        return _clone_render_output(rendered), bundle, vox.clone()

    def _copy_inputs(self, cam, voxel_features):
        self._g_vox.copy_(voxel_features, non_blocking=True)
        self._g_cam.R.copy_(cam.R, non_blocking=True)
        self._g_cam.T.copy_(cam.T, non_blocking=True)
        self._g_cam.focal_length.copy_(cam.focal_length, non_blocking=True)
        self._g_cam.principal_point.copy_(cam.principal_point, non_blocking=True)

    # ------------------------------------------------------------------ views -> voxel grid
    @torch.no_grad()
    def encode_views(self, *, image_rgb: torch.Tensor, camera: PerspectiveCameras, fg_probability=None, mask_crop=None,
                     sequence_name=None, n_targets: int = 1) -> torch.Tensor:
        """The encoder branch of forward (holo_diffusion_model.py:248-257,327-373): (B, 3, H, W) views -> (1, C, R, R, R)
        grid in [-1, 1].  Source views = the views of the first view's sequence after the n_targets targets (all of
        them when none is left, :293-306)."""
        dev = image_rgb.device
        ops.require_cuda(dev, "HoloDiffusionModel.encode_views")
        B = camera.R.shape[0]
        if self.mask_images and fg_probability is not None:   # preprocess_input (:248-257)
            image_rgb = _enc.mask_background(image_rgb, fg_probability, self.mask_threshold, self.bg_color)
        if B <= n_targets:
            n_targets = 1
        sel = _enc.select_sources(sequence_name, B, n_targets) if B > 1 else [0]

        def src(t):
            return None if t is None else t[sel]

        feats = self.image_feature_extractor(src(image_rgb), src(fg_probability))
        C, R = self.feature_size, self.resol
        cached = self.__dict__.get("_grid_pts")
        if cached is None or cached[0] != (R, float(self.volume_extent), str(dev)):
            cached = self.__dict__["_grid_pts"] = ((R, float(self.volume_extent), str(dev)),
                                                   _enc.coord_grid(R, self.volume_extent, dev))
        pts = cached[1]
        cams = camera[sel]
        cams = PerspectiveCameras(cams.focal_length, cams.principal_point, cams.R, cams.T).to(dev)
        vw = None if sequence_name is None else _enc.view_weights(sequence_name[:1], [sequence_name[i] for i in sel], dev)
        rows = _enc.pool_views(self.view_pooler, pts, cams, feats, src(mask_crop), vw, mapper=self.pooled_feature_mapper)
        grid_cf = torch.empty(C * R ** 3, device=dev)
        ops.act_range(rows, R ** 3, C, 1, None, grid_cf, None)   # tanh (:373) + (V, C) rows -> (1, C, R, R, R)
        return grid_cf.view(1, C, R, R, R)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, *, image_rgb: Optional[torch.Tensor] = None, camera: PerspectiveCameras,
                fg_probability=None, mask_crop=None, depth_map=None, sequence_name=None, frame_timestamp=None,
                evaluation_mode: EvaluationMode = EvaluationMode.EVALUATION, voxel_features: Optional[torch.Tensor] = None,
                **kwargs) -> Dict[str, Any]:
        if coerce_mode(evaluation_mode) != EvaluationMode.EVALUATION:
            raise NotImplementedError("training branch is a 'next' row (SURVEY 8f)")
        target_cameras = camera[[0]]  # n_targets = 1 (holo_diffusion_model.py:263-273,315)
        if image_rgb is not None:
            # fmt: off
            assert self.view_pooler_enabled, "view_pooler must be enabled to use image_rgb"
            assert voxel_features is None, "Cannot provide both image_rgb and voxel_features"
            assert self.image_feature_extractor is not None, "Need an image_feature_extractor"
            # fmt: on
            voxel_features = self.encode_views(image_rgb=image_rgb, camera=camera, fg_probability=fg_probability,
                                               mask_crop=mask_crop, sequence_name=sequence_name)
        if voxel_features is None:
            voxel_features = self.sample_random_voxel_features()
        dev = voxel_features.device
        ops.require_cuda(dev, "HoloDiffusionModel")
        assert voxel_features.shape[0] == 1, "only one single voxel grid is supported per GPU"
        assert voxel_features.shape[1] == self.feature_size, "Wrong voxel feature size!"
        if self._range_stats is None or self._range_stats.device != dev:
            self._range_stats = torch.empty(4, dtype=torch.int32, device=dev)
            self._t0 = torch.zeros(1, dtype=torch.int64, device=dev)
        voxel_features = voxel_features.contiguous().float()
        if self.use_cuda_graph:
            rendered, ray_bundle, voxel_features = self._graph_forward(target_cameras, voxel_features)
        else:
            rendered, ray_bundle, voxel_features = self._device_forward(target_cameras.to(dev), voxel_features)
        self._check_range("input / tanh(net_3d) output")  # the only host sync: 16 bytes, after all work is queued
        preds: Dict[str, Any] = {"rendered": rendered, "ray_bundle": ray_bundle, "voxel_features": voxel_features}
        preds["images_render"] = rendered.features.permute(0, 3, 1, 2)
        preds["depths_render"] = rendered.depths.permute(0, 3, 1, 2)
        preds["masks_render"] = rendered.masks.permute(0, 3, 1, 2)
        preds["implicitron_render"] = ImplicitronRender(image_render=preds["images_render"],
                                                        depth_render=preds["depths_render"],
                                                        mask_render=preds["masks_render"])
        # no "objective" key: the reference adds it only when _get_objective() returns one (holo_diffusion_model.py:
        # 529-537), which it does not in evaluation without target images
        return preds
