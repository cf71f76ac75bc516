"""``HoloDiffusionModel`` under the reference's module path (/root/reference/holo_diffusion/holo_diffusion_model.py:44-634).

With the Implicitron config system present this is a registered model class with the reference's fields, assembled
from the registry's plug-ins (``net_3d_class_type`` -> ``SimpleUnet3D`` facade, ``renderer_class_type`` ->
``HoloMultiPassEmissionAbsorptionRenderer`` facade, ``implicit_function_class_type`` -> ``HoloVoxelGridImplicitFunction``
facade) exactly where the reference assembles its own (``create_net_3d`` :113-126, ``create_diffusion`` :128-132,
``_construct_implicit_functions`` :134-169); ``forward`` keeps the keyword-only signature (:201-214) and runs the
sampling / evaluation branch (:376-457) on the B200 kernels.  Base class: pytorch3d's ``GenericModel`` when the real
library is installed (all its config fields then exist, so the shipped yaml files validate), else a small stand-in
that declares the fields this path reads.
"""
from dataclasses import field
from typing import Any, Dict, List, Optional

import torch

import holo_diffusion_b200 as _b200

from ._plugin import HAVE_CONFIG, Configurable, adopt, plain, registry

# populate the registry like the reference does (:33-39)
from .holo_multipass_ea import HoloMultiPassEmissionAbsorptionRenderer  # noqa: F401
from .holo_voxel_grid_implicit_function import HoloVoxelGridImplicitFunction  # noqa: F401
from .utils.diffusion_utils import ImplicitronGaussianDiffusion, Unet3DBase

if HAVE_CONFIG:
    from pytorch3d.implicitron.models.implicit_function.base import ImplicitFunctionBase
    from pytorch3d.implicitron.models.renderer.base import BaseRenderer

    try:
        from pytorch3d.implicitron.models.renderer.base import ImplicitFunctionWrapper
    except ImportError:   # the test stand-in has none
        ImplicitFunctionWrapper = _b200.ImplicitFunctionWrapper

    try:
        from pytorch3d.implicitron.models.generic_model import GenericModel as _Base
        _GENERIC = True
    except ImportError:
        _GENERIC = False

        class _Base(Configurable, torch.nn.Module):
            """The GenericModel fields this path reads (pytorch3d 0.7.4 defaults, configs/base.yaml overrides them)."""
            render_image_width: int = 400
            render_image_height: int = 400
            chunk_size_grid: int = 4096
            raysampler_class_type: str = "AdaptiveRaySampler"
            renderer_class_type: str = "MultiPassEmissionAbsorptionRenderer"
            implicit_function_class_type: str = "NeuralRadianceFieldImplicitFunction"
            view_pooler_enabled: bool = False
            image_feature_extractor_class_type: Optional[str] = None
            mask_images: bool = True
            mask_depths: bool = True
            mask_threshold: float = 0.5
            bg_color: tuple = (0.0, 0.0, 0.0)

            def __post_init__(self):
                rs = plain(getattr(self, f"raysampler_{self.raysampler_class_type}_args", {}) or {})
                if self.raysampler_class_type != "AdaptiveRaySampler":
                    raise NotImplementedError("only AdaptiveRaySampler (configs/base.yaml:120)")
                self.raysampler = _b200.AdaptiveRaySampler(image_width=self.render_image_width,
                                                           image_height=self.render_image_height, **rs)
                self.renderer = registry.get(BaseRenderer, self.renderer_class_type)(
                    **plain(getattr(self, f"renderer_{self.renderer_class_type}_args", {}) or {}))
                self.create_net_3d()
                self.create_diffusion()
                self._implicit_functions = self._construct_implicit_functions()
                self.image_feature_extractor = self.view_pooler = None
                if self.view_pooler_enabled:   # GenericModel.create_image_feature_extractor / create_view_pooler
                    if self.image_feature_extractor_class_type is not None:
                        if self.image_feature_extractor_class_type != "ResNetFeatureExtractor":
                            raise NotImplementedError(self.image_feature_extractor_class_type)
                        self.image_feature_extractor = _b200.encoder.ResNetFeatureExtractor(**plain(getattr(
                            self, "image_feature_extractor_ResNetFeatureExtractor_args", {}) or {}))
                    self.view_pooler = _b200.encoder.ViewPooler(**plain(getattr(self, "view_pooler_args", {}) or {}))

    @registry.register
    class HoloDiffusionModel(_Base):
        resol: int = 32
        volume_extent: float = 8.0
        feature_size: int = 128
        num_passes: int = 2
        net_3d_enabled: bool = True
        net_3d: Optional[Unet3DBase]
        net_3d_class_type: str = "SimpleUnet3D"
        diffusion_enabled: bool = True
        diffusion: ImplicitronGaussianDiffusion
        enable_bootstrap: bool = True
        bootstrap_prob: float = 0.5
        loss_weights: Dict[str, float] = field(default_factory=lambda: {
            "loss_rgb_mse": 1.0, "loss_prev_stage_rgb_mse": 1.0, "loss_mask_bce": 0.0, "loss_prev_stage_mask_bce": 0.0})
        log_vars: List[str] = field(default_factory=lambda: ["loss_rgb_psnr", "loss_rgb_mse", "loss_prev_stage_rgb_mse",
                                                             "loss_prev_stage_rgb_psnr", "objective", "epoch", "sec/it"])
        use_cuda_graph: bool = True

        def __post_init__(self):
            super().__post_init__()
            rs = self.raysampler
            if not isinstance(rs, _b200.AdaptiveRaySampler):   # pytorch3d's sampler object: same fields, our kernel
                rs = _b200.AdaptiveRaySampler(
                    image_width=self.render_image_width, image_height=self.render_image_height,
                    n_pts_per_ray_evaluation=rs.n_pts_per_ray_evaluation, scene_extent=rs.scene_extent,
                    scene_center=tuple(rs.scene_center))
            core = _b200.HoloDiffusionModel.from_parts(
                resol=self.resol, volume_extent=self.volume_extent, feature_size=self.feature_size, net_3d=self.net_3d,
                diffusion=self.diffusion, raysampler=rs, renderer=self.renderer,
                implicit_functions=self._implicit_functions, render_image_width=self.render_image_width,
                render_image_height=self.render_image_height, chunk_size_grid=self.chunk_size_grid,
                use_cuda_graph=self.use_cuda_graph,
                image_feature_extractor=getattr(self, "image_feature_extractor", None),
                view_pooler=getattr(self, "view_pooler", None) if getattr(self, "view_pooler_enabled", False) else None,
                mask_images=getattr(self, "mask_images", True), mask_threshold=getattr(self, "mask_threshold", 0.5),
                bg_color=tuple(getattr(self, "bg_color", (0.0, 0.0, 0.0))))
            if core.pooled_feature_mapper is not None:   # :113, registered on the facade under the reference's name
                self.pooled_feature_mapper = core.pooled_feature_mapper
            if self.net_3d is not None:   # the sampling loop replays the denoiser as one CUDA graph too
                _b200.renderer.impl_of(self.net_3d)._exec.use_cuda_graph = self.use_cuda_graph
            adopt(self, core)   # children (net_3d, renderer, _implicit_functions) are already registered on the facade

        def create_net_3d(self):
            self.net_3d = None
            if self.net_3d_enabled:
                args = dict(getattr(self, "net_3d_" + self.net_3d_class_type + "_args", {}) or {})
                args.update(in_channels=self.feature_size, out_channels=self.feature_size, image_size=self.resol)
                self.net_3d = registry.get(Unet3DBase, self.net_3d_class_type)(**args)

        def create_diffusion(self):
            self.diffusion = None
            if self.diffusion_enabled:
                self.diffusion = ImplicitronGaussianDiffusion(**dict(getattr(self, "diffusion_args", {}) or {}))

        def _construct_implicit_functions(self):
            if self.implicit_function_class_type != "HoloVoxelGridImplicitFunction":
                raise ValueError(f"{str(type(self))} supports only HoloVoxelGridImplicitFunction!")
            name = f"implicit_function_{self.implicit_function_class_type}_args"
            config = getattr(self, name, None)
            if config is None:
                raise ValueError(f"{name} not present")
            args = dict(config)
            args.update(resol=self.resol, volume_extent=self.volume_extent, n_hidden=self.feature_size, feature_dim=0)
            fn = ImplicitFunctionWrapper(registry.get(ImplicitFunctionBase, self.implicit_function_class_type)(**args))
            return torch.nn.ModuleList([fn for _ in range(self.num_passes)])   # ONE RenderMLP shared by all passes

        def sample_random_voxel_features_progressive(self):
            return self._impl.sample_random_voxel_features_progressive()

        def sample_random_voxel_features(self) -> torch.Tensor:
            return self._impl.sample_random_voxel_features()

        def forward(self, *, image_rgb: Optional[torch.Tensor] = None, camera, fg_probability=None, mask_crop=None,
                    depth_map=None, sequence_name=None, frame_timestamp=None, evaluation_mode=None,
                    voxel_features: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, Any]:
            mode = _b200.EvaluationMode.EVALUATION if evaluation_mode is None else evaluation_mode
            return self._impl.forward(image_rgb=image_rgb, camera=camera, fg_probability=fg_probability,
                                      mask_crop=mask_crop, depth_map=depth_map, sequence_name=sequence_name,
                                      frame_timestamp=frame_timestamp, evaluation_mode=mode,
                                      voxel_features=voxel_features, **kwargs)

else:
    HoloDiffusionModel = _b200.HoloDiffusionModel
