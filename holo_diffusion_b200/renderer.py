"""Renderer plug-ins: the reference's ImplicitFunction / Renderer surface over the fused CUDA render kernel.

Mirrors (class names, constructor fields, parameter names):
  * ``RenderMLP`` / ``HoloVoxelGridImplicitFunction`` -- /root/reference/holo_diffusion/holo_voxel_grid_implicit_function.py:48-269
  * ``MLPWithInputSkips`` parameter layout          -- /root/reference/holo_diffusion/custom_modules.py:44-160
  * ``HoloMultiPassEmissionAbsorptionRenderer``     -- /root/reference/holo_diffusion/holo_multipass_ea.py:15-125
  * ``EmissionAbsorptionRaymarcher`` settings       -- /root/reference/configs/base.yaml:149-159
Implicitron materialises (densities, features) between the implicit function and the ray marcher; here the
seam sits at the renderer: ``HoloMultiPassEmissionAbsorptionRenderer.forward`` recognises its own implicit function
and runs ONE kernel launch for all rays and all passes (``holo_render_fwd``).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .cameras import ImplicitronRayBundle


class EvaluationMode(enum.Enum):
    TRAINING = "training"
    EVALUATION = "evaluation"


@dataclass
class RendererOutput:
    features: torch.Tensor
    depths: torch.Tensor
    masks: torch.Tensor
    prev_stage: Optional["RendererOutput"] = None
    normals: Optional[torch.Tensor] = None
    points: Optional[torch.Tensor] = None
    weights: Optional[torch.Tensor] = None
    aux: Dict[str, Any] = field(default_factory=dict)


class _MLPParams(nn.Module):
    """Parameter layout of MLPWithInputSkips: ``mlp.{i}.0.{weight,bias}`` (custom_modules.py:91-113)."""

    def __init__(self, n_layers, input_dim, output_dim, skip_dim, hidden_dim, input_skips):
        super().__init__()
        layers = []
        for li in range(n_layers):
            din = hidden_dim if li > 0 else input_dim
            dout = hidden_dim if li + 1 < n_layers else output_dim
            if li > 0 and li in input_skips:
                din = hidden_dim + skip_dim
            lin = nn.Linear(din, dout)
            nn.init.xavier_uniform_(lin.weight)
            layers.append(nn.Sequential(lin, nn.Identity()))
        self.mlp = nn.ModuleList(layers)
        self.input_skips = tuple(input_skips)


class RenderMLP(nn.Module):
    def __init__(self, input_dims: int = 128, output_feature_dims: int = 3, output_vp_independent_feature_dims: int = 64,
                 feat_emb_dims: int = 0, dir_emb_dims: int = 4, dnet_num_layers: int = 4, dnet_hidden_dim: int = 256,
                 dnet_input_skips: Sequence[int] = (2,), rnet_num_layers: int = 1, rnet_hidden_dim: int = 128,
                 rnet_input_skips: Sequence[int] = (), activation_fn: str = "LEAKYRELU"):
        super().__init__()
        if feat_emb_dims != 0 or rnet_num_layers != 1 or output_feature_dims != 3:
            raise NotImplementedError("fused renderer covers the shipped RenderMLP: identity feature embedding, "
                                      "one radiance layer, RGB output")
        if str(activation_fn).upper().split(".")[-1] != "LEAKYRELU":
            raise NotImplementedError("only the LeakyReLU(0.2) hidden activation of the shipped configs is built")
        if output_vp_independent_feature_dims != 0:
            raise NotImplementedError("view-independent feature head: HoloDiffusionModel forces feature_dim=0 "
                                      "(holo_diffusion_model.py:156)")
        self.input_dims, self.dir_emb_dims, self.dnet_hidden_dim = input_dims, dir_emb_dims, dnet_hidden_dim
        E = 3 * (2 * dir_emb_dims + 1)
        self._density_net = _MLPParams(dnet_num_layers, input_dims, dnet_hidden_dim + 1, input_dims, dnet_hidden_dim,
                                       dnet_input_skips)
        self._radiance_net = _MLPParams(1, dnet_hidden_dim + E, 3, dnet_hidden_dim + E, rnet_hidden_dim, ())
        self._packed = None
        self._packed_key = None

    def packed(self):
        """Collapsed + packed weights for the render kernel, rebuilt only when a parameter changes."""
        ps = list(self.parameters())
        key = tuple((p._version, p.data_ptr()) for p in ps)
        if key != self._packed_key:
            layers = [(seq[0].weight.detach().float().contiguous(), seq[0].bias.detach().float().contiguous())
                      for seq in self._density_net.mlp]
            rl = self._radiance_net.mlp[0][0]
            self._packed = ops.collapse_and_pack_render_mlp(layers, self._density_net.input_skips,
                                                            rl.weight.detach().float().contiguous(),
                                                            rl.bias.detach().float().contiguous(), self.input_dims)
            self._packed_key = key
        return self._packed


class HoloVoxelGridImplicitFunction(nn.Module):
    def __init__(self, resol: int = 32, volume_extent: float = 8.0, n_hidden: int = 128, feature_dim: int = 64,
                 init_density_bias: float = 1e-4, render_normals: bool = False, render_mlp_args: Optional[dict] = None):
        super().__init__()
        if render_normals:
            raise NotImplementedError("render_normals (autograd normals) is not part of the built path")
        self.resol, self.volume_extent, self.n_hidden, self.feature_dim = resol, volume_extent, n_hidden, feature_dim
        args = dict(render_mlp_args or {})
        args.update(input_dims=n_hidden, output_feature_dims=3, output_vp_independent_feature_dims=feature_dim)
        self.render_mlp = RenderMLP(**args)

    @staticmethod
    def allows_multiple_passes() -> bool:
        return True

    def forward(self, *, ray_bundle=None, pts_3d=None, voxel_grid_features=None, **kwargs):
        raise NotImplementedError(
            "The per-point (densities, features) seam is fused away: render through "
            "HoloMultiPassEmissionAbsorptionRenderer, which launches holo_render_fwd for this implicit function.")


class ImplicitFunctionWrapper(nn.Module):
    """pytorch3d.implicitron ImplicitFunctionWrapper: holds the function as ``_fn`` and late-bound kwargs."""

    def __init__(self, fn: nn.Module):
        super().__init__()
        self._fn = fn
        self.bound_args: Dict[str, Any] = {}

    def bind_args(self, **kwargs):
        self.bound_args = kwargs

    def unbind_args(self):
        self.bound_args = {}


def grid_to_channels_last(grid_ncdhw: torch.Tensor) -> torch.Tensor:
    """(1,C,D,H,W) -> (D,H,W,C) with the transpose kernel (one 128-byte line per voxel corner at C=32)."""
    _, C, D, H, W = grid_ncdhw.shape
    V = D * H * W
    return ops.transpose2d(grid_ncdhw.contiguous().float().reshape(-1), C, V).view(D, H, W, C)


class HoloMultiPassEmissionAbsorptionRenderer(nn.Module):
    def __init__(self, n_pts_per_ray_fine_training: int = 64, n_pts_per_ray_fine_evaluation: int = 64,
                 stratified_sampling_coarse_training: bool = True, stratified_sampling_coarse_evaluation: bool = False,
                 append_coarse_samples_to_fine: bool = True, density_noise_std_train: float = 1.0,
                 return_weights: bool = False, raymarcher_class_type: str = "EmissionAbsorptionRaymarcher",
                 raymarcher_EmissionAbsorptionRaymarcher_args: Optional[dict] = None, use_tensor_cores: bool = True):
        super().__init__()
        self.use_tensor_cores = use_tensor_cores
        if raymarcher_class_type != "EmissionAbsorptionRaymarcher":
            raise NotImplementedError("only EmissionAbsorptionRaymarcher is fused")
        rm = dict(surface_thickness=1, bg_color=(0.0,), replicate_last_interval=False, background_opacity=1e10,
                  density_relu=True, blend_output=False)
        rm.update(raymarcher_EmissionAbsorptionRaymarcher_args or {})
        if rm["surface_thickness"] != 1 or rm["replicate_last_interval"] or not rm["density_relu"] or rm["blend_output"]:
            raise NotImplementedError("fused ray marcher covers the shipped settings (configs/base.yaml:150-159)")
        bg = tuple(float(x) for x in rm["bg_color"])
        self.bg_color = bg * 3 if len(bg) == 1 else bg
        self.background_opacity = float(rm["background_opacity"])
        self.n_pts_per_ray_fine_evaluation = n_pts_per_ray_fine_evaluation
        self.n_pts_per_ray_fine_training = n_pts_per_ray_fine_training
        self.stratified_sampling_coarse_evaluation = stratified_sampling_coarse_evaluation
        self.append_coarse_samples_to_fine = append_coarse_samples_to_fine
        self.density_noise_std_train = density_noise_std_train
        self.return_weights = return_weights

    def forward(self, ray_bundle: ImplicitronRayBundle, implicit_functions: List[ImplicitFunctionWrapper],
                evaluation_mode: EvaluationMode = EvaluationMode.EVALUATION, **kwargs) -> RendererOutput:
        if evaluation_mode != EvaluationMode.EVALUATION:
            raise NotImplementedError("training-mode rendering (density noise, stratified refinement) is a 'next' row")
        if self.stratified_sampling_coarse_evaluation:
            raise NotImplementedError("stratified coarse sampling in evaluation")
        n_passes = len(implicit_functions)
        if n_passes not in (1, 2):
            raise NotImplementedError("1 or 2 rendering passes")
        w = implicit_functions[0]
        if any(f is not w for f in implicit_functions):
            raise NotImplementedError("all passes must share one implicit function (holo_diffusion_model.py:165-169)")
        fn = w._fn
        if not isinstance(fn, HoloVoxelGridImplicitFunction):
            raise NotImplementedError("fused renderer needs HoloVoxelGridImplicitFunction")
        grid = w.bound_args.get("voxel_grid_features")
        grid_cl = w.bound_args.get("voxel_grid_features_channels_last")
        if grid_cl is None:
            assert grid is not None, "voxel_grid_features must be provided!"
            assert grid.shape[0] == 1, "only one voxel grid per process (holo_diffusion_model.py:326)"
            grid_cl = grid_to_channels_last(grid)
        packed, hidden, E, tc_image = fn.render_mlp.packed()
        spatial = ray_bundle.lengths.shape[:-1]
        S = ray_bundle.lengths.shape[-1]
        n = int(torch.Size(spatial).numel())
        out = ops.render_fwd(grid_cl, fn.volume_extent, packed, hidden, fn.render_mlp.dir_emb_dims,
                             ray_bundle.origins.reshape(n, 3).contiguous(), ray_bundle.directions.reshape(n, 3).contiguous(),
                             ray_bundle.lengths.reshape(n, S).contiguous(), n_passes=n_passes,
                             n_fine=self.n_pts_per_ray_fine_evaluation, add_input_samples=self.append_coarse_samples_to_fine,
                             bg=self.bg_color, background_opacity=self.background_opacity,
                             return_weights=self.return_weights, tc_image=tc_image if self.use_tensor_cores else None)

        def wrap(o, prev):
            return RendererOutput(features=o["features"].view(*spatial, 3), depths=o["depths"].view(*spatial, 1),
                                  masks=o["masks"].view(*spatial, 1), prev_stage=prev,
                                  weights=None if o["weights"] is None else o["weights"].view(*spatial, -1),
                                  aux={"lengths": o["lengths"]})

        prev = wrap(out["prev"], None) if out["prev"] is not None else None
        return wrap(out, prev)
