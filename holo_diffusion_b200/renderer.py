"""Renderer plug-ins: the reference's ImplicitFunction / Renderer surface over the fused CUDA render kernel.

Mirrors (class names, constructor fields, parameter names):
  * ``RenderMLP`` / ``HoloVoxelGridImplicitFunction`` -- /root/reference/holo_diffusion/holo_voxel_grid_implicit_function.py:48-269
  * ``MLPWithInputSkips`` parameter layout          -- /root/reference/holo_diffusion/custom_modules.py:44-160
  * ``HoloMultiPassEmissionAbsorptionRenderer``     -- /root/reference/holo_diffusion/holo_multipass_ea.py:15-125
  * ``EmissionAbsorptionRaymarcher`` settings       -- /root/reference/configs/base.yaml:149-159
  * ``EmissionAbsorptionRaymarcher`` / ``RayPointRefiner`` -- pytorch3d 0.7.4 (un-vendored), invoked at
    /root/reference/holo_diffusion/holo_multipass_ea.py:96-116
Implicitron materialises (densities, features) between the implicit function and the ray marcher.  Two paths:
  * fused (the hot path): ``HoloMultiPassEmissionAbsorptionRenderer.forward`` recognises its own implicit function
    and runs ONE kernel launch for all rays and all passes (``holo_render_fwd[_tc]``);
  * staged: every plug-in is callable on its own (``holo_if_fwd``, ``holo_ea_raymarch``, ``holo_ray_refine``) and
    ``_run_raymarcher`` chains them exactly as the reference does -- taken for ``render_normals``, the
    view-independent feature head, training-mode noise / stratified refinement, or a foreign plug-in.
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


def coerce_mode(mode) -> "EvaluationMode":
    """Accept pytorch3d's own EvaluationMode (a different enum class with the same values) or the plain string."""
    return mode if isinstance(mode, EvaluationMode) else EvaluationMode(getattr(mode, "value", mode))


def impl_of(obj):
    """The B200 module behind a registry facade (holo_diffusion/_plugin.py: `_impl` in the instance dict), or obj."""
    return getattr(obj, "__dict__", {}).get("_impl", obj)


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
            raise NotImplementedError("the kernels cover the shipped RenderMLP: identity feature embedding, "
                                      "one radiance layer, RGB output")
        if str(activation_fn).upper().split(".")[-1] != "LEAKYRELU":
            raise NotImplementedError("only the LeakyReLU(0.2) hidden activation of the shipped configs is built")
        if not 0 <= output_vp_independent_feature_dims <= 64:
            raise NotImplementedError("view-independent feature head of at most 64 outputs")
        self.input_dims, self.dir_emb_dims, self.dnet_hidden_dim = input_dims, dir_emb_dims, dnet_hidden_dim
        self.output_feature_dims = output_feature_dims
        self.output_vp_independent_feature_dims = output_vp_independent_feature_dims
        E = 3 * (2 * dir_emb_dims + 1)
        self._density_net = _MLPParams(dnet_num_layers, input_dims, dnet_hidden_dim + 1, input_dims, dnet_hidden_dim,
                                       dnet_input_skips)
        self._radiance_net = _MLPParams(1, dnet_hidden_dim + E, 3, dnet_hidden_dim + E, rnet_hidden_dim, ())
        self._feature_net = None
        if output_vp_independent_feature_dims > 0:  # holo_voxel_grid_implicit_function.py:94-105
            self._feature_net = _MLPParams(1, dnet_hidden_dim, output_vp_independent_feature_dims, dnet_hidden_dim,
                                           rnet_hidden_dim, ())
        self._packed = None
        self._packed_key = None

    def packed(self):
        """Collapsed + packed weights for the render kernels, rebuilt only when a parameter changes."""
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

    def head(self) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        if self._feature_net is None:
            return None
        lin = self._feature_net.mlp[0][0]
        return lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()

    @torch.no_grad()
    def forward(self, features: torch.Tensor, view_dirs: torch.Tensor):
        """-> densities (...,1), radiance (...,3), view-independent features (...,F) or None (:107-129)."""
        sp = features.shape[:-1]
        packed, hidden, _, _ = self.packed()
        dens, out = ops.render_mlp_fwd(features.reshape(-1, self.input_dims).contiguous().float(),
                                       view_dirs.expand(*sp, 3).reshape(-1, 3).contiguous().float(), packed, hidden,
                                       self.dir_emb_dims, self.head())
        vp = out[:, 3:].reshape(*sp, -1) if self._feature_net is not None else None
        return dens.view(*sp, 1), out[:, :3].reshape(*sp, 3), vp


def grid_to_channels_last(grid_ncdhw: torch.Tensor) -> torch.Tensor:
    """(1,C,D,H,W) -> (D,H,W,C) with the transpose kernel (one 128-byte line per voxel corner at C=32)."""
    _, C, D, H, W = grid_ncdhw.shape
    V = D * H * W
    return ops.transpose2d(grid_ncdhw.contiguous().float().reshape(-1), C, V).view(D, H, W, C)


class HoloVoxelGridImplicitFunction(nn.Module):
    def __init__(self, resol: int = 32, volume_extent: float = 8.0, n_hidden: int = 128, feature_dim: int = 64,
                 init_density_bias: float = 1e-4, render_normals: bool = False, render_mlp_args: Optional[dict] = None):
        super().__init__()
        self.resol, self.volume_extent, self.n_hidden, self.feature_dim = resol, volume_extent, n_hidden, feature_dim
        self.init_density_bias, self.render_normals = init_density_bias, render_normals
        args = dict(render_mlp_args or {})
        args.update(input_dims=n_hidden, output_feature_dims=3, output_vp_independent_feature_dims=feature_dim)
        self.render_mlp = RenderMLP(**args)

    @staticmethod
    def allows_multiple_passes() -> bool:
        return True

    differentiable = False   # True: autograd reaches the RenderMLP parameters even when the grid needs no gradient

    def forward(self, *, ray_bundle: Optional[ImplicitronRayBundle] = None, fun_viewpool=None, camera=None,
                global_code=None, run_id=None, pass_number=None, pts_3d: Optional[torch.Tensor] = None,
                voxel_grid_features: Optional[torch.Tensor] = None,
                voxel_grid_features_channels_last: Optional[torch.Tensor] = None, **kwargs):
        """Inference (no autograd graph) unless gradients are asked for: autograd enabled AND (the grid requires a
        gradient or ``differentiable`` is set) -> the differentiable kernels of autograd.py (training)."""
        if torch.is_grad_enabled() and (self.differentiable or
                                        (voxel_grid_features is not None and voxel_grid_features.requires_grad)):
            return self._forward_grad(ray_bundle=ray_bundle, pts_3d=pts_3d, voxel_grid_features=voxel_grid_features)
        return self._forward_nograd(ray_bundle=ray_bundle, fun_viewpool=fun_viewpool, camera=camera,
                                    global_code=global_code, run_id=run_id, pass_number=pass_number, pts_3d=pts_3d,
                                    voxel_grid_features=voxel_grid_features,
                                    voxel_grid_features_channels_last=voxel_grid_features_channels_last, **kwargs)

    def _forward_grad(self, *, ray_bundle, pts_3d, voxel_grid_features):
        from .autograd import ImplicitFunctionRays, compose_density_net
        if ray_bundle is None or pts_3d is not None or self.render_normals or self.render_mlp.head() is not None:
            raise NotImplementedError("differentiable implicit function: ray bundles only, no normals / feature head")
        assert voxel_grid_features is not None and voxel_grid_features.shape[0] == 1, "one NCDHW voxel grid"
        mlp = self.render_mlp
        layers = [(seq[0].weight, seq[0].bias) for seq in mlp._density_net.mlp]
        A, c = compose_density_net(layers, mlp._density_net.input_skips)
        rl = mlp._radiance_net.mlp[0][0]
        spatial = ray_bundle.lengths.shape
        S = spatial[-1]
        n = int(torch.Size(spatial[:-1]).numel())
        dens, rgb = ImplicitFunctionRays.apply(voxel_grid_features, A, c, rl.weight, rl.bias,
                                               ray_bundle.origins.reshape(n, 3), ray_bundle.directions.reshape(n, 3),
                                               ray_bundle.lengths.reshape(n, S), self.volume_extent, mlp.dir_emb_dims)
        return dens.view(*spatial, 1), rgb.view(*spatial, 3), {}

    @torch.no_grad()
    def _forward_nograd(self, *, ray_bundle: Optional[ImplicitronRayBundle] = None, fun_viewpool=None, camera=None,
                        global_code=None, run_id=None, pass_number=None, pts_3d: Optional[torch.Tensor] = None,
                        voxel_grid_features: Optional[torch.Tensor] = None,
                        voxel_grid_features_channels_last: Optional[torch.Tensor] = None, **kwargs):
        """-> densities (...,S,1), features (...,S,3[+feature_dim]), aux {"normals": (...,S,3)} (:182-269).
        One ``holo_if_fwd`` launch: ray points, trilinear sampling, RenderMLP and the analytic normals."""
        assert voxel_grid_features is not None or voxel_grid_features_channels_last is not None, \
            "voxel_grid_features must be provided!"
        assert ray_bundle is not None or pts_3d is not None, "either ray_bundle or pts_3d must be provided!"
        grid_cl = voxel_grid_features_channels_last
        if grid_cl is None:
            assert voxel_grid_features.shape[0] == 1, "only batch size of 1 is supported"
            grid_cl = grid_to_channels_last(voxel_grid_features)
        packed, hidden, _, _ = self.render_mlp.packed()
        kw = dict(head=self.render_mlp.head(), normals=self.render_normals)
        if pts_3d is None:
            spatial = ray_bundle.lengths.shape
            S = spatial[-1]
            n = int(torch.Size(spatial[:-1]).numel())
            dens, feats, nrm = ops.if_fwd(grid_cl, self.volume_extent, packed, hidden, self.render_mlp.dir_emb_dims,
                                          origins=ray_bundle.origins.reshape(n, 3).contiguous().float(),
                                          dirs=ray_bundle.directions.reshape(n, 3).contiguous().float(),
                                          lengths=ray_bundle.lengths.reshape(n, S).contiguous().float(), **kw)
        else:
            spatial = pts_3d.shape[:-1]
            S = spatial[-1]
            dirs = None
            if ray_bundle is not None:
                dirs = ray_bundle.directions.reshape(-1, 3).contiguous().float()
            dens, feats, nrm = ops.if_fwd(grid_cl, self.volume_extent, packed, hidden, self.render_mlp.dir_emb_dims,
                                          dirs=dirs, pts_3d=pts_3d.reshape(-1, 3).contiguous().float(), S=S, **kw)
        aux: Dict[str, Any] = {}
        if nrm is not None:
            aux["normals"] = nrm.view(*spatial, 3)
        return dens.view(*spatial, 1), feats.view(*spatial, feats.shape[-1]), aux


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

    def forward(self, *args, **kwargs):
        return self._fn(*args, **{**self.bound_args, **kwargs})


class EmissionAbsorptionRaymarcher(nn.Module):
    """pytorch3d 0.7.4 EmissionAbsorptionRaymarcher (configs/base.yaml:149-159) over ``holo_ea_raymarch``."""

    def __init__(self, surface_thickness: int = 1, bg_color: Sequence[float] = (0.0,), replicate_last_interval: bool = False,
                 background_opacity: float = 1e10, density_relu: bool = True, blend_output: bool = False):
        super().__init__()
        if surface_thickness != 1 or replicate_last_interval or not density_relu or blend_output:
            raise NotImplementedError("ray marcher kernels cover the shipped settings (configs/base.yaml:150-159)")
        self.surface_thickness, self.replicate_last_interval = surface_thickness, replicate_last_interval
        self.density_relu, self.blend_output = density_relu, blend_output
        self.bg_color = tuple(float(x) for x in bg_color)
        self.background_opacity = float(background_opacity)

    def forward(self, rays_densities: torch.Tensor, rays_features: torch.Tensor, aux: Dict[str, Any],
                ray_lengths: torch.Tensor, ray_deltas: Optional[torch.Tensor] = None, density_noise_std: float = 0.0,
                **kwargs) -> RendererOutput:
        if torch.is_grad_enabled() and (rays_densities.requires_grad or rays_features.requires_grad):
            return self._forward_grad(rays_densities, rays_features, aux, ray_lengths, ray_deltas, density_noise_std)
        return self._forward_nograd(rays_densities, rays_features, aux, ray_lengths, ray_deltas, density_noise_std, **kwargs)

    def _forward_grad(self, rays_densities, rays_features, aux, ray_lengths, ray_deltas, density_noise_std):
        from .autograd import EARaymarch
        if ray_deltas is not None or "normals" in aux:
            raise NotImplementedError("differentiable ray marcher: no explicit ray_deltas, no normals")
        spatial = ray_lengths.shape[:-1]
        S = ray_lengths.shape[-1]
        n = int(torch.Size(spatial).numel())
        Fd = rays_features.shape[-1]
        if len(self.bg_color) not in (1, Fd):
            raise ValueError(f"Wrong number of background color channels: {len(self.bg_color)} for {Fd} features")
        noise = None
        if density_noise_std > 0.0:   # rays_densities + randn_like * std before the relu (holo_multipass_ea.py:87-91)
            noise = (torch.randn(n, S, device=ray_lengths.device) * density_noise_std).contiguous()
        f, d, m, w = EARaymarch.apply(rays_densities.reshape(n, S), rays_features.reshape(n, S, Fd),
                                      ray_lengths.reshape(n, S), noise, tuple(self.bg_color), self.background_opacity)
        return RendererOutput(features=f.view(*spatial, Fd), depths=d.view(*spatial, 1), masks=m.view(*spatial, 1),
                              weights=w.view(*spatial, S), aux=dict(aux))

    @torch.no_grad()
    def _forward_nograd(self, rays_densities: torch.Tensor, rays_features: torch.Tensor, aux: Dict[str, Any],
                        ray_lengths: torch.Tensor, ray_deltas: Optional[torch.Tensor] = None, density_noise_std: float = 0.0,
                        **kwargs) -> RendererOutput:
        if ray_deltas is not None:
            raise NotImplementedError("explicit ray_deltas")
        spatial = ray_lengths.shape[:-1]
        S = ray_lengths.shape[-1]
        n = int(torch.Size(spatial).numel())
        Fd = rays_features.shape[-1]
        if len(self.bg_color) not in (1, Fd):
            raise ValueError(f"Wrong number of background color channels: {len(self.bg_color)} for {Fd} features")
        dens = rays_densities.reshape(n, S).contiguous().float()
        noise = None
        if density_noise_std > 0.0:
            noise = (torch.randn_like(dens) * density_noise_std).contiguous()
        aux = dict(aux)
        normals = aux.get("normals")
        o = ops.ea_raymarch(dens, rays_features.reshape(n, S, Fd).contiguous().float(),
                            ray_lengths.reshape(n, S).contiguous().float(), self.bg_color, self.background_opacity,
                            noise=noise, normals=None if normals is None else normals.reshape(n, S, 3).contiguous())
        if normals is not None:
            # the reference multiplies and sums in Python (holo_multipass_ea.py:104-109); the kernel did it already
            aux.pop("normals")
            aux["rendered_normals"] = o["normals"].view(*spatial, 3)
        return RendererOutput(features=o["features"].view(*spatial, Fd), depths=o["depths"].view(*spatial, 1),
                              masks=o["masks"].view(*spatial, 1), weights=o["weights"].view(*spatial, S), aux=aux)


class RayPointRefiner:
    """pytorch3d 0.7.4 RayPointRefiner (+ sample_pdf) over ``holo_ray_refine``; random_sampling draws the
    uniforms with torch's generator (training), otherwise u = linspace(0, 1, n_pts_per_ray)."""

    def __init__(self, n_pts_per_ray: int, random_sampling: bool, add_input_samples: bool = True):
        self.n_pts_per_ray, self.random_sampling, self.add_input_samples = n_pts_per_ray, random_sampling, add_input_samples

    @torch.no_grad()
    def __call__(self, input_ray_bundle: ImplicitronRayBundle, ray_weights: torch.Tensor, **kwargs) -> ImplicitronRayBundle:
        z = input_ray_bundle.lengths
        spatial = z.shape[:-1]
        S = z.shape[-1]
        n = int(torch.Size(spatial).numel())
        u = None
        if self.random_sampling:
            u = torch.rand(n, self.n_pts_per_ray, device=z.device)
        new = ops.ray_refine(z.reshape(n, S).contiguous().float(), ray_weights.reshape(n, S).contiguous().float(),
                             self.n_pts_per_ray, self.add_input_samples, u)
        return ImplicitronRayBundle(input_ray_bundle.origins, input_ray_bundle.directions, new.view(*spatial, -1),
                                    input_ray_bundle.xys, input_ray_bundle.camera_ids, input_ray_bundle.camera_counts)


class HoloMultiPassEmissionAbsorptionRenderer(nn.Module):
    def __init__(self, n_pts_per_ray_fine_training: int = 64, n_pts_per_ray_fine_evaluation: int = 64,
                 stratified_sampling_coarse_training: bool = True, stratified_sampling_coarse_evaluation: bool = False,
                 append_coarse_samples_to_fine: bool = True, density_noise_std_train: float = 1.0,
                 return_weights: bool = False, raymarcher_class_type: str = "EmissionAbsorptionRaymarcher",
                 raymarcher_EmissionAbsorptionRaymarcher_args: Optional[dict] = None, use_tensor_cores: bool = True,
                 fused: bool = True):
        super().__init__()
        self.use_tensor_cores, self.fused = use_tensor_cores, fused
        if raymarcher_class_type != "EmissionAbsorptionRaymarcher":
            raise NotImplementedError("only EmissionAbsorptionRaymarcher is built")
        self.raymarcher = EmissionAbsorptionRaymarcher(**(raymarcher_EmissionAbsorptionRaymarcher_args or {}))
        bg = self.raymarcher.bg_color
        self.bg_color = bg * 3 if len(bg) == 1 else bg
        self.background_opacity = self.raymarcher.background_opacity
        self.n_pts_per_ray_fine_evaluation = n_pts_per_ray_fine_evaluation
        self.n_pts_per_ray_fine_training = n_pts_per_ray_fine_training
        self.stratified_sampling_coarse_training = stratified_sampling_coarse_training
        self.stratified_sampling_coarse_evaluation = stratified_sampling_coarse_evaluation
        self.append_coarse_samples_to_fine = append_coarse_samples_to_fine
        self.density_noise_std_train = density_noise_std_train
        self.return_weights = return_weights
        self._refiners = {
            EvaluationMode.TRAINING: RayPointRefiner(n_pts_per_ray_fine_training, stratified_sampling_coarse_training,
                                                     append_coarse_samples_to_fine),
            EvaluationMode.EVALUATION: RayPointRefiner(n_pts_per_ray_fine_evaluation,
                                                       stratified_sampling_coarse_evaluation,
                                                       append_coarse_samples_to_fine),
        }

    # ------------------------------------------------------------------ fused path (one launch)
    def is_fused(self, implicit_functions: List[ImplicitFunctionWrapper], evaluation_mode: EvaluationMode) -> bool:
        """True when the whole multi-pass render of these plug-ins is covered by the single fused kernel."""
        if not self.fused or evaluation_mode != EvaluationMode.EVALUATION or self.stratified_sampling_coarse_evaluation:
            return False
        if len(implicit_functions) not in (1, 2) or type(self.raymarcher) is not EmissionAbsorptionRaymarcher:
            return False
        w = implicit_functions[0]   # ours or pytorch3d's ImplicitFunctionWrapper: `_fn` + late-bound `bound_args`
        if not (hasattr(w, "_fn") and hasattr(w, "bound_args")) or any(f is not w for f in implicit_functions):
            return False
        fn = impl_of(w._fn)
        if type(fn) is not HoloVoxelGridImplicitFunction or fn.render_normals or fn.render_mlp.head() is not None:
            return False
        return "voxel_grid_features" in w.bound_args or "voxel_grid_features_channels_last" in w.bound_args

    def _forward_fused(self, ray_bundle: ImplicitronRayBundle, implicit_functions) -> RendererOutput:
        n_passes = len(implicit_functions)
        w = implicit_functions[0]
        fn = impl_of(w._fn)
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

    # ------------------------------------------------------------------ staged path (the reference's recursion)
    def _run_raymarcher(self, ray_bundle, implicit_functions, prev_stage, evaluation_mode, pass_number=0):
        """holo_multipass_ea.py:79-125, stage by stage on the per-stage kernels."""
        density_noise_std = self.density_noise_std_train if evaluation_mode == EvaluationMode.TRAINING else 0.0
        if_output = implicit_functions[0](ray_bundle=ray_bundle, pass_number=pass_number)
        output = self.raymarcher(*if_output, ray_lengths=ray_bundle.lengths, density_noise_std=density_noise_std)
        output.prev_stage = prev_stage
        weights = output.weights
        if "rendered_normals" in output.aux:
            output.normals = output.aux.pop("rendered_normals")
        elif "normals" in output.aux:  # a foreign ray marcher that passed the per-point normals through
            output.normals = (output.aux.pop("normals") * weights[..., None]).sum(dim=-2)
        output.aux["lengths"] = ray_bundle.lengths
        if not self.return_weights:
            output.weights = None
        if len(implicit_functions) > 1:
            fine_ray_bundle = self._refiners[evaluation_mode](ray_bundle, weights.detach())
            output = self._run_raymarcher(fine_ray_bundle, implicit_functions[1:], output, evaluation_mode,
                                          pass_number=pass_number + 1)
        return output

    @staticmethod
    def _wants_grad(implicit_functions) -> bool:
        for w in implicit_functions:
            g = getattr(w, "bound_args", {}).get("voxel_grid_features")
            if (g is not None and g.requires_grad) or getattr(impl_of(getattr(w, "_fn", w)), "differentiable", False):
                return True
        return False

    def forward(self, ray_bundle: ImplicitronRayBundle, implicit_functions: List[ImplicitFunctionWrapper],
                evaluation_mode: EvaluationMode = EvaluationMode.EVALUATION, **kwargs) -> RendererOutput:
        if not implicit_functions:
            raise ValueError("EA renderer expects implicit functions")
        evaluation_mode = coerce_mode(evaluation_mode)
        if torch.is_grad_enabled() and self._wants_grad(implicit_functions):
            # training: the reference's recursion on the differentiable per-stage kernels (autograd.py); the refiner
            # inside stays under no_grad as in pytorch3d
            return self._run_raymarcher(ray_bundle, list(implicit_functions), None, evaluation_mode)
        with torch.no_grad():
            if self.is_fused(implicit_functions, evaluation_mode):
                return self._forward_fused(ray_bundle, implicit_functions)
            return self._run_raymarcher(ray_bundle, list(implicit_functions), None, evaluation_mode)
