"""Stand-in for MultiPassEmissionAbsorptionRenderer + EmissionAbsorptionRaymarcher + RayPointRefiner (TEST
INFRASTRUCTURE ONLY).  The subclass under test overrides _run_raymarcher; this base supplies what it reads:
self.raymarcher, self._refiners[mode], self.return_weights, and forward() that enters the recursion."""
import torch

from oracle import render_oracle as ro
from pytorch3d.implicitron.models.renderer.base import BaseRenderer, EvaluationMode, ImplicitronRayBundle, RendererOutput


class EmissionAbsorptionRaymarcher(torch.nn.Module):
    def __init__(self, bg_color=(0.0,), background_opacity: float = 1e10, surface_thickness: int = 1,
                 replicate_last_interval: bool = False, density_relu: bool = True, blend_output: bool = False):
        super().__init__()
        assert surface_thickness == 1 and not replicate_last_interval and density_relu and not blend_output
        self.bg_color, self.background_opacity = tuple(bg_color), background_opacity

    def forward(self, rays_densities, rays_features, aux, ray_lengths, ray_deltas=None, density_noise_std: float = 0.0, **kw):
        noise = None
        if density_noise_std > 0.0:
            noise = density_noise_std * torch.randn_like(rays_densities[..., 0])
        bg = self.bg_color if len(self.bg_color) > 1 else self.bg_color * rays_features.shape[-1]
        o = ro.ea_raymarch(rays_densities, rays_features, ray_lengths, bg, self.background_opacity, noise)
        return RendererOutput(features=o.features, depths=o.depths, masks=o.masks, weights=o.weights, aux=dict(aux))


class RayPointRefiner:
    def __init__(self, n_pts_per_ray: int, random_sampling: bool, add_input_samples: bool = True):
        self.n, self.random, self.add = n_pts_per_ray, random_sampling, add_input_samples

    def __call__(self, input_ray_bundle, ray_weights, **kw):
        u = None
        if self.random:
            u = torch.rand(*ray_weights.shape[:-1], self.n, dtype=ray_weights.dtype)
        z = ro.refine_lengths(input_ray_bundle.lengths, ray_weights, self.n, self.add, u)
        return ImplicitronRayBundle(input_ray_bundle.origins, input_ray_bundle.directions, z, input_ray_bundle.xys)


class MultiPassEmissionAbsorptionRenderer(BaseRenderer, torch.nn.Module):
    raymarcher_class_type: str = "EmissionAbsorptionRaymarcher"
    n_pts_per_ray_fine_training: int = 64
    n_pts_per_ray_fine_evaluation: int = 64
    stratified_sampling_coarse_training: bool = True
    stratified_sampling_coarse_evaluation: bool = False
    append_coarse_samples_to_fine: bool = True
    density_noise_std_train: float = 0.0
    return_weights: bool = False

    def __post_init__(self):
        self._refiners = {
            EvaluationMode.TRAINING: RayPointRefiner(self.n_pts_per_ray_fine_training,
                                                     self.stratified_sampling_coarse_training,
                                                     self.append_coarse_samples_to_fine),
            EvaluationMode.EVALUATION: RayPointRefiner(self.n_pts_per_ray_fine_evaluation,
                                                       self.stratified_sampling_coarse_evaluation,
                                                       self.append_coarse_samples_to_fine),
        }
        self.raymarcher = EmissionAbsorptionRaymarcher(**getattr(self, "raymarcher_EmissionAbsorptionRaymarcher_args", {}))

    def forward(self, ray_bundle, implicit_functions, evaluation_mode=EvaluationMode.EVALUATION, **kw):
        return self._run_raymarcher(ray_bundle, implicit_functions, None, evaluation_mode)
