"""``RenderMLP`` / ``HoloVoxelGridImplicitFunction`` under the reference's module path
(/root/reference/holo_diffusion/holo_voxel_grid_implicit_function.py:48-129, :148-269) over holo_diffusion_b200."""
from typing import Tuple

import torch

import holo_diffusion_b200 as _b200

from ._plugin import HAVE_CONFIG, Configurable, adopt, fields, plain, registry, run_auto_creation
from .custom_modules import HiddenActivation

COLOUR_DIMS: int = 3
_MLP_FIELDS = ("input_dims", "output_feature_dims", "output_vp_independent_feature_dims", "feat_emb_dims", "dir_emb_dims",
               "dnet_num_layers", "dnet_hidden_dim", "dnet_input_skips", "rnet_num_layers", "rnet_hidden_dim",
               "rnet_input_skips", "activation_fn")

if HAVE_CONFIG:
    from pytorch3d.implicitron.models.implicit_function.base import ImplicitFunctionBase

    class RenderMLP(Configurable, torch.nn.Module):
        input_dims: int = 128
        output_feature_dims: int = COLOUR_DIMS
        output_vp_independent_feature_dims: int = 64
        feat_emb_dims: int = 0
        dir_emb_dims: int = 4
        dnet_num_layers: int = 4
        dnet_hidden_dim: int = 256
        dnet_input_skips: Tuple[int, ...] = (2,)
        rnet_num_layers: int = 1
        rnet_hidden_dim: int = 128
        rnet_input_skips: Tuple[int, ...] = ()
        activation_fn: HiddenActivation = HiddenActivation.LEAKYRELU

        def __post_init__(self):
            impl = _b200.RenderMLP(**{k: plain(v) for k, v in fields(self, _MLP_FIELDS).items()})
            adopt(self, impl, ("_density_net", "_radiance_net", "_feature_net"))

        def forward(self, features: torch.Tensor, view_dirs: torch.Tensor):
            return self._impl.forward(features, view_dirs)

    @registry.register
    class HoloVoxelGridImplicitFunction(ImplicitFunctionBase, torch.nn.Module):
        resol: int = 32
        volume_extent: float = 8.0
        n_hidden: int = 128
        feature_dim: int = 64
        init_density_bias: float = 1e-4
        render_mlp: RenderMLP
        render_normals: bool = False

        def __post_init__(self):
            run_auto_creation(self)
            impl = _b200.HoloVoxelGridImplicitFunction(
                resol=self.resol, volume_extent=self.volume_extent, n_hidden=self.n_hidden, feature_dim=self.feature_dim,
                init_density_bias=self.init_density_bias, render_normals=self.render_normals,
                render_mlp_args=plain(getattr(self, "render_mlp_args", {})))
            impl.render_mlp = self.render_mlp._impl   # ONE set of weights: the facade's RenderMLP is the registered child
            adopt(self, impl)

        def create_render_mlp(self):
            args = dict(getattr(self, "render_mlp_args", {}))
            args.update(input_dims=self.n_hidden, output_feature_dims=COLOUR_DIMS,
                        output_vp_independent_feature_dims=self.feature_dim)
            self.render_mlp = RenderMLP(**args)

        @staticmethod
        def allows_multiple_passes() -> bool:
            return True

        def forward(self, **kwargs):
            self._impl.render_normals = self.render_normals   # "visualisation switch" may be flipped after construction
            return self._impl.forward(**kwargs)

else:
    RenderMLP = _b200.RenderMLP
    HoloVoxelGridImplicitFunction = _b200.HoloVoxelGridImplicitFunction
