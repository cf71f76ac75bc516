"""``Unet3DBase`` / ``SimpleUnet3D`` / ``ImplicitronGaussianDiffusion`` under the reference's module path
(/root/reference/holo_diffusion/utils/diffusion_utils.py:30-140) over holo_diffusion_b200."""
from typing import Optional, Tuple

import torch

import holo_diffusion_b200 as _b200

from .._plugin import HAVE_CONFIG, Configurable, ReplaceableBase, adopt, fields, plain, registry
from ..guided_diffusion.gaussian_diffusion import ModelMeanType, ModelVarType

_UNET_FIELDS = ("image_size", "in_channels", "out_channels", "model_channels", "num_res_blocks", "channel_mult",
                "attention_resolutions", "num_heads", "dropout", "homogeneous_resample")
_DIFF_FIELDS = ("beta_schedule_type", "num_steps", "beta_start_unscaled", "beta_end_unscaled", "model_mean_type",
                "model_var_type", "schedule_sampler_type")

if HAVE_CONFIG:

    class Unet3DBase(ReplaceableBase, torch.nn.Module):
        def forward(self, x: torch.Tensor, timesteps: torch.Tensor, cond_features: Optional[torch.Tensor] = None,
                    **kwargs) -> torch.Tensor:
            raise NotImplementedError()

    @registry.register
    class SimpleUnet3D(Unet3DBase):
        image_size: int = 64
        in_channels: int = 128
        out_channels: int = 128
        model_channels: int = 128
        num_res_blocks: int = 2
        channel_mult: Tuple[int, ...] = (1, 2, 4, 8)
        attention_resolutions: Tuple[int, ...] = (8, 16)
        num_heads: int = 2
        dropout: float = 0.0
        homogeneous_resample: bool = True

        def __post_init__(self):
            adopt(self, _b200.SimpleUnet3D(**{k: plain(v) for k, v in fields(self, _UNET_FIELDS).items()}), ("_net",))

        def forward(self, x, timesteps, cond_features=None):
            return self._impl.forward(x, timesteps, cond_features)

        def shard_attention(self, group=None, min_tokens: int = 1 << 14):
            self._impl.shard_attention(group, min_tokens)

    class ImplicitronGaussianDiffusion(Configurable):
        beta_schedule_type: str = "linear"
        num_steps: int = 1000
        beta_start_unscaled: float = 0.0001
        beta_end_unscaled: float = 0.02
        model_mean_type: ModelMeanType = ModelMeanType.START_X
        model_var_type: ModelVarType = ModelVarType.FIXED_SMALL
        schedule_sampler_type: str = "uniform"

        def __post_init__(self):
            self._impl = _b200.ImplicitronGaussianDiffusion(**{k: plain(v) for k, v in fields(self, _DIFF_FIELDS).items()})

        def __getattr__(self, name):   # q_sample, p_mean_variance, p_sample[_loop[_progressive]], ddim_*, sample_timesteps
            if name == "_impl":
                raise AttributeError(name)
            return getattr(self._impl, name)

else:
    Unet3DBase = _b200.Unet3DBase
    SimpleUnet3D = _b200.SimpleUnet3D
    ImplicitronGaussianDiffusion = _b200.ImplicitronGaussianDiffusion
