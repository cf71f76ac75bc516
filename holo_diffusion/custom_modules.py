"""Names of /root/reference/holo_diffusion/custom_modules.py over holo_diffusion_b200: the activation enum the configs
refer to (:31-34), ``LazyLinearWithXavierInit`` (:37-41) and the view-pooling aggregator ``MLPMeanFeatureAggregator``
(:162-281), registered like the reference's when the Implicitron config system is present."""
import enum

import torch

import holo_diffusion_b200.encoder as _enc

from ._plugin import HAVE_CONFIG, adopt, registry


class HiddenActivation(enum.Enum):   # custom_modules.py:31-34
    RELU = "relu"
    SOFTPLUS = "softplus"
    LEAKYRELU = "leakyrelu"


LazyLinearWithXavierInit = _enc.LazyLinearWithXavierInit

_AGG_FIELDS = ("exclude_target_view", "exclude_target_view_mask_features", "concatenate_output", "n_hidden", "dim_out",
               "n_layers", "n_harmonic_functions_ray", "checkpointed_mlp")

if HAVE_CONFIG:
    from pytorch3d.implicitron.models.view_pooler.feature_aggregator import FeatureAggregatorBase

    @registry.register
    class MLPMeanFeatureAggregator(torch.nn.Module, FeatureAggregatorBase):
        exclude_target_view_mask_features: bool = True
        n_hidden: int = 128
        dim_out: int = 128
        n_layers: int = 1
        n_harmonic_functions_ray: int = 3
        checkpointed_mlp: bool = True

        def __post_init__(self):
            super().__init__()
            impl = _enc.MLPMeanFeatureAggregator(**{k: getattr(self, k) for k in _AGG_FIELDS if hasattr(self, k)})
            adopt(self, impl, ("_first_sampled", "_first_mean", "_last", "_mlp"))

        def get_aggregated_feature_dim(self, feats_or_feats_dim=None):
            return self.dim_out

        def forward(self, *a, **k):   # the pooling runs fused (holo_diffusion_b200.encoder.pool_views)
            return self._impl.forward(*a, **k)

else:
    MLPMeanFeatureAggregator = _enc.MLPMeanFeatureAggregator
