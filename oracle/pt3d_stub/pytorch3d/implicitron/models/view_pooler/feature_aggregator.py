"""Stand-in leaves (TEST INFRASTRUCTURE ONLY) for pytorch3d.implicitron.models.view_pooler.feature_aggregator:
restated from memory of pytorch3d 0.7.4 -- unpinned arithmetic, delegating to oracle/encoder_oracle.py where it has
the same function."""
from enum import Enum

import torch

from oracle import encoder_oracle as eo
from pytorch3d.implicitron.tools.config import ReplaceableBase


class ReductionFunction(Enum):
    AVG = "avg"
    MAX = "max"
    STD = "std"
    STD_AVG = "std_avg"


class FeatureAggregatorBase(ReplaceableBase):
    exclude_target_view: bool = True
    exclude_target_view_mask_features: bool = True
    concatenate_output: bool = True


def _get_view_sampling_mask(n_cameras: int, pts_batch: int, device, exclude_target_view: bool):
    idx = torch.arange(n_cameras, dtype=torch.int64, device=device)[None].expand(pts_batch, n_cameras)
    if exclude_target_view:
        return (idx != torch.arange(pts_batch, dtype=torch.int64, device=device)[:, None]).float()
    return idx.new_ones(idx.shape).float()


def _mask_target_view_features(feats_sampled):
    one = next(iter(feats_sampled.values()))
    pts_batch, n_cameras = one.shape[:2]
    m = _get_view_sampling_mask(n_cameras, pts_batch, one.device, True)
    m = m.view(pts_batch, n_cameras, *([1] * (one.ndim - 2)))
    return {k: f * m for k, f in feats_sampled.items()}


def _avgmaxstd_reduction_function(x, w, reduction_functions, dim: int = 1):
    if list(reduction_functions) != [ReductionFunction.AVG]:
        raise NotImplementedError("stand-in: only the AVG reduction (what MLPMeanFeatureAggregator asks for)")
    return eo.wmean(x, w, dim=dim, eps=1e-2)
