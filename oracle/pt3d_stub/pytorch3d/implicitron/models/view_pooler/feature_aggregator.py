"""Names only: the view-pooling encoder is outside the hot path and never instantiated here."""
from enum import Enum


class FeatureAggregatorBase:
    pass


class ReductionFunction(Enum):
    AVG = "avg"
    MAX = "max"
    STD = "std"
    STD_AVG = "std_avg"


def _mask_target_view_features(*a, **k):
    raise NotImplementedError


def _get_view_sampling_mask(*a, **k):
    raise NotImplementedError


def _avgmaxstd_reduction_function(*a, **k):
    raise NotImplementedError
