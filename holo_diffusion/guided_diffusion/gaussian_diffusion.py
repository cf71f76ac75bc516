"""Configuration enums of /root/reference/holo_diffusion/guided_diffusion/gaussian_diffusion.py:67-107 (names only: the
yaml configs and ``ImplicitronGaussianDiffusion``'s fields refer to them)."""
import enum


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()
