"""holo_diffusion_b200 -- B200-native (sm_100a) implementation of HoloDiffusion's hot path.

The 3-D UNet denoise step and the volumetric renderer as hand-written CUDA kernels behind a C-ABI
(``include/holo_b200.h``, ``libholo_b200.so``), exposed through the reference's plug-in surface
(``SimpleUnet3D``, ``ImplicitronGaussianDiffusion``, ``HoloVoxelGridImplicitFunction``,
``HoloMultiPassEmissionAbsorptionRenderer``, ``HoloDiffusionModel.forward``).  No CPU fallback.
"""
from ._lib import HoloError, lib  # noqa: F401  (importing the package loads the CUDA library or fails loudly)

lib()

from .cameras import (AdaptiveRaySampler, ImplicitronRayBundle, PerspectiveCameras,  # noqa: E402,F401
                      get_simple_360_camera_trajectory, look_at_view_transform)
from .diffusion import ImplicitronGaussianDiffusion  # noqa: E402,F401
from . import encoder  # noqa: E402,F401
from .encoder import (AngleWeightedReductionFeatureAggregator, MLPMeanFeatureAggregator,  # noqa: E402,F401
                      ResNetFeatureExtractor, ViewPooler)
from .model import HoloDiffusionModel  # noqa: E402,F401
from .pipeline import ViewStream  # noqa: E402,F401
from .renderer import (EmissionAbsorptionRaymarcher, EvaluationMode,  # noqa: E402,F401
                       HoloMultiPassEmissionAbsorptionRenderer, HoloVoxelGridImplicitFunction, ImplicitFunctionWrapper,
                       RayPointRefiner, RendererOutput, RenderMLP)
from .unet import SimpleUnet3D, Unet3DBase  # noqa: E402,F401
