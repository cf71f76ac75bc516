from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, Optional

import torch

from pytorch3d.implicitron.tools.config import ReplaceableBase


class EvaluationMode(Enum):
    TRAINING = "training"
    EVALUATION = "evaluation"


@dataclass
class ImplicitronRayBundle:
    origins: torch.Tensor
    directions: torch.Tensor
    lengths: torch.Tensor
    xys: Optional[torch.Tensor] = None
    camera_ids: Optional[torch.Tensor] = None
    camera_counts: Optional[torch.Tensor] = None


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


class BaseRenderer(ReplaceableBase):
    pass
