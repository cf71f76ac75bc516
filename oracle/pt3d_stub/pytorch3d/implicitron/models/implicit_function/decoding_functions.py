from enum import Enum

import torch


class DecoderActivation(Enum):
    RELU = "relu"
    SOFTPLUS = "softplus"
    SIGMOID = "sigmoid"
    IDENTITY = "identity"


def _xavier_init(linear) -> None:
    torch.nn.init.xavier_uniform_(linear.weight.data)
