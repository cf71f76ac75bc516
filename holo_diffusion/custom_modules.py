"""Names of /root/reference/holo_diffusion/custom_modules.py that configs refer to."""
import enum


class HiddenActivation(enum.Enum):   # custom_modules.py:31-34
    RELU = "relu"
    SOFTPLUS = "softplus"
    LEAKYRELU = "leakyrelu"
