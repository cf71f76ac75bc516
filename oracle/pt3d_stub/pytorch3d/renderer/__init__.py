"""Stand-in leaves (TEST INFRASTRUCTURE ONLY): delegate to the oracle's restatement -- unpinned arithmetic."""
import torch

from oracle import render_oracle as ro


class HarmonicEmbedding(torch.nn.Module):
    def __init__(self, n_harmonic_functions: int = 6, omega_0: float = 1.0, logspace: bool = True, append_input: bool = True):
        super().__init__()
        assert omega_0 == 1.0 and logspace and append_input
        self.n = n_harmonic_functions

    def get_output_dim(self, input_dims: int = 3) -> int:
        return input_dims * (2 * self.n + 1)

    def forward(self, x):
        return ro.harmonic_embedding(x, self.n)


def ray_bundle_to_ray_points(ray_bundle):
    return ro.ray_points(ray_bundle)
