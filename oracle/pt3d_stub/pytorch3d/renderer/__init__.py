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


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees=True, up=((0, 1, 0),), **kw):
    """Stand-in leaf: R (N,3,3), T (N,3) from the oracle's restatement (unpinned arithmetic)."""
    assert degrees
    d, e, a = (torch.as_tensor(float(v), dtype=torch.float32).reshape(1) for v in (dist, elev, azim))
    return ro.look_at_rotation_translation(d, e, a, up=tuple(float(u) for u in up[0]), dtype=torch.float32)


class PerspectiveCameras:
    """Container only (NDC PerspectiveCameras): the in-tree code under test just constructs it."""

    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=None, T=None, **kw):
        self.focal_length, self.principal_point, self.R, self.T = focal_length, principal_point, R, T
