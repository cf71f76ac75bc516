"""Host-side camera model, fly-around trajectory and the evaluation ray sampler.

Mirrors the pieces of pytorch3d 0.7.4 the reference hot path touches (un-vendored there):
  * ``PerspectiveCameras`` (NDC, row-vector convention ``X_cam = X_world R + T``),
  * ``look_at_view_transform`` / ``so3_exp_map``,
  * ``get_simple_360_camera_trajectory`` -- /root/reference/holo_diffusion/utils/render_utils/flyaround.py:301-350,
  * ``AdaptiveRaySampler`` (full-grid, non-stratified evaluation) -- configs/base.yaml:129-140, invoked at
    /root/reference/holo_diffusion/holo_diffusion_model.py:442-448.
Camera parameters are a few floats per view and are prepared with torch on the host; the per-ray work
(unprojection, origins, unit directions, depths) is the CUDA kernel ``holo_raygen``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops


class PerspectiveCameras:
    def __init__(self, focal_length, principal_point=None, R=None, T=None, device=None):
        R = torch.as_tensor(R, dtype=torch.float32)
        n = R.shape[0]
        f = torch.as_tensor(focal_length, dtype=torch.float32).reshape(n, -1)
        if f.shape[1] == 1:
            f = f.expand(n, 2)
        pp = torch.zeros(n, 2) if principal_point is None else torch.as_tensor(principal_point, dtype=torch.float32)
        self.R, self.T = R.contiguous(), torch.as_tensor(T, dtype=torch.float32).reshape(n, 3).contiguous()
        self.focal_length, self.principal_point = f.contiguous(), pp.reshape(n, 2).contiguous()
        if device is not None:
            self.to(device)

    def __len__(self):
        return self.R.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = [idx]
        return PerspectiveCameras(self.focal_length[idx], self.principal_point[idx], self.R[idx], self.T[idx])

    def to(self, device):
        self.R, self.T = self.R.to(device), self.T.to(device)
        self.focal_length, self.principal_point = self.focal_length.to(device), self.principal_point.to(device)
        return self

    @property
    def device(self):
        return self.R.device

    def get_camera_center(self):
        return -torch.einsum("nj,nij->ni", self.T, self.R)


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, up=((0.0, 1.0, 0.0),)):
    """Camera on a sphere around the origin looking at it (degrees), pytorch3d convention."""
    d, e, a = torch.broadcast_tensors(torch.as_tensor(dist, dtype=torch.float32).reshape(-1),
                                      torch.as_tensor(elev, dtype=torch.float32).reshape(-1),
                                      torch.as_tensor(azim, dtype=torch.float32).reshape(-1))
    e, a = e * (math.pi / 180.0), a * (math.pi / 180.0)
    centre = torch.stack([d * e.cos() * a.sin(), d * e.sin(), d * e.cos() * a.cos()], -1)
    upv = torch.as_tensor(up, dtype=torch.float32).reshape(-1, 3).expand_as(centre)
    zax = F.normalize(-centre, dim=-1, eps=1e-5)
    xax = F.normalize(torch.cross(upv, zax, dim=-1), dim=-1, eps=1e-5)
    yax = F.normalize(torch.cross(zax, xax, dim=-1), dim=-1, eps=1e-5)
    R = torch.stack([xax, yax, zax], dim=2)
    T = -torch.einsum("nji,nj->ni", R, centre)
    return R, T


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    theta = (log_rot * log_rot).sum(-1).clamp(eps).sqrt()
    x, y, z = log_rot.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1).reshape(-1, 3, 3)
    s = (theta.sin() / theta)[:, None, None]
    c = ((1.0 - theta.cos()) / (theta * theta))[:, None, None]
    return torch.eye(3)[None] + s * K + c * (K @ K)


def get_simple_360_camera_trajectory(max_angle: float, n_flyaround_poses: int, camera_elevation: float,
                                     hemispherical_radius: float, up: Tuple[float, float, float],
                                     camera_focal_length: float,
                                     canonical_up: Tuple[float, float, float] = (0.0, -1.0, 0.0)) -> PerspectiveCameras:
    n = n_flyaround_poses
    az = torch.linspace(0, math.degrees(max_angle), n + 1)[:n]
    R, T = look_at_view_transform(torch.full((n,), float(hemispherical_radius)),
                                  torch.full((n,), math.degrees(camera_elevation)), az, up=(canonical_up,))
    axis = torch.cross(torch.tensor(canonical_up), torch.tensor(up), dim=-1)
    R = torch.bmm(so3_exp_map(axis[None]).expand(n, 3, 3), R)
    return PerspectiveCameras(torch.full((n, 1), float(camera_focal_length)), torch.zeros(n, 2), R, T)


@dataclass
class ImplicitronRayBundle:
    origins: torch.Tensor
    directions: torch.Tensor
    lengths: torch.Tensor
    xys: torch.Tensor
    camera_ids: Optional[torch.Tensor] = None
    camera_counts: Optional[torch.Tensor] = None


class AdaptiveRaySampler:
    """Evaluation-mode (full grid, non-stratified) ray sampler with the depth range adapted to the camera
    distance: [|C - scene_center| - scene_extent, ... + scene_extent].  Ray r = h * W + w, row-major."""

    def __init__(self, image_width: int = 256, image_height: int = 256, n_pts_per_ray_evaluation: int = 64,
                 scene_extent: float = 4.0, scene_center: Sequence[float] = (0.0, 0.0, 0.0), **unused):
        self.image_width, self.image_height = image_width, image_height
        self.n_pts_per_ray_evaluation = n_pts_per_ray_evaluation
        self.scene_extent, self.scene_center = scene_extent, tuple(scene_center)
        self._xy: Dict[Tuple[int, int, str], torch.Tensor] = {}

    @staticmethod
    def ndc_pixel_grid(H: int, W: int) -> torch.Tensor:
        """(H, W, 2) NDC (x, y) pixel centres; +x left, +y up; the short side spans [-1, 1]."""
        rx, ry = (W / H, 1.0) if W >= H else (1.0, H / W)
        xs = torch.linspace(rx - rx / W, -rx + rx / W, W, dtype=torch.float32)
        ys = torch.linspace(ry - ry / H, -ry + ry / H, H, dtype=torch.float32)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([xx, yy], -1)

    def xy_grid(self, device) -> torch.Tensor:
        key = (self.image_height, self.image_width, str(device))
        if key not in self._xy:
            self._xy[key] = self.ndc_pixel_grid(self.image_height, self.image_width).reshape(-1, 2).contiguous().to(device)
        return self._xy[key]

    def __call__(self, cameras: PerspectiveCameras, evaluation_mode=None, mask=None) -> ImplicitronRayBundle:
        if mask is not None:
            raise NotImplementedError("mask_sample ray sampling (training) is not part of the built path")
        dev = cameras.device
        ops.require_cuda(dev, "AdaptiveRaySampler (cameras)")
        H, W, S = self.image_height, self.image_width, self.n_pts_per_ray_evaluation
        xy = self.xy_grid(dev)
        o, d, l = ops.raygen(cameras.R, cameras.T, cameras.focal_length, cameras.principal_point, xy, S,
                             self.scene_extent, self.scene_center)
        B = len(cameras)
        return ImplicitronRayBundle(o.view(B, H, W, 3), d.view(B, H, W, 3), l.view(B, H, W, S),
                                    xy.view(1, H, W, 2).expand(B, H, W, 2))
