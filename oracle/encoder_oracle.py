"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the view-pooling encoder (views -> voxel grid), SURVEY.md 8(f) row 2.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.

What it follows:
  * /root/reference/holo_diffusion/holo_diffusion_model.py:327-373 -- the encoder branch of ``forward``: image features
    -> ``VolumeLocator(...).get_coord_grid()`` -> ``view_pooler(pts, camera, feats, masks)`` ->
    ``pooled_feature_mapper`` -> ``permute(0, 3, 1, 2).reshape(1, -1, R, R, R)`` -> tanh;
  * /root/reference/holo_diffusion/custom_modules.py:162-334 -- ``MLPMeanFeatureAggregator`` and
    ``_get_point_to_source_camera_ray_dirs`` (IN-TREE: pinned by tests/golden/make_encoder_intree_golden.py, which
    executes the reference classes over oracle/pt3d_stub and compares with ``mlp_mean_aggregate`` below);
  * pytorch3d 0.7.4 leaves, ABSENT from /root/reference (environment.yaml:139) and restated from memory -- PARITY
    UNPINNED for these: ``VolumeLocator.get_coord_grid`` (voxel centres, x fastest), ``ViewSampler.forward`` /
    ``project_points_and_sample`` (NDC projection with the eps-clamped divide, ``ndc_grid_sample`` = negate + scale the
    longer side + ``F.grid_sample(align_corners=False)``), ``cameras_points_cartesian_product``, ``wmean(eps=1e-2)``,
    ``_get_view_sampling_mask``, ``HarmonicEmbedding``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import render_oracle as ro


# ------------------------------------------------------------------------------------------------ pytorch3d leaves
def coord_grid(resol: int, volume_extent: float, dtype=torch.float32) -> torch.Tensor:
    """VolumeLocator(1, (R, R, R), voxel_size=extent / R).get_coord_grid().reshape(1, -1, 3) [pt3d-recalled]:
    world coordinates of the voxel CENTRES, local [-1, 1] (align_corners=True) scaled by (R - 1) / 2 * voxel_size;
    the flattened index is d * R^2 + h * R + w with (x, y, z) <-> (w, h, d)."""
    voxel = volume_extent / resol
    half = 0.5 * (resol - 1) * voxel
    lin = torch.linspace(-1.0, 1.0, resol, dtype=dtype) * half
    z, y, x = torch.meshgrid(lin, lin, lin, indexing="ij")
    return torch.stack([x, y, z], -1).reshape(1, -1, 3)


def project_ndc(cams: ro.OracleCameras, pts: torch.Tensor, eps: float = 1e-2) -> torch.Tensor:
    """camera.transform_points(pts, eps)[..., :2] for NDC PerspectiveCameras [pt3d-recalled]: (n_cam, P, 2).
    Homogeneous result (fx X + px Z, fy Y + py Z, 1, Z) divided by sign(Z) max(|Z|, eps)."""
    cam = torch.einsum("pi,nij->npj", pts, cams.R) + cams.T[:, None]
    z = cam[..., 2]
    sign = z.sign() + (z == 0).to(z)
    den = sign * z.abs().clamp(min=eps)
    x = (cams.focal[:, None, 0] * cam[..., 0] + cams.pp[:, None, 0] * z) / den
    y = (cams.focal[:, None, 1] * cam[..., 1] + cams.pp[:, None, 1] * z) / den
    return torch.stack([x, y], -1)


def ndc_grid_sample(feat: torch.Tensor, xy_ndc: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """pytorch3d ndc_grid_sample(align_corners=False) [pt3d-recalled]: feat (n, C, H, W), xy_ndc (n, P, 2) -> (n, C, P)."""
    H, W = feat.shape[2:]
    g = -xy_ndc.clone()
    aspect = min(H, W) / max(H, W)
    if H >= W:
        g[..., 1] = g[..., 1] * aspect
    else:
        g[..., 0] = g[..., 0] * aspect
    return F.grid_sample(feat, g[:, :, None], mode=mode, align_corners=False, padding_mode="zeros")[..., 0]


def sample_views(cams: ro.OracleCameras, pts: torch.Tensor, feats: Dict[str, torch.Tensor], masks: Optional[torch.Tensor],
                 masked_sampling: bool = False, sampling_mode: str = "bilinear", view_weight: Optional[torch.Tensor] = None,
                 eps: float = 1e-2) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """ViewSampler.forward for ONE point batch [pt3d-recalled]: pts (P, 3) -> {k: (1, n_cam, P, C_k)}, (1, n_cam, P, 1).
    view_weight (n_cam) is the camera_pts_mask (sequence id of the camera == sequence id of the points)."""
    xy = project_ndc(cams, pts, eps)
    n = xy.shape[0]
    vw = torch.ones(n, dtype=pts.dtype, device=pts.device) if view_weight is None else view_weight.to(pts)
    out = {k: ndc_grid_sample(f, xy, sampling_mode).permute(0, 2, 1)[None] * vw[None, :, None, None] for k, f in feats.items()}
    if masked_sampling:
        m = ndc_grid_sample(masks, xy, "nearest").permute(0, 2, 1)[None]
    else:
        m = torch.ones(1, n, pts.shape[0], 1, dtype=pts.dtype, device=pts.device)
    return out, m * vw[None, :, None, None]


def point_to_camera_ray_dirs(cams: ro.OracleCameras, pts: torch.Tensor) -> torch.Tensor:
    """_get_point_to_source_camera_ray_dirs (custom_modules.py:283-334): (1, n_cam, P, 3), unit vectors from every
    camera centre -T R^T to every point."""
    centre = -torch.einsum("nj,nij->ni", cams.T, cams.R)
    return F.normalize(pts[None] - centre[:, None], dim=-1)[None]


def wmean(x: torch.Tensor, w: torch.Tensor, dim: int = 1, eps: float = 1e-2) -> torch.Tensor:
    """pytorch3d implicitron wmean as _avg_reduction_function calls it [pt3d-recalled]."""
    return (x * w[..., None]).sum(dim, keepdim=True) / w[..., None].sum(dim, keepdim=True).clamp(eps)


# ------------------------------------------------------------------------------------------------ in-tree code
ACT = {"relu": torch.relu, "leakyrelu": ro.leaky, "softplus": F.softplus, "identity": lambda x: x}


def mlp_mean_aggregate(p: Dict[str, torch.Tensor], feats_sampled: Dict[str, torch.Tensor], masks_sampled: torch.Tensor,
                       cams: ro.OracleCameras, pts: torch.Tensor, n_harmonic_ray: int = 3, hidden_activation: str = "leakyrelu",
                       last_activation: str = "softplus") -> torch.Tensor:
    """MLPMeanFeatureAggregator.forward (custom_modules.py:205-281) with exclude_target_view(_mask_features) = False, as
    HoloDiffusionModel.__post_init__ forces (holo_diffusion_model.py:115-116): -> (1, 1, P, dim_out).
    p: state dict of the aggregator (``_first_sampled``, ``_first_mean``, ``_mlp.mlp.{i}.0``, ``_last``)."""
    w = masks_sampled[..., 0]                                              # :241 (sampling mask = ones)
    ray = ro.harmonic_embedding(point_to_camera_ray_dirs(cams, pts), n_harmonic_ray)   # :243-245
    x = torch.cat([*feats_sampled.values(), ray], -1) * w[..., None]       # :248-256
    mean = wmean(x, w)                                                     # :257-262 (AVG)
    y = F.linear(x, p["_first_sampled.weight"], p["_first_sampled.bias"]) + \
        F.linear(mean, p["_first_mean.weight"], p["_first_mean.bias"])    # :263
    n_layers = len([k for k in p if k.startswith("_mlp.mlp.") and k.endswith(".weight")])
    for li in range(n_layers):   # MLPWithInputSkips (custom_modules.py:91-113,133-160): the LAST layer gets the hidden
        y = F.linear(y, p[f"_mlp.mlp.{li}.0.weight"], p[f"_mlp.mlp.{li}.0.bias"])   # activation, the others the "last" one
        y = ACT[hidden_activation if li == n_layers - 1 else last_activation](y)
    out = F.linear(y, p["_last.weight"], p["_last.bias"])                  # :264
    return (out * torch.softmax(out[..., :1], dim=1)).sum(dim=1, keepdim=True)   # :265-267


def angle_weighted_aggregate(feats_sampled: Dict[str, torch.Tensor], masks_sampled: torch.Tensor, cams: ro.OracleCameras,
                             pts: torch.Tensor, gamma: float = 1.0, min_ray_angle_weight: float = 0.1,
                             with_std: bool = True) -> torch.Tensor:
    """pytorch3d AngleWeightedReductionFeatureAggregator.forward with reduction_functions = (AVG, STD), no target-view
    exclusion [pt3d-recalled, UNPINNED]: -> (1, 1, P, sum_k 2 C_k), per feature [wmean | sqrt(clamp(wvar, 1e-4))].
    The angle weight compares the ray to every source camera with the ray to camera 0 (the point batch's own view)."""
    d = point_to_camera_ray_dirs(cams, pts)                               # (1, n, P, 3)
    dots = (d * d[:, :1]).sum(-1)
    w = masks_sampled[..., 0] * ((0.5 * (dots + 1.0)) ** gamma + min_ray_angle_weight)
    out = []
    for f in feats_sampled.values():
        mu = wmean(f, w)
        out.append(mu)
        if with_std:
            out.append(wmean((f - mu) ** 2, w).clamp(1e-4).sqrt())
    return torch.cat(out, -1)


def make_aggregator_params(in_dim: int, n_hidden: int = 128, dim_out: int = 128, n_layers: int = 1, seed: int = 0,
                           gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Xavier-uniform weights, zero biases on the lazy layers (custom_modules.py:37-41), small random biases elsewhere
    so that every bias path is exercised."""
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i):
        a = gain * math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def bias(o):
        return (torch.rand(o, generator=g) * 2 - 1) * 0.1

    p = {"_first_sampled.weight": xavier(n_hidden, in_dim), "_first_sampled.bias": bias(n_hidden),
         "_first_mean.weight": xavier(n_hidden, in_dim), "_first_mean.bias": bias(n_hidden)}
    for li in range(n_layers):
        p[f"_mlp.mlp.{li}.0.weight"], p[f"_mlp.mlp.{li}.0.bias"] = xavier(n_hidden, n_hidden), bias(n_hidden)
    p["_last.weight"], p["_last.bias"] = xavier(dim_out, n_hidden), bias(dim_out)
    return p


def encode(cams: ro.OracleCameras, feats: Dict[str, torch.Tensor], masks: Optional[torch.Tensor],
           agg: Optional[Dict[str, torch.Tensor]], mapper_w: torch.Tensor, mapper_b: torch.Tensor, resol: int,
           volume_extent: float, n_harmonic_ray: int = 3, masked_sampling: bool = False,
           view_weight: Optional[torch.Tensor] = None, pts: Optional[torch.Tensor] = None, chunk: int = 4096,
           angle_args: Optional[dict] = None, **act) -> torch.Tensor:
    """holo_diffusion_model.py:341-373: -> voxel_features (1, C, R, R, R) in [-1, 1] (or (1, C, 1, P) for explicit pts).
    agg: MLPMeanFeatureAggregator state dict, or None for the angle-weighted reduction (angle_args)."""
    grid_pts = coord_grid(resol, volume_extent)[0] if pts is None else pts
    rows = []
    for i in range(0, grid_pts.shape[0], chunk):
        pc = grid_pts[i:i + chunk]
        fs, ms = sample_views(cams, pc, feats, masks, masked_sampling, view_weight=view_weight)
        if agg is None:
            rows.append(angle_weighted_aggregate(fs, ms, cams, pc, **(angle_args or {})))
        else:
            rows.append(mlp_mean_aggregate(agg, fs, ms, cams, pc, n_harmonic_ray, **act))
    pooled = torch.cat(rows, 2)                                            # (1, 1, P, dim_out)
    v = F.linear(pooled, mapper_w, mapper_b).permute(0, 3, 1, 2)           # :368-369
    if pts is None:
        v = v.reshape(1, -1, resol, resol, resol)
    return torch.tanh(v)                                                   # :373


def make_views(n_src: int, image_hw: Tuple[int, int] = (40, 40), stage_channels: Sequence[int] = (16, 16, 16, 16),
               seed: int = 0, radius: float = 10.0, focal: float = 3.2):
    """Synthetic source views: cameras on the evaluation orbit and feature maps shaped like ResNetFeatureExtractor's
    output (four stages at 1/4 .. 1/32 of the rescaled image, L2-normalised * 1/2; the mask; the normalised image)."""
    g = torch.Generator().manual_seed(seed)
    cams = ro.simple_360_cameras(max(n_src, 2), radius=radius, focal_length=focal)[list(range(n_src))]
    H, W = image_hw
    feats: Dict[str, torch.Tensor] = {}
    for si, c in enumerate(stage_channels):
        s = 2 ** (si + 2)
        f = torch.randn(n_src, c, max(H // s, 1), max(W // s, 1), generator=g)
        feats[f"res_layer_{si + 1}"] = F.normalize(f, dim=1) * (1.0 / math.sqrt(len(stage_channels)))
    fg = torch.rand(n_src, 1, H, W, generator=g)
    feats["mask"] = fg
    feats["image"] = torch.randn(n_src, 3, H, W, generator=g)
    mask_crop = (torch.rand(n_src, 1, H, W, generator=g) > 0.3).float()
    return cams, feats, mask_crop
