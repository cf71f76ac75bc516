"""CPU restatement of the HoloDiffusion volumetric renderer (TEST INFRASTRUCTURE ONLY).

Pinning: the in-tree logic (RenderMLP / MLPWithInputSkips, HoloVoxelGridImplicitFunction.forward, the multi-pass
recursion) is pinned against the reference's own code run on ``oracle/pt3d_stub`` (tests/golden/render_intree_ref.npz);
the pytorch3d leaves (ray sampler, cameras, grid sampling, ray marcher, refiner, harmonic embedding) are
**parity unpinned**: pytorch3d 0.7.4 is not available; see ``oracle/__init__.py``.

Every function cites the reference call site (``/root/reference`` relative) that fixes its
arguments and, where the arithmetic lives in pytorch3d 0.7.4, the pytorch3d file it follows.
All functions are dtype- and device-generic (fp32 = the reference's arithmetic, fp64 = noise-free twin; CUDA tensors
run the same torch ops through cuDNN / cuBLAS -- used by the full-size parity test, never by the product).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# cameras  (pytorch3d/renderer/cameras.py: look_at_view_transform, PerspectiveCameras)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleCameras:
    """NDC PerspectiveCameras, row-vector convention X_cam = X_world @ R + T."""

    R: torch.Tensor  # (N,3,3)
    T: torch.Tensor  # (N,3)
    focal: torch.Tensor  # (N,2)
    pp: torch.Tensor  # (N,2)

    def __len__(self):
        return self.R.shape[0]

    def __getitem__(self, i):
        if isinstance(i, int):
            i = [i]
        return OracleCameras(self.R[i], self.T[i], self.focal[i], self.pp[i])

    def centre(self):
        # C = -T R^T   (pytorch3d cameras.get_camera_center)
        return -torch.einsum("nj,nij->ni", self.T, self.R)

    def unproject(self, xy_depth: torch.Tensor) -> torch.Tensor:
        """(N,P,3) of (x_ndc, y_ndc, depth) -> world points (N,P,3)."""
        x, y, z = xy_depth.unbind(-1)
        xc = (x - self.pp[:, None, 0]) * z / self.focal[:, None, 0]
        yc = (y - self.pp[:, None, 1]) * z / self.focal[:, None, 1]
        cam = torch.stack([xc, yc, z], -1)
        return torch.einsum("npj,nij->npi", cam - self.T[:, None], self.R)


def look_at_rotation_translation(dist, elev_deg, azim_deg, up=(0.0, 1.0, 0.0), dtype=torch.float32):
    """pytorch3d look_at_view_transform (degrees=True, at=(0,0,0))."""
    e = torch.as_tensor(elev_deg, dtype=dtype) * (math.pi / 180.0)
    a = torch.as_tensor(azim_deg, dtype=dtype) * (math.pi / 180.0)
    d = torch.as_tensor(dist, dtype=dtype)
    C = torch.stack([d * torch.cos(e) * torch.sin(a), d * torch.sin(e), d * torch.cos(e) * torch.cos(a)], -1)
    C = C.reshape(-1, 3)
    up_t = torch.as_tensor(up, dtype=dtype).reshape(1, 3).expand_as(C)
    z = F.normalize(-C, dim=-1, eps=1e-5)
    x = F.normalize(torch.cross(up_t, z, dim=-1), dim=-1, eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=-1), dim=-1, eps=1e-5)
    R = torch.stack([x, y, z], dim=1).transpose(1, 2)  # columns are x,y,z
    T = -torch.einsum("nji,nj->ni", R, C)  # -R^T C
    return R, T


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """pytorch3d/transforms/so3.py so3_exp_map (Rodrigues)."""
    nrms = (log_rot * log_rot).sum(-1)
    ang = nrms.clamp(eps).sqrt()
    inv = 1.0 / ang
    fac1 = inv * ang.sin()
    fac2 = inv * inv * (1.0 - ang.cos())
    K = torch.zeros(log_rot.shape[0], 3, 3, dtype=log_rot.dtype)
    x, y, z = log_rot.unbind(-1)
    K[:, 0, 1], K[:, 0, 2] = -z, y
    K[:, 1, 0], K[:, 1, 2] = z, -x
    K[:, 2, 0], K[:, 2, 1] = -y, x
    K2 = K @ K
    return fac1[:, None, None] * K + fac2[:, None, None] * K2 + torch.eye(3, dtype=log_rot.dtype)[None]


CANONICAL_CO3D_UP_AXIS = (-0.0396, -0.8306, -0.5554)  # visualize_reconstruction.py:35


def simple_360_cameras(n_poses: int, max_angle=2 * math.pi, elevation=-math.pi / 6.0, radius=10.0,
                       up=CANONICAL_CO3D_UP_AXIS, focal_length=3.2, canonical_up=(0.0, -1.0, 0.0),
                       dtype=torch.float32) -> OracleCameras:
    """holo_diffusion/utils/render_utils/flyaround.py:301-350 (defaults generate_samples.py:46-48)."""
    max_deg = 360.0 * max_angle / (2 * math.pi)
    elev_deg = 360.0 * elevation / (2 * math.pi)
    az = torch.linspace(0, max_deg, n_poses + 1)[:n_poses]
    R, T = look_at_rotation_translation(torch.full_like(az, radius), torch.full_like(az, elev_deg), az,
                                        up=canonical_up, dtype=dtype)
    axis = torch.cross(torch.tensor(canonical_up, dtype=dtype), torch.tensor(up, dtype=dtype), dim=-1)
    Rp = so3_exp_map(axis[None])[0]
    R = torch.bmm(Rp[None].expand_as(R), R)
    return OracleCameras(R, T, torch.full((n_poses, 2), focal_length, dtype=dtype),
                         torch.zeros(n_poses, 2, dtype=dtype))


# ----------------------------------------------------------------------------------------
# rays  (pytorch3d renderer/implicit/raysampling.py, implicitron ray_sampler.py;
#        configured configs/base.yaml:129-140, invoked holo_diffusion_model.py:442-448)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleRayBundle:
    origins: torch.Tensor  # (B, ..., 3)
    directions: torch.Tensor  # (B, ..., 3)
    lengths: torch.Tensor  # (B, ..., S)
    xys: torch.Tensor  # (B, ..., 2)


def ndc_xy_grid(H: int, W: int, dtype=torch.float32) -> torch.Tensor:
    """NDCMultinomialRaysampler pixel-centre grid, (H,W,2) of (x,y); +x left, +y up."""
    if W >= H:
        rx, ry = W / H, 1.0
    else:
        rx, ry = 1.0, H / W
    hx, hy = rx / W, ry / H
    xs = torch.linspace(rx - hx, -rx + hx, W, dtype=dtype)
    ys = torch.linspace(ry - hy, -ry + hy, H, dtype=dtype)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([X, Y], -1)


def depth_bounds(cams: OracleCameras, scene_extent: float, scene_center=(0.0, 0.0, 0.0)):
    """implicitron camera_utils.get_min_max_depth_bounds (AdaptiveRaySampler)."""
    c = cams.centre()
    sc = torch.tensor(scene_center, dtype=c.dtype)[None]
    cd = ((c - sc) ** 2).sum(-1).clamp(0.001).sqrt().clamp(0.001)
    cd = cd.clamp(scene_extent + 1e-3)
    return cd - scene_extent, cd + scene_extent


def sample_rays(cams: OracleCameras, H: int, W: int, S: int, scene_extent: float = 4.0) -> OracleRayBundle:
    """Full-grid, non-stratified evaluation rays.  Ray r = h*W + w (row-major)."""
    B = len(cams)
    dt = cams.R.dtype
    xy = ndc_xy_grid(H, W, dt).reshape(1, H * W, 2).expand(B, -1, -1)
    n = H * W
    one = torch.ones(B, n, 1, dtype=dt)
    pts = torch.cat([torch.cat([xy, one], -1), torch.cat([xy, 2 * one], -1)], 1)
    w = cams.unproject(pts)
    p1, p2 = w[:, :n], w[:, n:]
    d = p2 - p1
    o = p1 - d
    d = F.normalize(d, dim=-1)
    mn, mx = depth_bounds(cams, scene_extent)
    lengths = mn[:, None] + torch.linspace(0, 1, S, dtype=dt)[None] * (mx - mn)[:, None]
    lengths = lengths[:, None, :].expand(B, n, S)
    return OracleRayBundle(o.reshape(B, H, W, 3), d.reshape(B, H, W, 3),
                           lengths.reshape(B, H, W, S).contiguous(), xy.reshape(B, H, W, 2).contiguous())


# ----------------------------------------------------------------------------------------
# implicit function  (holo_voxel_grid_implicit_function.py:182-269, custom_modules.py:44-160)
# ----------------------------------------------------------------------------------------
def ray_points(b: OracleRayBundle) -> torch.Tensor:
    """pytorch3d ray_bundle_to_ray_points; call site holo_voxel_grid_implicit_function.py:199-201."""
    return b.origins[..., None, :] + b.lengths[..., :, None] * b.directions[..., None, :]


def world_to_local(p: torch.Tensor, resol: int, extent: float) -> torch.Tensor:
    """VolumeLocator(align_corners=True, voxel_size=extent/resol): holo_voxel_grid_implicit_function.py:204-209."""
    scale = (resol - 1) * (extent / resol) * 0.5
    return p * torch.tensor(1.0 / scale, dtype=p.dtype)


def sample_grid(grid: torch.Tensor, p_local: torch.Tensor) -> torch.Tensor:
    """FullResolutionVoxelGrid.evaluate_world -> interpolate_volume -> F.grid_sample
    (bilinear, zeros, align_corners=True); call site holo_voxel_grid_implicit_function.py:217-221.
    grid (1,C,D,H,W), p_local (P,3) in (x->W, y->H, z->D) order -> (P,C)."""
    out = F.grid_sample(grid, p_local.view(1, -1, 1, 1, 3), mode="bilinear", padding_mode="zeros",
                        align_corners=True)
    return out[0, :, :, 0, 0].t()


def harmonic_embedding(v: torch.Tensor, n: int) -> torch.Tensor:
    """pytorch3d HarmonicEmbedding(n, omega_0=1, logspace=True, append_input=True)."""
    if n == 0:
        return v
    freq = 2.0 ** torch.arange(n, dtype=v.dtype, device=v.device)
    e = (v[..., None] * freq).reshape(*v.shape[:-1], -1)
    return torch.cat([e.sin(), e.cos(), v], -1)


def leaky(x):
    return F.leaky_relu(x, 0.2)


def density_net(params: Dict[str, torch.Tensor], feats: torch.Tensor) -> torch.Tensor:
    """MLPWithInputSkips.forward (custom_modules.py:133-160) for RenderMLP._density_net: the hidden activation
    lands on the LAST layer only (:108-112), skip concat cat((y, z)) at layer 2.  Returns (P, hidden + 1)."""
    n_layers = 1 + max(int(k.split(".")[2]) for k in params if k.startswith("_density_net.mlp."))
    skips = (2,)
    y = feats
    for li in range(n_layers):
        if li in skips:
            y = torch.cat((y, feats), -1)
        y = F.linear(y, params[f"_density_net.mlp.{li}.0.weight"], params[f"_density_net.mlp.{li}.0.bias"])
        if li == n_layers - 1:
            y = leaky(y)
    return y


def render_mlp(params: Dict[str, torch.Tensor], feats: torch.Tensor, dirs: torch.Tensor, dir_emb: int = 4,
               return_head: bool = False):
    """RenderMLP.forward (holo_voxel_grid_implicit_function.py:107-129).

    params keys follow the reference state dict: ``_density_net.mlp.{i}.0.{weight,bias}``,
    ``_radiance_net.mlp.0.0.{weight,bias}``, optional ``_feature_net.mlp.0.0.{weight,bias}`` (the single-layer
    view-independent head, :94-105; being the last layer it gets the LeakyReLU).  feats (P,C), dirs (P,3) used as
    given.  Returns densities (P,1), rgb (P,3) [, head features (P,F) or None]."""
    y = density_net(params, feats)
    mlp_feats, dens = y[..., :-1], y[..., -1:]
    pe = harmonic_embedding(dirs, dir_emb)
    r = F.linear(torch.cat([mlp_feats, pe], -1), params["_radiance_net.mlp.0.0.weight"],
                 params["_radiance_net.mlp.0.0.bias"])
    rgb = torch.sigmoid(leaky(r))
    if not return_head:
        return dens, rgb
    head = None
    if "_feature_net.mlp.0.0.weight" in params:
        head = leaky(F.linear(mlp_feats, params["_feature_net.mlp.0.0.weight"], params["_feature_net.mlp.0.0.bias"]))
    return dens, rgb, head


def density_normals(params, grid, pts: torch.Tensor, resol: int, extent: float) -> torch.Tensor:
    """RenderMLP.get_normals (holo_voxel_grid_implicit_function.py:131-145): normalize(d density / d point) with
    autograd through grid_sample and the density net."""
    with torch.enable_grad():
        x = pts.clone()
        x.requires_grad = True
        f = sample_grid(grid, world_to_local(x.reshape(-1, 3), resol, extent))
        y = density_net(params, f)[..., -1:].sum()
        g = torch.autograd.grad(y, x)[0]
    return F.normalize(g, dim=-1)


def implicit_function(params, grid, bundle: Optional[OracleRayBundle], resol: int, extent: float,
                      render_normals: bool = False, pts_3d: Optional[torch.Tensor] = None):
    """HoloVoxelGridImplicitFunction.forward: returns densities (...,S,1), features (...,S,3[+F])
    [, normals (...,S,3) when render_normals]."""
    pts = ray_points(bundle) if pts_3d is None else pts_3d
    sp = pts.shape[:-1]
    f = sample_grid(grid, world_to_local(pts.reshape(-1, 3), resol, extent))
    dirs = bundle.directions if bundle is not None else torch.ones(*sp[:-1], 3, dtype=pts.dtype, device=pts.device)
    d = F.normalize(dirs, dim=-1)[..., None, :].expand(*sp, 3).reshape(-1, 3)
    dens, rgb, head = render_mlp(params, f, d, return_head=True)
    feats = rgb if head is None else torch.cat([rgb, head], -1)
    out = dens.reshape(*sp, 1), feats.reshape(*sp, -1)
    if render_normals:
        return out + (density_normals(params, grid, pts, resol, extent),)
    return out


# ----------------------------------------------------------------------------------------
# ray marcher + refiner  (pytorch3d implicitron raymarcher.py / ray_point_refiner.py / sample_pdf.py;
#                         configured configs/base.yaml:141-159; invoked holo_multipass_ea.py:96-116)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleRenderOut:
    features: torch.Tensor
    depths: torch.Tensor
    masks: torch.Tensor
    weights: Optional[torch.Tensor] = None
    prev_stage: Optional["OracleRenderOut"] = None
    lengths: Optional[torch.Tensor] = None
    normals: Optional[torch.Tensor] = None


def ea_raymarch(dens, feats, lengths, bg=(1.0, 1.0, 1.0), background_opacity=1e10, noise=None) -> OracleRenderOut:
    """EmissionAbsorptionRaymarcher(surface_thickness=1, replicate_last_interval=False,
    density_relu=True, blend_output=False)."""
    deltas = torch.cat([lengths[..., 1:] - lengths[..., :-1],
                        background_opacity * torch.ones_like(lengths[..., :1])], -1)
    d = dens[..., 0]
    if noise is not None:
        d = d + noise
    d = torch.relu(d)
    wd = deltas * d
    capped = 1.0 - torch.exp(-wd)
    opac = 1.0 - torch.exp(-torch.cumsum(wd, -1))
    mask = opac[..., -1:]
    absorb = (1.0 - opac).roll(1, -1).clone()
    absorb[..., :1] = 1.0
    w = capped * absorb
    f = (w[..., None] * feats).sum(-2)
    depth = (w * lengths)[..., None].sum(-2)
    f = f + (1 - mask) * torch.tensor(bg, dtype=f.dtype, device=f.device)
    return OracleRenderOut(f, depth, mask, w, lengths=lengths)


def sample_pdf(bins, weights, N: int, u: Optional[torch.Tensor] = None, eps: float = 1e-5):
    """pytorch3d sample_pdf_python; deterministic u=linspace(0,1,N) unless ``u`` given."""
    weights = weights + eps
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0.0, 1.0, N, dtype=weights.dtype, device=weights.device).expand(*cdf.shape[:-1], N)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = cdf.gather(-1, below), cdf.gather(-1, above)
    b0, b1 = bins.gather(-1, below), bins.gather(-1, above)
    den = c1 - c0
    den = torch.where(den < eps, torch.ones_like(den), den)
    t = (u - c0) / den
    return b0 + t * (b1 - b0)


def refine_lengths(lengths, weights, n_fine: int, add_input=True, u=None):
    """RayPointRefiner.forward (n_pts_per_ray_fine_evaluation, append_coarse_samples_to_fine)."""
    z = lengths
    mid = 0.5 * (z[..., 1:] + z[..., :-1])
    S = z.shape[-1]
    zs = sample_pdf(mid.reshape(-1, S - 1), weights.reshape(-1, S)[..., 1:-1], n_fine,
                    None if u is None else u.reshape(-1, n_fine)).reshape(*z.shape[:-1], n_fine)
    zz = torch.cat([z, zs], -1) if add_input else zs
    return torch.sort(zz, -1)[0]


def render_multipass(params, grid, bundle: OracleRayBundle, resol: int, extent: float, n_passes: int = 2,
                     n_fine: int = 16, bg=(1.0, 1.0, 1.0), render_normals: bool = False, noise=None,
                     u=None) -> OracleRenderOut:
    """HoloMultiPassEmissionAbsorptionRenderer._run_raymarcher (holo_multipass_ea.py:79-125).
    Evaluation mode by default; training mode = ``noise`` (list per pass of density_noise_std * randn, :87-91)
    and ``u`` (stratified refinement uniforms) supplied by the caller."""
    prev = None
    b = bundle
    for p in range(n_passes):
        o = implicit_function(params, grid, b, resol, extent, render_normals=render_normals)
        dens, feats = o[0], o[1]
        out = ea_raymarch(dens, feats, b.lengths, bg, noise=None if noise is None else noise[p])
        if render_normals:
            out.normals = (o[2] * out.weights[..., None]).sum(-2)  # holo_multipass_ea.py:104-109
        out.prev_stage = prev
        prev = out
        if p + 1 < n_passes:
            b = OracleRayBundle(b.origins, b.directions, refine_lengths(b.lengths, out.weights, n_fine, u=u), b.xys)
    return out


def render_chunked(params, grid, bundle: OracleRayBundle, resol, extent, n_passes=2, n_fine=16,
                   chunk_size_grid: int = 4096, bg=(1.0, 1.0, 1.0)) -> OracleRenderOut:
    """GenericModel._render chunk loop (entered holo_diffusion_model.py:451-457): flatten rays,
    n_chunks = ceil(n_rays*S/chunk), rays_per_chunk = ceil(n_rays/n_chunks), slice, cat, reshape."""
    B = bundle.origins.shape[0]
    sp = bundle.origins.shape[1:-1]
    n = int(math.prod(sp))
    S = bundle.lengths.shape[-1]
    flat = OracleRayBundle(bundle.origins.reshape(B, n, 3), bundle.directions.reshape(B, n, 3),
                           bundle.lengths.reshape(B, n, S), bundle.xys.reshape(B, n, 2))
    n_chunks = max(1, math.ceil(n * S / chunk_size_grid)) if chunk_size_grid > 0 else 1
    per = math.ceil(n / n_chunks)
    outs: List[OracleRenderOut] = []
    for a in range(0, n, per):
        sl = slice(a, min(n, a + per))
        outs.append(render_multipass(params, grid, OracleRayBundle(flat.origins[:, sl], flat.directions[:, sl],
                                                                   flat.lengths[:, sl], flat.xys[:, sl]),
                                     resol, extent, n_passes, n_fine, bg))

    def cat(stage_list):
        if stage_list[0] is None:
            return None
        o = OracleRenderOut(
            torch.cat([s.features for s in stage_list], 1).reshape(B, *sp, -1),
            torch.cat([s.depths for s in stage_list], 1).reshape(B, *sp, -1),
            torch.cat([s.masks for s in stage_list], 1).reshape(B, *sp, -1),
            torch.cat([s.weights for s in stage_list], 1).reshape(B, *sp, -1),
            lengths=torch.cat([s.lengths for s in stage_list], 1).reshape(B, *sp, -1))
        o.prev_stage = cat([s.prev_stage for s in stage_list])
        return o

    return cat(outs)


# ----------------------------------------------------------------------------------------
# parameter fixtures
# ----------------------------------------------------------------------------------------
def make_render_mlp_params(in_dims: int, seed: int = 1, hidden: int = 256, dir_emb: int = 4,
                           density_scale: float = 8.0, density_bias: float = 0.5, dtype=torch.float32):
    """Xavier-uniform weights (pytorch3d _xavier_init) + default nn.Linear biases, then the density
    row scaled/biased so compositing is exercised (SURVEY.md section 4 testing traps)."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(hidden, in_dims), (hidden, hidden), (hidden, hidden + in_dims), (hidden + 1, hidden)]
    p = {}

    def lin(o, i):
        bound_w = math.sqrt(6.0 / (i + o))
        w = (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * bound_w
        bound_b = 1.0 / math.sqrt(i)
        b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * bound_b
        return w.to(dtype), b.to(dtype)

    for li, (o, i) in enumerate(shapes):
        w, b = lin(o, i)
        p[f"_density_net.mlp.{li}.0.weight"], p[f"_density_net.mlp.{li}.0.bias"] = w, b
    p["_density_net.mlp.3.0.weight"][-1] *= density_scale
    p["_density_net.mlp.3.0.bias"][-1] += density_bias
    w, b = lin(3, hidden + 3 * (2 * dir_emb + 1))
    p["_radiance_net.mlp.0.0.weight"], p["_radiance_net.mlp.0.0.bias"] = w, b
    return p
