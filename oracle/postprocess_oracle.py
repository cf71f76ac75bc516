"""CPU restatement of the per-view post-processing (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

  * make_depth_image: pytorch3d==0.7.4 implicitron/tools/vis_utils.py [pt3d-recalled; the library is absent from
    /root/reference and from this image => PARITY UNPINNED for this function], called from
    /root/reference/holo_diffusion/utils/render_utils/flyaround.py:470-477.
  * depth_frame: the compositing / channel repeat of _images_from_preds (flyaround.py:476-479).
  * frame_u8: clip + resize + 8-bit of _generate_prediction_videos (flyaround.py:588-595); the resize is torch's
    bilinear F.interpolate(align_corners=False) (PIL is the reference's resampler: frames differ by resampling kernel).
  * shade_depth: depth_to_shaded(method="mesh") (/root/reference/holo_diffusion/utils/render_utils/
    shaded_depth_render.py:143-206): _smooth_depth :15-24 and get_grid_mesh :248-280 follow the in-tree code; the
    render of that mesh from its own camera (mesh_render.py -> pytorch3d rasteriser + SoftGouraudShader, absent) is
    restated as per-vertex Gouraud shading read back at the vertex's own pixel: explicit triangle list, face normals
    scattered onto the vertices (pytorch3d Meshes.verts_normals: area-weighted), PointLights defaults.
"""
import math

import torch
import torch.nn.functional as F


def make_depth_image(depths, masks, max_quantile=0.98, min_quantile=0.02, min_out_depth=0.1, max_out_depth=0.9):
    """depths, masks (B,1,H,W) -> (B,1,H,W)."""
    normfacs = []
    for d, m in zip(depths, masks):
        ok = (d.reshape(-1) > 1e-6) * (m.reshape(-1) > 0.5)
        if ok.sum() <= 1:
            normfacs.append(torch.zeros(2).type_as(depths))
            continue
        dok = d.reshape(-1)[ok].reshape(-1)
        _maxk = max(int(round((1 - max_quantile) * (dok.numel()))), 1)
        _mink = max(int(round(min_quantile * (dok.numel()))), 1)
        normfac_max = dok.topk(k=_maxk, dim=-1).values[-1]
        normfac_min = dok.topk(k=_mink, dim=-1, largest=False).values[-1]
        normfacs.append(torch.stack([normfac_min, normfac_max]))
    normfacs = torch.stack(normfacs)
    _min, _max = normfacs[:, 0].view(-1, 1, 1, 1), normfacs[:, 1].view(-1, 1, 1, 1)
    depths = (depths - _min) / (_max - _min).clamp(1e-4)
    depths = ((depths * (max_out_depth - min_out_depth) + min_out_depth) * masks.float()).clamp(0.0, 1.0)
    return depths, normfacs


def depth_frame(depths, masks):
    v, nf = make_depth_image(depths, masks)
    v = v * masks + (1 - masks)
    return v.repeat(1, 3, 1, 1), nf


def frame_u8(img_chw, out_hw=None):
    x = img_chw.clamp(0.0, 1.0)[None]
    if x.shape[1] == 1:
        x = x.repeat(1, 3, 1, 1)
    if out_hw is not None and tuple(out_hw) != tuple(x.shape[-2:]):
        x = F.interpolate(x, size=tuple(out_hw), mode="bilinear", align_corners=False)
    return torch.round(x[0].clamp(0.0, 1.0) * 255.0).to(torch.uint8).permute(1, 2, 0).contiguous()


def smooth_depth(g, m, k):
    gm = torch.cat((g, m.float()), dim=1)
    gma = F.avg_pool2d(gm, 2 * k + 1, padding=k, stride=1)
    g, m = gma.split([gma.shape[1] - 1, 1], dim=1)
    return g / m.clamp(1e-4)


def grid_mesh_faces(mask_hw):
    he, wi = mask_hw.shape
    idx = torch.arange(he * wi).reshape(he, wi)
    fq = F.unfold(idx[None, None].float(), 2).long()[0]            # (4, n_quads): a, b, c, d
    mq = F.unfold((mask_hw[None, None] > 0.5).float(), 2).long()[0]
    fq = fq[:, mq.sum(0) == 4]
    tri1 = fq[:3].T[:, [0, 2, 1]]                                   # (a, c, b)
    tri2 = fq[1:].T[:, [0, 2, 1]][:, [0, 2, 1]]                     # (b, c, d) after the two swaps of the reference
    return torch.cat((tri1, tri2), dim=0)


def shade_depth(depth_hw, mask_hw, focal, pp, smoothing_kernel_size=0.005, mask_thr=0.5, depth_thr=1e-2,
                material=((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 0.9), 128.0), bg=(1.0, 1.0, 1.0),
                light=(0.5, 0.3, 0.2)):
    H, W = depth_hw.shape
    d = depth_hw.double()[None, None]
    ok = (mask_hw[None, None] > mask_thr) * (d > depth_thr)
    k = int(math.ceil(smoothing_kernel_size * math.sqrt(H ** 2 + W ** 2)))
    ds = smooth_depth(d, ok, k)[0, 0]
    rx, ry = (W / H, 1.0) if W >= H else (1.0, H / W)
    xs = rx - 2 * rx * (torch.arange(W, dtype=torch.float64) + 0.5) / W
    ys = ry - 2 * ry * (torch.arange(H, dtype=torch.float64) + 0.5) / H
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    verts = torch.stack([(xx - pp[0]) / focal[0] * ds, (yy - pp[1]) / focal[1] * ds, ds], -1).reshape(-1, 3)
    faces = grid_mesh_faces(ok[0, 0].float())
    out = torch.tensor(bg, dtype=torch.float64).view(3, 1).repeat(1, H * W)
    used = torch.zeros(H * W, dtype=torch.bool)
    if faces.shape[0]:
        v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
        fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
        vn = torch.zeros_like(verts)
        for j in range(3):
            vn.index_add_(0, faces[:, j], fn)
            used[faces[:, j]] = True
        vn = F.normalize(vn, dim=-1)
        l = F.normalize(-verts, dim=-1)
        flip = (vn * l).sum(-1, keepdim=True) < 0
        vn = torch.where(flip, -vn, vn)
        ndl = (vn * l).sum(-1).clamp(0.0)
        spec = (2 * ndl * ndl - 1).clamp(0.0) ** material[3]
        amb, dif, spc = (torch.tensor(c, dtype=torch.float64) for c in material[:3])
        col = (amb * light[0])[:, None] + (dif * light[1])[:, None] * ndl[None] + (spc * light[2])[:, None] * spec[None]
        out[:, used] = col.clamp(0.0, 1.0)[:, used]
    return out.view(3, H, W).float(), used.view(H, W).float(), k
