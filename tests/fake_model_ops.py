"""Torch / oracle stand-ins for the C-ABI entry points HoloDiffusionModel.forward and the fly-around post-processing
reach, so that the HOST logic (plug-in assembly, registry facades, flyaround loop, frame packing order) runs on a CPU
box.  Test infrastructure: the kernels themselves are checked by the -m gpu tests."""
import struct

import torch

import fake_unet_ops
from oracle import postprocess_oracle as po
from oracle import render_oracle as ro


def _enc(x: float) -> int:
    i = struct.unpack("i", struct.pack("f", float(x)))[0]
    return i ^ 0x7FFFFFFF if i < 0 else i


def range_init(stats):
    stats[0], stats[1], stats[2], stats[3] = _enc(float("inf")), _enc(float("-inf")), 0, 0


def _dec(i: int) -> float:
    i = int(i)
    if i < 0:
        i ^= 0x7FFFFFFF
    return struct.unpack("f", struct.pack("i", i))[0]


def act_range(x_cl, V, C, act, y_cl, y_cf, stats):
    y = torch.tanh(x_cl) if act == 1 else x_cl
    if stats is not None:
        stats[0] = _enc(min(_dec(stats[0]), float(y.min())))
        stats[1] = _enc(max(_dec(stats[1]), float(y.max())))
        stats[2] += int(torch.isnan(y).sum())
    if y_cl is not None:
        y_cl.copy_(y)
    if y_cf is not None:
        y_cf.copy_(y.t().reshape(-1))


def raygen(R, T, focal, pp, xy, S, scene_extent, scene_center=(0.0, 0.0, 0.0)):
    os_, ds_, ls_ = [], [], []
    for i in range(R.shape[0]):
        cam = ro.OracleCameras(R[i:i + 1], T[i:i + 1], focal[i:i + 1], pp[i:i + 1])
        z = torch.ones(xy.shape[0], 1)
        p1 = cam.unproject(torch.cat([xy, z], -1)[None])[0]
        p2 = cam.unproject(torch.cat([xy, 2 * z], -1)[None])[0]
        d = p2 - p1
        o = p1 - d
        mn, mx = ro.depth_bounds(cam, scene_extent, scene_center)
        ln = float(mn) + torch.linspace(0, 1, S) * (float(mx) - float(mn))
        os_.append(o), ds_.append(torch.nn.functional.normalize(d, dim=-1)), ls_.append(ln[None].expand(xy.shape[0], S))
    return torch.stack(os_), torch.stack(ds_), torch.stack(ls_).contiguous()


def collapse(layers, skips, rw, rb, C):
    p = {}
    for i, (w, b) in enumerate(layers):
        p[f"_density_net.mlp.{i}.0.weight"], p[f"_density_net.mlp.{i}.0.bias"] = w, b
    p["_radiance_net.mlp.0.0.weight"], p["_radiance_net.mlp.0.0.bias"] = rw, rb
    H = layers[-1][0].shape[0] - 1
    return p, H, rw.shape[1] - H, None


def render_fwd(grid_dhwc, volume_extent, packed, hidden, n_harm, origins, dirs, lengths, n_passes=1, n_fine=0,
               add_input_samples=True, bg=(1.0, 1.0, 1.0), background_opacity=1e10, return_weights=False, return_prev=True,
               tc_image=None):
    grid = grid_dhwc.permute(3, 0, 1, 2)[None]
    b = ro.OracleRayBundle(origins[None], dirs[None], lengths[None], torch.zeros(1, origins.shape[0], 2))
    o = ro.render_multipass(packed, grid, b, grid.shape[-1], volume_extent, n_passes, n_fine, bg)

    def d(x):
        return None if x is None else {"features": x.features[0], "depths": x.depths[0], "masks": x.masks[0],
                                       "weights": x.weights[0] if return_weights else None, "lengths": x.lengths[0]}

    out = d(o)
    out["prev"] = d(o.prev_stage) if return_prev else None
    return out


def transpose2d(src, rows, cols, out=None):
    r = src.view(rows, cols).t().contiguous().view(-1)
    if out is not None:
        out.copy_(r)
        return out
    return r


def depth_image(depth, mask, min_quantile=0.02, max_quantile=0.98, min_out=0.1, max_out=0.9, composite_white=True):
    v, nf = po.make_depth_image(depth[None, None], mask[None, None], max_quantile, min_quantile, min_out, max_out)
    if composite_white:
        v = v * mask[None, None] + (1 - mask[None, None])
    return v[0].repeat(3, 1, 1), nf[0]


def frame_u8(src_chw, out_hw=None, out=None):
    f = po.frame_u8(src_chw, out_hw)
    if out is not None:
        out.copy_(f)
        return out
    return f


def shade_depth(depth, mask, focal, pp, smooth_k, mask_thr=0.5, depth_thr=1e-2, material="medium", bg=(1.0, 1.0, 1.0),
                light=(0.5, 0.3, 0.2)):
    from holo_diffusion_b200.ops import MATERIALS
    o, m, _ = po.shade_depth(depth, mask, focal, pp, mask_thr=mask_thr, depth_thr=depth_thr, material=MATERIALS[material],
                             bg=bg, light=light)
    return o, m


def ddpm_step(model_out, x_t, noise, t, coef1, coef2, logvar, clip, x_prev, pred_x0=None):
    sh = [-1] + [1] * (x_t.ndim - 1)
    x0 = model_out.clamp(-1, 1) if clip else model_out
    mean = coef1[t].view(sh) * x0 + coef2[t].view(sh) * x_t
    if noise is not None:
        mean = mean + (t != 0).float().view(sh) * torch.exp(0.5 * logvar[t].view(sh)) * noise
    x_prev.copy_(mean)
    if pred_x0 is not None:
        pred_x0.copy_(x0)


def install(ops_module, setattr_fn=setattr):
    """Replace every entry point the model / fly-around path calls; also lifts the CUDA-only guard."""
    fake_unet_ops.install(ops_module, setattr_fn)
    for name, fn in dict(range_init=range_init, act_range=act_range, raygen=raygen, collapse_and_pack_render_mlp=collapse,
                         render_fwd=render_fwd, transpose2d=transpose2d, depth_image=depth_image, frame_u8=frame_u8,
                         shade_depth=shade_depth, ddpm_step=ddpm_step, require_cuda=lambda *a, **k: None).items():
        setattr_fn(ops_module, name, fn)
