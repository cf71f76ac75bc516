"""Torch / oracle stand-ins for the view-pooling entry points (holo_viewpool_*), written against the kernels' CONTRACT
in include/holo_b200.h (row layouts, padding, what each output holds), so that the HOST logic of
holo_diffusion_b200/encoder.py -- weight folding and packing, chunking, padding, activation placement, the mapper, the
model's encoder branch -- runs on a CPU box.  Test infrastructure; the kernels themselves are checked by the -m gpu tests."""
import torch
import torch.nn.functional as F

import fake_model_ops
import fake_unet_ops
from oracle import encoder_oracle as eo
from oracle import render_oracle as ro

ACTS = {"identity": lambda x: x, "relu": torch.relu, "leakyrelu": ro.leaky, "softplus": F.softplus}


def _sampled(pts, R, T, focal, pp, maps_cl, mask_map, view_weight, eps):
    cams = ro.OracleCameras(R, T, focal, pp)
    feats = {str(k): m.permute(0, 3, 1, 2) for k, m in enumerate(maps_cl)}
    masks = None if mask_map is None else mask_map[:, None]
    fs, ms = eo.sample_views(cams, pts, feats, masks, masked_sampling=mask_map is not None, view_weight=view_weight, eps=eps)
    return cams, fs, ms


def viewpool_sample(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, n_harmonic, Kpad, rows_per_view, x_hi, x_lo, mean_hi,
                    mean_lo, mask_map=None, view_weight=None, eps=1e-2, x_f32=None, mean_f32=None):
    cams, fs, ms = _sampled(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, mask_map, view_weight, eps)
    w = ms[0, ..., 0]                                                        # (n_src, P)
    ray = ro.harmonic_embedding(eo.point_to_camera_ray_dirs(cams, pts), n_harmonic)[0]
    x = torch.cat([*[f[0] for f in fs.values()], ray], -1) * w[..., None]    # (n_src, P, Kx)
    mean = (x * w[..., None]).sum(0) / w.sum(0).clamp(1e-2)[:, None]
    n_src, P, Kx = x.shape
    xv, xl = x_hi.view(n_src, rows_per_view, Kpad), x_lo.view(n_src, rows_per_view, Kpad)
    xv[:, :P] = 0
    xv[:, :P, :Kx] = x
    xl[:, :P] = 0
    mean_hi[:P] = 0
    mean_hi[:P, :Kx] = mean
    mean_lo[:P] = 0
    if x_f32 is not None:
        x_f32.copy_(x)
    if mean_f32 is not None:
        mean_f32.copy_(mean)


def viewpool_angle_reduce(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, Kpad, out_hi, out_lo, mask_map=None,
                          view_weight=None, eps=1e-2, gamma=1.0, min_ray_angle_weight=0.1, with_std=True, out_f32=None):
    cams, fs, ms = _sampled(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, mask_map, view_weight, eps)
    g = eo.angle_weighted_aggregate(fs, ms, cams, pts, gamma, min_ray_angle_weight, with_std)[0, 0]
    P = pts.shape[0]
    out_hi[:P] = 0
    out_hi[:P, : g.shape[1]] = g
    out_lo[:P] = 0
    if out_f32 is not None:
        out_f32.copy_(g)


def viewpool_act_split(y, point_term, n_views, rows_per_view, C, act, hi, lo):
    v = y.view(n_views, rows_per_view, C)
    if point_term is not None:
        v = v + point_term[None]
    hi.copy_(ACTS[act](v).reshape(hi.shape))
    lo.zero_()


def viewpool_reduce(z, n_views, rows_per_view, n_pts, C, out=None, out_hi=None, out_lo=None):
    v = z.view(n_views, rows_per_view, C)[:, :n_pts]
    g = (v * torch.softmax(v[..., :1], dim=0)).sum(0)
    if out is not None:
        out[:n_pts] = g
    if out_hi is not None:
        out_hi[:n_pts] = g
        out_lo[:n_pts] = 0


def gemm_tc_act(a_hi, a_lo, a_pitch, M, K, b_hi, b_lo, b_pitch, N, bias, row_term, row_term_rows, act, out_pitch, out=None,
                out_hi=None, out_lo=None, acc_scale=1.0):
    a = fake_unet_ops._view(a_hi, 0, a_pitch, M, K) + fake_unet_ops._view(a_lo, 0, a_pitch, M, K)
    b = fake_unet_ops._view(b_hi, 0, b_pitch, N, K) + fake_unet_ops._view(b_lo, 0, b_pitch, N, K)
    y = (a @ b.t()) * acc_scale
    if bias is not None:
        y = y + bias
    if row_term is not None:
        y = y + row_term[torch.arange(M) % row_term_rows]
    y = ACTS[act](y)
    if out is not None:
        fake_unet_ops._view(out, 0, out_pitch, M, N).copy_(y)
    if out_hi is not None:
        fake_unet_ops._view(out_hi, 0, out_pitch, M, N).copy_(y)
        out_lo.zero_()
    return 0


ALL = dict(gemm_tc_act=gemm_tc_act, viewpool_sample=viewpool_sample, viewpool_angle_reduce=viewpool_angle_reduce, viewpool_act_split=viewpool_act_split,
           viewpool_reduce=viewpool_reduce, gemm_tc=fake_unet_ops.gemm_tc, act_range=fake_model_ops.act_range,
           require_cuda=lambda device, who: None)


def install(ops_module, setattr_fn=setattr):
    for name, fn in ALL.items():
        setattr_fn(ops_module, name, fn)
