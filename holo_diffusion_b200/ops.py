"""Torch-tensor front end of the C-ABI (device pointers + sizes go straight to libholo_b200.so).

PyTorch is used for device memory and streams only; every arithmetic op on the hot path is one of the CUDA
kernels behind ``include/holo_b200.h``.  All tensors must be CUDA, contiguous and of the stated dtype --
violations raise instead of silently copying.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from ._lib import HoloError, lib

_f = ctypes.c_float
_i = ctypes.c_int
_ll = ctypes.c_longlong


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32, name: str = "tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise HoloError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise HoloError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise HoloError(f"{name}: expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def _ptr16(t: Optional[torch.Tensor], name: str = "tensor"):
    """Pointer to one half of a tensor-core operand pair, bf16 or fp16 (the dtype selects the format flag)."""
    return None if t is None else _ptr(t, t.dtype if t.dtype in (torch.bfloat16, torch.float16) else torch.bfloat16, name)


def _pair_f16(*ts) -> int:
    """1 if the operand halves passed are fp16, 0 if bf16; one call cannot mix them (nor can one UMMA)."""
    fl = {t.dtype == torch.float16 for t in ts if t is not None}
    if len(fl) > 1:
        raise HoloError("the halves of the operand pairs of one call must share a dtype (all bf16 or all fp16)")
    return 1 if fl == {True} else 0


def require_cuda(device, who: str):
    """There is no CPU fallback: anything that is not a CUDA device is an error (host-logic tests replace this
    predicate together with the entry points they stand in for)."""
    if torch.device(device).type != "cuda":
        raise HoloError(f"{who}: CUDA tensors only (no CPU fallback)")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _host3(v: Sequence[float]):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


# ------------------------------------------------------------------ renderer
def raygen(R, T, focal, pp, xy, S: int, scene_extent: float, scene_center=(0.0, 0.0, 0.0)):
    n_cam, n_rays = R.shape[0], xy.shape[0]
    dev = R.device
    origins = torch.empty(n_cam, n_rays, 3, device=dev)
    dirs = torch.empty(n_cam, n_rays, 3, device=dev)
    lengths = torch.empty(n_cam, n_rays, S, device=dev)
    lib().call("holo_raygen", _ptr(R), _ptr(T), _ptr(focal), _ptr(pp), _ptr(xy), n_cam, n_rays, S,
               float(scene_extent), ctypes.cast(_host3(scene_center), ctypes.c_void_p), _ptr(origins), _ptr(dirs),
               _ptr(lengths), _stream())
    return origins, dirs, lengths


def collapse_and_pack_render_mlp(density_layers, skips, radiance_w, radiance_b, C: int):
    """density_layers: list of (weight, bias) fp32 CUDA tensors; returns the packed fp32 buffer."""
    dev = radiance_w.device
    A = c = None
    rows = 0
    for li, (W, b) in enumerate(density_layers):
        out_dim, in_total = W.shape
        skip = 1 if (li in skips and li > 0) else 0
        A_out = torch.empty(out_dim, C, dtype=torch.float64, device=dev)
        c_out = torch.empty(out_dim, dtype=torch.float64, device=dev)
        lib().call("holo_affine_compose_f64", _ptr(W), _ptr(b), out_dim, in_total,
                   _ptr(A, torch.float64), _ptr(c, torch.float64), rows, C, skip,
                   _ptr(A_out, torch.float64), _ptr(c_out, torch.float64), _stream())
        A, c, rows = A_out, c_out, out_dim
    H = rows - 1
    E = radiance_w.shape[1] - H
    n = lib().cdll.holo_render_mlp_packed_floats(H, C, E)
    packed = torch.empty(n, device=dev)
    lib().call("holo_pack_render_mlp", _ptr(A, torch.float64), _ptr(c, torch.float64), _ptr(radiance_w),
               _ptr(radiance_b), H, C, E, _ptr(packed), _stream())
    tc_image = None
    if H == 256 and C in (16, 32) and E <= 27:
        tc_image = torch.empty(lib().cdll.holo_render_tc_image_bytes(), dtype=torch.uint8, device=dev)
        lib().call("holo_pack_render_mlp_tc", _ptr(A, torch.float64), _ptr(c, torch.float64), _ptr(radiance_w),
                   _ptr(radiance_b), H, C, E, _ptr(tc_image, torch.uint8), _stream())
    return packed, H, E, tc_image


def render_fwd(grid_dhwc, volume_extent: float, packed_mlp, hidden: int, n_harmonic: int, origins, dirs, lengths,
               n_passes: int = 1, n_fine: int = 0, add_input_samples: bool = True, bg=(1.0, 1.0, 1.0),
               background_opacity: float = 1e10, return_weights: bool = False, return_prev: bool = True, tc_image=None):
    """tc_image given -> the tcgen05 kernel (holo_render_fwd_tc), else the fp32 CUDA-core kernel (holo_render_fwd)."""
    D, H, W, C = grid_dhwc.shape
    n_rays, S = lengths.shape
    dev = grid_dhwc.device
    S_last = S if n_passes == 1 else (S + n_fine if add_input_samples else n_fine)
    out = {
        "features": torch.empty(n_rays, 3, device=dev),
        "depths": torch.empty(n_rays, 1, device=dev),
        "masks": torch.empty(n_rays, 1, device=dev),
        "weights": torch.empty(n_rays, S_last, device=dev) if return_weights else None,
        "lengths": torch.empty(n_rays, S_last, device=dev) if n_passes > 1 else lengths,
    }
    prev = None
    if n_passes > 1 and return_prev:
        prev = {
            "features": torch.empty(n_rays, 3, device=dev),
            "depths": torch.empty(n_rays, 1, device=dev),
            "masks": torch.empty(n_rays, 1, device=dev),
            "weights": torch.empty(n_rays, S, device=dev) if return_weights else None,
            "lengths": lengths,
        }
    tail = (_ptr(out["features"]), _ptr(out["depths"]), _ptr(out["masks"]), _ptr(out["weights"]),
            _ptr(out["lengths"]) if n_passes > 1 else None,
            _ptr(prev["features"]) if prev else None, _ptr(prev["depths"]) if prev else None,
            _ptr(prev["masks"]) if prev else None, _ptr(prev["weights"]) if prev else None)
    if tc_image is not None:
        scratch = torch.empty(S * n_rays, device=dev) if n_passes > 1 else None
        lib().call("holo_render_fwd_tc", _ptr(grid_dhwc), D, H, W, C, float(volume_extent), _ptr(tc_image, torch.uint8),
                   n_harmonic, _ptr(origins), _ptr(dirs), _ptr(lengths), n_rays, S, n_passes, n_fine,
                   1 if add_input_samples else 0, ctypes.cast(_host3(bg), ctypes.c_void_p), float(background_opacity),
                   *tail, _ptr(scratch), _stream())
    else:
        lib().call("holo_render_fwd", _ptr(grid_dhwc), D, H, W, C, float(volume_extent), _ptr(packed_mlp), hidden,
                   n_harmonic, _ptr(origins), _ptr(dirs), _ptr(lengths), n_rays, S, n_passes, n_fine,
                   1 if add_input_samples else 0, ctypes.cast(_host3(bg), ctypes.c_void_p), float(background_opacity),
                   *tail, _stream())
    out["prev"] = prev
    return out


def _hostf(v: Sequence[float]):
    return (ctypes.c_float * len(v))(*[float(x) for x in v])


def if_fwd(grid_dhwc, volume_extent: float, packed_mlp, hidden: int, n_harmonic: int, *, origins=None, dirs=None,
           lengths=None, pts_3d=None, S: int = 1, head=None, normals: bool = False):
    """HoloVoxelGridImplicitFunction.forward on n_points = n_rays * S points (holo_if_fwd).
    Either (origins, dirs, lengths (n_rays,S)) or pts_3d (n_points,3) [+ dirs (n_rays,3) or None = dummy ones].
    head = (weight (F,hidden), bias (F)) of the view-independent feature net or None."""
    D, H, W, C = grid_dhwc.shape
    dev = grid_dhwc.device
    if pts_3d is not None:
        n_points = pts_3d.shape[0]
    else:
        n_points = lengths.shape[0] * lengths.shape[1]
        S = lengths.shape[1]
    F = 0 if head is None else head[0].shape[0]
    dens = torch.empty(n_points, device=dev)
    feats = torch.empty(n_points, 3 + F, device=dev)
    nrm = torch.empty(n_points, 3, device=dev) if normals else None
    lib().call("holo_if_fwd", _ptr(grid_dhwc), D, H, W, C, float(volume_extent), _ptr(packed_mlp), hidden, n_harmonic,
               _ptr(head[0]) if head else None, _ptr(head[1]) if head else None, F, _ptr(origins), _ptr(dirs),
               _ptr(lengths), _ptr(pts_3d), n_points, S, _ptr(dens), _ptr(feats), _ptr(nrm), _stream())
    return dens, feats, nrm


def render_mlp_fwd(feats, view_dirs, packed_mlp, hidden: int, n_harmonic: int, head=None):
    """RenderMLP.forward on (n_points, C) features and (n_points, 3) view directions (holo_render_mlp_fwd)."""
    n_points, C = feats.shape
    dev = feats.device
    F = 0 if head is None else head[0].shape[0]
    dens = torch.empty(n_points, device=dev)
    out = torch.empty(n_points, 3 + F, device=dev)
    lib().call("holo_render_mlp_fwd", _ptr(feats), _ptr(view_dirs), n_points, C, _ptr(packed_mlp), hidden, n_harmonic,
               _ptr(head[0]) if head else None, _ptr(head[1]) if head else None, F, _ptr(dens), _ptr(out), _stream())
    return dens, out


def ea_raymarch(densities, features, lengths, bg: Sequence[float], background_opacity: float = 1e10, noise=None,
                normals=None):
    """EmissionAbsorptionRaymarcher on (n_rays,S) densities, (n_rays,S,Fd) features (holo_ea_raymarch)."""
    n_rays, S = lengths.shape
    Fd = features.shape[-1]
    dev = lengths.device
    out = {"features": torch.empty(n_rays, Fd, device=dev), "depths": torch.empty(n_rays, 1, device=dev),
           "masks": torch.empty(n_rays, 1, device=dev), "weights": torch.empty(n_rays, S, device=dev),
           "normals": torch.empty(n_rays, 3, device=dev) if normals is not None else None}
    bgv = _hostf(bg)
    lib().call("holo_ea_raymarch", _ptr(densities), _ptr(features), _ptr(lengths), _ptr(noise), _ptr(normals), n_rays,
               S, Fd, ctypes.cast(bgv, ctypes.c_void_p), len(bg), float(background_opacity), _ptr(out["features"]),
               _ptr(out["depths"]), _ptr(out["masks"]), _ptr(out["weights"]), _ptr(out["normals"]), _stream())
    return out


def ray_refine(lengths, weights, n_fine: int, add_input_samples: bool = True, u=None):
    """RayPointRefiner: (n_rays,S) depths + weights -> sorted (n_rays, S + n_fine) depths (holo_ray_refine)."""
    n_rays, S = lengths.shape
    S2 = S + n_fine if add_input_samples else n_fine
    out = torch.empty(n_rays, S2, device=lengths.device)
    lib().call("holo_ray_refine", _ptr(lengths), _ptr(weights), _ptr(u), n_rays, S, n_fine,
               1 if add_input_samples else 0, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------ denoiser
def transpose2d(src, rows: int, cols: int, out=None):
    if out is None:
        out = torch.empty(cols * rows, device=src.device)
    lib().call("holo_transpose2d", _ptr(src), _ptr(out), rows, cols, _stream())
    return out


def gn_stats(x1, C1, x2, C2, V, acc):
    lib().call("holo_gn_stats", _ptr(x1), C1, _ptr(x2), C2, V, _ptr(acc, torch.float64), _stream())


def gn_stats_pp(x1, C1, x2, C2, V, acc, acc_next):
    lib().call("holo_gn_stats_pp", _ptr(x1), C1, _ptr(x2), C2, V, _ptr(acc, torch.float64),
               _ptr(acc_next, torch.float64), _stream())


def gn_apply_fused(x1, C1, x2, C2, V, acc, gamma, beta, film, eps, silu: bool, y=None, y_hi=None, y_lo=None,
                   raw_hi=None, raw_lo=None):
    lib().call("holo_gn_apply_fused", _ptr(x1), C1, _ptr(x2), C2, V, _ptr(acc, torch.float64), _ptr(gamma), _ptr(beta),
               _ptr(film), float(eps), 1 if silu else 0, _ptr(y), _ptr16(y_hi), _ptr16(y_lo),
               _ptr16(raw_hi), _ptr16(raw_lo), _pair_f16(y_hi, y_lo, raw_hi, raw_lo), _stream())


def gn_apply_fused_ch(x1, C1, st1, x2, C2, st2, V, gamma, beta, film, eps, silu: bool, y=None, y_hi=None, y_lo=None,
                      raw_hi=None, raw_lo=None):
    lib().call("holo_gn_apply_fused_ch", _ptr(x1), C1, _ptr(st1, torch.float64), _ptr(x2), C2, _ptr(st2, torch.float64), V,
               _ptr(gamma), _ptr(beta), _ptr(film), float(eps), 1 if silu else 0, _ptr(y), _ptr16(y_hi),
               _ptr16(y_lo), _ptr16(raw_hi), _ptr16(raw_lo), _pair_f16(y_hi, y_lo, raw_hi, raw_lo), _stream())


def gn_finalize(acc, gamma, beta, film, C, V, a, b, eps=1e-5):
    lib().call("holo_gn_finalize", _ptr(acc, torch.float64), _ptr(gamma), _ptr(beta), _ptr(film), C, V, float(eps),
               _ptr(a), _ptr(b), _stream())


def gn_apply(x1, C1, x2, C2, V, a, b, silu: bool, y=None, y_hi=None, y_lo=None):
    lib().call("holo_gn_apply", _ptr(x1), C1, _ptr(x2), C2, V, _ptr(a), _ptr(b), 1 if silu else 0, _ptr(y),
               _ptr(y_hi, torch.bfloat16), _ptr(y_lo, torch.bfloat16), _stream())


def split_bf16(x, V, C, Cpad, hi, lo, ups_dims=None, x2=None, C2=0):
    """cat(x (V,C), x2 (V,C2)) fp32 -> hi / lo of shape (Vout,Cpad), both bf16 or both fp16 (by dtype);
    ups_dims=(D,H,W) folds a nearest x2 upsample."""
    d = ups_dims or (0, 0, 0)
    lib().call("holo_split_bf16", _ptr(x), C, _ptr(x2), C2, V, Cpad, 1 if ups_dims else 0, d[0], d[1], d[2],
               _ptr16(hi), _ptr16(lo), _pair_f16(hi, lo), _stream())


def _ptr_off(t, off_elems: int, dtype):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr() + off_elems * t.element_size())


def gemm_tc(a_hi, a_lo, a_off, a_pitch, M, K, b_hi, b_lo, b_off, b_pitch, N, bias, residual, out_pitch, out, out_off=0,
            out_hi=None, out_lo=None, out_is_zeroed: bool = False, acc_scale: float = 1.0) -> int:
    """out[m][n] (+out_off, row pitch out_pitch) = bias + residual + sum_k a[m][k] b[n][k] on tcgen05 (bf16x3)."""
    f16 = _pair_f16(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo)
    bf = torch.float16 if f16 else torch.bfloat16
    rc = lib().try_call("holo_gemm_tc", _ptr_off(a_hi, a_off, bf), _ptr_off(a_lo, a_off, bf), a_pitch, M, K,
                        _ptr_off(b_hi, b_off, bf), _ptr_off(b_lo, b_off, bf), b_pitch, N, _ptr(bias),
                        _ptr_off(residual, out_off, torch.float32), out_pitch, _ptr_off(out, out_off, torch.float32),
                        _ptr_off(out_hi, out_off, bf), _ptr_off(out_lo, out_off, bf), 1 if out_is_zeroed else 0,
                        FMT_F16 * f16, float(acc_scale), _stream())
    if rc not in (0, -3):
        raise HoloError(f"holo_gemm_tc failed ({rc}): {lib().cdll.holo_last_error().decode()}")
    return rc


def gemm_tc_act(a_hi, a_lo, a_pitch, M, K, b_hi, b_lo, b_pitch, N, bias, row_term, row_term_rows: int, act: str, out_pitch,
                out=None, out_hi=None, out_lo=None, acc_scale: float = 1.0) -> int:
    """out[m][n] = act(acc_scale * a b^T + bias[n] + row_term[m % row_term_rows][n]) as fp32 and / or an operand pair."""
    f16 = _pair_f16(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo)
    rc = lib().try_call("holo_gemm_tc_act", _ptr16(a_hi), _ptr16(a_lo), a_pitch, M, K, _ptr16(b_hi), _ptr16(b_lo), b_pitch, N,
                        _ptr(bias), _ptr(row_term), int(row_term_rows), VP_ACT[act], out_pitch, _ptr(out), _ptr16(out_hi),
                        _ptr16(out_lo), FMT_F16 * f16, float(acc_scale), _stream())
    if rc not in (0, -3):
        raise HoloError(f"holo_gemm_tc_act failed ({rc}): {lib().cdll.holo_last_error().decode()}")
    return rc


P_SCALE_F16 = 4096.0   # fp16 probability pairs are written as 4096 P; the P V GEMM takes acc_scale = 1 / 4096


def softmax_split(S, n_rows, T, scale2, P_hi, P_lo) -> float:
    """P = softmax(scale2 * S) as an operand pair; returns the power of two P was multiplied by (1 for bf16 halves,
    4096 for fp16 halves: pass its reciprocal to the P V GEMM as acc_scale)."""
    f16 = _pair_f16(P_hi, P_lo)
    p_scale = P_SCALE_F16 if f16 else 1.0
    lib().call("holo_softmax_split", _ptr(S), n_rows, T, float(scale2), _ptr16(P_hi), _ptr16(P_lo), f16, p_scale,
               _stream())
    return p_scale


def transpose_split(src, src_off, src_pitch, rows, cols, hi, lo):
    lib().call("holo_transpose_split_bf16", _ptr_off(src, src_off, torch.float32), src_pitch, rows, cols,
               _ptr16(hi), _ptr16(lo), _pair_f16(hi, lo), _stream())


def conv3d_simt(x1, C1, x2, C2, dims: Tuple[int, int, int], ksize, stride, ups, w, bias, residual, Cout, out):
    lib().call("holo_conv3d_simt", _ptr(x1), C1, _ptr(x2), C2, dims[0], dims[1], dims[2], ksize, stride,
               1 if ups else 0, _ptr(w), _ptr(bias), _ptr(residual), Cout, _ptr(out), _stream())


FMT_F16 = 1   # include/holo_b200.h HOLO_FMT_F16


def conv3d_tc(x_hi, x_lo, Cin, dims, ksize, w_hi, w_lo, bias, residual, Cout, out, out_hi=None, out_lo=None,
              stride: int = 1, stats=None, w_scale: float = 1.0, splitk_ws=None) -> int:
    """dims = INPUT volume.  Returns the library status (0 ok, 1 ok but `stats` not produced, -3 unsupported shape);
    other errors raise.  stats: optional zeroed fp64 (Cout, 2) tensor receiving per-channel (sum, sumsq) of the output.
    splitk_ws: fp32 scratch of splitk_ws_floats() elements: split-K launches then reduce deterministically through it (a
    second launch) and produce `stats` too.
    The operand format follows the dtype of the four halves (all bf16 or all fp16); the weight pair holds
    w_scale * w (a power of two; the kernel multiplies the accumulators by 1 / w_scale)."""
    fmt = FMT_F16 * _pair_f16(x_hi, x_lo, w_hi, w_lo, out_hi, out_lo)
    rc = lib().try_call("holo_conv3d_tc", _ptr16(x_hi), _ptr16(x_lo), Cin, dims[0],
                        dims[1], dims[2], ksize, stride, _ptr16(w_hi), _ptr16(w_lo), _ptr(bias),
                        _ptr(residual), Cout, _ptr(out), _ptr16(out_hi), _ptr16(out_lo),
                        _ptr(stats, torch.float64), fmt, 1.0 / float(w_scale), _ptr(splitk_ws), _stream())
    if rc not in (0, 1, -3):
        raise HoloError(f"holo_conv3d_tc failed ({rc}): {lib().cdll.holo_last_error().decode()}")
    return rc


def splitk_ws_floats() -> int:
    return int(lib().cdll.holo_conv3d_tc_splitk_bytes()) // 4


def conv3d_tc_skip(x_hi, x_lo, Cin, skip_hi, skip_lo, Cin_skip, dims, w_hi, w_lo, bias, residual, Cout, out, stats=None,
                   w_scale: float = 1.0, splitk_ws=None) -> int:
    """out = conv3^3(x) + conv1^1(skip) + bias (+ residual) in one launch (holo_conv3d_tc_skip); w = [Cout][27 Cin +
    Cin_skip] pairs.  Status as conv3d_tc."""
    fmt = FMT_F16 * _pair_f16(x_hi, x_lo, skip_hi, skip_lo, w_hi, w_lo)
    rc = lib().try_call("holo_conv3d_tc_skip", _ptr16(x_hi), _ptr16(x_lo), Cin, _ptr16(skip_hi), _ptr16(skip_lo), Cin_skip,
                        dims[0], dims[1], dims[2], _ptr16(w_hi), _ptr16(w_lo), _ptr(bias), _ptr(residual), Cout, _ptr(out),
                        _ptr(stats, torch.float64), fmt, 1.0 / float(w_scale), _ptr(splitk_ws), _stream())
    if rc not in (0, 1, -3):
        raise HoloError(f"holo_conv3d_tc_skip failed ({rc}): {lib().cdll.holo_last_error().decode()}")
    return rc


def v_transpose_split(qkv, T, heads, ch, vt_hi, vt_lo):
    lib().call("holo_v_transpose_split", _ptr(qkv), T, heads, ch, _ptr16(vt_hi), _ptr16(vt_lo), _pair_f16(vt_hi, vt_lo),
               _stream())


def attention_flash_workspace(T, heads, ch, kv_splits: int, device) -> Optional[torch.Tensor]:
    n = lib().cdll.holo_attention_flash_workspace_bytes(T, heads, ch, kv_splits)
    return torch.empty(n, dtype=torch.uint8, device=device) if n > 0 else None


def attention_flash(qkv_hi, qkv_lo, vt_hi, vt_lo, T, heads, ch, out=None, out_hi=None, out_lo=None,
                    softmax_scale: float = 0.0, q_begin: int = 0, q_count: int = 0, kv_splits: int = 1,
                    workspace=None) -> int:
    """Fused attention on tcgen05 (holo_attention_flash).  Returns 0, or -3 for a shape the kernel does not take.
    softmax_scale 0 = ch^-1/2; (q_begin, q_count) restricts the launch to a query range (0, 0 = all); kv_splits > 1
    shares the keys of a query tile between that many CTAs (workspace from attention_flash_workspace)."""
    rc = lib().try_call("holo_attention_flash", _ptr16(qkv_hi), _ptr16(qkv_lo), _ptr16(vt_hi), _ptr16(vt_lo), T,
                        heads, ch, _ptr(out), _ptr16(out_hi), _ptr16(out_lo),
                        _pair_f16(qkv_hi, qkv_lo, vt_hi, vt_lo, out_hi, out_lo), float(softmax_scale), int(q_begin),
                        int(q_count), int(kv_splits), _ptr(workspace, torch.uint8), _stream())
    if rc not in (0, -3):
        raise HoloError(f"holo_attention_flash failed ({rc}): {lib().cdll.holo_last_error().decode()}")
    return rc


def attention_simt(qkv, T, heads, ch, out):
    lib().call("holo_attention_simt", _ptr(qkv), T, heads, ch, _ptr(out), _stream())


_FREQS = {}


def timestep_freqs(dim: int, device) -> torch.Tensor:
    """th.exp(-math.log(10000) * th.arange(half) / half) on the CPU, then moved -- nn.py:119-121."""
    import math
    key = (dim, str(device))
    if key not in _FREQS:
        half = dim // 2
        _FREQS[key] = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(device)
    return _FREQS[key]


def timestep_embedding(t, dim, out):
    lib().call("holo_timestep_embedding", _ptr(t, torch.int64), t.numel(), dim, _ptr(timestep_freqs(dim, t.device)),
               _ptr(out), _stream())


def linear_rows(x, W, b, M, in_dim, out_dim, silu_in, silu_out, out):
    lib().call("holo_linear_rows", _ptr(x), _ptr(W), _ptr(b), M, in_dim, out_dim, 1 if silu_in else 0,
               1 if silu_out else 0, _ptr(out), _stream())


def ddpm_step(model_out, x_t, noise, t, coef1, coef2, logvar, clip: bool, x_prev, pred_x0=None):
    n_batch = t.numel()
    per = model_out.numel() // n_batch
    lib().call("holo_ddpm_step", _ptr(model_out), _ptr(x_t), _ptr(noise), _ptr(t, torch.int64), _ptr(coef1),
               _ptr(coef2), _ptr(logvar), per, n_batch, 1 if clip else 0, _ptr(x_prev), _ptr(pred_x0), _stream())


def ddim_step(model_out, x_t, noise, t, ac, ac_to, sqrt_recip, sqrt_recipm1, eta: float, clip: bool, x_out, pred_x0=None):
    n_batch = t.numel()
    lib().call("holo_ddim_step", _ptr(model_out), _ptr(x_t), _ptr(noise), _ptr(t, torch.int64), _ptr(ac), _ptr(ac_to),
               _ptr(sqrt_recip), _ptr(sqrt_recipm1), float(eta), model_out.numel() // n_batch, n_batch, 1 if clip else 0,
               _ptr(x_out), _ptr(pred_x0), _stream())


def q_sample(x0, noise, t, sqrt_ac, sqrt_1m_ac, out):
    n_batch = t.numel()
    lib().call("holo_q_sample", _ptr(x0), _ptr(noise), _ptr(t, torch.int64), _ptr(sqrt_ac), _ptr(sqrt_1m_ac),
               x0.numel() // n_batch, n_batch, _ptr(out), _stream())


def range_init(stats):
    lib().call("holo_range_init", _ptr(stats, torch.int32), _stream())


def act_range(x_cl, V, C, act: int, y_cl, y_cf, stats):
    lib().call("holo_act_range", _ptr(x_cl), V, C, act, _ptr(y_cl), _ptr(y_cf), _ptr(stats, torch.int32), _stream())


def decode_range(stats_host) -> Tuple[float, float, int]:
    """stats4 (int32, host) -> (min, max, nan_count)."""
    import struct

    def dec(i):
        i = int(i)
        if i < 0:
            i ^= 0x7FFFFFFF
        return struct.unpack("f", struct.pack("i", i))[0]

    return dec(stats_host[0]), dec(stats_host[1]), int(stats_host[2])


# ------------------------------------------------------------------ per-view post-processing (generate_samples.py)
def depth_image(depth, mask, min_quantile=0.02, max_quantile=0.98, min_out=0.1, max_out=0.9, composite_white=True):
    """(H, W) depth + mask -> ((3, H, W) depth visualisation, (2,) selected (min, max) depths)."""
    H, W = depth.shape[-2:]
    out = torch.empty(3, H, W, device=depth.device)
    nf = torch.empty(2, device=depth.device)
    lib().call("holo_depth_image", _ptr(depth), _ptr(mask), H * W, float(min_quantile), float(max_quantile),
               float(min_out), float(max_out), 1 if composite_white else 0, _ptr(nf), _ptr(out), _stream())
    return out, nf


def frame_u8(src_chw, out_hw=None, out=None):
    """(C, H, W) float image (C = 1 | 3) -> (h, w, 3) uint8 frame (clip, bilinear resize, round)."""
    C, H, W = src_chw.shape
    h, w = (H, W) if out_hw is None else out_hw
    if out is None:
        out = torch.empty(h, w, 3, dtype=torch.uint8, device=src_chw.device)
    lib().call("holo_frame_u8", _ptr(src_chw), C, H, W, h, w, _ptr(out, torch.uint8), _stream())
    return out


MATERIALS = {   # shaded_depth_render.py:83-98: (ambient, diffuse, specular, shininess)
    "high_contrast": ((0.5, 0.5, 0.5), (2.0, 2.0, 2.0), (1.0, 1.0, 0.9), 256.0),
    "medium": ((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 0.9), 128.0),
}


def shade_depth(depth, mask, focal, pp, smooth_k: int, mask_thr=0.5, depth_thr=1e-2, material="medium",
                bg=(1.0, 1.0, 1.0), light=(0.5, 0.3, 0.2)):
    """(H, W) depth + mask of one view, NDC focal (fx, fy) / principal point (px, py) as floats ->
    ((3, H, W) shaded render, (H, W) mask).  `light` = ambient / diffuse / specular strength of the point light at the
    camera centre (pytorch3d PointLights defaults: mesh_render.py:66-68 builds it with none given); the material colours
    multiply them (Gouraud: colour = ambient + diffuse * n.l + specular * (r.v)^shininess at every vertex)."""
    H, W = depth.shape[-2:]
    amb, dif, spec, shin = MATERIALS[material]
    m10 = [a * light[0] for a in amb] + [d * light[1] for d in dif] + [s * light[2] for s in spec] + [shin]
    dev = depth.device
    out, om = torch.empty(3, H, W, device=dev), torch.empty(H, W, device=dev)
    sd, ok = torch.empty(H, W, device=dev), torch.empty(H, W, dtype=torch.uint8, device=dev)
    lib().call("holo_shade_depth", _ptr(depth), _ptr(mask), H, W, float(focal[0]), float(focal[1]), float(pp[0]),
               float(pp[1]), int(smooth_k), float(mask_thr), float(depth_thr),
               ctypes.cast((ctypes.c_float * 10)(*m10), ctypes.c_void_p), ctypes.cast(_host3(bg), ctypes.c_void_p),
               _ptr(sd), _ptr(ok, torch.uint8), _ptr(out), _ptr(om), _stream())
    return out, om


# ------------------------------------------------------------------ view-pooling encoder (csrc/viewpool.cu)
class HoloFeatureMap(ctypes.Structure):   # include/holo_b200.h: holo_feature_map
    _fields_ = [("data", ctypes.c_void_p), ("channels", _i), ("height", _i), ("width", _i)]


VP_ACT = {"identity": 0, "relu": 1, "leakyrelu": 2, "softplus": 3}


def viewpool_sample(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, n_harmonic: int, Kpad: int, rows_per_view: int,
                    x_hi, x_lo, mean_hi, mean_lo, mask_map=None, view_weight=None, eps: float = 1e-2, x_f32=None,
                    mean_f32=None):
    """pts (P, 3); maps_cl: list of channels-last (n_src, H, W, C) feature maps; rows of X / mean as operand pairs."""
    n_src = cam_R.shape[0]
    arr = _feature_map_array(maps_cl, n_src)
    Hm, Wm = (mask_map.shape[-2], mask_map.shape[-1]) if mask_map is not None else (0, 0)
    lib().call("holo_viewpool_sample", _ptr(pts), pts.shape[0], _ptr(cam_R), _ptr(cam_T), _ptr(cam_focal), _ptr(cam_pp),
               n_src, ctypes.cast(arr, ctypes.c_void_p), len(maps_cl), _ptr(mask_map), Hm, Wm, _ptr(view_weight),
               int(n_harmonic), float(eps), int(Kpad), int(rows_per_view), _ptr16(x_hi), _ptr16(x_lo), _ptr16(mean_hi),
               _ptr16(mean_lo), _ptr(x_f32), _ptr(mean_f32), _pair_f16(x_hi, x_lo, mean_hi, mean_lo), _stream())


def _feature_map_array(maps_cl, n_src: int):
    arr = (HoloFeatureMap * len(maps_cl))()
    for k, m in enumerate(maps_cl):
        if m.dim() != 4 or m.shape[0] != n_src:
            raise HoloError(f"feature map {k}: expected (n_src={n_src}, H, W, C), got {tuple(m.shape)}")
        arr[k].data, arr[k].channels, arr[k].height, arr[k].width = _ptr(m, name=f"feature map {k}").value, m.shape[3], \
            m.shape[1], m.shape[2]
    return arr


def viewpool_angle_reduce(pts, cam_R, cam_T, cam_focal, cam_pp, maps_cl, Kpad: int, out_hi, out_lo, mask_map=None,
                          view_weight=None, eps: float = 1e-2, gamma: float = 1.0, min_ray_angle_weight: float = 0.1,
                          with_std: bool = True, out_f32=None):
    n_src = cam_R.shape[0]
    arr = _feature_map_array(maps_cl, n_src)
    Hm, Wm = (mask_map.shape[-2], mask_map.shape[-1]) if mask_map is not None else (0, 0)
    lib().call("holo_viewpool_angle_reduce", _ptr(pts), pts.shape[0], _ptr(cam_R), _ptr(cam_T), _ptr(cam_focal),
               _ptr(cam_pp), n_src, ctypes.cast(arr, ctypes.c_void_p), len(maps_cl), _ptr(mask_map), Hm, Wm,
               _ptr(view_weight), float(eps), float(gamma), float(min_ray_angle_weight), 1 if with_std else 0, int(Kpad),
               _ptr16(out_hi), _ptr16(out_lo), _ptr(out_f32), _pair_f16(out_hi, out_lo), _stream())


def viewpool_act_split(y, point_term, n_views: int, rows_per_view: int, C: int, act: str, hi, lo):
    lib().call("holo_viewpool_act_split", _ptr(y), _ptr(point_term), int(n_views), int(rows_per_view), int(C), VP_ACT[act],
               _ptr16(hi), _ptr16(lo), _pair_f16(hi, lo), _stream())


def viewpool_reduce(z, n_views: int, rows_per_view: int, n_pts: int, C: int, out=None, out_hi=None, out_lo=None):
    lib().call("holo_viewpool_reduce", _ptr(z), int(n_views), int(rows_per_view), int(n_pts), int(C), _ptr(out),
               _ptr16(out_hi), _ptr16(out_lo), _pair_f16(out_hi, out_lo), _stream())


# ------------------------------------------------------------------ whole-graph denoiser (csrc/unet_exec.cu)
class HoloUnetConfig(ctypes.Structure):   # include/holo_b200.h: holo_unet_config (same field order)
    _fields_ = [("in_channels", _i), ("model_channels", _i), ("out_channels", _i), ("num_res_blocks", _i), ("n_levels", _i),
                ("channel_mult", _i * 8), ("n_attention_resolutions", _i), ("attention_resolutions", _i * 8), ("num_heads", _i),
                ("D", _i), ("H", _i), ("W", _i), ("pair_f16", _i), ("fuse_skip", _i), ("attn_kv_split", _i),
                ("use_tensor_cores", _i), ("splitk_workspace", _i)]


class NativeUnet:
    """holo_unet_* behind a small object: the C++ executor of the whole UNet forward.  `state_dict` maps the reference's
    parameter names below `_net.` to fp32 CUDA tensors, which are BORROWED (kept alive here)."""

    def __init__(self, state_dict, in_channels, model_channels, out_channels, num_res_blocks, channel_mult,
                 attention_resolutions, num_heads, dims, pair_f16=True, fuse_skip=True, attn_kv_split=0,
                 use_tensor_cores=True, splitk_workspace=True):
        cm, ar = list(channel_mult), list(attention_resolutions)
        cfg = HoloUnetConfig(in_channels, model_channels, out_channels, num_res_blocks, len(cm), (_i * 8)(*cm), len(ar),
                             (_i * 8)(*ar), num_heads, dims[0], dims[1], dims[2], 1 if pair_f16 else 0,
                             1 if fuse_skip else 0, int(attn_kv_split), 1 if use_tensor_cores else 0,
                             1 if splitk_workspace else 0)
        h = ctypes.c_void_p()
        lib().call("holo_unet_create", ctypes.byref(cfg), ctypes.byref(h))
        self._h, self.cfg = h, cfg
        L = lib().cdll
        self.names = [L.holo_unet_param_name(h, i).decode() for i in range(L.holo_unet_param_count(h))]
        self._keep = {}
        self.set_params(state_dict)

    def set_params(self, state_dict, stream=None):
        dev = None
        for i, name in enumerate(self.names):
            t = state_dict[name].detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            dev = t.device
            self._keep[name] = t
            lib().call("holo_unet_set_param", self._h, name.encode(), _ptr(t), t.numel())
        self.packed = torch.empty(int(lib().cdll.holo_unet_packed_bytes(self._h)), dtype=torch.uint8, device=dev)
        self.workspace = torch.empty(int(lib().cdll.holo_unet_workspace_bytes(self._h)), dtype=torch.uint8, device=dev)
        lib().call("holo_unet_pack", self._h, _ptr(self.packed, torch.uint8), _stream())

    def forward_cl(self, x_cl, t, out=None):
        V = self.cfg.D * self.cfg.H * self.cfg.W
        if out is None:
            out = torch.empty(V, self.cfg.out_channels, device=x_cl.device)
        lib().call("holo_unet_fwd_cl", self._h, _ptr(x_cl), _ptr(t, torch.int64), _ptr(out),
                   _ptr(self.workspace, torch.uint8), _stream())
        return out

    def forward(self, x, t, out=None):
        c = self.cfg
        if out is None:
            out = torch.empty(1, c.out_channels, c.D, c.H, c.W, device=x.device)
        lib().call("holo_unet_fwd", self._h, _ptr(x), _ptr(t, torch.int64), _ptr(out), _ptr(self.workspace, torch.uint8),
                   _stream())
        return out

    def __del__(self):
        try:
            lib().cdll.holo_unet_destroy(self._h)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass
