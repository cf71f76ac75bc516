"""CPU stand-ins for the C-ABI entry points the UNet executor calls (TEST INFRASTRUCTURE ONLY): plain torch ops with the
same argument meaning, so that ``UNetExecutor``'s orchestration -- which tensor feeds which kernel, the un-materialised
skip concat, FiLM slices, epilogue statistics, residuals, zero-padded attention heads, the query-sharded attention
under gloo -- runs without a GPU and is compared with the oracle.  Operand pairs are kept in fp32 (hi = value, lo = 0):
set ``executor.pair_dtype = torch.float32``.  The kernels themselves are checked in tests/test_unet_gpu.py."""
import math

import torch
import torch.nn.functional as F


def _cl_to_ncdhw(x_cl, C, dims):
    D, H, W = dims
    return x_cl.reshape(D, H, W, C).permute(3, 0, 1, 2)[None]


def _ncdhw_to_cl(x):
    C = x.shape[1]
    return x[0].permute(1, 2, 3, 0).reshape(-1, C)


def _cat(x1, C1, x2, C2):
    return x1 if x2 is None or C2 == 0 else torch.cat([x1, x2], 1)


def timestep_embedding(t, dim, out):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    out.copy_(torch.cat([torch.cos(args), torch.sin(args)], -1))


def linear_rows(x, W, b, M, in_dim, out_dim, silu_in, silu_out, out):
    y = F.linear(F.silu(x) if silu_in else x, W, b)
    out.copy_((F.silu(y) if silu_out else y).reshape(out.shape))


def gn_stats_pp(x1, C1, x2, C2, V, acc, acc_next):
    x = _cat(x1, C1, x2, C2).double()
    C = x.shape[1]
    g = x.reshape(V, 32, C // 32)
    acc.view(8, 32, 2)[0, :, 0] += g.sum((0, 2))
    acc.view(8, 32, 2)[0, :, 1] += (g * g).sum((0, 2))
    acc_next.zero_()


def _apply(x, group_sum, group_sq, V, gamma, beta, film, eps, silu):
    C = x.shape[1]
    cpg = C // 32
    count = float(V) * cpg
    mean = group_sum / count
    var = (group_sq / count - mean * mean).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + eps)).float().repeat_interleave(cpg)
    mean = mean.float().repeat_interleave(cpg)
    a = rstd * gamma
    b = beta - mean * a
    if film is not None:
        sc, sh = 1.0 + film[:C], film[C:]
        a, b = a * sc, b * sc + sh
    y = x * a + b
    return F.silu(y) if silu else y


def _emit(val, y, y_hi, y_lo):
    if y is not None:
        y.copy_(val)
    if y_hi is not None:
        y_hi.copy_(val)
        y_lo.zero_()


def gn_apply_fused(x1, C1, x2, C2, V, acc, gamma, beta, film, eps, silu, y=None, y_hi=None, y_lo=None, raw_hi=None,
                   raw_lo=None):
    x = _cat(x1, C1, x2, C2)
    s = acc.view(8, 32, 2).sum(0)
    _emit(_apply(x, s[:, 0], s[:, 1], V, gamma, beta, film, eps, silu), y, y_hi, y_lo)
    if raw_hi is not None:
        _emit(x, None, raw_hi, raw_lo)


def gn_apply_fused_ch(x1, C1, st1, x2, C2, st2, V, gamma, beta, film, eps, silu, y=None, y_hi=None, y_lo=None,
                      raw_hi=None, raw_lo=None):
    x = _cat(x1, C1, x2, C2)
    st = st1.view(C1, 2) if x2 is None or C2 == 0 else torch.cat([st1.view(C1, 2), st2.view(C2, 2)], 0)
    C = x.shape[1]
    g = st.view(32, C // 32, 2).sum(1)
    _emit(_apply(x, g[:, 0], g[:, 1], V, gamma, beta, film, eps, silu), y, y_hi, y_lo)
    if raw_hi is not None:
        _emit(x, None, raw_hi, raw_lo)


def split_bf16(x, V, C, Cpad, hi, lo, ups_dims=None, x2=None, C2=0):
    v = _cat(x, C, x2, C2)
    if ups_dims is not None:
        v = _ncdhw_to_cl(F.interpolate(_cl_to_ncdhw(v, v.shape[1], ups_dims), scale_factor=2, mode="nearest"))
    hi.zero_()
    hi[:, : v.shape[1]] = v
    lo.zero_()


def conv3d_tc(x_hi, x_lo, Cin, dims, ksize, w_hi, w_lo, bias, residual, Cout, out, out_hi=None, out_lo=None, stride=1,
              stats=None, w_scale=1.0, splitk_ws=None):
    x = _cl_to_ncdhw(x_hi + x_lo, Cin, dims)
    w = ((w_hi + w_lo) / w_scale).reshape(Cout, ksize, ksize, ksize, Cin).permute(0, 4, 1, 2, 3)
    y = _ncdhw_to_cl(F.conv3d(x, w, bias, stride=stride, padding=ksize // 2))
    if residual is not None:
        y = y + residual
    if out is not None:
        out.copy_(y)
    if out_hi is not None:
        out_hi.copy_(y)
        out_lo.zero_()
    if stats is not None:
        st = stats.view(Cout, 2)
        st[:, 0] += y.double().sum(0)
        st[:, 1] += (y.double() ** 2).sum(0)
    return 0


def conv3d_tc_skip(x_hi, x_lo, Cin, skip_hi, skip_lo, Cin_skip, dims, w_hi, w_lo, bias, residual, Cout, out, stats=None,
                   w_scale=1.0, splitk_ws=None):
    w = (w_hi + w_lo) / w_scale                               # [Cout][27 Cin + Cin_skip]
    w3 = w[:, : 27 * Cin].reshape(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3)
    w1 = w[:, 27 * Cin:].reshape(Cout, Cin_skip, 1, 1, 1)
    y = F.conv3d(_cl_to_ncdhw(x_hi + x_lo, Cin, dims), w3, bias, padding=1) + \
        F.conv3d(_cl_to_ncdhw(skip_hi + skip_lo, Cin_skip, dims), w1)
    y = _ncdhw_to_cl(y)
    if residual is not None:
        y = y + residual
    out.copy_(y)
    if stats is not None:
        st = stats.view(Cout, 2)
        st[:, 0] += y.double().sum(0)
        st[:, 1] += (y.double() ** 2).sum(0)
    return 0


def conv3d_simt(x1, C1, x2, C2, dims, ksize, stride, ups, w, bias, residual, Cout, out):
    x = _cl_to_ncdhw(_cat(x1, C1, x2, C2), C1 + C2, dims)
    if ups:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    wt = w.reshape(ksize, ksize, ksize, C1 + C2, Cout).permute(4, 3, 0, 1, 2)   # [tap][Cin][Cout]
    y = _ncdhw_to_cl(F.conv3d(x, wt, bias, stride=stride, padding=ksize // 2))
    out.copy_(y if residual is None else y + residual)


def _legacy_attention(q, k, v, scale2):   # (heads, T, ch) each
    w = torch.softmax(torch.einsum("htc,hsc->hts", q, k) * scale2, -1)
    return torch.einsum("hts,hsc->htc", w, v)


def attention_simt(qkv, T, heads, ch, out):
    q, k, v = qkv.reshape(T, heads, 3, ch).permute(2, 1, 0, 3)
    out.copy_(_legacy_attention(q, k, v, 1.0 / math.sqrt(ch)).permute(1, 0, 2).reshape(T, heads * ch))


def v_transpose_split(qkv, T, heads, ch, vt_hi, vt_lo):
    vt_hi.copy_(qkv.reshape(T, heads, 3, ch)[:, :, 2].reshape(T, heads * ch).t())
    vt_lo.zero_()


def attention_flash_workspace(T, heads, ch, kv_splits, device):
    return None


def attention_flash(qkv_hi, qkv_lo, vt_hi, vt_lo, T, heads, ch, out=None, out_hi=None, out_lo=None, softmax_scale=0.0,
                    q_begin=0, q_count=0, kv_splits=1, workspace=None):
    if q_count <= 0:
        q_begin, q_count = 0, T
    qkv = (qkv_hi + qkv_lo).reshape(T, heads, 3, ch)
    q, k = qkv[q_begin:q_begin + q_count, :, 0].permute(1, 0, 2), qkv[:, :, 1].permute(1, 0, 2)
    v = (vt_hi + vt_lo).reshape(heads, ch, T).permute(0, 2, 1)
    o = _legacy_attention(q, k, v, softmax_scale if softmax_scale > 0 else 1.0 / math.sqrt(ch))
    o = o.permute(1, 0, 2).reshape(q_count, heads * ch)
    if out is not None:
        out[q_begin:q_begin + q_count] = o
    if out_hi is not None:
        out_hi[q_begin:q_begin + q_count] = o
        out_lo[q_begin:q_begin + q_count] = 0
    return 0


def _view(t, off, pitch, rows, cols):
    return t.reshape(-1)[off:].as_strided((rows, cols), (pitch, 1))


def gemm_tc(a_hi, a_lo, a_off, a_pitch, M, K, b_hi, b_lo, b_off, b_pitch, N, bias, residual, out_pitch, out, out_off=0,
            out_hi=None, out_lo=None, out_is_zeroed=False, acc_scale=1.0):
    a = _view(a_hi, a_off, a_pitch, M, K) + _view(a_lo, a_off, a_pitch, M, K)
    b = _view(b_hi, b_off, b_pitch, N, K) + _view(b_lo, b_off, b_pitch, N, K)
    y = (a @ b.t()) * acc_scale
    if bias is not None:
        y = y + bias
    if residual is not None:
        y = y + _view(residual, out_off, out_pitch, M, N)
    o = _view(out, out_off, out_pitch, M, N)
    o.copy_(o + y if out_is_zeroed else y)
    return 0


def softmax_split(S, n_rows, T, scale2, P_hi, P_lo):
    P_hi.copy_(torch.softmax(S * scale2, -1))
    P_lo.zero_()
    return 1.0


def transpose_split(src, src_off, src_pitch, rows, cols, hi, lo):
    hi.copy_(_view(src, src_off, src_pitch, rows, cols).t())
    lo.zero_()


ALL = dict(timestep_embedding=timestep_embedding, linear_rows=linear_rows, gn_stats_pp=gn_stats_pp,
           gn_apply_fused=gn_apply_fused, gn_apply_fused_ch=gn_apply_fused_ch, split_bf16=split_bf16, conv3d_tc=conv3d_tc,
           conv3d_tc_skip=conv3d_tc_skip, conv3d_simt=conv3d_simt, attention_simt=attention_simt, v_transpose_split=v_transpose_split,
           attention_flash=attention_flash, attention_flash_workspace=attention_flash_workspace, gemm_tc=gemm_tc, softmax_split=softmax_split, transpose_split=transpose_split)


def install(ops_module, setattr_fn=setattr):
    for name, fn in ALL.items():
        setattr_fn(ops_module, name, fn)
    setattr_fn(ops_module, "NativeUnet", None)   # the C++ whole-graph executor has no stand-in: walk the blocks in Python
