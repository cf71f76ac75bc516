"""View-pooling encoder: source views -> voxel grid.

Mirrors the encoder branch of the reference's ``HoloDiffusionModel.forward``
(/root/reference/holo_diffusion/holo_diffusion_model.py:327-373):

    img_feats = image_feature_extractor(source images, source fg masks)                :341-347
    grid_xyz  = VolumeLocator(1, (R, R, R), voxel_size = extent / R).get_coord_grid()  :350-356
    pooled    = view_pooler(pts=grid_xyz, camera=source cameras, feats=img_feats, masks=mask_crop)   :358-367
    grid      = tanh(pooled_feature_mapper(pooled)).permute(0, 3, 1, 2).reshape(1, -1, R, R, R)      :368-373

with the in-tree ``MLPMeanFeatureAggregator`` (custom_modules.py:162-281; configs/hydrant.yaml:184) or pytorch3d's
``AngleWeightedReductionFeatureAggregator`` (the view pooler's default, which configs/base.yaml keeps) as aggregator.

Everything from the projection of the R^3 grid points to the tanh runs on the library's kernels: ``holo_viewpool_sample``
/ ``holo_viewpool_angle_reduce`` (project, bilinear-sample every feature map, ray-direction embedding, weights, mean),
``holo_gemm_tc`` (every Linear layer, tcgen05, fp32 operands as 16-bit hi/lo pairs), ``holo_viewpool_act_split``,
``holo_viewpool_reduce`` (softmax-weighted sum over the views), ``holo_act_range`` (tanh, both layouts, range statistics)
-- chunked over the points so that the (views x points x 128) intermediates stay L2-sized.  The image feature extractor
in front (torchvision ResNet34 in the reference: pytorch3d ``ResNetFeatureExtractor``) is a library network on both
sides (cuDNN), restated here so that the shipped configs resolve.

The pytorch3d leaves this follows (ViewSampler / project_points_and_sample / ndc_grid_sample / VolumeLocator / wmean /
the angle-weighted aggregator / ResNetFeatureExtractor) are restated from memory of pytorch3d 0.7.4: parity unpinned,
like the renderer's leaves.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .cameras import PerspectiveCameras
from .renderer import _MLPParams

MASK_FEATURE_NAME, IMAGE_FEATURE_NAME = "mask", "image"     # pytorch3d feature_extractor names
PAIR_DTYPE = torch.float16   # 16-bit format of the GEMM operand pairs (holo_conv3d_tc operand_fmt); bf16 also works


def _ceil(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class LazyLinearWithXavierInit(nn.LazyLinear):
    """custom_modules.py:37-41: Xavier weights and a zero bias once the input width is known."""

    def reset_parameters(self) -> None:
        if not self.has_uninitialized_params() and self.in_features != 0:
            nn.init.xavier_uniform_(self.weight.data)
            self.bias.data[:] = 0.0


def _materialize(lin: nn.Module, in_dim: int, device):
    """Give a lazy Linear its shape without running it (the reference's first forward does that implicitly)."""
    if isinstance(lin, nn.modules.lazy.LazyModuleMixin) and lin.has_uninitialized_params():
        lin.to(device)
        with torch.no_grad():
            lin._infer_parameters(lin, (torch.empty(1, in_dim, device=device),))   # shapes, reset_parameters, class swap
    if lin.weight.shape[1] != in_dim:
        raise ValueError(f"Linear expects {lin.weight.shape[1]} inputs, the pooled row has {in_dim}")


def _pack_pair(w: torch.Tensor, n_pad: int, k_pad: int, pair_dtype) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """(N, K) weights -> zero-padded K-major hi/lo pair of s * w (s: power of two keeping the fp16 lo half normal)."""
    wp = torch.zeros(n_pad, k_pad, device=w.device, dtype=torch.float32)
    wp[: w.shape[0], : w.shape[1]] = w.float()
    scale = 1.0
    if pair_dtype == torch.float16:
        amax = float(wp.abs().max())
        if amax > 0 and math.isfinite(amax):
            scale = 2.0 ** (9 - math.floor(math.log2(amax)))
    ws = wp * scale
    hi = ws.to(pair_dtype)
    return hi, (ws - hi.float()).to(pair_dtype), scale


def _pad_bias(b: torch.Tensor, n_pad: int) -> torch.Tensor:
    out = torch.zeros(n_pad, device=b.device, dtype=torch.float32)
    out[: b.shape[0]] = b.float()
    return out


class _Gemm:
    """One Linear layer on holo_gemm_tc: out = a W^T + b."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], k_pad: int, pair_dtype, n_mult: int = 64):
        self.n, self.k = w.shape
        self.n_pad, self.k_pad = _ceil(self.n, n_mult), k_pad
        self.hi, self.lo, self.scale = _pack_pair(w, self.n_pad, k_pad, pair_dtype)
        self.bias = None if b is None else _pad_bias(b, self.n_pad)

    def __call__(self, a_hi, a_lo, rows: int, out, out_hi=None, out_lo=None, act: Optional[str] = None, row_term=None,
                 row_term_rows: int = 0):
        if act is not None or row_term is not None:   # activation / per-point term folded into the epilogue
            rc = ops.gemm_tc_act(a_hi, a_lo, self.k_pad, rows, self.k_pad, self.hi, self.lo, self.k_pad, self.n_pad, self.bias,
                                 row_term, row_term_rows, act or "identity", self.n_pad, out, out_hi, out_lo,
                                 acc_scale=1.0 / self.scale)
        else:
            rc = ops.gemm_tc(a_hi, a_lo, 0, self.k_pad, rows, self.k_pad, self.hi, self.lo, 0, self.k_pad, self.n_pad,
                             self.bias, None, self.n_pad, out, 0, out_hi, out_lo, acc_scale=1.0 / self.scale)
        if rc != 0:
            raise ops.HoloError(f"holo_gemm_tc rejected a {rows} x {self.k_pad} x {self.n_pad} Linear layer: "
                                f"{ops.lib().cdll.holo_last_error().decode()}")


# ------------------------------------------------------------------------------------------------ aggregators
class MLPMeanFeatureAggregator(nn.Module):
    """custom_modules.py:162-281.  Parameter names as in the reference (``_first_sampled``, ``_first_mean``,
    ``_mlp.mlp.{i}.0``, ``_last``) so that its checkpoints load."""

    def __init__(self, exclude_target_view: bool = True, exclude_target_view_mask_features: bool = True,
                 concatenate_output: bool = True, n_hidden: int = 128, dim_out: int = 128, n_layers: int = 1,
                 n_harmonic_functions_ray: int = 3, checkpointed_mlp: bool = True):
        super().__init__()
        if n_layers < 1:
            raise NotImplementedError("MLPMeanFeatureAggregator with an empty MLP")
        self.exclude_target_view, self.exclude_target_view_mask_features = exclude_target_view, exclude_target_view_mask_features
        self.concatenate_output = concatenate_output
        self.n_hidden, self.dim_out, self.n_layers = n_hidden, dim_out, n_layers
        self.n_harmonic_functions_ray, self.checkpointed_mlp = n_harmonic_functions_ray, checkpointed_mlp
        self._first_sampled = LazyLinearWithXavierInit(n_hidden)
        self._first_mean = LazyLinearWithXavierInit(n_hidden)
        self._last = nn.Linear(n_hidden, dim_out)
        nn.init.xavier_uniform_(self._last.weight)
        self._mlp = _MLPParams(n_layers, n_hidden, n_hidden, n_hidden, n_hidden, ())

    def get_aggregated_feature_dim(self, feats_or_feats_dim=None) -> int:
        return self.dim_out

    def forward(self, *a, **k):
        raise NotImplementedError("the aggregator runs fused with the view sampling: call ViewPooler.forward / pool_views")


class AngleWeightedReductionFeatureAggregator(nn.Module):
    """pytorch3d's default aggregator of the view pooler (no parameters): angle-weighted mean and std over the views."""

    def __init__(self, exclude_target_view: bool = True, exclude_target_view_mask_features: bool = True,
                 concatenate_output: bool = True, reduction_functions: Sequence = ("AVG", "STD"),
                 weight_by_ray_angle_gamma: float = 1.0, min_ray_angle_weight: float = 0.1):
        super().__init__()
        self.exclude_target_view, self.exclude_target_view_mask_features = exclude_target_view, exclude_target_view_mask_features
        self.concatenate_output = concatenate_output
        self.reduction_functions = tuple(reduction_functions)
        self.weight_by_ray_angle_gamma, self.min_ray_angle_weight = weight_by_ray_angle_gamma, min_ray_angle_weight

    def get_aggregated_feature_dim(self, feats_or_feats_dim) -> int:
        n = feats_or_feats_dim if isinstance(feats_or_feats_dim, int) else sum(f.shape[1] for f in feats_or_feats_dim.values())
        return n * len(self.reduction_functions)

    def forward(self, *a, **k):
        raise NotImplementedError("the aggregator runs fused with the view sampling: call ViewPooler.forward / pool_views")


def _reductions(agg) -> Tuple[str, ...]:
    return tuple(str(getattr(r, "name", r)).upper().split(".")[-1] for r in agg.reduction_functions)


class ViewSampler(nn.Module):
    def __init__(self, masked_sampling: bool = False, sampling_mode: str = "bilinear"):
        super().__init__()
        self.masked_sampling, self.sampling_mode = masked_sampling, sampling_mode


AGGREGATORS = {"MLPMeanFeatureAggregator": MLPMeanFeatureAggregator,
               "AngleWeightedReductionFeatureAggregator": AngleWeightedReductionFeatureAggregator}


# ------------------------------------------------------------------------------------------------ the fused pooling
class _PoolPlan:
    """Packed weights of one (aggregator, mapper) pair; rebuilt when a parameter changes."""

    def __init__(self):
        self.key = None


def _param_key(mods: Sequence[Optional[nn.Module]], extra) -> tuple:
    ps = [p for m in mods if m is not None for p in m.parameters()]
    lazy = nn.parameter.UninitializedParameter
    return (tuple((-1, id(p)) if isinstance(p, lazy) else (p._version, p.data_ptr()) for p in ps), extra)


def _scatter_cols(w: torch.Tensor, cols: torch.Tensor, k: int) -> torch.Tensor:
    """(N, len(cols)) weights -> (N, k) with the columns moved to `cols` and zeros elsewhere."""
    out = torch.zeros(w.shape[0], k, device=w.device, dtype=w.dtype)
    out[:, cols.to(w.device)] = w
    return out


def _build_mlp_mean_plan(agg, mapper: Optional[nn.Linear], Kx: int, cols: torch.Tensor, Kx_pad: int, Kpad: int, pair_dtype,
                         device) -> dict:
    _materialize(agg._first_sampled, Kx, device)
    _materialize(agg._first_mean, Kx, device)
    layers = [seq[0] for seq in agg._mlp.mlp]
    d = torch.float64
    w0, b0 = layers[0].weight.detach().to(d), layers[0].bias.detach().to(d)
    # mlp_in = first_sampled(x) + first_mean(mean) feeds the MLP's first Linear directly (custom_modules.py:263-264 with
    # MLPWithInputSkips.forward :133-160): the three Linear layers fold into two matrices and one bias (fp64 products)
    A = _scatter_cols(w0 @ agg._first_sampled.weight.detach().to(d), cols, Kx_pad)
    Bm = _scatter_cols(w0 @ agg._first_mean.weight.detach().to(d), cols, Kx_pad)
    bias0 = w0 @ (agg._first_sampled.bias.detach().to(d) + agg._first_mean.bias.detach().to(d)) + b0
    H = _ceil(A.shape[0], 64)
    plan = {"first": _Gemm(A, None, Kpad, pair_dtype), "mean": _Gemm(Bm, bias0, Kpad, pair_dtype), "hidden": [], "acts": []}
    n = len(layers)
    # activation placement of MLPWithInputSkips (custom_modules.py:108-112): the last layer gets the hidden activation
    # (LeakyReLU 0.2), the others the "last" one (Softplus)
    plan["acts"] = ["leakyrelu" if i == n - 1 else "softplus" for i in range(n)]
    for lin in layers[1:]:
        plan["hidden"].append(_Gemm(lin.weight.detach(), lin.bias.detach(), H, pair_dtype))
        H = _ceil(lin.weight.shape[0], 64)
    plan["last"] = _Gemm(agg._last.weight.detach(), agg._last.bias.detach(), H, pair_dtype)
    plan["D"] = agg._last.weight.shape[0]
    if mapper is not None:
        _materialize(mapper, plan["D"], device)
    plan["mapper"] = None if mapper is None else _Gemm(mapper.weight.detach(), mapper.bias.detach(), plan["last"].n_pad,
                                                       pair_dtype, n_mult=16)
    return plan


class ViewPooler(nn.Module):
    """pytorch3d ``ViewPooler`` (view_sampler + feature_aggregator) with the fused CUDA path behind it."""

    def __init__(self, view_sampler_args: Optional[dict] = None,
                 feature_aggregator_class_type: str = "AngleWeightedReductionFeatureAggregator", **aggregator_args):
        super().__init__()
        self.view_sampler = ViewSampler(**(view_sampler_args or {}))
        if feature_aggregator_class_type not in AGGREGATORS:
            raise NotImplementedError(f"feature aggregator {feature_aggregator_class_type} is not built "
                                      f"(have {sorted(AGGREGATORS)})")
        a = aggregator_args.get(f"feature_aggregator_{feature_aggregator_class_type}_args") or {}
        self.feature_aggregator = AGGREGATORS[feature_aggregator_class_type](**dict(a))

    def get_aggregated_feature_dim(self, feats) -> int:
        return self.feature_aggregator.get_aggregated_feature_dim(feats)

    @torch.no_grad()
    def forward(self, *, pts: torch.Tensor, seq_id_pts, camera: PerspectiveCameras, seq_id_camera, feats: Dict[str, torch.Tensor],
                masks: Optional[torch.Tensor], **kwargs) -> torch.Tensor:
        """-> (1, 1, n_pts, aggregated dim), as ViewPooler.forward of pytorch3d for ONE batch of points."""
        if pts.shape[0] != 1:
            raise NotImplementedError("one voxel grid per GPU (holo_diffusion_model.py:326)")
        vw = view_weights(seq_id_pts, seq_id_camera, pts.device)
        return pool_views(self, pts[0].reshape(-1, 3), camera, feats, masks, vw, mapper=None)[None, None]


def view_weights(seq_id_pts, seq_id_camera, device) -> Optional[torch.Tensor]:
    """camera_pts_mask of ViewSampler.forward: 1 where the camera comes from the points' sequence (None: all do)."""
    if seq_id_pts is None or seq_id_camera is None:
        return None
    s0 = seq_id_pts[0]
    m = [1.0 if s == s0 else 0.0 for s in seq_id_camera]
    return None if all(v == 1.0 for v in m) else torch.tensor(m, device=device)


def coord_grid(resol: int, volume_extent: float, device) -> torch.Tensor:
    """VolumeLocator(1, (R, R, R), voxel_size = extent / R).get_coord_grid() flattened: (R^3, 3) voxel centres, x fastest."""
    half = 0.5 * (resol - 1) * (volume_extent / resol)
    lin = torch.linspace(-1.0, 1.0, resol, device=device) * half
    z, y, x = torch.meshgrid(lin, lin, lin, indexing="ij")
    return torch.stack([x, y, z], -1).reshape(-1, 3).contiguous()


def _chunk_points(n_src: int) -> int:
    """Points per chunk.  Measured on B200 (profiles/r02j-r02l): bigger is faster -- the (view, point) activations do
    not stay in L2 at any useful chunk size, while every chunk costs six launches of host time -- so the default takes up
    to 4 Mi (view, point) rows at once: the whole 64^3 grid for up to 16 views, ~6.5 GB of operand / activation buffers."""
    env = os.environ.get("HOLO_VIEWPOOL_CHUNK")
    if env:
        return max(128, _ceil(int(env), 128))
    return max(1024, ((1 << 22) // max(n_src, 1)) // 128 * 128)


@torch.no_grad()
def pool_views(pooler, pts: torch.Tensor, camera, feats: Dict[str, torch.Tensor], masks: Optional[torch.Tensor],
               view_weight: Optional[torch.Tensor], mapper: Optional[nn.Module], out_cl: Optional[torch.Tensor] = None,
               pair_dtype=None, debug: Optional[dict] = None) -> torch.Tensor:
    """pts (P, 3) -> pooled rows.  mapper None: (P, aggregated dim) fp32 (the view pooler's own output);
    mapper given: (P, mapper outputs) fp32 rows of ``pooled_feature_mapper(pooled)`` BEFORE the tanh (channels-last
    grid rows; written into out_cl when given).  `pooler`: anything with ``.view_sampler.masked_sampling`` and
    ``.feature_aggregator`` (this module's classes, or pytorch3d's own objects built by the config system)."""
    dev = pts.device
    ops.require_cuda(dev, "pool_views")
    pair_dtype = PAIR_DTYPE if pair_dtype is None else pair_dtype
    # HOLO_VIEWPOOL_FUSE_ACT=1 folds the activations (and the per-point mean term) into the GEMM epilogues
    # (holo_gemm_tc_act).  Validated, bit-identical -- and SLOWER on B200 (4.25 vs 3.60 ms per 64^3 grid,
    # profiles/r02l): with K = 128 these GEMMs are epilogue-bound, and the epilogue is 4 warps per SM, so the hi/lo
    # conversion of 128 columns per row costs more there than as a full-occupancy pass at 0.8 of the HBM copy rate.
    fuse_act = os.environ.get("HOLO_VIEWPOOL_FUSE_ACT", "0") == "1"
    agg = pooler.feature_aggregator
    kind = type(agg).__name__
    if getattr(agg, "exclude_target_view", False) or getattr(agg, "exclude_target_view_mask_features", False):
        raise NotImplementedError("target-view exclusion (HoloDiffusionModel switches it off: holo_diffusion_model.py:115-116)")
    if getattr(pooler.view_sampler, "sampling_mode", "bilinear") != "bilinear":
        raise NotImplementedError("view sampling modes other than bilinear")
    pts = pts.contiguous().float()
    P = pts.shape[0]
    n_src = camera.R.shape[0]
    # channels-last maps (a tap = contiguous channels), every map zero-padded to a multiple of 4 channels: a lane of the
    # gather kernels owns 4 columns of ONE map and reads a tap as one float4.  `cols`: where the reference's columns sit
    # in the padded row -- the weight matrices get zero columns at the padding, nothing else changes.
    maps, chans, starts = [], [], []
    for f in feats.values():
        m = f.detach().float().permute(0, 2, 3, 1)
        c = m.shape[3]
        starts.append(sum(mm.shape[3] for mm in maps))
        chans.append(c)
        maps.append((nn.functional.pad(m, (0, _ceil(c, 4) - c)) if c % 4 else m).contiguous())
    F_ = sum(chans)
    F_pad = sum(m.shape[3] for m in maps)
    mask_map = None
    if getattr(pooler.view_sampler, "masked_sampling", False):
        if masks is None:
            raise ValueError("masked_sampling needs the masks")
        mask_map = masks.detach().float().reshape(n_src, masks.shape[-2], masks.shape[-1]).contiguous()
    cam = (camera.R.float().contiguous(), camera.T.float().contiguous(),
           camera.focal_length.float().expand(n_src, 2).contiguous(), camera.principal_point.float().expand(n_src, 2).contiguous())
    mlp_mean = kind == "MLPMeanFeatureAggregator"
    if not mlp_mean and kind != "AngleWeightedReductionFeatureAggregator":
        raise NotImplementedError(f"feature aggregator {kind} is not built")
    if mlp_mean:
        n_harm = agg.n_harmonic_functions_ray
        E = 3 * (2 * n_harm + 1)
        Kx, Kx_pad = F_ + E, F_pad + E
        cols = [st + c for st, ch in zip(starts, chans) for c in range(ch)] + [F_pad + e for e in range(E)]
    else:
        red = _reductions(agg)
        if red not in (("AVG",), ("AVG", "STD")):
            raise NotImplementedError(f"reduction functions {red}: AVG and AVG+STD are built")
        per = len(red)
        Kx, Kx_pad = F_ * per, F_pad * per
        # pooled row of the reference: [mu_k | std_k] per feature; padded: the same with C_k rounded up
        cols = [per * st + r * maps[k].shape[3] + c for k, (st, ch) in enumerate(zip(starts, chans)) for r in range(per)
                for c in range(ch)]
    cols_cache = pooler.__dict__.setdefault("_holo_cols", {})
    cols_t = cols_cache.get((tuple(cols), str(dev)))
    if cols_t is None:   # a host -> device copy: once per layout, not per call
        cols_cache.clear()
        cols_t = cols_cache[(tuple(cols), str(dev))] = torch.tensor(cols, device=dev, dtype=torch.long)
    Kpad = _ceil(Kx_pad, 64)
    if Kpad > 256:
        raise NotImplementedError(f"{Kx_pad} pooled columns: the kernels hold rows of up to 256")
    plans = pooler.__dict__.setdefault("_holo_plans", {})
    extra = (tuple(cols), pair_dtype, str(dev))
    key = _param_key([agg, mapper], extra)
    plan = plans.get("plan")
    if plan is None or plan["key"] != key:
        if mlp_mean:
            plan = _build_mlp_mean_plan(agg, mapper, Kx, cols_t, Kx_pad, Kpad, pair_dtype, dev)
        else:
            if mapper is not None:
                _materialize(mapper, Kx, dev)
            plan = {"mapper": None if mapper is None else _Gemm(_scatter_cols(mapper.weight.detach(), cols_t, Kx_pad),
                                                                mapper.bias.detach(), Kpad, pair_dtype, n_mult=16), "D": Kx}
        plan["key"] = _param_key([agg, mapper], extra)   # lazy layers materialised: new pointers
        plans["plan"] = plan
    mp = plan["mapper"]
    n_out = plan["D"] if mp is None else mp.n
    n_out_pad = _ceil(plan["D"], 64) if mp is None else mp.n_pad
    chunk = min(_chunk_points(n_src), _ceil(P, 128))
    P_pad = _ceil(P, chunk)
    direct = out_cl is not None and mp is not None and out_cl.shape == (P_pad, n_out_pad) and out_cl.is_contiguous()
    rows_out = out_cl if direct else torch.empty(P_pad, n_out_pad, device=dev)
    # work buffers of one chunk, kept on the pooler between calls (zero-filled once: rows / columns that no kernel
    # writes -- the tail of a ragged last chunk -- must hold finite values for the GEMMs)
    ws_key = (n_src, chunk, Kpad, pair_dtype, str(dev), mlp_mean, key[1] if not mlp_mean else tuple(
        [plan["first"].n_pad, plan["last"].n_pad] + [g.n_pad for g in plan["hidden"]]))
    ws = pooler.__dict__.get("_holo_ws")
    if ws is None or ws["key"] != ws_key:
        z16 = lambda r, c: (torch.zeros(r, c, device=dev, dtype=pair_dtype), torch.zeros(r, c, device=dev, dtype=pair_dtype))  # noqa: E731
        ws = {"key": ws_key}
        if mlp_mean:
            Hp, Dp = plan["first"].n_pad, plan["last"].n_pad
            widths = [Hp] + [g.n_pad for g in plan["hidden"]]
            ws["x"], ws["m"], ws["h"], ws["g"] = z16(n_src * chunk, Kpad), z16(chunk, Kpad), z16(n_src * chunk, max(widths)), z16(chunk, Dp)
            ws["h2"] = z16(n_src * chunk, max(widths)) if plan["hidden"] else None   # a GEMM cannot write its own operand
            ws["y"] = torch.zeros(n_src * chunk, max(widths + [Dp]), device=dev)
            ws["mterm"], ws["pooled"] = torch.zeros(chunk, Hp, device=dev), torch.zeros(chunk, Dp, device=dev)
        else:
            ws["g"], ws["pooled"] = z16(chunk, Kpad), torch.zeros(chunk, Kx_pad, device=dev)
        pooler.__dict__["_holo_ws"] = ws
    (g_hi, g_lo), pooled = ws["g"], ws["pooled"] if (mp is None or debug is not None) else None
    if mlp_mean:
        Hp, Dp = plan["first"].n_pad, plan["last"].n_pad
        (x_hi, x_lo), (m_hi, m_lo), (h_hi, h_lo), y, mterm = ws["x"], ws["m"], ws["h"], ws["y"], ws["mterm"]
    for c0 in range(0, P, chunk):
        n = min(chunk, P - c0)
        pc = pts[c0:c0 + n]
        if mlp_mean:
            dbg_x = dbg_m = None
            if debug is not None and c0 == 0:
                dbg_x, dbg_m = torch.empty(n_src, n, Kx_pad, device=dev), torch.empty(n, Kx_pad, device=dev)
            ops.viewpool_sample(pc, *cam, maps, n_harm, Kpad, chunk, x_hi, x_lo, m_hi, m_lo, mask_map=mask_map,
                                view_weight=view_weight, x_f32=dbg_x, mean_f32=dbg_m)
            rows = n_src * chunk
            plan["mean"](m_hi, m_lo, chunk, mterm)
            hh, hl = h_hi.view(-1)[: rows * Hp].view(rows, Hp), h_lo.view(-1)[: rows * Hp].view(rows, Hp)
            if fuse_act:   # H = act(X A^T + M[p]) straight out of the GEMM epilogue as the next operand pair
                plan["first"](x_hi, x_lo, rows, None, hh, hl, act=plan["acts"][0], row_term=mterm, row_term_rows=chunk)
            else:
                y0 = y.view(-1)[: rows * Hp].view(rows, Hp)
                plan["first"](x_hi, x_lo, rows, y0)
                ops.viewpool_act_split(y0, mterm, n_src, chunk, Hp, plan["acts"][0], hh, hl)
            if debug is not None and c0 == 0:
                debug.update(x=dbg_x[..., cols_t], mean=dbg_m[..., cols_t], mterm=mterm[:, : plan["first"].n].clone())
            for li, g in enumerate(plan["hidden"]):
                # ping-pong between the two halves of the operand buffer (hidden widths are equal: n_hidden)
                nh, nl = (ws["h2"][0], ws["h2"][1]) if li % 2 == 0 else (h_hi, h_lo)
                nh, nl = nh.view(-1)[: rows * g.n_pad].view(rows, g.n_pad), nl.view(-1)[: rows * g.n_pad].view(rows, g.n_pad)
                if fuse_act:
                    g(hh, hl, rows, None, nh, nl, act=plan["acts"][li + 1])
                else:
                    yi = y.view(-1)[: rows * g.n_pad].view(rows, g.n_pad)
                    g(hh, hl, rows, yi)
                    ops.viewpool_act_split(yi, None, n_src, chunk, g.n_pad, plan["acts"][li + 1], nh, nl)
                hh, hl = nh, nl
            z = y.view(-1)[: rows * Dp].view(rows, Dp)
            plan["last"](hh, hl, rows, z)
            ops.viewpool_reduce(z, n_src, chunk, n, Dp, out=pooled, out_hi=g_hi, out_lo=g_lo)
            if debug is not None and c0 == 0:
                debug.update(z=z.view(n_src, chunk, Dp)[:, :n, : plan["D"]].clone(), pooled=pooled[:n, : plan["D"]].clone())
        else:
            ops.viewpool_angle_reduce(pc, *cam, maps, Kpad, g_hi, g_lo, mask_map=mask_map, view_weight=view_weight,
                                      gamma=float(agg.weight_by_ray_angle_gamma),
                                      min_ray_angle_weight=float(agg.min_ray_angle_weight), with_std=len(red) == 2,
                                      out_f32=None if pooled is None else pooled[:n])
            if debug is not None and c0 == 0:
                debug.update(pooled=pooled[:n][:, cols_t])
        if mp is None:
            rows_out[c0:c0 + n, : plan["D"]] = pooled[:n, : plan["D"]] if mlp_mean else pooled[:n][:, cols_t]
        else:
            mp(g_hi, g_lo, chunk, rows_out[c0:c0 + chunk])
    if direct:
        return rows_out
    return rows_out[:P, :n_out].contiguous() if (P_pad != P or n_out_pad != n_out) else rows_out


# ------------------------------------------------------------------------------------------------ image features
class ResNetFeatureExtractor(nn.Module):
    """pytorch3d ``ResNetFeatureExtractor`` restated (configs/base.yaml:161-164 sets proj_dim 16, image_rescale 0.32):
    torchvision ResNet stem + stages, a 1x1 projection and an L2 normalisation per requested stage, plus the (rescaled)
    masks and images as extra features.  A library network (cuDNN) on both sides; ``pretrained`` weights cannot be
    downloaded here -- load them through the model's state dict (``stem.*``, ``layers.*``, ``proj_layers.*``)."""

    _DIMS = {"resnet18": (64, 128, 256, 512), "resnet34": (64, 128, 256, 512), "resnet50": (256, 512, 1024, 2048),
             "resnet101": (256, 512, 1024, 2048), "resnet152": (256, 512, 1024, 2048)}
    _MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    def __init__(self, name: str = "resnet34", pretrained: bool = True, stages: Sequence[int] = (1, 2, 3, 4),
                 normalize_image: bool = True, image_rescale: float = 128.0 / 800.0, first_max_pool: bool = True,
                 proj_dim: int = 32, l2_norm: bool = True, add_masks: bool = True, add_images: bool = True,
                 global_average_pool: bool = False, feature_rescale: float = 1.0):
        super().__init__()
        import torchvision
        if global_average_pool:
            raise NotImplementedError("global_average_pool")
        self.stages, self.normalize_image, self.image_rescale = tuple(stages), normalize_image, image_rescale
        self.proj_dim, self.l2_norm, self.add_masks, self.add_images = proj_dim, l2_norm, add_masks, add_images
        self.feature_rescale = feature_rescale
        self._feat_dim: Dict[str, int] = {}
        self.layers, self.proj_layers = nn.ModuleList(), nn.ModuleList()
        if len(self.stages) > 0:
            net = getattr(torchvision.models, name)(weights=None)   # no network: `pretrained` comes with the checkpoint
            self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool) if first_max_pool else \
                nn.Sequential(net.conv1, net.bn1, net.relu)
            for stage in range(max(self.stages)):
                dim = self._DIMS[name][stage]
                if (stage + 1) in self.stages:
                    if proj_dim > 0:
                        self.proj_layers.append(nn.Conv2d(dim, proj_dim, 1))
                        dim = proj_dim
                    else:
                        self.proj_layers.append(nn.Identity())
                    self._feat_dim[f"res_layer_{stage + 1}"] = dim
                else:
                    self.proj_layers.append(nn.Identity())
                self.layers.append(getattr(net, f"layer{stage + 1}"))
        if add_masks:
            self._feat_dim[MASK_FEATURE_NAME] = 1
        if add_images:
            self._feat_dim[IMAGE_FEATURE_NAME] = 3
        self.register_buffer("_resnet_mean", torch.tensor(self._MEAN).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("_resnet_std", torch.tensor(self._STD).view(1, 3, 1, 1), persistent=False)

    def get_feat_dims(self) -> int:
        return sum(self._feat_dim.values())

    @torch.no_grad()
    def forward(self, imgs: Optional[torch.Tensor], masks: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, torch.Tensor]:
        feats: Dict[str, torch.Tensor] = {}
        if imgs is not None:
            if self.normalize_image:
                imgs = (imgs - self._resnet_mean) / self._resnet_std
            if not math.isclose(self.image_rescale, 1.0):
                imgs = nn.functional.interpolate(imgs, scale_factor=self.image_rescale, mode="bilinear")
            if len(self.stages) > 0:
                feat = self.stem(imgs)
                for stage, (layer, proj) in enumerate(zip(self.layers, self.proj_layers)):
                    feat = layer(feat)
                    if (stage + 1) in self.stages:
                        f = proj(feat)
                        if self.l2_norm:
                            f = nn.functional.normalize(f, dim=1) * (1.0 / math.sqrt(len(self.stages)))
                        feats[f"res_layer_{stage + 1}"] = f
        if self.add_masks:
            if masks is None:
                raise ValueError("add_masks needs the foreground masks")
            feats[MASK_FEATURE_NAME] = masks
        if self.add_images and imgs is not None:
            feats[IMAGE_FEATURE_NAME] = imgs
        if self.feature_rescale != 1.0:
            feats = {k: self.feature_rescale * f for k, f in feats.items()}
        return feats


def mask_background(image_rgb: torch.Tensor, fg_probability: torch.Tensor, mask_threshold: float, bg_color) -> torch.Tensor:
    """pytorch3d preprocess_input's image masking (holo_diffusion_model.py:248-257 calls it with the model's mask_images /
    mask_threshold / bg_color): pixels whose foreground probability is below the threshold take the background colour."""
    fg = (fg_probability > mask_threshold).to(image_rgb)
    bg = image_rgb.new_tensor(bg_color).view(1, 3, 1, 1)
    return fg * image_rgb + (1.0 - fg) * bg


def select_sources(sequence_name: Optional[List[str]], batch_size: int, n_targets: int) -> List[int]:
    """safe_slice_sources of the reference's forward (holo_diffusion_model.py:278-311): the views of the first view's
    sequence after the targets; everything when nothing is left."""
    if sequence_name is None:
        sel = list(range(n_targets, batch_size))
    else:
        ok = [i for i, s in enumerate(sequence_name) if s == sequence_name[0]]
        sel = ok[n_targets:]
    return sel if sel else list(range(batch_size))
