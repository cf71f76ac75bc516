// Backward of the staged renderer (renderer half): what `objective.backward()`
// (/root/reference/trainer/training_loop.py:518-556) propagates through
//   EmissionAbsorptionRaymarcher.forward            (pytorch3d 0.7.4; called holo_multipass_ea.py:96-100, density noise :87-91)
//   HoloVoxelGridImplicitFunction.forward           (holo_voxel_grid_implicit_function.py:182-269: trilinear grid_sample :210-221,
//                                                    RenderMLP :107-129)
// for the 1024 `mask_sample` rays of a training step (configs/base.yaml:132-134).  The ray-point refiner runs under
// no_grad in the reference (ray_point_refiner), the ray lengths carry no gradient.
//
//   holo_ea_raymarch_bwd : dL/d(features, depths, masks[, weights]) -> dL/d(densities), dL/d(ray features)
//   holo_if_bwd          : dL/d(densities, rgb) -> dL/d(voxel grid) (trilinear scatter-add) and dL/d(collapsed MLP):
//                          W_eff (H+1, C), b_eff (H+1), radiance layer Wr (3, H+E), br (3).  The chain from (W_eff, b_eff)
//                          to the four density-net layers is the (tiny) affine composition, differentiated on the host side.
// Forward activations are recomputed (nothing but the inputs is saved).
#include "common.cuh"
#include "render_device.cuh"
#include "../../include/holo_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
// EmissionAbsorptionRaymarcher backward, one thread per ray.
//   wd_k = delta_k relu(sigma_k + noise_k);  C_k = sum_{j<k} wd_j;  w_k = (1 - e^{-wd_k}) e^{-C_k}
//   features = sum_k w_k f_k + (1 - mask) bg;  depth = sum_k w_k t_k;  mask = 1 - e^{-C_S}
//   d wd_k = gw_k e^{-wd_k} e^{-C_k} - sum_{m>k} gw_m w_m + gmask_total e^{-C_S},   gw_k = gF . f_k + gD t_k + gW_k
// ------------------------------------------------------------------------------------------------
__global__ void ea_bwd_kernel(const float* __restrict__ dens, const float* __restrict__ noise, const float* __restrict__ feats,
                              const float* __restrict__ lengths, int n_rays, int S, int Fd, float bg_opacity,
                              const float* __restrict__ bg /*[Fd] device*/, const float* __restrict__ g_feat,
                              const float* __restrict__ g_depth, const float* __restrict__ g_mask,
                              const float* __restrict__ g_weights, float* __restrict__ d_dens, float* __restrict__ d_feats) {
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float* z = lengths + (size_t)ray * S;
    const float* sg = dens + (size_t)ray * S;
    const float* nz = noise ? noise + (size_t)ray * S : nullptr;
    const float* f = feats + (size_t)ray * S * Fd;
    const float* gF = g_feat + (size_t)ray * Fd;
    const float gD = g_depth ? g_depth[ray] : 0.f;
    // forward recompute: C_k = sum_{j<k} wd_j, parked in this ray's slice of d_dens until the reverse sweep overwrites it
    float* ck_buf = d_dens + (size_t)ray * S;
    float cum = 0.f;
    for (int s = 0; s < S; ++s) {
        const float delta = (s + 1 < S) ? (z[s + 1] - z[s]) : bg_opacity;
        const float d = sg[s] + (nz ? nz[s] : 0.f);
        ck_buf[s] = cum;
        cum += delta * fmaxf(d, 0.f);
    }
    const float T_all = expf(-cum);            // 1 - mask
    float g_m = g_mask ? g_mask[ray] : 0.f;    // dL/dmask, plus the background term of the features
    for (int c = 0; c < Fd; ++c) g_m -= gF[c] * bg[c];
    const float g_cum_all = g_m * T_all;       // mask = 1 - e^{-C_S}
    // reverse sweep: suffix = sum_{m>k} gw_m w_m
    float suffix = 0.f;
    for (int s = S - 1; s >= 0; --s) {
        const float delta = (s + 1 < S) ? (z[s + 1] - z[s]) : bg_opacity;
        const float d = sg[s] + (nz ? nz[s] : 0.f);
        const float wd = delta * fmaxf(d, 0.f);
        const float Tk = expf(-ck_buf[s]), ewd = expf(-wd);
        const float w = (1.f - ewd) * Tk;
        float gw = gD * z[s] + (g_weights ? g_weights[(size_t)ray * S + s] : 0.f);
        for (int c = 0; c < Fd; ++c) {
            gw = fmaf(gF[c], f[(size_t)s * Fd + c], gw);
            d_feats[((size_t)ray * S + s) * Fd + c] = w * gF[c];
        }
        const float g_wd = gw * ewd * Tk - suffix + g_cum_all;
        ck_buf[s] = d > 0.f ? delta * g_wd : 0.f;   // relu backward: 0 at d <= 0
        suffix = fmaf(gw, w, suffix);
    }
}

// ------------------------------------------------------------------------------------------------
// Implicit function backward: a warp handles 32 points; parameter gradients are reduced over the warp with shuffles,
// accumulated in shared memory per CTA and flushed with one global atomic per value.
// ------------------------------------------------------------------------------------------------
struct IfBwdParams {
    const float* grid;
    int D, Hh, Ww;
    float isx, isy, isz;
    const float* mlp;
    int Hd, n_harm;
    const float* origins;
    const float* dirs;
    const float* lengths;
    long long P;
    int S;
    const float* g_dens;   // (P)
    const float* g_rgb;    // (P, 3)
    float* d_grid;         // (D,H,W,C), accumulated (caller zeroes)
    float* d_W;            // (Hd+1, C)
    float* d_b;            // (Hd+1)
    float* d_Wr;           // (3, Hd + E)
    float* d_br;           // (3)
};

__device__ __forceinline__ float leaky_grad(float pre) { return pre > 0.f ? 1.f : 0.2f; }

template <int C>
__global__ void __launch_bounds__(128) if_points_bwd_kernel(IfBwdParams P) {
    extern __shared__ __align__(16) float smem[];
    const int Hd = P.Hd;
    const int E = 3 * (2 * P.n_harm + 1);
    float* sW = smem;
    float4* sEp = reinterpret_cast<float4*>(sW + (Hd + 1) * C);
    float* sSig = reinterpret_cast<float*>(sEp + Hd);
    float* sDir = sSig + 4;
    int n_pack = (Hd + 1) * C + 4 * Hd + 4 + 3 * E + 3;
    n_pack = (n_pack + 3) / 4 * 4;
    float* gW = smem + n_pack;                 // (Hd+1, C)
    float* gb = gW + (Hd + 1) * C;             // (Hd+1)
    float* gWr = gb + (Hd + 1);                // (3, Hd+E)
    float* gbr = gWr + 3 * (Hd + E);           // (3)
    const int n_grad = (Hd + 1) * C + (Hd + 1) + 3 * (Hd + E) + 3;
    for (int i = threadIdx.x; i < n_pack; i += blockDim.x) smem[i] = P.mlp[i];
    for (int i = threadIdx.x; i < n_grad; i += blockDim.x) gW[i] = 0.f;
    __syncthreads();
    const float b_sigma = sSig[0];
    const float* wsig = sW + (size_t)Hd * C;
    const int lane = threadIdx.x & 31;
    const long long n_warp_iters = (P.P + 31) / 32;

    for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; wi < n_warp_iters;
         wi += ((long long)gridDim.x * blockDim.x) / 32) {
        const long long p = wi * 32 + lane;
        const bool ok = p < P.P;
        const long long pp = ok ? p : 0;
        const long long ray = pp / P.S;
        // ---- forward recompute: direction embedding, sampled feature, corner table
        float d[3] = {P.dirs[ray * 3 + 0], P.dirs[ray * 3 + 1], P.dirs[ray * 3 + 2]};
        const float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
        float dn[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
        float rd[3];
        dir_radiance_const(sDir, E, P.n_harm, dn, rd);
        const float zl = P.lengths[pp];
        const float lx = (P.origins[ray * 3 + 0] + zl * d[0]) * P.isx;
        const float ly = (P.origins[ray * 3 + 1] + zl * d[1]) * P.isy;
        const float lz = (P.origins[ray * 3 + 2] + zl * d[2]) * P.isz;
        float x[C];
        sample_trilinear<C>(P.grid, P.D, P.Hh, P.Ww, lx, ly, lz, x);
        // pass 1: radiance pre-activations
        float r0 = rd[0], r1 = rd[1], r2 = rd[2];
        for (int j = 0; j < Hd; ++j) {
            const float4* wrow = reinterpret_cast<const float4*>(sW + (size_t)j * C);
            const float4 ep = sEp[j];
            float a = ep.w;
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 w = wrow[c4];
                a = fmaf(w.x, x[c4 * 4 + 0], a), a = fmaf(w.y, x[c4 * 4 + 1], a);
                a = fmaf(w.z, x[c4 * 4 + 2], a), a = fmaf(w.w, x[c4 * 4 + 3], a);
            }
            const float h = holo_leaky(a);
            r0 = fmaf(ep.x, h, r0), r1 = fmaf(ep.y, h, r1), r2 = fmaf(ep.z, h, r2);
        }
        float gr[3];
        {
            const float r[3] = {r0, r1, r2};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float s = 1.f / (1.f + expf(-holo_leaky(r[i])));
                gr[i] = ok ? P.g_rgb[pp * 3 + i] * s * (1.f - s) * leaky_grad(r[i]) : 0.f;
            }
        }
        // radiance bias and direction block: d rd = gr
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float t = warp_sum(gr[i]);
            if (lane == 0) atomicAdd(&gbr[i], t);
        }
        for (int c = 0; c < 3; ++c) {
            float freq = 1.f;
            for (int k = 0; k < P.n_harm; ++k) {
                const float e = dn[c] * freq;
                const float sn = sinf(e), cs = cosf(e);
                const int ms = c * P.n_harm + k, mc = 3 * P.n_harm + c * P.n_harm + k;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float ts = warp_sum(gr[i] * sn), tc = warp_sum(gr[i] * cs);
                    if (lane == 0) atomicAdd(&gWr[i * (Hd + E) + Hd + ms], ts), atomicAdd(&gWr[i * (Hd + E) + Hd + mc], tc);
                }
                freq *= 2.f;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float t = warp_sum(gr[i] * dn[c]);
                if (lane == 0) atomicAdd(&gWr[i * (Hd + E) + Hd + 6 * P.n_harm + c], t);
            }
        }
        // pass 2: hidden units again -> dWr (hidden part), da_j, dW_eff, db_eff, dx
        float dx[C];
#pragma unroll
        for (int c = 0; c < C; ++c) dx[c] = 0.f;
        for (int j = 0; j <= Hd; ++j) {
            const float4* wrow = reinterpret_cast<const float4*>(sW + (size_t)j * C);
            float a = j < Hd ? sEp[j].w : b_sigma;
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 w = wrow[c4];
                a = fmaf(w.x, x[c4 * 4 + 0], a), a = fmaf(w.y, x[c4 * 4 + 1], a);
                a = fmaf(w.z, x[c4 * 4 + 2], a), a = fmaf(w.w, x[c4 * 4 + 3], a);
            }
            float da;
            if (j < Hd) {
                const float4 ep = sEp[j];
                const float h = holo_leaky(a);
                const float t0 = warp_sum(gr[0] * h), t1 = warp_sum(gr[1] * h), t2 = warp_sum(gr[2] * h);
                if (lane == 0)
                    atomicAdd(&gWr[0 * (Hd + E) + j], t0), atomicAdd(&gWr[1 * (Hd + E) + j], t1), atomicAdd(&gWr[2 * (Hd + E) + j], t2);
                da = (ep.x * gr[0] + ep.y * gr[1] + ep.z * gr[2]) * leaky_grad(a);
            } else {
                da = ok ? P.g_dens[pp] * leaky_grad(a) : 0.f;
            }
            const float tb = warp_sum(da);
            if (lane == 0) atomicAdd(&gb[j], tb);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float t = warp_sum(da * x[c]);
                if (lane == 0) atomicAdd(&gW[j * C + c], t);
                dx[c] = fmaf(da, sW[(size_t)j * C + c], dx[c]);
            }
        }
        // ---- scatter dx into the grid with the trilinear weights (grid_sampler_3d_backward w.r.t. the input)
        if (ok) {
            const float ix = ((lx + 1.f) / 2.f) * (float)(P.Ww - 1);
            const float iy = ((ly + 1.f) / 2.f) * (float)(P.Hh - 1);
            const float iz = ((lz + 1.f) / 2.f) * (float)(P.D - 1);
            const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
            const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)P.Ww + 1.f);
            const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)P.Hh + 1.f);
            const int z0 = (int)fminf(fmaxf(fz0, -2.f), (float)P.D + 1.f);
            const float wx1 = ix - fx0, wy1 = iy - fy0, wz1 = iz - fz0;
            const float wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy, wz0 = (fz0 + 1.f) - iz;
#pragma unroll
            for (int corner = 0; corner < 8; ++corner) {
                const int ddx = corner & 1, ddy = (corner >> 1) & 1, ddz = corner >> 2;
                const int xx = x0 + ddx, yy = y0 + ddy, zz = z0 + ddz;
                const float w = (ddx ? wx1 : wx0) * (ddy ? wy1 : wy0) * (ddz ? wz1 : wz0);
                if (xx >= 0 && xx < P.Ww && yy >= 0 && yy < P.Hh && zz >= 0 && zz < P.D) {
                    float* gp = P.d_grid + (((size_t)zz * P.Hh + yy) * P.Ww + xx) * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) atomicAdd(gp + c, w * dx[c]);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (Hd + 1) * C; i += blockDim.x) atomicAdd(&P.d_W[i], gW[i]);
    for (int i = threadIdx.x; i < Hd + 1; i += blockDim.x) atomicAdd(&P.d_b[i], gb[i]);
    for (int i = threadIdx.x; i < 3 * (Hd + E); i += blockDim.x) atomicAdd(&P.d_Wr[i], gWr[i]);
    for (int i = threadIdx.x; i < 3; i += blockDim.x) atomicAdd(&P.d_br[i], gbr[i]);
}

template <int C>
int launch_if_bwd(const IfBwdParams& P, cudaStream_t st) {
    const int E = 3 * (2 * P.n_harm + 1);
    size_t n_pack = (size_t)(P.Hd + 1) * C + 4 * P.Hd + 4 + 3 * E + 3;
    n_pack = (n_pack + 3) / 4 * 4;
    const size_t n_grad = (size_t)(P.Hd + 1) * C + (P.Hd + 1) + 3 * (P.Hd + E) + 3;
    const size_t smem = (n_pack + n_grad) * sizeof(float);
    if (smem > 227 * 1024) {
        holo_set_error("holo_if_bwd: hidden %d x C %d does not fit shared memory", P.Hd, C);
        return HOLO_ERR_UNSUPPORTED;
    }
    int blocks = holo_cdiv(P.P, 128);
    if (blocks > 148 * 2) blocks = 148 * 2;
    auto k = if_points_bwd_kernel<C>;
    HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "holo_if_bwd");
    k<<<blocks, 128, smem, st>>>(P);
    HOLO_CHECK_LAUNCH("holo_if_bwd");
    return HOLO_OK;
}

}  // namespace

extern "C" int holo_ea_raymarch_bwd(const float* densities, const float* features, const float* lengths,
                                    const float* density_noise, int n_rays, int S, int feat_dim, const float* bg_dev,
                                    float background_opacity, const float* grad_features, const float* grad_depths,
                                    const float* grad_masks, const float* grad_weights, float* d_densities,
                                    float* d_features, void* stream) {
    if (n_rays == 0) return HOLO_OK;
    HOLO_CHECK_ARG(densities && features && lengths && bg_dev && grad_features && d_densities && d_features,
                   "holo_ea_raymarch_bwd: null pointer");
    HOLO_CHECK_ARG(S >= 1 && feat_dim >= 1 && feat_dim <= 128, "holo_ea_raymarch_bwd: 1 <= feat_dim <= 128");
    ea_bwd_kernel<<<holo_cdiv(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        densities, density_noise, features, lengths, n_rays, S, feat_dim, background_opacity, bg_dev, grad_features, grad_depths,
        grad_masks, grad_weights, d_densities, d_features);
    HOLO_CHECK_LAUNCH("holo_ea_raymarch_bwd");
    return HOLO_OK;
}

extern "C" int holo_if_bwd(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent, const float* packed_mlp,
                           int hidden, int n_harmonic, const float* origins, const float* dirs, const float* lengths,
                           long long n_points, int S, const float* grad_densities, const float* grad_rgb, float* d_grid_dhwc,
                           float* d_W_eff, float* d_b_eff, float* d_Wr, float* d_br, void* stream) {
    if (n_points == 0) return HOLO_OK;
    HOLO_CHECK_ARG(grid_dhwc && packed_mlp && origins && dirs && lengths && grad_densities && grad_rgb && d_grid_dhwc &&
                       d_W_eff && d_b_eff && d_Wr && d_br,
                   "holo_if_bwd: null pointer");
    HOLO_CHECK_ARG(D > 1 && H > 1 && W > 1 && S >= 1 && n_points % S == 0, "holo_if_bwd: bad shape");
    IfBwdParams P;
    memset(&P, 0, sizeof(P));
    P.grid = grid_dhwc, P.D = D, P.Hh = H, P.Ww = W;
    P.isx = 1.0f / ((float)(W - 1) * (volume_extent / (float)W) * 0.5f);
    P.isy = 1.0f / ((float)(H - 1) * (volume_extent / (float)H) * 0.5f);
    P.isz = 1.0f / ((float)(D - 1) * (volume_extent / (float)D) * 0.5f);
    P.mlp = packed_mlp, P.Hd = hidden, P.n_harm = n_harmonic;
    P.origins = origins, P.dirs = dirs, P.lengths = lengths, P.P = n_points, P.S = S;
    P.g_dens = grad_densities, P.g_rgb = grad_rgb;
    P.d_grid = d_grid_dhwc, P.d_W = d_W_eff, P.d_b = d_b_eff, P.d_Wr = d_Wr, P.d_br = d_br;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 8: return launch_if_bwd<8>(P, st);
        case 16: return launch_if_bwd<16>(P, st);
        case 32: return launch_if_bwd<32>(P, st);
        case 64: return launch_if_bwd<64>(P, st);
        default:
            holo_set_error("holo_if_bwd: unsupported channel count %d (8, 16, 32, 64)", C);
            return HOLO_ERR_UNSUPPORTED;
    }
}
