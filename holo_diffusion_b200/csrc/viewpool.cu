// View-pooling encoder (views -> voxel grid), SURVEY.md section 8(f) row 2: the gather / reduction kernels around the
// tcgen05 GEMMs (holo_gemm_tc) that carry the aggregator's Linear layers.
//
//   reference: holo_diffusion_model.py:327-373 (VolumeLocator grid -> view_pooler -> pooled_feature_mapper -> tanh),
//              custom_modules.py:162-334 (MLPMeanFeatureAggregator, _get_point_to_source_camera_ray_dirs), and the
//              pytorch3d 0.7.4 leaves it calls (ViewSampler / project_points_and_sample / ndc_grid_sample /
//              HarmonicEmbedding / wmean), restated from memory -- see oracle/encoder_oracle.py.
//
// Per chunk of points the host (holo_diffusion_b200/encoder.py) runs
//   holo_viewpool_sample   X[s][p] = cat(sampled features, harmonic(ray dir)) * w[s][p]  (16-bit hi/lo pair rows, K padded)
//                          mean[p] = sum_s X[s][p] w[s][p] / max(sum_s w[s][p], 1e-2)
//   holo_gemm_tc           Y = X A^T,  M = mean B^T + b           (first Linear pair folded into the MLP's first Linear)
//   holo_viewpool_act_split  H = act(Y[s][p] + M[p]) as a pair
//   holo_gemm_tc           ... further MLP layers ..., Z = H W_last^T + b_last
//   holo_viewpool_reduce   G[p] = sum_s softmax_s(Z[s][p][0]) Z[s][p]   (pair)
//   holo_gemm_tc           grid rows = G W_map^T + b_map ; holo_act_range applies tanh and writes both layouts.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "../../include/holo_b200.h"

namespace {

constexpr int VP_MAX_MAPS = 8;
constexpr int VP_MAX_K = 256;      // padded row length of X (features + ray embedding)

struct VpMap {
    const float* data;   // channels-last (n_src, H, W, C)
    int C, H, W, c0;     // c0: first column of this map in the concatenated row
};

struct VpParams {
    const float* pts;            // (n_pts, 3) world
    long long n_pts;
    const float *R, *T, *focal, *pp;   // (n_src,3,3) (n_src,3) (n_src,2) (n_src,2)
    int n_src;
    VpMap maps[VP_MAX_MAPS];
    int n_maps, F;               // F = sum of map channels
    const float* mask_map;       // optional (n_src, Hm, Wm): ViewSampler.masked_sampling (nearest)
    int Hm, Wm;
    const float* view_weight;    // optional (n_src): 1 where the camera belongs to the points' sequence
    int n_harm, E, Kx, Kpad;
    float eps;
    long long rows_per_view;
    uint16_t *x_hi, *x_lo, *m_hi, *m_lo;
    float *x_f32, *m_f32;        // optional fp32 copies (tests / the un-fused ViewPooler API): (n_src, n_pts, Kx), (n_pts, Kx)
    int pair_f16;
};

// F.grid_sample coordinates of an NDC location (pytorch3d ndc_to_grid_sample_coords: negate, scale the longer side)
__device__ __forceinline__ void ndc_to_pixel(float x_ndc, float y_ndc, int H, int W, float& ix, float& iy) {
    float gx = -x_ndc, gy = -y_ndc;
    if (H >= W) gy *= (float)W / (float)H;
    else gx *= (float)H / (float)W;
    ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f;   // align_corners = False
    iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
}

__device__ __forceinline__ float bilinear_cl(const float* __restrict__ img, int C, int H, int W, int c, float ix, float iy) {
    // zeros padding: a tap outside the image contributes nothing
    if (!(ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H)) return 0.f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const bool xa = x0 >= 0, xb = x1 < W, ya = y0 >= 0, yb = y1 < H;
    float v = 0.f;
    if (ya && xa) v += __ldg(img + ((size_t)y0 * W + x0) * C + c) * (wx0 * wy0);
    if (ya && xb) v += __ldg(img + ((size_t)y0 * W + x1) * C + c) * (wx1 * wy0);
    if (yb && xa) v += __ldg(img + ((size_t)y1 * W + x0) * C + c) * (wx0 * wy1);
    if (yb && xb) v += __ldg(img + ((size_t)y1 * W + x1) * C + c) * (wx1 * wy1);
    return v;
}

__device__ __forceinline__ float nearest_1ch(const float* __restrict__ img, int H, int W, float ix, float iy) {
    if (!(ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H)) return 0.f;
    const int x = (int)nearbyintf(ix), y = (int)nearbyintf(iy);   // round half to even, as F.grid_sample(mode="nearest")
    if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
    return __ldg(img + (size_t)y * W + x);
}

// Per (view, point): NDC projection, aggregation weight w = view_weight * sampled mask, unit vector centre -> point
__device__ __forceinline__ void view_geometry(const VpParams& P, int s, float px, float py, float pz, float& x_ndc,
                                              float& y_ndc, float& vw, float& w, float& dx, float& dy, float& dz) {
    const float* R = P.R + s * 9;
    const float t0 = __ldg(P.T + s * 3), t1 = __ldg(P.T + s * 3 + 1), t2 = __ldg(P.T + s * 3 + 2);
    // X_cam = X_world R + T (row vectors)
    const float xc = px * __ldg(R + 0) + py * __ldg(R + 3) + pz * __ldg(R + 6) + t0;
    const float yc = px * __ldg(R + 1) + py * __ldg(R + 4) + pz * __ldg(R + 7) + t1;
    const float zc = px * __ldg(R + 2) + py * __ldg(R + 5) + pz * __ldg(R + 8) + t2;
    // NDC projection with the sign-preserving clamp of the homogeneous divide (Transform3d.transform_points eps)
    const float sgn = zc < 0.f ? -1.f : 1.f;
    const float den = sgn * fmaxf(fabsf(zc), P.eps);
    x_ndc = (__ldg(P.focal + s * 2) * xc + __ldg(P.pp + s * 2) * zc) / den;
    y_ndc = (__ldg(P.focal + s * 2 + 1) * yc + __ldg(P.pp + s * 2 + 1) * zc) / den;
    vw = P.view_weight ? __ldg(P.view_weight + s) : 1.f;
    w = vw;
    if (P.mask_map) {
        float ix, iy;
        ndc_to_pixel(x_ndc, y_ndc, P.Hm, P.Wm, ix, iy);
        w *= nearest_1ch(P.mask_map + (size_t)s * P.Hm * P.Wm, P.Hm, P.Wm, ix, iy);
    }
    // ray direction point <- camera centre, centre = -T R^T (custom_modules.py:283-304), F.normalize eps 1e-12
    const float cx = -(t0 * __ldg(R + 0) + t1 * __ldg(R + 1) + t2 * __ldg(R + 2));
    const float cy = -(t0 * __ldg(R + 3) + t1 * __ldg(R + 4) + t2 * __ldg(R + 5));
    const float cz = -(t0 * __ldg(R + 6) + t1 * __ldg(R + 7) + t2 * __ldg(R + 8));
    dx = px - cx, dy = py - cy, dz = pz - cz;
    const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    dx *= inv, dy *= inv, dz *= inv;
}

__device__ __forceinline__ float sample_column(const VpParams& P, int s, int j, float x_ndc, float y_ndc) {
    VpMap m = P.maps[0];
#pragma unroll
    for (int k = 1; k < VP_MAX_MAPS; ++k)   // static indices: the parameter struct stays in the constant bank
        if (k < P.n_maps && j >= P.maps[k].c0) m = P.maps[k];
    float ix, iy;
    ndc_to_pixel(x_ndc, y_ndc, m.H, m.W, ix, iy);
    return bilinear_cl(m.data + (size_t)s * m.H * m.W * m.C, m.C, m.H, m.W, j - m.c0, ix, iy);
}

// One warp per point; lanes walk the columns of the row, the view loop is inside (the mean needs all views).
__global__ void __launch_bounds__(256) viewpool_sample_kernel(const VpParams P) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int ROUNDS = VP_MAX_K / 32;
    for (long long p = warp0; p < P.n_pts; p += n_warps) {
        const float px = __ldg(P.pts + p * 3), py = __ldg(P.pts + p * 3 + 1), pz = __ldg(P.pts + p * 3 + 2);
        float acc[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) acc[r] = 0.f;
        float wsum = 0.f;
        for (int s = 0; s < P.n_src; ++s) {
            float x_ndc, y_ndc, vw, w, dx, dy, dz;
            view_geometry(P, s, px, py, pz, x_ndc, y_ndc, vw, w, dx, dy, dz);
            const size_t row = ((size_t)s * P.rows_per_view + (size_t)p) * P.Kpad;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int j = r * 32 + lane;
                if (j < P.Kpad) {
                    float v = 0.f;
                    if (j < P.F) {
                        v = sample_column(P, s, j, x_ndc, y_ndc) * vw;
                    } else if (j < P.Kx) {
                        // HarmonicEmbedding(append_input): [sin(d_i 2^f)] (i major, f minor), [cos(...)], d
                        const int e = j - P.F, nh3 = 3 * P.n_harm;
                        if (e < 2 * nh3) {
                            const int q = e < nh3 ? e : e - nh3;
                            const int comp = q / P.n_harm, f = q - comp * P.n_harm;
                            const float d = comp == 0 ? dx : (comp == 1 ? dy : dz);
                            const float a = d * (float)(1 << f);
                            v = e < nh3 ? sinf(a) : cosf(a);
                        } else {
                            const int comp = e - 2 * nh3;
                            v = comp == 0 ? dx : (comp == 1 ? dy : dz);
                        }
                    }
                    v *= w;
                    acc[r] += v * w;
                    uint16_t h, l;
                    holo_split1(v, P.pair_f16 != 0, h, l);
                    P.x_hi[row + j] = h, P.x_lo[row + j] = l;
                    if (P.x_f32 && j < P.Kx) P.x_f32[((size_t)s * P.n_pts + (size_t)p) * P.Kx + j] = v;
                }
            }
            wsum += w;
        }
        const float invw = 1.f / fmaxf(wsum, 1e-2f);   // wmean(..., eps=1e-2)
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int j = r * 32 + lane;
            if (j < P.Kpad) {
                const float m = acc[r] * invw;
                uint16_t h, l;
                holo_split1(m, P.pair_f16 != 0, h, l);
                P.m_hi[(size_t)p * P.Kpad + j] = h, P.m_lo[(size_t)p * P.Kpad + j] = l;
                if (P.m_f32 && j < P.Kx) P.m_f32[(size_t)p * P.Kx + j] = m;
            }
        }
    }
}

// AngleWeightedReductionFeatureAggregator (pytorch3d 0.7.4, restated from memory; reduction_functions = AVG, STD):
//   w[s] = mask[s] * ((0.5 (d_s . d_0 + 1))^gamma + min_weight),  d_s = unit vector camera s -> point (camera 0 = the
//   first camera of the point batch), mu = sum w x / max(sum w, 1e-2), std = sqrt(max(sum w (x - mu)^2 / max(sum w, 1e-2), 1e-4));
//   row layout [mu_k | std_k] per feature map k.  Two passes over the views (the second re-gathers: the maps are L2 resident).
__global__ void __launch_bounds__(256) viewpool_angle_kernel(const VpParams P, float gamma, float min_weight, int with_std,
                                                             float* __restrict__ out_f32) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int ROUNDS = VP_MAX_K / 32;
    const int per = with_std ? 2 : 1;
    for (long long p = warp0; p < P.n_pts; p += n_warps) {
        const float px = __ldg(P.pts + p * 3), py = __ldg(P.pts + p * 3 + 1), pz = __ldg(P.pts + p * 3 + 2);
        float mu[ROUNDS], var[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) mu[r] = 0.f, var[r] = 0.f;
        float wsum = 0.f, d0x = 0.f, d0y = 0.f, d0z = 0.f;
        for (int pass = 0; pass < per; ++pass) {
            for (int s = 0; s < P.n_src; ++s) {
                float x_ndc, y_ndc, vw, w, dx, dy, dz;
                view_geometry(P, s, px, py, pz, x_ndc, y_ndc, vw, w, dx, dy, dz);
                if (s == 0) d0x = dx, d0y = dy, d0z = dz;
                const float a01 = 0.5f * (dx * d0x + dy * d0y + dz * d0z + 1.f);
                w *= (gamma == 1.f ? a01 : powf(a01, gamma)) + min_weight;
                if (pass == 0) wsum += w;
#pragma unroll
                for (int r = 0; r < ROUNDS; ++r) {
                    const int j = r * 32 + lane;
                    if (j < P.F) {
                        const float v = sample_column(P, s, j, x_ndc, y_ndc) * vw;
                        if (pass == 0) mu[r] += v * w;
                        else var[r] += (v - mu[r]) * (v - mu[r]) * w;
                    }
                }
            }
            if (pass == 0) {
                const float invw = 1.f / fmaxf(wsum, 1e-2f);
#pragma unroll
                for (int r = 0; r < ROUNDS; ++r) mu[r] *= invw;
            }
        }
        const float invw = 1.f / fmaxf(wsum, 1e-2f);
        // zero the padding columns, then scatter [mu_k | std_k]
        for (int j = per * P.F + lane; j < P.Kpad; j += 32) P.m_hi[(size_t)p * P.Kpad + j] = 0, P.m_lo[(size_t)p * P.Kpad + j] = 0;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int j = r * 32 + lane;
            if (j < P.F) {
                VpMap m = P.maps[0];
#pragma unroll
                for (int k = 1; k < VP_MAX_MAPS; ++k)
                    if (k < P.n_maps && j >= P.maps[k].c0) m = P.maps[k];
                const int col_mu = per * m.c0 + (j - m.c0), col_sd = col_mu + m.C;
                uint16_t h, l;
                holo_split1(mu[r], P.pair_f16 != 0, h, l);
                P.m_hi[(size_t)p * P.Kpad + col_mu] = h, P.m_lo[(size_t)p * P.Kpad + col_mu] = l;
                if (out_f32) out_f32[(size_t)p * per * P.F + col_mu] = mu[r];
                if (with_std) {
                    const float sd = sqrtf(fmaxf(var[r] * invw, 1e-4f));
                    holo_split1(sd, P.pair_f16 != 0, h, l);
                    P.m_hi[(size_t)p * P.Kpad + col_sd] = h, P.m_lo[(size_t)p * P.Kpad + col_sd] = l;
                    if (out_f32) out_f32[(size_t)p * per * P.F + col_sd] = sd;
                }
            }
        }
    }
}

__device__ __forceinline__ float vp_act(float x, int act) {
    switch (act) {
        case 1: return fmaxf(x, 0.f);
        case 2: return holo_leaky(x);
        case 3: return x > 20.f ? x : log1pf(expf(x));   // torch.nn.Softplus(beta=1, threshold=20)
        default: return x;
    }
}

// H = act(Y[s][p] + M[p]) -> hi/lo pair; four columns per thread
__global__ void __launch_bounds__(256) viewpool_act_split_kernel(const float* __restrict__ y, const float* __restrict__ pt_term,
                                                                 long long rows, long long rows_per_view, int C4, int act,
                                                                 uint2* __restrict__ hi, uint2* __restrict__ lo, int pair_f16) {
    const long long total = rows * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(y) + i);
        if (pt_term) {
            const long long row = i / C4;
            const int c4 = (int)(i - row * C4);
            const float4 t = __ldg(reinterpret_cast<const float4*>(pt_term) + (row % rows_per_view) * C4 + c4);
            v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
        }
        v.x = vp_act(v.x, act), v.y = vp_act(v.y, act), v.z = vp_act(v.z, act), v.w = vp_act(v.w, act);
        uint2 h, l;
        holo_split2(v.x, v.y, pair_f16 != 0, h.x, l.x);
        holo_split2(v.z, v.w, pair_f16 != 0, h.y, l.y);
        hi[i] = h, lo[i] = l;
    }
}

// G[p] = sum_s softmax_s(Z[s][p][0]) Z[s][p][:]  (custom_modules.py:262-264); one warp per point
__global__ void __launch_bounds__(256) viewpool_reduce_kernel(const float* __restrict__ z, int n_views, long long rows_per_view,
                                                              long long n_pts, int C, float* __restrict__ out,
                                                              uint2* __restrict__ out_hi, uint2* __restrict__ out_lo,
                                                              int pair_f16) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int C4 = C >> 2;
    for (long long p = warp0; p < n_pts; p += n_warps) {
        float mx = -INFINITY;
        for (int s = lane; s < n_views; s += 32) mx = fmaxf(mx, __ldg(z + ((size_t)s * rows_per_view + p) * C));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float den = 0.f;
        for (int s = lane; s < n_views; s += 32) den += expf(__ldg(z + ((size_t)s * rows_per_view + p) * C) - mx);
        den = warp_sum(den);
        const float inv = 1.f / den;
        for (int c4 = lane; c4 < C4; c4 += 32) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < n_views; ++s) {
                const float* row = z + ((size_t)s * rows_per_view + p) * C;
                const float ws = expf(__ldg(row) - mx) * inv;
                const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c4);
                a.x += v.x * ws, a.y += v.y * ws, a.z += v.z * ws, a.w += v.w * ws;
            }
            if (out) reinterpret_cast<float4*>(out)[(size_t)p * C4 + c4] = a;
            if (out_hi) {
                uint2 h, l;
                holo_split2(a.x, a.y, pair_f16 != 0, h.x, l.x);
                holo_split2(a.z, a.w, pair_f16 != 0, h.y, l.y);
                out_hi[(size_t)p * C4 + c4] = h, out_lo[(size_t)p * C4 + c4] = l;
            }
        }
    }
}

}  // namespace

static int fill_params(const char* who, VpParams& P, const float* pts, long long n_pts, const float* R, const float* T,
                       const float* focal, const float* pp, int n_src, const holo_feature_map* maps, int n_maps,
                       const float* mask_map, int Hm, int Wm, const float* view_weight, float eps) {
    if (!(pts && R && T && focal && pp && maps)) {
        holo_set_error("%s: null arg", who);
        return HOLO_ERR_ARG;
    }
    if (!(n_pts > 0 && n_src > 0 && n_maps > 0 && n_maps <= VP_MAX_MAPS) || (mask_map && (Hm <= 0 || Wm <= 0))) {
        holo_set_error("%s: n_pts=%lld n_src=%d n_maps=%d (1..%d), mask map %dx%d", who, n_pts, n_src, n_maps, VP_MAX_MAPS, Hm, Wm);
        return HOLO_ERR_ARG;
    }
    P.pts = pts, P.n_pts = n_pts, P.R = R, P.T = T, P.focal = focal, P.pp = pp, P.n_src = n_src;
    int c0 = 0;
    for (int k = 0; k < n_maps; ++k) {
        if (!(maps[k].data && maps[k].channels > 0 && maps[k].height > 0 && maps[k].width > 0)) {
            holo_set_error("%s: feature map %d is empty", who, k);
            return HOLO_ERR_ARG;
        }
        P.maps[k].data = maps[k].data, P.maps[k].C = maps[k].channels, P.maps[k].H = maps[k].height;
        P.maps[k].W = maps[k].width, P.maps[k].c0 = c0;
        c0 += maps[k].channels;
    }
    for (int k = n_maps; k < VP_MAX_MAPS; ++k) P.maps[k] = P.maps[0];
    P.n_maps = n_maps, P.F = c0;
    P.mask_map = mask_map, P.Hm = Hm, P.Wm = Wm, P.view_weight = view_weight, P.eps = eps;
    P.n_harm = 0, P.E = 0, P.Kx = c0, P.Kpad = 0, P.rows_per_view = n_pts;
    P.x_hi = P.x_lo = P.m_hi = P.m_lo = nullptr, P.x_f32 = P.m_f32 = nullptr, P.pair_f16 = 0;
    return HOLO_OK;
}

extern "C" int holo_viewpool_sample(const float* pts, long long n_pts, const float* R, const float* T, const float* focal,
                                    const float* pp, int n_src, const holo_feature_map* maps, int n_maps,
                                    const float* mask_map, int Hm, int Wm, const float* view_weight, int n_harmonic,
                                    float eps, int Kpad, long long rows_per_view, void* x_hi, void* x_lo, void* mean_hi,
                                    void* mean_lo, float* x_f32, float* mean_f32, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(x_hi && x_lo && mean_hi && mean_lo, "holo_viewpool_sample: null output");
    HOLO_CHECK_ARG(n_harmonic >= 0 && n_harmonic <= 16, "holo_viewpool_sample: n_harmonic=%d", n_harmonic);
    HOLO_CHECK_ARG(rows_per_view >= n_pts, "holo_viewpool_sample: rows_per_view %lld < n_pts %lld", rows_per_view, n_pts);
    VpParams P;
    const int rc = fill_params("holo_viewpool_sample", P, pts, n_pts, R, T, focal, pp, n_src, maps, n_maps, mask_map, Hm, Wm,
                               view_weight, eps);
    if (rc) return rc;
    P.n_harm = n_harmonic, P.E = 3 * (2 * n_harmonic + 1), P.Kx = P.F + P.E, P.Kpad = Kpad;
    if (Kpad < P.Kx || Kpad > VP_MAX_K || Kpad % 8) {
        holo_set_error("holo_viewpool_sample: row of %d features + %d ray-embedding columns needs %d <= Kpad <= %d, "
                       "Kpad %% 8 == 0 (got %d)", P.F, P.E, P.Kx, VP_MAX_K, Kpad);
        return HOLO_ERR_UNSUPPORTED;
    }
    P.rows_per_view = rows_per_view;
    P.x_hi = (uint16_t*)x_hi, P.x_lo = (uint16_t*)x_lo, P.m_hi = (uint16_t*)mean_hi, P.m_lo = (uint16_t*)mean_lo;
    P.x_f32 = x_f32, P.m_f32 = mean_f32, P.pair_f16 = pair_f16;
    long long blocks = (n_pts + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    viewpool_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P);
    HOLO_CHECK_LAUNCH("holo_viewpool_sample");
    return HOLO_OK;
}

extern "C" int holo_viewpool_angle_reduce(const float* pts, long long n_pts, const float* R, const float* T,
                                          const float* focal, const float* pp, int n_src, const holo_feature_map* maps,
                                          int n_maps, const float* mask_map, int Hm, int Wm, const float* view_weight,
                                          float eps, float gamma, float min_ray_angle_weight, int with_std, int Kpad,
                                          void* out_hi, void* out_lo, float* out_f32, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(out_hi && out_lo, "holo_viewpool_angle_reduce: null output");
    VpParams P;
    const int rc = fill_params("holo_viewpool_angle_reduce", P, pts, n_pts, R, T, focal, pp, n_src, maps, n_maps, mask_map, Hm,
                               Wm, view_weight, eps);
    if (rc) return rc;
    const int cols = (with_std ? 2 : 1) * P.F;
    if (P.F > VP_MAX_K || Kpad < cols || Kpad % 8) {
        holo_set_error("holo_viewpool_angle_reduce: %d feature columns (<= %d), %d output columns need Kpad >= that, "
                       "Kpad %% 8 == 0 (got %d)", P.F, VP_MAX_K, cols, Kpad);
        return HOLO_ERR_UNSUPPORTED;
    }
    P.Kpad = Kpad, P.m_hi = (uint16_t*)out_hi, P.m_lo = (uint16_t*)out_lo, P.pair_f16 = pair_f16;
    long long blocks = (n_pts + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    viewpool_angle_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, gamma, min_ray_angle_weight, with_std ? 1 : 0,
                                                                              out_f32);
    HOLO_CHECK_LAUNCH("holo_viewpool_angle_reduce");
    return HOLO_OK;
}

extern "C" int holo_viewpool_act_split(const float* y, const float* point_term, int n_views, long long rows_per_view, int C,
                                       int act, void* hi, void* lo, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(y && hi && lo && n_views > 0 && rows_per_view > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 3,
                   "holo_viewpool_act_split: bad args (C=%d must be a multiple of 4, act=%d in 0..3)", C, act);
    const long long rows = (long long)n_views * rows_per_view, total = rows * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    viewpool_act_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, point_term, rows, rows_per_view, C / 4, act,
                                                                                  (uint2*)hi, (uint2*)lo, pair_f16);
    HOLO_CHECK_LAUNCH("holo_viewpool_act_split");
    return HOLO_OK;
}

extern "C" int holo_viewpool_reduce(const float* z, int n_views, long long rows_per_view, long long n_pts, int C, float* out,
                                    void* out_hi, void* out_lo, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(z && (out || out_hi) && n_views > 0 && n_pts > 0 && rows_per_view >= n_pts && C > 0 && C % 4 == 0,
                   "holo_viewpool_reduce: bad args (C=%d must be a multiple of 4)", C);
    HOLO_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr), "holo_viewpool_reduce: hi/lo outputs come together");
    long long blocks = (n_pts + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    viewpool_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, n_views, rows_per_view, n_pts, C, out,
                                                                               (uint2*)out_hi, (uint2*)out_lo, pair_f16);
    HOLO_CHECK_LAUNCH("holo_viewpool_reduce");
    return HOLO_OK;
}
