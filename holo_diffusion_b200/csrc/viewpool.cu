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

__device__ __forceinline__ float nearest_1ch(const float* __restrict__ img, int H, int W, float ix, float iy) {
    if (!(ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H)) return 0.f;
    const int x = (int)nearbyintf(ix), y = (int)nearbyintf(iy);   // round half to even, as F.grid_sample(mode="nearest")
    if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
    return __ldg(img + (size_t)y * W + x);
}

// Cameras staged in shared memory once per block: R(9) T(3) focal(2) pp(2) centre(3) view weight(1)
constexpr int CAM_FLOATS = 20;

__device__ __forceinline__ void stage_cameras(const VpParams& P, float* cam) {
    for (int s = threadIdx.x; s < P.n_src; s += blockDim.x) {
        float* c = cam + s * CAM_FLOATS;
        const float* R = P.R + s * 9;
        for (int i = 0; i < 9; ++i) c[i] = R[i];
        const float t0 = P.T[s * 3], t1 = P.T[s * 3 + 1], t2 = P.T[s * 3 + 2];
        c[9] = t0, c[10] = t1, c[11] = t2;
        c[12] = P.focal[s * 2], c[13] = P.focal[s * 2 + 1], c[14] = P.pp[s * 2], c[15] = P.pp[s * 2 + 1];
        // centre = -T R^T (custom_modules.py:283-304: "does not produce nans randomly unlike get_camera_center()")
        c[16] = -(t0 * R[0] + t1 * R[1] + t2 * R[2]);
        c[17] = -(t0 * R[3] + t1 * R[4] + t2 * R[5]);
        c[18] = -(t0 * R[6] + t1 * R[7] + t2 * R[8]);
        c[19] = P.view_weight ? P.view_weight[s] : 1.f;
    }
    __syncthreads();
}

// Per (view, point): NDC projection, aggregation weight w = view_weight * sampled mask, unit vector centre -> point
__device__ __forceinline__ void view_geometry(const VpParams& P, const float* __restrict__ c, int s, float px, float py,
                                              float pz, float& x_ndc, float& y_ndc, float& vw, float& w, float& dx,
                                              float& dy, float& dz) {
    // X_cam = X_world R + T (row vectors)
    const float xc = px * c[0] + py * c[3] + pz * c[6] + c[9];
    const float yc = px * c[1] + py * c[4] + pz * c[7] + c[10];
    const float zc = px * c[2] + py * c[5] + pz * c[8] + c[11];
    // NDC projection with the sign-preserving clamp of the homogeneous divide (Transform3d.transform_points eps)
    const float den = (zc < 0.f ? -1.f : 1.f) * fmaxf(fabsf(zc), P.eps);
    const float rden = 1.f / den;
    x_ndc = (c[12] * xc + c[14] * zc) * rden;
    y_ndc = (c[13] * yc + c[15] * zc) * rden;
    vw = c[19];
    w = vw;
    if (P.mask_map) {
        float ix, iy;
        ndc_to_pixel(x_ndc, y_ndc, P.Hm, P.Wm, ix, iy);
        w *= nearest_1ch(P.mask_map + (size_t)s * P.Hm * P.Wm, P.Hm, P.Wm, ix, iy);
    }
    dx = px - c[16], dy = py - c[17], dz = pz - c[18];   // F.normalize, eps 1e-12
    const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    dx *= inv, dy *= inv, dz *= inv;
}

// Four consecutive columns of one feature map, owned by one lane for the whole kernel.  Every map's channel count is a
// multiple of 4 (the host pads with zero channels), so a group never straddles two maps and its taps are one float4.
struct ColGroup {
    const float* base;        // map data + first channel of the group (view 0)
    size_t view_stride;       // H * W * C floats
    int C, H, W;
    float ax, bx, ay, by;     // pixel = ndc * a + b  (negation, aspect scaling and the align_corners=False shift folded)
    int kind;                 // 0 feature, 1 ray embedding, 2 zero padding, 3 beyond the row
};

__device__ __forceinline__ ColGroup make_group(const VpParams& P, int j0, int n_cols) {
    ColGroup g;
    g.base = nullptr, g.view_stride = 0, g.C = g.H = g.W = 1, g.ax = g.bx = g.ay = g.by = 0.f;
    if (j0 >= n_cols) {
        g.kind = 3;
    } else if (j0 < P.F) {
        VpMap m = P.maps[0];
#pragma unroll
        for (int k = 1; k < VP_MAX_MAPS; ++k)   // static indices: the parameter struct stays in the constant bank
            if (k < P.n_maps && j0 >= P.maps[k].c0) m = P.maps[k];
        g.kind = 0, g.base = m.data + (j0 - m.c0), g.view_stride = (size_t)m.H * m.W * m.C, g.C = m.C, g.H = m.H, g.W = m.W;
        const float sx = m.H >= m.W ? 1.f : (float)m.H / (float)m.W, sy = m.H >= m.W ? (float)m.W / (float)m.H : 1.f;
        g.ax = -0.5f * sx * (float)m.W, g.bx = 0.5f * ((float)m.W - 1.f);
        g.ay = -0.5f * sy * (float)m.H, g.by = 0.5f * ((float)m.H - 1.f);
    } else {
        g.kind = j0 < P.Kx ? 1 : 2;
    }
    return g;
}

// bilinear, zeros padding: a tap outside the image contributes nothing
__device__ __forceinline__ float4 sample_group(const ColGroup& g, int s, float x_ndc, float y_ndc) {
    const float ix = fmaf(x_ndc, g.ax, g.bx), iy = fmaf(y_ndc, g.ay, g.by);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(ix > -1.f && ix < (float)g.W && iy > -1.f && iy < (float)g.H)) return v;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx1 = ix - fx0, wx0 = 1.f - wx1, wy1 = iy - fy0, wy0 = 1.f - wy1;
    const bool xa = x0 >= 0, xb = x0 + 1 < g.W, ya = y0 >= 0, yb = y0 + 1 < g.H;
    const float* p = g.base + (size_t)s * g.view_stride + ((ptrdiff_t)y0 * g.W + x0) * g.C;
    const int dxs = g.C, dys = g.W * g.C;
    if (ya && xa) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); const float k = wx0 * wy0; v.x += t.x * k, v.y += t.y * k, v.z += t.z * k, v.w += t.w * k; }
    if (ya && xb) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + dxs)); const float k = wx1 * wy0; v.x += t.x * k, v.y += t.y * k, v.z += t.z * k, v.w += t.w * k; }
    if (yb && xa) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + dys)); const float k = wx0 * wy1; v.x += t.x * k, v.y += t.y * k, v.z += t.z * k, v.w += t.w * k; }
    if (yb && xb) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + dys + dxs)); const float k = wx1 * wy1; v.x += t.x * k, v.y += t.y * k, v.z += t.z * k, v.w += t.w * k; }
    return v;
}

// HarmonicEmbedding(append_input): [sin(d_i 2^f)] (i major, f minor), [cos(...)], d.  Column e of the embedding is
// decoded ONCE per thread (EmbSlot); in the view loop lane e evaluates its own column and the lanes that own embedding
// column groups collect their four values with shuffles -- one sinf/cosf per lane instead of four on six lanes.
struct EmbSlot {
    int comp;      // 0..2: component of the direction; -1: no column
    int fn;        // 0 sin, 1 cos, 2 identity
    float scale;   // 2^f
};

__device__ __forceinline__ EmbSlot make_emb_slot(int e, int n_harm, int E) {
    EmbSlot t;
    t.comp = -1, t.fn = 2, t.scale = 1.f;
    if (e >= E) return t;
    const int nh3 = 3 * n_harm;
    if (e < 2 * nh3) {
        const int q = e < nh3 ? e : e - nh3;
        t.comp = q / n_harm;
        t.scale = (float)(1 << (q - t.comp * n_harm));
        t.fn = e < nh3 ? 0 : 1;
    } else {
        t.comp = e - 2 * nh3;
    }
    return t;
}

__device__ __forceinline__ float eval_emb(const EmbSlot& t, float dx, float dy, float dz) {
    if (t.comp < 0) return 0.f;
    const float a = (t.comp == 0 ? dx : (t.comp == 1 ? dy : dz)) * t.scale;
    return t.fn == 0 ? sinf(a) : (t.fn == 1 ? cosf(a) : a);
}

__device__ __forceinline__ void store_pair4(uint16_t* hi, uint16_t* lo, size_t at, float4 v, bool f16) {
    uint2 h, l;
    holo_split2(v.x, v.y, f16, h.x, l.x);
    holo_split2(v.z, v.w, f16, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + at) = h;
    *reinterpret_cast<uint2*>(lo + at) = l;
}

constexpr int VP_GROUPS = VP_MAX_K / 128;   // 4-column groups per lane (template parameter of the kernels: 1 for rows <= 128)

// One warp per point; every lane owns up to two 4-column groups of the row; the view loop is inside (the mean needs
// all views).  Rows: [features of map 0 | map 1 | ... | sin | cos | dir | zero padding to Kpad].
template <int VP_GROUPS>
__global__ void __launch_bounds__(256) viewpool_sample_kernel(const VpParams P) {
    extern __shared__ float cam[];
    stage_cameras(P, cam);
    const int lane = threadIdx.x & 31;
    const bool f16 = P.pair_f16 != 0;
    ColGroup grp[VP_GROUPS];
#pragma unroll
    for (int g = 0; g < VP_GROUPS; ++g) grp[g] = make_group(P, (g * 32 + lane) * 4, P.Kpad);
    const EmbSlot slot0 = make_emb_slot(lane, P.n_harm, P.E), slot1 = make_emb_slot(32 + lane, P.n_harm, P.E);
    const bool two_slots = P.E > 32;   // E = 3 (2 n + 1) <= 64 (host check)
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp0; p < P.n_pts; p += n_warps) {
        const float px = __ldg(P.pts + p * 3), py = __ldg(P.pts + p * 3 + 1), pz = __ldg(P.pts + p * 3 + 2);
        float4 acc[VP_GROUPS];
#pragma unroll
        for (int g = 0; g < VP_GROUPS; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        float wsum = 0.f;
        for (int s = 0; s < P.n_src; ++s) {
            float x_ndc, y_ndc, vw, w, dx, dy, dz;
            view_geometry(P, cam + s * CAM_FLOATS, s, px, py, pz, x_ndc, y_ndc, vw, w, dx, dy, dz);
            const size_t row = ((size_t)s * P.rows_per_view + (size_t)p) * P.Kpad;
            const float ev0 = eval_emb(slot0, dx, dy, dz), ev1 = two_slots ? eval_emb(slot1, dx, dy, dz) : 0.f;
#pragma unroll
            for (int g = 0; g < VP_GROUPS; ++g) {
                const int j0 = (g * 32 + lane) * 4;
                // every lane takes part in the shuffles; lanes without an embedding group read column 0
                const int e = grp[g].kind == 1 ? j0 - P.F : 0;
                float em[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ei = e + i;
                    const float a = __shfl_sync(0xffffffffu, ev0, ei & 31);
                    const float b = two_slots ? __shfl_sync(0xffffffffu, ev1, ei & 31) : 0.f;
                    em[i] = ei < P.E ? (ei < 32 ? a : b) : 0.f;
                }
                if (grp[g].kind == 3) continue;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grp[g].kind == 0) {
                    v = sample_group(grp[g], s, x_ndc, y_ndc);
                    v.x *= vw, v.y *= vw, v.z *= vw, v.w *= vw;
                } else if (grp[g].kind == 1) {
                    v = make_float4(em[0], em[1], em[2], em[3]);
                }
                v.x *= w, v.y *= w, v.z *= w, v.w *= w;
                acc[g].x += v.x * w, acc[g].y += v.y * w, acc[g].z += v.z * w, acc[g].w += v.w * w;
                store_pair4(P.x_hi, P.x_lo, row + j0, v, f16);
                if (P.x_f32) {
                    float* o = P.x_f32 + ((size_t)s * P.n_pts + (size_t)p) * P.Kx;
                    if (j0 < P.Kx) o[j0] = v.x;
                    if (j0 + 1 < P.Kx) o[j0 + 1] = v.y;
                    if (j0 + 2 < P.Kx) o[j0 + 2] = v.z;
                    if (j0 + 3 < P.Kx) o[j0 + 3] = v.w;
                }
            }
            wsum += w;
        }
        const float invw = 1.f / fmaxf(wsum, 1e-2f);   // wmean(..., eps=1e-2)
#pragma unroll
        for (int g = 0; g < VP_GROUPS; ++g) {
            if (grp[g].kind == 3) continue;
            const int j0 = (g * 32 + lane) * 4;
            const float4 m = make_float4(acc[g].x * invw, acc[g].y * invw, acc[g].z * invw, acc[g].w * invw);
            store_pair4(P.m_hi, P.m_lo, (size_t)p * P.Kpad + j0, m, f16);
            if (P.m_f32) {
                float* o = P.m_f32 + (size_t)p * P.Kx;
                if (j0 < P.Kx) o[j0] = m.x;
                if (j0 + 1 < P.Kx) o[j0 + 1] = m.y;
                if (j0 + 2 < P.Kx) o[j0 + 2] = m.z;
                if (j0 + 3 < P.Kx) o[j0 + 3] = m.w;
            }
        }
    }
}

// AngleWeightedReductionFeatureAggregator (pytorch3d 0.7.4, restated from memory; reduction_functions = AVG, STD):
//   w[s] = mask[s] * ((0.5 (d_s . d_0 + 1))^gamma + min_weight),  d_s = unit vector camera s -> point (camera 0 = the
//   first camera of the point batch), mu = sum w x / max(sum w, 1e-2), std = sqrt(max(sum w (x - mu)^2 / max(sum w, 1e-2), 1e-4));
//   row layout [mu_k | std_k] per feature map k.  Two passes over the views (the second re-gathers: the maps are L2 resident).
template <int VP_GROUPS>
__global__ void __launch_bounds__(256) viewpool_angle_kernel(const VpParams P, float gamma, float min_weight, int with_std,
                                                             float* __restrict__ out_f32) {
    extern __shared__ float cam[];
    stage_cameras(P, cam);
    const int lane = threadIdx.x & 31;
    const bool f16 = P.pair_f16 != 0;
    const int per = with_std ? 2 : 1;
    ColGroup grp[VP_GROUPS];
    int col_mu[VP_GROUPS];
#pragma unroll
    for (int g = 0; g < VP_GROUPS; ++g) {
        const int j0 = (g * 32 + lane) * 4;
        grp[g] = make_group(P, j0, P.F);
        VpMap m = P.maps[0];
#pragma unroll
        for (int k = 1; k < VP_MAX_MAPS; ++k)
            if (k < P.n_maps && j0 >= P.maps[k].c0) m = P.maps[k];
        col_mu[g] = per * m.c0 + (j0 - m.c0);
    }
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp0; p < P.n_pts; p += n_warps) {
        const float px = __ldg(P.pts + p * 3), py = __ldg(P.pts + p * 3 + 1), pz = __ldg(P.pts + p * 3 + 2);
        float4 mu[VP_GROUPS], var[VP_GROUPS];
#pragma unroll
        for (int g = 0; g < VP_GROUPS; ++g) mu[g] = var[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        float wsum = 0.f, d0x = 0.f, d0y = 0.f, d0z = 0.f;
        for (int pass = 0; pass < per; ++pass) {
            for (int s = 0; s < P.n_src; ++s) {
                float x_ndc, y_ndc, vw, w, dx, dy, dz;
                view_geometry(P, cam + s * CAM_FLOATS, s, px, py, pz, x_ndc, y_ndc, vw, w, dx, dy, dz);
                if (s == 0) d0x = dx, d0y = dy, d0z = dz;
                const float a01 = 0.5f * (dx * d0x + dy * d0y + dz * d0z + 1.f);
                w *= (gamma == 1.f ? a01 : powf(a01, gamma)) + min_weight;
                if (pass == 0) wsum += w;
#pragma unroll
                for (int g = 0; g < VP_GROUPS; ++g) {
                    if (grp[g].kind != 0) continue;
                    float4 v = sample_group(grp[g], s, x_ndc, y_ndc);
                    v.x *= vw, v.y *= vw, v.z *= vw, v.w *= vw;
                    if (pass == 0) {
                        mu[g].x += v.x * w, mu[g].y += v.y * w, mu[g].z += v.z * w, mu[g].w += v.w * w;
                    } else {
                        const float ex = v.x - mu[g].x, ey = v.y - mu[g].y, ez = v.z - mu[g].z, ew = v.w - mu[g].w;
                        var[g].x += ex * ex * w, var[g].y += ey * ey * w, var[g].z += ez * ez * w, var[g].w += ew * ew * w;
                    }
                }
            }
            if (pass == 0) {
                const float invw = 1.f / fmaxf(wsum, 1e-2f);
#pragma unroll
                for (int g = 0; g < VP_GROUPS; ++g) mu[g].x *= invw, mu[g].y *= invw, mu[g].z *= invw, mu[g].w *= invw;
            }
        }
        const float invw = 1.f / fmaxf(wsum, 1e-2f);
        // zero the padding columns, then scatter [mu_k | std_k]
        for (int j = per * P.F + lane; j < P.Kpad; j += 32) P.m_hi[(size_t)p * P.Kpad + j] = 0, P.m_lo[(size_t)p * P.Kpad + j] = 0;
#pragma unroll
        for (int g = 0; g < VP_GROUPS; ++g) {
            if (grp[g].kind != 0) continue;
            store_pair4(P.m_hi, P.m_lo, (size_t)p * P.Kpad + col_mu[g], mu[g], f16);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)p * per * P.F + col_mu[g]) = mu[g];
            if (with_std) {
                const float4 sd = make_float4(sqrtf(fmaxf(var[g].x * invw, 1e-4f)), sqrtf(fmaxf(var[g].y * invw, 1e-4f)),
                                              sqrtf(fmaxf(var[g].z * invw, 1e-4f)), sqrtf(fmaxf(var[g].w * invw, 1e-4f)));
                store_pair4(P.m_hi, P.m_lo, (size_t)p * P.Kpad + col_mu[g] + grp[g].C, sd, f16);
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)p * per * P.F + col_mu[g] + grp[g].C) = sd;
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
    if (!(n_pts > 0 && n_src > 0 && n_src <= 512 && n_maps > 0 && n_maps <= VP_MAX_MAPS) || (mask_map && (Hm <= 0 || Wm <= 0))) {
        holo_set_error("%s: n_pts=%lld n_src=%d (1..512) n_maps=%d (1..%d), mask map %dx%d", who, n_pts, n_src, n_maps, VP_MAX_MAPS, Hm, Wm);
        return HOLO_ERR_ARG;
    }
    P.pts = pts, P.n_pts = n_pts, P.R = R, P.T = T, P.focal = focal, P.pp = pp, P.n_src = n_src;
    int c0 = 0;
    for (int k = 0; k < n_maps; ++k) {
        if (!(maps[k].data && maps[k].channels > 0 && maps[k].height > 0 && maps[k].width > 0)) {
            holo_set_error("%s: feature map %d is empty", who, k);
            return HOLO_ERR_ARG;
        }
        if (maps[k].channels % 4 || ((uintptr_t)maps[k].data & 15)) {
            holo_set_error("%s: feature map %d has %d channels: pad to a multiple of 4 (zero channels) and align to 16 "
                           "bytes -- a lane reads the 4 channels of its column group as one float4", who, k, maps[k].channels);
            return HOLO_ERR_UNSUPPORTED;
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
    HOLO_CHECK_ARG(n_harmonic >= 0 && n_harmonic <= 10, "holo_viewpool_sample: n_harmonic=%d (0..10: 3 (2 n + 1) <= 64 columns)",
                   n_harmonic);
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
    const size_t smem = (size_t)n_src * CAM_FLOATS * sizeof(float);
    if (Kpad <= 128) viewpool_sample_kernel<1><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(P);
    else viewpool_sample_kernel<2><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(P);
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
    const size_t smem = (size_t)n_src * CAM_FLOATS * sizeof(float);
    if (P.F <= 128)
        viewpool_angle_kernel<1><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(P, gamma, min_ray_angle_weight,
                                                                                         with_std ? 1 : 0, out_f32);
    else
        viewpool_angle_kernel<2><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(P, gamma, min_ray_angle_weight,
                                                                                         with_std ? 1 : 0, out_f32);
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
