// Device functions shared by the fused SIMT renderer (render.cu) and the per-stage plug-in kernels
// (render_stages.cu): trilinear voxel sampling, the collapsed RenderMLP decode, torch.linspace.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// torch.linspace(0, 1, n)[i] in fp32 (ATen computes the upper half from the end point)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace01(int i, int n) {
    if (n == 1) return 0.0f;
    float step = 1.0f / (float)(n - 1);
    return (i < n / 2) ? step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

template <int C>
__device__ __forceinline__ void sample_trilinear(const float* __restrict__ grid, int D, int H, int W, float lx,
                                                 float ly, float lz, float (&f)[C]) {
    // ATen grid_sampler_3d, bilinear, zeros padding, align_corners=True
    float ix = ((lx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((ly + 1.f) / 2.f) * (float)(H - 1);
    float iz = ((lz + 1.f) / 2.f) * (float)(D - 1);
    float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
    // clamp before the int conversion so that far-away points cannot overflow (they are out of range anyway)
    int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)W + 1.f);
    int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)H + 1.f);
    int z0 = (int)fminf(fmaxf(fz0, -2.f), (float)D + 1.f);
    float x1f = fx0 + 1.f, y1f = fy0 + 1.f, z1f = fz0 + 1.f;
    float wx0 = x1f - ix, wx1 = ix - fx0;
    float wy0 = y1f - iy, wy1 = iy - fy0;
    float wz0 = z1f - iz, wz1 = iz - fz0;
#pragma unroll
    for (int c = 0; c < C; ++c) f[c] = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
        int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        float w = (dx ? wx1 : wx0) * (dy ? wy1 : wy0) * (dz ? wz1 : wz0);
        if (xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D) {
            const float4* p = reinterpret_cast<const float4*>(grid + (((size_t)zz * H + yy) * W + xx) * C);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                float4 v = __ldg(p + c4);
                f[c4 * 4 + 0] += v.x * w;
                f[c4 * 4 + 1] += v.y * w;
                f[c4 * 4 + 2] += v.z * w;
                f[c4 * 4 + 3] += v.w * w;
            }
        }
    }
}

// Decode NP points at once: sigma_raw and rgb (after sigmoid) from features.
template <int C, int NP>
__device__ __forceinline__ void decode_points(const float* __restrict__ sW, const float4* __restrict__ sEp,
                                              float b_sigma, int Hd, const float (&x)[NP][C], const float (&rd)[3],
                                              float (&sigma)[NP], float (&rgb)[NP][3]) {
    float r[NP][3];
#pragma unroll
    for (int p = 0; p < NP; ++p) r[p][0] = rd[0], r[p][1] = rd[1], r[p][2] = rd[2];
    for (int j = 0; j < Hd; ++j) {
        const float4* wrow = reinterpret_cast<const float4*>(sW + (size_t)j * C);
        float4 ep = sEp[j];
        float a0[NP], a1[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) a0[p] = ep.w, a1[p] = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
            float4 w = wrow[c4];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                a0[p] = fmaf(w.x, x[p][c4 * 4 + 0], a0[p]);
                a1[p] = fmaf(w.y, x[p][c4 * 4 + 1], a1[p]);
                a0[p] = fmaf(w.z, x[p][c4 * 4 + 2], a0[p]);
                a1[p] = fmaf(w.w, x[p][c4 * 4 + 3], a1[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float h = holo_leaky(a0[p] + a1[p]);
            r[p][0] = fmaf(ep.x, h, r[p][0]);
            r[p][1] = fmaf(ep.y, h, r[p][1]);
            r[p][2] = fmaf(ep.z, h, r[p][2]);
        }
    }
    {
        const float4* wrow = reinterpret_cast<const float4*>(sW + (size_t)Hd * C);
        float a0[NP], a1[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) a0[p] = b_sigma, a1[p] = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
            float4 w = wrow[c4];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                a0[p] = fmaf(w.x, x[p][c4 * 4 + 0], a0[p]);
                a1[p] = fmaf(w.y, x[p][c4 * 4 + 1], a1[p]);
                a0[p] = fmaf(w.z, x[p][c4 * 4 + 2], a0[p]);
                a1[p] = fmaf(w.w, x[p][c4 * 4 + 3], a1[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) sigma[p] = holo_leaky(a0[p] + a1[p]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int i = 0; i < 3; ++i) rgb[p][i] = 1.f / (1.f + expf(-holo_leaky(r[p][i])));
}


// Per-ray constant of the radiance pre-activation: br + Wr[:, H:] . HarmonicEmbedding(dn)
// (pytorch3d HarmonicEmbedding(n_harm, omega_0=1, logspace=True, append_input=True): [sin | cos | input],
// per-coordinate-major frequencies; RenderMLP.forward, holo_voxel_grid_implicit_function.py:117-120).
// sDir: [3][E] direction block of the radiance weight followed by br[3].
__device__ __forceinline__ void dir_radiance_const(const float* __restrict__ sDir, int E, int nh, const float (&dn)[3],
                                                   float (&rd)[3]) {
    const float* br = sDir + 3 * E;
    rd[0] = br[0], rd[1] = br[1], rd[2] = br[2];
    for (int c = 0; c < 3; ++c) {
        float freq = 1.f;
        for (int k = 0; k < nh; ++k) {
            float e = dn[c] * freq;
            float sn = sinf(e), cs = cosf(e);
            int ms = c * nh + k, mc = 3 * nh + c * nh + k;
#pragma unroll
            for (int i = 0; i < 3; ++i) rd[i] += sDir[i * E + ms] * sn + sDir[i * E + mc] * cs;
            freq *= 2.f;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) rd[i] += sDir[i * E + 6 * nh + c] * dn[c];
    }
}

// d(density pre-activation)/d(local coords) of a trilinearly sampled point, the analytic form of what
// RenderMLP.get_normals obtains with autograd (holo_voxel_grid_implicit_function.py:131-145): the density net is
// affine in the sampled feature (row `wsig` of the collapsed net), so
//   d pre / d ix = sum_corners (+-1) w_other_axes (wsig . v_corner),   in-bounds corners only
// (ATen grid_sampler_3d_backward, bilinear / zeros / align_corners=True: gix_mult = (W-1)/2).
// Returns pre = wsig . x (without the bias) and g = gradient w.r.t. the LOCAL coordinates (lx, ly, lz).
template <int C>
__device__ __forceinline__ void trilinear_density_grad(const float* __restrict__ grid, int D, int H, int W, float lx,
                                                       float ly, float lz, const float* __restrict__ wsig, float& pre,
                                                       float (&g)[3]) {
    float ix = ((lx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((ly + 1.f) / 2.f) * (float)(H - 1);
    float iz = ((lz + 1.f) / 2.f) * (float)(D - 1);
    float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
    int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)W + 1.f);
    int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)H + 1.f);
    int z0 = (int)fminf(fmaxf(fz0, -2.f), (float)D + 1.f);
    float wx0 = (fx0 + 1.f) - ix, wx1 = ix - fx0;
    float wy0 = (fy0 + 1.f) - iy, wy1 = iy - fy0;
    float wz0 = (fz0 + 1.f) - iz, wz1 = iz - fz0;
    pre = 0.f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
        int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        if (xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D) {
            const float4* p = reinterpret_cast<const float4*>(grid + (((size_t)zz * H + yy) * W + xx) * C);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                float4 v = __ldg(p + c4);
                float4 w = *reinterpret_cast<const float4*>(wsig + c4 * 4);
                d0 = fmaf(w.x, v.x, d0), d1 = fmaf(w.y, v.y, d1);
                d0 = fmaf(w.z, v.z, d0), d1 = fmaf(w.w, v.w, d1);
            }
            float dot = d0 + d1;
            float wx = dx ? wx1 : wx0, wy = dy ? wy1 : wy0, wz = dz ? wz1 : wz0;
            pre = fmaf(wx * wy * wz, dot, pre);
            gx = fmaf((dx ? 1.f : -1.f) * wy * wz, dot, gx);
            gy = fmaf((dy ? 1.f : -1.f) * wx * wz, dot, gy);
            gz = fmaf((dz ? 1.f : -1.f) * wx * wy, dot, gz);
        }
    }
    g[0] = gx * (0.5f * (float)(W - 1));
    g[1] = gy * (0.5f * (float)(H - 1));
    g[2] = gz * (0.5f * (float)(D - 1));
}
