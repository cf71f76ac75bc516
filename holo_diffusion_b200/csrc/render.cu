// Fused volumetric renderer for sm_100a: ray generation, density-net collapse, and the one-kernel
// trilinear-sample -> RenderMLP -> emission-absorption compositing -> importance refinement -> 2nd pass.
//
// Reference path replaced (citations relative to /root/reference):
//   holo_diffusion/holo_voxel_grid_implicit_function.py:182-269  (HoloVoxelGridImplicitFunction.forward)
//   holo_diffusion/holo_voxel_grid_implicit_function.py:107-129  (RenderMLP.forward)
//   holo_diffusion/custom_modules.py:91-113,133-160              (MLPWithInputSkips)
//   holo_diffusion/holo_multipass_ea.py:79-125                   (_run_raymarcher, recursion over passes)
//   pytorch3d 0.7.4 EmissionAbsorptionRaymarcher / RayPointRefiner / sample_pdf / NDCMultinomialRaysampler
//   (un-vendored; arithmetic per SURVEY.md Appendix A).
#include "common.cuh"
#include "render_device.cuh"
#include "../../include/holo_b200.h"

// ------------------------------------------------------------------------------------------------
// Ray generation (NDCMultinomialRaysampler + AdaptiveRaySampler depth bounds), eval / full-grid mode.
//   configs/base.yaml:129-140; invoked holo_diffusion_model.py:442-448.
// ------------------------------------------------------------------------------------------------
__global__ void raygen_kernel(const float* __restrict__ R, const float* __restrict__ T,
                              const float* __restrict__ focal, const float* __restrict__ pp,
                              const float* __restrict__ xy, int n_cam, int n_rays, int S, float scene_extent,
                              float cx, float cy, float cz, float* __restrict__ origins,
                              float* __restrict__ dirs, float* __restrict__ lengths) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_cam * n_rays;
    if (gid < total) {
        int cam = (int)(gid / n_rays);
        int r = (int)(gid % n_rays);
        const float* Rc = R + cam * 9;
        const float* Tc = T + cam * 3;
        float fx = focal[cam * 2], fy = focal[cam * 2 + 1], px = pp[cam * 2], py = pp[cam * 2 + 1];
        float x = xy[r * 2], y = xy[r * 2 + 1];
        float w1[3], w2[3];
#pragma unroll
        for (int zi = 0; zi < 2; ++zi) {
            float z = (float)(zi + 1);
            float c0 = (x - px) * z / fx - Tc[0];
            float c1 = (y - py) * z / fy - Tc[1];
            float c2 = z - Tc[2];
            float* w = zi ? w2 : w1;
            // X_world = (X_cam - T) R^T
#pragma unroll
            for (int i = 0; i < 3; ++i) w[i] = c0 * Rc[i * 3 + 0] + c1 * Rc[i * 3 + 1] + c2 * Rc[i * 3 + 2];
        }
        float d[3] = {w2[0] - w1[0], w2[1] - w1[1], w2[2] - w1[2]};
        float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            origins[gid * 3 + i] = w1[i] - d[i];
            dirs[gid * 3 + i] = d[i] / nrm;
        }
    }
    // lengths: coalesced over (ray, s)
    long long tot_l = total * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot_l;
         i += (long long)gridDim.x * blockDim.x) {
        int cam = (int)(i / ((long long)n_rays * S));
        int s = (int)(i % S);
        const float* Rc = R + cam * 9;
        const float* Tc = T + cam * 3;
        float C[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) C[k] = -(Tc[0] * Rc[k * 3 + 0] + Tc[1] * Rc[k * 3 + 1] + Tc[2] * Rc[k * 3 + 2]);
        float d2 = (C[0] - cx) * (C[0] - cx) + (C[1] - cy) * (C[1] - cy) + (C[2] - cz) * (C[2] - cz);
        float cd = fmaxf(sqrtf(fmaxf(d2, 0.001f)), 0.001f);
        cd = fmaxf(cd, scene_extent + 1e-3f);
        float mn = cd - scene_extent, mx = cd + scene_extent;
        lengths[i] = mn + linspace01(s, S) * (mx - mn);
    }
}

extern "C" int holo_raygen(const float* R, const float* T, const float* focal, const float* pp, const float* xy,
                           int n_cam, int n_rays, int S, float scene_extent, const float* scene_center3_host,
                           float* origins, float* dirs, float* lengths, void* stream) {
    HOLO_CHECK_ARG(n_cam > 0 && n_rays > 0 && S > 0, "holo_raygen: bad sizes");
    long long total = (long long)n_cam * n_rays * S;
    int threads = 256;
    int blocks = holo_cdiv((long long)n_cam * n_rays, threads);
    int want = holo_cdiv(total, threads * 4);
    if (want > blocks) blocks = want;
    float cx = scene_center3_host ? scene_center3_host[0] : 0.f;
    float cy = scene_center3_host ? scene_center3_host[1] : 0.f;
    float cz = scene_center3_host ? scene_center3_host[2] : 0.f;
    raygen_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(R, T, focal, pp, xy, n_cam, n_rays, S, scene_extent,
                                                                cx, cy, cz, origins, dirs, lengths);
    HOLO_CHECK_LAUNCH("holo_raygen");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Density-net collapse (SURVEY.md section 7 hard part 3): MLPWithInputSkips gives layers 0..n-2 the Identity
// activation (custom_modules.py:108-112), so the net is affine up to the final LeakyReLU.  We carry
// y = A x + c in fp64: one call composes a Linear (optionally preceded by the skip concat cat((y, x))).
//   A_out = W[:, :rows] A_in + W[:, rows:rows+C]   (second term only when skip)
//   c_out = W[:, :rows] c_in + b
// A_in == nullptr means "identity" (first layer): A_out = W, c_out = b.
// ------------------------------------------------------------------------------------------------
__global__ void affine_compose_kernel(const float* __restrict__ W, const float* __restrict__ b, int out_dim,
                                      int in_total, const double* __restrict__ A_in,
                                      const double* __restrict__ c_in, int rows, int C, int skip,
                                      double* __restrict__ A_out, double* __restrict__ c_out) {
    int o = blockIdx.x;
    for (int col = threadIdx.x; col <= C; col += blockDim.x) {
        double acc;
        if (col < C) {
            acc = 0.0;
            if (A_in == nullptr) {
                acc = (double)W[(size_t)o * in_total + col];
            } else {
                for (int k = 0; k < rows; ++k) acc += (double)W[(size_t)o * in_total + k] * A_in[(size_t)k * C + col];
                if (skip) acc += (double)W[(size_t)o * in_total + rows + col];
            }
            A_out[(size_t)o * C + col] = acc;
        } else {
            acc = (double)b[o];
            if (A_in != nullptr)
                for (int k = 0; k < rows; ++k) acc += (double)W[(size_t)o * in_total + k] * c_in[k];
            c_out[o] = acc;
        }
    }
}

extern "C" int holo_affine_compose_f64(const float* W, const float* b, int out_dim, int in_total, const double* A_in,
                                       const double* c_in, int rows, int C, int skip, double* A_out, double* c_out,
                                       void* stream) {
    HOLO_CHECK_ARG(out_dim > 0 && C > 0, "holo_affine_compose_f64: bad sizes");
    if (A_in == nullptr)
        HOLO_CHECK_ARG(in_total == C, "holo_affine_compose_f64: first layer must consume C inputs");
    else
        HOLO_CHECK_ARG(in_total == rows + (skip ? C : 0), "holo_affine_compose_f64: in_total mismatch");
    affine_compose_kernel<<<out_dim, 128, 0, (cudaStream_t)stream>>>(W, b, out_dim, in_total, A_in, c_in, rows, C,
                                                                    skip, A_out, c_out);
    HOLO_CHECK_LAUNCH("holo_affine_compose_f64");
    return HOLO_OK;
}

// Pack the collapsed density net + radiance layer into the layout the render kernel stages in shared memory:
//   [0, (H+1)*C)            W_eff rows (row j = hidden unit j, row H = density)
//   [.., +4*H)              per hidden unit float4 (wr[0][j], wr[1][j], wr[2][j], b_eff[j])
//   [.., +4)                (b_eff[H], 0, 0, 0)
//   [.., +3*E+3 -> pad 4)   radiance direction block: Wr[i][H + m] (i-major, m < E) then br[3]
__global__ void pack_render_mlp_kernel(const double* __restrict__ A, const double* __restrict__ c,
                                       const float* __restrict__ Wr, const float* __restrict__ br, int H, int C,
                                       int E, float* __restrict__ out) {
    int n_w = (H + 1) * C;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = gridDim.x * blockDim.x;
    for (int i = tid; i < n_w; i += stride) out[i] = (float)A[i];
    float* ep = out + n_w;
    for (int j = tid; j < H; j += stride) {
        ep[j * 4 + 0] = Wr[0 * (H + E) + j];
        ep[j * 4 + 1] = Wr[1 * (H + E) + j];
        ep[j * 4 + 2] = Wr[2 * (H + E) + j];
        ep[j * 4 + 3] = (float)c[j];
    }
    float* sg = ep + 4 * H;
    if (tid == 0) {
        sg[0] = (float)c[H];
        sg[1] = sg[2] = sg[3] = 0.f;
    }
    float* dr = sg + 4;
    for (int i = tid; i < 3 * E; i += stride) dr[i] = Wr[(i / E) * (H + E) + H + (i % E)];
    if (tid < 3) dr[3 * E + tid] = br[tid];
}

extern "C" long long holo_render_mlp_packed_floats(int H, int C, int E) {
    long long n = (long long)(H + 1) * C + 4LL * H + 4 + 3LL * E + 3;
    return (n + 3) / 4 * 4;
}

extern "C" int holo_pack_render_mlp(const double* A_eff, const double* c_eff, const float* Wr, const float* br, int H,
                                    int C, int E, float* packed, void* stream) {
    HOLO_CHECK_ARG(H > 0 && C > 0 && C % 4 == 0 && E >= 3, "holo_pack_render_mlp: bad sizes (C must be a multiple of 4)");
    pack_render_mlp_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(A_eff, c_eff, Wr, br, H, C, E, packed);
    HOLO_CHECK_LAUNCH("holo_pack_render_mlp");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// The fused render kernel (SIMT fp32 version).  One thread per ray, RT rays per CTA.
// ------------------------------------------------------------------------------------------------
struct RenderParams {
    const float* grid;  // channels-last (D,H,W,C)
    int D, Hh, Ww;
    float inv_scale;    // 1 / ((R-1) * voxel_size / 2), same for the three axes (cubic voxels)
    float inv_scale_y, inv_scale_z;
    const float* mlp;   // packed, see pack_render_mlp_kernel
    int Hd;             // hidden width (256)
    int n_harm;         // harmonic functions of the direction embedding (4)
    const float* origins;
    const float* dirs;
    const float* lengths;  // (n_rays, S)
    int n_rays, S, n_fine, add_input;
    float bg[3];
    float bg_opacity;
    // outputs of the LAST pass
    float* features;  // (n,3)
    float* depths;    // (n)
    float* masks;     // (n)
    float* weights;   // (n,S_last) or null
    float* lengths_out;  // (n,S_last) or null (only meaningful for 2 passes)
    // outputs of the previous stage (only for 2 passes; may be null)
    float* p_features;
    float* p_depths;
    float* p_masks;
    float* p_weights;  // (n,S) or null
    int n_passes;
};

template <int C, int NP, int RT>
__global__ void __launch_bounds__(RT) render_fused_kernel(RenderParams P) {
    extern __shared__ __align__(16) float smem[];
    const int Hd = P.Hd;
    const int E = 3 * (2 * P.n_harm + 1);
    float* sW = smem;                                        // (Hd+1)*C
    float4* sEp = reinterpret_cast<float4*>(sW + (Hd + 1) * C);  // Hd float4
    float* sSig = reinterpret_cast<float*>(sEp + Hd);        // 4
    float* sDir = sSig + 4;                                  // 3E+3 (padded)
    int n_pack = (Hd + 1) * C + 4 * Hd + 4 + 3 * E + 3;
    n_pack = (n_pack + 3) / 4 * 4;
    float* sRay = smem + n_pack;  // per-ray column arrays, [slot][RT]
    for (int i = threadIdx.x; i < n_pack; i += RT) smem[i] = P.mlp[i];
    __syncthreads();

    const int ray = blockIdx.x * RT + threadIdx.x;
    if (ray >= P.n_rays) return;
    const int tid = threadIdx.x;
    const float b_sigma = sSig[0];

    float o[3], d[3], dn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = P.origins[(size_t)ray * 3 + i], d[i] = P.dirs[(size_t)ray * 3 + i];
    {
        // F.normalize(directions, dim=-1) (holo_voxel_grid_implicit_function.py:239)
        float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
#pragma unroll
        for (int i = 0; i < 3; ++i) dn[i] = d[i] / nrm;
    }
    // per-ray part of the radiance layer: br + Wr[:, H:] . PE(dn)
    float rd[3];
    dir_radiance_const(sDir, E, P.n_harm, dn, rd);

    const int S1 = P.S;
    const int S2 = P.add_input ? P.S + P.n_fine : P.n_fine;
    const float* zin = P.lengths + (size_t)ray * S1;

    for (int pass = 0; pass < P.n_passes; ++pass) {
        const int S = pass == 0 ? S1 : S2;
        const bool last = pass == P.n_passes - 1;
        // z accessor: pass 0 from global, pass 1 from the per-ray shared column
        auto Z = [&](int s) -> float { return pass == 0 ? zin[s] : sRay[(size_t)s * RT + tid]; };
        float cum = 0.f, opac_prev = 0.f, af[3] = {0.f, 0.f, 0.f}, ad = 0.f;
        float* wout = last ? P.weights : P.p_weights;
        for (int s0 = 0; s0 < S; s0 += NP) {
            float x[NP][C], zs[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                int s = min(s0 + p, S - 1);
                zs[p] = Z(s);
                float px = o[0] + zs[p] * d[0], py = o[1] + zs[p] * d[1], pz = o[2] + zs[p] * d[2];
                sample_trilinear<C>(P.grid, P.D, P.Hh, P.Ww, px * P.inv_scale, py * P.inv_scale_y,
                                    pz * P.inv_scale_z, x[p]);
            }
            float sg[NP], rgb[NP][3];
            decode_points<C, NP>(sW, sEp, b_sigma, Hd, x, rd, sg, rgb);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                int s = s0 + p;
                if (s < S) {
                    float delta = (s + 1 < S) ? (Z(s + 1) - zs[p]) : P.bg_opacity;
                    float wd = delta * fmaxf(sg[p], 0.f);
                    float capped = 1.f - expf(-wd);
                    cum += wd;
                    float opac = 1.f - expf(-cum);
                    float absorb = (s == 0) ? 1.f : (1.f - opac_prev);
                    float w = capped * absorb;
                    af[0] += w * rgb[p][0];
                    af[1] += w * rgb[p][1];
                    af[2] += w * rgb[p][2];
                    ad += w * zs[p];
                    opac_prev = opac;
                    if (wout) wout[(size_t)ray * S + s] = w;
                    if (!last) sRay[(size_t)(S2 + s) * RT + tid] = w;  // keep for the refiner
                }
            }
        }
        {
            float mask = opac_prev;
            float* F = last ? P.features : P.p_features;
            float* Dp = last ? P.depths : P.p_depths;
            float* M = last ? P.masks : P.p_masks;
            if (F) {
                F[(size_t)ray * 3 + 0] = af[0] + (1.f - mask) * P.bg[0];
                F[(size_t)ray * 3 + 1] = af[1] + (1.f - mask) * P.bg[1];
                F[(size_t)ray * 3 + 2] = af[2] + (1.f - mask) * P.bg[2];
            }
            if (Dp) Dp[ray] = ad;
            if (M) M[ray] = mask;
        }
        if (!last) {
            // ---- RayPointRefiner + sample_pdf (deterministic u), SURVEY.md A8 ----
            float* wc = sRay + (size_t)S2 * RT + tid;  // column of S1 weights -> becomes the cdf
            const float eps = 1e-5f;
            float tot = 0.f;
            for (int i = 1; i <= S1 - 2; ++i) tot += wc[(size_t)i * RT] + eps;
            float run = 0.f;
            wc[0] = 0.f;
            for (int i = 1; i <= S1 - 2; ++i) {
                run += (wc[(size_t)i * RT] + eps) / tot;
                wc[(size_t)i * RT] = run;  // cdf[i], i in [0, S1-2]
            }
            const int ncdf = S1 - 1;
            float* zc = sRay + tid;  // merged lengths column, S2 entries
            int base = 0;
            if (P.add_input) {
                for (int s = 0; s < S1; ++s) zc[(size_t)s * RT] = zin[s];
                base = S1;
            }
            int inds = 0;
            for (int j = 0; j < P.n_fine; ++j) {
                float u = linspace01(j, P.n_fine);
                while (inds < ncdf && wc[(size_t)inds * RT] <= u) ++inds;  // searchsorted(right=True)
                int below = max(inds - 1, 0), above = min(inds, ncdf - 1);
                float c0 = wc[(size_t)below * RT], c1 = wc[(size_t)above * RT];
                float b0 = 0.5f * (zin[below + 1] + zin[below]);
                float b1 = 0.5f * (zin[above + 1] + zin[above]);
                float den = c1 - c0;
                if (den < eps) den = 1.f;
                float t = (u - c0) / den;
                zc[(size_t)(base + j) * RT] = b0 + t * (b1 - b0);
            }
            // torch.sort(cat(z, z_new)): insertion sort (input is two nearly sorted runs)
            for (int i = 1; i < S2; ++i) {
                float v = zc[(size_t)i * RT];
                int k = i - 1;
                while (k >= 0 && zc[(size_t)k * RT] > v) {
                    zc[(size_t)(k + 1) * RT] = zc[(size_t)k * RT];
                    --k;
                }
                zc[(size_t)(k + 1) * RT] = v;
            }
            if (P.lengths_out)
                for (int s = 0; s < S2; ++s) P.lengths_out[(size_t)ray * S2 + s] = zc[(size_t)s * RT];
        }
    }
}

template <int C, int NP>
static int launch_render(const RenderParams& P, cudaStream_t st) {
    const int E = 3 * (2 * P.n_harm + 1);
    size_t n_pack = (size_t)(P.Hd + 1) * C + 4 * P.Hd + 4 + 3 * E + 3;
    n_pack = (n_pack + 3) / 4 * 4;
    int S2 = P.add_input ? P.S + P.n_fine : P.n_fine;
    size_t per_ray = P.n_passes > 1 ? (size_t)(S2 + P.S) : 0;
    size_t sm256 = (n_pack + per_ray * 256) * sizeof(float);
    size_t sm128 = (n_pack + per_ray * 128) * sizeof(float);
    size_t sm64 = (n_pack + per_ray * 64) * sizeof(float);
    const size_t lim = 227 * 1024;
    if (sm256 <= lim) {
        auto k = render_fused_kernel<C, NP, 256>;
        HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm256), "holo_render_fwd");
        k<<<holo_cdiv(P.n_rays, 256), 256, sm256, st>>>(P);
    } else if (sm128 <= lim) {
        auto k = render_fused_kernel<C, NP, 128>;
        HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm128), "holo_render_fwd");
        k<<<holo_cdiv(P.n_rays, 128), 128, sm128, st>>>(P);
    } else if (sm64 <= lim) {
        auto k = render_fused_kernel<C, NP, 64>;
        HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm64), "holo_render_fwd");
        k<<<holo_cdiv(P.n_rays, 64), 64, sm64, st>>>(P);
    } else {
        holo_set_error("holo_render_fwd: S=%d n_fine=%d needs too much shared memory", P.S, P.n_fine);
        return HOLO_ERR_UNSUPPORTED;
    }
    HOLO_CHECK_LAUNCH("holo_render_fwd");
    return HOLO_OK;
}

extern "C" int holo_render_fwd(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent,
                               const float* packed_mlp, int hidden, int n_harmonic, const float* origins,
                               const float* dirs, const float* lengths, int n_rays, int S, int n_passes, int n_fine,
                               int add_input_samples, const float* bg3_host, float background_opacity,
                               float* features, float* depths, float* masks, float* weights, float* lengths_out,
                               float* prev_features, float* prev_depths, float* prev_masks, float* prev_weights,
                               void* stream) {
    if (n_rays == 0) return HOLO_OK;
    HOLO_CHECK_ARG(grid_dhwc && packed_mlp && origins && dirs && lengths, "holo_render_fwd: null input");
    HOLO_CHECK_ARG(features && depths && masks, "holo_render_fwd: null output");
    HOLO_CHECK_ARG(n_passes == 1 || n_passes == 2, "holo_render_fwd: n_passes must be 1 or 2 (got %d)", n_passes);
    HOLO_CHECK_ARG(S >= 2 && (n_passes == 1 || (n_fine >= 1 && S >= 3)), "holo_render_fwd: bad S/n_fine");
    HOLO_CHECK_ARG(D > 1 && H > 1 && W > 1, "holo_render_fwd: grid must be at least 2^3");
    RenderParams P;
    P.grid = grid_dhwc;
    P.D = D, P.Hh = H, P.Ww = W;
    // VolumeLocator: local = world * (1 / ((R-1) * (extent/R) / 2)) per axis (x->W, y->H, z->D)
    P.inv_scale = 1.0f / ((float)(W - 1) * (volume_extent / (float)W) * 0.5f);
    P.inv_scale_y = 1.0f / ((float)(H - 1) * (volume_extent / (float)H) * 0.5f);
    P.inv_scale_z = 1.0f / ((float)(D - 1) * (volume_extent / (float)D) * 0.5f);
    P.mlp = packed_mlp;
    P.Hd = hidden;
    P.n_harm = n_harmonic;
    P.origins = origins, P.dirs = dirs, P.lengths = lengths;
    P.n_rays = n_rays, P.S = S, P.n_fine = n_passes > 1 ? n_fine : 0, P.add_input = add_input_samples ? 1 : 0;
    P.bg[0] = bg3_host ? bg3_host[0] : 1.f, P.bg[1] = bg3_host ? bg3_host[1] : 1.f, P.bg[2] = bg3_host ? bg3_host[2] : 1.f;
    P.bg_opacity = background_opacity;
    P.features = features, P.depths = depths, P.masks = masks, P.weights = weights, P.lengths_out = lengths_out;
    P.p_features = prev_features, P.p_depths = prev_depths, P.p_masks = prev_masks, P.p_weights = prev_weights;
    P.n_passes = n_passes;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 8: return launch_render<8, 2>(P, st);
        case 16: return launch_render<16, 2>(P, st);
        case 32: return launch_render<32, 2>(P, st);
        case 64: return launch_render<64, 1>(P, st);
        case 128: return launch_render<128, 1>(P, st);
        default:
            holo_set_error("holo_render_fwd: unsupported channel count %d (8,16,32,64,128)", C);
            return HOLO_ERR_UNSUPPORTED;
    }
}
