// Per-view post-processing of generate_samples.py on the GPU: what the reference does on
// the host for every rendered frame before it reaches the video writer --
//   /root/reference/holo_diffusion/utils/render_utils/flyaround.py:422-488  (_images_from_preds: depth image via
//       pytorch3d vis_utils.make_depth_image, white compositing, channel repeat, .cpu())
//   /root/reference/holo_diffusion/utils/render_utils/flyaround.py:558-610  (_generate_prediction_videos: clip to [0, 1],
//       resize, 8-bit frames)
//   /root/reference/holo_diffusion/utils/render_utils/shaded_depth_render.py:143-206 + mesh_render.py (depth -> smoothed
//       depth -> camera-space mesh -> Phong-shaded render from the same camera), here as a screen-space kernel.
// Once a view costs ~9 ms these host passes (quantiles by topk on 65k pixels, a mesh rasteriser, PIL resizes, one
// blocking .cpu() per key) dominate the wall clock of the unmodified script.
#include "common.cuh"
#include "../../include/holo_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// make_depth_image: normalise the depth between its (min_quantile, max_quantile) order statistics over the pixels with
// depth > 1e-6 and mask > 0.5 -- EXACT selection (k-th smallest / k-th largest as pytorch3d's topk), by a 4-pass
// 8-bit radix select on the bit patterns (positive floats order like their bits).  One CTA.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool px_ok(float d, float m) { return d > 1e-6f && m > 0.5f; }

// k-th smallest (k >= 1) of the keys key(i) = bits(d_i) ^ flip over the ok pixels; flip = 0xffffffff selects the k-th
// largest.  All threads of the CTA call it; result broadcast through shared memory.
__device__ uint32_t radix_select(const float* __restrict__ d, const float* __restrict__ m, int n, uint32_t flip,
                                 unsigned long long k, uint32_t* hist /*[256]*/, uint32_t* bcast /*[2]*/) {
    uint32_t prefix = 0, prefix_mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float di = d[i];
            if (!px_ok(di, m[i])) continue;
            const uint32_t key = __float_as_uint(di) ^ flip;
            if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long run = 0;
            uint32_t b = 0;
            for (; b < 256; ++b) {
                if (run + hist[b] >= k) break;
                run += hist[b];
            }
            if (b == 256) b = 255;   // cannot happen for 1 <= k <= n_ok
            bcast[0] = b;
            bcast[1] = (uint32_t)(k - run);
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        prefix_mask |= 0xffu << shift;
        k = bcast[1];
        __syncthreads();
    }
    return prefix ^ flip;
}

__global__ void depth_normfac_kernel(const float* __restrict__ d, const float* __restrict__ m, int n, double min_q,
                                     double max_q, float* __restrict__ normfac /*[2] = (min, max)*/) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bcast[2];
    __shared__ unsigned long long n_ok_s;
    if (threadIdx.x == 0) n_ok_s = 0;
    __syncthreads();
    unsigned long long cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += px_ok(d[i], m[i]) ? 1 : 0;
    atomicAdd(&n_ok_s, cnt);
    __syncthreads();
    const unsigned long long n_ok = n_ok_s;
    if (n_ok <= 1) {   // "if ok.sum() <= 1: normfacs.append(zeros(2))"
        if (threadIdx.x == 0) normfac[0] = 0.f, normfac[1] = 0.f;
        return;
    }
    // max(int(round(q * numel)), 1): Python's round() is half-to-even = rint() on the double product
    long long maxk = (long long)rint((1.0 - max_q) * (double)n_ok);
    long long mink = (long long)rint(min_q * (double)n_ok);
    if (maxk < 1) maxk = 1;
    if (mink < 1) mink = 1;
    if (maxk > (long long)n_ok) maxk = (long long)n_ok;
    if (mink > (long long)n_ok) mink = (long long)n_ok;
    const uint32_t vmin = radix_select(d, m, n, 0u, (unsigned long long)mink, hist, bcast);
    const uint32_t vmax = radix_select(d, m, n, 0xffffffffu, (unsigned long long)maxk, hist, bcast);
    if (threadIdx.x == 0) normfac[0] = __uint_as_float(vmin), normfac[1] = __uint_as_float(vmax);
}

// depth image = ((d - min) / clamp(max - min, 1e-4) * (hi - lo) + lo) * mask, clamped to [0, 1]; then the reference's
// white compositing v * mask + (1 - mask) (flyaround.py:476-477) and the repeat to 3 channels (:478-479)
__global__ void depth_image_kernel(const float* __restrict__ d, const float* __restrict__ m, int n,
                                   const float* __restrict__ normfac, float lo, float hi, int composite_white,
                                   float* __restrict__ out3 /*(3, n)*/) {
    const float mn = normfac[0], mx = normfac[1];
    const float inv = 1.0f / fmaxf(mx - mn, 1e-4f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float mi = m[i];
        float v = (d[i] - mn) * inv;
        v = fminf(fmaxf((v * (hi - lo) + lo) * mi, 0.f), 1.f);
        if (composite_white) v = v * mi + (1.f - mi);
        out3[i] = v, out3[n + i] = v, out3[2 * n + i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// frame packing: (C in {1, 3}, H, W) float -> (h, w, 3) uint8, clip to [0, 1], bilinear resize (pixel centres aligned
// as F.interpolate(align_corners=False)), round to nearest -- the numpy / PIL leg of VideoWriter.write_frame
// ---------------------------------------------------------------------------------------------------------------
__global__ void frame_u8_kernel(const float* __restrict__ src, int C, int H, int W, int h, int w,
                                uint8_t* __restrict__ dst) {
    const int n = h * w;
    const float sy = (float)H / (float)h, sx = (float)W / (float)w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i % w;
        float fy = ((float)y + 0.5f) * sy - 0.5f, fx = ((float)x + 0.5f) * sx - 0.5f;
        fy = fmaxf(fy, 0.f), fx = fmaxf(fx, 0.f);
        const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float wy = fy - (float)y0, wx = fx - (float)x0;
        auto tap = [](float t) { return fminf(fmaxf(t, 0.f), 1.f); };   // the reference clips BEFORE it resizes
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* p = src + (size_t)(C == 1 ? 0 : c) * H * W;
            const float v = (1.f - wy) * ((1.f - wx) * tap(p[y0 * W + x0]) + wx * tap(p[y0 * W + x1])) +
                            wy * ((1.f - wx) * tap(p[y1 * W + x0]) + wx * tap(p[y1 * W + x1]));
            dst[(size_t)i * 3 + c] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shaded depth, screen space.  The reference unprojects the (box-smoothed) depth map with the view's intrinsics into a
// camera-space vertex grid, triangulates the quads whose four pixels are valid, and renders that mesh from the SAME
// camera with a point light at the camera centre and a Phong material (shaded_depth_render.py:46-141).  Seen from its
// own camera every pixel shows its own vertex, so the render reduces to: vertex normal = normalised sum of the
// (area-weighted) normals of the incident triangles, Phong shading at the vertex, white elsewhere.
// Pass 1: box-smoothed depth (avg_pool2d of [depth * ok, ok], window 2k+1, zero padding; :15-24).
// Pass 2: vertex, triangle normals, shading.
// ---------------------------------------------------------------------------------------------------------------
__global__ void smooth_depth_kernel(const float* __restrict__ d, const float* __restrict__ m, int H, int W, int k,
                                    float mask_thr, float depth_thr, float* __restrict__ ds, uint8_t* __restrict__ ok_out) {
    const int n = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i % W;
        float sd = 0.f, sm = 0.f;
        for (int dy = -k; dy <= k; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -k; dx <= k; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const float dv = d[yy * W + xx];
                const float okv = (m[yy * W + xx] > mask_thr && dv > depth_thr) ? 1.f : 0.f;
                // the reference pools the RAW depth (cat((g, m)) pools g itself, not g * m) and the mask separately
                sd += dv, sm += okv;
            }
        }
        const float area = (float)((2 * k + 1) * (2 * k + 1));   // avg_pool2d counts the zero padding
        ds[i] = (sd / area) / fmaxf(sm / area, 1e-4f);
        ok_out[i] = (m[i] > mask_thr && d[i] > depth_thr) ? 1 : 0;
    }
}

struct ShadeParams {
    int H, W;
    float fx, fy, px, py;          // NDC intrinsics of the view
    float ndc_rx, ndc_ry;          // NDC half-extent of the image (short side = 1)
    float ambient[3], diffuse[3], specular[3];
    float shininess;
    float bg[3];
};

__device__ __forceinline__ float3 unproject(const ShadeParams& P, const float* ds, int y, int x) {
    // NDC pixel centre (+x left, +y up), camera-space point at depth z along the unit-depth ray ... the reference
    // replaces the ray LENGTHS by the depth (shaded_depth_render.py:171-179): p = origin + depth * direction with the
    // direction normalised by NDCGridRaysampler's default (un-normalised: direction has z = 1)
    const float xn = P.ndc_rx - (2.f * P.ndc_rx) * ((float)x + 0.5f) / (float)P.W;
    const float yn = P.ndc_ry - (2.f * P.ndc_ry) * ((float)y + 0.5f) / (float)P.H;
    const float z = ds[y * P.W + x];
    return make_float3((xn - P.px) / P.fx * z, (yn - P.py) / P.fy * z, z);
}
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__global__ void shade_depth_kernel(const float* __restrict__ ds, const uint8_t* __restrict__ ok, ShadeParams P,
                                   float* __restrict__ out3 /*(3, H, W)*/, float* __restrict__ out_mask /*(H, W)*/) {
    const int n = P.H * P.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / P.W, x = i % P.W;
        float3 nsum = make_float3(0.f, 0.f, 0.f);
        bool any = false;
        // the four quads that touch pixel (y, x): quad (qy, qx) spans pixels (qy..qy+1, qx..qx+1) and exists when all
        // four are valid; it is split into triangles (a, c, b) and (b, c, d) with a = (qy, qx), b = (qy, qx+1),
        // c = (qy+1, qx), d = (qy+1, qx+1) (get_grid_mesh, shaded_depth_render.py:248-280, winding as there)
        for (int qy = y - 1; qy <= y; ++qy)
            for (int qx = x - 1; qx <= x; ++qx) {
                if (qy < 0 || qx < 0 || qy + 1 >= P.H || qx + 1 >= P.W) continue;
                if (!(ok[qy * P.W + qx] && ok[qy * P.W + qx + 1] && ok[(qy + 1) * P.W + qx] && ok[(qy + 1) * P.W + qx + 1]))
                    continue;
                const float3 a = unproject(P, ds, qy, qx), b = unproject(P, ds, qy, qx + 1);
                const float3 c = unproject(P, ds, qy + 1, qx), d = unproject(P, ds, qy + 1, qx + 1);
                const bool in1 = (y == qy && x == qx) || (y == qy && x == qx + 1) || (y == qy + 1 && x == qx);   // a, b, c
                const bool in2 = !(y == qy && x == qx);                                                         // b, c, d
                // face normal ~ cross(v1 - v0, v2 - v0), un-normalised = area weighting (pytorch3d verts_normals)
                if (in1) {
                    const float3 fn = cross3(sub3(c, a), sub3(b, a));
                    nsum.x += fn.x, nsum.y += fn.y, nsum.z += fn.z;
                }
                if (in2) {
                    const float3 fn = cross3(sub3(c, b), sub3(d, b));
                    nsum.x += fn.x, nsum.y += fn.y, nsum.z += fn.z;
                }
                any = true;
            }
        float r = P.bg[0], g = P.bg[1], bl = P.bg[2], mk = 0.f;
        if (any) {
            const float inv = rsqrtf(fmaxf(dot3(nsum, nsum), 1e-24f));
            float3 nrm = make_float3(nsum.x * inv, nsum.y * inv, nsum.z * inv);
            const float3 p = unproject(P, ds, y, x);
            const float pl = rsqrtf(fmaxf(dot3(p, p), 1e-24f));
            const float3 l = make_float3(-p.x * pl, -p.y * pl, -p.z * pl);   // light = view direction: camera at the origin
            if (dot3(nrm, l) < 0.f) nrm = make_float3(-nrm.x, -nrm.y, -nrm.z);   // the surface faces its own camera
            const float ndl = fmaxf(dot3(nrm, l), 0.f);
            // Phong: reflect(l, n) . v with v = l  ->  2 (n.l)^2 - 1
            const float rv = fmaxf(2.f * ndl * ndl - 1.f, 0.f);
            const float sp = powf(rv, P.shininess);
            r = P.ambient[0] + P.diffuse[0] * ndl + P.specular[0] * sp;
            g = P.ambient[1] + P.diffuse[1] * ndl + P.specular[1] * sp;
            bl = P.ambient[2] + P.diffuse[2] * ndl + P.specular[2] * sp;
            mk = 1.f;
        }
        out3[i] = fminf(fmaxf(r, 0.f), 1.f), out3[n + i] = fminf(fmaxf(g, 0.f), 1.f), out3[2 * n + i] = fminf(fmaxf(bl, 0.f), 1.f);
        if (out_mask) out_mask[i] = mk;
    }
}

inline int grid_for(int n) {
    int b = holo_cdiv(n, 256);
    return b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b);
}

}  // namespace

extern "C" int holo_depth_image(const float* depth, const float* mask, int n_pixels, float min_quantile,
                                float max_quantile, float min_out_depth, float max_out_depth, int composite_white,
                                float* normfac2, float* out_3n, void* stream) {
    HOLO_CHECK_ARG(depth && mask && normfac2 && out_3n && n_pixels > 0, "holo_depth_image: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    depth_normfac_kernel<<<1, 1024, 0, st>>>(depth, mask, n_pixels, (double)min_quantile, (double)max_quantile, normfac2);
    HOLO_CHECK_LAUNCH("holo_depth_image (quantiles)");
    depth_image_kernel<<<grid_for(n_pixels), 256, 0, st>>>(depth, mask, n_pixels, normfac2, min_out_depth, max_out_depth,
                                                           composite_white, out_3n);
    HOLO_CHECK_LAUNCH("holo_depth_image");
    return HOLO_OK;
}

extern "C" int holo_frame_u8(const float* src_chw, int C, int H, int W, int out_h, int out_w, void* dst_hw3_u8,
                             void* stream) {
    HOLO_CHECK_ARG(src_chw && dst_hw3_u8 && (C == 1 || C == 3) && H > 0 && W > 0 && out_h > 0 && out_w > 0,
                   "holo_frame_u8: bad args (C must be 1 or 3)");
    frame_u8_kernel<<<grid_for(out_h * out_w), 256, 0, (cudaStream_t)stream>>>(src_chw, C, H, W, out_h, out_w,
                                                                               (uint8_t*)dst_hw3_u8);
    HOLO_CHECK_LAUNCH("holo_frame_u8");
    return HOLO_OK;
}

extern "C" int holo_shade_depth(const float* depth, const float* mask, int H, int W, float fx, float fy, float px,
                                float py, int smooth_k, float mask_thr, float depth_thr, const float* material10_host,
                                const float* bg3_host, float* scratch_depth, void* scratch_ok_u8, float* out_3hw,
                                float* out_mask, void* stream) {
    HOLO_CHECK_ARG(depth && mask && material10_host && bg3_host && scratch_depth && scratch_ok_u8 && out_3hw && H > 1 &&
                       W > 1 && smooth_k >= 0,
                   "holo_shade_depth: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    smooth_depth_kernel<<<grid_for(H * W), 256, 0, st>>>(depth, mask, H, W, smooth_k, mask_thr, depth_thr, scratch_depth,
                                                         (uint8_t*)scratch_ok_u8);
    HOLO_CHECK_LAUNCH("holo_shade_depth (smooth)");
    ShadeParams P;
    P.H = H, P.W = W, P.fx = fx, P.fy = fy, P.px = px, P.py = py;
    P.ndc_rx = W >= H ? (float)W / (float)H : 1.f;
    P.ndc_ry = W >= H ? 1.f : (float)H / (float)W;
    for (int i = 0; i < 3; ++i)
        P.ambient[i] = material10_host[i], P.diffuse[i] = material10_host[3 + i], P.specular[i] = material10_host[6 + i],
        P.bg[i] = bg3_host[i];
    P.shininess = material10_host[9];
    shade_depth_kernel<<<grid_for(H * W), 256, 0, st>>>(scratch_depth, (const uint8_t*)scratch_ok_u8, P, out_3hw, out_mask);
    HOLO_CHECK_LAUNCH("holo_shade_depth");
    return HOLO_OK;
}
