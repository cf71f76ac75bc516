// Streaming (HBM-bound) kernels of the denoiser: layout changes, GroupNorm(32) statistics / affine+SiLU,
// timestep embedding + small linears, the fused DDPM ancestral step, tanh + range check.
//
// Reference path replaced (relative to /root/reference/holo_diffusion):
//   guided_diffusion/nn.py:23-25,99-106      GroupNorm32(32, C), eps 1e-5
//   guided_diffusion/unet.py:184,208,248-252 GN -> SiLU, FiLM  out_norm(h) * (1 + scale) + shift
//   guided_diffusion/nn.py:109-127           timestep_embedding
//   guided_diffusion/unet.py:646-650,199-205 time_embed / emb_layers linears
//   guided_diffusion/gaussian_diffusion.py:318,229-251,498-506  clamp, posterior mean, ancestral sample
//   holo_diffusion_model.py:424-428          tanh + range asserts
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/holo_b200.h"

// ------------------------------------------------------------------------------------------------
// layout: (C, V) channels-first <-> (V, C) channels-last, smem-tiled transpose
// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int swap_xy) {
    // src is (rows, cols) row-major, dst is (cols, rows)
    __shared__ float tile[32][33];
    // the longer axis rides grid.x (2^31 - 1 blocks); grid.y is limited to 65535 (a 128^3 grid has V / 32 = 65536)
    const int bx = swap_xy ? blockIdx.y : blockIdx.x, by = swap_xy ? blockIdx.x : blockIdx.y;
    int c0 = bx * 32, r0 = by * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

extern "C" int holo_transpose2d(const float* src, float* dst, int rows, int cols, void* stream) {
    HOLO_CHECK_ARG(src && dst && rows > 0 && cols > 0, "holo_transpose2d: bad args");
    const int swap_xy = rows > cols;
    dim3 grid(holo_cdiv(swap_xy ? rows : cols, 32), holo_cdiv(swap_xy ? cols : rows, 32));
    HOLO_CHECK_ARG(grid.y <= 65535, "holo_transpose2d: both dimensions exceed 65535 * 32 (%d x %d)", rows, cols);
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, rows, cols, swap_xy);
    HOLO_CHECK_LAUNCH("holo_transpose2d");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics over a channels-last tensor made of up to two sources (the skip concat, unet.py:829).
// Every block reduces a slab of voxels; per-channel partials in fp32, combined per group and accumulated
// into global fp64 (sum, sumsq), 8 replicas x 32 groups x 2.  acc (512 doubles) must be zeroed (holo_gn_finalize
// re-zeroes it after use).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                        int C2, long long V, int vox_per_block,
                                                        double* __restrict__ acc, double* __restrict__ acc_zero) {
    extern __shared__ float sh[];  // [vstep][2C] per-thread partials, then [2C] totals
    holo_pdl_trigger();   // opt-in PDL build only (common.cuh): no dependent access before the wait
    holo_pdl_wait();
    // ping-pong accumulators: while this GroupNorm accumulates into `acc`, block 0 clears the buffer the NEXT one
    // will use (its previous reader, the preceding fused apply, has completed in stream order)
    if (acc_zero && blockIdx.x == 0)
        for (int i = threadIdx.x; i < 512; i += blockDim.x) acc_zero[i] = 0.0;
    const int C = C1 + C2;
    const int cq = C / 4;  // float4 lanes per voxel
    long long v0 = (long long)blockIdx.x * vox_per_block;
    long long v1 = v0 + vox_per_block;
    if (v1 > V) v1 = V;
    // thread t owns float4 lane (t % cq) and walks voxels with stride blockDim/cq
    const int lane = threadIdx.x % cq;
    const int vrow = threadIdx.x / cq;
    const int vstep = blockDim.x / cq;
    const int c = lane * 4;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    if (vrow < vstep) {
        const float* base = (c < C1) ? x1 + c : x2 + (c - C1);
        const int pitch = (c < C1) ? C1 : C2;
#pragma unroll 4
        for (long long v = v0 + vrow; v < v1; v += vstep) {
            float4 a = __ldg(reinterpret_cast<const float4*>(base + v * pitch));
            s.x += a.x, s.y += a.y, s.z += a.z, s.w += a.w;
            q.x = fmaf(a.x, a.x, q.x), q.y = fmaf(a.y, a.y, q.y), q.z = fmaf(a.z, a.z, q.z), q.w = fmaf(a.w, a.w, q.w);
        }
        float* row = sh + (size_t)vrow * 2 * C;
        *reinterpret_cast<float4*>(row + c) = s;
        *reinterpret_cast<float4*>(row + C + c) = q;
    }
    __syncthreads();
    // column sums over the vstep rows: thread t < 2C owns one (channel, sum|sumsq) column
    for (int col = threadIdx.x; col < 2 * C; col += blockDim.x) {
        float tot = 0.f;
        for (int r = 0; r < vstep; ++r) tot += sh[(size_t)r * 2 * C + col];
        sh[col] = tot;  // row 0 becomes the totals (each column is owned by one thread)
    }
    __syncthreads();
    const int cpg = C / 32;
    if (threadIdx.x < 64) {
        const int g = threadIdx.x % 32, which = threadIdx.x / 32;  // which: 0 sum, 1 sumsq
        double a = 0;
        for (int k = 0; k < cpg; ++k) a += (double)sh[which * C + g * cpg + k];
        // 8 replicas of the accumulator spread the same-address fp64 atomics of ~1000 CTAs
        atomicAdd(&acc[(blockIdx.x & 7) * 64 + g * 2 + which], a);
    }
}

static int gn_stats_launch(const float* x1, int C1, const float* x2, int C2, long long V, double* acc64,
                           double* acc_zero, void* stream) {
    int C = C1 + C2;
    HOLO_CHECK_ARG(x1 && acc64 && V > 0, "holo_gn_stats: bad args");
    HOLO_CHECK_ARG(C % 32 == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C <= 1024, "holo_gn_stats: C=%d+%d unsupported", C1, C2);
    HOLO_CHECK_ARG(C2 == 0 || x2, "holo_gn_stats: second source missing");
    int threads = 256;
    if (C / 4 > threads) threads = C / 4;
    // ~4 CTAs per SM for the big tensors; the coarse levels (V = 64 .. 4096) still get several CTAs
    long long vpb = (V + 148 * 4 - 1) / (148 * 4);
    if (vpb < 8) vpb = 8;
    int blocks = holo_cdiv(V, vpb);
    const int vstep = threads / (C / 4);
    size_t smem = (size_t)(vstep > 1 ? vstep : 1) * 2 * C * sizeof(float);
    holo_launch(gn_stats_kernel, dim3(blocks), dim3(threads), (size_t)smem, (cudaStream_t)stream, x1, C1, x2, C2, V,
                (int)vpb, acc64, acc_zero);
    HOLO_CHECK_LAUNCH("holo_gn_stats");
    return HOLO_OK;
}

extern "C" int holo_gn_stats(const float* x1, int C1, const float* x2, int C2, long long V, double* acc64,
                             void* stream) {
    return gn_stats_launch(x1, C1, x2, C2, V, acc64, nullptr, stream);
}
extern "C" int holo_gn_stats_pp(const float* x1, int C1, const float* x2, int C2, long long V, double* acc64,
                                double* acc64_next, void* stream) {
    HOLO_CHECK_ARG(acc64_next && acc64_next != acc64, "holo_gn_stats_pp: need a distinct buffer to clear");
    return gn_stats_launch(x1, C1, x2, C2, V, acc64, acc64_next, stream);
}

// per-channel affine from the statistics:  y = x * a[c] + b[c]
//   plain GN : a = rstd*gamma,            b = beta - mean*rstd*gamma
//   FiLM     : a *= (1+scale[c]),         b = b*(1+scale[c]) + shift[c]      (unet.py:248-252)
__global__ void gn_finalize_kernel(double* __restrict__ acc, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ film, int C,
                                   double count, float eps, float* __restrict__ a, float* __restrict__ b) {
    // single block: after every thread has read the statistics they are re-zeroed for the next GroupNorm
    __shared__ double s_acc[64];
    if (threadIdx.x < 64) {
        double a = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) a += acc[r * 64 + threadIdx.x];
        s_acc[threadIdx.x] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) acc[i] = 0.0;
    int cpg = C / 32;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        double mean = s_acc[g * 2] / count;
        double var = s_acc[g * 2 + 1] / count - mean * mean;
        if (var < 0) var = 0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float aa = rstd * gamma[c];
        float bb = beta[c] - (float)mean * aa;
        if (film) {
            float sc = 1.f + film[c], sh = film[C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        a[c] = aa, b[c] = bb;
    }
}
__global__ void zero_f64_kernel(double* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

extern "C" int holo_gn_finalize(double* acc64, const float* gamma, const float* beta, const float* film_scale_shift,
                                int C, long long V, float eps, float* a, float* b, void* stream) {
    HOLO_CHECK_ARG(acc64 && gamma && beta && a && b && C % 32 == 0, "holo_gn_finalize: bad args");
    double count = (double)V * (double)(C / 32);
    gn_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(acc64, gamma, beta, film_scale_shift, C, count, eps, a, b);
    HOLO_CHECK_LAUNCH("holo_gn_finalize");
    return HOLO_OK;
}

// y = act(x*a+b) over the (possibly concatenated) channels-last tensor.  Optional bf16 hi/lo split outputs
// feed the tensor-core convolution (x = hi + lo + O(2^-17 x)).
template <bool SILU>
__global__ void gn_apply_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                long long V, const float* __restrict__ a, const float* __restrict__ b,
                                float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo) {
    const int C = C1 + C2;
    const long long total4 = V * (C / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        long long v = i / (C / 4);
        int c = (int)(i % (C / 4)) * 4;
        float4 xv = (c < C1) ? *reinterpret_cast<const float4*>(x1 + v * C1 + c)
                             : *reinterpret_cast<const float4*>(x2 + v * C2 + (c - C1));
        float4 av = *reinterpret_cast<const float4*>(a + c);
        float4 bv = *reinterpret_cast<const float4*>(b + c);
        float4 r;
        r.x = xv.x * av.x + bv.x, r.y = xv.y * av.y + bv.y, r.z = xv.z * av.z + bv.z, r.w = xv.w * av.w + bv.w;
        if (SILU) r.x = holo_silu(r.x), r.y = holo_silu(r.y), r.z = holo_silu(r.z), r.w = holo_silu(r.w);
        if (y) *reinterpret_cast<float4*>(y + v * C + c) = r;
        if (y_hi) {
            float rr[4] = {r.x, r.y, r.z, r.w};
            uint16_t hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat16 h = __float2bfloat16_rn(rr[k]);
                __nv_bfloat16 l = __float2bfloat16_rn(rr[k] - __bfloat162float(h));
                hi[k] = *reinterpret_cast<uint16_t*>(&h);
                lo[k] = *reinterpret_cast<uint16_t*>(&l);
            }
            *reinterpret_cast<uint2*>(y_hi + v * C + c) =
                make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
            *reinterpret_cast<uint2*>(y_lo + v * C + c) =
                make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
        }
    }
}

// GroupNorm apply with the finalize step folded in: every CTA derives the per-channel affine (gamma, beta, FiLM,
// statistics) into shared memory, then streams its share of the tensor.  Saves one launch per GroupNorm.
template <bool SILU>
__global__ void gn_apply_fused_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                      long long V, const double* __restrict__ acc, const double* __restrict__ ch1,
                                      const double* __restrict__ ch2, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ film, float eps,
                                      float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                                      uint16_t* __restrict__ raw_hi, uint16_t* __restrict__ raw_lo, int pair_f16) {
    extern __shared__ float s_ab[];  // a[C], b[C]
    __shared__ double s_acc[64];
    const int C = C1 + C2;
    const int cpg = C / 32;
    holo_pdl_trigger();   // opt-in PDL build only (common.cuh): the statistics below come from the predecessor
    holo_pdl_wait();
    if (threadIdx.x < 64) {
        double t = 0;
        if (acc) {  // group statistics from holo_gn_stats (8 replicas)
#pragma unroll
            for (int r = 0; r < 8; ++r) t += acc[r * 64 + threadIdx.x];
        } else {    // per-channel statistics left by the producing convolutions' epilogues
            const int g = threadIdx.x / 2, which = threadIdx.x % 2;
            for (int k = 0; k < cpg; ++k) {
                const int c = g * cpg + k;
                t += (c < C1) ? ch1[c * 2 + which] : ch2[(c - C1) * 2 + which];
            }
        }
        s_acc[threadIdx.x] = t;
    }
    __syncthreads();
    const double count = (double)V * (double)cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        double mean = s_acc[g * 2] / count;
        double var = s_acc[g * 2 + 1] / count - mean * mean;
        if (var < 0) var = 0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float aa = rstd * gamma[c];
        float bb = beta[c] - (float)mean * aa;
        if (film) {
            float sc = 1.f + film[c], sh = film[C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        s_ab[c] = aa, s_ab[C + c] = bb;
    }
    __syncthreads();
    if (((C1 | C2) & 7) == 0 && ((long long)gridDim.x * blockDim.x) % (C / 8) == 0) {
        // fast path: 8 channels per thread and a stride that is a multiple of the channel-group count, so the thread's
        // channel slice (and its affine, kept in registers) never changes and the loop carries no division
        const int q8 = C / 8;
        const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int c = (int)(tid % q8) * 8;
        const long long vstep = ((long long)gridDim.x * blockDim.x) / q8;
        float a8[8], b8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a8[k] = s_ab[c + k], b8[k] = s_ab[C + c + k];
        const float* src = (c < C1) ? x1 + c : x2 + (c - C1);
        const int pitch = (c < C1) ? C1 : C2;
        for (long long v = tid / q8; v < V; v += vstep) {
            const float4 u0 = *reinterpret_cast<const float4*>(src + v * pitch);
            const float4 u1 = *reinterpret_cast<const float4*>(src + v * pitch + 4);
            float r[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            if (raw_hi) {
                // the un-normalised tensor as a bf16 hi/lo pair too (operand of the ResBlock's 1x1 skip convolution,
                // unet.py:222,255): saves the separate split pass re-reading the concat
                uint32_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) holo_split2(r[2 * k], r[2 * k + 1], pair_f16 != 0, h[k], l[k]);
                *reinterpret_cast<uint4*>(raw_hi + v * C + c) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(raw_lo + v * C + c) = make_uint4(l[0], l[1], l[2], l[3]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                r[k] = fmaf(r[k], a8[k], b8[k]);
                if (SILU) r[k] = __fdividef(r[k], 1.0f + __expf(-r[k]));  // 2 MUFU ops; ~1e-6 relative
            }
            if (y) {
                *reinterpret_cast<float4*>(y + v * C + c) = make_float4(r[0], r[1], r[2], r[3]);
                *reinterpret_cast<float4*>(y + v * C + c + 4) = make_float4(r[4], r[5], r[6], r[7]);
            }
            if (y_hi) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) holo_split2(r[2 * k], r[2 * k + 1], pair_f16 != 0, h[k], l[k]);
                *reinterpret_cast<uint4*>(y_hi + v * C + c) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(y_lo + v * C + c) = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
        return;
    }
    const long long total4 = V * (C / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        long long v = i / (C / 4);
        int c = (int)(i % (C / 4)) * 4;
        float4 xv = (c < C1) ? *reinterpret_cast<const float4*>(x1 + v * C1 + c)
                             : *reinterpret_cast<const float4*>(x2 + v * C2 + (c - C1));
        float4 av = *reinterpret_cast<const float4*>(s_ab + c);
        float4 bv = *reinterpret_cast<const float4*>(s_ab + C + c);
        float4 r;
        r.x = xv.x * av.x + bv.x, r.y = xv.y * av.y + bv.y, r.z = xv.z * av.z + bv.z, r.w = xv.w * av.w + bv.w;
        if (SILU) r.x = holo_silu(r.x), r.y = holo_silu(r.y), r.z = holo_silu(r.z), r.w = holo_silu(r.w);
        if (y) *reinterpret_cast<float4*>(y + v * C + c) = r;
        if (y_hi) {
            uint32_t h2[2], l2[2];
            holo_split2(r.x, r.y, pair_f16 != 0, h2[0], l2[0]);
            holo_split2(r.z, r.w, pair_f16 != 0, h2[1], l2[1]);
            *reinterpret_cast<uint2*>(y_hi + v * C + c) = make_uint2(h2[0], h2[1]);
            *reinterpret_cast<uint2*>(y_lo + v * C + c) = make_uint2(l2[0], l2[1]);
        }
    }
}

static int gn_apply_fused_launch(const float* x1, int C1, const float* x2, int C2, long long V, const double* acc64,
                                 const double* ch1, const double* ch2, const float* gamma, const float* beta,
                                 const float* film_scale_shift, float eps, int silu, float* y, void* y_hi_bf16,
                                 void* y_lo_bf16, void* raw_hi_bf16, void* raw_lo_bf16, int pair_f16, void* stream) {
    int C = C1 + C2;
    HOLO_CHECK_ARG(x1 && (acc64 || ch1) && gamma && beta && (y || y_hi_bf16) && V > 0, "holo_gn_apply_fused: bad args");
    HOLO_CHECK_ARG((raw_hi_bf16 == nullptr) == (raw_lo_bf16 == nullptr), "holo_gn_apply_fused: raw hi/lo come together");
    HOLO_CHECK_ARG(!raw_hi_bf16 || ((C1 | C2) & 7) == 0, "holo_gn_apply_fused: raw split needs channel counts % 8 == 0");
    HOLO_CHECK_ARG(acc64 || C2 == 0 || ch2, "holo_gn_apply_fused_ch: statistics of the second source missing");
    HOLO_CHECK_ARG(C % 32 == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C <= 4096, "holo_gn_apply_fused: C=%d+%d unsupported", C1, C2);
    HOLO_CHECK_ARG((y_hi_bf16 == nullptr) == (y_lo_bf16 == nullptr), "holo_gn_apply_fused: hi/lo must come together");
    long long total4 = V * (C / 4);
    int blocks = holo_cdiv(total4, 256 * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (C % 8 == 0) {
        // the kernel's fast path wants (blocks * 256) % (C / 8) == 0: round the grid to a multiple of q8 / gcd(q8, 256)
        int q8 = C / 8, g = q8, b = 256;
        while (b) {
            int t = g % b;
            g = b, b = t;
        }
        const int m = q8 / g;
        blocks = blocks < m ? m : blocks / m * m;
    }
    size_t smem = 2 * (size_t)C * sizeof(float);
    if (silu)
        holo_launch(gn_apply_fused_kernel<true>, dim3(blocks), dim3(256), smem, (cudaStream_t)stream, x1, C1, x2, C2, V,
                    acc64, ch1, ch2, gamma, beta, film_scale_shift, eps, y, (uint16_t*)y_hi_bf16, (uint16_t*)y_lo_bf16,
                    (uint16_t*)raw_hi_bf16, (uint16_t*)raw_lo_bf16, pair_f16);
    else
        holo_launch(gn_apply_fused_kernel<false>, dim3(blocks), dim3(256), smem, (cudaStream_t)stream, x1, C1, x2, C2, V,
                    acc64, ch1, ch2, gamma, beta, film_scale_shift, eps, y, (uint16_t*)y_hi_bf16, (uint16_t*)y_lo_bf16,
                    (uint16_t*)raw_hi_bf16, (uint16_t*)raw_lo_bf16, pair_f16);
    HOLO_CHECK_LAUNCH("holo_gn_apply_fused");
    return HOLO_OK;
}

extern "C" int holo_gn_apply_fused(const float* x1, int C1, const float* x2, int C2, long long V, const double* acc64,
                                   const float* gamma, const float* beta, const float* film_scale_shift, float eps,
                                   int silu, float* y, void* y_hi_bf16, void* y_lo, void* raw_hi_bf16,
                                   void* raw_lo, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(acc64, "holo_gn_apply_fused: null statistics");
    return gn_apply_fused_launch(x1, C1, x2, C2, V, acc64, nullptr, nullptr, gamma, beta, film_scale_shift, eps, silu, y,
                                 y_hi_bf16, y_lo, raw_hi_bf16, raw_lo, pair_f16, stream);
}

extern "C" int holo_gn_apply_fused_ch(const float* x1, int C1, const double* ch_stats1, const float* x2, int C2,
                                      const double* ch_stats2, long long V, const float* gamma, const float* beta,
                                      const float* film_scale_shift, float eps, int silu, float* y, void* y_hi_bf16,
                                      void* y_lo, void* raw_hi_bf16, void* raw_lo, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(ch_stats1, "holo_gn_apply_fused_ch: null statistics");
    return gn_apply_fused_launch(x1, C1, x2, C2, V, nullptr, ch_stats1, ch_stats2, gamma, beta, film_scale_shift, eps,
                                 silu, y, y_hi_bf16, y_lo, raw_hi_bf16, raw_lo, pair_f16, stream);
}

extern "C" int holo_gn_apply(const float* x1, int C1, const float* x2, int C2, long long V, const float* a,
                             const float* b, int silu, float* y, void* y_hi_bf16, void* y_lo_bf16, void* stream) {
    int C = C1 + C2;
    HOLO_CHECK_ARG(x1 && a && b && (y || y_hi_bf16) && V > 0, "holo_gn_apply: bad args");
    HOLO_CHECK_ARG(C1 % 4 == 0 && C2 % 4 == 0, "holo_gn_apply: channel counts must be multiples of 4");
    HOLO_CHECK_ARG((y_hi_bf16 == nullptr) == (y_lo_bf16 == nullptr), "holo_gn_apply: hi/lo must come together");
    long long total4 = V * (C / 4);
    int blocks = holo_cdiv(total4, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (silu)
        gn_apply_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x1, C1, x2, C2, V, a, b, y, (uint16_t*)y_hi_bf16,
                                                                        (uint16_t*)y_lo_bf16);
    else
        gn_apply_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x1, C1, x2, C2, V, a, b, y,
                                                                         (uint16_t*)y_hi_bf16, (uint16_t*)y_lo_bf16);
    HOLO_CHECK_LAUNCH("holo_gn_apply");
    return HOLO_OK;
}

// fp32 -> bf16 hi/lo split of a plain channels-last tensor (operands that do not come out of a GroupNorm).
// Options: zero-pad the channel count up to Cpad (the tensor-core conv walks K in 64-channel slabs), and fold a
// nearest x2 upsample of the (D,H,W) volume (Upsample.forward, unet.py:94-97) so the upsampled fp32 tensor is
// never written.
__global__ void split_bf16_kernel(const float* __restrict__ x, int C1, const float* __restrict__ x2, int C2,
                                  long long Vout, int Cpad, int ups, int Din, int Hin, int Win,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int pair_f16) {
    const int C = C1 + C2;
    const int q = Cpad / 4;
    const long long total = Vout * q;
    holo_pdl_trigger();   // opt-in PDL build only (common.cuh)
    holo_pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long v = i / q;
        int c = (int)(i % q) * 4;
        long long vs = v;
        if (ups) {
            int Wo = Win * 2, Ho = Hin * 2;
            int ow = (int)(v % Wo), oh = (int)((v / Wo) % Ho), od = (int)(v / ((long long)Wo * Ho));
            vs = ((long long)(od >> 1) * Hin + (oh >> 1)) * Win + (ow >> 1);
        }
        float4 r = make_float4(0, 0, 0, 0);
        if (c < C1) r = *reinterpret_cast<const float4*>(x + vs * C1 + c);
        else if (c < C) r = *reinterpret_cast<const float4*>(x2 + vs * C2 + (c - C1));
        uint32_t h2[2], l2[2];
        holo_split2(r.x, r.y, pair_f16 != 0, h2[0], l2[0]);
        holo_split2(r.z, r.w, pair_f16 != 0, h2[1], l2[1]);
        *reinterpret_cast<uint2*>(hi + v * Cpad + c) = make_uint2(h2[0], h2[1]);
        *reinterpret_cast<uint2*>(lo + v * Cpad + c) = make_uint2(l2[0], l2[1]);
    }
}

extern "C" int holo_split_bf16(const float* x1, int C1, const float* x2, int C2, long long V, int Cpad, int upsample2x,
                               int Din, int Hin, int Win, void* hi_bf16, void* lo, int pair_f16, void* stream) {
    const int C = C1 + C2;
    void* lo_bf16 = lo;
    HOLO_CHECK_ARG(x1 && hi_bf16 && lo_bf16 && V > 0 && C1 > 0 && C1 % 4 == 0 && C2 % 4 == 0 && Cpad % 4 == 0 && Cpad >= C,
                   "holo_split_bf16: channel counts must be multiples of 4, Cpad >= C1 + C2");
    HOLO_CHECK_ARG(C2 == 0 || x2, "holo_split_bf16: second source missing");
    long long Vout = V;
    if (upsample2x) {
        HOLO_CHECK_ARG((long long)Din * Hin * Win == V, "holo_split_bf16: dims do not match V");
        Vout = V * 8;
    }
    int blocks = holo_cdiv(Vout * (Cpad / 4), 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    holo_launch(split_bf16_kernel, dim3(blocks), dim3(256), (size_t)0, (cudaStream_t)stream, x1, C1, x2, C2, Vout, Cpad,
                upsample2x, Din, Hin, Win, (uint16_t*)hi_bf16, (uint16_t*)lo_bf16, pair_f16);
    HOLO_CHECK_LAUNCH("holo_split_bf16");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// timestep embedding + small dense layers (M = batch of timesteps, tiny)
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int n, int dim,
                                          const float* __restrict__ freqs, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int half = dim / 2;
    if (i < n * half) {
        int b = i / half, k = i % half;
        // freqs = th.exp(-log(10000) * arange(half) / half) is built on the host exactly as the reference does
        // (nn.py:119-121, a CPU tensor moved to the device); args = t.float() * freqs
        float arg = (float)t[b] * freqs[k];
        out[b * dim + k] = cosf(arg);
        out[b * dim + half + k] = sinf(arg);
    }
}

extern "C" int holo_timestep_embedding(const long long* t_i64, int n, int dim, const float* freqs, float* out,
                                       void* stream) {
    HOLO_CHECK_ARG(t_i64 && out && freqs && n > 0 && dim > 0 && dim % 2 == 0, "holo_timestep_embedding: bad args");
    timestep_embedding_kernel<<<holo_cdiv(n * dim / 2, 128), 128, 0, (cudaStream_t)stream>>>(t_i64, n, dim, freqs, out);
    HOLO_CHECK_LAUNCH("holo_timestep_embedding");
    return HOLO_OK;
}

// out[m][o] = act_out( b[o] + sum_i W[o][i] * act_in(x[m][i]) ); one warp per (m, o)
__global__ void linear_rows_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                   const float* __restrict__ b, int M, int in_dim, int out_dim, int silu_in,
                                   int silu_out, float* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int lane = threadIdx.x % 32;
    if (warp >= M * out_dim) return;
    int m = warp / out_dim, o = warp % out_dim;
    float acc = 0.f;
    for (int i = lane; i < in_dim; i += 32) {
        float xv = x[(size_t)m * in_dim + i];
        if (silu_in) xv = holo_silu(xv);
        acc = fmaf(W[(size_t)o * in_dim + i], xv, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        acc += b ? b[o] : 0.f;
        out[(size_t)m * out_dim + o] = silu_out ? holo_silu(acc) : acc;
    }
}

extern "C" int holo_linear_rows(const float* x, const float* W, const float* b, int M, int in_dim, int out_dim,
                                int silu_in, int silu_out, float* out, void* stream) {
    HOLO_CHECK_ARG(x && W && out && M > 0 && in_dim > 0 && out_dim > 0, "holo_linear_rows: bad args");
    long long warps = (long long)M * out_dim;
    linear_rows_kernel<<<holo_cdiv(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, W, b, M, in_dim, out_dim,
                                                                                     silu_in, silu_out, out);
    HOLO_CHECK_LAUNCH("holo_linear_rows");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// DDPM ancestral step (START_X / FIXED_SMALL): tables are device fp32 copies of the fp64 schedule,
// indexed by the device-resident timestep (no host round trip, no per-step H2D of the tables).
//   x0 = clamp(model_out, -1, 1);  mean = c1[t]*x0 + c2[t]*x_t;  x_{t-1} = mean + [t!=0]*exp(0.5*logvar[t])*eps
// ------------------------------------------------------------------------------------------------
__global__ void ddpm_step_kernel(const float* __restrict__ model_out, const float* __restrict__ x_t,
                                 const float* __restrict__ noise, const long long* __restrict__ t,
                                 const float* __restrict__ coef1, const float* __restrict__ coef2,
                                 const float* __restrict__ logvar, long long per_sample4, int n_batch, int clip,
                                 float* __restrict__ x_prev, float* __restrict__ pred_x0) {
    long long total = per_sample4 * n_batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int bidx = (int)(i / per_sample4);
        long long tt = t[bidx];
        float c1 = coef1[tt], c2 = coef2[tt];
        float sd = (tt != 0) ? expf(0.5f * logvar[tt]) : 0.f;
        float4 m = reinterpret_cast<const float4*>(model_out)[i];
        float4 x = reinterpret_cast<const float4*>(x_t)[i];
        float4 e = noise ? reinterpret_cast<const float4*>(noise)[i] : make_float4(0, 0, 0, 0);
        if (clip) {
            m.x = fminf(fmaxf(m.x, -1.f), 1.f), m.y = fminf(fmaxf(m.y, -1.f), 1.f);
            m.z = fminf(fmaxf(m.z, -1.f), 1.f), m.w = fminf(fmaxf(m.w, -1.f), 1.f);
        }
        float4 r;
        r.x = (c1 * m.x + c2 * x.x) + sd * e.x;
        r.y = (c1 * m.y + c2 * x.y) + sd * e.y;
        r.z = (c1 * m.z + c2 * x.z) + sd * e.z;
        r.w = (c1 * m.w + c2 * x.w) + sd * e.w;
        reinterpret_cast<float4*>(x_prev)[i] = r;
        if (pred_x0) reinterpret_cast<float4*>(pred_x0)[i] = m;
    }
}

extern "C" int holo_ddpm_step(const float* model_out, const float* x_t, const float* noise, const long long* t_i64,
                              const float* coef1, const float* coef2, const float* logvar, long long per_sample,
                              int n_batch, int clip_denoised, float* x_prev, float* pred_xstart, void* stream) {
    HOLO_CHECK_ARG(model_out && x_t && t_i64 && coef1 && coef2 && logvar && x_prev, "holo_ddpm_step: null arg");
    HOLO_CHECK_ARG(per_sample > 0 && per_sample % 4 == 0 && n_batch > 0, "holo_ddpm_step: per_sample must be a multiple of 4");
    long long total = per_sample / 4 * n_batch;
    int blocks = holo_cdiv(total, 256 * 2);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ddpm_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(model_out, x_t, noise, t_i64, coef1, coef2, logvar,
                                                               per_sample / 4, n_batch, clip_denoised, x_prev,
                                                               pred_xstart);
    HOLO_CHECK_LAUNCH("holo_ddpm_step");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// DDIM step (gaussian_diffusion.py:645-693) and its reverse ODE (:695-731) for a START_X model:
//   x0 = clamp(model_out);  eps = (sqrt_recip_ac[t] x_t - x0) / sqrt_recipm1_ac[t]
//   sigma = eta sqrt((1 - ab_to) / (1 - ab)) sqrt(1 - ab / ab_to)          (ab_to = alphas_cumprod_prev[t])
//   x_{t-1} = x0 sqrt(ab_to) + sqrt(1 - ab_to - sigma^2) eps + [t != 0] sigma noise
// reverse: ab_to = alphas_cumprod_next[t], sigma = 0, no noise.
// ------------------------------------------------------------------------------------------------
__global__ void ddim_step_kernel(const float* __restrict__ model_out, const float* __restrict__ x_t,
                                 const float* __restrict__ noise, const long long* __restrict__ t,
                                 const float* __restrict__ ac, const float* __restrict__ ac_to,
                                 const float* __restrict__ sqrt_recip, const float* __restrict__ sqrt_recipm1,
                                 float eta, long long per_sample, int n_batch, int clip,
                                 float* __restrict__ x_out, float* __restrict__ pred_x0) {
    long long total = per_sample * n_batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long tt = t[i / per_sample];
        float ab = ac[tt], abt = ac_to[tt];
        float sigma = 0.f;
        if (eta != 0.f) sigma = eta * sqrtf((1.f - abt) / (1.f - ab)) * sqrtf(1.f - ab / abt);
        float m = model_out[i];
        if (clip) m = fminf(fmaxf(m, -1.f), 1.f);
        float x = x_t[i];
        float eps = (sqrt_recip[tt] * x - m) / sqrt_recipm1[tt];
        float r = m * sqrtf(abt) + sqrtf(1.f - abt - sigma * sigma) * eps;
        if (noise && tt != 0) r += sigma * noise[i];
        x_out[i] = r;
        if (pred_x0) pred_x0[i] = m;
    }
}

extern "C" int holo_ddim_step(const float* model_out, const float* x_t, const float* noise, const long long* t_i64,
                              const float* alphas_cumprod, const float* alphas_cumprod_to,
                              const float* sqrt_recip_alphas_cumprod, const float* sqrt_recipm1_alphas_cumprod,
                              float eta, long long per_sample, int n_batch, int clip_denoised, float* x_out,
                              float* pred_xstart, void* stream) {
    HOLO_CHECK_ARG(model_out && x_t && t_i64 && alphas_cumprod && alphas_cumprod_to && sqrt_recip_alphas_cumprod &&
                       sqrt_recipm1_alphas_cumprod && x_out,
                   "holo_ddim_step: null arg");
    HOLO_CHECK_ARG(per_sample > 0 && n_batch > 0 && eta >= 0.f, "holo_ddim_step: bad sizes / eta");
    int blocks = holo_cdiv(per_sample * n_batch, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ddim_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(model_out, x_t, noise, t_i64, alphas_cumprod,
                                                               alphas_cumprod_to, sqrt_recip_alphas_cumprod,
                                                               sqrt_recipm1_alphas_cumprod, eta, per_sample, n_batch,
                                                               clip_denoised, x_out, pred_xstart);
    HOLO_CHECK_LAUNCH("holo_ddim_step");
    return HOLO_OK;
}

// q_sample: x_t = sqrt_ac[t]*x0 + sqrt_1m_ac[t]*eps   (gaussian_diffusion.py:209-227)
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const long long* __restrict__ t, const float* __restrict__ sa,
                                const float* __restrict__ s1, long long per_sample, int n_batch,
                                float* __restrict__ out) {
    long long total = per_sample * n_batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long tt = t[i / per_sample];
        out[i] = sa[tt] * x0[i] + s1[tt] * noise[i];
    }
}
extern "C" int holo_q_sample(const float* x0, const float* noise, const long long* t_i64, const float* sqrt_ac,
                             const float* sqrt_1m_ac, long long per_sample, int n_batch, float* out, void* stream) {
    HOLO_CHECK_ARG(x0 && noise && t_i64 && sqrt_ac && sqrt_1m_ac && out && per_sample > 0 && n_batch > 0, "holo_q_sample: bad args");
    int blocks = holo_cdiv(per_sample * n_batch, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    q_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x0, noise, t_i64, sqrt_ac, sqrt_1m_ac, per_sample, n_batch, out);
    HOLO_CHECK_LAUNCH("holo_q_sample");
    return HOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Elementwise activation + range statistics, both layouts written in one pass.
//   act: 0 none, 1 tanh, 2 clamp[-1,1].  stats (device, 4 ints): ordered-int min, ordered-int max, nan count, unused.
// Input is channels-last (V,C); outputs: channels-last (for the renderer) and/or channels-first (for the API).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

__global__ void act_range_kernel(const float* __restrict__ x, long long V, int C, int act,
                                 float* __restrict__ y_cl, float* __restrict__ y_cf, int* __restrict__ stats) {
    __shared__ float tile[32][33];
    __shared__ float s_mn[8], s_mx[8];
    __shared__ int s_nan[8];
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    // tile: 32 voxels x 32 channels; blockDim (32, 8); every block walks several voxel tiles (grid-stride)
    for (long long v0 = (long long)blockIdx.x * 32; v0 < V; v0 += (long long)gridDim.x * 32) {
        for (int c0 = 0; c0 < C; c0 += 32) {
            for (int i = threadIdx.y; i < 32; i += 8) {
                long long v = v0 + i;
                int c = c0 + threadIdx.x;
                if (v < V && c < C) {
                    float a = x[v * C + c];
                    if (act == 1) a = tanhf(a);
                    else if (act == 2) a = fminf(fmaxf(a, -1.f), 1.f);
                    if (a != a) nan++;
                    mn = fminf(mn, a), mx = fmaxf(mx, a);
                    if (y_cl) y_cl[v * C + c] = a;
                    tile[i][threadIdx.x] = a;
                }
            }
            __syncthreads();
            if (y_cf) {
                for (int i = threadIdx.y; i < 32; i += 8) {
                    int c = c0 + i;
                    long long v = v0 + threadIdx.x;
                    if (v < V && c < C) y_cf[(size_t)c * V + v] = tile[threadIdx.x][i];
                }
            }
            __syncthreads();
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            nan += __shfl_xor_sync(0xffffffffu, nan, o);
        }
        if (threadIdx.x == 0) s_mn[threadIdx.y] = mn, s_mx[threadIdx.y] = mx, s_nan[threadIdx.y] = nan;
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {  // one set of atomics per block
            for (int w = 1; w < 8; ++w) mn = fminf(mn, s_mn[w]), mx = fmaxf(mx, s_mx[w]), nan += s_nan[w];
            if (mn != INFINITY) atomicMin(&stats[0], float_to_ordered(mn));
            if (mx != -INFINITY) atomicMax(&stats[1], float_to_ordered(mx));
            if (nan) atomicAdd(&stats[2], nan);
        }
    }
}

__global__ void range_init_kernel(int* stats) {
    stats[0] = 0x7fffffff;            // +max ordered
    stats[1] = (int)0x80000000;       // min ordered
    stats[2] = 0;
    stats[3] = 0;
}

extern "C" int holo_range_init(int* stats4, void* stream) {
    HOLO_CHECK_ARG(stats4, "holo_range_init: null");
    range_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats4);
    HOLO_CHECK_LAUNCH("holo_range_init");
    return HOLO_OK;
}

extern "C" int holo_act_range(const float* x_cl, long long V, int C, int act, float* y_cl, float* y_cf, int* stats4,
                              void* stream) {
    HOLO_CHECK_ARG(x_cl && V > 0 && C > 0 && act >= 0 && act <= 2, "holo_act_range: bad args");
    int blocks = holo_cdiv(V, 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    act_range_kernel<<<blocks, dim3(32, 8), 0, (cudaStream_t)stream>>>(x_cl, V, C, act, y_cl, y_cf, stats4);
    HOLO_CHECK_LAUNCH("holo_act_range");
    return HOLO_OK;
}
