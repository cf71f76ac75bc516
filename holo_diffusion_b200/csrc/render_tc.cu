// Fused volumetric renderer, tensor-core version (sm_100a): the 256-wide hidden layer of the collapsed RenderMLP
// runs on tcgen05 (M = 128 rays of one depth step, N = 256 hidden units, K = 32 features as a 3xBF16 split),
// everything else of a ray -- trilinear gather, density, emission-absorption compositing, importance
// re-sampling, second pass -- stays in registers of the thread that owns the ray.
//
// Same contract and reference citations as render.cu (holo_render_fwd); this kernel is selected for C in {16, 32}
// with the shipped 256-wide density net.  Per depth step and group of 128 rays:
//   ray threads : sample x (fp32) -> sigma (fp32 CUDA cores) -> write [x_hi | x_lo] as one 128-byte swizzled row
//   MMA warp    : D[128 x 256] = [x_hi|x_lo].[W_hi|W_hi]^T + x_hi.W_lo^T + 1.[b_hi + b_lo]^T      (7 UMMAs, TMEM)
//   ray threads : tcgen05.ld the row, rgb_pre = lin(x) + sum_j 0.4 wr_j |t_j|   (leaky(t) = 0.6 t + 0.4 |t|, the
//                 linear part 0.6 wr.(W x + b) is a 3 x C map folded at pack time), sigmoid, compositing.
// Two groups per CTA ping-pong so one group's MMA hides behind the other's CUDA-core work.  The refiner is
// streamed: the coarse weights go to a [S][n_rays] scratch, the inverse-cdf samples are generated in increasing
// order and merged with the coarse depths on the fly (both runs are sorted), so no per-ray arrays are needed.
#include "common.cuh"
#include "../../include/holo_b200.h"
#include <cuda_bf16.h>

namespace {

constexpr int HID = 256;
constexpr int GROUP = 128;            // rays per group (= UMMA M)
constexpr int RAY_THREADS = 2 * GROUP;      // rays per CTA
// warp roles: [0, 256) samplers (thread = ray), [256, 512) epilogue/compositing (thread = ray), [512, 576) MMA issuers
constexpr int THREADS = 2 * RAY_THREADS + 128;  // warps 16,17 issue MMAs; 18,19 only complete the warpgroup for setmaxnreg

// shared-memory image (byte offsets; every UMMA tile 1024-aligned)
constexpr int OFF_A0 = 0;                 // [128][128B]  group 0 operand rows
constexpr int OFF_A1 = 16 * 1024;         // group 1
constexpr int OFF_B1 = 32 * 1024;         // [256][128B]  [W_hi | W_hi]
constexpr int OFF_B2 = 64 * 1024;         // [256][128B]  [W_lo | 0]
constexpr int OFF_ONES = 96 * 1024;       // [128][128B]  [1, 1, 0, ...]
constexpr int OFF_BB = 112 * 1024;        // [256][128B]  [b_hi, b_lo, 0, ...]
constexpr int OFF_F32 = 144 * 1024;       // fp32 parameter block (see pack kernel)
constexpr int F32_WSIG = 0;               // [32] density row of W_eff
constexpr int F32_LIN = 32;               // [3][32]  0.6 * wr . W_eff[:256]
constexpr int F32_MISC = 128;             // b_sigma, lin_const[3]
constexpr int F32_EP = 132;               // [256][3] 0.4 * wr[i][j] stored j-major
constexpr int F32_DIR = 132 + 768;        // [3][E] + br[3]   (E <= 27)
constexpr int F32_COUNT = F32_DIR + 3 * 27 + 3;
constexpr int IMG_BYTES = OFF_F32 + ((F32_COUNT * 4 + 15) / 16) * 16;
constexpr int OFF_A2 = ((IMG_BYTES + 1023) / 1024) * 1024;  // second operand buffer of group 0 / 1 (double buffering)
constexpr int OFF_A3 = OFF_A2 + 16 * 1024;
constexpr int OFF_PR = OFF_A3 + 16 * 1024;   // [group][buf][6][128] floats: sigma, lin0..2, z, z_next handed S -> E
constexpr int PR_FLOATS = 6 * GROUP;
constexpr int OFF_WTOT = OFF_PR + 2 * 2 * PR_FLOATS * 4;  // [group][128] cdf normaliser handed E -> S between passes
constexpr int CI_STRIDE = 20;                              // floats per ray: 8 corner voxel indices + 8 weights (+ pad)
constexpr int OFF_CI = OFF_WTOT + 2 * GROUP * 4;           // [8 sampler warps][32 rays][CI_STRIDE]
constexpr int OFF_BAR = OFF_CI + 8 * 32 * CI_STRIDE * 4;   // mbarriers + tmem slot
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    // one leader lane; unlike `lane == 0` the compiler keeps descriptors in uniform registers (no R2UR waterfall
    // loop around every UTCHMMA / UTMALDG)
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld and pin the register uses behind it (zero-instruction "launder" asms)
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i += 8)
        asm volatile("" : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]),
                          "+r"(r[i + 6]), "+r"(r[i + 7]));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ float linspace01(int i, int n) {
    if (n == 1) return 0.0f;
    float step = 1.0f / (float)(n - 1);
    return (i < n / 2) ? step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

struct RenderTcParams {
    const float* grid;
    int D, Hh, Ww, C;
    float inv_x, inv_y, inv_z;
    const uint8_t* image;  // packed smem image (IMG_BYTES)
    int n_harm;
    const float* origins;
    const float* dirs;
    const float* lengths;
    int n_rays, S, n_fine, add_input, n_passes;
    float bg[3];
    float bg_opacity;
    float* features;
    float* depths;
    float* masks;
    float* weights;
    float* lengths_out;
    float* p_features;
    float* p_depths;
    float* p_masks;
    float* p_weights;
    float* scratch_w;  // [S][n_rays] coarse weights (two passes only)
};

template <int C>
__global__ void __launch_bounds__(THREADS, 1) render_tc_kernel(RenderTcParams P) {
    // Warp-specialised pipeline per group of 128 rays (two groups per CTA):
    //   sampler warps  : depth stream (coarse / merged fine) -> trilinear gather -> sigma, linear radiance part ->
    //                    operand row [x_hi | x_lo] into A[buf], scalars into PR[buf]          (step s + 1)
    //   MMA warp       : 7 UMMAs into the group's TMEM accumulator                            (step s)
    //   epilogue warps : TMEM row -> radiance -> emission-absorption compositing, outputs     (step s - 1)
    // so the three stages of consecutive depth steps overlap and 16 ray warps (instead of 8) hide each other's
    // gather / TMEM / barrier latencies.
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    // per group g: a_full[2], pr_free[2], tm_full, tm_free, pass_done
    auto a_full = [&](int g, int b) { return bars + g * 8 + b; };
    auto pr_free = [&](int g, int b) { return bars + g * 8 + 2 + b; };
    auto tm_full = [&](int g) { return bars + g * 8 + 4; };
    auto tm_free = [&](int g) { return bars + g * 8 + 5; };
    auto pass_done = [&](int g) { return bars + g * 8 + 6; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const float* sF = reinterpret_cast<const float*>(smem + OFF_F32);
    float* sPR = reinterpret_cast<float*>(smem + OFF_PR);
    float* sWtot = reinterpret_cast<float*>(smem + OFF_WTOT);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.image) + (OFF_B1 / 16);
        uint4* dst = reinterpret_cast<uint4*>(smem + OFF_B1);
        for (int i = tid; i < (IMG_BYTES - OFF_B1) / 16; i += THREADS) dst[i] = src[i];
    }
    {   // operand buffers start zeroed: for C = 16 the chunks of channels 16..31 are never written
        uint4* z0 = reinterpret_cast<uint4*>(smem + OFF_A0);
        uint4* z1 = reinterpret_cast<uint4*>(smem + OFF_A2);
        for (int i = tid; i < 2 * 16 * 1024 / 16; i += THREADS) z0[i] = make_uint4(0, 0, 0, 0), z1[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int g = 0; g < 2; ++g) {
            mbar_init(a_full(g, 0), GROUP), mbar_init(a_full(g, 1), GROUP);
            mbar_init(pr_free(g, 0), GROUP), mbar_init(pr_free(g, 1), GROUP);
            mbar_init(tm_full(g), 1), mbar_init(tm_free(g), GROUP), mbar_init(pass_done(g), GROUP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged B tiles are read by the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int S1 = P.S;
    const int S2 = P.add_input ? P.S + P.n_fine : P.n_fine;
    const int total_steps = S1 + (P.n_passes > 1 ? S2 : 0);
    const float eps = 1e-5f;

    // register re-balancing (per warpgroup of 4 warps).  The CTA is launched with 96 registers x 640 threads; the MMA
    // warpgroup drops to 24 and releases 4 x 32 x 72 = 9216 registers, which lets the 16 ray warps grow by
    // 9216 / 512 = 18 -> 112 (a request the pool cannot cover would block forever).
    if (warp < 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");

    if (warp >= 18) {
        // idle filler warps of the MMA warpgroup
    } else if (warp >= 16) {
        // ============================ MMA issuer of group g ============================
        const int g = warp - 16;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(GROUP >> 4) << 24);
        const uint32_t a0 = smem_u32(smem + (g ? OFF_A1 : OFF_A0)), a1 = smem_u32(smem + (g ? OFF_A3 : OFF_A2));
        const uint32_t b1 = smem_u32(smem + OFF_B1), b2 = smem_u32(smem + OFF_B2);
        const uint32_t on = smem_u32(smem + OFF_ONES), bb = smem_u32(smem + OFF_BB);
        const uint32_t d = tmem_base + (uint32_t)(g * HID);
        for (int gs = 0; gs < total_steps; ++gs) {
            const int buf = gs & 1;
            mbar_wait(a_full(g, buf), (gs >> 1) & 1);
            if (gs >= 1) mbar_wait(tm_free(g), (gs - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t a = buf ? a1 : a0;
                umma(d, sw128_desc(on), sw128_desc(bb), idesc, 0);                       // bias (starts the tile)
                umma(d, sw128_desc(a), sw128_desc(b2), idesc, 1);                        // x_hi . W_lo, K 0..15
                umma(d, sw128_desc(a + 32), sw128_desc(b2 + 32), idesc, 1);              //              K 16..31
#pragma unroll
                for (int k = 0; k < 4; ++k)                                              // [x_hi|x_lo] . [W_hi|W_hi]
                    umma(d, sw128_desc(a + 32 * k), sw128_desc(b1 + 32 * k), idesc, 1);
                umma_commit(tm_full(g));
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ============================ samplers ============================
        // Lane l OWNS ray l of the warp (depth stream), but the gather is warp-cooperative: QPR = C/4 lanes share one
        // ray, each loading its own 16-byte channel quad of the 8 corner lines.  A warp-level LDG.128 then touches
        // 32/QPR lines instead of 32 (the thread-per-ray gather spent ~2000 L1 wavefronts per warp and step -- it, not
        // the MLP, bounded the kernel), and every lane ends up holding 4 channels of the sampled feature, which is
        // exactly the granularity of the operand row pieces, so no transposition is needed.
        constexpr int QPR = C / 4;        // lanes per ray
        constexpr int RPI = 32 / QPR;     // rays per cooperative iteration
        const int g = warp / 4;
        const int row = tid - g * GROUP;  // row owned by this lane
        const int q = lane % QPR, sub = lane / QPR;
        const long long ray_raw = (long long)blockIdx.x * RAY_THREADS + tid;
        const size_t ray = ray_raw < P.n_rays ? (size_t)ray_raw : (size_t)(P.n_rays - 1);  // idle lanes shadow the last ray
        uint8_t* a_base0 = smem + (g ? OFF_A1 : OFF_A0);
        uint8_t* a_base1 = smem + (g ? OFF_A3 : OFF_A2);
        const float b_sigma = sF[F32_MISC];
        // this lane's slice of the density row and of the linear radiance map (channels 4q .. 4q+3)
        const float4 ws = *reinterpret_cast<const float4*>(sF + F32_WSIG + 4 * q);
        const float4 la = *reinterpret_cast<const float4*>(sF + F32_LIN + 4 * q);
        const float4 lb = *reinterpret_cast<const float4*>(sF + F32_LIN + 32 + 4 * q);
        const float4 lc = *reinterpret_cast<const float4*>(sF + F32_LIN + 64 + 4 * q);
        float o[3], d[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) o[i] = P.origins[ray * 3 + i], d[i] = P.dirs[ray * 3 + i];
        const float* zin = P.lengths + ray * S1;
        const int D = P.D, H = P.Hh, W = P.Ww;
        float* ci_warp = reinterpret_cast<float*>(smem + OFF_CI) + (size_t)warp * 32 * CI_STRIDE;
        int gs = 0;
        for (int pass = 0; pass < P.n_passes; ++pass) {
            const int S = pass == 0 ? S1 : S2;
            // ---- depth stream: coarse depths, or the on-the-fly merge of coarse depths and inverse-cdf samples
            int ic = 0, jf = 0, inds = 0;
            float c_prev = 0.f, c_cur = 0.f, f_next = INFINITY, w_tot = 1.f;
            const int ncdf = S1 - 1;
            const float* wsc = P.scratch_w + ray;  // column of coarse weights, stride n_rays
            if (pass == 1) {
                mbar_wait(pass_done(g), 0);        // the epilogue threads finished the coarse pass (weights, w_tot)
                w_tot = sWtot[g * GROUP + row];
            }
            auto gen_fine = [&]() {               // next inverse-cdf sample (sample_pdf, deterministic u)
                if (jf >= P.n_fine) {
                    f_next = INFINITY;
                    return;
                }
                float u = linspace01(jf, P.n_fine);
                while (inds < ncdf && c_cur <= u) {  // searchsorted(cdf, u, right=True)
                    c_prev = c_cur;
                    ++inds;
                    if (inds < ncdf) c_cur = c_prev + (wsc[(size_t)inds * P.n_rays] + eps) / w_tot;
                }
                int below = max(inds - 1, 0), above = min(inds, ncdf - 1);
                float c0 = c_prev, c1 = (inds < ncdf) ? c_cur : c_prev;
                float b0 = 0.5f * (zin[below + 1] + zin[below]);
                float b1 = 0.5f * (zin[above + 1] + zin[above]);
                float den = c1 - c0;
                if (den < eps) den = 1.f;
                float t = (u - c0) / den;
                f_next = b0 + t * (b1 - b0);
            };
            float z_last = -INFINITY;
            auto next_z = [&]() -> float {
                float z;
                if (pass == 0) {
                    z = zin[min(ic, S1 - 1)];
                    ++ic;
                } else {
                    float zc = (P.add_input && ic < S1) ? zin[ic] : INFINITY;
                    if (zc <= f_next) {
                        z = zc;
                        ++ic;
                    } else {
                        z = f_next;
                        ++jf;
                        gen_fine();
                    }
                    z = fmaxf(z, z_last);  // the two runs are sorted up to 1-ulp ties (torch.sort in the refiner)
                }
                z_last = z;
                return z;
            };
            if (pass == 1) gen_fine();
            float z_cur = next_z();
            float z_nxt = (S > 1) ? next_z() : z_cur;
            for (int s = 0; s < S; ++s, ++gs) {
                const int buf = gs & 1;
                // ---- owner lane: corner voxel indices (clamped) and trilinear weights of its point, once per step
                {
                    const float lx = (o[0] + z_cur * d[0]) * P.inv_x, ly = (o[1] + z_cur * d[1]) * P.inv_y,
                                lz = (o[2] + z_cur * d[2]) * P.inv_z;
                    // ATen grid_sampler_3d, bilinear, zeros padding, align_corners=True
                    const float ix = ((lx + 1.f) / 2.f) * (float)(W - 1);
                    const float iy = ((ly + 1.f) / 2.f) * (float)(H - 1);
                    const float iz = ((lz + 1.f) / 2.f) * (float)(D - 1);
                    const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
                    const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)W + 1.f);
                    const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)H + 1.f);
                    const int z0 = (int)fminf(fmaxf(fz0, -2.f), (float)D + 1.f);
                    const float wx1 = ix - fx0, wy1 = iy - fy0, wz1 = iz - fz0;
                    const float wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy, wz0 = (fz0 + 1.f) - iz;
                    int idx[8];
                    float wt[8];
#pragma unroll
                    for (int corner = 0; corner < 8; ++corner) {
                        const int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
                        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
                        const float w = (dx ? wx1 : wx0) * (dy ? wy1 : wy0) * (dz ? wz1 : wz0);
                        const bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D;
                        wt[corner] = ok ? w : 0.f;   // zeros padding: clamped address, weight 0 (0 * finite = 0 exactly)
                        const int xc = min(max(xx, 0), W - 1), yc = min(max(yy, 0), H - 1), zc = min(max(zz, 0), D - 1);
                        idx[corner] = (zc * H + yc) * W + xc;
                    }
                    float* ci = ci_warp + lane * CI_STRIDE;
                    *reinterpret_cast<int4*>(ci + 0) = make_int4(idx[0], idx[1], idx[2], idx[3]);
                    *reinterpret_cast<int4*>(ci + 4) = make_int4(idx[4], idx[5], idx[6], idx[7]);
                    *reinterpret_cast<float4*>(ci + 8) = make_float4(wt[0], wt[1], wt[2], wt[3]);
                    *reinterpret_cast<float4*>(ci + 12) = make_float4(wt[4], wt[5], wt[6], wt[7]);
                }
                __syncwarp();
                // the slot must have been drained by the epilogue of step gs - 2 before anything is written
                if (gs >= 2) mbar_wait(pr_free(g, buf), ((gs >> 1) - 1) & 1);
                uint8_t* a_base = buf ? a_base1 : a_base0;
                float* prb = sPR + (size_t)(g * 2 + buf) * PR_FLOATS;
#pragma unroll 2
                for (int it = 0; it < QPR; ++it) {
                    const int rl = it * RPI + sub;          // lane that owns the ray handled now
                    const float* ci = ci_warp + rl * CI_STRIDE;
                    const int4 i0 = *reinterpret_cast<const int4*>(ci + 0), i1 = *reinterpret_cast<const int4*>(ci + 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(ci + 8), w1 = *reinterpret_cast<const float4*>(ci + 12);
                    const int idx[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
                    const float wt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    float4 v[8];
#pragma unroll
                    for (int corner = 0; corner < 8; ++corner)   // 8 coalesced 16-byte gathers in flight
                        v[corner] = __ldg(reinterpret_cast<const float4*>(P.grid + (size_t)idx[corner] * C) + q);
                    float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int corner = 0; corner < 8; ++corner) {
                        x4.x = fmaf(v[corner].x, wt[corner], x4.x), x4.y = fmaf(v[corner].y, wt[corner], x4.y);
                        x4.z = fmaf(v[corner].z, wt[corner], x4.z), x4.w = fmaf(v[corner].w, wt[corner], x4.w);
                    }
                    // partial dot products over this lane's 4 channels ...
                    float pv[4];
                    pv[0] = fmaf(ws.x, x4.x, fmaf(ws.y, x4.y, fmaf(ws.z, x4.z, ws.w * x4.w)));
                    pv[1] = fmaf(la.x, x4.x, fmaf(la.y, x4.y, fmaf(la.z, x4.z, la.w * x4.w)));
                    pv[2] = fmaf(lb.x, x4.x, fmaf(lb.y, x4.y, fmaf(lb.z, x4.z, lb.w * x4.w)));
                    pv[3] = fmaf(lc.x, x4.x, fmaf(lc.y, x4.y, fmaf(lc.z, x4.z, lc.w * x4.w)));
                    // ... reduce-scattered over the QPR lanes of the ray: after the butterfly each lane pair holds ONE
                    // complete sum (4 + 2 [+ 1] shuffles instead of 12)
                    int vidx = 0;
                    {
                        const bool up = (q & (QPR / 2)) != 0;
                        const float s0 = up ? pv[0] : pv[2], s1 = up ? pv[1] : pv[3];
                        float k0 = up ? pv[2] : pv[0], k1 = up ? pv[3] : pv[1];
                        k0 += __shfl_xor_sync(0xffffffffu, s0, QPR / 2);
                        k1 += __shfl_xor_sync(0xffffffffu, s1, QPR / 2);
                        vidx = up ? 2 : 0;
                        const bool up2 = (q & (QPR / 4)) != 0;
                        const float s2 = up2 ? k0 : k1;
                        float k = up2 ? k1 : k0;
                        k += __shfl_xor_sync(0xffffffffu, s2, QPR / 4);
                        vidx += up2 ? 1 : 0;
                        if (QPR == 8) k += __shfl_xor_sync(0xffffffffu, k, 1);
                        pv[0] = k;
                    }
                    // operand row pieces: [x_hi (chunks 0..3) | x_lo (chunks 4..7)], this lane owns 8 bytes of each
                    const int R = (warp % 4) * 32 + rl;   // row of the group's M tile
                    const float h0 = __bfloat162float(__float2bfloat16_rn(x4.x)), h1 = __bfloat162float(__float2bfloat16_rn(x4.y));
                    const float h2 = __bfloat162float(__float2bfloat16_rn(x4.z)), h3 = __bfloat162float(__float2bfloat16_rn(x4.w));
                    const uint2 hi = make_uint2(pack_bf16x2(x4.x, x4.y), pack_bf16x2(x4.z, x4.w));
                    const uint2 lo = make_uint2(pack_bf16x2(x4.x - h0, x4.y - h1), pack_bf16x2(x4.z - h2, x4.w - h3));
                    uint8_t* arow = a_base + (size_t)(R / 8) * 1024 + (R % 8) * 128;
                    const int swz = R % 8;
                    *reinterpret_cast<uint2*>(arow + (((q >> 1) ^ swz) * 16) + (q & 1) * 8) = hi;
                    *reinterpret_cast<uint2*>(arow + (((4 + (q >> 1)) ^ swz) * 16) + (q & 1) * 8) = lo;
                    if (QPR == 4 || (q & 1) == 0)
                        prb[vidx * GROUP + R] = vidx == 0 ? holo_leaky(b_sigma + pv[0]) : pv[0];
                }
                __syncwarp();   // the corner scratch is rewritten by the owners in the next step
                prb[4 * GROUP + row] = z_cur, prb[5 * GROUP + row] = z_nxt;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(a_full(g, buf));
                z_cur = z_nxt;
                z_nxt = (s + 2 < S) ? next_z() : z_cur;
            }
        }
    } else {
        // ============================ epilogue + compositing ============================
        const int et = tid - RAY_THREADS;
        const int g = et / GROUP;
        const int row = et - g * GROUP;
        const long long ray_raw = (long long)blockIdx.x * RAY_THREADS + et;
        const bool valid = ray_raw < P.n_rays;
        const size_t ray = valid ? (size_t)ray_raw : (size_t)(P.n_rays - 1);
        const uint32_t taddr = tmem_base + ((uint32_t)((warp % 4) * 32) << 16) + (uint32_t)(g * HID);
        float dn[3];
        {
            float d0 = P.dirs[ray * 3], d1 = P.dirs[ray * 3 + 1], d2 = P.dirs[ray * 3 + 2];
            float nrm = fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-12f);
            dn[0] = d0 / nrm, dn[1] = d1 / nrm, dn[2] = d2 / nrm;
        }
        // per-ray constant of the radiance pre-activation: br + Wr[:, H:] . PE(dn) + 0.6 * wr . b_eff
        float rd[3];
        {
            const int nh = P.n_harm, E = 3 * (2 * nh + 1);
            const float* sDir = sF + F32_DIR;
            const float* br = sDir + 3 * E;
            rd[0] = br[0] + sF[F32_MISC + 1], rd[1] = br[1] + sF[F32_MISC + 2], rd[2] = br[2] + sF[F32_MISC + 3];
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
        int gs = 0;
        float w_tot = 0.f;  // sum_{i=1}^{S1-2} (w_i + eps), accumulated during the coarse pass
        for (int pass = 0; pass < P.n_passes; ++pass) {
            const int S = pass == 0 ? S1 : S2;
            const bool last = pass == P.n_passes - 1;
            float cum = 0.f, opac_prev = 0.f, af[3] = {0.f, 0.f, 0.f}, ad = 0.f;
            float* wout = last ? P.weights : P.p_weights;
            float* lout = (pass == 1) ? P.lengths_out : nullptr;
            for (int s = 0; s < S; ++s, ++gs) {
                const int buf = gs & 1;
                mbar_wait(a_full(g, buf), (gs >> 1) & 1);   // direct S -> E ordering for the PR scalars
                const float* pr = sPR + (size_t)(g * 2 + buf) * PR_FLOATS + row;
                const float sig = pr[0];
                float r0 = rd[0] + pr[1 * GROUP], r1 = rd[1] + pr[2 * GROUP], r2 = rd[2] + pr[3 * GROUP];
                const float z_s = pr[4 * GROUP], z_n = pr[5 * GROUP];
                mbar_wait(tm_full(g), gs & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float q0 = 0.f, q1 = 0.f, q2 = 0.f;  // second set of accumulators: 6 independent FMA chains
                {
                    uint32_t ta[32], tb[32];
                    auto consume = [&](const uint32_t (&t)[32], int j0) {
                        const float4* ep = reinterpret_cast<const float4*>(sF + F32_EP + j0 * 3);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {  // 4 hidden units = 12 coefficients = 3 float4
                            float4 e0 = ep[q * 3 + 0], e1 = ep[q * 3 + 1], e2 = ep[q * 3 + 2];
                            float t0 = fabsf(__uint_as_float(t[q * 4 + 0])), t1 = fabsf(__uint_as_float(t[q * 4 + 1]));
                            float t2 = fabsf(__uint_as_float(t[q * 4 + 2])), t3 = fabsf(__uint_as_float(t[q * 4 + 3]));
                            r0 = fmaf(e0.x, t0, r0), r1 = fmaf(e0.y, t0, r1), r2 = fmaf(e0.z, t0, r2);
                            q0 = fmaf(e0.w, t1, q0), q1 = fmaf(e1.x, t1, q1), q2 = fmaf(e1.y, t1, q2);
                            r0 = fmaf(e1.z, t2, r0), r1 = fmaf(e1.w, t2, r1), r2 = fmaf(e2.x, t2, r2);
                            q0 = fmaf(e2.y, t3, q0), q1 = fmaf(e2.z, t3, q1), q2 = fmaf(e2.w, t3, q2);
                        }
                    };
                    tmem_ld32_async(taddr, ta);
#pragma unroll 1
                    for (int j0 = 0; j0 < HID; j0 += 64) {  // the load of the next 32 columns overlaps the FMAs
                        tmem_ld_wait(ta);
                        tmem_ld32_async(taddr + (uint32_t)(j0 + 32), tb);
                        consume(ta, j0);
                        tmem_ld_wait(tb);
                        if (j0 + 64 < HID) tmem_ld32_async(taddr + (uint32_t)(j0 + 64), ta);
                        consume(tb, j0 + 32);
                    }
                }
                // accumulator and scalar slot are free: release them before the (long) compositing math
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(tm_free(g));
                mbar_arrive(pr_free(g, buf));
                // ex2.approx-based exp (rel. error ~2^-22, far inside the 1e-4 parity bar) for the three sigmoids; the
                // opacity terms below keep the full-precision expf because 1 - exp(-x) cancels for thin samples
                float rgb0 = __fdividef(1.f, 1.f + __expf(-holo_leaky(r0 + q0)));
                float rgb1 = __fdividef(1.f, 1.f + __expf(-holo_leaky(r1 + q1)));
                float rgb2 = __fdividef(1.f, 1.f + __expf(-holo_leaky(r2 + q2)));
                float delta = (s + 1 < S) ? (z_n - z_s) : P.bg_opacity;
                float wd = delta * fmaxf(sig, 0.f);
                float capped = 1.f - expf(-wd);
                cum += wd;
                float opac = 1.f - expf(-cum);
                float absorb = (s == 0) ? 1.f : (1.f - opac_prev);
                float w = capped * absorb;
                af[0] += w * rgb0, af[1] += w * rgb1, af[2] += w * rgb2;
                ad += w * z_s;
                opac_prev = opac;
                if (valid) {
                    if (wout) wout[ray * S + s] = w;
                    if (lout) lout[ray * S + s] = z_s;
                    if (!last) P.scratch_w[(size_t)s * P.n_rays + ray] = w;
                }
                if (!last && s >= 1 && s <= S1 - 2) w_tot += w + eps;
            }
            if (valid) {
                float mask = opac_prev;
                float* F = last ? P.features : P.p_features;
                float* Dp = last ? P.depths : P.p_depths;
                float* M = last ? P.masks : P.p_masks;
                if (F) {
                    F[ray * 3 + 0] = af[0] + (1.f - mask) * P.bg[0];
                    F[ray * 3 + 1] = af[1] + (1.f - mask) * P.bg[1];
                    F[ray * 3 + 2] = af[2] + (1.f - mask) * P.bg[2];
                }
                if (Dp) Dp[ray] = ad;
                if (M) M[ray] = mask;
            }
            if (!last) {
                sWtot[g * GROUP + row] = w_tot;
                __threadfence_block();
                mbar_arrive(pass_done(g));   // release: the coarse weights (global scratch) and w_tot are visible
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// ---------------------------------------------------------------------------------------------------------
// pack: collapsed density net (fp64) + radiance layer -> the shared-memory image above
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t sw128_off(int row, int byte_in_row) {
    int chunk = byte_in_row / 16, within = byte_in_row % 16;
    return (size_t)(row / 8) * 1024 + (row % 8) * 128 + ((chunk ^ (row % 8)) * 16) + within;
}

__global__ void pack_render_tc_fill_kernel(const double* __restrict__ A, const double* __restrict__ c,
                                           const float* __restrict__ Wr, const float* __restrict__ br, int C, int E,
                                           uint8_t* __restrict__ img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    __nv_bfloat16* b1 = reinterpret_cast<__nv_bfloat16*>(img + OFF_B1);
    __nv_bfloat16* b2 = reinterpret_cast<__nv_bfloat16*>(img + OFF_B2);
    __nv_bfloat16* on = reinterpret_cast<__nv_bfloat16*>(img + OFF_ONES);
    __nv_bfloat16* bb = reinterpret_cast<__nv_bfloat16*>(img + OFF_BB);
    float* f32 = reinterpret_cast<float*>(img + OFF_F32);
    for (int i = tid; i < HID * 32; i += nth) {
        int j = i / 32, k = i % 32;
        float w = (k < C) ? (float)A[(size_t)j * C + k] : 0.f;
        __nv_bfloat16 h = __float2bfloat16_rn(w);
        __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        b1[sw128_off(j, k * 2) / 2] = h;
        b1[sw128_off(j, (32 + k) * 2) / 2] = h;
        b2[sw128_off(j, k * 2) / 2] = l;
    }
    for (int j = tid; j < HID; j += nth) {
        float b = (float)c[j];
        __nv_bfloat16 h = __float2bfloat16_rn(b);
        bb[sw128_off(j, 0) / 2] = h;
        bb[sw128_off(j, 2) / 2] = __float2bfloat16_rn(b - __bfloat162float(h));
        // 0.4 * wr[i][j], j-major triples
        f32[F32_EP + j * 3 + 0] = 0.4f * Wr[0 * (HID + E) + j];
        f32[F32_EP + j * 3 + 1] = 0.4f * Wr[1 * (HID + E) + j];
        f32[F32_EP + j * 3 + 2] = 0.4f * Wr[2 * (HID + E) + j];
    }
    for (int r = tid; r < GROUP; r += nth) {
        on[sw128_off(r, 0) / 2] = __float2bfloat16_rn(1.f);
        on[sw128_off(r, 2) / 2] = __float2bfloat16_rn(1.f);
    }
    for (int k = tid; k < 32; k += nth) f32[F32_WSIG + k] = (k < C) ? (float)A[(size_t)HID * C + k] : 0.f;
    for (int i = tid; i < 3 * 32; i += nth) {
        int ch = i / 32, k = i % 32;
        double acc = 0.0;
        if (k < C)
            for (int j = 0; j < HID; ++j) acc += (double)Wr[ch * (HID + E) + j] * A[(size_t)j * C + k];
        f32[F32_LIN + i] = (float)(0.6 * acc);
    }
    if (tid < 3) {
        double acc = 0.0;
        for (int j = 0; j < HID; ++j) acc += (double)Wr[tid * (HID + E) + j] * c[j];
        f32[F32_MISC + 1 + tid] = (float)(0.6 * acc);
        f32[F32_DIR + 3 * E + tid] = br[tid];
    }
    if (tid == 0) f32[F32_MISC] = (float)c[HID];
    for (int i = tid; i < 3 * E; i += nth) f32[F32_DIR + i] = Wr[(i / E) * (HID + E) + HID + (i % E)];
}

}  // namespace

extern "C" long long holo_render_tc_image_bytes(void) { return IMG_BYTES; }

extern "C" int holo_pack_render_mlp_tc(const double* A_eff, const double* c_eff, const float* Wr, const float* br, int H,
                                       int C, int E, void* image, void* stream) {
    HOLO_CHECK_ARG(A_eff && c_eff && Wr && br && image, "holo_pack_render_mlp_tc: null arg");
    if (H != HID || !(C == 16 || C == 32) || E > 27 || E < 3) {
        holo_set_error("holo_pack_render_mlp_tc: needs hidden=256, C in {16,32}, E<=27 (got %d, %d, %d)", H, C, E);
        return HOLO_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HOLO_CUDA(cudaMemsetAsync(image, 0, IMG_BYTES, st), "holo_pack_render_mlp_tc");
    pack_render_tc_fill_kernel<<<32, 256, 0, st>>>(A_eff, c_eff, Wr, br, C, E, (uint8_t*)image);
    HOLO_CHECK_LAUNCH("holo_pack_render_mlp_tc");
    return HOLO_OK;
}

extern "C" int holo_render_fwd_tc(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent,
                                  const void* tc_image, int n_harmonic, const float* origins, const float* dirs,
                                  const float* lengths, int n_rays, int S, int n_passes, int n_fine,
                                  int add_input_samples, const float* bg3_host, float background_opacity,
                                  float* features, float* depths, float* masks, float* weights, float* lengths_out,
                                  float* prev_features, float* prev_depths, float* prev_masks, float* prev_weights,
                                  float* scratch_weights, void* stream) {
    if (n_rays == 0) return HOLO_OK;
    HOLO_CHECK_ARG(grid_dhwc && tc_image && origins && dirs && lengths, "holo_render_fwd_tc: null input");
    HOLO_CHECK_ARG(features && depths && masks, "holo_render_fwd_tc: null output");
    HOLO_CHECK_ARG(n_passes == 1 || n_passes == 2, "holo_render_fwd_tc: n_passes must be 1 or 2");
    HOLO_CHECK_ARG(S >= 2 && (n_passes == 1 || (n_fine >= 1 && S >= 3 && scratch_weights)),
                   "holo_render_fwd_tc: bad S/n_fine or missing scratch (S*n_rays floats)");
    HOLO_CHECK_ARG(D > 1 && H > 1 && W > 1, "holo_render_fwd_tc: grid must be at least 2^3");
    if (!(C == 16 || C == 32)) {
        holo_set_error("holo_render_fwd_tc: C=%d unsupported (16, 32)", C);
        return HOLO_ERR_UNSUPPORTED;
    }
    RenderTcParams P;
    P.grid = grid_dhwc, P.D = D, P.Hh = H, P.Ww = W, P.C = C;
    P.inv_x = 1.0f / ((float)(W - 1) * (volume_extent / (float)W) * 0.5f);
    P.inv_y = 1.0f / ((float)(H - 1) * (volume_extent / (float)H) * 0.5f);
    P.inv_z = 1.0f / ((float)(D - 1) * (volume_extent / (float)D) * 0.5f);
    P.image = (const uint8_t*)tc_image;
    P.n_harm = n_harmonic;
    P.origins = origins, P.dirs = dirs, P.lengths = lengths;
    P.n_rays = n_rays, P.S = S, P.n_fine = n_passes > 1 ? n_fine : 0, P.add_input = add_input_samples ? 1 : 0;
    P.n_passes = n_passes;
    P.bg[0] = bg3_host ? bg3_host[0] : 1.f, P.bg[1] = bg3_host ? bg3_host[1] : 1.f, P.bg[2] = bg3_host ? bg3_host[2] : 1.f;
    P.bg_opacity = background_opacity;
    P.features = features, P.depths = depths, P.masks = masks, P.weights = weights, P.lengths_out = lengths_out;
    P.p_features = prev_features, P.p_depths = prev_depths, P.p_masks = prev_masks, P.p_weights = prev_weights;
    P.scratch_w = scratch_weights;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = holo_cdiv(n_rays, RAY_THREADS);
    if (C == 32) {
        HOLO_CUDA(cudaFuncSetAttribute(render_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                  "holo_render_fwd_tc");
        render_tc_kernel<32><<<blocks, THREADS, SMEM_BYTES, st>>>(P);
    } else {
        HOLO_CUDA(cudaFuncSetAttribute(render_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                  "holo_render_fwd_tc");
        render_tc_kernel<16><<<blocks, THREADS, SMEM_BYTES, st>>>(P);
    }
    HOLO_CHECK_LAUNCH("holo_render_fwd_tc");
    return HOLO_OK;
}
