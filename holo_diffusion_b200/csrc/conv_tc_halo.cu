// tcgen05 3x3x3 convolution with a HALO-RESIDENT activation tile (sm_100a).
//
// conv_tc.cu re-loads the 128-voxel operand box once per tap (27x) and the whole weight tensor once per 128 voxels;
// at 64^3 that makes the layer L2->SMEM bandwidth bound (~9 TB/s measured) at ~22% tensor-pipe activity.  Here a CTA
// owns an 8(w) x 16(h) x 2(d) output tile (two M=128 accumulators in TMEM) and, per 32-channel half-slab and per
// kw, loads ONE box 8(w) x 18(h) x 4(d) (shifted by kw-1 in w).  Rows of that box are ordered w + 8*(h' + 18*d'), so
// the 8-row core groups of a tap (kh, kd) and depth slice dd start at 512 B * (kh + 18*(kd + dd)) and follow each
// other at a uniform 512-byte stride for h = 0..15 -- exactly the K-major SWIZZLE_64B canonical layout, so all 9
// (kh, kd) taps x 2 slices are UMMA operand views of the same shared-memory bytes.  Activation traffic drops 4x,
// weight traffic 2x (one B tile feeds both accumulators): 432 KB per 128 voxels instead of 1306 KB at Cin = 64.
//
// Same math as conv_tc.cu (3xBF16, fp32 TMEM accumulation) and same epilogue; used by holo_conv3d_tc for
// ksize 3, stride 1, W % 8 == 0, H % 16 == 0, D % 2 == 0, Cin % 32 == 0, Cout % 64 == 0.
#include "common.cuh"
#include "../../include/holo_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>

namespace {

constexpr int TW = 8, TH = 16, TD = 2;
constexpr int HH = TH + 2, HD = TD + 2;        // halo extents in h and d (w is handled by the kw shift)
constexpr int HS = 32;                          // channels per half-slab = one 64-byte swizzle row
constexpr int BLOCK_N = 64;
constexpr int A_ROWS = TW * HH * HD;            // 576
constexpr int A_BYTES = A_ROWS * 64;            // 36864 per operand (hi or lo)
constexpr int A_STAGE = 2 * A_BYTES;            // 73728
constexpr int B_BYTES = BLOCK_N * 64;           // 4096 per operand
constexpr int B_ENTRY = 2 * B_BYTES;            // 8192
constexpr int NB = 8;                           // B ring entries
constexpr int NA = 2;                           // A stages
constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_ENTRY + 1024 /*align*/ + 512 /*barriers*/;
constexpr int NUM_THREADS = 192;

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major SWIZZLE_64B operand descriptor: 64-byte rows, 8-row groups 512 B apart (SBO), version 1, layout type 4
__device__ __forceinline__ uint64_t sw64_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct HaloParams {
    int Cin, D, H, W, Cout;
    const float* bias;
    const float* residual;
    float* out;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    HaloParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem + NA * A_STAGE;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + NB * B_ENTRY);
    uint64_t* a_empty = a_full + NA;
    uint64_t* b_full = a_empty + NA;
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_w = P.W / TW, tiles_h = P.H / TH;
    const int tile = blockIdx.x;
    const int w0 = (tile % tiles_w) * TW;
    const int h0 = ((tile / tiles_w) % tiles_h) * TH;
    const int d0 = (tile / (tiles_w * tiles_h)) * TD;
    const int n0 = blockIdx.y * BLOCK_N;
    const int halves = P.Cin / HS;
    const int n_stages = halves * 3;  // (half-slab, kw)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
        for (int s = 0; s < NA; ++s) mbar_init(&a_full[s], 1), mbar_init(&a_empty[s], 1);
        for (int s = 0; s < NB; ++s) mbar_init(&b_full[s], 1), mbar_init(&b_empty[s], 1);
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int sb = 0;
            uint32_t pb = 0;
            for (int st = 0; st < n_stages; ++st) {
                const int half = st / 3, kw = st % 3;
                const int sa = st % NA;
                const uint32_t pa = (st / NA) & 1;
                mbar_wait(&a_empty[sa], pa ^ 1);
                uint8_t* a = smem + sa * A_STAGE;
                mbar_expect_tx(&a_full[sa], A_STAGE);
                tma_load_4d(a, &map_a_hi, &a_full[sa], half * HS, w0 + kw - 1, h0 - 1, d0 - 1);
                tma_load_4d(a + A_BYTES, &map_a_lo, &a_full[sa], half * HS, w0 + kw - 1, h0 - 1, d0 - 1);
                for (int t9 = 0; t9 < 9; ++t9) {
                    const int kd = t9 / 3, kh = t9 % 3;
                    const int tap = (kd * 3 + kh) * 3 + kw;
                    mbar_wait(&b_empty[sb], pb ^ 1);
                    uint8_t* b = smem_b + sb * B_ENTRY;
                    mbar_expect_tx(&b_full[sb], B_ENTRY);
                    const int kk = tap * P.Cin + half * HS;
                    tma_load_2d(b, &map_b_hi, &b_full[sb], kk, n0);
                    tma_load_2d(b + B_BYTES, &map_b_lo, &b_full[sb], kk, n0);
                    if (++sb == NB) sb = 0, pb ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // x_hi . [w_hi | w_lo]^T (N = 128) then x_lo . w_hi^T (N = 64) into the first half: see conv_tc.cu
        constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int sb = 0;
        uint32_t pb = 0;
        for (int st = 0; st < n_stages; ++st) {
            const int sa = st % NA;
            const uint32_t pa = (st / NA) & 1;
            mbar_wait(&a_full[sa], pa);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(smem + sa * A_STAGE), a_lo = a_hi + A_BYTES;
            for (int t9 = 0; t9 < 9; ++t9) {
                const int kd = t9 / 3, kh = t9 % 3;
                mbar_wait(&b_full[sb], pb);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t b_hi = smem_u32(smem_b + sb * B_ENTRY);  // w_lo rows follow at + B_BYTES
#pragma unroll
                    for (int dd = 0; dd < TD; ++dd) {
                        // operand view of tap (kh, kd) for depth slice dd: 16 groups of 8 rows, 512 B apart
                        const uint32_t off = 512u * (uint32_t)(kh + HH * (kd + dd));
                        const uint32_t d_tmem = tmem_base + (uint32_t)(dd * 2 * BLOCK_N);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {  // 32 channels = 2 K-steps of 16
                            const uint32_t ko = k * 32;
                            const uint64_t dah = sw64_desc(a_hi + off + ko), dal = sw64_desc(a_lo + off + ko);
                            const uint64_t dbh = sw64_desc(b_hi + ko);
                            umma_bf16(d_tmem, dah, dbh, idesc1, (st | t9 | k) != 0);
                            umma_bf16(d_tmem, dal, dbh, idesc2, 1);
                        }
                    }
                    umma_commit(&b_empty[sb]);
                    if (t9 == 8) umma_commit(&a_empty[sa]);
                    if (st == n_stages - 1 && t9 == 8) umma_commit(acc_full);
                }
                __syncwarp();
                if (++sb == NB) sb = 0, pb ^= 1;
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp % 4;
        const int r = q * 32 + lane;
        const int w = w0 + (r % TW), h = h0 + (r / TW);
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int dd = 0; dd < TD; ++dd) {
            const size_t v = ((size_t)(d0 + dd) * P.H + h) * P.W + w;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                uint32_t acc[16], acc2[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dd * 2 * BLOCK_N + c0), acc);
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dd * 2 * BLOCK_N + BLOCK_N + c0), acc2);
                const int n = n0 + c0;
                float vals[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) vals[j] = __uint_as_float(acc[j]) + __uint_as_float(acc2[j]);
                if (P.bias) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n) + j4);
                        vals[j4 * 4 + 0] += b.x, vals[j4 * 4 + 1] += b.y, vals[j4 * 4 + 2] += b.z, vals[j4 * 4 + 3] += b.w;
                    }
                }
                if (P.residual) {
                    const float4* rp = reinterpret_cast<const float4*>(P.residual + v * P.Cout + n);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 b = __ldg(rp + j4);
                        vals[j4 * 4 + 0] += b.x, vals[j4 * 4 + 1] += b.y, vals[j4 * 4 + 2] += b.z, vals[j4 * 4 + 3] += b.w;
                    }
                }
                if (P.out) {
                    float4* op = reinterpret_cast<float4*>(P.out + v * P.Cout + n);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        op[j4] = make_float4(vals[j4 * 4], vals[j4 * 4 + 1], vals[j4 * 4 + 2], vals[j4 * 4 + 3]);
                }
                if (P.out_hi) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        __nv_bfloat16 h0b = __float2bfloat16_rn(vals[2 * j]), h1b = __float2bfloat16_rn(vals[2 * j + 1]);
                        __nv_bfloat16 l0b = __float2bfloat16_rn(vals[2 * j] - __bfloat162float(h0b));
                        __nv_bfloat16 l1b = __float2bfloat16_rn(vals[2 * j + 1] - __bfloat162float(h1b));
                        hi[j] = (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
                        lo[j] = (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
                    }
                    uint4* hp = reinterpret_cast<uint4*>(P.out_hi + v * P.Cout + n);
                    uint4* lp = reinterpret_cast<uint4*>(P.out_lo + v * P.Cout + n);
                    hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]), hp[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    lp[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]), lp[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// returns HOLO_ERR_UNSUPPORTED when the shape is not eligible (the caller then uses conv_tc.cu)
int holo_conv3d_tc_halo(const void* x_hi, const void* x_lo, int Cin, int D, int H, int W, const void* w_hi,
                        const void* w_lo, const float* bias, const float* residual, int Cout, float* out,
                        void* out_hi, void* out_lo, cudaStream_t st) {
    if (W % TW || H % TH || D % TD || Cin % HS || Cout % BLOCK_N) return HOLO_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        holo_set_error("holo_conv3d_tc: cuTensorMapEncodeTiled unavailable");
        return HOLO_ERR_CUDA;
    }
    CUtensorMap ah, al, bh, bl;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {HS, TW, HH, HD};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r1 = enc(&ah, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_hi), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&al, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_lo), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
            holo_set_error("holo_conv3d_tc: activation tensor map failed (%d, %d)", (int)r1, (int)r2);
            return HOLO_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)27 * Cin, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)27 * Cin * 2};
        cuuint32_t box[2] = {HS, BLOCK_N};
        cuuint32_t es[2] = {1, 1};
        CUresult r1 = enc(&bh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_hi), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&bl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_lo), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
            holo_set_error("holo_conv3d_tc: weight tensor map failed (%d, %d)", (int)r1, (int)r2);
            return HOLO_ERR_CUDA;
        }
    }
    HaloParams P;
    P.Cin = Cin, P.D = D, P.H = H, P.W = W, P.Cout = Cout;
    P.bias = bias, P.residual = residual, P.out = out;
    P.out_hi = (__nv_bfloat16*)out_hi, P.out_lo = (__nv_bfloat16*)out_lo;
    HOLO_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
              "holo_conv3d_tc");
    const int tiles = (W / TW) * (H / TH) * (D / TD);
    conv_tc_halo_kernel<<<dim3(tiles, Cout / BLOCK_N), NUM_THREADS, SMEM_BYTES, st>>>(ah, al, bh, bl, P);
    HOLO_CHECK_LAUNCH("holo_conv3d_tc");
    return HOLO_OK;
}
