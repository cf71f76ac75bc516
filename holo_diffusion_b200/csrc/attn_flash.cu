// Fused (flash-style) QKVAttention over voxel tokens on tcgen05 for sm_100a: the T x T logits never leave the SM.
//
// Reference op replaced: QKVAttentionLegacy.forward -- /root/reference/holo_diffusion/guided_diffusion/unet.py:438-455
//   qkv (B*H, 3*ch, T) -> q, k, v;  w = softmax((q s)^T (k s));  a = w v,   s = ch^-1/4.
//
// One CTA = 128 queries of one head.  Keys/values stream through shared memory in 64-key tiles (TMA, SWIZZLE_128B):
//   S = Q K^T        tcgen05.mma kind::f16, 3xBF16 split (hi.hi + hi.lo + lo.hi), fp32 accumulator in TMEM (double-buffered)
//   P = exp2(s^2 log2e (S - rowmax))  fp32 in registers, written back to shared memory as a bf16 hi/lo A operand
//   O += P V         tcgen05.mma, 3xBF16 split, fp32 accumulator in TMEM across all key tiles
// The softmax is EXACT two-pass rather than online: pass A streams K only and reduces the row maxima, pass B
// recomputes S, exponentiates against the final maximum and accumulates O with no rescaling of the TMEM
// accumulator (1.5x the tensor work of one pass; the tensor pipe is not what bounds this kernel at T <= 4096).
//
// Pass A only needs a stabiliser close to the true maximum (softmax is invariant to it), so it multiplies the bf16
// hi parts only (1 MMA per K step instead of 3).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = softmax + epilogue:
// two warps per TMEM lane quarter, each thread owns one query row and HALF of the tile's 64 key columns (one warp
// per scheduler was issue/latency bound: 130 us per T = 4096 block, see profiles/).
#include "common.cuh"
#include "../../include/holo_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>

namespace {

constexpr int BM = 128;  // queries per CTA (UMMA M)
constexpr int BN = 64;   // keys per tile = one 128-byte swizzle row of P
constexpr int NTHREADS = 320;
constexpr int SOFTMAX_THREADS = 256;
constexpr int TMEM_COLS = 512;  // S0 [0,64) | S1 [64,128) | O chain A [128, +ch) | O chain B [.., +ch) | O sum [.., +ch)
constexpr int O_COL = 128;

template <int CH>
struct FCfg {
    static constexpr int SLABS = CH / 64;
    static constexpr int Q_HALF = SLABS * BM * 128;  // bytes of q_hi (then q_lo)
    static constexpr int K_HALF = SLABS * BN * 128;
    static constexpr int V_HALF = CH * 128;
    static constexpr int Q_BYTES = 2 * Q_HALF, K_BYTES = 2 * K_HALF, V_BYTES = 2 * V_HALF;
    static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
    // the K/V ring must cover the TMA round trip: a stage is held from its load until P V of that tile retires
    // (3 stages stalled every tile of pass B by ~0.4 us at ch = 64)
    static constexpr int STAGES = CH == 64 ? 4 : 2;
    static constexpr int P_HALF = BM * 128;
    static constexpr int P_BYTES = 2 * P_HALF;       // one P tile: [p_hi | p_lo]
    static constexpr int PBUF = CH == 64 ? 2 : 1;    // P double-buffered where shared memory allows (ch = 64): with a
                                                     // single buffer softmax(i+1) -> P V(i) -> softmax... serialises
    static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + PBUF * P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------- PTX wrappers (same conventions as conv_tc.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major SWIZZLE_128B operand descriptor: 8-row groups 1024 B apart (see conv_tc.cu)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16 (format 1) or fp16 (format 0), both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_16(int n, bool f16) {
    return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n\t"
        "tcgen05.wait::st.sync.aligned;" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// probabilities (a, b) in [0, 1] -> packed hi and lo halves (bf16x2, or fp16x2 without the saturation holo_split2
// needs for unbounded values), a in the low half: 2 packed converts + 2 subtractions (+ 1 unpack for fp16)
template <bool F16>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (F16) {
        const __half2 h = __floats2half2_rn(a, b);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    } else {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
        const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    }
}
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, relative error 2^-22
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct FlashParams {
    int T, C, heads;
    float scale_log2;  // ch^-1/2 * log2(e): logits are (q s).(k s) = s^2 q.k
    float* out;        // (T, C) fp32 or null
    __nv_bfloat16* out_hi;  // (T, C) hi/lo split of the result (same 16-bit format as the inputs), or null
    __nv_bfloat16* out_lo;
    int o_chunk;            // key tiles per O accumulation chain (see "chunked O accumulation" in the kernel)
    int q_tile0;            // first 128-query tile of this launch (query-sharded attention: a rank owns a tile range)
    // split-KV (gridDim.z > 1): CTA z handles key tiles [z * kt_per_split, ...) with its OWN stabiliser m_z and writes
    // the un-normalised O_z (part_o: [z][T][C] fp32) and (m_z * scale_log2, rowsum_z) (part_ml: [z][heads][T] float2);
    // flash_combine_kernel merges: O = sum_z 2^(m_z - M) O_z / sum_z 2^(m_z - M) l_z.  Fills the 148 SMs when
    // (T / 128) x heads is small (64 CTAs at T = 4096, 8 at T = 512).
    int kt_per_split;
    float* part_o;
    float2* part_ml;
};

// F16: every operand pair (q, k, v^T in, P inside, the output pair) has fp16 halves instead of bf16.  The logits'
// absolute error is what softmax turns into a relative error of P: 2^-17 |q||k| with bf16 pairs (6e-5 through the
// 64^3 UNet), 2^-22 with fp16 pairs.
template <int CH, bool F16>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                  const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                  const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                  FlashParams P) {
    using F = FCfg<CH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* q_smem = smem;                                   // [q_hi slabs | q_lo slabs]
    uint8_t* kv_smem = smem + F::Q_BYTES;                     // STAGES x [k_hi | k_lo | v_hi | v_lo]
    uint8_t* p_smem = kv_smem + F::STAGES * F::STAGE_BYTES;   // PBUF x [p_hi | p_lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + F::PBUF * F::P_BYTES);
    uint64_t* kv_full = bars;                     // [STAGES]
    uint64_t* kv_empty = kv_full + F::STAGES;     // [STAGES]
    uint64_t* s_full = kv_empty + F::STAGES;      // [2]
    uint64_t* s_empty = s_full + 2;               // [2]
    uint64_t* p_full = s_empty + 2;               // [2]
    uint64_t* p_empty = p_full + 2;               // [2]
    uint64_t* q_full = p_empty + 2;
    uint64_t* o_full = q_full + 1;                // [2] chain buffer A / B holds a finished chain
    uint64_t* o_free = o_full + 2;                // [2] the softmax warps have folded it into the running sum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m0 = (P.q_tile0 + (int)blockIdx.x) * BM;
    const int head = blockIdx.y;
    // split-KV is an OPT-IN BUILD (-DHOLO_ENABLE_SPLIT_KV, HOLO_BUILD_EXPERIMENTAL=1): it has not run on a GPU yet, and
    // the default build keeps the device code that has (identical SASS to the validated revision)
    const int kt0 = (int)blockIdx.z * P.kt_per_split;                     // first key tile of this CTA's share
    const int NT = min(P.T / BN - kt0, P.kt_per_split);                   // key tiles per pass (all of them unsplit)
    const int n_jobs = 2 * NT;      // S jobs: pass A (max) then pass B (exp + PV)
    const int col_q = head * 3 * CH, col_k = col_q + CH;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v_lo) : "memory");
        for (int s = 0; s < F::STAGES; ++s) mbar_init(&kv_full[s], 1), mbar_init(&kv_empty[s], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&s_full[b], 1), mbar_init(&s_empty[b], SOFTMAX_THREADS / 32);
        for (int b = 0; b < 2; ++b) mbar_init(&p_full[b], SOFTMAX_THREADS), mbar_init(&p_empty[b], 1);
        mbar_init(q_full, 1);
        for (int b = 0; b < 2; ++b) mbar_init(&o_full[b], 1), mbar_init(&o_free[b], SOFTMAX_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    holo_pdl_trigger();   // opt-in PDL build only (common.cuh)
    holo_pdl_wait();

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(q_full, (uint32_t)F::Q_BYTES);
#pragma unroll
            for (int sl = 0; sl < F::SLABS; ++sl) {
                tma_load_2d(q_smem + sl * (BM * 128), &map_q_hi, q_full, col_q + sl * 64, m0);
                tma_load_2d(q_smem + F::Q_HALF + sl * (BM * 128), &map_q_lo, q_full, col_q + sl * 64, m0);
            }
            for (int j = 0; j < n_jobs; ++j) {
                const int stage = j % F::STAGES;
                const uint32_t phase = (uint32_t)(j / F::STAGES) & 1u;
                const bool with_v = j >= NT;
                const int n0 = (kt0 + (with_v ? j - NT : j)) * BN;
                mbar_wait(&kv_empty[stage], phase ^ 1u);
                uint8_t* st = kv_smem + stage * F::STAGE_BYTES;
                mbar_expect_tx(&kv_full[stage], (uint32_t)(F::K_BYTES + (with_v ? F::V_BYTES : 0)));
#pragma unroll
                for (int sl = 0; sl < F::SLABS; ++sl) {
                    tma_load_2d(st + sl * (BN * 128), &map_k_hi, &kv_full[stage], col_k + sl * 64, n0);
                    tma_load_2d(st + F::K_HALF + sl * (BN * 128), &map_k_lo, &kv_full[stage], col_k + sl * 64, n0);
                }
                if (with_v) {
                    tma_load_2d(st + F::K_BYTES, &map_v_hi, &kv_full[stage], n0, head * CH);
                    tma_load_2d(st + F::K_BYTES + F::V_HALF, &map_v_lo, &kv_full[stage], n0, head * CH);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc_s = idesc_16(BN, F16);
        constexpr uint32_t idesc_o = idesc_16(CH, F16);
        const uint32_t q_base = smem_u32(q_smem);
        const uint32_t p_base = smem_u32(p_smem);
        // S job j: wait for its K tile and for the softmax warps to have drained the S buffer, then 3 x SLABS x 4 UMMAs
        auto issue_s = [&](int j) {
            const int stage = j % F::STAGES;
            const int b = j & 1;
            mbar_wait(&kv_full[stage], (uint32_t)(j / F::STAGES) & 1u);
            mbar_wait(&s_empty[b], ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)(b * BN);
                const uint32_t k_base = smem_u32(kv_smem + stage * F::STAGE_BYTES);
#pragma unroll
                for (int sl = 0; sl < F::SLABS; ++sl)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t a_hi = q_base + sl * (BM * 128) + k * 32, a_lo = a_hi + F::Q_HALF;
                        const uint32_t b_hi = k_base + sl * (BN * 128) + k * 32, b_lo = b_hi + F::K_HALF;
                        const uint64_t dah = sw128_desc(a_hi), dbh = sw128_desc(b_hi);
                        umma(d, dah, dbh, idesc_s, (sl | k) != 0);
                        if (j >= NT) {  // pass B: the full 3xBF16 product; pass A: hi.hi is enough for the stabiliser
                            umma(d, dah, sw128_desc(b_lo), idesc_s, 1);
                            umma(d, sw128_desc(a_lo), dbh, idesc_s, 1);
                        }
                    }
                umma_commit(&s_full[b]);
                if (j < NT) umma_commit(&kv_empty[stage]);  // pass A: the stage holds K only, free it now
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        tc_fence_after();
        for (int j = 0; j < NT; ++j) issue_s(j);  // pass A
        issue_s(NT);                               // pass B prologue
        // CHUNKED O ACCUMULATION: the tensor core truncates every add into the fp32 TMEM accumulator (a bias of
        // ~3e-8 of the running sum per MMA, 12 MMAs per key tile): over T = 2 M keys (BASELINE cfg #5) one chain would
        // be 1e-2 off.  O is therefore accumulated in chains of P.o_chunk key tiles, alternating between two TMEM
        // buffers that start from zero; the softmax warps fold every finished chain into a running sum kept in a third
        // TMEM region (tcgen05.ld / add in registers, round-to-nearest / tcgen05.st) while the next chain runs.
        const int oc = P.o_chunk;
        for (int i = 0; i < NT; ++i) {
            const int j = NT + i;
            if (i + 1 < NT) issue_s(j + 1);        // S of the next tile overlaps the softmax of this one
            const int stage = j % F::STAGES;
            const int pb = i % F::PBUF;
            const int chain = i / oc, ob = chain & 1, use = chain >> 1;
            const bool first = (i % oc) == 0, last = (i % oc) == oc - 1 || i == NT - 1;
            if (first && use > 0) {                // the chain that used this buffer two chains ago has been folded
                mbar_wait(&o_free[ob], (uint32_t)(use - 1) & 1u);
                tc_fence_after();
            }
            mbar_wait(&p_full[pb], (uint32_t)(i / F::PBUF) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + O_COL + (uint32_t)(ob * CH);
                const uint32_t v_base = smem_u32(kv_smem + stage * F::STAGE_BYTES + F::K_BYTES);
#pragma unroll
                for (int k = 0; k < BN / 16; ++k) {
                    const uint32_t a_hi = p_base + pb * F::P_BYTES + k * 32, a_lo = a_hi + F::P_HALF;
                    const uint32_t b_hi = v_base + k * 32, b_lo = b_hi + F::V_HALF;
                    const uint64_t dah = sw128_desc(a_hi), dbh = sw128_desc(b_hi);
                    umma(d, dah, dbh, idesc_o, (!first) || k != 0);
                    umma(d, dah, sw128_desc(b_lo), idesc_o, 1);
                    umma(d, sw128_desc(a_lo), dbh, idesc_o, 1);
                }
                umma_commit(&p_empty[pb]);
                umma_commit(&kv_empty[stage]);
                if (last) umma_commit(&o_full[ob]);
            }
            __syncwarp();
        }
    } else {
        // ================= softmax + epilogue =================
        const int q = warp % 4;             // TMEM lane quarter of this warp
        const int half = (warp - 2) / 4;    // which 32 of the tile's 64 key columns (and which half of O's columns)
        const int row = q * 32 + lane;      // query row of the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* xch = reinterpret_cast<float*>(p_smem);  // [2][128] exchange between the two column halves of a row
        float mx = -INFINITY;
        uint32_t v[32];
        // ---- pass A: row maxima of the (bf16 hi.hi) logits
        for (int j = 0; j < NT; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[b], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            tmem_ld32(lane_addr + (uint32_t)(b * BN + half * 32), v);
#pragma unroll
            for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
        }
        xch[half * BM + row] = mx;                       // the P buffer is idle until pass B
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(xch[row], xch[BM + row]);
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read before P tiles overwrite the scratch
        // ---- pass B: P = exp2(scale (S - max)), row sums, P -> shared memory as the bf16 hi/lo A operand of P V
        const float msc = mx * P.scale_log2;
        float l0 = 0.f, l1 = 0.f;
        uint8_t* prow = p_smem + (size_t)(row / 8) * 1024 + (row % 8) * 128;
        const int swz = row % 8;
        // chunked O accumulation (see the MMA warp): fold a finished chain into the running sum in TMEM; this thread
        // owns its query row x half of the channels
        const int oc = P.o_chunk;
        const int n_chains = (NT + oc - 1) / oc;
        const uint32_t o_sum = lane_addr + (uint32_t)(O_COL + 2 * CH);
        auto fold = [&](int chain) {
            const int ob = chain & 1;
            mbar_wait(&o_full[ob], (uint32_t)(chain >> 1) & 1u);
            tc_fence_after();
            const uint32_t src = lane_addr + (uint32_t)(O_COL + ob * CH);
#pragma unroll 1
            for (int c0 = half * (CH / 2); c0 < (half + 1) * (CH / 2); c0 += 16) {
                uint32_t a[16];
                tmem_ld16(src + (uint32_t)c0, a);
                if (chain > 0) {
                    uint32_t bsum[16];
                    tmem_ld16(o_sum + (uint32_t)c0, bsum);
#pragma unroll
                    for (int k = 0; k < 16; ++k) a[k] = __float_as_uint(__uint_as_float(a[k]) + __uint_as_float(bsum[k]));
                }
                tmem_st16(o_sum + (uint32_t)c0, a);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[ob]);
        };
        for (int i = 0; i < NT; ++i) {
            const int j = NT + i;
            const int b = j & 1;
            if (i > oc && i % oc == 1) fold(i / oc - 1);   // one tile late: the chain's last P V has retired by now
            mbar_wait(&s_full[b], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            uint32_t hi[16], lo[16];  // 32 keys = 16 bf16x2 each
            tmem_ld32(lane_addr + (uint32_t)(b * BN + half * 32), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);      // S buffer drained (values live in registers now)
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(v[c]), P.scale_log2, -msc));
                const float p1 = ex2_approx(fmaf(__uint_as_float(v[c + 1]), P.scale_log2, -msc));
                l0 += p0, l1 += p1;
                split2<F16>(p0, p1, hi[c / 2], lo[c / 2]);
            }
            const int pb = i % F::PBUF;
            mbar_wait(&p_empty[pb], ((uint32_t)(i / F::PBUF) & 1u) ^ 1u);  // the P V that last read this buffer retired
            uint8_t* prow_hi = prow + pb * F::P_BYTES;
            uint8_t* prow_lo = prow_hi + F::P_HALF;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                  // 16-byte chunk ck = keys 8 ck .. 8 ck + 7
                const int off = ((half * 4 + c) ^ swz) * 16;
                *reinterpret_cast<uint4*>(prow_hi + off) = make_uint4(hi[c * 4], hi[c * 4 + 1], hi[c * 4 + 2], hi[c * 4 + 3]);
                *reinterpret_cast<uint4*>(prow_lo + off) = make_uint4(lo[c * 4], lo[c * 4 + 1], lo[c * 4 + 2], lo[c * 4 + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA (async proxy)
            mbar_arrive(&p_full[pb]);
        }
        // ---- epilogue: O / rowsum -> global (fp32 and / or the bf16 hi/lo split the projection conv consumes)
        if (n_chains >= 2 && NT - (n_chains - 1) * oc < 2) fold(n_chains - 2);   // not reached inside the loop
        const int last_chain = n_chains - 1;
        mbar_wait(&o_full[last_chain & 1], (uint32_t)(last_chain >> 1) & 1u);   // all MMAs retired: P buffer free again
        tc_fence_after();
        const uint32_t o_last = lane_addr + (uint32_t)(O_COL + (last_chain & 1) * CH);
        // 32 channels of O = last chain (+ the folded earlier chains)
        auto load_o = [&](int c0, uint32_t (&out)[32]) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t a[16];
                tmem_ld16(o_last + (uint32_t)(c0 + 16 * hh), a);
                if (n_chains >= 2) {
                    uint32_t bsum[16];
                    tmem_ld16(o_sum + (uint32_t)(c0 + 16 * hh), bsum);
#pragma unroll
                    for (int k = 0; k < 16; ++k) a[k] = __float_as_uint(__uint_as_float(a[k]) + __uint_as_float(bsum[k]));
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) out[16 * hh + k] = a[k];
            }
        };
        xch[half * BM + row] = l0 + l1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float lsum = xch[row] + xch[BM + row];
        const int t = m0 + row;
        const bool ok = t < P.T;
        const size_t obase = (size_t)t * P.C + (size_t)head * CH;
        if (gridDim.z > 1) {
            // split-KV: this CTA's share of the keys only -- leave O un-normalised for flash_combine_kernel
            if (ok && half == 0) P.part_ml[((size_t)blockIdx.z * P.heads + head) * P.T + t] = make_float2(msc, lsum);
            float* po = P.part_o + (size_t)blockIdx.z * P.T * P.C + obase;
#pragma unroll 1
            for (int c0 = half * (CH / 2); c0 < (half + 1) * (CH / 2); c0 += 32) {
                load_o(c0, v);
                if (ok) {
                    float4* op = reinterpret_cast<float4*>(po + c0);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4)
                        op[c4] = make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]),
                                             __uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3]));
                }
            }
            tc_fence_before();
        } else
        {
        const float inv = 1.0f / lsum;
#pragma unroll 1
        for (int c0 = half * (CH / 2); c0 < (half + 1) * (CH / 2); c0 += 32) {
            load_o(c0, v);
            if (ok) {
                float f[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]) * inv;
                if (P.out) {
                    float4* op = reinterpret_cast<float4*>(P.out + obase + c0);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) op[c4] = make_float4(f[c4 * 4], f[c4 * 4 + 1], f[c4 * 4 + 2], f[c4 * 4 + 3]);
                }
                if (P.out_hi) {
                    uint32_t oh[16], ol[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) holo_split2(f[2 * c], f[2 * c + 1], F16, oh[c], ol[c]);
                    uint4* hp = reinterpret_cast<uint4*>(P.out_hi + obase + c0);
                    uint4* lp = reinterpret_cast<uint4*>(P.out_lo + obase + c0);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        hp[c4] = make_uint4(oh[c4 * 4], oh[c4 * 4 + 1], oh[c4 * 4 + 2], oh[c4 * 4 + 3]);
                        lp[c4] = make_uint4(ol[c4 * 4], ol[c4 * 4 + 1], ol[c4 * 4 + 2], ol[c4 * 4 + 3]);
                    }
                }
            }
        }
        tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// Merge of the split-KV partials: thread = 8 channels of one query row.
__global__ void flash_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, int splits,
                                     int T, int heads, int ch, int q_begin, int q_count, float* __restrict__ out,
                                     uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int pair_f16) {
    const int C = heads * ch;
    const int c8 = C / 8;
    const long long total = (long long)q_count * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = q_begin + (int)(i / c8);
        const int c = (int)(i % c8) * 8;
        const int head = c / ch;
        float M = -INFINITY;
        for (int z = 0; z < splits; ++z) M = fmaxf(M, part_ml[((size_t)z * heads + head) * T + t].x);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float l = 0.f;
        for (int z = 0; z < splits; ++z) {
            const float2 ml = part_ml[((size_t)z * heads + head) * T + t];
            const float w = exp2f(ml.x - M);
            l = fmaf(w, ml.y, l);
            const float4* po = reinterpret_cast<const float4*>(part_o + ((size_t)z * T + t) * C + c);
            const float4 a = po[0], b = po[1];
            acc[0] = fmaf(w, a.x, acc[0]), acc[1] = fmaf(w, a.y, acc[1]), acc[2] = fmaf(w, a.z, acc[2]), acc[3] = fmaf(w, a.w, acc[3]);
            acc[4] = fmaf(w, b.x, acc[4]), acc[5] = fmaf(w, b.y, acc[5]), acc[6] = fmaf(w, b.z, acc[6]), acc[7] = fmaf(w, b.w, acc[7]);
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= inv;
        const size_t o = (size_t)t * C + c;
        if (out) {
            *reinterpret_cast<float4*>(out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(out + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        if (out_hi) {
            uint32_t h[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) holo_split2(acc[2 * k], acc[2 * k + 1], pair_f16 != 0, h[k], lo[k]);
            *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// V (T, ch) slices of the head-major qkv tensor -> V^T (heads*ch, T) bf16 hi/lo, K-major for the P V product
__global__ void v_transpose_split_kernel(const float* __restrict__ qkv, int T, int heads, int ch,
                                         uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int pair_f16) {
    __shared__ float tile[32][33];
    const int C = heads * ch;
    const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;   // tokens on grid.x: T / 32 exceeds grid.y's 65535 at 128^3
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        float val = 0.f;
        if (t < T && c < C) val = qkv[(size_t)t * 3 * C + (size_t)(c / ch) * 3 * ch + 2 * ch + (c % ch)];
        tile[i][threadIdx.x] = val;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (t < T && c < C) {
            holo_split1(tile[threadIdx.x][i], pair_f16 != 0, hi[(size_t)c * T + t], lo[(size_t)c * T + t]);
        }
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

// 2-D bf16 tensor (rows, cols), row pitch in elements, box (64 cols = 128 B, box_rows), SWIZZLE_128B
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long pitch, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int CH, bool F16>
int launch_flash(const CUtensorMap* maps, const FlashParams& P, int q_tiles, int splits, cudaStream_t st) {
    auto k = attn_flash_kernel<CH, F16>;
    HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, FCfg<CH>::SMEM_BYTES),
              "holo_attention_flash");
    dim3 grid((unsigned)q_tiles, (unsigned)P.heads, (unsigned)splits);
    holo_launch(k, grid, dim3(NTHREADS), (size_t)FCfg<CH>::SMEM_BYTES, st, maps[0], maps[1], maps[2], maps[3], maps[4],
                maps[5], P);
    HOLO_CHECK_LAUNCH("holo_attention_flash");
    return HOLO_OK;
}

}  // namespace

extern "C" int holo_v_transpose_split(const float* qkv_cl, int T, int heads, int ch, void* vt_hi_bf16,
                                      void* vt_lo_bf16, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(qkv_cl && vt_hi_bf16 && vt_lo_bf16 && T > 0 && heads > 0 && ch > 0, "holo_v_transpose_split: bad args");
    dim3 grid(holo_cdiv(T, 32), holo_cdiv((long long)heads * ch, 32));
    v_transpose_split_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(qkv_cl, T, heads, ch, (uint16_t*)vt_hi_bf16,
                                                                            (uint16_t*)vt_lo_bf16, pair_f16);
    HOLO_CHECK_LAUNCH("holo_v_transpose_split");
    return HOLO_OK;
}

extern "C" int holo_attention_flash(const void* qkv_hi_bf16, const void* qkv_lo_bf16, const void* vt_hi_bf16,
                                    const void* vt_lo_bf16, int T, int heads, int ch, float* out_cl, void* out_hi_bf16,
                                    void* out_lo, int pair_f16, float softmax_scale, int q_begin, int q_count,
                                    int kv_splits, void* workspace, void* stream) {
    void* out_lo_bf16 = out_lo;
    HOLO_CHECK_ARG(qkv_hi_bf16 && qkv_lo_bf16 && vt_hi_bf16 && vt_lo_bf16, "holo_attention_flash: null input");
    HOLO_CHECK_ARG(out_cl || out_hi_bf16, "holo_attention_flash: no output requested");
    HOLO_CHECK_ARG((out_hi_bf16 == nullptr) == (out_lo_bf16 == nullptr), "holo_attention_flash: hi/lo outputs come together");
    if (!(ch == 64 || ch == 128) || T < BN || T % BN != 0 || heads < 1) {
        holo_set_error("holo_attention_flash: unsupported shape T=%d heads=%d ch=%d (ch 64|128, T %% 64 == 0)", T, heads, ch);
        return HOLO_ERR_UNSUPPORTED;
    }
    if (q_count <= 0) q_begin = 0, q_count = T;
    HOLO_CHECK_ARG(q_begin >= 0 && q_begin % BM == 0 && q_begin + q_count <= T && (q_count % BM == 0 || q_begin + q_count == T),
                   "holo_attention_flash: the query range [%d, %d) must start on a multiple of 128 and end on one or at T",
                   q_begin, q_begin + q_count);
    const int C = heads * ch;
    CUtensorMap maps[6];
    int e = make_map(&maps[0], qkv_hi_bf16, T, 3LL * C, 3LL * C, BM);
    if (!e) e = make_map(&maps[1], qkv_lo_bf16, T, 3LL * C, 3LL * C, BM);
    if (!e) e = make_map(&maps[2], qkv_hi_bf16, T, 3LL * C, 3LL * C, BN);
    if (!e) e = make_map(&maps[3], qkv_lo_bf16, T, 3LL * C, 3LL * C, BN);
    if (!e) e = make_map(&maps[4], vt_hi_bf16, C, T, T, ch);
    if (!e) e = make_map(&maps[5], vt_lo_bf16, C, T, T, ch);
    if (e) {
        holo_set_error("holo_attention_flash: cuTensorMapEncodeTiled failed (%d)", e);
        return HOLO_ERR_CUDA;
    }
    FlashParams P;
    P.T = T, P.C = C, P.heads = heads;
    P.scale_log2 = (softmax_scale > 0.f ? softmax_scale : 1.0f / sqrtf((float)ch)) * 1.4426950408889634f;
    P.out = out_cl, P.out_hi = (__nv_bfloat16*)out_hi_bf16, P.out_lo = (__nv_bfloat16*)out_lo_bf16;
    P.q_tile0 = q_begin / BM;
    static const int o_chunk_env = [] {   // key tiles per O accumulation chain; >= 2 (64 tiles = 768 MMAs: bias ~2e-5)
        const char* e = getenv("HOLO_ATTN_O_CHUNK");
        const int v = e ? atoi(e) : 64;
        return v < 2 ? (1 << 30) : v;
    }();
    P.o_chunk = o_chunk_env;
    const int q_tiles = (q_count + BM - 1) / BM;
    // split-KV: kv_splits CTAs share the key tiles of one (query tile, head); needs the caller's workspace
    const int nt_all = T / BN;
    int splits = kv_splits < 1 ? 1 : (kv_splits > nt_all ? nt_all : kv_splits);
    P.kt_per_split = (nt_all + splits - 1) / splits;
    splits = (nt_all + P.kt_per_split - 1) / P.kt_per_split;
    P.part_o = nullptr, P.part_ml = nullptr;
    if (splits > 1) {
        HOLO_CHECK_ARG(workspace, "holo_attention_flash: kv_splits > 1 needs a workspace of "
                                  "holo_attention_flash_workspace_bytes(T, heads, ch, kv_splits)");
        P.part_o = reinterpret_cast<float*>(workspace);
        P.part_ml = reinterpret_cast<float2*>(P.part_o + (size_t)splits * T * C);
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (pair_f16)
        rc = ch == 64 ? launch_flash<64, true>(maps, P, q_tiles, splits, st) : launch_flash<128, true>(maps, P, q_tiles, splits, st);
    else
        rc = ch == 64 ? launch_flash<64, false>(maps, P, q_tiles, splits, st)
                      : launch_flash<128, false>(maps, P, q_tiles, splits, st);
    if (rc != HOLO_OK || splits == 1) return rc;
    const long long total = (long long)q_count * (C / 8);
    int blocks = holo_cdiv(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    flash_combine_kernel<<<blocks, 256, 0, st>>>(P.part_o, P.part_ml, splits, T, heads, ch, q_begin, q_count, out_cl,
                                                 (uint16_t*)out_hi_bf16, (uint16_t*)out_lo_bf16, pair_f16);
    HOLO_CHECK_LAUNCH("holo_attention_flash (combine)");
    return HOLO_OK;
}

extern "C" long long holo_attention_flash_workspace_bytes(int T, int heads, int ch, int kv_splits) {
    if (kv_splits <= 1) return 0;
    const long long C = (long long)heads * ch;
    return (long long)kv_splits * T * C * 4 + (long long)kv_splits * heads * T * 8;
}
