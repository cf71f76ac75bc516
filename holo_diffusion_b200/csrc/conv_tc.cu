// tcgen05 implicit-GEMM 3-D convolution for sm_100a: TMA-staged D x H x W halo tiles, UMMA (kind::f16, bf16
// operands, fp32 accumulators in TMEM), 3xBF16 operand split for fp32-grade accuracy.
//
//   out[v][n] = bias[n] + residual[v][n] + sum_{tap, c} x[v + off(tap)][c] * w[n][tap][c]
//
// GEMM view: M = 128 output voxels per CTA (an 8(w) x 4(h) x 4(d) box), N = BLOCK_N output channels,
// K = taps * Cin walked in 64-channel (128-byte) slabs.  For every slab the TMA producer loads the box shifted
// by the tap offset straight out of the channels-last activation; out-of-range rows/columns/slices are
// zero-filled by the TMA unit, which implements padding=1 with no im2col buffer and no bounds code.  The
// operands are pre-split x = hi + lo (bf16 each, |x - hi - lo| <= 2^-17 |x|) and each slab issues
// lo*hi + hi*lo + hi*hi into the same TMEM accumulator -- error ~1e-5 of fp32 through the whole UNet
// (measured), at 3 MMAs per product instead of fp32 CUDA-core math.
//
// Reference ops replaced: nn.Conv3d 3^3 / 1^3 stride 1 (and nn.Conv1d k=1) --
// /root/reference/holo_diffusion/guided_diffusion/unet.py:185,211,222,383,392,657,792.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// (TMEM -> registers -> bias/residual -> global, optionally also the bf16 hi/lo split of the result).
#include "common.cuh"
#include "../../include/holo_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

namespace {

constexpr int TILE_W = 8, TILE_H = 4, TILE_D = 4;
constexpr int BLOCK_M = TILE_W * TILE_H * TILE_D;  // 128
constexpr int SLAB = 64;                           // bf16 channels per K slab = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * 128;        // 16 KB
constexpr int NUM_THREADS = 192;

template <int BLOCK_N>
struct Cfg {
    static constexpr int B_TILE_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BLOCK_N <= 64) ? 4 : (BLOCK_N <= 128 ? 3 : 2);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*GN partials*/ + 16 /*last-slice flag*/;
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // [hi*hi + lo*hi | hi*lo]
};

// ---------------------------------------------------------------- PTX wrappers
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (=1024 B between 8-row
// groups) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A format [7,10) and B format
// [10,13) (0 = fp16, 1 = bf16; the hardware faults on a bf16 x fp16 mix), both K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_f16, bool b_f16) {
    return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BLOCK_M >> 4) << 24);
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
// Two TMEM loads (the [hi.hi + lo.hi] and [hi.lo] halves of W accumulator columns) and ONE wait in a single asm
// statement: the loads overlap each other, and no consumer of the registers can be scheduled before the wait.
__device__ __forceinline__ void tmem_ld_pair16(uint32_t ta, uint32_t tb, uint32_t (&a)[16], uint32_t (&b)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
          "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(b[0]), "=r"(b[1]),
          "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]), "=r"(b[10]),
          "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
        : "r"(ta), "r"(tb)
        : "memory");
}
struct TcParams {
    int Cin, D, H, W, ksize, Cout;   // D, H, W = OUTPUT volume
    int tw, th, td;                  // voxel box of one M tile: 8x4x4 (128 rows) or 4x4x4 (64 rows, upper half idle)
    int stride;                      // 1 | 2 (input coordinate = out * stride + tap - pad)
    int iters_per_split;             // split-K: gridDim.z slices of the (tap, slab) loop; atomics into a zeroed out
    int m_tiles, n_blocks, nsplit;   // work-item grid walked by the persistent CTAs
    int chunk;                       // (tap, slab) iterations per TMEM accumulation chain (see "chunked accumulation")
    int stages;                      // pipeline depth actually used (<= Cfg::STAGES); short K loops take fewer stages
                                     // and less shared memory so that several CTAs share an SM
    long long out_pitch;  // elements between consecutive output rows (= Cout for dense tensors)
    const float* bias;
    const float* residual;
    float* out;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    double* stats;  // optional per-output-channel (sum, sumsq) for the GroupNorm that consumes this tensor
    float* partials;     // deterministic split-K: the K slices park their partial tiles here ([slice][tile][128][BLOCK_N])
    int k2_slabs;     // fused 1x1 "skip" operand: Cin2 / 64 extra K iterations after the taps x slabs main loop, reading
                      // the SECOND activation pair at the output voxel itself (no tap offset); 0 = none
    long long* trace; // debug (holo_debug_conv_trace): CTA 0 stamps clock64 at [0] entry, [1] set-up done, [2] first TMA
                      // issued, [3] first operands landed, [4] last MMA issued, [5] first accumulator ready, [6] first item
                      // written, [7] exit
    int fmt;          // 0 = bf16 pairs, HOLO_FMT_F16 = fp16 pairs (all four operand halves)
    float acc_scale;  // accumulators are multiplied by this before bias / residual (undoes the weights' 2^e scale)
    // ---- conv_tc_kernel<N, true> only (holo_gemm_tc_act; appended so that the default instantiations' parameter offsets
    // and code stay what the round's full GPU runs measured)
    int epi_act;      // activation applied after bias / residual: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 Softplus
    unsigned res_mod; // the residual is a (res_mod, Cout) table indexed by output row % res_mod; 0 = row for row
};

template <int BLOCK_N, bool EPI = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const __grid_constant__ CUtensorMap map_a2_hi, const __grid_constant__ CUtensorMap map_a2_lo, TcParams P) {
    // PERSISTENT: each CTA walks work items (M tile, N block, K split) with a stride of gridDim.x.  The TMA producer
    // runs ahead across item boundaries and the accumulator is double-buffered in TMEM, so the epilogue of item i
    // (TMEM -> registers -> global) overlaps the main loop of item i+1 and the ~5 us of per-CTA prologue / epilogue
    // that a one-tile-per-CTA launch pays 14x per SM is paid once.
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_stages = P.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + n_stages * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + n_stages;
    uint64_t* tmem_full_bar = empty_bar + n_stages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_stat = reinterpret_cast<float*>(smem + n_stages * C::STAGE_BYTES + 256);  // [2][BLOCK_N] sum, sumsq

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef HOLO_CONV_TRACE   // debug build only (-DHOLO_CONV_TRACE): the stamps sit in the single-thread issue loops
    const bool tr = P.trace != nullptr && blockIdx.x == 0;
#else
    constexpr bool tr = false;
#endif
    if (tr && threadIdx.x == 0) P.trace[0] = clock64();
    const int tiles_w = P.W / P.tw, tiles_h = P.H / P.th;
    const int rows = P.tw * P.th * P.td;
    const int taps = P.ksize * P.ksize * P.ksize;
    const int pad = P.ksize / 2;
    const int slabs = P.Cin / SLAB;
    const int k_main = taps * slabs;
    const int k_total = k_main + P.k2_slabs;   // the fused skip operand extends the K loop (weights are concatenated)
    const bool split = P.nsplit > 1;
    const int n_items = P.m_tiles * P.n_blocks * P.nsplit;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
        if (P.k2_slabs) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2_lo) : "memory");
        }
        for (int s = 0; s < n_stages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&tmem_full_bar[b], 1), mbar_init(&tmem_empty_bar[b], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tr && threadIdx.x == 0) P.trace[1] = clock64();
    holo_pdl_trigger();   // (opt-in PDL build) the next kernel may be scheduled; it waits for this grid to complete
    holo_pdl_wait();      // everything above overlapped the predecessor's tail; no dependent access before this line

    // work item -> (M tile, N block, K slice); M tiles fastest so that concurrently running CTAs share weights in L2
    auto decode = [&](int item, int& w0, int& h0, int& d0, int& n0, int& it_begin, int& it_end, int& z) {
        const int mt = item % P.m_tiles;
        const int rest = item / P.m_tiles;
        const int nb = rest % P.n_blocks;
        z = rest / P.n_blocks;
        w0 = (mt % tiles_w) * P.tw;
        h0 = ((mt / tiles_w) % tiles_h) * P.th;
        d0 = (mt / (tiles_w * tiles_h)) * P.td;
        n0 = nb * BLOCK_N;
        it_begin = z * P.iters_per_split;
        it_end = min(k_total, it_begin + P.iters_per_split);
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = 2u * (uint32_t)rows * 128u + 2u * (uint32_t)C::B_TILE_BYTES;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int w0, h0, d0, n0, it_begin, it_end, z;
                decode(item, w0, h0, d0, n0, it_begin, it_end, z);
                for (int it = it_begin; it < it_end; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    int kk;
                    if (it < k_main) {
                        const int tap = it / slabs, slab = it % slabs;
                        const int kd = tap / (P.ksize * P.ksize), kh = (tap / P.ksize) % P.ksize, kw = tap % P.ksize;
                        const int c0 = slab * SLAB;
                        const int xw = w0 * P.stride + kw - pad, xh = h0 * P.stride + kh - pad, xd = d0 * P.stride + kd - pad;
                        tma_load_4d(st, &map_a_hi, &full_bar[stage], c0, xw, xh, xd);
                        tma_load_4d(st + A_TILE_BYTES, &map_a_lo, &full_bar[stage], c0, xw, xh, xd);
                        kk = tap * P.Cin + c0;
                    } else {   // fused 1x1 skip operand (stride 1): the box of the output tile itself
                        const int c0 = (it - k_main) * SLAB;
                        tma_load_4d(st, &map_a2_hi, &full_bar[stage], c0, w0, h0, d0);
                        tma_load_4d(st + A_TILE_BYTES, &map_a2_lo, &full_bar[stage], c0, w0, h0, d0);
                        kk = taps * P.Cin + c0;
                    }
                    tma_load_2d(st + 2 * A_TILE_BYTES, &map_b_hi, &full_bar[stage], kk, n0);
                    tma_load_2d(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &map_b_lo, &full_bar[stage], kk, n0);
                    if (tr && item == (int)blockIdx.x && it == it_begin) P.trace[2] = clock64();
                    if (++stage == n_stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // Two UMMAs per K step instead of three: the B tile holds [w_hi rows | w_lo rows] back to back, so
        //   x_hi . [w_hi | w_lo]^T  (N = 2*BLOCK_N)  fills columns [0,BN) with hi*hi and [BN,2BN) with hi*lo,
        //   x_lo . w_hi^T           (N = BLOCK_N)    adds lo*hi to columns [0,BN);
        // the epilogue adds the two halves.  14 KB instead of 18 KB of shared-memory operand reads per K step
        // (the kernel is bound by the SMEM operand bandwidth of the SS-mode UMMA at N = 64).
        const bool f16 = (P.fmt & HOLO_FMT_F16) != 0;   // A and B of one MMA must share the format
        const uint32_t idesc1 = make_idesc(2 * BLOCK_N, f16, f16);    // x_hi . [w_hi | w_lo]
        const uint32_t idesc2 = make_idesc(BLOCK_N, f16, f16);        // x_lo . w_hi
        // CHUNKED ACCUMULATION: the tensor core adds every K = 16 step into the fp32 TMEM accumulator with truncation, a
        // bias that grows linearly with the length of the chain.  The (tap, slab) loop of an
        // item is therefore cut into chains of P.chunk iterations; each chain starts from zero in the other TMEM
        // buffer and the epilogue warps sum the chains in registers (round-to-nearest) while the next chain runs.
        int stage = 0;
        uint32_t phase = 0;
        int cl = 0;   // chains issued by this CTA so far (TMEM buffer = cl & 1)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int w0, h0, d0, n0, it_begin, it_end, z;
            decode(item, w0, h0, d0, n0, it_begin, it_end, z);
            for (int cb = it_begin; cb < it_end; cb += P.chunk, ++cl) {
                const int ce = min(it_end, cb + P.chunk);
                const int buf = cl & 1;
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * C::TMEM_COLS);
                mbar_wait(&tmem_empty_bar[buf], ((cl >> 1) & 1) ^ 1);  // epilogue drained this accumulator
                tc_fence_after();
                for (int it = cb; it < ce; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (tr && lane == 0 && item == (int)blockIdx.x && it == it_begin) P.trace[3] = clock64();
                    if (elect_one()) {
                        const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint32_t a_lo = a_hi + A_TILE_BYTES;
                        const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;  // b_lo follows at + B_TILE_BYTES
#pragma unroll
                        for (int k = 0; k < SLAB / 16; ++k) {
                            const uint32_t ko = k * 32;  // 16 bf16 = 32 bytes inside the 128-byte swizzle row
                            const uint64_t dah = make_kmajor_sw128_desc(a_hi + ko), dal = make_kmajor_sw128_desc(a_lo + ko);
                            const uint64_t dbh = make_kmajor_sw128_desc(b_hi + ko);
                            umma_bf16(d_tmem, dah, dbh, idesc1, (it > cb) || (k != 0));
                            umma_bf16(d_tmem, dal, dbh, idesc2, 1);
                        }
                        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                        if (it == ce - 1) umma_commit(&tmem_full_bar[buf]);
                    }
                    __syncwarp();
                    if (++stage == n_stages) stage = 0, phase ^= 1;
                }
            }
        }
        if (tr && lane == 0) P.trace[4] = clock64();
    } else {
        // ================= epilogue =================
        const int q = warp % 4;  // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        const bool row_ok = r < rows;
        const int et = threadIdx.x - 64;  // 0..127 within the epilogue warps
        const bool ws = split && P.partials != nullptr;   // deterministic split-K through a workspace (see below)
        const bool do_stats = P.stats != nullptr && !split;
        int stat_n0 = -1;
        auto flush_stats = [&]() {   // all 128 epilogue threads: smem partials -> global fp64, then clear
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (stat_n0 >= 0)
                for (int i = et; i < 2 * BLOCK_N; i += 128) {
                    const int which = i / BLOCK_N, c = i % BLOCK_N;
                    atomicAdd(&P.stats[(size_t)(stat_n0 + c) * 2 + which], (double)s_stat[i]);
                }
            for (int i = et; i < 2 * BLOCK_N; i += 128) s_stat[i] = 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        // GroupNorm statistics of the tensor being written: column sums of 16 columns over the warp's 32 rows by a butterfly
        // reduce-scatter (16 values -> one complete column per lane pair), then shared atomics
        auto add_stats = [&](const float (&vals)[16], int c0) {
            float sv[16], qv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = row_ok ? vals[j] : 0.f, qv[j] = sv[j] * sv[j];
            int col = 0;
#pragma unroll
            for (int lvl = 16, n_keep = 8; lvl >= 2; lvl >>= 1, n_keep >>= 1) {
                const bool up = (lane & lvl) != 0;
#pragma unroll
                for (int i = 0; i < n_keep; ++i) {
                    const float ss = up ? sv[i] : sv[i + n_keep], sq = up ? qv[i] : qv[i + n_keep];
                    float ks = up ? sv[i + n_keep] : sv[i], kq = up ? qv[i + n_keep] : qv[i];
                    ks += __shfl_xor_sync(0xffffffffu, ss, lvl);
                    kq += __shfl_xor_sync(0xffffffffu, sq, lvl);
                    sv[i] = ks, qv[i] = kq;
                }
                col += up ? n_keep : 0;
            }
            sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
            qv[0] += __shfl_xor_sync(0xffffffffu, qv[0], 1);
            if ((lane & 1) == 0) {
                atomicAdd(&s_stat[c0 + col], sv[0]);
                atomicAdd(&s_stat[BLOCK_N + c0 + col], qv[0]);
            }
        };
        if (do_stats) flush_stats();  // clears the partials
        int cl = 0;
        float accv[BLOCK_N];   // sum of the finished chains of the current item (used when an item has several)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int w0, h0, d0, n0, it_begin, it_end, z;
            decode(item, w0, h0, d0, n0, it_begin, it_end, z);
            if (do_stats && n0 != stat_n0) {
                if (stat_n0 >= 0) flush_stats();
                stat_n0 = n0;
            }
            const int w = w0 + (r % P.tw), h = h0 + ((r / P.tw) % P.th), d = d0 + r / (P.tw * P.th);
            const size_t v = row_ok ? ((size_t)d * P.H + h) * P.W + w : 0;
            const bool lead = z == 0;
            const int n_chains = (it_end - it_begin + P.chunk - 1) / P.chunk;
            // ---- every chain: TMEM -> registers (summed, round-to-nearest), hand the buffer straight back so that the
            // MMA warp never waits for the global-memory part of the epilogue
            for (int ch = 0; ch < n_chains; ++ch, ++cl) {
                const int buf = cl & 1;
                const uint32_t t_base = tmem_base + (uint32_t)(buf * C::TMEM_COLS) + ((uint32_t)(q * 32) << 16);
                mbar_wait(&tmem_full_bar[buf], (cl >> 1) & 1);
                tc_fence_after();
                if (tr && et == 0 && cl == 0) P.trace[5] = clock64();
                // pairs of 16-column loads with one wait each (32-column pairs measured no faster and cost 60 registers)
#pragma unroll
                for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                    uint32_t acc[16], acc2[16];
                    tmem_ld_pair16(t_base + (uint32_t)c0, t_base + (uint32_t)(BLOCK_N + c0), acc, acc2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float s2 = __uint_as_float(acc[j]) + __uint_as_float(acc2[j]);
                        accv[c0 + j] = ch == 0 ? s2 : accv[c0 + j] + s2;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
            // ---- scale, bias, residual, outputs, statistics of 16 finished columns
            auto emit = [&](float (&vals)[16], int c0, bool add_bias_res, bool atomic) {
                const int n = n0 + c0;
                if (P.bias && add_bias_res) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n) + j4);
                        vals[j4 * 4 + 0] += b.x, vals[j4 * 4 + 1] += b.y, vals[j4 * 4 + 2] += b.z, vals[j4 * 4 + 3] += b.w;
                    }
                }
                if (P.residual && add_bias_res && row_ok) {
                    size_t rrow = v;
                    if constexpr (EPI) rrow = P.res_mod ? (size_t)((unsigned)v % P.res_mod) : v;
                    const float4* rp = reinterpret_cast<const float4*>(P.residual + rrow * P.out_pitch + n);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 b = __ldg(rp + j4);
                        vals[j4 * 4 + 0] += b.x, vals[j4 * 4 + 1] += b.y, vals[j4 * 4 + 2] += b.z, vals[j4 * 4 + 3] += b.w;
                    }
                }
                if constexpr (EPI) {   // activation of the consumer folded into this epilogue (holo_gemm_tc_act)
                    if (P.epi_act == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vals[j] = fmaxf(vals[j], 0.f);
                    } else if (P.epi_act == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vals[j] = holo_leaky(vals[j]);
                    } else if (P.epi_act == 3) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vals[j] = vals[j] > 20.f ? vals[j] : log1pf(expf(vals[j]));
                    }
                }
                if (row_ok) {
                    if (P.out && atomic) {
                        float* op = P.out + v * P.out_pitch + n;
#pragma unroll
                        for (int j = 0; j < 16; ++j) atomicAdd(op + j, vals[j]);
                    } else if (P.out) {
                        float4* op = reinterpret_cast<float4*>(P.out + v * P.out_pitch + n);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            op[j4] = make_float4(vals[j4 * 4], vals[j4 * 4 + 1], vals[j4 * 4 + 2], vals[j4 * 4 + 3]);
                    }
                    if (P.out_hi) {
                        uint32_t hi[8], lo[8];   // the result as an operand pair in the format of this call's operands
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            holo_split2(vals[2 * j], vals[2 * j + 1], (P.fmt & HOLO_FMT_F16) != 0, hi[j], lo[j]);
                        uint4* hp = reinterpret_cast<uint4*>(P.out_hi + v * P.out_pitch + n);
                        uint4* lp = reinterpret_cast<uint4*>(P.out_lo + v * P.out_pitch + n);
                        hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]), hp[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                        lp[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]), lp[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
                if (do_stats && !atomic) add_stats(vals, c0);
            };
            if (!ws) {
#pragma unroll
                for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                    float vals[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) vals[j] = accv[c0 + j] * P.acc_scale;
                    emit(vals, c0, lead, split);   // split without a workspace: fp32 atomics into the zeroed output
                }
            } else {
                // DETERMINISTIC SPLIT-K: every K slice parks its raw partial tile in the workspace ([slice][tile][128][N],
                // plain stores: no zero-fill, no atomics); splitk_reduce_kernel then sums the slices in slice order, applies
                // scale / bias / residual, writes the output once and accumulates the GroupNorm statistics.
                const int mt = item % P.m_tiles, nb = (item / P.m_tiles) % P.n_blocks;
                const size_t tile_floats = (size_t)BLOCK_M * BLOCK_N;
                const size_t tile_id = (size_t)mt * P.n_blocks + nb;
                float* mine = P.partials + ((size_t)z * P.m_tiles * P.n_blocks + tile_id) * tile_floats + (size_t)r * BLOCK_N;
#pragma unroll
                for (int c4 = 0; c4 < BLOCK_N / 4; ++c4)
                    __stcg(reinterpret_cast<float4*>(mine) + c4,
                           make_float4(accv[c4 * 4], accv[c4 * 4 + 1], accv[c4 * 4 + 2], accv[c4 * 4 + 3]));
            }
            if (tr && et == 0 && item == (int)blockIdx.x) P.trace[6] = clock64();
        }
        if (do_stats) flush_stats();
    }
    __syncthreads();
    if (tr && threadIdx.x == 0) P.trace[7] = clock64();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * C::TMEM_COLS));
    }
}

// Second half of the deterministic split-K: out[v][n] = acc_scale * sum_z partial[z][tile(v, n)][row(v)][n % BLOCK_N] + bias
// + residual, in slice order; per-channel (sum, sumsq) of the result for the consumer GroupNorm.  Thread = one voxel row of
// a tile x 4 channels; a warp covers 32 rows of the same 4 channels, so the statistics reduce with shuffles.
struct ReduceParams {
    const float* partials;
    int nsplit, m_tiles, n_blocks, block_n;
    int D, H, W, tw, th, td, rows, Cout;
    long long out_pitch;
    float acc_scale;
    const float* bias;
    const float* residual;
    float* out;
    double* stats;
};
__global__ void __launch_bounds__(256) splitk_reduce_kernel(ReduceParams P) {
    const int lane = threadIdx.x & 31;
    const int c4_per_tile = P.block_n / 4;
    const long long row_groups = (long long)P.m_tiles * P.n_blocks * c4_per_tile * (BLOCK_M / 32);   // warps of work
    const size_t tile_floats = (size_t)BLOCK_M * P.block_n;
    const size_t slice_stride = (size_t)P.m_tiles * P.n_blocks * tile_floats;
    const int tiles_w = P.W / P.tw, tiles_h = P.H / P.th;
    for (long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; wg < row_groups;
         wg += ((long long)gridDim.x * blockDim.x) / 32) {
        const int c4 = (int)(wg % c4_per_tile);
        const int rq = (int)((wg / c4_per_tile) % (BLOCK_M / 32));
        const long long tile_id = wg / ((long long)c4_per_tile * (BLOCK_M / 32));
        const int nb = (int)(tile_id % P.n_blocks), mt = (int)(tile_id / P.n_blocks);
        const int r = rq * 32 + lane;
        const bool row_ok = r < P.rows;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok) {
            const float* base = P.partials + (size_t)tile_id * tile_floats + (size_t)r * P.block_n + c4 * 4;
            // 8 slices in flight per thread (the loads are independent, the adds keep the slice order)
            int z = 0;
            for (; z + 8 <= P.nsplit; z += 8) {
                float4 t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) t[k] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(z + k) * slice_stride));
#pragma unroll
                for (int k = 0; k < 8; ++k) acc.x += t[k].x, acc.y += t[k].y, acc.z += t[k].z, acc.w += t[k].w;
            }
            for (; z < P.nsplit; ++z) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(base + (size_t)z * slice_stride));
                acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
            }
            const int n = nb * P.block_n + c4 * 4;
            acc.x *= P.acc_scale, acc.y *= P.acc_scale, acc.z *= P.acc_scale, acc.w *= P.acc_scale;
            if (P.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n));
                acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
            }
            const int w0 = (mt % tiles_w) * P.tw, h0 = ((mt / tiles_w) % tiles_h) * P.th, d0 = (mt / (tiles_w * tiles_h)) * P.td;
            const int w = w0 + (r % P.tw), h = h0 + ((r / P.tw) % P.th), d = d0 + r / (P.tw * P.th);
            const size_t v = ((size_t)d * P.H + h) * P.W + w;
            if (P.residual) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(P.residual + v * P.out_pitch + n));
                acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
            }
            *reinterpret_cast<float4*>(P.out + v * P.out_pitch + n) = acc;
        }
        if (P.stats) {   // uniform per warp
            float sv[4] = {acc.x, acc.y, acc.z, acc.w};
            const int n = nb * P.block_n + c4 * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float s1 = warp_sum(row_ok ? sv[k] : 0.f), s2 = warp_sum(row_ok ? sv[k] * sv[k] : 0.f);
                if (lane == 0) {
                    atomicAdd(&P.stats[(size_t)(n + k) * 2], (double)s1);
                    atomicAdd(&P.stats[(size_t)(n + k) * 2 + 1], (double)s2);
                }
            }
        }
    }
}

long long* g_conv_trace = nullptr;   // holo_debug_conv_trace

// ---------------------------------------------------------------- host side
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

int make_act_map(CUtensorMap* m, const void* base, int C, long long pitch, int D, int H, int W, int tw, int th, int td,
                 int stride) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
    // with an element stride s the box spans s*t input elements and the TMA keeps every s-th one (t of them)
    cuuint32_t box[4] = {SLAB, (cuuint32_t)(tw * stride), (cuuint32_t)(th * stride), (cuuint32_t)(td * stride)};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, (cuuint32_t)stride};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int make_w_map(CUtensorMap* m, const void* base, int Ktot, long long pitch, int Cout, int block_n) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {SLAB, (cuuint32_t)block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int BLOCK_N, bool EPI = false>
int launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
           const CUtensorMap& a2h, const CUtensorMap& a2l, const TcParams& P, int tiles, int nsplit, cudaStream_t st) {
    auto k = conv_tc_kernel<BLOCK_N, EPI>;
    static bool attr_set = false;   // per instantiation; the driver calls below cost ~10 us of host time per launch
    if (!attr_set) {
        HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BLOCK_N>::SMEM_BYTES),
                  "holo_conv3d_tc");
        attr_set = true;
    }
    TcParams Q = P;
    Q.m_tiles = tiles, Q.n_blocks = P.Cout / BLOCK_N, Q.nsplit = nsplit;
    Q.stages = Q.iters_per_split < Cfg<BLOCK_N>::STAGES ? Q.iters_per_split : Cfg<BLOCK_N>::STAGES;
    const int smem = Q.stages * Cfg<BLOCK_N>::STAGE_BYTES + 1024 + 256 + 1024 + 16;
    // persistent grid: as many CTAs as fit on the chip (shared memory and the 512 TMEM columns bound the residency)
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    int occ_smem = (227 * 1024) / (smem + 1024);
    int occ_tmem = 512 / (2 * Cfg<BLOCK_N>::TMEM_COLS);
    int occ = occ_smem < occ_tmem ? occ_smem : occ_tmem;
    // registers bound the residency too (the epilogue keeps BLOCK_N partial sums per thread); one query per pipeline depth
    static int occ_cache[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int& occ_api = occ_cache[Q.stages & 7];
    if (occ_api == 0 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_api, k, NUM_THREADS, (size_t)smem) != cudaSuccess)
        occ_api = -1;
    if (occ_api > 0 && occ_api < occ) occ = occ_api;
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    const long long items = (long long)Q.m_tiles * Q.n_blocks * Q.nsplit;
    long long grid = (long long)n_sm * occ;
    if (grid > items) grid = items;
    holo_launch(k, dim3((unsigned)grid), dim3(NUM_THREADS), (size_t)smem, st, ah, al, bh, bl, a2h, a2l, Q);
    HOLO_CHECK_LAUNCH("holo_conv3d_tc");
    return HOLO_OK;
}

}  // namespace

extern "C" long long holo_conv3d_tc_splitk_bytes(void) { return 64LL << 20; }

static int conv_tc_impl(const char* who, const void* x_hi, const void* x_lo, int Cin, long long x_pitch, int Din, int Hin,
                        int Win, int ksize, int stride, const void* w_hi, const void* w_lo, long long w_pitch,
                        const float* bias, const float* residual, int Cout, long long out_pitch, float* out,
                        void* out_hi_bf16, void* out_lo_bf16, void* stream, int out_is_zeroed = 0,
                        double* stats = nullptr, int fmt = 0, float acc_scale = 1.0f, const void* x2_hi = nullptr,
                        const void* x2_lo = nullptr, int Cin2 = 0, float* splitk_partials = nullptr, int epi_act = 0,
                        long long res_mod = 0) {
    if (!(x_hi && x_lo && w_hi && w_lo && (out || out_hi_bf16))) {
        holo_set_error("%s: null arg", who);
        return HOLO_ERR_ARG;
    }
    if ((out_hi_bf16 == nullptr) != (out_lo_bf16 == nullptr)) {
        holo_set_error("%s: hi/lo outputs come together", who);
        return HOLO_ERR_ARG;
    }
    const int D = Din / stride, H = Hin / stride, W = Win / stride;  // output volume
    int tw, th, td;
    if (W % 8 == 0 && H % 4 == 0 && D % 4 == 0) tw = 8, th = 4, td = 4;
    else if (W % 4 == 0 && H % 4 == 0 && D % 4 == 0) tw = 4, th = 4, td = 4;
    else tw = 0, th = 0, td = 0;
    const bool stride_ok = stride == 1 || (stride == 2 && ksize == 3 && Din % 2 == 0 && Hin % 2 == 0 && Win % 2 == 0);
    if (fmt & ~HOLO_FMT_F16) {
        holo_set_error("%s: unknown operand_fmt bits 0x%x", who, fmt);
        return HOLO_ERR_ARG;
    }
    if (!(ksize == 1 || ksize == 3) || !stride_ok || tw == 0 || Cin % SLAB || Cout % 16 || x_pitch % 8 || w_pitch % 8 ||
        out_pitch % 4) {
        holo_set_error("%s: unsupported shape Cin=%d Cout=%d dims=%dx%dx%d k=%d stride=%d", who, Cin, Cout, Din, Hin, Win,
                       ksize, stride);
        return HOLO_ERR_UNSUPPORTED;
    }
    int block_n;
    if (Cout % 128 == 0) block_n = 128;
    else if (Cout % 64 == 0) block_n = 64;
    else if (Cout % 32 == 0) block_n = 32;
    else block_n = 16;
    // small volumes: prefer more CTAs over wider tiles
    const int tiles = (D / td) * (H / th) * (W / tw);
    if (block_n == 128 && tiles * (Cout / 128) < 148) block_n = 64;
    // short-K plain GEMMs (the view-pooling encoder's Linear layers: 2 K-steps per output tile) are bound by the epilogue;
    // N = 64 tiles would leave TMEM and registers for two CTAs per SM.  Measured SLOWER (four GEMMs of the encoder: 1.78 vs
    // 1.53 ms, profiles/r02l/encoder_gemm_n64.json): twice the A-operand traffic outweighs the extra epilogue warps.  Off.
    static const int gemm_n64 = [] {
        const char* e = getenv("HOLO_GEMM_SHORTK_N64");
        return e ? atoi(e) : 0;
    }();
    if (gemm_n64 && block_n == 128 && ksize == 1 && Cin2 == 0 && Cin <= 256 && strncmp(who, "holo_gemm", 9) == 0) block_n = 64;
    const int taps = ksize * ksize * ksize;
    if (Cin2 && (!x2_hi || !x2_lo || Cin2 % SLAB || stride != 1)) {
        holo_set_error("%s: the fused skip operand needs both halves, Cin_skip %% 64 == 0 and stride 1", who);
        return HOLO_ERR_ARG;
    }
    const int k_total = taps * (Cin / SLAB) + Cin2 / SLAB;
    // split-K: when the M x N grid cannot fill the 148 SMs and K is long, slice the (tap, slab) loop across
    // gridDim.z and accumulate with fp32 atomics into a zeroed output (>= 4 iterations per slice)
    int nsplit = 1, per = k_total;
    const int base = tiles * (Cout / block_n);
    const bool epi = epi_act != 0 || res_mod != 0;   // activation / row-table epilogue: one slice owns the whole K loop
    if (base < 120 && k_total >= 8 && out && !out_hi_bf16 && !epi && (out_pitch == Cout || out_is_zeroed)) {
        static const int target = [] {   // work items aimed at (2 per SM); HOLO_SPLITK_TARGET for tuning
            const char* e = getenv("HOLO_SPLITK_TARGET");
            const int v = e ? atoi(e) : 296;
            return v < 148 ? 148 : v;
        }();
        int want = (target + base - 1) / base;
        int maxs = k_total / 4;
        nsplit = want < maxs ? want : maxs;
        if (nsplit < 1) nsplit = 1;
        per = (k_total + nsplit - 1) / nsplit;
        nsplit = (k_total + per - 1) / per;
    }
    // Upper bound on the slices of the workspace path (HOLO_SPLITK_MAX; the reduce launch reads every slice once: 24 vs 40
    // slices measured 9.31 vs 9.27 ms per step, profiles/r02g)
    static const int splitk_max = [] {
        const char* e = getenv("HOLO_SPLITK_MAX");
        const int v = e ? atoi(e) : 64;
        return v < 2 ? 2 : v;
    }();
    if (nsplit > splitk_max && splitk_partials) {
        per = (k_total + splitk_max - 1) / splitk_max;
        nsplit = (k_total + per - 1) / per;
    }
    CUtensorMap ah, al, bh, bl, a2h, a2l;
    int e = make_act_map(&ah, x_hi, Cin, x_pitch, Din, Hin, Win, tw, th, td, stride);
    if (!e) e = make_act_map(&al, x_lo, Cin, x_pitch, Din, Hin, Win, tw, th, td, stride);
    if (!e) e = make_w_map(&bh, w_hi, taps * Cin + Cin2, w_pitch, Cout, block_n);
    if (!e) e = make_w_map(&bl, w_lo, taps * Cin + Cin2, w_pitch, Cout, block_n);
    if (Cin2) {
        if (!e) e = make_act_map(&a2h, x2_hi, Cin2, Cin2, Din, Hin, Win, tw, th, td, 1);
        if (!e) e = make_act_map(&a2l, x2_lo, Cin2, Cin2, Din, Hin, Win, tw, th, td, 1);
    } else {
        a2h = ah, a2l = al;   // never dereferenced (k2_slabs = 0)
    }
    if (e) {
        holo_set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, e);
        return HOLO_ERR_CUDA;
    }
    TcParams P;
    P.Cin = Cin, P.D = D, P.H = H, P.W = W, P.ksize = ksize, P.Cout = Cout, P.out_pitch = out_pitch;
    P.tw = tw, P.th = th, P.td = td, P.stride = stride, P.iters_per_split = per;
    P.bias = bias, P.residual = residual, P.out = out;
    P.out_hi = (__nv_bfloat16*)out_hi_bf16, P.out_lo = (__nv_bfloat16*)out_lo_bf16;
    // split-K through the caller's workspace (deterministic, no zero-fill, statistics and operand pairs from the last
    // slice) when it is given and large enough; otherwise fp32 atomics into a zeroed output
    const long long ws_need = (long long)nsplit * tiles * (Cout / block_n) * BLOCK_M * block_n * (long long)sizeof(float);
    const bool ws = nsplit > 1 && splitk_partials && ws_need <= holo_conv3d_tc_splitk_bytes() && out && !out_hi_bf16;
    P.partials = ws ? splitk_partials : nullptr;
    P.trace = g_conv_trace;
    P.stats = (nsplit == 1 && out_pitch == Cout) ? stats : nullptr;
    P.fmt = fmt, P.acc_scale = acc_scale, P.k2_slabs = Cin2 / SLAB;
    P.epi_act = epi_act, P.res_mod = (unsigned)res_mod;
    if (epi && (block_n < 64 || nsplit > 1 || epi_act < 0 || epi_act > 3 || res_mod < 0 || res_mod >= (1LL << 31) ||
                (long long)D * H * W >= (1LL << 32))) {
        holo_set_error("%s: the fused activation / row-table epilogue takes Cout %% 64 == 0, un-split K, act 0..3 (got Cout=%d "
                       "act=%d slices=%d)", who, Cout, epi_act, nsplit);
        return HOLO_ERR_UNSUPPORTED;
    }
    // chunked accumulation (see the MMA issuer): chains of ~HOLO_CONV_CHUNK (tap, slab) iterations, 0 = one chain per item
    // (default 9 = three chains for a 27-tap x 1-slab item), balanced so that no short tail chain is left: every chain
    // boundary costs ~0.3 us of tensor-pipe time (the TMEM -> register flush shares the TMEM port, profiles/r02b)
    static const int chunk_env = [] {
        const char* e = getenv("HOLO_CONV_CHUNK");
        return e ? atoi(e) : 9;
    }();
    if (chunk_env > 0 && per > chunk_env) {
        const int n_chains = (per + chunk_env - 1) / chunk_env;
        P.chunk = (per + n_chains - 1) / n_chains;
    } else {
        P.chunk = 1 << 30;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (nsplit > 1 && !ws && !out_is_zeroed)
        HOLO_CUDA(cudaMemsetAsync(out, 0, (size_t)D * H * W * Cout * sizeof(float), st), who);
    int rc;
    if (epi) {
        rc = block_n == 128 ? launch<128, true>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st)
                            : launch<64, true>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st);
    } else switch (block_n) {
        case 128: rc = launch<128>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st); break;
        case 64: rc = launch<64>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st); break;
        case 32: rc = launch<32>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st); break;
        default: rc = launch<16>(ah, al, bh, bl, a2h, a2l, P, tiles, nsplit, st); break;
    }
    if (rc == HOLO_OK && ws) {
        ReduceParams R;
        R.partials = splitk_partials, R.nsplit = nsplit, R.m_tiles = tiles, R.n_blocks = Cout / block_n, R.block_n = block_n;
        R.D = D, R.H = H, R.W = W, R.tw = tw, R.th = th, R.td = td, R.rows = tw * th * td, R.Cout = Cout;
        R.out_pitch = out_pitch, R.acc_scale = acc_scale, R.bias = bias, R.residual = residual, R.out = out;
        R.stats = out_pitch == Cout ? stats : nullptr;
        const long long warps = (long long)tiles * (Cout / block_n) * (block_n / 4) * (BLOCK_M / 32);
        long long blocks = (warps + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(R);
        HOLO_CHECK_LAUNCH(who);
        return (stats && !R.stats) ? 1 : HOLO_OK;
    }
    if (rc == HOLO_OK && stats && !P.stats) return 1;  // done, but the statistics were not produced (split-K / pitch)
    return rc;
}

int holo_conv3d_tc_halo(const void* x_hi, const void* x_lo, int Cin, int D, int H, int W, const void* w_hi,
                        const void* w_lo, const float* bias, const float* residual, int Cout, float* out,
                        void* out_hi, void* out_lo, cudaStream_t st);

extern "C" int holo_conv3d_tc(const void* x_hi, const void* x_lo, int Cin, int D, int H, int W, int ksize, int stride,
                              const void* w_hi, const void* w_lo, const float* bias, const float* residual, int Cout,
                              float* out, void* out_hi_bf16, void* out_lo_bf16, double* stats_ch, int operand_fmt,
                              float acc_scale, float* splitk_partials, void* stream) {
    const int taps = ksize * ksize * ksize;
    // Optional (HOLO_CONV_HALO=1): halo-resident activation tile (conv_tc_halo.cu), 3x less L2->SMEM traffic.
    // Measured on B200 it ties the tap-reload kernel before and loses to it after that kernel became persistent
    // (both are bound by the SS-mode UMMA operand fetch at N = 64, not by L2), so it is off by default.
    static const bool no_halo = getenv("HOLO_CONV_HALO") == nullptr;
    if (!no_halo && !stats_ch && operand_fmt == 0 && acc_scale == 1.0f && ksize == 3 && stride == 1 && x_hi && x_lo && w_hi && w_lo && (out || out_hi_bf16) &&
        (out_hi_bf16 == nullptr) == (out_lo_bf16 == nullptr) && W % 8 == 0 && H % 16 == 0 && D % 2 == 0 &&
        Cout % 64 == 0 && Cin % 64 == 0 && (long long)(W / 8) * (H / 16) * (D / 2) * (Cout / 64) >= 120) {
        int rc = holo_conv3d_tc_halo(x_hi, x_lo, Cin, D, H, W, w_hi, w_lo, bias, residual, Cout, out, out_hi_bf16,
                                     out_lo_bf16, (cudaStream_t)stream);
        if (rc != HOLO_ERR_UNSUPPORTED) return rc;
    }
    return conv_tc_impl("holo_conv3d_tc", x_hi, x_lo, Cin, Cin, D, H, W, ksize, stride, w_hi, w_lo,
                        (long long)taps * Cin, bias, residual, Cout, Cout, out, out_hi_bf16, out_lo_bf16, stream, 0,
                        stats_ch, operand_fmt, acc_scale, nullptr, nullptr, 0, splitk_partials);
}



// ResBlock tail in one launch: out = conv3^3(x) + conv1^1(skip_x) + bias (+ residual): the 1x1 skip connection
// (unet.py:222,255) rides the same TMEM accumulator as Cin_skip / 64 extra K iterations, so its fp32 result is never
// written nor re-read.  Weights: [Cout][27 * Cin + Cin_skip] pairs (the 3^3 taps, then the 1x1 columns), one common
// scale; bias = conv bias + skip bias (the caller adds them).
extern "C" int holo_conv3d_tc_skip(const void* x_hi, const void* x_lo, int Cin, const void* skip_hi, const void* skip_lo,
                                   int Cin_skip, int D, int H, int W, const void* w_hi, const void* w_lo,
                                   const float* bias, const float* residual, int Cout, float* out, double* stats_ch,
                                   int operand_fmt, float acc_scale, float* splitk_partials, void* stream) {
    if (!skip_hi || !skip_lo || Cin_skip <= 0) {
        holo_set_error("holo_conv3d_tc_skip: the skip operand is missing");
        return HOLO_ERR_ARG;
    }
    return conv_tc_impl("holo_conv3d_tc_skip", x_hi, x_lo, Cin, Cin, D, H, W, 3, 1, w_hi, w_lo,
                        27LL * Cin + Cin_skip, bias, residual, Cout, Cout, out, nullptr, nullptr, stream, 0, stats_ch,
                        operand_fmt, acc_scale, skip_hi, skip_lo, Cin_skip, splitk_partials);
}

// Plain GEMM on the same kernel: out[m][n] = bias[n] + residual[m][n] + sum_k a[m][k] * b[n][k]  (both K-major,
// bf16 hi/lo pairs, arbitrary row pitches).  M % 128 == 0, K % 64 == 0, N % 16 == 0.
extern "C" int holo_gemm_tc(const void* a_hi, const void* a_lo, long long a_pitch, int M, int K, const void* b_hi,
                            const void* b_lo, long long b_pitch, int N, const float* bias, const float* residual,
                            long long out_pitch, float* out, void* out_hi_bf16, void* out_lo_bf16, int out_is_zeroed,
                            int operand_fmt, float acc_scale, void* stream) {
    if (M % BLOCK_M) {
        holo_set_error("holo_gemm_tc: M=%d must be a multiple of 128", M);
        return HOLO_ERR_UNSUPPORTED;
    }
    return conv_tc_impl("holo_gemm_tc", a_hi, a_lo, K, a_pitch, M / 32, TILE_H, TILE_W, 1, 1, b_hi, b_lo, b_pitch, bias,
                        residual, N, out_pitch, out, out_hi_bf16, out_lo_bf16, stream, out_is_zeroed, nullptr, operand_fmt,
                        acc_scale);
}

// holo_gemm_tc with the consumer's elementwise work folded into the epilogue:
//   out[m][n] = act(acc_scale * sum_k a[m][k] b[n][k] + bias[n] + row_term[m % row_term_rows][n])
// written as fp32 and / or as the next GEMM's operand pair.  The view-pooling encoder's Linear layers use it
// (custom_modules.py:255-264: the per-point mean term is shared by the point's rows of every view).
extern "C" int holo_gemm_tc_act(const void* a_hi, const void* a_lo, long long a_pitch, int M, int K, const void* b_hi,
                                const void* b_lo, long long b_pitch, int N, const float* bias, const float* row_term,
                                long long row_term_rows, int act, long long out_pitch, float* out, void* out_hi,
                                void* out_lo, int operand_fmt, float acc_scale, void* stream) {
    if (M % BLOCK_M) {
        holo_set_error("holo_gemm_tc_act: M=%d must be a multiple of 128", M);
        return HOLO_ERR_UNSUPPORTED;
    }
    if (row_term && row_term_rows <= 0) {
        holo_set_error("holo_gemm_tc_act: row_term needs row_term_rows > 0");
        return HOLO_ERR_ARG;
    }
    if (!row_term && !act)   // nothing to fold: the plain kernel
        return conv_tc_impl("holo_gemm_tc_act", a_hi, a_lo, K, a_pitch, M / 32, TILE_H, TILE_W, 1, 1, b_hi, b_lo, b_pitch, bias,
                            nullptr, N, out_pitch, out, out_hi, out_lo, stream, 0, nullptr, operand_fmt, acc_scale);
    return conv_tc_impl("holo_gemm_tc_act", a_hi, a_lo, K, a_pitch, M / 32, TILE_H, TILE_W, 1, 1, b_hi, b_lo, b_pitch, bias,
                        row_term, N, out_pitch, out, out_hi, out_lo, stream, 0, nullptr, operand_fmt, acc_scale,
                        nullptr, nullptr, 0, nullptr, act, row_term ? row_term_rows : 0);
}

// Debug: subsequent convolution launches make CTA 0 write 8 clock64 stamps (TcParams::trace) into dev_buf8 (8 int64 on
// the device); NULL switches it off.  Used by tools/conv_micro.py to see where a small launch spends its time.
extern "C" int holo_debug_conv_trace(void* dev_buf8) {
    g_conv_trace = reinterpret_cast<long long*>(dev_buf8);
    return HOLO_OK;
}
