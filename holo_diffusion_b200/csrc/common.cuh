// Shared helpers for the holo_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define HOLO_OK 0
#define HOLO_ERR_ARG -1
#define HOLO_ERR_CUDA -2
#define HOLO_ERR_UNSUPPORTED -3

// thread-local last error text, exported through holo_last_error()
void holo_set_error(const char* fmt, ...);

#define HOLO_CHECK_ARG(cond, ...)          \
    do {                                   \
        if (!(cond)) {                     \
            holo_set_error(__VA_ARGS__);   \
            return HOLO_ERR_ARG;           \
        }                                  \
    } while (0)

#define HOLO_CHECK_LAUNCH(name)                                                       \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            holo_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
            return HOLO_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

#define HOLO_CUDA(call, name)                                                         \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            holo_set_error("%s: %s", name, cudaGetErrorString(e__));                  \
            return HOLO_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

static inline int holo_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (PDL), switched on at run time with HOLO_PDL=1.  A kernel launched through
// holo_launch may start while its predecessor in the stream is still running; it therefore executes holo_pdl_wait()
// -- which returns when every prerequisite grid has completed and flushed its memory -- before its first dependent
// global access (reads AND writes: the caching allocator recycles buffers), and only its prologue (barrier init, TMEM
// allocation, tensor-map prefetch) and its scheduling overlap the predecessor's tail.  Aimed at the launches of the
// coarse UNet levels, which keep 16-64 CTAs on 148 SMs for 10-25 us each.  Without the launch attribute both
// griddepcontrol instructions are no-ops.
#include <cstdlib>
#include <utility>
__device__ __forceinline__ void holo_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void holo_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch of a kernel that contains holo_pdl_wait().  Errors surface through cudaGetLastError (HOLO_CHECK_LAUNCH).
template <typename... KArgs, typename... Args>
static inline void holo_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                               Args&&... args) {
    static const bool on = [] {
        const char* e = getenv("HOLO_PDL");
        return e && e[0] == '1';
    }();
    if (on) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        return;
    }
    kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float holo_silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float holo_leaky(float x) { return x > 0.0f ? x : 0.2f * x; }

// Operand split x = hi + lo of the tensor-core kernels, two values per 32-bit word.
//   bf16 pair: fp32's range, hi + lo exact to 2^-17 |x|                      ("3xBF16")
//   fp16 pair (pair_f16): hi + lo exact to 2^-22 |x| for 6e-5 <= |x| (absolute 3e-8 below), same MMA rate -- what the
//     convolutions use: through the whole UNet the distance to exact arithmetic drops from 7e-5 to 4e-6, i.e. to
//     fp32's own.  Range: conversions saturate, |x| <= 131008 is representable (hi and lo both at 65504); a tcgen05
//     kind::f16 MMA wants A and B in the SAME 16-bit format (bf16 x fp16 raises an illegal-instruction fault), so a
//     bf16 hi half for range is not an option.
__device__ __forceinline__ void holo_split2(float a, float b, bool pair_f16, uint32_t& h, uint32_t& l) {
    if (pair_f16) {
        const float m = 65504.f;
        const __half2 hh = __floats2half2_rn(fminf(fmaxf(a, -m), m), fminf(fmaxf(b, -m), m));
        h = *reinterpret_cast<const uint32_t*>(&hh);
        const float2 hf = __half22float2(hh);
        const __half2 lh = __floats2half2_rn(fminf(fmaxf(a - hf.x, -m), m), fminf(fmaxf(b - hf.y, -m), m));
        l = *reinterpret_cast<const uint32_t*>(&lh);
    } else {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
        h = *reinterpret_cast<const uint32_t*>(&hb);
        const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xffff0000u);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(ra, rb);
        l = *reinterpret_cast<const uint32_t*>(&lb);
    }
}

__device__ __forceinline__ void holo_split1(float a, bool pair_f16, uint16_t& h, uint16_t& l) {
    uint32_t h2, l2;
    holo_split2(a, 0.f, pair_f16, h2, l2);
    h = (uint16_t)(h2 & 0xffffu), l = (uint16_t)(l2 & 0xffffu);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
