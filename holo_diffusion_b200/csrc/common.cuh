// Shared helpers for the holo_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
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

__device__ __forceinline__ float holo_silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float holo_leaky(float x) { return x > 0.0f ? x : 0.2f * x; }

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
