// Helpers that turn QKVAttentionLegacy (/root/reference/holo_diffusion/guided_diffusion/unet.py:438-455) into
// tensor-core GEMMs on holo_gemm_tc:   S = Q K^T  ->  P = softmax(s^2 S)  ->  O = P V,
// with s = ch^-1/4 folded into the softmax argument ((q s)(k s) = s^2 q k).  The fp32 softmax runs over one
// full key row per CTA and emits P directly as the hi/lo pair (bf16 or fp16 halves) the second GEMM consumes.
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/holo_b200.h"

namespace {

// one CTA per query row; T <= 16384
__global__ void __launch_bounds__(256) softmax_split_kernel(const float* __restrict__ S, int T, float scale2,
                                                            uint16_t* __restrict__ P_hi,
                                                            uint16_t* __restrict__ P_lo, int pair_f16, float p_scale) {
    extern __shared__ float row[];
    __shared__ float red[8];
    const size_t base = (size_t)blockIdx.x * T;
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    float mx = -INFINITY;
    for (int i = tid * 4; i < T; i += 256 * 4) {
        float4 v = *reinterpret_cast<const float4*>(S + base + i);
        v.x *= scale2, v.y *= scale2, v.z *= scale2, v.w *= scale2;
        *reinterpret_cast<float4*>(row + i) = v;
        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid * 4; i < T; i += 256 * 4) {
        float4 v = *reinterpret_cast<float4*>(row + i);
        v.x = expf(v.x - mx), v.y = expf(v.y - mx), v.z = expf(v.z - mx), v.w = expf(v.w - mx);
        *reinterpret_cast<float4*>(row + i) = v;
        sum += (v.x + v.y) + (v.z + v.w);
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float inv = p_scale / sum;   // p_scale (a power of two) lifts small probabilities out of fp16's subnormals
    for (int i = tid * 4; i < T; i += 256 * 4) {
        float4 v = *reinterpret_cast<float4*>(row + i);
        uint32_t h0, l0, h1, l1;
        holo_split2(v.x * inv, v.y * inv, pair_f16 != 0, h0, l0);
        holo_split2(v.z * inv, v.w * inv, pair_f16 != 0, h1, l1);
        *reinterpret_cast<uint2*>(P_hi + base + i) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(P_lo + base + i) = make_uint2(l0, l1);
    }
}

// src (rows, cols) fp32 with row pitch -> dst_hi/lo (cols, rows) 16-bit halves: V (T, ch) -> V^T (ch, T), K-major for
// the PV GEMM
__global__ void transpose_split_kernel(const float* __restrict__ src, long long src_pitch, int rows, int cols,
                                       uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int pair_f16) {
    __shared__ float tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * src_pitch + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) {
            holo_split1(tile[threadIdx.x][i], pair_f16 != 0, hi[(size_t)c * rows + r], lo[(size_t)c * rows + r]);
        }
    }
}

}  // namespace

extern "C" int holo_softmax_split(const float* S, int n_rows, int T, float scale2, void* P_hi_bf16, void* P_lo_bf16,
                                  int pair_f16, float p_scale, void* stream) {
    HOLO_CHECK_ARG(S && P_hi_bf16 && P_lo_bf16 && n_rows > 0 && T > 0 && T % 4 == 0 && T <= 16384,
                   "holo_softmax_split: T must be a multiple of 4 and <= 16384");
    size_t smem = (size_t)T * sizeof(float);
    if (smem > 48 * 1024)
        HOLO_CUDA(cudaFuncSetAttribute(softmax_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "holo_softmax_split");
    softmax_split_kernel<<<n_rows, 256, smem, (cudaStream_t)stream>>>(S, T, scale2, (uint16_t*)P_hi_bf16,
                                                                      (uint16_t*)P_lo_bf16, pair_f16, p_scale);
    HOLO_CHECK_LAUNCH("holo_softmax_split");
    return HOLO_OK;
}

extern "C" int holo_transpose_split_bf16(const float* src, long long src_pitch, int rows, int cols, void* hi_bf16,
                                         void* lo_bf16, int pair_f16, void* stream) {
    HOLO_CHECK_ARG(src && hi_bf16 && lo_bf16 && rows > 0 && cols > 0, "holo_transpose_split_bf16: bad args");
    dim3 grid(holo_cdiv(cols, 32), holo_cdiv(rows, 32));
    transpose_split_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, src_pitch, rows, cols,
                                                                          (uint16_t*)hi_bf16, (uint16_t*)lo_bf16, pair_f16);
    HOLO_CHECK_LAUNCH("holo_transpose_split_bf16");
    return HOLO_OK;
}
