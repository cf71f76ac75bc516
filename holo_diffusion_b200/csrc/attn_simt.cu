// fp32 CUDA-core streaming-softmax attention over voxel tokens, head-major "legacy" qkv layout.
//
// Reference op replaced: QKVAttentionLegacy.forward, /root/reference/holo_diffusion/guided_diffusion/unet.py:438-455
//   qkv (N, H*3*ch, T) -> per head [q | k | v]; weight = softmax((q*s)^T (k*s)), s = ch^-1/4, fp32 softmax; a = weight v.
// Here qkv is channels-last (T, H*3*ch) and the output is channels-last (T, H*ch).  The T x T matrix is never
// materialised (the reference writes it twice per block, 134 MB at T = 4096).
#include "common.cuh"
#include "../../include/holo_b200.h"

namespace {
constexpr int BQ = 32, BKV = 32, NT = 256;

template <int CH>
__global__ void __launch_bounds__(NT) attn_simt_kernel(const float* __restrict__ qkv, int T, int heads,
                                                        float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                       // [BQ][CH+1]
    float* Ks = Qs + BQ * (CH + 1);       // [BKV][CH+1]
    float* Vs = Ks + BKV * (CH + 1);      // [BKV][CH]
    float* Ss = Vs + BKV * CH;            // [BQ][BKV+1]
    float* s_alpha = Ss + BQ * (BKV + 1); // [BQ]
    float* s_m = s_alpha + BQ;            // [BQ]
    float* s_l = s_m + BQ;                // [BQ]
    const int tid = threadIdx.x;
    const int h = blockIdx.y;
    const int q0 = blockIdx.x * BQ;
    const int stride = heads * 3 * CH;
    const float scale = 1.0f / sqrtf(sqrtf((float)CH));
    const float* qb = qkv + (size_t)h * 3 * CH;
    for (int i = tid; i < BQ * CH; i += NT) {
        int r = i / CH, c = i % CH;
        Qs[r * (CH + 1) + c] = (q0 + r < T) ? qb[(size_t)(q0 + r) * stride + c] * scale : 0.f;
    }
    if (tid < BQ) s_m[tid] = -INFINITY, s_l[tid] = 0.f;
    // O tile: thread owns query oq and channels oc + 8*i
    constexpr int OPT = CH / 8;
    const int oq = tid / 8, oc = tid % 8;
    float o[OPT];
#pragma unroll
    for (int i = 0; i < OPT; ++i) o[i] = 0.f;
    // S tile: thread owns (sq0, sq0+1) x (sk0, sk0+1)
    const int sq0 = (tid / 16) * 2, sk0 = (tid % 16) * 2;
    __syncthreads();
    for (int k0 = 0; k0 < T; k0 += BKV) {
        for (int i = tid; i < BKV * CH; i += NT) {
            int r = i / CH, c = i % CH;
            bool ok = k0 + r < T;
            Ks[r * (CH + 1) + c] = ok ? qb[(size_t)(k0 + r) * stride + CH + c] * scale : 0.f;
            Vs[r * CH + c] = ok ? qb[(size_t)(k0 + r) * stride + 2 * CH + c] : 0.f;
        }
        __syncthreads();
        {
            float s00 = 0, s01 = 0, s10 = 0, s11 = 0;
            const float* qa = Qs + sq0 * (CH + 1);
            const float* qb2 = qa + (CH + 1);
            const float* ka = Ks + sk0 * (CH + 1);
            const float* kb = ka + (CH + 1);
#pragma unroll 8
            for (int c = 0; c < CH; ++c) {
                float a = qa[c], b = qb2[c], x = ka[c], y = kb[c];
                s00 = fmaf(a, x, s00), s01 = fmaf(a, y, s01), s10 = fmaf(b, x, s10), s11 = fmaf(b, y, s11);
            }
            bool v0 = k0 + sk0 < T, v1 = k0 + sk0 + 1 < T;
            Ss[sq0 * (BKV + 1) + sk0] = v0 ? s00 : -INFINITY;
            Ss[sq0 * (BKV + 1) + sk0 + 1] = v1 ? s01 : -INFINITY;
            Ss[(sq0 + 1) * (BKV + 1) + sk0] = v0 ? s10 : -INFINITY;
            Ss[(sq0 + 1) * (BKV + 1) + sk0 + 1] = v1 ? s11 : -INFINITY;
        }
        __syncthreads();
        {
            // online softmax: warp w handles rows w*4 .. w*4+3, lane = key
            int w = tid / 32, lane = tid % 32;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                int r = w * 4 + rr;
                float s = Ss[r * (BKV + 1) + lane];
                float mx = s;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                float m_old = s_m[r];
                float m_new = fmaxf(m_old, mx);
                float p = (s == -INFINITY) ? 0.f : expf(s - m_new);
                float ps = warp_sum(p);
                Ss[r * (BKV + 1) + lane] = p;
                if (lane == 0) {
                    float alpha = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
                    s_alpha[r] = alpha;
                    s_m[r] = m_new;
                    s_l[r] = s_l[r] * alpha + ps;
                }
            }
        }
        __syncthreads();
        {
            float alpha = s_alpha[oq];
#pragma unroll
            for (int i = 0; i < OPT; ++i) o[i] *= alpha;
            for (int k = 0; k < BKV; ++k) {
                float p = Ss[oq * (BKV + 1) + k];
                const float* vr = Vs + k * CH + oc;
#pragma unroll
                for (int i = 0; i < OPT; ++i) o[i] = fmaf(p, vr[8 * i], o[i]);
            }
        }
        __syncthreads();
    }
    if (q0 + oq < T) {
        float inv = 1.0f / s_l[oq];
        float* op = out + (size_t)(q0 + oq) * heads * CH + h * CH + oc;
#pragma unroll
        for (int i = 0; i < OPT; ++i) op[8 * i] = o[i] * inv;
    }
}

template <int CH>
int launch_attn(const float* qkv, int T, int heads, float* out, cudaStream_t st) {
    size_t smem = (size_t)(BQ * (CH + 1) + BKV * (CH + 1) + BKV * CH + BQ * (BKV + 1) + 3 * BQ) * sizeof(float);
    auto k = attn_simt_kernel<CH>;
    HOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "holo_attention_simt");
    k<<<dim3(holo_cdiv(T, BQ), heads), NT, smem, st>>>(qkv, T, heads, out);
    HOLO_CHECK_LAUNCH("holo_attention_simt");
    return HOLO_OK;
}
}  // namespace

extern "C" int holo_attention_simt(const float* qkv_cl, int T, int heads, int ch, float* out_cl, void* stream) {
    HOLO_CHECK_ARG(qkv_cl && out_cl && T > 0 && heads > 0, "holo_attention_simt: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    switch (ch) {
        case 8: return launch_attn<8>(qkv_cl, T, heads, out_cl, st);
        case 16: return launch_attn<16>(qkv_cl, T, heads, out_cl, st);
        case 32: return launch_attn<32>(qkv_cl, T, heads, out_cl, st);
        case 64: return launch_attn<64>(qkv_cl, T, heads, out_cl, st);
        case 128: return launch_attn<128>(qkv_cl, T, heads, out_cl, st);
        case 256: return launch_attn<256>(qkv_cl, T, heads, out_cl, st);
        default:
            holo_set_error("holo_attention_simt: unsupported head width %d", ch);
            return HOLO_ERR_UNSUPPORTED;
    }
}
