// Whole-graph denoiser behind the C-ABI: holo_unet_create / set_param / pack / fwd.
//
// A native (C++) executor of guided_diffusion's UNetModel(dims=3, use_scale_shift_norm=True, resblock_updown=False,
// conv_resample=True, homogeneous_resample=True, num_head_channels=-1) --
// /root/reference/holo_diffusion/guided_diffusion/unet.py:566-837 as configured by SimpleUnet3D,
// /root/reference/holo_diffusion/utils/diffusion_utils.py:41-86 -- over the kernels of this library: the same launch
// sequence as the Python executor (holo_diffusion_b200/unet.py::UNetExecutor), so that a host with no Python can run
// the denoiser, and a Python host spends ~1 ms instead of ~12 ms of interpreter time per evaluation.
//
//   holo_unet_create(cfg)            builds the block list, the parameter table (names = the reference's state-dict
//                                    keys below `_net.`) and, by a dry run, the workspace plan
//   holo_unet_set_param(name, ptr)   fp32 device pointers of the checkpoint tensors (borrowed, not copied)
//   holo_unet_pack(packed, stream)   operand-pair weights ([Cout][tap][Cin_pad] fp16/bf16 hi/lo of 2^e w), CUDA-core
//                                    layouts, concatenated FiLM projection, sinusoid frequencies -> caller's buffer
//   holo_unet_fwd[_cl](x, t, out, workspace, stream)   one evaluation; asynchronous, CUDA-graph capturable
//
// The library allocates no device memory: `packed` and `workspace` are the caller's (sizes from *_bytes()).
#include "common.cuh"
#include "../../include/holo_b200.h"

#include <cmath>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr int GN_GROUPS = 32;
constexpr float GN_EPS = 1e-5f;   // nn.GroupNorm default (normalization(), nn.py:92-99)

// ---------------------------------------------------------------- pack kernels
// w (Cout, Cin, taps) fp32 -> hi / lo (Cout_p, taps, Cin_pad) 16-bit pair of scale * w, K-major rows for TMA / UMMA.
// Head padding (attention heads narrower than 64 channels run zero-padded): mode 1 = output rows are (head, q|k|v, c)
// with c < chp and map to the source row (head, q|k|v, c) for c < ch; mode 2 = input channels are (head, c).
struct PadSpec {
    int mode, heads, ch, chp;
};
__device__ __forceinline__ int src_row(int r, const PadSpec& p) {
    if (p.mode != 1) return r;
    const int c = r % p.chp, j = (r / p.chp) % 3, h = r / (3 * p.chp);
    return c < p.ch ? (h * 3 + j) * p.ch + c : -1;
}
__device__ __forceinline__ int src_col(int k, const PadSpec& p) {
    if (p.mode != 2) return k;
    const int c = k % p.chp, h = k / p.chp;
    return c < p.ch ? h * p.ch + c : -1;
}
__global__ void pack_pairs_kernel(const float* __restrict__ w, int Cout_src, int Cin_src, int taps, int Cout_p, int Cin_p,
                                  int Cin_pad, long long row_pitch, long long col_off, float scale, int pair_f16,
                                  PadSpec pad, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const long long total = (long long)Cout_p * taps * Cin_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % Cin_pad), tap = (int)((i / Cin_pad) % taps), r = (int)(i / ((long long)Cin_pad * taps));
        float v = 0.f;
        if (k < Cin_p) {
            const int sr = src_row(r, pad), sk = src_col(k, pad);
            if (sr >= 0 && sk >= 0) v = w[((size_t)sr * Cin_src + sk) * taps + tap] * scale;
        }
        uint16_t h, l;
        holo_split1(v, pair_f16 != 0, h, l);
        const size_t o = (size_t)r * row_pitch + col_off + (size_t)tap * Cin_pad + k;
        hi[o] = h, lo[o] = l;
    }
}
// w (Cout, Cin, taps) -> (taps, Cin_p, Cout_p) fp32 for the CUDA-core kernels
__global__ void pack_simt_kernel(const float* __restrict__ w, int Cin_src, int taps, int Cout_p, int Cin_p, PadSpec pad,
                                 float* __restrict__ out) {
    const long long total = (long long)taps * Cin_p * Cout_p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % Cout_p), k = (int)((i / Cout_p) % Cin_p), tap = (int)(i / ((long long)Cout_p * Cin_p));
        const int sr = src_row(r, pad), sk = src_col(k, pad);
        out[i] = (sr >= 0 && sk >= 0) ? w[((size_t)sr * Cin_src + sk) * taps + tap] : 0.f;
    }
}
__global__ void pack_bias_kernel(const float* __restrict__ b, const float* __restrict__ b2, int n, PadSpec pad,
                                 float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = src_row(i, pad);
        out[i] = (s >= 0 ? b[s] : 0.f) + (b2 ? b2[i] : 0.f);
    }
}
__global__ void absmax_kernel(const float* __restrict__ w, long long n, float* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // non-negative floats
}

inline int grid_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// ---------------------------------------------------------------- model description
struct Param {
    std::string name;
    long long numel;
    const float* p = nullptr;
};

struct Conv {   // one nn.Conv3d / nn.Conv1d (k = 1)
    int w = -1, b = -1;                 // parameter indices
    int cout = 0, cin = 0, taps = 1;    // as stored
    PadSpec pad{0, 0, 0, 0};
    int cout_p = 0, cin_p = 0, cin_pad = 0;   // after head padding / after padding K to 64
    bool need_tc = false, need_simt = false;
    size_t off_hi = 0, off_lo = 0, off_simt = 0, off_bias = 0;   // byte offsets into the packed buffer
    float scale = 1.f;
    int amax_slot = -1;
};
struct Norm {
    int g = -1, b = -1, C = 0;
};
struct Res {
    int cin, cout;
    Norm n1, n2;
    Conv c1, c2, skip;
    bool has_skip = false, fused = false;
    int emb_w = -1, emb_b = -1;
    int film_off = 0;   // offset of this block's 2*cout FiLM values
    // fused tail [Cout][27 cout + cin] pairs
    size_t f_hi = 0, f_lo = 0, f_bias = 0;
    float f_scale = 1.f;
    int f_amax = -1;
};
struct Attn {
    Norm n;
    Conv qkv, proj;          // unpadded
    Conv qkv_p, proj_p;      // zero-padded heads (only when used)
    int C, heads;
    bool padded = false;
};
enum Kind { K_RES, K_ATTN, K_DOWN, K_UP, K_CONV };
struct Layer {
    Kind kind;
    int idx;
};

struct Dims {
    int d, h, w;
    long long V() const { return (long long)d * h * w; }
};

struct Act {   // a channels-last activation made of one or two sources (the un-materialised skip concat)
    float* x1 = nullptr;
    int c1 = 0;
    float* x2 = nullptr;
    int c2 = 0;
    Dims dims{0, 0, 0};
    double* st1 = nullptr;
    double* st2 = nullptr;
    int C() const { return c1 + c2; }
};

struct Unet {
    holo_unet_config cfg;
    std::vector<Param> params;
    std::map<std::string, int> by_name;
    std::vector<Res> res;
    std::vector<Attn> attn;
    std::vector<Conv> convs;   // Downsample.op / Upsample.conv / input conv / out conv
    std::vector<std::vector<Layer>> input_blocks, output_blocks;
    std::vector<Layer> middle;
    Norm out_norm;
    int out_conv = -1;
    int te_w0 = -1, te_b0 = -1, te_w2 = -1, te_b2 = -1;
    int emb_dim = 0, film_total = 0;
    // packed buffer layout
    size_t off_freqs = 0, off_film_w = 0, off_film_b = 0, off_amax = 0, packed_bytes = 0;
    int n_amax = 0;
    const uint8_t* packed = nullptr;
    bool packed_ok = false;
    // workspace plan
    size_t ws_bytes = 0;

    int add_param(const std::string& name, long long numel) {
        params.push_back({name, numel, nullptr});
        by_name[name] = (int)params.size() - 1;
        return (int)params.size() - 1;
    }
};

inline bool tile_ok(const Dims& d) { return d.w % 4 == 0 && d.h % 4 == 0 && d.d % 4 == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Conv make_conv(Unet& u, const std::string& prefix, int cout, int cin, int taps) {
    Conv c;
    c.cout = cout, c.cin = cin, c.taps = taps;
    c.w = u.add_param(prefix + ".weight", (long long)cout * cin * taps);
    c.b = u.add_param(prefix + ".bias", cout);
    c.cout_p = cout, c.cin_p = cin, c.cin_pad = (cin + 63) / 64 * 64;
    return c;
}
Norm make_norm(Unet& u, const std::string& prefix, int C) {
    Norm n;
    n.C = C;
    n.g = u.add_param(prefix + ".weight", C);
    n.b = u.add_param(prefix + ".bias", C);
    return n;
}
int make_res(Unet& u, const std::string& prefix, int cin, int cout) {
    Res r;
    r.cin = cin, r.cout = cout;
    r.n1 = make_norm(u, prefix + ".in_layers.0", cin);
    r.c1 = make_conv(u, prefix + ".in_layers.2", cout, cin, 27);
    r.emb_w = u.add_param(prefix + ".emb_layers.1.weight", (long long)2 * cout * u.emb_dim);
    r.emb_b = u.add_param(prefix + ".emb_layers.1.bias", 2 * cout);
    r.n2 = make_norm(u, prefix + ".out_layers.0", cout);
    r.c2 = make_conv(u, prefix + ".out_layers.3", cout, cout, 27);
    r.has_skip = cin != cout;
    if (r.has_skip) r.skip = make_conv(u, prefix + ".skip_connection", cout, cin, 1);
    r.film_off = u.film_total;
    u.film_total += 2 * cout;
    u.res.push_back(r);
    return (int)u.res.size() - 1;
}
int make_attn(Unet& u, const std::string& prefix, int C) {
    Attn a;
    a.C = C, a.heads = u.cfg.num_heads;
    a.n = make_norm(u, prefix + ".norm", C);
    a.qkv = make_conv(u, prefix + ".qkv", 3 * C, C, 1);
    a.proj = make_conv(u, prefix + ".proj_out", C, C, 1);
    u.attn.push_back(a);
    return (int)u.attn.size() - 1;
}
int make_plain_conv(Unet& u, const std::string& prefix, int cout, int cin, int taps) {
    u.convs.push_back(make_conv(u, prefix, cout, cin, taps));
    return (int)u.convs.size() - 1;
}

bool in_list(const int* v, int n, int x) {
    for (int i = 0; i < n; ++i)
        if (v[i] == x) return true;
    return false;
}

// UNetModel.__init__ (unet.py:638-797) for the SimpleUnet3D settings: the same module tree, hence the same names
void build(Unet& u) {
    const holo_unet_config& c = u.cfg;
    const int mc = c.model_channels;
    u.emb_dim = 4 * mc;
    u.te_w0 = u.add_param("time_embed.0.weight", (long long)u.emb_dim * mc);
    u.te_b0 = u.add_param("time_embed.0.bias", u.emb_dim);
    u.te_w2 = u.add_param("time_embed.2.weight", (long long)u.emb_dim * u.emb_dim);
    u.te_b2 = u.add_param("time_embed.2.bias", u.emb_dim);
    int width = c.channel_mult[0] * mc;
    std::vector<int> skip_widths;
    u.input_blocks.push_back({{K_CONV, make_plain_conv(u, "input_blocks.0.0", width, c.in_channels, 27)}});
    skip_widths.push_back(width);
    int ds = 1, bi = 1;
    for (int level = 0; level < c.n_levels; ++level) {
        const int target = c.channel_mult[level] * mc;
        for (int i = 0; i < c.num_res_blocks; ++i, ++bi) {
            std::vector<Layer> ls;
            const std::string p = "input_blocks." + std::to_string(bi);
            ls.push_back({K_RES, make_res(u, p + ".0", width, target)});
            width = target;
            if (in_list(c.attention_resolutions, c.n_attention_resolutions, ds)) ls.push_back({K_ATTN, make_attn(u, p + ".1", width)});
            u.input_blocks.push_back(ls);
            skip_widths.push_back(width);
        }
        if (level + 1 < c.n_levels) {
            const std::string p = "input_blocks." + std::to_string(bi);
            u.input_blocks.push_back({{K_DOWN, make_plain_conv(u, p + ".0.op", width, width, 27)}});
            ++bi;
            skip_widths.push_back(width);
            ds *= 2;
        }
    }
    u.middle.push_back({K_RES, make_res(u, "middle_block.0", width, width)});
    u.middle.push_back({K_ATTN, make_attn(u, "middle_block.1", width)});
    u.middle.push_back({K_RES, make_res(u, "middle_block.2", width, width)});
    int oi = 0;
    for (int level = c.n_levels - 1; level >= 0; --level) {
        const int target = c.channel_mult[level] * mc;
        for (int i = 0; i < c.num_res_blocks + 1; ++i, ++oi) {
            std::vector<Layer> ls;
            const std::string p = "output_blocks." + std::to_string(oi);
            const int sw = skip_widths.back();
            skip_widths.pop_back();
            ls.push_back({K_RES, make_res(u, p + ".0", width + sw, target)});
            width = target;
            int j = 1;
            if (in_list(c.attention_resolutions, c.n_attention_resolutions, ds)) {
                ls.push_back({K_ATTN, make_attn(u, p + "." + std::to_string(j), width)});
                ++j;
            }
            if (level > 0 && i == c.num_res_blocks) {
                ls.push_back({K_UP, make_plain_conv(u, p + "." + std::to_string(j) + ".conv", width, width, 27)});
                ds /= 2;
            }
            u.output_blocks.push_back(ls);
        }
    }
    u.out_norm = make_norm(u, "out.0", width);
    u.out_conv = make_plain_conv(u, "out.2", c.out_channels, c.channel_mult[0] * mc, 27);
}

// a CUDA runtime call inside Exec: the first failure is kept in `rc`
#define HOLO_CUDA_VOID(call)                                              \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess && rc == HOLO_OK) {                        \
            holo_set_error("holo_unet_fwd: %s", cudaGetErrorString(e__)); \
            rc = HOLO_ERR_CUDA;                                           \
        }                                                                 \
    } while (0)

// ---------------------------------------------------------------- executor (DRY = shape walk + workspace plan only)
struct Arena {   // first-fit free list over the caller's workspace; identical decisions in the dry and the real run
    uint8_t* base = nullptr;
    size_t peak = 0;
    std::vector<std::pair<size_t, size_t>> live;   // (offset, bytes), sorted by offset
    void* alloc(size_t bytes) {
        bytes = align_up(bytes ? bytes : 1, 1024);
        size_t off = 0;
        size_t pos = 0;
        for (; pos < live.size(); ++pos) {
            if (live[pos].first - off >= bytes) break;
            off = live[pos].first + live[pos].second;
        }
        live.insert(live.begin() + pos, {off, bytes});
        if (off + bytes > peak) peak = off + bytes;
        return base + off;
    }
    void release(const void* p) {
        if (!p) return;
        const size_t off = (size_t)((const uint8_t*)p - base);
        for (size_t i = 0; i < live.size(); ++i)
            if (live[i].first == off) {
                live.erase(live.begin() + i);
                return;
            }
    }
};

struct Exec {
    Unet& u;
    bool dry;
    Arena ar;
    cudaStream_t st;
    int rc = HOLO_OK;
    double* acc[2] = {nullptr, nullptr};
    int acc_i = 0;
    double* arena64 = nullptr;
    size_t arena_off = 0, arena_n = 1 << 18;
    float* film_all = nullptr;
    float* splitk_ws = nullptr;   // scratch of the deterministic split-K convolutions
    int fmt;       // HOLO_FMT_F16 or 0
    int pf16;

    Exec(Unet& u_, bool dry_, void* ws, cudaStream_t s) : u(u_), dry(dry_), st(s) {
        // the dry run plans against a fake non-null base so that offset 0 is an ordinary (releasable) allocation
        ar.base = dry ? reinterpret_cast<uint8_t*>((uintptr_t)1 << 30) : (uint8_t*)ws;
        pf16 = u.cfg.pair_f16 ? 1 : 0;
        fmt = pf16 ? HOLO_FMT_F16 : 0;
    }
    template <typename T>
    T* alloc(size_t n) {
        return (T*)ar.alloc(n * sizeof(T));
    }
    void free_(const void* p) { ar.release(p); }
    void check(int r) {
        if (r < 0 && rc == HOLO_OK) rc = r;
    }
    const float* P(int idx) const { return u.params[idx].p; }
    const uint8_t* pk(size_t off) const { return u.packed + off; }

    double* stats_slice(int cout) {
        const size_t n = 2 * (size_t)cout;
        if (arena_off + n > arena_n) return nullptr;
        double* s = arena64 + arena_off;
        arena_off += n;
        return s;
    }
    bool tc_ok(const Conv& c, const Dims& out) const { return c.cout_p % 16 == 0 && c.cout_p <= 4096 && tile_ok(out); }

    // ---- GroupNorm (+FiLM) (+SiLU) of a one- or two-source activation -> fp32 y or the operand pair (+ raw pair)
    void gn(const Act& a, const Norm& n, const float* film, bool silu, float* y, uint16_t* y_hi, uint16_t* y_lo,
            uint16_t* r_hi, uint16_t* r_lo) {
        const long long V = a.dims.V();
        if (a.st1 && (!a.x2 || a.st2)) {
            if (!dry)
                check(holo_gn_apply_fused_ch(a.x1, a.c1, a.st1, a.x2, a.c2, a.st2, V, P(n.g), P(n.b), film, GN_EPS, silu, y, y_hi,
                                             y_lo, r_hi, r_lo, pf16, st));
            return;
        }
        double* cur = acc[acc_i];
        double* nxt = acc[acc_i ^ 1];
        acc_i ^= 1;
        if (!dry) {
            check(holo_gn_stats_pp(a.x1, a.c1, a.x2, a.c2, V, cur, nxt, st));
            check(holo_gn_apply_fused(a.x1, a.c1, a.x2, a.c2, V, cur, P(n.g), P(n.b), film, GN_EPS, silu, y, y_hi, y_lo, r_hi,
                                      r_lo, pf16, st));
        }
    }

    Act conv_tc(const Conv& c, const uint16_t* hi, const uint16_t* lo, const Dims& in, const float* residual, int stride,
                bool want_stats, uint16_t** o_hi = nullptr, uint16_t** o_lo = nullptr) {
        const int k = c.taps == 27 ? 3 : 1;
        Act r;
        r.dims = {in.d / stride, in.h / stride, in.w / stride};
        r.c1 = c.cout_p;
        const long long Vo = r.dims.V();
        r.x1 = alloc<float>((size_t)Vo * c.cout_p);
        if (o_hi) *o_hi = alloc<uint16_t>((size_t)Vo * c.cout_p), *o_lo = alloc<uint16_t>((size_t)Vo * c.cout_p);
        double* sts = want_stats ? stats_slice(c.cout_p) : nullptr;
        int ret = 1;
        if (!dry) {
            ret = holo_conv3d_tc(hi, lo, c.cin_pad, in.d, in.h, in.w, k, stride, pk(c.off_hi), pk(c.off_lo),
                                 (const float*)pk(c.off_bias), residual, c.cout_p, r.x1, o_hi ? *o_hi : nullptr,
                                 o_lo ? *o_lo : nullptr, sts, fmt, 1.0f / c.scale, splitk_ws, st);
            check(ret);
        }
        r.st1 = (ret == 0) ? sts : nullptr;
        return r;
    }

    Act conv_simt(const Conv& c, const Act& a, int stride, bool ups, const float* residual, const float* pre) {
        const int k = c.taps == 27 ? 3 : 1;
        Act r;
        const Dims& d = a.dims;
        if (ups) r.dims = {2 * d.d, 2 * d.h, 2 * d.w};
        else if (stride == 2) r.dims = {(d.d - 1) / 2 + 1, (d.h - 1) / 2 + 1, (d.w - 1) / 2 + 1};
        else r.dims = d;
        r.c1 = c.cout_p;
        r.x1 = alloc<float>((size_t)r.dims.V() * c.cout_p);
        if (!dry) {
            if (pre)
                check(holo_conv3d_simt(pre, c.cin_p, nullptr, 0, d.d, d.h, d.w, k, stride, ups, (const float*)pk(c.off_simt),
                                       (const float*)pk(c.off_bias), residual, c.cout_p, r.x1, st));
            else
                check(holo_conv3d_simt(a.x1, a.c1, a.x2, a.c2, d.d, d.h, d.w, k, stride, ups, (const float*)pk(c.off_simt),
                                       (const float*)pk(c.off_bias), residual, c.cout_p, r.x1, st));
        }
        return r;
    }

    // conv(SiLU(GN(act)))  -- the in_layers / out_layers / out pattern; raw_*: also the pair of the raw input
    Act conv_norm(Conv& c, const Act& a, const Norm& n, const float* film, const float* residual, uint16_t** raw_hi,
                  uint16_t** raw_lo) {
        const bool tc = tc_ok(c, a.dims) && c.cin_p % 64 == 0 && u.cfg.use_tensor_cores;
        const long long V = a.dims.V();
        const bool raw_ok = raw_hi && tc && a.c1 % 8 == 0 && a.c2 % 8 == 0;
        if (raw_hi) *raw_hi = nullptr, *raw_lo = nullptr;
        if (tc) {
            c.need_tc = true;
            uint16_t* y_hi = alloc<uint16_t>((size_t)V * a.C());
            uint16_t* y_lo = alloc<uint16_t>((size_t)V * a.C());
            if (raw_ok) *raw_hi = alloc<uint16_t>((size_t)V * a.C()), *raw_lo = alloc<uint16_t>((size_t)V * a.C());
            gn(a, n, film, true, nullptr, y_hi, y_lo, raw_ok ? *raw_hi : nullptr, raw_ok ? *raw_lo : nullptr);
            Act r = conv_tc(c, y_hi, y_lo, a.dims, residual, 1, true);
            free_(y_hi), free_(y_lo);
            return r;
        }
        c.need_simt = true;
        float* y = alloc<float>((size_t)V * a.C());
        gn(a, n, film, true, y, nullptr, nullptr, nullptr, nullptr);
        Act r = conv_simt(c, a, 1, false, residual, y);
        free_(y);
        return r;
    }

    // conv on a raw activation (first conv, skip 1x1, Upsample / Downsample convs)
    Act conv_raw(Conv& c, const Act& a, int stride, bool ups, const float* residual) {
        const Dims ind = ups ? Dims{2 * a.dims.d, 2 * a.dims.h, 2 * a.dims.w} : a.dims;
        const bool even = ind.d % 2 == 0 && ind.h % 2 == 0 && ind.w % 2 == 0;
        const Dims od = {ind.d / stride, ind.h / stride, ind.w / stride};
        if ((stride == 1 || even) && tc_ok(c, od) && u.cfg.use_tensor_cores) {
            c.need_tc = true;
            const long long Vo = ind.V();
            uint16_t* hi = alloc<uint16_t>((size_t)Vo * c.cin_pad);
            uint16_t* lo = alloc<uint16_t>((size_t)Vo * c.cin_pad);
            if (!dry)
                check(holo_split_bf16(a.x1, a.c1, a.x2, a.c2, a.dims.V(), c.cin_pad, ups ? 1 : 0, a.dims.d, a.dims.h, a.dims.w,
                                      hi, lo, pf16, st));
            Act r = conv_tc(c, hi, lo, ind, residual, stride, true);
            free_(hi), free_(lo);
            return r;
        }
        c.need_simt = true;
        return conv_simt(c, a, stride, ups, residual, nullptr);
    }

    Act res_block(Res& b, const Act& a) {
        const float* film = film_all + b.film_off;
        uint16_t *raw_hi = nullptr, *raw_lo = nullptr;
        Act h = conv_norm(b.c1, a, b.n1, nullptr, nullptr, b.has_skip ? &raw_hi : nullptr, b.has_skip ? &raw_lo : nullptr);
        const float* skip = nullptr;
        float* skip_owned = nullptr;
        if (!b.has_skip) {
            skip = a.x1;
        } else {
            if (u.cfg.fuse_skip && u.cfg.use_tensor_cores && raw_hi && tc_ok(b.c2, a.dims) && b.c2.cin % 64 == 0 &&
                b.skip.cin % 64 == 0 && b.skip.cin == a.C()) {
                b.fused = true;
                const long long V = h.dims.V();
                uint16_t* y_hi = alloc<uint16_t>((size_t)V * h.c1);
                uint16_t* y_lo = alloc<uint16_t>((size_t)V * h.c1);
                gn(h, b.n2, film, true, nullptr, y_hi, y_lo, nullptr, nullptr);
                Act r;
                r.dims = h.dims, r.c1 = b.cout;
                r.x1 = alloc<float>((size_t)V * b.cout);
                double* sts = stats_slice(b.cout);
                int ret = 1;
                if (!dry) {
                    ret = holo_conv3d_tc_skip(y_hi, y_lo, b.c2.cin, raw_hi, raw_lo, b.skip.cin, h.dims.d, h.dims.h, h.dims.w,
                                              pk(b.f_hi), pk(b.f_lo), (const float*)pk(b.f_bias), nullptr, b.cout, r.x1, sts, fmt,
                                              1.0f / b.f_scale, splitk_ws, st);
                    check(ret);
                }
                r.st1 = ret == 0 ? sts : nullptr;
                free_(y_hi), free_(y_lo), free_(raw_hi), free_(raw_lo), free_(h.x1);
                return r;
            }
            Act s;
            if (raw_hi && tc_ok(b.skip, a.dims) && b.skip.cin_pad == a.C()) {
                b.skip.need_tc = true;
                s = conv_tc(b.skip, raw_hi, raw_lo, a.dims, nullptr, 1, true);
            } else {
                s = conv_raw(b.skip, a, 1, false, nullptr);
            }
            skip = skip_owned = s.x1;
        }
        free_(raw_hi), free_(raw_lo);
        Act r = conv_norm(b.c2, h, b.n2, film, skip, nullptr, nullptr);
        free_(h.x1), free_(skip_owned);
        return r;
    }

    static int kv_split_auto(int T, int heads) {
        const int ctas = ((T + 127) / 128) * heads;
        int s = (T / 64) / 2;
        const int by_sm = 148 / ctas;
        if (by_sm < s) s = by_sm;
        return s < 1 ? 1 : s;
    }

    Act attention(Attn& b, const Act& a) {
        const int T = (int)a.dims.V(), C = a.C(), heads = b.heads, ch = C / heads;
        const bool use_tc = u.cfg.use_tensor_cores != 0;
        int chp = ch < 64 ? 64 : ch;
        bool tc = use_tc && T % 128 == 0 && chp % 64 == 0 && C % 64 == 0;
        Conv *cq = &b.qkv, *cp = &b.proj;
        if (tc && chp != ch) {
            if (!b.padded) {
                b.padded = true;
                b.qkv_p = b.qkv, b.proj_p = b.proj;
                b.qkv_p.pad = {1, heads, ch, chp}, b.qkv_p.cout_p = heads * 3 * chp;
                b.proj_p.pad = {2, heads, ch, chp}, b.proj_p.cin_p = heads * chp, b.proj_p.cin_pad = (heads * chp + 63) / 64 * 64;
                b.qkv_p.need_tc = b.qkv_p.need_simt = b.proj_p.need_tc = b.proj_p.need_simt = false;
            }
            cq = &b.qkv_p, cp = &b.proj_p;
        } else {
            chp = ch;
        }
        const int Cp = heads * chp;
        Act flat = a;
        flat.dims = {1, 1, T};
        Act out;
        if (!tc) {
            cq->need_simt = cp->need_simt = true;
            float* y = alloc<float>((size_t)T * C);
            gn(flat, b.n, nullptr, false, y, nullptr, nullptr, nullptr, nullptr);
            Act qkv = conv_simt(*cq, flat, 1, false, nullptr, y);
            free_(y);
            float* att = alloc<float>((size_t)T * C);
            if (!dry) check(holo_attention_simt(qkv.x1, T, heads, ch, att, st));
            free_(qkv.x1);
            Act aa;
            aa.x1 = att, aa.c1 = C, aa.dims = flat.dims;
            out = conv_simt(*cp, aa, 1, false, a.x1, nullptr);
            free_(att);
            out.dims = a.dims;
            return out;
        }
        cq->need_tc = cp->need_tc = true;
        const Dims g = {T / 32, 4, 8};   // GEMM view of the token axis for the TMA box
        uint16_t* y_hi = alloc<uint16_t>((size_t)T * C);
        uint16_t* y_lo = alloc<uint16_t>((size_t)T * C);
        gn(flat, b.n, nullptr, false, nullptr, y_hi, y_lo, nullptr, nullptr);
        uint16_t *q_hi, *q_lo;
        Act qkv = conv_tc(*cq, y_hi, y_lo, g, nullptr, 1, false, &q_hi, &q_lo);
        free_(y_hi), free_(y_lo);
        uint16_t* a_hi = alloc<uint16_t>((size_t)T * Cp);
        uint16_t* a_lo = alloc<uint16_t>((size_t)T * Cp);
        if (chp == 64 || chp == 128) {
            uint16_t* vt_hi = alloc<uint16_t>((size_t)Cp * T);
            uint16_t* vt_lo = alloc<uint16_t>((size_t)Cp * T);
            const int splits = u.cfg.attn_kv_split > 0 ? u.cfg.attn_kv_split : kv_split_auto(T, heads);
            const long long wsb = holo_attention_flash_workspace_bytes(T, heads, chp, splits);
            void* ws = wsb ? (void*)alloc<uint8_t>((size_t)wsb) : nullptr;
            if (!dry) {
                check(holo_v_transpose_split(qkv.x1, T, heads, chp, vt_hi, vt_lo, pf16, st));
                check(holo_attention_flash(q_hi, q_lo, vt_hi, vt_lo, T, heads, chp, nullptr, a_hi, a_lo, pf16,
                                           1.0f / sqrtf((float)ch), 0, 0, splits, ws, st));
            }
            free_(vt_hi), free_(vt_lo), free_(ws);
        } else {
            // three-launch pipeline per head: S = Q K^T, P = softmax(S / sqrt(ch)), O = P V
            float* S = alloc<float>((size_t)T * T);
            uint16_t* P_hi = alloc<uint16_t>((size_t)T * T);
            uint16_t* P_lo = alloc<uint16_t>((size_t)T * T);
            uint16_t* vt_hi = alloc<uint16_t>((size_t)ch * T);
            uint16_t* vt_lo = alloc<uint16_t>((size_t)ch * T);
            float* att = alloc<float>((size_t)T * C);
            const float p_scale = pf16 ? 4096.f : 1.f;
            if (!dry) {
                HOLO_CUDA_VOID(cudaMemsetAsync(att, 0, (size_t)T * C * sizeof(float), st));
                for (int h = 0; h < heads; ++h) {
                    const size_t base = (size_t)h * 3 * ch;
                    check(holo_gemm_tc(q_hi + base, q_lo + base, 3LL * C, T, ch, q_hi + base + ch, q_lo + base + ch, 3LL * C, T,
                                       nullptr, nullptr, T, S, nullptr, nullptr, 0, fmt, 1.0f, st));
                    check(holo_softmax_split(S, T, T, 1.0f / sqrtf((float)ch), P_hi, P_lo, pf16, p_scale, st));
                    check(holo_transpose_split_bf16(qkv.x1 + base + 2 * ch, 3LL * C, T, ch, vt_hi, vt_lo, pf16, st));
                    check(holo_gemm_tc(P_hi, P_lo, T, T, T, vt_hi, vt_lo, T, ch, nullptr, nullptr, C, att + (size_t)h * ch, nullptr,
                                       nullptr, 1, fmt, 1.0f / p_scale, st));
                }
                check(holo_split_bf16(att, C, nullptr, 0, T, C, 0, 0, 0, 0, a_hi, a_lo, pf16, st));
            }
            free_(S), free_(P_hi), free_(P_lo), free_(vt_hi), free_(vt_lo), free_(att);
        }
        free_(qkv.x1), free_(q_hi), free_(q_lo);
        out = conv_tc(*cp, a_hi, a_lo, g, a.x1, 1, true);
        free_(a_hi), free_(a_lo);
        out.dims = a.dims;
        return out;
    }

    // runs one nn.Sequential; frees every intermediate activation it creates (never its input)
    Act run(std::vector<Layer>& seq, const Act& in) {
        Act cur = in;
        bool owned = false;
        for (Layer& l : seq) {
            Act nxt;
            switch (l.kind) {
                case K_RES: nxt = res_block(u.res[l.idx], cur); break;
                case K_ATTN: nxt = attention(u.attn[l.idx], cur); break;
                case K_DOWN: nxt = conv_raw(u.convs[l.idx], cur, 2, false, nullptr); break;
                case K_UP: nxt = conv_raw(u.convs[l.idx], cur, 1, true, nullptr); break;
                default: nxt = conv_raw(u.convs[l.idx], cur, 1, false, nullptr); break;
            }
            if (owned) free_(cur.x1);
            cur = nxt;
            owned = true;
        }
        return cur;
    }

    // x_cl (V, Cin) channels-last -> out_cl (V, Cout)
    void forward_cl(const float* x_cl, const long long* t_dev, float* out_cl) {
        const holo_unet_config& c = u.cfg;
        acc[0] = alloc<double>(512), acc[1] = alloc<double>(512);
        arena64 = alloc<double>(arena_n);
        arena_off = 0, acc_i = 0;
        const int mc = c.model_channels, E = u.emb_dim;
        float* e0 = alloc<float>(mc);
        float* e1 = alloc<float>(E);
        float* emb = alloc<float>(E);
        film_all = alloc<float>(u.film_total);
        splitk_ws = u.cfg.splitk_workspace ? alloc<float>((size_t)holo_conv3d_tc_splitk_bytes() / 4) : nullptr;
        if (!dry) {
            HOLO_CUDA_VOID(cudaMemsetAsync(acc[0], 0, 512 * sizeof(double), st));
            HOLO_CUDA_VOID(cudaMemsetAsync(arena64, 0, arena_n * sizeof(double), st));
            check(holo_timestep_embedding(t_dev, 1, mc, (const float*)pk(u.off_freqs), e0, st));
            check(holo_linear_rows(e0, P(u.te_w0), P(u.te_b0), 1, mc, E, 0, 1, e1, st));
            check(holo_linear_rows(e1, P(u.te_w2), P(u.te_b2), 1, E, E, 0, 0, emb, st));
            check(holo_linear_rows(emb, (const float*)pk(u.off_film_w), (const float*)pk(u.off_film_b), 1, E, u.film_total, 1, 0,
                                   film_all, st));
        }
        Act act;
        act.x1 = const_cast<float*>(x_cl), act.c1 = c.in_channels, act.dims = {c.D, c.H, c.W};
        std::vector<Act> skips;
        for (auto& blk : u.input_blocks) {
            Act nxt = run(blk, act);
            skips.push_back(nxt);   // kept alive until its output block has consumed it
            act = nxt;
        }
        Act mid = run(u.middle, act);
        act = mid;
        bool act_owned = true;   // `mid` is not a skip
        for (auto& blk : u.output_blocks) {
            Act s = skips.back();
            skips.pop_back();
            Act cat;
            cat.x1 = act.x1, cat.c1 = act.c1, cat.dims = act.dims, cat.x2 = s.x1, cat.c2 = s.c1, cat.st1 = act.st1, cat.st2 = s.st1;
            Act nxt = run(blk, cat);
            if (act_owned) free_(act.x1);
            free_(s.x1);
            act = nxt;
            act_owned = true;
        }
        // out: GroupNorm + SiLU + conv (unet.py:789-793), written straight into the caller's buffer
        Conv& oc = u.convs[u.out_conv];
        Act fin = conv_norm(oc, act, u.out_norm, nullptr, nullptr, nullptr, nullptr);
        if (!dry)
            HOLO_CUDA_VOID(cudaMemcpyAsync(out_cl, fin.x1, (size_t)fin.dims.V() * oc.cout_p * sizeof(float),
                                           cudaMemcpyDeviceToDevice, st));
        free_(fin.x1);
        if (act_owned) free_(act.x1);
        free_(e0), free_(e1), free_(emb), free_(film_all), free_(splitk_ws), free_(acc[0]), free_(acc[1]), free_(arena64);
    }

};

// ---------------------------------------------------------------- packed-buffer layout
void plan_packed(Unet& u) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    u.n_amax = 0;
    auto plan_conv = [&](Conv& c) {
        c.off_bias = take((size_t)c.cout_p * 4);
        if (c.need_tc) {
            c.off_hi = take((size_t)c.cout_p * c.taps * c.cin_pad * 2);
            c.off_lo = take((size_t)c.cout_p * c.taps * c.cin_pad * 2);
            c.amax_slot = u.n_amax++;
        }
        if (c.need_simt) c.off_simt = take((size_t)c.taps * c.cin_p * c.cout_p * 4);
    };
    for (Conv& c : u.convs) plan_conv(c);
    for (Res& r : u.res) {
        plan_conv(r.c1);
        if (r.fused) {
            r.f_bias = take((size_t)r.cout * 4);
            r.f_hi = take((size_t)r.cout * (27 * r.c2.cin + r.skip.cin) * 2);
            r.f_lo = take((size_t)r.cout * (27 * r.c2.cin + r.skip.cin) * 2);
            r.f_amax = u.n_amax;
            u.n_amax += 2;   // the 3^3 part and the 1x1 part (the common scale follows the larger)
        } else {
            plan_conv(r.c2);
            if (r.has_skip) plan_conv(r.skip);
        }
    }
    for (Attn& a : u.attn) {
        if (a.padded) plan_conv(a.qkv_p), plan_conv(a.proj_p);
        else plan_conv(a.qkv), plan_conv(a.proj);
    }
    u.off_freqs = take((size_t)(u.cfg.model_channels / 2) * 4);
    u.off_film_w = take((size_t)u.film_total * u.emb_dim * 4);
    u.off_film_b = take((size_t)u.film_total * 4);
    u.off_amax = take((size_t)u.n_amax * 4);
    u.packed_bytes = off;
}

float pow2_scale(float amax, bool f16) {
    // fp16 pair of s w, s = 2^e with s max|w| in [2^9, 2^10): hi + lo exact to 2^-22 |w| and lo stays normal
    if (!f16 || !(amax > 0.f) || !std::isfinite(amax)) return 1.f;
    return std::ldexp(1.f, 9 - (int)std::floor(std::log2(amax)));
}

}  // namespace

extern "C" int holo_unet_create(const holo_unet_config* cfg, void** handle) {
    HOLO_CHECK_ARG(cfg && handle, "holo_unet_create: null argument");
    const holo_unet_config& c = *cfg;
    HOLO_CHECK_ARG(c.n_levels >= 1 && c.n_levels <= 8 && c.n_attention_resolutions >= 0 && c.n_attention_resolutions <= 8 &&
                       c.num_res_blocks >= 1 && c.num_heads >= 1 && c.in_channels >= 1 && c.out_channels >= 1 &&
                       c.model_channels >= 32 && c.model_channels % 32 == 0 && c.D >= 1 && c.H >= 1 && c.W >= 1,
                   "holo_unet_create: bad configuration");
    const int down = 1 << (c.n_levels - 1);
    HOLO_CHECK_ARG(c.D % down == 0 && c.H % down == 0 && c.W % down == 0,
                   "holo_unet_create: the grid (%d, %d, %d) must be divisible by 2^(levels - 1) = %d", c.D, c.H, c.W, down);
    Unet* u = new Unet();
    u->cfg = c;
    build(*u);
    // dry run: decides, from the static shapes, which kernel every convolution takes (=> which packed layouts exist)
    // and the peak of the workspace
    Exec ex(*u, true, nullptr, nullptr);
    ex.forward_cl(nullptr, nullptr, nullptr);
    u->ws_bytes = ex.ar.peak + (size_t)2 * c.D * c.H * c.W * (c.in_channels > c.out_channels ? c.in_channels : c.out_channels) * 4 + 4096;
    plan_packed(*u);
    *handle = u;
    return HOLO_OK;
}

extern "C" int holo_unet_destroy(void* handle) {
    delete reinterpret_cast<Unet*>(handle);
    return HOLO_OK;
}

extern "C" int holo_unet_param_count(void* handle) {
    return handle ? (int)reinterpret_cast<Unet*>(handle)->params.size() : HOLO_ERR_ARG;
}
extern "C" const char* holo_unet_param_name(void* handle, int i) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    return (u && i >= 0 && i < (int)u->params.size()) ? u->params[i].name.c_str() : nullptr;
}
extern "C" long long holo_unet_param_numel(void* handle, int i) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    return (u && i >= 0 && i < (int)u->params.size()) ? u->params[i].numel : -1;
}
extern "C" int holo_unet_set_param(void* handle, const char* name, const float* dev_ptr, long long numel) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    HOLO_CHECK_ARG(u && name && dev_ptr, "holo_unet_set_param: null argument");
    auto it = u->by_name.find(name);
    HOLO_CHECK_ARG(it != u->by_name.end(), "holo_unet_set_param: unknown parameter '%s'", name);
    Param& p = u->params[it->second];
    HOLO_CHECK_ARG(p.numel == numel, "holo_unet_set_param: '%s' has %lld elements, expected %lld", name, numel, p.numel);
    p.p = dev_ptr;
    u->packed_ok = false;
    return HOLO_OK;
}
extern "C" long long holo_unet_packed_bytes(void* handle) {
    return handle ? (long long)reinterpret_cast<Unet*>(handle)->packed_bytes : -1;
}
extern "C" long long holo_unet_workspace_bytes(void* handle) {
    return handle ? (long long)reinterpret_cast<Unet*>(handle)->ws_bytes : -1;
}

// Builds every derived weight layout in `packed`.  Synchronises `stream` once (the per-layer power-of-two scales of the
// fp16 pairs are derived from max|w| on the host, as the Python executor does).
extern "C" int holo_unet_pack(void* handle, void* packed_dev, void* stream) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    HOLO_CHECK_ARG(u && packed_dev, "holo_unet_pack: null argument");
    for (const Param& p : u->params) HOLO_CHECK_ARG(p.p, "holo_unet_pack: parameter '%s' was not set", p.name.c_str());
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* pk = reinterpret_cast<uint8_t*>(packed_dev);
    const bool f16 = u->cfg.pair_f16 != 0;
    float* amax_dev = reinterpret_cast<float*>(pk + u->off_amax);
    HOLO_CUDA(cudaMemsetAsync(amax_dev, 0, (size_t)(u->n_amax ? u->n_amax : 1) * 4, st), "holo_unet_pack");
    auto amax_of = [&](const Conv& c, int slot) {
        absmax_kernel<<<grid_for(u->params[c.w].numel), 256, 0, st>>>(u->params[c.w].p, u->params[c.w].numel, amax_dev + slot);
    };
    std::vector<Conv*> all;
    for (Conv& c : u->convs) all.push_back(&c);
    for (Res& r : u->res) {
        all.push_back(&r.c1);
        if (r.fused) {
            amax_of(r.c2, r.f_amax), amax_of(r.skip, r.f_amax + 1);
        } else {
            all.push_back(&r.c2);
            if (r.has_skip) all.push_back(&r.skip);
        }
    }
    for (Attn& a : u->attn) {
        if (a.padded) all.push_back(&a.qkv_p), all.push_back(&a.proj_p);
        else all.push_back(&a.qkv), all.push_back(&a.proj);
    }
    for (Conv* c : all)
        if (c->need_tc) amax_of(*c, c->amax_slot);
    std::vector<float> amax((size_t)(u->n_amax ? u->n_amax : 1));
    HOLO_CUDA(cudaMemcpyAsync(amax.data(), amax_dev, (size_t)u->n_amax * 4, cudaMemcpyDeviceToHost, st), "holo_unet_pack");
    HOLO_CUDA(cudaStreamSynchronize(st), "holo_unet_pack");
    for (Conv* c : all) {
        const float* w = u->params[c->w].p;
        pack_bias_kernel<<<grid_for(c->cout_p), 256, 0, st>>>(u->params[c->b].p, nullptr, c->cout_p, c->pad,
                                                              reinterpret_cast<float*>(pk + c->off_bias));
        if (c->need_tc) {
            c->scale = pow2_scale(amax[c->amax_slot], f16);
            pack_pairs_kernel<<<grid_for((long long)c->cout_p * c->taps * c->cin_pad), 256, 0, st>>>(
                w, c->cout, c->cin, c->taps, c->cout_p, c->cin_p, c->cin_pad, (long long)c->taps * c->cin_pad, 0, c->scale, f16,
                c->pad, reinterpret_cast<uint16_t*>(pk + c->off_hi), reinterpret_cast<uint16_t*>(pk + c->off_lo));
        }
        if (c->need_simt)
            pack_simt_kernel<<<grid_for((long long)c->taps * c->cin_p * c->cout_p), 256, 0, st>>>(
                w, c->cin, c->taps, c->cout_p, c->cin_p, c->pad, reinterpret_cast<float*>(pk + c->off_simt));
    }
    for (Res& r : u->res) {
        if (!r.fused) continue;
        const float am = fmaxf(amax[r.f_amax], amax[r.f_amax + 1]);
        r.f_scale = pow2_scale(am, f16);
        const long long pitch = 27LL * r.c2.cin + r.skip.cin;
        uint16_t* hi = reinterpret_cast<uint16_t*>(pk + r.f_hi);
        uint16_t* lo = reinterpret_cast<uint16_t*>(pk + r.f_lo);
        const PadSpec none{0, 0, 0, 0};
        pack_pairs_kernel<<<grid_for((long long)r.cout * 27 * r.c2.cin), 256, 0, st>>>(
            u->params[r.c2.w].p, r.cout, r.c2.cin, 27, r.cout, r.c2.cin, r.c2.cin, pitch, 0, r.f_scale, f16, none, hi, lo);
        pack_pairs_kernel<<<grid_for((long long)r.cout * r.skip.cin), 256, 0, st>>>(
            u->params[r.skip.w].p, r.cout, r.skip.cin, 1, r.cout, r.skip.cin, r.skip.cin, pitch, 27LL * r.c2.cin, r.f_scale, f16,
            none, hi, lo);
        pack_bias_kernel<<<grid_for(r.cout), 256, 0, st>>>(u->params[r.c2.b].p, u->params[r.skip.b].p, r.cout, none,
                                                           reinterpret_cast<float*>(pk + r.f_bias));
    }
    // sinusoid frequencies exp(-ln(1e4) i / half), computed on the host in fp32 as the reference does (nn.py:119-121)
    const int half = u->cfg.model_channels / 2;
    std::vector<float> freqs((size_t)half);
    for (int i = 0; i < half; ++i) freqs[i] = expf(-logf(10000.f) * (float)i / (float)half);
    HOLO_CUDA(cudaMemcpyAsync(pk + u->off_freqs, freqs.data(), (size_t)half * 4, cudaMemcpyHostToDevice, st), "holo_unet_pack");
    // concatenated FiLM projection (emb_layers[1] of every ResBlock, unet.py:199-205): one launch per evaluation
    for (const Res& r : u->res) {
        HOLO_CUDA(cudaMemcpyAsync(pk + u->off_film_w + (size_t)r.film_off * u->emb_dim * 4, u->params[r.emb_w].p,
                                  (size_t)2 * r.cout * u->emb_dim * 4, cudaMemcpyDeviceToDevice, st),
                  "holo_unet_pack");
        HOLO_CUDA(cudaMemcpyAsync(pk + u->off_film_b + (size_t)r.film_off * 4, u->params[r.emb_b].p, (size_t)2 * r.cout * 4,
                                  cudaMemcpyDeviceToDevice, st),
                  "holo_unet_pack");
    }
    HOLO_CUDA(cudaStreamSynchronize(st), "holo_unet_pack");   // `freqs` (host) must outlive its copy
    HOLO_CHECK_LAUNCH("holo_unet_pack");
    u->packed = pk;
    u->packed_ok = true;
    return HOLO_OK;
}

extern "C" int holo_unet_fwd_cl(void* handle, const float* x_cl, const long long* t_dev, float* out_cl, void* workspace,
                                void* stream) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    HOLO_CHECK_ARG(u && x_cl && t_dev && out_cl && workspace, "holo_unet_fwd_cl: null argument");
    HOLO_CHECK_ARG(u->packed_ok, "holo_unet_fwd_cl: call holo_unet_pack first (and again after holo_unet_set_param)");
    Exec ex(*u, false, workspace, (cudaStream_t)stream);
    ex.forward_cl(x_cl, t_dev, out_cl);
    return ex.rc;
}

extern "C" int holo_unet_fwd(void* handle, const float* x_ncdhw, const long long* t_dev, float* out_ncdhw, void* workspace,
                             void* stream) {
    Unet* u = reinterpret_cast<Unet*>(handle);
    HOLO_CHECK_ARG(u && x_ncdhw && t_dev && out_ncdhw && workspace, "holo_unet_fwd: null argument");
    HOLO_CHECK_ARG(u->packed_ok, "holo_unet_fwd: call holo_unet_pack first (and again after holo_unet_set_param)");
    const holo_unet_config& c = u->cfg;
    const long long V = (long long)c.D * c.H * c.W;
    const int cmax = c.in_channels > c.out_channels ? c.in_channels : c.out_channels;
    // the two channels-last staging tensors live at the END of the workspace (holo_unet_workspace_bytes counts them)
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    const size_t tail = (size_t)2 * V * cmax * 4 + 4096;
    float* x_cl = reinterpret_cast<float*>(ws + align_up(u->ws_bytes - tail, 1024));
    float* y_cl = x_cl + (size_t)V * cmax;
    int rc = holo_transpose2d(x_ncdhw, x_cl, c.in_channels, (int)V, stream);
    if (rc) return rc;
    Exec ex(*u, false, workspace, (cudaStream_t)stream);
    ex.forward_cl(x_cl, t_dev, y_cl);
    if (ex.rc) return ex.rc;
    return holo_transpose2d(y_cl, out_ncdhw, (int)V, c.out_channels, stream);
}
