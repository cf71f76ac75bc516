/* holo_b200.h -- C-ABI of the B200-native HoloDiffusion hot path (libholo_b200.so).
 *
 * The reference (facebookresearch/holo_diffusion) has no FFI: its Python calls torch / pytorch3d ops directly.
 * Every entry point below therefore names the reference Python call site (relative to /root/reference) whose
 * arithmetic it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer, the
 *     library never allocates; sizes of packed buffers are reported by the *_floats / *_bytes helpers.
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on that stream, no hidden syncs,
 *     CUDA-graph capturable (holo_conv3d_tc additionally encodes TMA descriptors on the host per call).
 *   - return 0 on success, negative on error; holo_last_error() gives the thread-local message.
 *   - activations are fp32 channels-last: a (D,H,W) voxel grid with C channels is float[D*H*W][C]
 *     ("DHWC"); V = D*H*W.  N (batch) is looped by the caller (the reference renders one grid per process,
 *     holo_diffusion_model.py:326).
 */
#ifndef HOLO_B200_H
#define HOLO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HOLO_B200_VERSION 122

int holo_version(void);
const char* holo_last_error(void);
int holo_device_info(int* sm_major, int* sm_minor, int* n_sm);

/* ---------------------------------------------------------------------------------------------------------
 * Renderer
 * ------------------------------------------------------------------------------------------------------- */

/* Full-grid evaluation rays for n_cam NDC perspective cameras.
 * Replaces self.raysampler(target_cameras, evaluation_mode) -- holo_diffusion_model.py:442-448
 * (pytorch3d AdaptiveRaySampler / NDCMultinomialRaysampler, configs/base.yaml:129-140).
 * R (n_cam,3,3) row-vector convention X_cam = X_world R + T; xy (n_rays,2) NDC pixel centres, ray r = h*W + w.
 * origins/dirs (n_cam,n_rays,3), lengths (n_cam,n_rays,S). */
int holo_raygen(const float* R, const float* T, const float* focal, const float* pp, const float* xy, int n_cam,
                int n_rays, int S, float scene_extent, const float* scene_center3_host, float* origins, float* dirs,
                float* lengths, void* stream);

/* One composition step of the density-net collapse in fp64 (y = A x + c through a Linear, optional skip concat).
 * Replaces MLPWithInputSkips.forward for the activation-free layers -- custom_modules.py:108-112,133-160.
 * A_in == NULL: first layer (A_out = W, c_out = b). */
int holo_affine_compose_f64(const float* W, const float* b, int out_dim, int in_total, const double* A_in,
                            const double* c_in, int rows, int C, int skip, double* A_out, double* c_out,
                            void* stream);

/* Pack collapsed density net (A_eff (H+1,C), c_eff (H+1)) + radiance layer Wr (3, H+E), br (3) for the render
 * kernel.  RenderMLP parameters: holo_voxel_grid_implicit_function.py:62-105. */
long long holo_render_mlp_packed_floats(int H, int C, int E);
int holo_pack_render_mlp(const double* A_eff, const double* c_eff, const float* Wr, const float* br, int H, int C,
                         int E, float* packed, void* stream);

/* The fused renderer: per ray, n_passes x (trilinear sample -> RenderMLP -> emission-absorption compositing),
 * with importance re-sampling of the ray depths between passes.
 * Replaces HoloMultiPassEmissionAbsorptionRenderer._run_raymarcher (holo_multipass_ea.py:79-125) including
 *   HoloVoxelGridImplicitFunction.forward (holo_voxel_grid_implicit_function.py:182-269),
 *   EmissionAbsorptionRaymarcher (configs/base.yaml:149-159) and RayPointRefiner (configs/base.yaml:142-146),
 * and the GenericModel._render chunk loop (holo_diffusion_model.py:451-457): one launch covers all rays.
 * grid_dhwc (D,H,W,C); origins/dirs (n_rays,3); lengths (n_rays,S).
 * Outputs of the last pass: features (n_rays,3), depths (n_rays), masks (n_rays), optional weights
 * (n_rays,S_last), optional lengths_out (n_rays,S_last).  prev_* = outputs of pass 0 when n_passes == 2
 * (RendererOutput.prev_stage); any of them may be NULL.  S_last = S (+ n_fine if add_input_samples). */
int holo_render_fwd(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent,
                    const float* packed_mlp, int hidden, int n_harmonic, const float* origins, const float* dirs,
                    const float* lengths, int n_rays, int S, int n_passes, int n_fine, int add_input_samples,
                    const float* bg3_host, float background_opacity, float* features, float* depths, float* masks,
                    float* weights, float* lengths_out, float* prev_features, float* prev_depths, float* prev_masks,
                    float* prev_weights, void* stream);

/* Tensor-core version of the fused renderer (same contract and outputs as holo_render_fwd): the 256-wide hidden
 * layer runs on tcgen05 as a 3xBF16 split (M = 128 rays of one depth step), density / compositing / re-sampling stay
 * in fp32 registers.  C in {16, 32}, hidden = 256.  tc_image: holo_render_tc_image_bytes() bytes written by
 * holo_pack_render_mlp_tc from the collapsed net; scratch_weights: S * n_rays floats (two passes only). */
long long holo_render_tc_image_bytes(void);
int holo_pack_render_mlp_tc(const double* A_eff, const double* c_eff, const float* Wr, const float* br, int H, int C,
                            int E, void* image, void* stream);
int holo_render_fwd_tc(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent, const void* tc_image,
                       int n_harmonic, const float* origins, const float* dirs, const float* lengths, int n_rays,
                       int S, int n_passes, int n_fine, int add_input_samples, const float* bg3_host,
                       float background_opacity, float* features, float* depths, float* masks, float* weights,
                       float* lengths_out, float* prev_features, float* prev_depths, float* prev_masks,
                       float* prev_weights, float* scratch_weights, void* stream);

/* Per-stage renderer ops: the reference's plug-in seams as separately callable kernels.  The fused launches above
 * are the hot path; these carry `render_normals`, the view-independent feature head, explicit `pts_3d`,
 * training-mode density noise / stratified refinement and user-supplied stages in between.
 *
 * holo_if_fwd replaces HoloVoxelGridImplicitFunction.forward (holo_voxel_grid_implicit_function.py:182-269):
 * points = origins + lengths * dirs over (n_rays, S) -- or explicit pts_3d (n_points, 3), in which case dirs may be
 * NULL (dummy all-ones directions, :228-236) -- trilinear sample, RenderMLP; n_points = n_rays * S.
 * head_w (n_head, hidden) / head_b: the single-layer view-independent feature net (:94-105), n_head in 0..64.
 * densities (n_points), features (n_points, 3 + n_head) = [rgb | head]; normals (n_points, 3) or NULL =
 * RenderMLP.get_normals (:131-145), computed analytically (the density net is affine in the sampled feature). */
int holo_if_fwd(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent, const float* packed_mlp,
                int hidden, int n_harmonic, const float* head_w, const float* head_b, int n_head,
                const float* origins, const float* dirs, const float* lengths, const float* pts_3d,
                long long n_points, int S, float* densities, float* features, float* normals, void* stream);
/* RenderMLP.forward(features, view_dirs) (holo_voxel_grid_implicit_function.py:107-129): feats (n_points, C),
 * view_dirs (n_points, 3) used as given. */
int holo_render_mlp_fwd(const float* feats, const float* view_dirs, long long n_points, int C,
                        const float* packed_mlp, int hidden, int n_harmonic, const float* head_w,
                        const float* head_b, int n_head, float* densities, float* features, void* stream);
/* EmissionAbsorptionRaymarcher.forward (pytorch3d 0.7.4; configs/base.yaml:149-159; invoked
 * holo_multipass_ea.py:96-100) for surface_thickness 1, replicate_last_interval False, density_relu True,
 * blend_output False.  densities (n_rays,S), features (n_rays,S,feat_dim), lengths (n_rays,S); density_noise
 * (n_rays,S) or NULL is added before the relu (density_noise_std * randn drawn by the caller, :87-91);
 * normals (n_rays,S,3) or NULL -> out_normals (n_rays,3) = sum_s w n (:104-109).  bg_host: n_bg = 1 or feat_dim. */
int holo_ea_raymarch(const float* densities, const float* features, const float* lengths,
                     const float* density_noise, const float* normals, int n_rays, int S, int feat_dim,
                     const float* bg_host, int n_bg, float background_opacity, float* out_features,
                     float* out_depths, float* out_masks, float* out_weights, float* out_normals, void* stream);
/* RayPointRefiner.forward + sample_pdf (pytorch3d 0.7.4; configs/base.yaml:142-146; invoked
 * holo_multipass_ea.py:115-116).  u (n_rays, n_fine) caller-drawn uniforms (training, stratified) or NULL =
 * linspace(0, 1, n_fine) (evaluation).  lengths_out (n_rays, S + n_fine) sorted (n_fine if !add_input_samples). */
int holo_ray_refine(const float* lengths, const float* weights, const float* u, int n_rays, int S, int n_fine,
                    int add_input_samples, float* lengths_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Denoiser (guided_diffusion UNetModel, unet.py:566-837) building blocks
 * ------------------------------------------------------------------------------------------------------- */

/* (rows, cols) -> (cols, rows); NCDHW <-> DHWC is transpose2d over (C, V). */
int holo_transpose2d(const float* src, float* dst, int rows, int cols, void* stream);

/* GroupNorm32(32, C) -- nn.py:23-25,99-106 -- over channels-last x = cat(x1 (V,C1), x2 (V,C2)) (th.cat, unet.py:829).
 * stats accumulates (sum, sumsq) per group in acc64[8][32][2] (8 replicas; must start zeroed; finalize re-zeroes it);
 * finalize turns them into the per-channel affine y = x*a + b, folding gamma/beta and, when
 * film_scale_shift (2C: scale then shift) is given, the FiLM h*(1+scale)+shift of unet.py:248-252;
 * apply writes act(x*a+b) (act = SiLU, unet.py:184,208) as fp32 and/or as a bf16 hi/lo pair. */
int holo_gn_stats(const float* x1, int C1, const float* x2, int C2, long long V, double* acc64, void* stream);
int holo_gn_finalize(double* acc64, const float* gamma, const float* beta, const float* film_scale_shift, int C,
                     long long V, float eps, float* a, float* b, void* stream);
int holo_gn_apply(const float* x1, int C1, const float* x2, int C2, long long V, const float* a, const float* b,
                  int silu, float* y, void* y_hi_bf16, void* y_lo_bf16, void* stream);
/* Two-launch form used by the UNet executor: stats_pp accumulates into acc64 and clears acc64_next (ping-pong, so no
 * separate zeroing launch); apply_fused derives the affine inside every CTA (finalize folded in) and applies it. */
int holo_gn_stats_pp(const float* x1, int C1, const float* x2, int C2, long long V, double* acc64, double* acc64_next,
                     void* stream);
int holo_gn_apply_fused(const float* x1, int C1, const float* x2, int C2, long long V, const double* acc64,
                        const float* gamma, const float* beta, const float* film_scale_shift, float eps, int silu,
                        float* y, void* y_hi, void* y_lo, void* raw_hi, void* raw_lo, int pair_f16,
                        void* stream);
/* raw_hi/raw_lo (optional, C1 % 8 == C2 % 8 == 0): the UN-normalised cat(x1, x2) as a hi/lo pair as well -- the
 * operand of the ResBlock's 1x1 skip convolution (unet.py:222,255) -- written in the same pass over the tensor.
 * pair_f16 = 1 writes the pairs (y and raw) as saturating fp16 halves instead of bf16 (holo_conv3d_tc operand_fmt).
 * Same, from the per-channel statistics the producing convolutions left behind (holo_conv3d_tc stats_ch): no
 * statistics pass over the tensor at all.  ch_stats2 belongs to the second source of the concat. */
int holo_gn_apply_fused_ch(const float* x1, int C1, const double* ch_stats1, const float* x2, int C2,
                           const double* ch_stats2, long long V, const float* gamma, const float* beta,
                           const float* film_scale_shift, float eps, int silu, float* y, void* y_hi,
                           void* y_lo, void* raw_hi, void* raw_lo, int pair_f16, void* stream);
/* fp32 cat(x1 (V,C1), x2 (V,C2)) -> hi/lo pair (Vout,Cpad): consumes the skip concat in place, zero-pads channels
 * to Cpad; upsample2x folds F.interpolate(nearest, x2) of the (Din,Hin,Win) volume (Upsample.forward,
 * unet.py:94-97), Vout = 8 V.  The halves are bf16, or saturating fp16 with pair_f16 = 1. */
int holo_split_bf16(const float* x1, int C1, const float* x2, int C2, long long V, int Cpad, int upsample2x, int Din,
                    int Hin, int Win, void* hi, void* lo, int pair_f16, void* stream);

/* Exact-fp32 implicit-GEMM convolution (nn.Conv3d 3^3/1^3, stride 1|2, padding k/2; nn.Conv1d k=1):
 * unet.py:185,211,222,657,792 / Downsample :129-131 / Upsample :89-97 (upsample2x folds F.interpolate nearest x2)
 * / AttentionBlock qkv, proj_out :383,392.  Weights pre-packed as [tap][Cin][Cout]; out = conv + bias (+ residual). */
int holo_conv3d_simt(const float* x1, int C1, const float* x2, int C2, int Din, int Hin, int Win, int ksize,
                     int stride, int upsample2x, const float* w_tap_cin_cout, const float* bias,
                     const float* residual, int Cout, float* out, void* stream);

/* tcgen05 (5th-gen tensor core) implicit-GEMM convolution, 3-term split operands (hi.hi + hi.lo + lo.hi), fp32 TMEM
 * accumulation.  Same contract as holo_conv3d_simt for ksize 1|3, stride 1|2 (Downsample.op, unet.py:129-131), one
 * source; operands are 16-bit hi/lo pairs: x_hi/x_lo (V,Cin) channels-last over the INPUT volume (D,H,W), w_hi/w_lo
 * [Cout][tap][Cin] (K-major).  operand_fmt selects the number format of all four halves (one kind::f16 UMMA cannot
 * mix bf16 with fp16): 0 = bf16 pairs (fp32's range, exact to 2^-17: "3xBF16"); HOLO_FMT_F16 = fp16 pairs (exact to
 * 2^-22; activations saturate beyond |x| = 131008; the weights are the pair of s*w with s a power of two chosen by
 * the caller so that s*max|w| ~ 2^10, which keeps w_lo out of the subnormals, and acc_scale = 1/s undoes it:
 * acc_scale multiplies the accumulators before bias and residual; pass 1 otherwise).  fp16 pairs bring the UNet's
 * distance to exact arithmetic from 7e-5 to 4e-6 -- fp32's own -- at the same MMA count.
 * out_hi/out_lo (optional): the result as an operand pair in the same format (the attention's q, k).
 * Cin % 64 == 0, Cout % 16 == 0, output dims multiples of (4,4,4).  Returns HOLO_ERR_UNSUPPORTED (-3) for
 * shapes it does not take.  stats_ch (optional, [Cout][2] doubles, pre-zeroed): per-channel (sum, sumsq) of the output
 * for the GroupNorm that consumes it, accumulated in the epilogue.  Small grids are split over K: with splitk_partials
 * (holo_conv3d_tc_splitk_bytes() bytes of scratch) every K slice parks its partial tile there and a second launch sums
 * the slices in slice order, applies scale / bias / residual and writes `out` once -- a deterministic result, no
 * zero-fill, statistics produced; without it the slices add into a zeroed `out` with fp32 atomics (summation order varies
 * run to run at the 1e-7 level) and the call returns 1 instead of 0 = done, but the statistics were NOT produced.  Long K loops are accumulated in chains (HOLO_CONV_CHUNK iterations, default 9) that
 * the epilogue sums in registers: the tensor core truncates every add into the TMEM accumulator. */
#define HOLO_FMT_F16 1
int holo_conv3d_tc(const void* x_hi, const void* x_lo, int Cin, int D, int H, int W, int ksize, int stride,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, int Cout, float* out,
                   void* out_hi, void* out_lo, double* stats_ch, int operand_fmt, float acc_scale,
                   float* splitk_partials, void* stream);
long long holo_conv3d_tc_splitk_bytes(void);
/* Debug aid (library built with -DHOLO_CONV_TRACE, e.g. HOLO_NVCC_FLAGS=-DHOLO_CONV_TRACE python
 * holo_diffusion_b200/build.py; a no-op otherwise): CTA 0 of every following tcgen05 convolution launch writes 8 clock64 stamps (entry, set-up done, first
 * TMA issued, first operands landed, last MMA issued, first accumulator ready, first item written, exit) into the 8
 * int64 at dev_buf8; NULL switches it off. */
int holo_debug_conv_trace(void* dev_buf8);

/* ResBlock tail in ONE launch (ResBlock._forward, unet.py:254-256, with a 1x1 skip_connection :222):
 *   out = conv3^3(x) + conv1^1(skip_x) + bias (+ residual)
 * The skip convolution rides the 3^3 convolution's TMEM accumulator as Cin_skip / 64 extra K iterations.  w_hi / w_lo:
 * [Cout][27 * Cin + Cin_skip] pairs (taps first, then the 1x1 columns) under one common scale; bias = the sum of the two
 * biases.  Same shape rules, formats, statistics and return codes as holo_conv3d_tc (stride 1, Cin_skip % 64 == 0).
 * The executor's default for ResBlocks with a skip convolution (HOLO_FUSE_SKIP=0 turns it off). */
int holo_conv3d_tc_skip(const void* x_hi, const void* x_lo, int Cin, const void* skip_hi, const void* skip_lo,
                        int Cin_skip, int D, int H, int W, const void* w_hi, const void* w_lo, const float* bias,
                        const float* residual, int Cout, float* out, double* stats_ch, int operand_fmt,
                        float acc_scale, float* splitk_partials, void* stream);

/* Plain GEMM on the tcgen05 kernel: out[m][n] = bias[n] + residual[m][n] + sum_k a[m][k] b[n][k]; a, b are 16-bit
 * hi/lo pairs (operand_fmt as for holo_conv3d_tc; out_hi/out_lo are written in the same format), K-major with
 * arbitrary row pitches (elements).  M % 128 == 0, K % 64 == 0, N % 16 == 0.
 * out_is_zeroed = 1 promises a zero-filled fp32 `out`, which lets small grids split K and accumulate with atomics.
 * With holo_softmax_split / holo_transpose_split_bf16 it carries QKVAttentionLegacy (unet.py:438-455) on tensor
 * cores: S = Q K^T, P = softmax(scale2 * S) (fp32, one key row per CTA, emitted as a hi/lo pair), O = P V. */
int holo_gemm_tc(const void* a_hi, const void* a_lo, long long a_pitch, int M, int K, const void* b_hi,
                 const void* b_lo, long long b_pitch, int N, const float* bias, const float* residual,
                 long long out_pitch, float* out, void* out_hi, void* out_lo, int out_is_zeroed, int operand_fmt,
                 float acc_scale, void* stream);
/* The same GEMM with the consumer's elementwise work in its epilogue:
 *   out[m][n] = act(acc_scale * sum_k a[m][k] b[n][k] + bias[n] + row_term[m % row_term_rows][n]),
 * act: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 Softplus; row_term (row_term_rows, N) fp32 with the output's pitch, or NULL;
 * fp32 `out` and / or the operand pair out_hi / out_lo of the next GEMM.  N % 64 == 0, K is never split.
 * Carries the Linear layers of MLPMeanFeatureAggregator (custom_modules.py:255-264): the mean term of a point is
 * shared by that point's row in every view (rows are view-major), the hidden activations never visit HBM as fp32. */
int holo_gemm_tc_act(const void* a_hi, const void* a_lo, long long a_pitch, int M, int K, const void* b_hi,
                     const void* b_lo, long long b_pitch, int N, const float* bias, const float* row_term,
                     long long row_term_rows, int act, long long out_pitch, float* out, void* out_hi, void* out_lo,
                     int operand_fmt, float acc_scale, void* stream);
/* P = p_scale * softmax(scale2 * S): with fp16 halves pass p_scale = 4096 (normalised probabilities of long rows sit
 * in fp16's subnormal range) and give the P V GEMM acc_scale = 1 / 4096; 1 otherwise. */
int holo_softmax_split(const float* S, int n_rows, int T, float scale2, void* P_hi, void* P_lo, int pair_f16,
                       float p_scale, void* stream);
int holo_transpose_split_bf16(const float* src, long long src_pitch, int rows, int cols, void* hi, void* lo,
                              int pair_f16, void* stream);

/* Fused (flash-style) QKVAttentionLegacy.forward -- unet.py:438-455 -- on tcgen05: one launch for all heads, the
 * T x T logits stay in TMEM / shared memory (exact two-pass softmax, 3-term operand splits, fp32 accumulation).
 * qkv_hi/lo (T, heads*3*ch) hi/lo pair of the head-major [q|k|v] tensor (the qkv convolution's split output);
 * vt_hi/lo (heads*ch, T) hi/lo pair of V^T, produced by holo_v_transpose_split from the fp32 qkv tensor.
 * out_cl (T, heads*ch) fp32 and/or out_hi/lo its split (what the projection conv consumes); either may be NULL.
 * pair_f16 selects the 16-bit format of EVERY pair (inputs, the probabilities inside, the output pair): 0 = bf16
 * halves, 1 = fp16 halves (logits exact to 2^-22 instead of 2^-17 -- softmax turns their absolute error into a
 * relative error of the probabilities).  softmax_scale: logits are softmax_scale * q.k; <= 0 selects the reference's
 * ch^-1/2 (heads narrower than 64 channels run zero-padded to 64 and pass their true ch^-1/2 here).
 * q_begin / q_count: only queries [q_begin, q_begin + q_count) are computed and written (rows of the full-size
 * outputs) -- the query-sharded multi-GPU form, every rank holding all keys / values; q_begin % 128 == 0, the range
 * ends on a multiple of 128 or at T; q_count <= 0 = all queries.
 * kv_splits > 1 (opt-in): that many CTAs share the key tiles of one (query tile, head), each with its own softmax
 * stabiliser, and a small merge kernel combines the partial results -- fills the 148 SMs when (T / 128) x heads is
 * small (64 CTAs at T = 4096).  Needs `workspace` of holo_attention_flash_workspace_bytes() bytes (device memory).
 * ch in {64, 128}, T % 64 == 0; other shapes return HOLO_ERR_UNSUPPORTED (-3). */
int holo_v_transpose_split(const float* qkv_cl, int T, int heads, int ch, void* vt_hi, void* vt_lo, int pair_f16,
                           void* stream);
int holo_attention_flash(const void* qkv_hi, const void* qkv_lo, const void* vt_hi,
                         const void* vt_lo, int T, int heads, int ch, float* out_cl, void* out_hi,
                         void* out_lo, int pair_f16, float softmax_scale, int q_begin, int q_count, int kv_splits,
                         void* workspace, void* stream);
long long holo_attention_flash_workspace_bytes(int T, int heads, int ch, int kv_splits);

/* QKVAttentionLegacy.forward -- unet.py:438-455.  qkv_cl (T, heads*3*ch) head-major [q|k|v]; out_cl (T, heads*ch). */
int holo_attention_simt(const float* qkv_cl, int T, int heads, int ch, float* out_cl, void* stream);

/* timestep_embedding -- nn.py:109-127 -- and the small Linear layers (time_embed unet.py:646-650, emb_layers :199-205):
 * out[m][o] = act_out(b[o] + sum_i W[o][i] * act_in(x[m][i])) with act = SiLU when the flag is set.
 * freqs (dim/2): exp(-ln(1e4) * i / (dim/2)), built by the caller on the host as the reference does (nn.py:119-121). */
int holo_timestep_embedding(const long long* t_i64, int n, int dim, const float* freqs, float* out, void* stream);
int holo_linear_rows(const float* x, const float* W, const float* b, int M, int in_dim, int out_dim, int silu_in,
                     int silu_out, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Backward of the staged renderer (renderer half): what objective.backward()
 * (trainer/training_loop.py:518-556) propagates through the ray marcher and the implicit function for the mask_sample
 * rays of a training step (configs/base.yaml:132-134).  The refiner runs under no_grad; ray lengths carry no gradient.
 * ------------------------------------------------------------------------------------------------------- */

/* EmissionAbsorptionRaymarcher backward.  Inputs as holo_ea_raymarch (density_noise = the noise the forward added, or
 * NULL); bg_dev (feat_dim) on the DEVICE; grad_* = dL/d(out_features (n,F), out_depths (n), out_masks (n), out_weights
 * (n,S)), the last three optional.  d_densities (n,S), d_features (n,S,F). */
int holo_ea_raymarch_bwd(const float* densities, const float* features, const float* lengths, const float* density_noise,
                         int n_rays, int S, int feat_dim, const float* bg_dev, float background_opacity,
                         const float* grad_features, const float* grad_depths, const float* grad_masks,
                         const float* grad_weights, float* d_densities, float* d_features, void* stream);

/* HoloVoxelGridImplicitFunction backward for ray points (same inputs as holo_if_fwd in ray mode, no feature head):
 * grad_densities (P), grad_rgb (P,3) -> ACCUMULATED into caller-zeroed d_grid_dhwc (D,H,W,C) (trilinear scatter-add,
 * grid_sampler_3d backward), d_W_eff (hidden+1, C), d_b_eff (hidden+1) of the collapsed density net and d_Wr (3, hidden +
 * E), d_br (3) of the radiance layer.  Forward activations are recomputed. */
int holo_if_bwd(const float* grid_dhwc, int D, int H, int W, int C, float volume_extent, const float* packed_mlp,
                int hidden, int n_harmonic, const float* origins, const float* dirs, const float* lengths,
                long long n_points, int S, const float* grad_densities, const float* grad_rgb, float* d_grid_dhwc,
                float* d_W_eff, float* d_b_eff, float* d_Wr, float* d_br, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Whole-graph denoiser: UNetModel.forward (unet.py:800-837) as configured by SimpleUnet3D
 * (utils/diffusion_utils.py:41-86) in ONE call, for hosts without Python (csrc/unet_exec.cu walks the block list and
 * issues the launches above in the order the Python executor does).
 *   holo_unet_create         block list + parameter table + workspace plan from the configuration
 *   holo_unet_param_*        enumerate the parameters: names are the reference's state-dict keys below `_net.`
 *                            ("input_blocks.1.0.in_layers.2.weight", ...), numel their element counts
 *   holo_unet_set_param      fp32 DEVICE pointer of one checkpoint tensor (borrowed; must stay valid)
 *   holo_unet_pack           derived weight layouts (16-bit operand pairs, CUDA-core layouts, concatenated FiLM
 *                            projection) into the caller's `packed` buffer of holo_unet_packed_bytes(); synchronises the
 *                            stream once; call again after parameters change
 *   holo_unet_fwd            x (1, C, D, H, W) fp32, t (1) int64 on the device -> out (1, C_out, D, H, W); asynchronous,
 *                            graph-capturable; `workspace` of holo_unet_workspace_bytes() device bytes
 *   holo_unet_fwd_cl         the same on channels-last tensors (V, C) -> (V, C_out) (no layout passes)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct holo_unet_config {
    int in_channels, model_channels, out_channels, num_res_blocks;
    int n_levels;
    int channel_mult[8];
    int n_attention_resolutions;
    int attention_resolutions[8];   /* downsampling factors that carry attention (SimpleUnet3D.attention_resolutions) */
    int num_heads;
    int D, H, W;                    /* grid of one sample */
    int pair_f16;                   /* 1 = fp16 operand pairs (default of the Python executor), 0 = bf16 pairs */
    int fuse_skip;                  /* 1 = ResBlock tails with a skip convolution as one launch (holo_conv3d_tc_skip) */
    int attn_kv_split;              /* 0 = auto (fill the SMs), n >= 1 = CTAs sharing the keys of one query tile */
    int use_tensor_cores;           /* 0 = exact-fp32 CUDA-core kernels everywhere */
    int splitk_workspace;           /* 1 = split-K convolutions reduce deterministically through scratch (and produce the
                                       GroupNorm statistics); 0 = fp32 atomics + separate statistics passes */
} holo_unet_config;
int holo_unet_create(const holo_unet_config* cfg, void** handle);
int holo_unet_destroy(void* handle);
int holo_unet_param_count(void* handle);
const char* holo_unet_param_name(void* handle, int index);
long long holo_unet_param_numel(void* handle, int index);
int holo_unet_set_param(void* handle, const char* name, const float* dev_ptr, long long numel);
long long holo_unet_packed_bytes(void* handle);
long long holo_unet_workspace_bytes(void* handle);
int holo_unet_pack(void* handle, void* packed_dev, void* stream);
int holo_unet_fwd(void* handle, const float* x_ncdhw, const long long* t_dev, float* out_ncdhw, void* workspace,
                  void* stream);
int holo_unet_fwd_cl(void* handle, const float* x_cl, const long long* t_dev, float* out_cl, void* workspace,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Diffusion (gaussian_diffusion.py) and model glue (holo_diffusion_model.py)
 * ------------------------------------------------------------------------------------------------------- */

/* p_sample for ModelMeanType.START_X / ModelVarType.FIXED_SMALL -- gaussian_diffusion.py:459-508 (+ :318,:229-251).
 * coef1/coef2/logvar: device fp32 copies of posterior_mean_coef1/2 and posterior_log_variance_clipped (:150-187),
 * indexed on the device by t (n_batch int64).  noise may be NULL (treated as 0). */
int holo_ddpm_step(const float* model_out, const float* x_t, const float* noise, const long long* t_i64,
                   const float* coef1, const float* coef2, const float* logvar, long long per_sample, int n_batch,
                   int clip_denoised, float* x_prev, float* pred_xstart, void* stream);
/* ddim_sample -- gaussian_diffusion.py:645-693 -- and ddim_reverse_sample -- :695-731 -- for a START_X model.
 * alphas_cumprod_to = alphas_cumprod_prev (forward; noise (may be NULL) scaled by sigma(eta)) or alphas_cumprod_next
 * (reverse ODE: pass eta = 0, noise = NULL). */
int holo_ddim_step(const float* model_out, const float* x_t, const float* noise, const long long* t_i64,
                   const float* alphas_cumprod, const float* alphas_cumprod_to,
                   const float* sqrt_recip_alphas_cumprod, const float* sqrt_recipm1_alphas_cumprod, float eta,
                   long long per_sample, int n_batch, int clip_denoised, float* x_out, float* pred_xstart,
                   void* stream);
/* q_sample -- gaussian_diffusion.py:209-227. */
int holo_q_sample(const float* x0, const float* noise, const long long* t_i64, const float* sqrt_ac,
                  const float* sqrt_1m_ac, long long per_sample, int n_batch, float* out, void* stream);

/* tanh / clamp + the range statistics behind `assert voxel_features.min() >= -1 and ...max() <= 1`
 * (holo_diffusion_model.py:381,424-428): one pass writes the channels-last copy (renderer input), the
 * channels-first copy (API tensor) and min/max/nan-count into stats4 (ordered-int encoding; see range_init). */
int holo_range_init(int* stats4, void* stream);
int holo_act_range(const float* x_cl, long long V, int C, int act, float* y_cl, float* y_cf, int* stats4,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Per-view post-processing of generate_samples.py
 * ------------------------------------------------------------------------------------------------------- */

/* Depth visualisation of one view: pytorch3d vis_utils.make_depth_image (exact order statistics at min_quantile /
 * 1 - max_quantile over the pixels with depth > 1e-6 and mask > 0.5, normalisation into [min_out, max_out], * mask,
 * clamp) followed by the white compositing v * mask + (1 - mask) and the repeat to 3 channels of
 * _images_from_preds -- utils/render_utils/flyaround.py:470-479.  depth, mask (n_pixels); normfac2 (2) receives the
 * selected (min, max); out_3n (3, n_pixels). */
int holo_depth_image(const float* depth, const float* mask, int n_pixels, float min_quantile, float max_quantile,
                     float min_out_depth, float max_out_depth, int composite_white, float* normfac2, float* out_3n,
                     void* stream);

/* One video frame: (C in {1,3}, H, W) float -> (out_h, out_w, 3) uint8: clip to [0, 1], bilinear resize, round.
 * Replaces rendered_pred[k][0].clip(0, 1).cpu().numpy() + VideoWriter.write_frame(resize) --
 * utils/render_utils/flyaround.py:588-595. */
int holo_frame_u8(const float* src_chw, int C, int H, int W, int out_h, int out_w, void* dst_hw3_u8, void* stream);

/* Shaded depth render of one view in screen space: box-smoothed depth (window 2 smooth_k + 1) -> camera-space vertex
 * grid through the NDC intrinsics (fx, fy, px, py) -> area-weighted vertex normals of the valid quads' triangles ->
 * Phong shading with a point light at the camera (material10 = ambient rgb, diffuse rgb, specular rgb, shininess),
 * background bg3 elsewhere.  Replaces depth_to_shaded(method="mesh") -- utils/render_utils/shaded_depth_render.py:143-206
 * (mesh construction :248-280, rendered from the mesh's own camera by mesh_render.py).  scratch_depth (H*W) floats,
 * scratch_ok_u8 (H*W) bytes; out_3hw (3,H,W); out_mask (H*W) or NULL. */
int holo_shade_depth(const float* depth, const float* mask, int H, int W, float fx, float fy, float px, float py,
                     int smooth_k, float mask_thr, float depth_thr, const float* material10_host, const float* bg3_host,
                     float* scratch_depth, void* scratch_ok_u8, float* out_3hw, float* out_mask, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * View-pooling encoder: source views -> voxel grid.
 *   holo_diffusion_model.py:327-373: VolumeLocator grid points -> view_pooler (pytorch3d ViewSampler +
 *   custom_modules.py:162-334 MLPMeanFeatureAggregator) -> pooled_feature_mapper -> tanh.
 * The Linear layers run on holo_gemm_tc; these three kernels are the gather and the reductions around them.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct holo_feature_map {
    const float* data;          /* channels-LAST (n_src, height, width, channels) fp32 on the device */
    int channels, height, width;
} holo_feature_map;

/* Project n_pts world points into the n_src source cameras, sample every feature map there (bilinear, zeros padding,
 * align_corners = False, NDC -> grid_sample coordinates as pytorch3d ndc_grid_sample), append the harmonic embedding
 * of the unit vector camera centre -> point, and multiply by the aggregation weight
 *   w[s][p] = view_weight[s] * (mask_map ? nearest(mask_map[s]) : 1)
 * -- ViewSampler.forward / project_points_and_sample (eps: clamp of the perspective divide) and
 * custom_modules.py:241-260,283-334.  Rows: X[(s * rows_per_view + p) * Kpad + j] as a 16-bit hi/lo pair (pair_f16 as
 * for holo_conv3d_tc), columns [features of map 0 | map 1 | ... | sin | cos | dir | zero padding to Kpad];
 * mean[p * Kpad + j] = sum_s X w / max(sum_s w, 1e-2) (wmean of _avgmaxstd_reduction_function, AVG).
 * x_f32 (n_src, n_pts, F + E) / mean_f32 (n_pts, F + E): optional fp32 copies.  maps: HOST array. Kpad <= 256.
 * Every map's channel count must be a multiple of 4 (pad with zero channels; data 16-byte aligned): one lane owns four
 * consecutive columns of one map and reads each bilinear tap as a float4. */
int holo_viewpool_sample(const float* pts, long long n_pts, const float* R, const float* T, const float* focal,
                         const float* pp, int n_src, const holo_feature_map* maps, int n_maps, const float* mask_map,
                         int Hm, int Wm, const float* view_weight, int n_harmonic, float eps, int Kpad,
                         long long rows_per_view, void* x_hi, void* x_lo, void* mean_hi, void* mean_lo, float* x_f32,
                         float* mean_f32, int pair_f16, void* stream);
/* pytorch3d AngleWeightedReductionFeatureAggregator (the view_pooler's default aggregator, configs/base.yaml:165-168
 * leaves it in place; restated from memory of pytorch3d 0.7.4): same projection / sampling as holo_viewpool_sample,
 *   w[s][p] = view_weight[s] * mask * ((0.5 (d_s . d_0 + 1))^gamma + min_ray_angle_weight)
 * (d_s: unit vector camera s -> point; camera 0 is the reference view), reduced over the views to
 * [wmean | sqrt(max(wvar, 1e-4))] per feature map (with_std = 0: the mean alone): out row p = hi/lo pair of
 * [mu_0 | std_0 | mu_1 | std_1 | ... | zero padding to Kpad]; out_f32 (n_pts, (1 + with_std) * F) optional. */
int holo_viewpool_angle_reduce(const float* pts, long long n_pts, const float* R, const float* T, const float* focal,
                               const float* pp, int n_src, const holo_feature_map* maps, int n_maps,
                               const float* mask_map, int Hm, int Wm, const float* view_weight, float eps, float gamma,
                               float min_ray_angle_weight, int with_std, int Kpad, void* out_hi, void* out_lo,
                               float* out_f32, int pair_f16, void* stream);
/* H = act(Y[s][p] + point_term[p]) as a hi/lo pair; Y (n_views * rows_per_view, C), point_term (rows_per_view, C) or
 * NULL.  act: 0 identity, 1 ReLU, 2 LeakyReLU(0.2), 3 Softplus -- the activations of MLPWithInputSkips
 * (custom_modules.py:62-88) between the aggregator's Linear layers (:255-261). */
int holo_viewpool_act_split(const float* y, const float* point_term, int n_views, long long rows_per_view, int C,
                            int act, void* hi, void* lo, int pair_f16, void* stream);
/* G[p] = sum_s softmax_s(Z[s][p][0]) Z[s][p][:] -- custom_modules.py:262-264; Z (n_views, rows_per_view, C);
 * out (n_pts, C) fp32 and/or its hi/lo pair (the operand of the pooled_feature_mapper GEMM). */
int holo_viewpool_reduce(const float* z, int n_views, long long rows_per_view, long long n_pts, int C, float* out,
                         void* out_hi, void* out_lo, int pair_f16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HOLO_B200_H */
