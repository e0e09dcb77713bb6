/* b200sd.h — C ABI of libb200sd.so, the sm_100a compute library behind the local-GPU worker.
 *
 * Boundary being replaced (reference = papuSpartan/stable-diffusion-webui-distributed @ 8fd65ebd):
 *   scripts/spartan/worker.py:288-504  Worker.request()  — the reference posts the job to a remote sdwui
 *   (`session.post(full_url("txt2img"|"img2img"))`, worker.py:432-435) whose process_images() runs
 *   UNet x steps + VAE decode.  The reference contains none of that arithmetic; these
 *   entry points are what a local executor binds instead of the HTTP call: every function below is one
 *   op of the per-step eps-prediction / final decode (SURVEY.md §8 a-ext x1..x11), taking raw device
 *   pointers and an explicit stream.
 *
 * Conventions
 *   - caller owns all device memory; the library never allocates, never synchronises, never throws
 *   - activations are NHWC ("pixels x channels") fp16 (dtype 0) or bf16 (dtype 1); `ld*` / `pitch` are row
 *     pitches in ELEMENTS, so channel slices of wider buffers (skip-concat buffers) are addressed in place
 *   - weights are pre-packed [N][K] (K contiguous); conv weights [Cout][ky][kx][Cin]
 *   - all pointers 16-byte aligned, pitches multiples of 8 elements
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream)
 *   - return value: 0 ok, <0 error (B200SD_ERR_*); the call launched nothing if it failed
 */
#ifndef B200SD_H_
#define B200SD_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200SD_OK 0
#define B200SD_ERR_INVALID (-1)
#define B200SD_ERR_CUDA (-2)
#define B200SD_ERR_TMAP (-3)
#define B200SD_ERR_UNSUPPORTED (-4)

#define B200SD_F16 0
#define B200SD_BF16 1

/* epilogue flags for b200sd_linear / b200sd_conv2d */
#define B200SD_EPI_GEGLU 1 /* tile columns [0,bn/2) = value, [bn/2,bn) = gate: out = v * gelu_erf(g) */
#define B200SD_EPI_SILU 2  /* out = silu(acc + bias (+ residual)) */

typedef struct b200sd_epilogue {
  const float* bias;      /* [groups][N] fp32 or NULL */
  int bias_group_rows;    /* output rows sharing one bias row (H*W for a per-image bias); <=0: one row */
  const void* residual;   /* [M][N_out] same dtype as the output, or NULL */
  long long ldr;          /* residual row pitch (elements) */
  int flags;              /* B200SD_EPI_* */
} b200sd_epilogue;

/* library / build identification: returns a static string "b200sd <version> sm_100a" */
const char* b200sd_version(void);
/* debug: a device buffer of 16 x 64 int64 that one CTA of every later b200sd_attention launch fills with clock64()
 * stamps of its per-tile pipeline events (tools/attn_trace.py); NULL switches it off (default). */
int b200sd_debug_attention_trace(void* device_buffer);
/* debug only: clock64 timeline of CTA 0 of the GEMM / conv kernel (tools/gemm_trace.py; trace-enabled builds). */
int b200sd_debug_gemm_trace(void* device_buffer);

/* ---- tensor-core ops (tcgen05 + TMA) ----------------------------------------------------------- */

/* D[M,N_out] = epi(A[M,K] . Wt[N,K]^T).  Linear layers and 1x1 convs (upstream ldm CrossAttention.to_q/k/v/
 * to_out, FeedForward.net, SpatialTransformer.proj_in/out, ResBlock.skip_connection).
 * K % 64 == 0, N % block_n == 0, block_n in {32,64,...,256}.  max_ctas <= 0: one CTA per SM. */
int b200sd_linear(const void* A, long long lda, const void* Wt, void* D, long long ldd, int M, int N, int K,
                  int block_n, const b200sd_epilogue* epi, int dtype, int max_ctas, void* stream);

/* NHWC convolution as implicit GEMM: X[NB,Hin,Win,C] (channel pitch `pitch_c`), Wt[Cout][k*k*C],
 * ksize in {1,3}, stride in {1,2}, zero padding `pad` before / `pad_end` after each spatial dim.
 * D rows are output pixels in (n, y, x) order.  (upstream ResBlock.in_layers/out_layers conv, Upsample.conv,
 * Downsample.op, AutoencoderKL Decoder/Encoder convs.)  C % 64 == 0. */
int b200sd_conv2d(const void* X, long long pitch_c, int NB, int Hin, int Win, int C, const void* Wt, int ksize,
                  int stride, int pad, int pad_end, void* D, long long ldd, int Cout, int block_n,
                  const b200sd_epilogue* epi, int dtype, int max_ctas, void* stream);

/* O[b,s,h*d] = softmax(Q K^T * scale) V per (batch, head); Q/K/V rows are tokens, head h occupies columns
 * [h*d_pad, h*d_pad+d) (zero padded to d_pad, a multiple of 64).  (upstream CrossAttention.forward)
 * v_ones_col != 0 (needs d < d_pad): V[:, h*d_pad + d] == 1 for every head — the P.V tensor-core product then also
 * yields the softmax denominators (column d of the accumulator), so no CUDA-core row sums are computed. */
int b200sd_attention(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                     void* O, long long ldo, int B, int heads, int Sq, int Skv, int d, int d_pad, float scale,
                     int v_ones_col, int dtype, void* stream);

/* ---- HBM-bound ops ------------------------------------------------------------------------------- */

/* GroupNorm statistics: stats[n][g] = (sum, sumsq) over the group's channels and all HW pixels, bit-reproducible
 * run to run (per-CTA partial sums combined in a fixed order; no floating-point atomics).
 * `stats` holds b200sd_groupnorm_stats_floats() floats: the [NB][G][2] results first, then the kernel's scratch.
 * The buffer must be zero-filled once when it is allocated (the scratch contains arrival counters that the kernel
 * leaves at zero); it may be shared by successive calls on one stream. (upstream GroupNorm32 / Normalize) */
long long b200sd_groupnorm_stats_floats(int NB, int HW, int C, int G);
int b200sd_groupnorm_stats(const void* X, long long pitch, int NB, int HW, int C, int G, float* stats, int dtype,
                           void* stream);
/* Y = (X - mean) * rstd * gamma + beta, optional SiLU; mean/rstd from `stats`. */
int b200sd_groupnorm_apply(const void* X, long long pitch_x, void* Y, long long pitch_y, int NB, int HW, int C, int G,
                           const float* stats, const float* gamma, const float* beta, float eps, int silu, int dtype,
                           void* stream);
/* GroupNorm in one call.  mode 1: b200sd_groupnorm_stats + b200sd_groupnorm_apply.  mode 2: the one-pass kernel — each CTA
 * keeps its ~48 KB slab of an image in shared memory while the image's CTAs agree on the statistics (1 read + 1 write
 * instead of 2 reads + 1 write) — or B200SD_ERR_UNSUPPORTED when the shape is not eligible (C <= 2048, G <= 64 and all
 * slabs of ONE image resident on the device at the same time: a function of HW and C only, never of NB, so an image's bits
 * do not depend on the batch it travels in).  mode 0: mode 1, unless B200SD_GN_FUSED=1 and the shape is eligible (measured
 * on B200 the one-pass kernel is the slower one at the UNet's shapes: the cross-CTA agreement costs more than the second
 * read it saves).  `stats` as for b200sd_groupnorm_stats. */
int b200sd_groupnorm(const void* X, long long pitch_x, void* Y, long long pitch_y, int NB, int HW, int C, int G,
                     float* stats, const float* gamma, const float* beta, float eps, int silu, int mode, int dtype,
                     void* stream);
/* 1 if modes 0 / 2 take the one-pass kernel for this shape on the current device. */
int b200sd_groupnorm_is_fused(int NB, int HW, int C, int G, int dtype);
/* measurement aid (tools/norm_sweep.py): slab KB of the one-pass kernel, 8..160 (0 = keep; default 48, env
 * B200SD_GN_FUSED_KB) — call b200sd_groupnorm_stats_floats again afterwards — and whether the statistics kernel walks the
 * tensor back to front (-1 = keep; default 1, env B200SD_GN_REVERSE). */
int b200sd_debug_gn_config(int slab_kb, int reverse_stats);
/* LayerNorm over the last dim of [rows, C]. (upstream BasicTransformerBlock.norm1/2/3) */
int b200sd_layernorm(const void* X, long long ldx, void* Y, long long ldy, int rows, int C, const float* gamma,
                     const float* beta, float eps, int dtype, void* stream);
/* nearest-neighbour 2x upsample, NHWC. (upstream Upsample before its conv) */
int b200sd_upsample2x(const void* X, long long pitch_x, void* Y, long long pitch_y, int NB, int H, int W, int C,
                      int dtype, void* stream);
/* row softmax of an fp16/bf16 matrix in place, fp32 math (VAE mid-block attention, d=512 single head). */
int b200sd_softmax_rows(void* S, long long lds, int rows, int cols, float scale, int dtype, void* stream);
/* Y[rows, C] = silu(X) elementwise (emb_layers SiLU). */
int b200sd_silu(const void* X, void* Y, long long n, int dtype, void* stream);

/* ---- sampler / conditioning / output ------------------------------------------------------------- */

/* sinusoidal timestep embedding (cos first, then sin), out [T][dim] fp16/bf16. (ldm timestep_embedding) */
int b200sd_timestep_embedding(const float* t, int T, int dim, void* out, long long ldo, int dtype, void* stream);
/* table[step][c] (fp32) = conv_bias[c] + emb[step][c]; folds the per-step time embedding into conv biases. */
int b200sd_fold_bias(const void* emb, long long lde, const float* bias, float* table, int T, int C, int dtype,
                     void* stream);
/* cur[0..n) = table[*step_counter][0..n) : selects this sampler step's rows on the device (graph replay safe). */
int b200sd_select_step(const float* table, long long row_len, const int* step_counter, float* cur, void* stream);
/* latents fp32 NHWC [B,HW,4] -> UNet input [2B,HW,pitch] (cond half and uncond half identical, channels >= 4
 * untouched: they are zero from allocation) */
int b200sd_pack_unet_input(const float* x, void* xin, long long pitch, int B, int HW, float in_scale, int dtype,
                           void* stream);
/* classifier-free guidance + one DDIM (eta = 0) update, then re-pack the next UNet input and advance
 * *step_counter.  eps [2B,HW,pitch_e] (cond first, uncond second), coef[step] = {sqrt(a_t), sqrt(1-a_t),
 * sqrt(a_prev), sqrt(1-a_prev)}.  (sdwui CFGDenoiser + sd_samplers_timesteps_impl.ddim) */
int b200sd_cfg_ddim_step(const void* eps, long long pitch_e, float* x, void* xin, long long pitch_x, int B, int HW,
                         float cfg_scale, const float* coef, int* step_counter, int dtype, void* stream);
/* Euler-ancestral step on sigma-space latents (k-diffusion sample_euler_ancestral); noise may be NULL when
 * sigma_up == 0.  coef[step] = {sigma, sigma_next_down, sigma_up, in_scale_next}. */
int b200sd_cfg_euler_a_step(const void* eps, long long pitch_e, float* x, const float* noise, void* xin,
                            long long pitch_x, int B, int HW, float cfg_scale, const float* coef, int* step_counter,
                            int dtype, void* stream);
/* DPM-Solver++(2M) step on sigma-space latents (k-diffusion sample_dpmpp_2m; sdwui "DPM++ 2M" / "DPM++ 2M Karras"):
 * old_denoised [B,HW,4] fp32 carries the previous step's x0 prediction (ignored when c2 == 0).
 * coef[step] = {sigma, sigma_next/sigma, c1, c2, in_scale_next, 0, 0, 0} — 8 floats per row. */
int b200sd_cfg_dpmpp_2m_step(const void* eps, long long pitch_e, float* x, float* old_denoised, void* xin,
                             long long pitch_x, int B, int HW, float cfg_scale, const float* coef, int* step_counter,
                             int dtype, void* stream);
/* ---- generic sampler building blocks: every sampler of the reference's ETA table (scripts/spartan/worker.py:75-94)
 * beyond the four fused ones above is, per model evaluation, one b200sd_cfg_eps plus a few b200sd_latent_lincomb on
 * fp32 NHWC latents [B,HW,4], with coefficient rows selected on the device by *step_counter (graph replay safe). */
/* e = eps_uncond + cfg_scale * (eps_cond - eps_uncond), fp32 [B,HW,4].  (sdwui CFGDenoiser; equals k-diffusion
 * to_d(x, sigma, denoised) for the eps-prediction CompVisDenoiser) */
int b200sd_cfg_eps(const void* eps, long long pitch_e, float* e, int B, int HW, float cfg_scale, int dtype, void* stream);
/* dst = sum_{k<n_src} c[k] * srcs[k], c = coef + (*step_counter) * ld + col0.  srcs / idx_strides are HOST arrays of
 * n_src (<= 8) device pointers / element strides: a source with idx_strides[k] != 0 is a stack of tensors and tensor
 * (int)coef[(*step_counter) * ld + idx_col] of it is read (per-step noise draws).  xin != NULL: dst * c[n_src] is also
 * packed as the next UNet input [2B,HW,pitch_x].  dst may alias a source. */
int b200sd_latent_lincomb(float* dst, const float* const* srcs, const long long* idx_strides, int n_src,
                          const float* coef, int ld, int col0, int idx_col, const int* step_counter, void* xin,
                          long long pitch_x, int B, int HW, int dtype, void* stream);
/* *step_counter += 1 on the device. */
int b200sd_bump_step(int* step_counter, void* stream);
/* decoded image [B,HW,pitch] (first 3 channels RGB in [-1,1]) -> uint8 [B,HW,3]:
 * trunc(255 * clamp((v+1)/2, 0, 1))  (sdwui process_images_inner) */
int b200sd_quantize_u8(const void* img, long long pitch, unsigned char* out, int B, int HW, int dtype, void* stream);

/* img2img input: uint8 [B,HW,3] RGB -> [B,HW,pitch] with channel c < 3 = 2*x/255 - 1 (channels >= 3 untouched: zero
 * from allocation).  (sdwui StableDiffusionProcessingImg2Img.init) */
int b200sd_image_to_nhwc(const unsigned char* img, void* out, long long pitch, int B, int HW, int dtype, void* stream);
/* VAE encoder moments [B,HW,pitch] (first 4 channels = posterior mean) -> scaled latents fp32 [B,HW,4] = mean * scale
 * (AutoencoderKL.encode(...).mean * scale_factor) */
int b200sd_unpack_latent(const void* moments, long long pitch, float* x, int B, int HW, float scale, int dtype,
                         void* stream);
/* hires fix, "Latent" upscaler: fp32 NHWC latents [B,H*W,4] -> [B,Ho*Wo,4], bilinear, half-pixel centres, no antialias
 * (torch.nn.functional.interpolate(mode="bilinear") in sdwui StableDiffusionProcessingTxt2Img.sample_hr_pass) */
int b200sd_resize_latent_bilinear(const float* x, float* y, int B, int H, int W, int Ho, int Wo, void* stream);

/* inpainting: x[b,p,:] = x[b,p,:] * latmask[p] + init[b,p,:] * (1 - latmask[p]) on fp32 NHWC latents [B,HW,4]; latmask [HW]
 * is the request's latent-resolution mask (1 = repaint).  (sdwui CFGDenoiser.apply_blend, before every model call of the
 * timestep samplers, and once more after sampling) */
int b200sd_blend_latent(float* x, const float* init, const float* latmask, int B, int HW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SD_H_ */
