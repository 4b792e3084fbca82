/*
 * rgm_b200 -- C ABI of the B200-native (sm_100a) rule-guided sampling hot path.
 *
 * The reference (yjhuangcd/rule-guided-music) is pure Python/PyTorch and has no FFI of its own; its drop-in boundary
 * for this path is four Python callables: the denoiser `model(x, t, **kw)`, the latent decoder
 * `embed_model.decode(z)`, the rule programs `FUNC_DICT[name](roll)` / `LOSS_DICT[name]`, and the sampler methods that
 * call them.  Every entry point below names the reference callable(s) it replaces (file:line under the reference
 * repository).  The Python mirror of the reference API (rule_guided_music_b200/) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*); compute calls do not synchronise
 *     (a handle's workspace grows on first use of a larger batch: one cudaMalloc, nothing is freed until *_destroy;
 *     rgm_*_reserve pre-sizes)
 *   - return 0 on success, negative on error; rgm_last_error() returns the message for the calling thread
 *   - handles own their packed weights and workspace; one handle is used by one host thread at a time
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails with an error
 */
#ifndef RGM_B200_H_
#define RGM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------------ */
const char* rgm_last_error(void);
int rgm_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
unsigned long long rgm_launch_count(void);
/* 0 when the current device is sm_100 (B200); negative with an error message otherwise */
int rgm_check_device(void);
/* Optional per-launch device timer used by bench.py for the roofline figures: while enabled, every kernel launch of the
 * library is bracketed by CUDA events on its stream.  rgm_prof_enable(1) clears and starts, (0) stops.
 * rgm_prof_summary synchronises the device and writes one line per kernel family into a HOST buffer:
 * name \t launches \t total_ms \t flops_algorithmic \t flops_executed \t bytes_algorithmic */
int rgm_prof_enable(int on);
int rgm_prof_summary(char* buf_host, int cap);

/* ---- denoiser: DiTRotary.forward (guided_diffusion/dit.py:538-634; registry DiT_models dit.py:969-983) -------- */
typedef struct rgm_dit rgm_dit;
/* DiTRotary.__init__ (dit.py:545-576): label_rows = num_classes + 1 (LabelEmbedder table, dit.py:79-80) or 0 when the
 * model has no label embedder; latent_w = input_size[1] (16); mlp_hidden = int(hidden * mlp_ratio) */
int rgm_dit_create(rgm_dit** out, int depth, int hidden, int heads, int patch, int in_channels, int out_channels,
                   int label_rows, int latent_w, int mlp_hidden);
/* 2 (default) = consecutive sample chunks alternate between two workspaces / streams so one chunk's attention and
 * LayerNorm passes overlap the other's GEMMs; 1 = serial (per-launch profiling). */
int rgm_dit_set_lanes(rgm_dit* h, int lanes);
int rgm_dit_destroy(rgm_dit* h);
/* Pre-size the handle's workspaces and rotary table for batches of up to B samples of latent height H (what the first
 * rgm_dit_forward of that size would allocate).  Buffers only grow and a grown buffer's predecessor stays allocated
 * until rgm_dit_destroy, so a CUDA graph captured earlier keeps replaying into valid memory; growth while the caller's
 * stream is being captured is refused with an error instead -- reserve (or run the shape once eagerly) before capture. */
int rgm_dit_reserve(rgm_dit* h, int B, int H);
/* model.load_state_dict(sd, strict=False) (scripts/sample_rule.py:71-73), one tensor per call: `key` is the reference
 * state-dict key, `src` the fp32 tensor on the device.  Returns 0 = stored (converted to the kernel layout),
 * 1 = key not part of this path (ignored), negative = error (element count mismatch).
 * Extra key "__timestep_freqs": the 128 sinusoid frequencies of TimestepEmbedder.timestep_embedding (dit.py:57-59). */
int rgm_dit_load(rgm_dit* h, const char* key, const float* src, long long numel, void* stream);
/* model(x, t, y) (dit.py:618-634): x f32 [B, C, H, latent_w]; t f32 [B] (already mapped / rescaled by the caller,
 * respace.py:123-128); y int64 [B] label rows or NULL (dit.py:627-629); out f32 [B, C_out, H, latent_w].
 * H * latent_w / patch must be 64, 128, 192 or 256 tokens. */
int rgm_dit_forward(rgm_dit* h, const float* x, const float* t, const long long* y, float* out, int B, int H,
                    void* stream);

/* ---- latent decoder: AutoencoderKL.decode (taming/models/klvae_pedal.py:80-85) -------------------------------- */
typedef struct rgm_vae rgm_vae;
/* Decoder.__init__ (taming/modules/diffusionmodules/model.py:436-504) with attn_resolutions = [] (mid attention only) */
int rgm_vae_create(rgm_vae** out, int ch, const int* ch_mult_host, int n_levels, int num_res_blocks, int z_channels,
                   int out_ch);
int rgm_vae_destroy(rgm_vae* h);
/* Diagnostic of the convolutions that normalise their own output (GroupNorm statistics exchanged between CTAs while the
 * accumulators wait in tensor memory): 0 in a healthy run; 1 if a wait for an image's other tiles ever gave up; 2 if a
 * partial sum left the range of the fixed-point statistics (a GroupNorm group with an rms above ~250 over a whole image:
 * create the handle with RGM_GN_EPI=0 for such weights); negative on error.  Synchronises the device.  After a non-zero
 * value the handle's results are invalid and further decode / encode calls fail. */
int rgm_vae_gn_timeouts(rgm_vae* h);
/* Pre-size the activation buffers for decodes / encodes of up to n_tiles 16x16 latent tiles (same contract as
 * rgm_dit_reserve). */
int rgm_vae_reserve(rgm_vae* h, int n_tiles);
/* 2 (default): consecutive tile chunks alternate between two internal streams so GroupNorm passes overlap convolutions;
 * 1: strictly serial on the caller's stream (used for per-kernel timing) */
int rgm_vae_set_lanes(rgm_vae* h, int lanes);
/* AutoencoderKL.init_from_ckpt (klvae_pedal.py:50-59), one tensor per call; keys "post_quant_conv.*", "decoder.*".
 * Same return convention as rgm_dit_load. */
int rgm_vae_load(rgm_vae* h, const char* key, const float* src, long long numel, void* stream);
/* AutoencoderKL.encode_save(x, range_fix=False) (klvae_pedal.py:60-68): Encoder (model.py:342-433) + quant_conv.
 * x f32 NCHW [n, 3, 128, 128] (piano-roll tiles in [-1, 1]) -> moments f32 NCHW [n, 2*z_channels, 16, 16]
 * (mean | log-variance).  Weights: the checkpoint's encoder.* and quant_conv.* keys through rgm_vae_load. */
int rgm_vae_encode(rgm_vae* h, const float* x, float* moments, int n, void* stream);

/* _decode(pred_zstart, embed_model, scale_factor) (guided_diffusion/gaussian_diffusion.py:1347-1358):
 * lat f32 [n_cand, 4, Hlat, 16] -> roll f32 [n_cand, roll_ch, 128, 8*Hlat], roll_ch in [1, out_ch] (the rules read
 * channel 0 only, music_rules.py:31,56).  embed_model.decode(z [n,4,16,16]) is the case Hlat = 16, scale_factor = 1
 * with z transposed to [n,4,time,pitch]. */
int rgm_vae_decode_latents(rgm_vae* h, const float* lat, float scale_factor, float* roll, int n_cand, int Hlat,
                           int roll_ch, void* stream);

/* ---- rule programs (music_rule_guidance/music_rules.py, rule_maps.py) ----------------------------------------- */
/* FUNC_DICT["pitch_hist"] = total_pitch_class_histogram (music_rules.py:29-43): roll f32 [n, ch, 128, L] (channel 0
 * is read AND piano-masked in place, music_rules.py:23-26) -> hist f32 [n, 12] */
int rgm_rule_pitch_hist(float* roll, float* hist, int n, int ch, int L, void* stream);
/* FUNC_DICT["note_density" | "note_density_hr_*" | "note_density_pixel"] = note_density(interval, horizontal_scale)
 * (music_rules.py:46-83; quantize_factor = 1): out f32 [n, 2*L/interval]; channel 0 is thresholded in place */
int rgm_rule_note_density(float* roll, float* out, int n, int ch, int L, int interval, float horizontal_scale,
                          void* stream);
/* scg_sample's per-rule update `total_log_prob += -LOSS_DICT[rule](gen, target.repeat(N,1)) * weight`
 * (gaussian_diffusion.py:532-538): gen f32 [n, K], target f32 [B, K] (row i uses target i % B), total f32 [n];
 * kind 0 = mse_loss_mean (rule_maps.py:17-18), 1 = zero_one_loss_mean (rule_maps.py:21-22) */
int rgm_rule_loss_accum(const float* gen, const float* target, float* total, int n, int B, int K, int kind,
                        float weight, void* stream);

/* ---- stochastic control guidance (guided_diffusion/gaussian_diffusion.py:491-554) ------------------------------ */
/* :510-514  cand[n, b, :] = mean[b, :] + g[b] * noise[n, b, :]   (elems = C*H*W per sample) */
int rgm_scg_fanout(const float* mean, const float* g, const float* noise, float* cand, int N, int B, long long elems,
                   void* stream);
/* :359-364  x0 = a[b]*x - c[b]*eps (a = sqrt_recip_alphas_cumprod[t], c = sqrt_recipm1_alphas_cumprod[t]); clamp: +-1 */
int rgm_x0_from_eps(const float* x, const float* eps, const float* a, const float* c, float* x0, int B,
                    long long elems, int clamp, void* stream);
/* :539-554  max_ind = total.view(N, B).argmax(0) (first maximal index); out[b] = cand[max_ind[b], b]; idx int64 [B] */
int rgm_scg_select(const float* total, const float* cand, float* out, long long* idx, int N, int B, long long elems,
                   void* stream);

/* ---- one whole SCG step from a C host --------------------------------------------------- */
/* GaussianDiffusion.__init__ (gaussian_diffusion.py:142-186) for the (respaced) betas of SpacedDiffusion
 * (respace.py:63-95): float64 on the host in numpy's operation order, cast to fp32 once (what _extract_into_tensor
 * :1331-1344 does per lookup), written to out_device as RGM_COEF_ROWS rows of T floats, row r at out_device + r*T. */
enum {
  RGM_COEF_BETAS = 0,
  RGM_COEF_ALPHAS_CUMPROD,
  RGM_COEF_ALPHAS_CUMPROD_PREV,
  RGM_COEF_SQRT_ALPHAS_CUMPROD,
  RGM_COEF_SQRT_ONE_MINUS_ALPHAS_CUMPROD,
  RGM_COEF_SQRT_RECIP_ALPHAS_CUMPROD,
  RGM_COEF_SQRT_RECIPM1_ALPHAS_CUMPROD,
  RGM_COEF_POSTERIOR_VARIANCE,
  RGM_COEF_POSTERIOR_LOG_VARIANCE_CLIPPED,
  RGM_COEF_POSTERIOR_MEAN_COEF1,
  RGM_COEF_POSTERIOR_MEAN_COEF2,
  RGM_COEF_FIXED_LARGE_VARIANCE,     /* append(posterior_variance[1], betas[1:])   (:316-329) */
  RGM_COEF_FIXED_LARGE_LOG_VARIANCE,
  RGM_COEF_LOG_BETAS,
  RGM_COEF_ROWS
};
int rgm_coeff_tables(const double* betas_host, int T, float* out_device, void* stream);
/* the same tables into a HOST buffer (no device needed; what the CPU-side tests pin against numpy) */
int rgm_coeff_tables_host(const double* betas_host, int T, float* out_host);
/* ddim_sample between the denoiser call and the noise (:921-944): pred_xstart = clip(sqrt_recip*x - sqrt_recipm1*eps),
 * eps re-derived from it, sigma[b] = eta*sqrt((1-abar_prev)/(1-abar))*sqrt(1-abar/abar_prev),
 * mean_pred = pred_xstart*sqrt(abar_prev) + sqrt(1-abar_prev-sigma^2)*eps.  t_index int64 [B] indexes the tables. */
int rgm_ddim_mean(const float* x, const float* eps, const float* coeff_tables, int T, const long long* t_index,
                  float eta, int clip_denoised, float* pred_xstart, float* mean_pred, float* sigma, int B,
                  long long elems, void* stream);

/* One rule of model_kwargs["rule"] with its scg_kwargs weight (FUNC_DICT / LOSS_DICT entries, rule_maps.py:5-38). */
enum { RGM_RULE_PITCH_HIST = 0, RGM_RULE_NOTE_DENSITY = 1, RGM_RULE_NOTE_DENSITY_CLASS = 2 };
typedef struct rgm_rule_spec {
  int kind;               /* RGM_RULE_* */
  int interval;           /* note density: window in roll columns (128; 16 for note_density_pixel) */
  float horizontal_scale; /* note density: 5 (default), 1 / 2 for the _hr_ variants, 1 for the class variant */
  int loss_kind;          /* 0 = mse_loss_mean, 1 = zero_one_loss_mean */
  float weight;           /* scg_kwargs.get(rule_name, 1.) */
  const float* target;    /* DEVICE f32 [B, K]: K = 12 (pitch_hist) or 2 * L / interval (class targets as floats) */
} rgm_rule_spec;

/* scg_sample (gaussian_diffusion.py:491-554) in one call.  The handle owns the candidate / roll scratch (grown on first
 * use or by rgm_scg_reserve; same no-free contract as rgm_dit_reserve).  vae may be NULL (embed_model=None, :523). */
typedef struct rgm_scg rgm_scg;
int rgm_scg_create(rgm_scg** out, rgm_dit* dit, rgm_vae* vae);
int rgm_scg_reserve(rgm_scg* h, int N, int B, int C, int H, int W);
int rgm_scg_destroy(rgm_scg* h);
/* mean f32 [B,C,H,W] (mean_pred / p_mean_var["mean"]), g f32 [B] (sigma or exp(0.5*log_variance)), noise f32
 * [N,B,C,H,W] (th.randn_like, :512), t_model f32 [B] = the timestep the denoiser is called with (already mapped by
 * respace.py:123-128), y int64 [B] or NULL, x0_a / x0_c f32 [B] = sqrt_recip_alphas_cumprod[t] /
 * sqrt_recipm1_alphas_cumprod[t] (:359-364), rules_host = n_rules specs in dict order (the rule programs write through
 * the roll, so order matters).  out_sample f32 [B,C,H,W] = the chosen x_(t-1), out_index int64 [B] = max_ind,
 * out_scores f32 [N*B] = total_log_prob (candidate-major) or NULL. */
int rgm_scg_step(rgm_scg* h, const float* mean, const float* g, const float* noise, const float* t_model,
                 const long long* y, const float* x0_a, const float* x0_c, float scale_factor,
                 const rgm_rule_spec* rules_host, int n_rules, int N, int B, int C, int H, int W, float* out_sample,
                 long long* out_index, float* out_scores, void* stream);

/* ---- building blocks (exposed for the parity tests) ---------------------------------------------------------- */
/* out32[M,N] = A16[M,K] . B16[N,K]^T + bias[N]      (torch.nn.functional.linear; reference dit.py:256,286,324-326)
 * block_n: 0 = choose, else 32 / 128 / 256 */
int rgm_gemm_f16(const void* a16, const void* b16, const float* bias, float* out32, int M, int N, int K, int block_n,
                 void* stream);
/* NHWC fp16 convolution with fp32 accumulation           (torch.nn.Conv2d; reference model.py:38-53,78-137)
 * kind: 0 = 1x1, 1 = 3x3 pad 1, 2 = nearest-2x upsample + 3x3 pad 1, 3 = Downsample: pad (0,1,0,1) + 3x3 stride 2
 * (reference model.py:55-75)  (weights packed by rgm_pack_conv_weight)
 * x16 [n,H,W,Cin], out16 [n,H',W',Cout], optional resid16 like out16; gn_part may be NULL */
int rgm_conv_f16(const void* x16, const void* w16_packed, const float* bias, const void* resid16, void* out16,
                 int n_img, int H, int W, int Cin, int Cout, int kind, int block_n, float* gn_part, void* stream);
/* GroupNorm apply (+ swish) as its own pass: y16 = swish(a*x16 + b) with ab = (a, b) f32 pairs per (image, channel)
 * (reference model.py:34-35, 29-31, the `h = nonlinearity(norm(x))` of ResnetBlock :119-120) */
int rgm_gn_apply_f16(const void* x16, const float* ab, void* y16, int n_img, int HW, int C, int swish, void* stream);
/* conv3x3(swish(GroupNorm(x))) in ONE kernel (reference model.py:117-137): x16_raw is the un-normalised NHWC tensor
 * [n, H, 128, Cin], ab_in its GroupNorm affine [n][Cin] (a, b); weights / bias / residual / output / gn_part as in
 * rgm_conv_f16 with kind 1.  W must be 128, H even, Cout 128. */
int rgm_conv_gn_f16(const void* x16_raw, const float* ab_in, const void* w16_packed, const float* bias,
                    const void* resid16, void* out16, int n_img, int H, int W, int Cin, int Cout, float* gn_part,
                    void* stream);
/* swish(GroupNorm_32(conv(x))) in ONE kernel with the normalisation applied to the convolution's OWN output
 * (reference model.py:119-124: h = conv1(..); h = norm2(h); h = nonlinearity(h)): the CTAs of an image exchange their
 * partial statistics while the accumulators wait in tensor memory, so the raw output is never written.  kind 0 / 1
 * (kind 2, the upsample conv with W a multiple of 32, in the dual form only),
 * Cout in {128, 256, 512}, H*W a multiple of 256; gamma / beta f32 [Cout]; gn_scratch: n * 512 bytes of device
 * scratch (per-group statistics accumulators that also count arrivals, zeroed by the call); gn_err: device int that is set to 1
 * if a wait gives up (never in a healthy run).  swish = 0 applies the norm only.
 * raw16 == NULL: only out16 = swish(norm(conv(x))) is written (the accumulators wait in tensor memory for the statistics).
 * raw16 != NULL ("dual" form, model.py:117-137 where a block's output feeds both the next shortcut and the next norm1):
 * raw16 = conv(x) + resid16 (resid16 may be NULL, or alias raw16) and out16 = swish(norm(raw16)), the normalised copy
 * written one tile later from the warp's own L2-resident rows. */
int rgm_conv_norm_f16(const void* x16, const void* w16_packed, const float* bias, const float* gamma, const float* beta,
                      const void* resid16, void* raw16, void* out16, int n_img, int H, int W, int Cin, int Cout, int kind,
                      int swish, void* gn_scratch, int* gn_err, void* stream);
/* weight fp32 [Cout,Cin,kh,kw] (torch layout) -> packed fp16 rows for rgm_conv_f16; cin_pad >= Cin (multiple of 64),
 * cout_pad >= Cout. Output size: kind 0: cout_pad*cin_pad; kind 1, 3: cout_pad*9*cin_pad; kind 2: 4*cout_pad*4*cin_pad */
int rgm_pack_conv_weight(const float* w32, void* w16_packed, int Cout, int Cin, int cout_pad, int cin_pad, int kind,
                         void* stream);
/* softmax(q k^T * scale) v per (sample, head) (dit.py:274-277): q,k fp16 [B,heads,T,dh], vt fp16 [B,heads,dh,T],
 * out fp16 [B*T, heads*dh]; T in {64, 128, 192, 256} */
int rgm_attention_f16(const void* q16, const void* k16, const void* vt16, void* out16, int B, int heads, int T, int dh,
                      float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGM_B200_H_ */
