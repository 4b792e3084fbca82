// Launchers of the memory-bound helper kernels around the tensor-core GEMMs (aux_kernels.cu, attention.cu,
// rules.cu).  All take the caller's stream and return the launch status; none synchronises.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>

namespace rgm {

// ---- DiT (reference guided_diffusion/dit.py) ------------------------------------------------------------------
// FlattenPatchify1D gather (dit.py:219-224): x f32 [B,C,H,W] -> tok fp16 [B*T, kpad], T = H*W/P,
// token j = (time j / tpt, pitch part j % tpt), feature = pitch_local*C + channel; features >= C*P are zero.
cudaError_t launch_patchify(const float* x, __half* tok, int B, int C, int H, int W, int P, int kpad,
                            cudaStream_t s);
// TimestepEmbedder.timestep_embedding (dit.py:47-65): emb fp16 [B,2*half] = [cos(t f) | sin(t f)]
cudaError_t launch_timestep_embedding(const float* t, const float* freqs, __half* emb, int B, int half, int ld,
                                      cudaStream_t s);
// rotary tables (rotary_embedding_torch restated): (cos, sin) [T, nfreq] of pos * freq
cudaError_t launch_rope_table(const float* freqs, float2* cs, int T, int nfreq, cudaStream_t s);
// LayerNorm(no affine, eps) + modulate (dit.py:25-26, 321-335): out16[row] = LN(x[row])*(1+scale[b]) + shift[b],
// b = row / rows_per_sample; shift/scale rows have stride mod_ld
cudaError_t launch_ln_modulate(const float* x, const float* shift, const float* scale, int mod_ld, __half* out,
                               long long rows, int D, int rows_per_sample, float eps, cudaStream_t s);
// multi-head attention on tcgen05 (dit.py:263-288): q,k fp16 [B,heads,T,dh] (rotary already applied), vt fp16
// [B,heads,dh,T]; out fp16 [B*T, heads*dh].  T in {128, 256}, dh % 8 == 0, dh <= 128.
cudaError_t launch_attention(const __half* q, const __half* k, const __half* vt, __half* out, int B, int heads, int T,
                             int dh, float scale, cudaStream_t s, std::string* err);

// ---- VAE decoder (reference taming/modules/diffusionmodules/model.py) -----------------------------------------
// _decode re-tiling (gaussian_diffusion.py:1350-1353) + post_quant_conv 1x1 (klvae_pedal.py:81) + conv_in 3x3
// (model.py:514), fp32 math: latents f32 [n_cand,4,Hlat,16] -> out fp16 NHWC [n_tiles,16,16,Cout] for tiles
// tile0 .. tile0+n_tiles-1 in tile-major order g = k*n_cand + cand.
cudaError_t launch_vae_stem(const float* lat, float scale, const float* pq_w, const float* pq_b,
                            const float* cin_w, const float* cin_b, __half* out, int n_cand, int Hlat, int tile0,
                            int n_tiles, int Cout, cudaStream_t s);
// GroupNorm(32 groups, eps) statistics -> per (image, channel) affine (a, b) with GN(x) = a*x + b
//   from the tensor itself (x fp16 NHWC [n, HW, C])
cudaError_t launch_gn_stats(const __half* x, const float* gamma, const float* beta, float2* ab, int n, int HW, int C,
                            float eps, cudaStream_t s);
//   from the partial sums a conv epilogue wrote (gemm_tc.cuh gn_part): slots_per_img 128-row slots per image and
//   parity, n_par parities (4 for the upsample conv), slot stride between parities = par_stride slots
cudaError_t launch_gn_finalize(const float* part, const float* gamma, const float* beta, float2* ab, int n,
                               int slots_per_img, int n_par, long long par_stride, int C, int HW_out, float eps,
                               cudaStream_t s);
// y = swish(a*x + b) (or a*x + b when swish == 0), fp16 NHWC
cudaError_t launch_gn_apply(const __half* x, const float2* ab, __half* y, int n, int HW, int C, int swish,
                            cudaStream_t s);
// norm_out + swish + conv_out (3x3, C -> out_ch <= 4, fp32 weights [out_ch,C,3,3]) + roll assembly, CUDA cores:
// x fp16 NHWC [n,128,128,C] (raw, before GroupNorm), ab its GroupNorm affine, roll f32 [n_cand, roll_ch, 128, roll_len]
// bfrag = the conv_out weights packed by launch_vae_out_pack (mma.m16n8k16 B fragments, 9 * C/16 * 32 uint2 in all:
// the channel-0 packing followed by the three-channel packing); out_ch must be 3
cudaError_t launch_vae_out(const __half* x, const float2* ab, const void* bfrag, const float* bias, float* roll, int n,
                           int C, int out_ch, int tile0, int n_cand, int roll_len, int roll_ch, cudaStream_t s);
cudaError_t launch_vae_out_pack(const float* w, void* bfrag, int C, int out_ch, cudaStream_t s);
// ---- VAE encoder (reference model.py:342-433, klvae_pedal.py:60-68) ---------------------------------------------
// Encoder.conv_in 3x3 (model.py:355-359, 3 -> Cout channels), fp32 math: x f32 NCHW [n, Cin<=4, 128, 128] ->
// out fp16 NHWC [n, 128, 128, Cout]
cudaError_t launch_vae_enc_stem(const float* x, const float* w, const float* b, __half* out, int n, int Cin, int Cout,
                                cudaStream_t s);
// quant_conv 1x1 (klvae_pedal.py:62): h f32 NHWC [n, HW, ld] (first C channels) -> moments f32 NCHW [n, C, HW]
cudaError_t launch_vae_quant(const float* h, int ld, const float* w, const float* b, float* moments, int n, int HW,
                             int C, cudaStream_t s);
// row softmax fp32 [rows, cols] -> fp16 (AttnBlock, model.py:183)
cudaError_t launch_softmax_rows(const float* x, __half* y, long long rows, int cols, cudaStream_t s);
// batched transpose fp16 [n, R, C] -> [n, C, R]
cudaError_t launch_transpose(const __half* x, __half* y, int n, int R, int C, cudaStream_t s);

// ---- rules and SCG selection (reference music_rule_guidance/music_rules.py, gaussian_diffusion.py:531-554) -----
// total_pitch_class_histogram (music_rules.py:29-43) on channel 0 of roll f32 [n, ch, 128, L]; writes the piano mask
// through to the roll like the reference; hist f32 [n,12]
cudaError_t launch_rule_pitch_hist(float* roll, float* hist, int n, int ch, int L, cudaStream_t s);
// note_density (music_rules.py:46-83): out f32 [n, 2*L/interval] = [vertical | horizontal]; thresholds channel 0 of
// the roll in place like the reference
cudaError_t launch_rule_note_density(float* roll, float* out, int n, int ch, int L, int interval, float hscale,
                                     cudaStream_t s);
// total[n] += weight * -(loss(gen[n,:], target[n % B,:]));  kind 0: mean squared error, 1: mean (gen != target)
cudaError_t launch_rule_loss_accum(const float* gen, const float* target, float* total, int n, int B, int K, int kind,
                                   float weight, cudaStream_t s);
// first-max argmax over N candidates per sample and gather of the winner (gaussian_diffusion.py:539-554)
cudaError_t launch_scg_select(const float* total, const float* cand, float* out, long long* idx, int N, int B,
                              long long elems, cudaStream_t s);

// ---- sampler elementwise (reference guided_diffusion/gaussian_diffusion.py) ------------------------------------
// candidates[n,b,:] = mean[b,:] + g[b]*noise[n,b,:]   (:510-514; g is per-sample because t is per-sample)
cudaError_t launch_scg_fanout(const float* mean, const float* g, const float* noise, float* cand, int N, int B,
                              long long elems, cudaStream_t s);
// x0 = a[b]*x - c[b]*eps  (:359-364), optional clamp to [-1,1]
cudaError_t launch_x0_from_eps(const float* x, const float* eps, const float* a, const float* c, float* x0, int B,
                               long long elems, int clamp, cudaStream_t s);

unsigned long long aux_launch_count();
unsigned long long attention_launch_count();
unsigned long long rules_launch_count();

}  // namespace rgm
