// C ABI: one stochastic-control-guidance step for a host that is not Python (include/rgm_b200.h).
//
//   rgm_coeff_tables   GaussianDiffusion.__init__ tables (guided_diffusion/gaussian_diffusion.py:142-186) in float64 on
//                      the host exactly as numpy computes them, cast to fp32 once, resident on the device
//   rgm_ddim_mean      the elementwise part of ddim_sample between the denoiser and the noise (:921-944)
//   rgm_scg_step       scg_sample (:491-554): fan-out, ONE denoiser call over N*B candidates, x0, fused _decode,
//                      rule programs in the caller's order, weighted -loss, first-max argmax, gather
//
// rgm_scg_step chains exactly the kernels the Python mirror launches (gaussian_diffusion.py of this package), so its
// result is bit-identical to the Python-orchestrated step (tests/test_capi_gpu.py).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rgm_b200.h"
#include "api_util.h"
#include "aux_kernels.h"

namespace rgm {

namespace {

struct Scg {
  rgm_dit* dit = nullptr;
  rgm_vae* vae = nullptr;
  GrowBuf cand, eps, x0, roll, rep, gen, total;
};

// out[i] = src[i % B] for the per-candidate copies of per-sample scalars (t.repeat(N), y.repeat(N), coefficient gathers)
__global__ void repeat_kernel(const float* __restrict__ t, const long long* __restrict__ y, const float* __restrict__ a,
                              const float* __restrict__ c, float* __restrict__ t_rep, long long* __restrict__ y_rep,
                              float* __restrict__ a_rep, float* __restrict__ c_rep, int B, int NB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB) return;
  const int b = i % B;
  t_rep[i] = t[b];
  if (y) y_rep[i] = y[b];
  a_rep[i] = a[b];
  c_rep[i] = c[b];
}

// torch.bucketize(x, bounds) (right = False): number of bounds strictly below x; written as fp32 class indices
// (music_rules.py:86-94; the first half of a row uses the vertical bounds, the second half the horizontal ones)
__constant__ float c_vt_bounds[7] = {1.29f, 2.7578125f, 3.61f, 4.4921875f, 5.28125f, 6.1171875f, 7.22f};
__constant__ float c_hr_bounds[7] = {1.8f, 2.6f, 3.2f, 3.6f, 4.4f, 4.8f, 5.8f};

__global__ void nd_classes_kernel(float* __restrict__ nd, long long total, int K, float hscale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const bool vertical = (i % K) < K / 2;
  const float v = nd[i];
  int cls = 0;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float bound = vertical ? c_vt_bounds[j] : __fdiv_rn(c_hr_bounds[j], hscale);
    cls += (bound < v) ? 1 : 0;
  }
  nd[i] = (float)cls;
}

// ddim_sample's elementwise block (:921-944), one rounding per torch op (no contraction), coefficients per sample
__global__ void ddim_mean_kernel(const float* __restrict__ x, const float* __restrict__ eps_model,
                                 const float* __restrict__ tab, int T, const long long* __restrict__ t, float eta,
                                 int clip, float* __restrict__ pred_xstart, float* __restrict__ mean_pred,
                                 float* __restrict__ sigma_out, long long total, long long elems) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / elems);
    const long long ti = t[b];
    const float ab = tab[RGM_COEF_ALPHAS_CUMPROD * T + ti], abp = tab[RGM_COEF_ALPHAS_CUMPROD_PREV * T + ti];
    const float sr = tab[RGM_COEF_SQRT_RECIP_ALPHAS_CUMPROD * T + ti];
    const float srm1 = tab[RGM_COEF_SQRT_RECIPM1_ALPHAS_CUMPROD * T + ti];
    float x0 = __fsub_rn(__fmul_rn(sr, x[i]), __fmul_rn(srm1, eps_model[i]));       // _predict_xstart_from_eps
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float e = __fdiv_rn(__fsub_rn(__fmul_rn(sr, x[i]), x0), srm1);           // _predict_eps_from_xstart
    // sigma = eta * sqrt((1 - abp) / (1 - ab)) * sqrt(1 - ab / abp)
    const float s1 = __fsqrt_rn(__fdiv_rn(__fsub_rn(1.f, abp), __fsub_rn(1.f, ab)));
    const float s2 = __fsqrt_rn(__fsub_rn(1.f, __fdiv_rn(ab, abp)));
    const float sigma = __fmul_rn(__fmul_rn(eta, s1), s2);
    // mean_pred = x0 * sqrt(abp) + sqrt(1 - abp - sigma^2) * eps
    const float k = __fsqrt_rn(__fsub_rn(__fsub_rn(1.f, abp), __fmul_rn(sigma, sigma)));
    mean_pred[i] = __fadd_rn(__fmul_rn(x0, __fsqrt_rn(abp)), __fmul_rn(k, e));
    pred_xstart[i] = x0;
    if (i == (long long)b * elems) sigma_out[b] = sigma;
  }
}

}  // namespace
}  // namespace rgm

using namespace rgm;

extern "C" {

int rgm_coeff_tables_host(const double* betas_host, int T, float* out_host) {
  if (!betas_host || !out_host || T < 2) return set_error("rgm_coeff_tables: bad arguments (need T >= 2 betas)");
  for (int i = 0; i < T; ++i)
    if (!(betas_host[i] > 0.0 && betas_host[i] <= 1.0)) return set_error("rgm_coeff_tables: betas must be in (0, 1]");
  // float64, the operation order of gaussian_diffusion.py:152-186 (np.cumprod is a sequential product)
  std::vector<double> ac(T), acp(T), pv(T);
  double run = 1.0;
  for (int i = 0; i < T; ++i) {
    run *= (1.0 - betas_host[i]);
    ac[i] = run;
    acp[i] = i == 0 ? 1.0 : ac[i - 1];
  }
  auto row = [&](int r) { return out_host + (size_t)r * T; };
  for (int i = 0; i < T; ++i) {
    const double b = betas_host[i];
    pv[i] = b * (1.0 - acp[i]) / (1.0 - ac[i]);
    row(RGM_COEF_BETAS)[i] = (float)b;
    row(RGM_COEF_ALPHAS_CUMPROD)[i] = (float)ac[i];
    row(RGM_COEF_ALPHAS_CUMPROD_PREV)[i] = (float)acp[i];
    row(RGM_COEF_SQRT_ALPHAS_CUMPROD)[i] = (float)std::sqrt(ac[i]);
    row(RGM_COEF_SQRT_ONE_MINUS_ALPHAS_CUMPROD)[i] = (float)std::sqrt(1.0 - ac[i]);
    row(RGM_COEF_SQRT_RECIP_ALPHAS_CUMPROD)[i] = (float)std::sqrt(1.0 / ac[i]);
    row(RGM_COEF_SQRT_RECIPM1_ALPHAS_CUMPROD)[i] = (float)std::sqrt(1.0 / ac[i] - 1.0);
    row(RGM_COEF_POSTERIOR_VARIANCE)[i] = (float)pv[i];
    row(RGM_COEF_POSTERIOR_MEAN_COEF1)[i] = (float)(b * std::sqrt(acp[i]) / (1.0 - ac[i]));
    row(RGM_COEF_POSTERIOR_MEAN_COEF2)[i] = (float)((1.0 - acp[i]) * std::sqrt(1.0 - b) / (1.0 - ac[i]));
    row(RGM_COEF_LOG_BETAS)[i] = (float)std::log(b);
  }
  for (int i = 0; i < T; ++i) {
    // posterior_log_variance_clipped = log(append(pv[1], pv[1:])); FIXED_LARGE = append(pv[1], betas[1:]) (:316-329)
    const double pvc = i == 0 ? pv[1] : pv[i];
    const double fl = i == 0 ? pv[1] : betas_host[i];
    row(RGM_COEF_POSTERIOR_LOG_VARIANCE_CLIPPED)[i] = (float)std::log(pvc);
    row(RGM_COEF_FIXED_LARGE_VARIANCE)[i] = (float)fl;
    row(RGM_COEF_FIXED_LARGE_LOG_VARIANCE)[i] = (float)std::log(fl);
  }
  return 0;
}

int rgm_coeff_tables(const double* betas_host, int T, float* out_device, void* stream) {
  if (!out_device || T < 2) return set_error("rgm_coeff_tables: bad arguments (need T >= 2 betas)");
  std::vector<float> out((size_t)RGM_COEF_ROWS * T);
  if (rgm_coeff_tables_host(betas_host, T, out.data()) != 0) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the staging vector dies with this call: a pageable cudaMemcpyAsync returns after the source has been consumed
  return check_cuda(cudaMemcpyAsync(out_device, out.data(), out.size() * sizeof(float), cudaMemcpyHostToDevice, st),
                    "rgm_coeff_tables");
}

int rgm_ddim_mean(const float* x, const float* eps, const float* coeff_tables, int T, const long long* t_index,
                  float eta, int clip_denoised, float* pred_xstart, float* mean_pred, float* sigma, int B,
                  long long elems, void* stream) {
  if (rgm_check_device()) return -1;
  if (!x || !eps || !coeff_tables || !t_index || !pred_xstart || !mean_pred || !sigma || B <= 0 || elems <= 0)
    return set_error("rgm_ddim_mean: bad arguments");
  const long long total = (long long)B * elems;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  ddim_mean_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, coeff_tables, T, t_index, eta,
                                                                         clip_denoised, pred_xstart, mean_pred, sigma,
                                                                         total, elems);
  g_aux_launches++;
  return check_cuda(cudaGetLastError(), "rgm_ddim_mean");
}

int rgm_scg_create(rgm_scg** out, rgm_dit* dit, rgm_vae* vae) {
  if (rgm_check_device()) return -1;
  if (!out || !dit) return set_error("rgm_scg_create: null argument (the decoder may be NULL, the denoiser may not)");
  Scg* s = new Scg();
  s->dit = dit;
  s->vae = vae;
  *out = reinterpret_cast<rgm_scg*>(s);
  return 0;
}

int rgm_scg_destroy(rgm_scg* h) {
  if (h) {
    cudaDeviceSynchronize();
    delete reinterpret_cast<Scg*>(h);
  }
  return 0;
}

static int scg_reserve(Scg* s, int N, int B, int C, int H, int W, int kmax, cudaStream_t st) {
  const size_t NB = (size_t)N * B, elems = (size_t)C * H * W;
  RGM_CUDA_OK(s->cand.reserve(NB * elems * sizeof(float), st));
  RGM_CUDA_OK(s->eps.reserve(NB * elems * sizeof(float), st));
  RGM_CUDA_OK(s->x0.reserve(NB * elems * sizeof(float), st));
  if (s->vae) RGM_CUDA_OK(s->roll.reserve(NB * 128 * (size_t)(8 * H) * sizeof(float), st));
  RGM_CUDA_OK(s->rep.reserve(NB * (3 * sizeof(float) + sizeof(long long)) + 64, st));
  RGM_CUDA_OK(s->gen.reserve(NB * (size_t)kmax * sizeof(float), st));
  RGM_CUDA_OK(s->total.reserve(NB * sizeof(float), st));
  return 0;
}

int rgm_scg_reserve(rgm_scg* h, int N, int B, int C, int H, int W) {
  if (!h || N <= 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error("rgm_scg_reserve: bad arguments");
  Scg* s = reinterpret_cast<Scg*>(h);
  if (scg_reserve(s, N, B, C, H, W, 2 * (8 * H), nullptr) != 0) return -1;
  if (rgm_dit_reserve(s->dit, N * B, H) != 0) return -1;
  if (s->vae && rgm_vae_reserve(s->vae, N * B * (H / 16)) != 0) return -1;
  return 0;
}

int rgm_scg_step(rgm_scg* h, const float* mean, const float* g, const float* noise, const float* t_model,
                 const long long* y, const float* x0_a, const float* x0_c, float scale_factor,
                 const rgm_rule_spec* rules_host, int n_rules, int N, int B, int C, int H, int W, float* out_sample,
                 long long* out_index, float* out_scores, void* stream) {
  if (rgm_check_device()) return -1;
  if (!h || !mean || !g || !noise || !t_model || !x0_a || !x0_c || !out_sample || !out_index)
    return set_error("rgm_scg_step: null argument");
  if (N <= 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0 || n_rules < 0 || (n_rules > 0 && !rules_host))
    return set_error("rgm_scg_step: bad sizes");
  Scg* s = reinterpret_cast<Scg*>(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int NB = N * B;
  const long long elems = (long long)C * H * W;
  const int L = s->vae ? 8 * H : W;  // roll length (without a decoder the rules read the latent itself, :524)
  int kmax = 12;
  for (int r = 0; r < n_rules; ++r) {
    const rgm_rule_spec& rs = rules_host[r];
    if (!rs.target) return set_error("rgm_scg_step: rule without a target");
    if (rs.kind == RGM_RULE_PITCH_HIST) continue;
    if (rs.kind != RGM_RULE_NOTE_DENSITY && rs.kind != RGM_RULE_NOTE_DENSITY_CLASS)
      return set_error("rgm_scg_step: unknown rule kind (chord rules are host callables, see INTEGRATION.md)");
    if (rs.interval <= 0 || L % rs.interval != 0) return set_error("rgm_scg_step: note-density interval must divide the roll length");
    kmax = kmax > 2 * (L / rs.interval) ? kmax : 2 * (L / rs.interval);
  }
  if (scg_reserve(s, N, B, C, H, W, kmax, st) != 0) return -1;
  float* cand = static_cast<float*>(s->cand.p);
  float* eps = static_cast<float*>(s->eps.p);
  float* x0 = static_cast<float*>(s->x0.p);
  float* total = static_cast<float*>(s->total.p);
  float* gen = static_cast<float*>(s->gen.p);
  long long* y_rep = static_cast<long long*>(s->rep.p);
  float* t_rep = reinterpret_cast<float*>(y_rep + NB);
  float* a_rep = t_rep + NB;
  float* c_rep = a_rep + NB;

  // :510-514 candidates, :515-518 one denoiser call over all of them, :519 x0
  RGM_CUDA_OK(launch_scg_fanout(mean, g, noise, cand, N, B, elems, st));
  repeat_kernel<<<(NB + 255) / 256, 256, 0, st>>>(t_model, y, x0_a, x0_c, t_rep, y_rep, a_rep, c_rep, B, NB);
  g_aux_launches++;
  RGM_CUDA_OK(cudaGetLastError());
  if (rgm_dit_forward(s->dit, cand, t_rep, y ? y_rep : nullptr, eps, NB, H, stream) != 0) return -1;
  RGM_CUDA_OK(launch_x0_from_eps(cand, eps, a_rep, c_rep, x0, NB, elems, 0, st));
  // :524 _decode (channel 0 only: the rule programs read nothing else)
  float* roll = x0;
  int roll_ch = C;
  if (s->vae) {
    roll = static_cast<float*>(s->roll.p);
    roll_ch = 1;
    if (rgm_vae_decode_latents(s->vae, x0, scale_factor, roll, NB, H, 1, stream) != 0) return -1;
  } else if (H != 128) {
    return set_error("rgm_scg_step: without a decoder the latent itself is the roll and must have 128 rows");
  }
  // :531-538 rules in the caller's order (they write through the roll), weighted -loss
  RGM_CUDA_OK(cudaMemsetAsync(total, 0, (size_t)NB * sizeof(float), st));
  for (int r = 0; r < n_rules; ++r) {
    const rgm_rule_spec& rs = rules_host[r];
    int K = 12;
    if (rs.kind == RGM_RULE_PITCH_HIST) {
      RGM_CUDA_OK(launch_rule_pitch_hist(roll, gen, NB, roll_ch, L, st));
    } else {
      K = 2 * (L / rs.interval);
      RGM_CUDA_OK(launch_rule_note_density(roll, gen, NB, roll_ch, L, rs.interval, rs.horizontal_scale, st));
      if (rs.kind == RGM_RULE_NOTE_DENSITY_CLASS) {
        const long long tot = (long long)NB * K;
        nd_classes_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(gen, tot, K, rs.horizontal_scale);
        g_aux_launches++;
        RGM_CUDA_OK(cudaGetLastError());
      }
    }
    RGM_CUDA_OK(launch_rule_loss_accum(gen, rs.target, total, NB, B, K, rs.loss_kind, rs.weight, st));
  }
  // :539-554 first-max argmax over the N candidates of each sample, gather
  RGM_CUDA_OK(launch_scg_select(total, cand, out_sample, out_index, N, B, elems, st));
  if (out_scores)
    RGM_CUDA_OK(cudaMemcpyAsync(out_scores, total, (size_t)NB * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
