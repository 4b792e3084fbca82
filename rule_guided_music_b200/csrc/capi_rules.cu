// C ABI: rule programs, SCG fan-out / selection and the attention building block (include/rgm_b200.h).
#include <string>

#include "../../include/rgm_b200.h"
#include "api_util.h"
#include "aux_kernels.h"

using namespace rgm;

#define RGM_ENTRY()                  \
  if (rgm_check_device()) return -1; \
  cudaStream_t st = static_cast<cudaStream_t>(stream)

extern "C" {

int rgm_rule_pitch_hist(float* roll, float* hist, int n, int ch, int L, void* stream) {
  RGM_ENTRY();
  if (!roll || !hist) return set_error("rgm_rule_pitch_hist: null argument");
  return check_cuda(launch_rule_pitch_hist(roll, hist, n, ch, L, st), "rgm_rule_pitch_hist");
}

int rgm_rule_note_density(float* roll, float* out, int n, int ch, int L, int interval, float horizontal_scale,
                          void* stream) {
  RGM_ENTRY();
  if (!roll || !out) return set_error("rgm_rule_note_density: null argument");
  if (L % 128 != 0 || interval <= 0 || L % interval != 0 ||
      (interval < 128 ? (128 % interval) != 0 : (interval % 128) != 0))
    return set_error("rgm_rule_note_density: L must be a multiple of 128 and interval a divisor or multiple of 128");
  return check_cuda(launch_rule_note_density(roll, out, n, ch, L, interval, horizontal_scale, st),
                    "rgm_rule_note_density");
}

int rgm_rule_loss_accum(const float* gen, const float* target, float* total, int n, int B, int K, int kind,
                        float weight, void* stream) {
  RGM_ENTRY();
  if (!gen || !target || !total || B <= 0 || K <= 0) return set_error("rgm_rule_loss_accum: bad argument");
  return check_cuda(launch_rule_loss_accum(gen, target, total, n, B, K, kind, weight, st), "rgm_rule_loss_accum");
}

int rgm_scg_fanout(const float* mean, const float* g, const float* noise, float* cand, int N, int B, long long elems,
                   void* stream) {
  RGM_ENTRY();
  if (!mean || !g || !noise || !cand) return set_error("rgm_scg_fanout: null argument");
  return check_cuda(launch_scg_fanout(mean, g, noise, cand, N, B, elems, st), "rgm_scg_fanout");
}

int rgm_x0_from_eps(const float* x, const float* eps, const float* a, const float* c, float* x0, int B,
                    long long elems, int clamp, void* stream) {
  RGM_ENTRY();
  if (!x || !eps || !a || !c || !x0) return set_error("rgm_x0_from_eps: null argument");
  return check_cuda(launch_x0_from_eps(x, eps, a, c, x0, B, elems, clamp, st), "rgm_x0_from_eps");
}

int rgm_scg_select(const float* total, const float* cand, float* out, long long* idx, int N, int B, long long elems,
                   void* stream) {
  RGM_ENTRY();
  if (!total || !cand || !out || !idx) return set_error("rgm_scg_select: null argument");
  return check_cuda(launch_scg_select(total, cand, out, idx, N, B, elems, st), "rgm_scg_select");
}

int rgm_attention_f16(const void* q16, const void* k16, const void* vt16, void* out16, int B, int heads, int T, int dh,
                      float scale, void* stream) {
  RGM_ENTRY();
  std::string err;
  if (launch_attention(static_cast<const __half*>(q16), static_cast<const __half*>(k16),
                       static_cast<const __half*>(vt16), static_cast<__half*>(out16), B, heads, T, dh, scale, st,
                       &err) != cudaSuccess)
    return set_error(err);
  return 0;
}

}  // extern "C"
