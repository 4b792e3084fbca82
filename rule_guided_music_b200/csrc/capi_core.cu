// C ABI: library-level entry points and the GEMM / convolution building blocks (include/rgm_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rgm_b200.h"
#include "api_util.h"
#include "aux_kernels.h"
#include "gemm_host.h"

namespace rgm {

thread_local std::string g_last_error;

int set_error(const std::string& m) {
  g_last_error = m;
  return -1;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return set_error(std::string(what) + ": " + cudaGetErrorString(e));
}

std::atomic<unsigned long long> g_aux_launches{0};

// ---- optional per-launch timer (rgm_prof_*) -------------------------------------------------------------------
std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec {
  std::string name;
  double flops_alg, flops_exec, bytes;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace

void prof_open(const char* name, double flops_alg, double flops_exec, double bytes, cudaStream_t st) {
  ProfRec r{name, flops_alg, flops_exec, bytes, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void prof_close(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, st);
}

// weight fp32 [Cout,Cin,kh,kw] -> packed fp16 [rows][taps*cin_pad]
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cin,
                                        int cout_pad, int cin_pad, int kind) {
  const int taps = kind == 0 ? 1 : ((kind == 1 || kind == 3) ? 9 : 4);  // 3 = stride-2 Downsample conv: plain 9 taps
  const int npar = kind == 2 ? 4 : 1;
  const long long K = (long long)taps * cin_pad;
  const long long total = (long long)npar * cout_pad * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int t = (int)((i / cin_pad) % taps);
    const long long row = i / K;
    const int co = (int)(row % cout_pad);
    const int par = (int)(row / cout_pad);
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      if (kind == 0) {
        v = w[(long long)co * Cin + ci];
      } else if (kind == 1 || kind == 3) {
        v = w[((long long)co * Cin + ci) * 9 + t];
      } else {
        // nearest-2x upsample then 3x3: output pixel (2i+ph, 2j+pw) reads low-res rows {i+ph-1, i+ph}; tap a = 0/1.
        // ph = 0: a=0 <- ky {0}, a=1 <- ky {1,2};  ph = 1: a=0 <- ky {0,1}, a=1 <- ky {2}.   Same for columns.
        const int ph = par >> 1, pw = par & 1, a = t >> 1, b = t & 1;
        const int ky0 = ph == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
        const int ky1 = ph == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int kx0 = pw == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
        const int kx1 = pw == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        const float* wp = w + ((long long)co * Cin + ci) * 9;
        for (int ky = ky0; ky <= ky1; ++ky)
          for (int kx = kx0; kx <= kx1; ++kx) v += wp[ky * 3 + kx];
      }
    }
    out[i] = __float2half_rn(v);
  }
}

cudaError_t launch_pack_conv_weight(const float* w32, __half* w16, int Cout, int Cin, int cout_pad, int cin_pad,
                                    int kind, cudaStream_t st) {
  pack_conv_weight_kernel<<<296, 256, 0, st>>>(w32, w16, Cout, Cin, cout_pad, cin_pad, kind);
  g_aux_launches++;
  return cudaGetLastError();
}

}  // namespace rgm

using namespace rgm;

extern "C" {

const char* rgm_last_error(void) { return g_last_error.c_str(); }
int rgm_version(void) { return 100; }
unsigned long long rgm_launch_count(void) {
  return gemm_launch_count() + g_aux_launches.load() + aux_launch_count() + attention_launch_count() +
         rules_launch_count();
}

int rgm_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_on.store(on ? 1 : 0);
  return 0;
}

int rgm_prof_summary(char* buf_host, int cap) {
  if (!buf_host || cap <= 0) return set_error("rgm_prof_summary: bad buffer");
  if (cudaDeviceSynchronize() != cudaSuccess) return set_error("rgm_prof_summary: device error");
  struct Agg {
    long long n = 0;
    double ms = 0, fa = 0, fe = 0, by = 0;
  };
  std::map<std::string, Agg> agg;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
      Agg& a = agg[r.name];
      a.n++;
      a.ms += ms;
      a.fa += r.flops_alg;
      a.fe += r.flops_exec;
      a.by += r.bytes;
    }
  }
  std::string out;
  char line[512];
  for (auto& kv : agg) {
    snprintf(line, sizeof line, "%s\t%lld\t%.6f\t%.6e\t%.6e\t%.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms,
             kv.second.fa, kv.second.fe, kv.second.by);
    out += line;
  }
  if ((int)out.size() + 1 > cap) return set_error("rgm_prof_summary: buffer too small");
  memcpy(buf_host, out.c_str(), out.size() + 1);
  return 0;
}

int rgm_check_device(void) {
  // cudaGetDeviceProperties costs milliseconds: query each device once.
  static std::atomic<int> ok_mask{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error("rgm_b200: no CUDA device (this library has no CPU path)");
  if (dev < 31 && (ok_mask.load(std::memory_order_relaxed) >> dev) & 1) return 0;
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
    return set_error("rgm_b200: cannot query the device");
  if (major != 10) {
    char buf[160];
    snprintf(buf, sizeof buf, "rgm_b200: device %d is sm_%d%d; this library is built for sm_100a only", dev, major,
             minor);
    return set_error(buf);
  }
  if (dev < 31) ok_mask.fetch_or(1 << dev);
  return 0;
}

int rgm_gemm_f16(const void* a16, const void* b16, const float* bias, float* out32, int M, int N, int K, int block_n,
                 void* stream) {
  if (rgm_check_device()) return -1;
  GemmDesc d;
  d.A = static_cast<const __half*>(a16);
  d.n_img = 1;
  d.H = 1;
  d.W = M;
  d.C = K;
  d.lda = K;
  d.B = static_cast<const __half*>(b16);
  d.rows_b = N;
  d.N = N;
  d.conv = CONV_1x1;
  d.epi = EPI_F32;
  d.block_n = block_n;
  d.e.out = out32;
  d.e.ldo = N;
  d.e.bias = bias;
  d.e.alpha = 1.f;
  if (getenv("RGM_DEBUG_SKIP_STORE")) d.e.act = 99;  // development knob (see gemm_tc.cuh)
  if (const char* tr = getenv("RGM_DEBUG_TRACE_PTR")) d.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0));
  std::string err;
  if (launch_gemm(d, static_cast<cudaStream_t>(stream), &err) != cudaSuccess) return set_error(err);
  return 0;
}

int rgm_conv_f16(const void* x16, const void* w16_packed, const float* bias, const void* resid16, void* out16,
                 int n_img, int H, int W, int Cin, int Cout, int kind, int block_n, float* gn_part, void* stream) {
  if (rgm_check_device()) return -1;
  GemmDesc d;
  d.A = static_cast<const __half*>(x16);
  d.n_img = n_img;
  d.H = H;
  d.W = W;
  d.C = Cin;
  d.lda = Cin;
  d.B = static_cast<const __half*>(w16_packed);
  d.N = Cout;
  d.rows_b = (kind == CONV_UP2 ? 4 : 1) * Cout;
  d.conv = kind;
  d.epi = EPI_F16;
  d.block_n = block_n;
  d.e.out = out16;
  d.e.ldo = Cout;
  d.e.bias = bias;
  d.e.alpha = 1.f;
  d.e.resid = static_cast<const __half*>(resid16);
  d.e.ldr = Cout;
  d.e.gn_part = gn_part;
  if (const char* tr = getenv("RGM_DEBUG_TRACE_PTR")) d.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0));
  if (kind == CONV_UP2) {
    d.e.up2 = 1;
    d.e.upH = H;
    d.e.upW = W;
  }
  std::string err;
  if (launch_gemm(d, static_cast<cudaStream_t>(stream), &err) != cudaSuccess) return set_error(err);
  return 0;
}

int rgm_gn_apply_f16(const void* x16, const float* ab, void* y16, int n_img, int HW, int C, int swish, void* stream) {
  if (rgm_check_device()) return -1;
  if (!x16 || !ab || !y16) return set_error("rgm_gn_apply_f16: null argument");
  return check_cuda(launch_gn_apply(static_cast<const __half*>(x16), reinterpret_cast<const float2*>(ab),
                                    static_cast<__half*>(y16), n_img, HW, C, swish, static_cast<cudaStream_t>(stream)),
                    "rgm_gn_apply_f16");
}

int rgm_conv_gn_f16(const void* x16_raw, const float* ab_in, const void* w16_packed, const float* bias,
                    const void* resid16, void* out16, int n_img, int H, int W, int Cin, int Cout, float* gn_part,
                    void* stream) {
  if (rgm_check_device()) return -1;
  if (!x16_raw || !ab_in || !w16_packed || !out16) return set_error("rgm_conv_gn_f16: null argument");
  GemmDesc d;
  d.A = static_cast<const __half*>(x16_raw);
  d.n_img = n_img;
  d.H = H;
  d.W = W;
  d.C = Cin;
  d.lda = Cin;
  d.B = static_cast<const __half*>(w16_packed);
  d.N = Cout;
  d.rows_b = Cout;
  d.conv = CONV_3x3;
  d.epi = EPI_F16;
  d.e.out = out16;
  d.e.ldo = Cout;
  d.e.bias = bias;
  d.e.alpha = 1.f;
  d.e.resid = static_cast<const __half*>(resid16);
  d.e.ldr = Cout;
  d.e.gn_part = gn_part;
  if (const char* tr = getenv("RGM_DEBUG_TRACE_PTR")) d.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0));
  if (!conv_gn_shape_ok(d))
    return set_error("rgm_conv_gn_f16: needs a 3x3 conv on [n, H even, 128, Cin % 64 == 0] with 128 output features");
  std::string err;
  if (launch_conv_gn(d, reinterpret_cast<const float2*>(ab_in), static_cast<cudaStream_t>(stream), &err) != cudaSuccess)
    return set_error(err);
  return 0;
}

int rgm_conv_norm_f16(const void* x16, const void* w16_packed, const float* bias, const float* gamma, const float* beta,
                      const void* resid16, void* raw16, void* out16, int n_img, int H, int W, int Cin, int Cout, int kind,
                      int swish, void* gn_scratch, int* gn_err, void* stream) {
  if (rgm_check_device()) return -1;
  if (!x16 || !w16_packed || !gamma || !beta || !out16 || !gn_scratch || !gn_err)
    return set_error("rgm_conv_norm_f16: null argument");
  GemmDesc d;
  d.A = static_cast<const __half*>(x16);
  d.n_img = n_img;
  d.H = H;
  d.W = W;
  d.C = Cin;
  d.lda = Cin;
  d.B = static_cast<const __half*>(w16_packed);
  d.N = Cout;
  d.rows_b = (kind == CONV_UP2 ? 4 : 1) * Cout;
  d.conv = kind;
  d.epi = EPI_F16;
  d.e.ldo = Cout;
  d.e.bias = bias;
  d.e.alpha = 1.f;
  if (kind == CONV_UP2) {
    if (raw16 == nullptr) return set_error("rgm_conv_norm_f16: the upsample conv needs the dual form (raw16)");
    d.e.up2 = 1;
    d.e.upH = H;
    d.e.upW = W;
  }
  if (!gemm_gn_fuse_supported(d))
    return set_error("rgm_conv_norm_f16: needs a 3x3 / 1x1 conv with 128 / 256 / 512 output features on images of a multiple "
                     "of 256 pixels that span no more tiles than there are resident CTAs");
  if (resid16 != nullptr && raw16 == nullptr)
    return set_error("rgm_conv_norm_f16: a residual needs raw16 (the raw tensor is what the residual is added to)");
  if (raw16 != nullptr) {  // dual form: raw tensor (+ residual) and its normalised copy
    d.e.out = raw16;
    d.e.gn_out2 = static_cast<__half*>(out16);
    d.e.resid = static_cast<const __half*>(resid16);
    d.e.ldr = Cout;
  } else {
    d.e.out = out16;
  }
  d.e.gn_sums = static_cast<unsigned long long*>(gn_scratch);
  d.e.gn_gamma = gamma;
  d.e.gn_beta = beta;
  d.e.gn_eps = 1e-6f;
  d.e.gn_swish = swish;
  d.e.gn_err = gn_err;
  if (const char* tr = getenv("RGM_DEBUG_TRACE_PTR")) d.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RGM_CUDA_OK(cudaMemsetAsync(gn_scratch, 0, gn_scratch_bytes(n_img), st));
  std::string err;
  if (launch_gemm(d, st, &err) != cudaSuccess) return set_error(err);
  return 0;
}

int rgm_pack_conv_weight(const float* w32, void* w16_packed, int Cout, int Cin, int cout_pad, int cin_pad, int kind,
                         void* stream) {
  if (rgm_check_device()) return -1;
  if (kind < 0 || kind > 3 || cin_pad < Cin || cout_pad < Cout) return set_error("rgm_pack_conv_weight: bad arguments");
  return check_cuda(launch_pack_conv_weight(w32, static_cast<__half*>(w16_packed), Cout, Cin, cout_pad, cin_pad, kind,
                                            static_cast<cudaStream_t>(stream)),
                    "rgm_pack_conv_weight");
}

}  // extern "C"
