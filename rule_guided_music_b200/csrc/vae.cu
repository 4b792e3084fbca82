// taming KL-VAE decoder on sm_100a (reference taming/models/klvae_pedal.py:80-85 and
// taming/modules/diffusionmodules/model.py:436-537): post_quant_conv + Decoder, driven per chunk of 16x16 latent
// tiles, activations NHWC fp16, every convolution an implicit GEMM on tcgen05 (gemm_tc.cuh) with
//   - bias and the ResnetBlock / AttnBlock residual add in the epilogue             model.py:117-137, 168-192
//   - GroupNorm partial statistics of the stored tensor emitted by the epilogue     model.py:34-35
//   - nearest-2x upsample folded into the convolution (4 parity sub-convolutions)   model.py:49-53
//   - conv_out scattering straight into the piano roll [cand, ch, 128, 8*Hlat]      gaussian_diffusion.py:1355
// GroupNorm apply + swish is one fp16 -> fp16 pass (aux_kernels.cu).  The latent re-tiling of _decode
// (gaussian_diffusion.py:1347-1358), post_quant_conv and conv_in are one fp32 kernel.
// The ENCODER (model.py:342-433 + quant_conv, klvae_pedal.py:60-68; used once per run by scripts/edit.py through
// gaussian_diffusion._encode :1382-1395) reuses every block: conv_in is an fp32 CUDA-core kernel (3 input channels),
// Downsample is the stride-2 implicit GEMM (TMA element strides), conv_out + quant_conv end in fp32.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/rgm_b200.h"
#include "api_util.h"
#include "aux_kernels.h"
#include "gemm_host.h"

namespace rgm {


namespace {

struct Conv {
  int cin = 0, cout = 0, k = 3, cout_pad = 0;
  int kind = CONV_3x3;
  __half* w = nullptr;  // packed
  float* b = nullptr;   // [cout_pad]
};
struct Norm {
  int c = 0;
  float *gamma = nullptr, *beta = nullptr;
};
struct Res {
  Norm n1, n2;
  Conv c1, c2, nin;
  bool has_nin = false;
};

using DevBuf = GrowBuf;  // api_util.h: grows by retiring, never frees what a captured graph may reference

}  // namespace

struct Vae {
  int ch = 128, n_levels = 4, nres = 2, zc = 4, out_ch = 3;
  std::vector<int> mult;
  // stem (fp32)
  float *pq_w = nullptr, *pq_b = nullptr, *cin_w = nullptr, *cin_b = nullptr;
  int block_in0 = 0;
  Res mid1, mid2;
  Norm attn_norm;
  Conv attn_q, attn_k, attn_v, attn_proj;
  std::vector<std::vector<Res>> up;     // [level][block]
  std::vector<Conv> upsample;           // [level] (level 0 unused)
  Norm norm_out;
  float *cout_w = nullptr, *cout_b = nullptr;  // conv_out stays fp32 at load; it runs fused on mma.sync (vae_out_kernel)
  void* cout_bfrag = nullptr;                  // ... from fp16 B fragments packed once per weight load
  int cout_cin = 0;
  // encoder (model.py:342-433) + quant_conv
  int in_ch = 3;
  float *e_cin_w = nullptr, *e_cin_b = nullptr, *q_w = nullptr, *q_b = nullptr;
  std::vector<std::vector<Res>> down;   // [level][block]
  std::vector<Conv> downsample;         // [level] (last level unused)
  Res e_mid1, e_mid2;
  Norm e_attn_norm, e_norm_out;
  Conv e_attn_q, e_attn_k, e_attn_v, e_attn_proj, e_cout;
  std::map<std::string, std::pair<float*, long long>> f32_keys;  // key -> (dst, numel)
  std::map<std::string, Conv*> conv_keys;                        // "<name>.weight" -> conv
  std::vector<void*> allocs;
  // Two independent chunk pipelines ("lanes"), each with its own activation buffers and stream: consecutive chunks
  // alternate between them, so the HBM-bound GroupNorm passes of one chunk run while the tensor-bound convolutions
  // of the other occupy the tcgen05 pipe (a GEMM CTA takes all shared memory of an SM but leaves registers and
  // thread slots for an elementwise block).
  struct Lane {
    // abbuf[2]: GroupNorm affines ping-pong -- a convolution reads its input's affine from one (fused normalise, or the
    // gn_apply pass before it) while its output's affine is written into the other
    DevBuf act[4], gnpart, abbuf[2], attn_s;
    DevBuf gncount;  // statistics accumulators (with arrival counts) of the GroupNorm-in-epilogue convolutions
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
  } lane[2];
  cudaEvent_t fork = nullptr;
  int n_lanes = 2;
  int chunk_tiles = 256;  // tiles per decode chunk (RGM_VAE_CHUNK): 256 vs 128 measured -0.5 % of the step (fewer launch tails)
  // conv1 of a ResnetBlock applies norm2 + swish to its own output inside its epilogue (gemm_tc.cuh,
  // gn_epilogue_loop): RGM_GN_EPI=0 restores the separate normalise pass
  bool gn_epi = true;
  // ... and conv2 / proj_out write the raw tensor AND its copy normalised for the next block ("dual" form): RGM_GN_DUAL=0
  // keeps the separate pass for those.  The upsample convs can do the same (RGM_GN_DUAL_UP=1) but lose: with K = 1024 the
  // mainloop is shorter than the two-pass epilogue (measured 1.62 vs 0.99 + 0.43 ms per 64x64 -> 128x128 launch).
  bool gn_dual = true, gn_dual_up = false;
  int gn_dual_min = 256;  // narrowest layer that uses the dual form (RGM_GN_DUAL_MIN)
  // flag "a GroupNorm-in-epilogue wait gave up" in pinned, mapped host memory: the kernels write it through gn_err (the
  // device alias), the host reads gn_err_h without a synchronisation at the start of every decode / encode call, so a
  // run that produced garbage fails loudly at the next call (and rgm_vae_gn_timeouts reports it after a sync)
  int* gn_err = nullptr;
  int* gn_err_h = nullptr;

  ~Vae() {
    for (void* p : allocs) cudaFree(p);
    if (gn_err_h) cudaFreeHost(gn_err_h);
    for (auto& l : lane) {
      if (l.stream) cudaStreamDestroy(l.stream);
      if (l.done) cudaEventDestroy(l.done);
    }
    if (fork) cudaEventDestroy(fork);
  }
  template <typename T>
  T* alloc(long long n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)n * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, (size_t)n * sizeof(T));
    allocs.push_back(p);
    return static_cast<T*>(p);
  }
  bool make_conv(Conv& c, const std::string& name, int cin, int cout, int k, int kind) {
    c.cin = cin;
    c.cout = cout;
    c.k = k;
    c.kind = kind;
    c.cout_pad = cout % 128 == 0 ? cout : (int)((cout + 31) / 32 * 32);
    const int taps = kind == CONV_1x1 ? 1 : (kind == CONV_UP2 ? 4 : 9);
    const int npar = kind == CONV_UP2 ? 4 : 1;
    c.w = alloc<__half>((long long)npar * c.cout_pad * taps * cin);
    c.b = alloc<float>(c.cout_pad);
    if (!c.w || !c.b) return false;
    conv_keys[name + ".weight"] = &c;
    f32_keys[name + ".bias"] = {c.b, cout};
    return true;
  }
  bool make_norm(Norm& n, const std::string& name, int c) {
    n.c = c;
    n.gamma = alloc<float>(c);
    n.beta = alloc<float>(c);
    if (!n.gamma || !n.beta) return false;
    f32_keys[name + ".weight"] = {n.gamma, c};
    f32_keys[name + ".bias"] = {n.beta, c};
    return true;
  }
  bool make_res(Res& r, const std::string& p, int cin, int cout) {
    r.has_nin = cin != cout;
    bool ok = make_norm(r.n1, p + ".norm1", cin) && make_conv(r.c1, p + ".conv1", cin, cout, 3, CONV_3x3) &&
              make_norm(r.n2, p + ".norm2", cout) && make_conv(r.c2, p + ".conv2", cout, cout, 3, CONV_3x3);
    if (ok && r.has_nin) ok = make_conv(r.nin, p + ".nin_shortcut", cin, cout, 1, CONV_1x1);
    return ok;
  }
};

namespace {

int vae_build(Vae* m) {
  int block_in = m->ch * m->mult[m->n_levels - 1];
  m->block_in0 = block_in;
  m->pq_w = m->alloc<float>(m->zc * m->zc);
  m->pq_b = m->alloc<float>(m->zc);
  m->cin_w = m->alloc<float>((long long)block_in * m->zc * 9);
  m->cin_b = m->alloc<float>(block_in);
  if (!m->pq_w || !m->pq_b || !m->cin_w || !m->cin_b) return set_error("rgm_vae_create: out of memory");
  m->f32_keys["post_quant_conv.weight"] = {m->pq_w, (long long)m->zc * m->zc};
  m->f32_keys["post_quant_conv.bias"] = {m->pq_b, m->zc};
  m->f32_keys["decoder.conv_in.weight"] = {m->cin_w, (long long)block_in * m->zc * 9};
  m->f32_keys["decoder.conv_in.bias"] = {m->cin_b, block_in};
  bool ok = m->make_res(m->mid1, "decoder.mid.block_1", block_in, block_in) &&
            m->make_norm(m->attn_norm, "decoder.mid.attn_1.norm", block_in) &&
            m->make_conv(m->attn_q, "decoder.mid.attn_1.q", block_in, block_in, 1, CONV_1x1) &&
            m->make_conv(m->attn_k, "decoder.mid.attn_1.k", block_in, block_in, 1, CONV_1x1) &&
            m->make_conv(m->attn_v, "decoder.mid.attn_1.v", block_in, block_in, 1, CONV_1x1) &&
            m->make_conv(m->attn_proj, "decoder.mid.attn_1.proj_out", block_in, block_in, 1, CONV_1x1) &&
            m->make_res(m->mid2, "decoder.mid.block_2", block_in, block_in);
  m->up.resize(m->n_levels);
  m->upsample.resize(m->n_levels);
  for (int lvl = m->n_levels - 1; ok && lvl >= 0; --lvl) {
    const int block_out = m->ch * m->mult[lvl];
    m->up[lvl].resize(m->nres + 1);
    for (int b = 0; ok && b <= m->nres; ++b) {
      ok = m->make_res(m->up[lvl][b], "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(b), block_in,
                       block_out);
      block_in = block_out;
    }
    if (ok && lvl != 0)
      ok = m->make_conv(m->upsample[lvl], "decoder.up." + std::to_string(lvl) + ".upsample.conv", block_in, block_in, 3,
                        CONV_UP2);
  }
  ok = ok && m->make_norm(m->norm_out, "decoder.norm_out", block_in);
  if (cudaHostAlloc(reinterpret_cast<void**>(&m->gn_err_h), sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *m->gn_err_h = 0;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->gn_err), m->gn_err_h, 0) != cudaSuccess) m->gn_err = nullptr;
  }
  ok = ok && m->gn_err != nullptr;
  m->cout_cin = block_in;
  m->cout_w = m->alloc<float>((long long)m->out_ch * block_in * 9);
  m->cout_b = m->alloc<float>(m->out_ch);
  m->cout_bfrag = m->alloc<uint2>(9LL * (block_in / 16) * 32);
  if (!ok || !m->cout_w || !m->cout_b || !m->cout_bfrag) return set_error("rgm_vae_create: out of memory");
  m->f32_keys["decoder.conv_out.weight"] = {m->cout_w, (long long)m->out_ch * block_in * 9};
  m->f32_keys["decoder.conv_out.bias"] = {m->cout_b, m->out_ch};

  // ---- encoder ---------------------------------------------------------------------------------------------
  const int zz = 2 * m->zc;  // double_z: mean and log-variance
  m->e_cin_w = m->alloc<float>((long long)m->ch * m->in_ch * 9);
  m->e_cin_b = m->alloc<float>(m->ch);
  m->q_w = m->alloc<float>((long long)zz * zz);
  m->q_b = m->alloc<float>(zz);
  if (!m->e_cin_w || !m->e_cin_b || !m->q_w || !m->q_b) return set_error("rgm_vae_create: out of memory");
  m->f32_keys["encoder.conv_in.weight"] = {m->e_cin_w, (long long)m->ch * m->in_ch * 9};
  m->f32_keys["encoder.conv_in.bias"] = {m->e_cin_b, m->ch};
  m->f32_keys["quant_conv.weight"] = {m->q_w, (long long)zz * zz};
  m->f32_keys["quant_conv.bias"] = {m->q_b, zz};
  m->down.resize(m->n_levels);
  m->downsample.resize(m->n_levels);
  block_in = m->ch;
  for (int lvl = 0; ok && lvl < m->n_levels; ++lvl) {
    const int block_out = m->ch * m->mult[lvl];
    m->down[lvl].resize(m->nres);
    for (int b = 0; ok && b < m->nres; ++b) {
      ok = m->make_res(m->down[lvl][b], "encoder.down." + std::to_string(lvl) + ".block." + std::to_string(b), block_in,
                       block_out);
      block_in = block_out;
    }
    if (ok && lvl != m->n_levels - 1)
      ok = m->make_conv(m->downsample[lvl], "encoder.down." + std::to_string(lvl) + ".downsample.conv", block_in,
                        block_in, 3, CONV_DOWN2);
  }
  ok = ok && m->make_res(m->e_mid1, "encoder.mid.block_1", block_in, block_in) &&
       m->make_norm(m->e_attn_norm, "encoder.mid.attn_1.norm", block_in) &&
       m->make_conv(m->e_attn_q, "encoder.mid.attn_1.q", block_in, block_in, 1, CONV_1x1) &&
       m->make_conv(m->e_attn_k, "encoder.mid.attn_1.k", block_in, block_in, 1, CONV_1x1) &&
       m->make_conv(m->e_attn_v, "encoder.mid.attn_1.v", block_in, block_in, 1, CONV_1x1) &&
       m->make_conv(m->e_attn_proj, "encoder.mid.attn_1.proj_out", block_in, block_in, 1, CONV_1x1) &&
       m->make_res(m->e_mid2, "encoder.mid.block_2", block_in, block_in) &&
       m->make_norm(m->e_norm_out, "encoder.norm_out", block_in) &&
       m->make_conv(m->e_cout, "encoder.conv_out", block_in, zz, 3, CONV_3x3);
  if (!ok) return set_error("rgm_vae_create: out of memory");
  return 0;
}

struct Ctx {
  Vae* m;
  Vae::Lane* L;
  cudaStream_t st;
  int nt;        // tiles in this chunk
  int ab_i = 0;  // which of L->abbuf holds the affine of the tensor about to be normalised
  float2* ab() const { return static_cast<float2*>(L->abbuf[ab_i].p); }
  float2* ab_other() const { return static_cast<float2*>(L->abbuf[ab_i ^ 1].p); }
};

#define RGM_VGEMM_OK(desc)                                                                    \
  do {                                                                                        \
    std::string _err;                                                                         \
    if (launch_gemm(desc, c.st, &_err) != cudaSuccess) return set_error("rgm_vae: " + _err);  \
  } while (0)

// conv on NHWC fp16 [nt, H, H, cin] -> out [nt, H', H', cout]; optional residual.  `next` = the GroupNorm that consumes
// `out` (or null): the epilogue then emits the partial statistics of the stored tensor and a small finalize launch folds
// them into that norm's per-(tile, channel) affine, so no statistics pass over the tensor follows.  (Folding inside the
// epilogue -- last-arriving warp per image behind a device-scope fence -- was measured in round 2: the fences and the
// serial fold cost 2-10x on the short-K convolutions; profiles/README.md.)
// `out_norm` (instead of `next`): `out` receives swish(out_norm(conv(x))) -- the statistics are exchanged between the CTAs
// of an image while the accumulators wait in tensor memory, and the raw convolution output is never written.
GemmDesc conv_desc(const Ctx& c, const Conv& cv, const __half* x, int H, const __half* resid, __half* out) {
  GemmDesc d;
  d.A = x;
  d.n_img = c.nt;
  d.H = H;
  d.W = H;
  d.C = cv.cin;
  d.lda = cv.cin;
  d.B = cv.w;
  d.N = cv.cout;
  d.rows_b = (cv.kind == CONV_UP2 ? 4 : 1) * cv.cout;
  d.conv = cv.kind;
  d.epi = EPI_F16;
  d.e.out = out;
  d.e.ldo = cv.cout;
  d.e.bias = cv.b;
  d.e.alpha = 1.f;
  d.e.resid = resid;
  d.e.ldr = cv.cout;
  if (cv.kind == CONV_UP2) {
    d.e.up2 = 1;
    d.e.upH = H;
    d.e.upW = H;
  }
  return d;
}

// `norm_copy` with `out_norm` ("dual" form): `out` receives the raw tensor (+ residual) as usual and `norm_copy` its
// normalised copy, written by the same launch one tile later from the warp's own L2-resident rows.
int run_conv(Ctx& c, const Conv& cv, const __half* x, int H, const __half* resid, __half* out, const Norm* next,
             bool fuse_input_norm = false, const Norm* out_norm = nullptr, __half* norm_copy = nullptr,
             bool norm_swish = true) {
  GemmDesc d = conv_desc(c, cv, x, H, resid, out);
  if (next != nullptr) {
    if (next->c != cv.cout || cv.cout % 128 != 0) return set_error("rgm_vae: GroupNorm partials need a 128-multiple channel count");
    d.e.gn_part = static_cast<float*>(c.L->gnpart.p);
  }
  if (out_norm != nullptr) {
    if (next != nullptr || out_norm->c != cv.cout || (resid != nullptr && norm_copy == nullptr))
      return set_error("rgm_vae: bad GroupNorm-in-epilogue request");
    d.e.gn_sums = static_cast<unsigned long long*>(c.L->gncount.p);
    d.e.gn_gamma = out_norm->gamma;
    d.e.gn_beta = out_norm->beta;
    d.e.gn_eps = 1e-6f;
    d.e.gn_swish = norm_swish ? 1 : 0;
    d.e.gn_err = c.m->gn_err;
    d.e.gn_out2 = norm_copy;
    RGM_CUDA_OK(cudaMemsetAsync(c.L->gncount.p, 0, gn_scratch_bytes(c.nt), c.st));
  }
  if (fuse_input_norm) {  // x is RAW: normalise + swish with the current affine inside the operand path (conv_gn.cuh)
    if (out_norm != nullptr) return set_error("rgm_vae: the consumer-side kernel cannot normalise its own output");
    std::string err;
    if (launch_conv_gn(d, c.ab(), c.st, &err) != cudaSuccess) return set_error("rgm_vae: " + err);
  } else {
    RGM_VGEMM_OK(d);
  }
  if (next != nullptr) {
    // fold the epilogue's partial sums into the consumer norm's per-(tile, channel) affine, in the OTHER affine buffer
    // (this convolution may still be reading the current one), which then becomes the current one
    const int oH = cv.kind == CONV_DOWN2 ? H / 2 : H;  // GEMM rows are output pixels (low-res ones for UP2)
    const int npar = cv.kind == CONV_UP2 ? 4 : 1;
    const int low = oH * oH;
    if (low % 128 != 0) return set_error("rgm_vae: GroupNorm partials need images of a multiple of 128 pixels");
    RGM_CUDA_OK(launch_gn_finalize(static_cast<float*>(c.L->gnpart.p), next->gamma, next->beta, c.ab_other(), c.nt,
                                   low / 128, npar, (long long)c.nt * low / 128, next->c, npar * low, 1e-6f, c.st));
    c.ab_i ^= 1;
  }
  return 0;
}

// whether conv(swish(norm(x))) of this layer can run as one kernel
bool can_fuse_norm(const Ctx& c, const Conv& cv, int H) {
  GemmDesc d;
  d.n_img = c.nt;
  d.H = H;
  d.W = H;
  d.C = cv.cin;
  d.lda = cv.cin;
  d.N = cv.cout;
  d.conv = cv.kind;
  d.epi = EPI_F16;
  return conv_gn_supported(d);
}

// whether this convolution can apply the GroupNorm that consumes its output inside its own epilogue
bool can_fuse_out_norm(const Ctx& c, const Conv& cv, int H) {
  if (!c.m->gn_epi) return false;
  return gemm_gn_fuse_supported(conv_desc(c, cv, nullptr, H, nullptr, nullptr));
}

// Which of the lane's four activation buffers holds the current tensor x and -- when its producer already wrote the
// normalised copy the next GroupNorm would compute ("dual" epilogue) -- that copy (else -1).
struct Act {
  int x = 0;
  int xn = -1;
};

int take_free(bool (&used)[4]) {
  for (int i = 0; i < 4; ++i)
    if (!used[i]) {
      used[i] = true;
      return i;
    }
  return -1;  // cannot happen: at most three buffers are live at any point of a block
}

// ResnetBlock (model.py:117-137) on buf[a.x] (its norm1 affine current in L->abbuf -- its producer's epilogue put it
// there -- unless stats_from_tensor, or already applied: buf[a.xn]).  On return a.x is the block's output and a.xn its
// copy normalised by `next` (with swish iff next_swish) when want_copy and the layer qualifies; otherwise the affine of
// `next` is left current.  next == nullptr: nobody normalises the output (an upsample / downsample conv follows).
int run_res(Ctx& c, const Res& r, __half* const* buf, Act& a, int H, const Norm* next, bool next_swish = true,
            bool want_copy = true, bool stats_from_tensor = false) {
  const int HW = H * H;
  bool used[4] = {false, false, false, false};
  used[a.x] = true;
  if (a.xn >= 0) used[a.xn] = true;
  const __half* x = buf[a.x];
  if (stats_from_tensor && a.xn < 0)  // x did not come out of a GEMM epilogue (the stems): one direct statistics pass
    RGM_CUDA_OK(launch_gn_stats(x, r.n1.gamma, r.n1.beta, c.ab(), c.nt, HW, r.n1.c, 1e-6f, c.st));
  // f1 / f2: the opt-in consumer-side kernel (RGM_CONV_GN=1, conv_gn.cuh) normalises the INPUT of conv1 / conv2 in its
  // operand path; its epilogue is the plain one, so a convolution that runs on it cannot also normalise its OUTPUT
  const bool f1 = a.xn < 0 && can_fuse_norm(c, r.c1, H);
  const bool e1 = !f1 && can_fuse_out_norm(c, r.c1, H);  // conv1 writes swish(norm2(conv1(.))) itself: no pass over h
  const bool f2 = !e1 && can_fuse_norm(c, r.c2, H);
  int in1 = a.x;  // (f1: the raw tensor, normalised inside the operand path)
  if (a.xn >= 0) {
    in1 = a.xn;
  } else if (!f1) {
    in1 = take_free(used);
    RGM_CUDA_OK(launch_gn_apply(x, c.ab(), buf[in1], c.nt, HW, r.n1.c, 1, c.st));
  }
  const int h = take_free(used);
  if (run_conv(c, r.c1, buf[in1], H, nullptr, buf[h], e1 ? nullptr : &r.n2, f1, e1 ? &r.n2 : nullptr)) return -1;
  if (in1 != a.x) used[in1] = false;  // the normalised input is dead
  int in2 = h;
  if (!e1 && !f2) {
    in2 = take_free(used);
    RGM_CUDA_OK(launch_gn_apply(buf[h], c.ab(), buf[in2], c.nt, HW, r.n2.c, 1, c.st));
    used[h] = false;
  }
  const int o = take_free(used);
  const __half* resid = x;
  if (r.has_nin) {
    if (run_conv(c, r.nin, x, H, nullptr, buf[o], nullptr)) return -1;
    resid = buf[o];  // conv2 adds the shortcut it finds in `out` and overwrites it (same thread reads then writes)
  }
  // (not for the 128-feature convolutions: their K = 1152 mainloop on one CTA is shorter than the two-pass epilogue --
  // measured in the step 0.90 ms against 0.58 + 0.22 ms for convolution + normalise pass; the 256- and 512-feature
  // layers on CTA pairs gain 0.06 / 0.015 / 0.006 ms per launch)
  const bool dual =
      next != nullptr && want_copy && !f2 && c.m->gn_dual && r.c2.cout >= c.m->gn_dual_min && can_fuse_out_norm(c, r.c2, H);
  const int on = dual ? take_free(used) : -1;  // x, conv2's input and the output are live: the fourth buffer is free
  if (run_conv(c, r.c2, buf[in2], H, resid, buf[o], dual ? nullptr : next, f2, dual ? next : nullptr,
               dual ? buf[on] : nullptr, next_swish))
    return -1;
  a.x = o;
  a.xn = on;
  return 0;
}

// AttnBlock (model.py:168-192) at 16x16: single head over the 256 positions of a tile, on buf[a.x] (the affine of `norm`
// current, or already applied -- without swish -- in buf[a.xn]).  On return a.x = x + proj_out(attention) and a.xn its
// copy normalised by `next` (or -1 with the affine of `next` current).  Uses all four buffers.
int run_mid_attn(Ctx& c, const Norm& norm, const Conv& cq, const Conv& ck, const Conv& cvv, const Conv& cproj,
                 __half* const* buf, Act& a, int H, int C, const Norm* next) {
  const int nt = c.nt;
  cudaStream_t st = c.st;
  Vae::Lane* L = c.L;
  const int HW = H * H;
  bool used[4] = {false, false, false, false};
  used[a.x] = true;
  if (a.xn >= 0) used[a.xn] = true;
  __half* x = buf[a.x];
  if (norm.c != C) return set_error("rgm_vae: attention norm width");
  int ihn = a.xn;
  if (ihn < 0) {
    ihn = take_free(used);
    RGM_CUDA_OK(launch_gn_apply(x, c.ab(), buf[ihn], nt, HW, C, 0, st));  // the affine of `norm` is current
  }
  __half* hn = buf[ihn];
  const int b1 = take_free(used), b2 = take_free(used);
  // q, k, v and v^T, P, attention output carve two buffers (each holds >= 4 tensors of this size: buffers are sized for
  // 128x128x256)
  const long long tsz = (long long)nt * HW * C;
  __half* q = buf[b1];
  __half* k = buf[b1] + tsz;
  __half* v = buf[b1] + 2 * tsz;
  __half* vT = buf[b2];
  __half* P = buf[b2] + tsz;
  __half* ao = buf[b2] + 2 * tsz;
  if (run_conv(c, cq, hn, H, nullptr, q, nullptr)) return -1;
  if (run_conv(c, ck, hn, H, nullptr, k, nullptr)) return -1;
  if (run_conv(c, cvv, hn, H, nullptr, v, nullptr)) return -1;
  RGM_CUDA_OK(launch_transpose(v, vT, nt, HW, C, st));
  float* S = static_cast<float*>(L->attn_s.p);
  {
    GemmDesc d;  // S[b] = q[b] k[b]^T * C^-0.5
    d.A = q;
    d.n_img = nt;
    d.H = 1;
    d.W = HW;
    d.C = C;
    d.lda = C;
    d.B = k;
    d.rows_b = HW;
    d.b_batch = nt;
    d.N = HW;
    d.conv = CONV_1x1;
    d.epi = EPI_F32;
    d.e.out = S;
    d.e.ldo = HW;
    d.e.alpha = 1.0f / sqrtf((float)C);
    RGM_VGEMM_OK(d);
  }
  RGM_CUDA_OK(launch_softmax_rows(S, P, (long long)nt * HW, HW, st));
  {
    GemmDesc d;  // ao[b] = P[b] v[b]  (B operand = v^T [C, HW])
    d.A = P;
    d.n_img = nt;
    d.H = 1;
    d.W = HW;
    d.C = HW;
    d.lda = HW;
    d.B = vT;
    d.rows_b = C;
    d.b_batch = nt;
    d.N = C;
    d.conv = CONV_1x1;
    d.epi = EPI_F16;
    d.e.out = ao;
    d.e.ldo = C;
    d.e.alpha = 1.f;
    RGM_VGEMM_OK(d);
  }
  // x + proj_out(attention) over the (dead) normalised input; q / k / v are dead too: their buffer takes the copy
  const bool dual = next != nullptr && c.m->gn_dual && can_fuse_out_norm(c, cproj, H);
  if (run_conv(c, cproj, ao, H, x, hn, dual ? nullptr : next, false, dual ? next : nullptr, dual ? buf[b1] : nullptr, true))
    return -1;
  a.x = ihn;
  a.xn = dual ? b1 : -1;
  return 0;
}

int decode_chunk(Vae* m, Vae::Lane* L, const float* lat, float scale, float* roll, int n_cand, int Hlat, int roll_ch,
                 int tile0, int nt, cudaStream_t st) {
  Ctx c{m, L, st, nt};
  __half* buf[4];
  for (int i = 0; i < 4; ++i) buf[i] = static_cast<__half*>(L->act[i].p);
  int H = 16;
  int C = m->block_in0;
  // stem
  RGM_CUDA_OK(launch_vae_stem(lat, scale, m->pq_w, m->pq_b, m->cin_w, m->cin_b, buf[0], n_cand, Hlat, tile0, nt, C,
                              st));
  // mid.block_1: the stem has no GEMM epilogue, so its GroupNorm statistics come from a direct pass
  Act a;  // the stem wrote buf[0]
  if (run_res(c, m->mid1, buf, a, H, &m->attn_norm, /*next_swish=*/false, true, /*stats_from_tensor=*/true)) return -1;
  // mid.attn_1 (model.py:168-192)
  if (run_mid_attn(c, m->attn_norm, m->attn_q, m->attn_k, m->attn_v, m->attn_proj, buf, a, H, C, &m->mid2.n1)) return -1;
  // mid.block_2
  if (run_res(c, m->mid2, buf, a, H, &m->up[m->n_levels - 1][0].n1)) return -1;
  for (int lvl = m->n_levels - 1; lvl >= 0; --lvl) {
    for (int b = 0; b <= m->nres; ++b) {
      // who normalises this block's output: the next block, norm_out after the last one (inside vae_out_kernel, from the
      // raw tensor and the affine: no copy), nobody before an upsample conv
      const bool last = b == m->nres && lvl == 0;
      const Norm* next = b < m->nres ? &m->up[lvl][b + 1].n1 : (lvl == 0 ? &m->norm_out : nullptr);
      if (run_res(c, m->up[lvl][b], buf, a, H, next, true, /*want_copy=*/!last)) return -1;
    }
    if (lvl != 0) {
      const int o = (a.x + 1) & 3, on = (a.x + 2) & 3;  // (a.xn is -1 here)
      const Norm* next = &m->up[lvl - 1][0].n1;
      const bool dual = m->gn_dual_up && can_fuse_out_norm(c, m->upsample[lvl], H);  // (off: see Vae::gn_dual_up)
      if (run_conv(c, m->upsample[lvl], buf[a.x], H, nullptr, buf[o], dual ? nullptr : next, false, dual ? next : nullptr,
                   dual ? buf[on] : nullptr, true))
        return -1;
      a.x = o;
      a.xn = dual ? on : -1;
      H *= 2;
    }
  }
  const int cur = a.x;
  // norm_out + swish + conv_out, assembled into the roll: one fused CUDA-core kernel (aux_kernels.cu)
  {
    if (H != 128) return set_error("rgm_vae: decoder output is not 128x128 (the roll kernels assume 128x128 tiles)");
    RGM_CUDA_OK(launch_vae_out(buf[cur], c.ab(), m->cout_bfrag, m->cout_b, roll, nt, m->norm_out.c, m->out_ch, tile0, n_cand,
                               8 * Hlat, roll_ch, st));
  }
  return 0;
}


// Encoder + quant_conv for tiles [t0, t0+nt): x f32 NCHW [n, in_ch, 128, 128] -> moments f32 NCHW [n, 2*zc, 16, 16]
int encode_chunk(Vae* m, Vae::Lane* L, const float* x, float* moments, int t0, int nt, cudaStream_t st) {
  Ctx c{m, L, st, nt};
  __half* buf[4];
  for (int i = 0; i < 4; ++i) buf[i] = static_cast<__half*>(L->act[i].p);
  int H = 128;
  RGM_CUDA_OK(launch_vae_enc_stem(x + (long long)t0 * m->in_ch * 128 * 128, m->e_cin_w, m->e_cin_b, buf[0], nt, m->in_ch,
                                  m->ch, st));
  Act a;  // the stem wrote buf[0]
  bool first = true;
  for (int lvl = 0; lvl < m->n_levels; ++lvl) {
    const bool last_lvl = lvl == m->n_levels - 1;
    for (int b = 0; b < m->nres; ++b) {
      // who normalises this block's output: the next block, mid.block_1 after the last level, nobody before a Downsample
      const Norm* next = b < m->nres - 1 ? &m->down[lvl][b + 1].n1 : (last_lvl ? &m->e_mid1.n1 : nullptr);
      if (run_res(c, m->down[lvl][b], buf, a, H, next, true, true, first)) return -1;
      first = false;
    }
    if (!last_lvl) {
      const int o = (a.x + 1) & 3;  // (a.xn is -1 here)
      if (run_conv(c, m->downsample[lvl], buf[a.x], H, nullptr, buf[o], &m->down[lvl + 1][0].n1)) return -1;  // H -> H/2
      a.x = o;
      a.xn = -1;
      H /= 2;
    }
  }
  if (H != 16) return set_error("rgm_vae_encode: encoder output is not 16x16");
  const int C = m->e_mid1.c1.cin;
  if (run_res(c, m->e_mid1, buf, a, H, &m->e_attn_norm, /*next_swish=*/false)) return -1;
  if (run_mid_attn(c, m->e_attn_norm, m->e_attn_q, m->e_attn_k, m->e_attn_v, m->e_attn_proj, buf, a, H, C, &m->e_mid2.n1))
    return -1;
  // norm_out is applied by the pass below, from the raw tensor and the affine: no normalised copy
  if (run_res(c, m->e_mid2, buf, a, H, &m->e_norm_out, true, /*want_copy=*/false)) return -1;
  const int cur = a.x;
  // norm_out + swish, conv_out (C -> 2*zc, fp32 out, feature dim padded to 32), quant_conv 1x1 -> NCHW moments
  __half* t = buf[(cur + 1) & 3];
  RGM_CUDA_OK(launch_gn_apply(buf[cur], c.ab(), t, nt, H * H, C, 1, st));
  float* h32 = static_cast<float*>(L->attn_s.p);
  {
    const Conv& cv = m->e_cout;
    GemmDesc d;
    d.A = t;
    d.n_img = nt;
    d.H = H;
    d.W = H;
    d.C = cv.cin;
    d.lda = cv.cin;
    d.B = cv.w;
    d.N = cv.cout_pad;
    d.rows_b = cv.cout_pad;
    d.conv = CONV_3x3;
    d.epi = EPI_F32;
    d.block_n = 32;
    d.e.out = h32;
    d.e.ldo = cv.cout_pad;
    d.e.bias = cv.b;
    d.e.alpha = 1.f;
    RGM_VGEMM_OK(d);
  }
  const int zz = 2 * m->zc;
  RGM_CUDA_OK(launch_vae_quant(h32, m->e_cout.cout_pad, m->q_w, m->q_b, moments + (long long)t0 * zz * H * H, nt, H * H,
                               zz, st));
  return 0;
}

// largest activation of either network per tile, in elements: 128x128 pixels x (channels of level 1)
long long vae_max_act(const Vae* m) {
  long long maxc = 0;
  for (int l = 0; l < m->n_levels; ++l) {
    const long long hw = (long long)(16 << (m->n_levels - 1 - l)) * (16 << (m->n_levels - 1 - l));
    const long long cmax = (long long)m->ch * m->mult[l < m->n_levels - 1 ? l + 1 : l];
    maxc = std::max(maxc, hw * std::max(cmax, (long long)m->ch * m->mult[l]));
  }
  return std::max(maxc, 4LL * 256 * m->block_in0);  // the attention block carves 3-4 tensors out of one buffer
}

// Size the activation / statistics buffers of `lanes` lanes for chunks of `chunk` tiles.  Buffers only grow, and a
// grown buffer's predecessor stays allocated (GrowBuf), so graphs captured earlier stay valid.
int vae_reserve(Vae* m, int chunk, int lanes, cudaStream_t st) {
  const long long maxc = vae_max_act(m);
  for (int l = 0; l < lanes; ++l) {
    Vae::Lane& L = m->lane[l];
    for (int i = 0; i < 4; ++i) RGM_CUDA_OK(L.act[i].reserve((size_t)chunk * maxc * sizeof(__half), st));
    RGM_CUDA_OK(L.gnpart.reserve((size_t)chunk * maxc / 8 + 1024, st));
    for (int i = 0; i < 2; ++i) RGM_CUDA_OK(L.abbuf[i].reserve((size_t)chunk * 512 * sizeof(float2) * 2, st));
    RGM_CUDA_OK(L.attn_s.reserve((size_t)chunk * 256 * 256 * sizeof(float), st));
    RGM_CUDA_OK(L.gncount.reserve(gn_scratch_bytes(chunk) + 1024, st));
  }
  return 0;
}

}  // namespace
}  // namespace rgm

using namespace rgm;

extern "C" {

int rgm_vae_create(rgm_vae** out, int ch, const int* ch_mult, int n_levels, int num_res_blocks, int z_channels,
                   int out_ch) {
  if (rgm_check_device()) return -1;
  if (!out || !ch_mult || n_levels < 1 || n_levels > 8) return set_error("rgm_vae_create: bad arguments");
  if (z_channels != 4) return set_error("rgm_vae_create: the fused stem is written for z_channels = 4");
  if (out_ch != 3) return set_error("rgm_vae_create: out_ch must be 3 (piano roll, onset, pedal: the fused conv_out kernel is written for it)");
  Vae* m = new Vae();
  m->ch = ch;
  m->n_levels = n_levels;
  m->nres = num_res_blocks;
  m->zc = z_channels;
  m->out_ch = out_ch;
  m->in_ch = out_ch;  // the autoencoder's input and output are the same 3-channel piano roll (f8-all-onset.yaml:9-10)
  m->mult.assign(ch_mult, ch_mult + n_levels);
  for (int i = 0; i < n_levels; ++i)
    if ((ch * ch_mult[i]) % 128 != 0) {
      delete m;
      return set_error("rgm_vae_create: every level's channel count must be a multiple of 128");
    }
  if ((16 << (n_levels - 1)) != 128) {
    delete m;
    return set_error("rgm_vae_create: this build decodes 16x16 latent tiles to 128x128 (4 levels)");
  }
  if (const char* e = getenv("RGM_VAE_CHUNK")) m->chunk_tiles = atoi(e) > 0 ? atoi(e) : m->chunk_tiles;
  if (const char* e = getenv("RGM_VAE_LANES")) m->n_lanes = atoi(e) >= 2 ? 2 : 1;
  if (const char* e = getenv("RGM_GN_EPI")) m->gn_epi = atoi(e) != 0;
  if (const char* e = getenv("RGM_GN_DUAL")) m->gn_dual = atoi(e) != 0;
  if (const char* e = getenv("RGM_GN_DUAL_UP")) m->gn_dual_up = atoi(e) != 0;
  if (const char* e = getenv("RGM_GN_DUAL_MIN")) m->gn_dual_min = atoi(e);
  if (vae_build(m) != 0) {
    delete m;
    return -1;
  }
  *out = reinterpret_cast<rgm_vae*>(m);
  return 0;
}

int rgm_vae_set_lanes(rgm_vae* h, int lanes) {
  if (!h) return set_error("rgm_vae_set_lanes: null handle");
  reinterpret_cast<Vae*>(h)->n_lanes = lanes >= 2 ? 2 : 1;
  return 0;
}

int rgm_vae_reserve(rgm_vae* h, int n_tiles) {
  if (!h) return set_error("rgm_vae_reserve: null handle");
  Vae* m = reinterpret_cast<Vae*>(h);
  if (n_tiles <= 0) return 0;
  const int chunk = m->chunk_tiles < n_tiles ? m->chunk_tiles : n_tiles;
  const int lanes = (m->n_lanes > 1 && n_tiles > chunk) ? 2 : 1;
  return vae_reserve(m, chunk, lanes, nullptr);
}

int rgm_vae_gn_timeouts(rgm_vae* h) {
  if (!h) return set_error("rgm_vae_gn_timeouts: null handle");
  Vae* m = reinterpret_cast<Vae*>(h);
  if (cudaDeviceSynchronize() != cudaSuccess) return set_error("rgm_vae_gn_timeouts: device error");
  return *static_cast<volatile int*>(m->gn_err_h);
}

int rgm_vae_destroy(rgm_vae* h) {
  if (h) {
    cudaDeviceSynchronize();
    delete reinterpret_cast<Vae*>(h);
  }
  return 0;
}

int rgm_vae_load(rgm_vae* h, const char* key, const float* src, long long numel, void* stream) {
  if (!h || !key || !src) return set_error("rgm_vae_load: null argument");
  Vae* m = reinterpret_cast<Vae*>(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const std::string k(key);
  auto f = m->f32_keys.find(k);
  if (f != m->f32_keys.end()) {
    if (numel != f->second.second)
      return set_error("rgm_vae_load: " + k + ": expected " + std::to_string(f->second.second) + " elements");
    RGM_CUDA_OK(cudaMemcpyAsync(f->second.first, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (k == "decoder.conv_out.weight")
      RGM_CUDA_OK(launch_vae_out_pack(m->cout_w, m->cout_bfrag, m->cout_cin, m->out_ch, st));
    return 0;
  }
  auto cv = m->conv_keys.find(k);
  if (cv == m->conv_keys.end()) return 1;  // encoder / loss / quant_conv tensors: not on this path
  const Conv& c = *cv->second;
  if (numel != (long long)c.cout * c.cin * c.k * c.k)
    return set_error("rgm_vae_load: " + k + ": unexpected element count");
  return check_cuda(launch_pack_conv_weight(src, c.w, c.cout, c.cin, c.cout_pad, c.cin, c.kind, st), "rgm_vae_load");
}

int rgm_vae_encode(rgm_vae* h, const float* x, float* moments, int n, void* stream) {
  if (rgm_check_device()) return -1;
  if (!h || !x || !moments) return set_error("rgm_vae_encode: null argument");
  Vae* m = reinterpret_cast<Vae*>(h);
  if (n <= 0) return 0;
  if (*static_cast<volatile int*>(m->gn_err_h))
    return set_error("rgm_vae_encode: an earlier call's in-epilogue GroupNorm failed (a wait gave up, or statistics left "
                     "the fixed-point range: rgm_vae_gn_timeouts); its results are invalid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunk = m->chunk_tiles < n ? m->chunk_tiles : n;
  if (vae_reserve(m, chunk, 1, st) != 0) return -1;
  Vae::Lane& L = m->lane[0];
  for (int t0 = 0; t0 < n; t0 += chunk) {
    const int nt = (n - t0) < chunk ? (n - t0) : chunk;
    if (encode_chunk(m, &L, x, moments, t0, nt, st) != 0) return -1;
  }
  return 0;
}

int rgm_vae_decode_latents(rgm_vae* h, const float* lat, float scale_factor, float* roll, int n_cand, int Hlat,
                           int roll_ch, void* stream) {
  if (rgm_check_device()) return -1;
  if (!h || !lat || !roll) return set_error("rgm_vae_decode_latents: null argument");
  Vae* m = reinterpret_cast<Vae*>(h);
  if (n_cand <= 0) return 0;
  if (Hlat % 16 != 0 || Hlat <= 0) return set_error("rgm_vae_decode_latents: latent length must be a multiple of 16");
  if (roll_ch < 1 || roll_ch > m->out_ch) return set_error("rgm_vae_decode_latents: roll_ch out of range");
  if (*static_cast<volatile int*>(m->gn_err_h))
    return set_error("rgm_vae_decode_latents: an earlier call's in-epilogue GroupNorm failed (a wait gave up, or "
                     "statistics left the fixed-point range: rgm_vae_gn_timeouts); its results are invalid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int total = n_cand * (Hlat / 16);
  const int chunk = m->chunk_tiles < total ? m->chunk_tiles : total;
  const int n_chunks = (total + chunk - 1) / chunk;
  const int lanes = (m->n_lanes > 1 && n_chunks > 1) ? 2 : 1;
  if (vae_reserve(m, chunk, lanes, st) != 0) return -1;
  if (lanes == 1) {
    for (int t0 = 0; t0 < total; t0 += chunk) {
      const int nt = (total - t0) < chunk ? (total - t0) : chunk;
      if (decode_chunk(m, &m->lane[0], lat, scale_factor, roll, n_cand, Hlat, roll_ch, t0, nt, st) != 0) return -1;
    }
    return 0;
  }
  // fork: both lane streams start after everything already enqueued on the caller's stream ...
  if (!m->fork) RGM_CUDA_OK(cudaEventCreateWithFlags(&m->fork, cudaEventDisableTiming));
  for (int l = 0; l < 2; ++l) {
    if (!m->lane[l].stream) RGM_CUDA_OK(cudaStreamCreateWithFlags(&m->lane[l].stream, cudaStreamNonBlocking));
    if (!m->lane[l].done) RGM_CUDA_OK(cudaEventCreateWithFlags(&m->lane[l].done, cudaEventDisableTiming));
  }
  RGM_CUDA_OK(cudaEventRecord(m->fork, st));
  for (int l = 0; l < 2; ++l) RGM_CUDA_OK(cudaStreamWaitEvent(m->lane[l].stream, m->fork, 0));
  int ci = 0;
  for (int t0 = 0; t0 < total; t0 += chunk, ++ci) {
    const int nt = (total - t0) < chunk ? (total - t0) : chunk;
    Vae::Lane& L = m->lane[ci & 1];
    if (decode_chunk(m, &L, lat, scale_factor, roll, n_cand, Hlat, roll_ch, t0, nt, L.stream) != 0) return -1;
  }
  // ... and join: the caller's stream continues after both lanes
  for (int l = 0; l < 2; ++l) {
    RGM_CUDA_OK(cudaEventRecord(m->lane[l].done, m->lane[l].stream));
    RGM_CUDA_OK(cudaStreamWaitEvent(st, m->lane[l].done, 0));
  }
  return 0;
}

}  // extern "C"
