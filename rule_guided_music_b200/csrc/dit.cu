// DiTRotary forward on sm_100a (reference guided_diffusion/dit.py:538-634): the handle owns fp16-packed weights and
// a workspace; rgm_dit_forward enqueues, per chunk of samples,
//   patchify gather -> 2 GEMMs (token MLP)                                   dit.py:217-227
//   timestep embedding -> 2 GEMMs (+ label row, SiLU) -> ONE adaLN GEMM for all 28 blocks + final layer
//                                                                            dit.py:58-70, 95-100, 327-333, 367-373
//   per block: LN+modulate -> QKV GEMM (bias + RoPE + head split in the epilogue) -> tcgen05 attention ->
//              proj GEMM (bias, gate, residual in the epilogue) -> LN+modulate -> fc1 GEMM (GELU-tanh) ->
//              fc2 GEMM (bias, gate, residual)                               dit.py:332-336, 263-288
//   LN+modulate -> final GEMM with the unpatchify scatter to NCHW            dit.py:372-376, 613-616
// Residual stream, LayerNorm statistics, softmax and all accumulators are fp32; GEMM operands are fp16.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rgm_b200.h"
#include "api_util.h"
#include "aux_kernels.h"
#include "gemm_host.h"

namespace rgm {

__global__ void convert_pad_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long rows, int csrc,
                                   int cdst) {
  const long long total = rows * cdst;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cdst);
    const long long r = i / cdst;
    dst[i] = __float2half_rn(c < csrc ? src[r * csrc + c] : 0.f);
  }
}

static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

using Workspace = GrowBuf;  // api_util.h: grows by retiring, never frees what a captured graph may reference

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(long long n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += (size_t)n * sizeof(T);
    return p;
  }
};

struct Dit {
  int depth, D, heads, dh, P, C, Cout, label_rows, W, mlp, freq_dim = 256;
  int K0, nfin, nada, rot, tpt;
  // fp16 weights
  __half *w_x0, *w_x2, *w_t0, *w_t2, *w_ada, *w_fin;
  std::vector<__half*> w_qkv, w_proj, w_fc1, w_fc2;
  // fp32
  float *b_x0, *b_x2, *b_t0, *b_t2, *b_ada, *b_fin, *ytab, *rope_freqs, *t_freqs;
  std::vector<float*> b_qkv, b_proj, b_fc1, b_fc2;
  void* w16_arena = nullptr;
  void* f32_arena = nullptr;
  // rotary tables for the last T used
  GrowBuf rope_buf;            // (cos, sin) [rope_T, rot/2]
  float2* rope_cs = nullptr;   // = rope_buf.p
  int rope_T = 0;
  bool rope_dirty = true;
  // Two chunk pipelines ("lanes", like the VAE decoder's): consecutive sample chunks alternate between two workspaces
  // and two streams, so one chunk's latency- and HBM-bound kernels (attention, LayerNorm, the residual epilogues) run
  // while the other chunk's GEMMs hold the tensor pipe.
  Workspace ws[2];
  cudaStream_t lane_stream[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, lane_done[2] = {nullptr, nullptr};
  int n_lanes = 2;
  int chunk = 256;

  ~Dit() {
    if (w16_arena) cudaFree(w16_arena);
    if (f32_arena) cudaFree(f32_arena);
    for (int l = 0; l < 2; ++l) {
      if (lane_stream[l]) cudaStreamDestroy(lane_stream[l]);
      if (lane_done[l]) cudaEventDestroy(lane_done[l]);
    }
    if (fork) cudaEventDestroy(fork);
  }
};

static int dit_alloc(Dit* m) {
  const int D = m->D;
  m->dh = D / m->heads;
  m->rot = (int)(m->dh * 0.5);  // dit.py:571 rotary_dim = int(hidden // heads * 0.5)
  m->tpt = m->W / m->P;
  m->K0 = (int)round_up((long long)m->C * m->P, 64);
  m->nfin = (int)round_up((long long)m->P * m->Cout, 32);
  m->nada = m->depth * 6 * D + 2 * D;
  for (int pass = 0; pass < 2; ++pass) {
    Carver h(pass ? m->w16_arena : nullptr), f(pass ? m->f32_arena : nullptr);
    m->w_x0 = h.take<__half>(256LL * m->K0);
    m->w_x2 = h.take<__half>((long long)D * 256);
    m->w_t0 = h.take<__half>((long long)D * m->freq_dim);
    m->w_t2 = h.take<__half>((long long)D * D);
    m->w_ada = h.take<__half>((long long)m->nada * D);
    m->w_fin = h.take<__half>((long long)m->nfin * D);
    m->w_qkv.resize(m->depth);
    m->w_proj.resize(m->depth);
    m->w_fc1.resize(m->depth);
    m->w_fc2.resize(m->depth);
    m->b_qkv.resize(m->depth);
    m->b_proj.resize(m->depth);
    m->b_fc1.resize(m->depth);
    m->b_fc2.resize(m->depth);
    for (int i = 0; i < m->depth; ++i) {
      m->w_qkv[i] = h.take<__half>(3LL * D * D);
      m->w_proj[i] = h.take<__half>((long long)D * D);
      m->w_fc1[i] = h.take<__half>((long long)m->mlp * D);
      m->w_fc2[i] = h.take<__half>((long long)D * m->mlp);
      m->b_qkv[i] = f.take<float>(3LL * D);
      m->b_proj[i] = f.take<float>(D);
      m->b_fc1[i] = f.take<float>(m->mlp);
      m->b_fc2[i] = f.take<float>(D);
    }
    m->b_x0 = f.take<float>(256);
    m->b_x2 = f.take<float>(D);
    m->b_t0 = f.take<float>(D);
    m->b_t2 = f.take<float>(D);
    m->b_ada = f.take<float>(m->nada);
    m->b_fin = f.take<float>(m->nfin);
    m->ytab = f.take<float>((long long)(m->label_rows > 0 ? m->label_rows : 1) * D);
    m->rope_freqs = f.take<float>(m->rot / 2 > 0 ? m->rot / 2 : 1);
    m->t_freqs = f.take<float>(m->freq_dim / 2);
    if (pass == 0) {
      RGM_CUDA_OK(cudaMalloc(&m->w16_arena, h.off + 256));
      RGM_CUDA_OK(cudaMalloc(&m->f32_arena, f.off + 256));
      RGM_CUDA_OK(cudaMemset(m->w16_arena, 0, h.off + 256));
      RGM_CUDA_OK(cudaMemset(m->f32_arena, 0, f.off + 256));
    }
  }
  return 0;
}

struct LoadTarget {
  __half* h = nullptr;
  float* f = nullptr;
  long long rows = 0;
  int csrc = 0, cdst = 0;  // fp16 targets: [rows, csrc] -> [rows, cdst]
  long long numel = 0;
};

static bool dit_resolve(Dit* m, const std::string& key, LoadTarget* t) {
  const int D = m->D;
  auto W = [&](__half* p, long long rows, int cols, int cdst = 0) {
    t->h = p;
    t->rows = rows;
    t->csrc = cols;
    t->cdst = cdst ? cdst : cols;
    t->numel = rows * cols;
    return true;
  };
  auto F = [&](float* p, long long n) {
    t->f = p;
    t->numel = n;
    return true;
  };
  if (key == "x_embedder.MLP.0.weight") return W(m->w_x0, 256, m->C * m->P, m->K0);
  if (key == "x_embedder.MLP.0.bias") return F(m->b_x0, 256);
  if (key == "x_embedder.MLP.2.weight") return W(m->w_x2, D, 256);
  if (key == "x_embedder.MLP.2.bias") return F(m->b_x2, D);
  if (key == "t_embedder.mlp.0.weight") return W(m->w_t0, D, m->freq_dim);
  if (key == "t_embedder.mlp.0.bias") return F(m->b_t0, D);
  if (key == "t_embedder.mlp.2.weight") return W(m->w_t2, D, D);
  if (key == "t_embedder.mlp.2.bias") return F(m->b_t2, D);
  if (key == "y_embedder.embedding_table.weight") return m->label_rows > 0 && F(m->ytab, (long long)m->label_rows * D);
  if (key == "rotary_emb.freqs") {
    m->rope_dirty = true;
    return F(m->rope_freqs, m->rot / 2);
  }
  if (key == "__timestep_freqs") return F(m->t_freqs, m->freq_dim / 2);
  if (key == "final_layer.linear.weight") return W(m->w_fin, m->P * m->Cout, D);
  if (key == "final_layer.linear.bias") return F(m->b_fin, m->P * m->Cout);
  if (key == "final_layer.adaLN_modulation.1.weight") return W(m->w_ada + (long long)m->depth * 6 * D * D, 2 * D, D);
  if (key == "final_layer.adaLN_modulation.1.bias") return F(m->b_ada + (long long)m->depth * 6 * D, 2 * D);
  if (key.rfind("blocks.", 0) == 0) {
    const size_t dot = key.find('.', 7);
    if (dot == std::string::npos) return false;
    const int i = atoi(key.substr(7, dot - 7).c_str());
    if (i < 0 || i >= m->depth) return false;
    const std::string rest = key.substr(dot + 1);
    if (rest == "attn.qkv.weight") return W(m->w_qkv[i], 3LL * D, D);
    if (rest == "attn.qkv.bias") return F(m->b_qkv[i], 3LL * D);
    if (rest == "attn.proj.weight") return W(m->w_proj[i], D, D);
    if (rest == "attn.proj.bias") return F(m->b_proj[i], D);
    if (rest == "mlp.fc1.weight") return W(m->w_fc1[i], m->mlp, D);
    if (rest == "mlp.fc1.bias") return F(m->b_fc1[i], m->mlp);
    if (rest == "mlp.fc2.weight") return W(m->w_fc2[i], D, m->mlp);
    if (rest == "mlp.fc2.bias") return F(m->b_fc2[i], D);
    if (rest == "adaLN_modulation.1.weight") return W(m->w_ada + (long long)i * 6 * D * D, 6LL * D, D);
    if (rest == "adaLN_modulation.1.bias") return F(m->b_ada + (long long)i * 6 * D, 6LL * D);
  }
  return false;
}

#define RGM_GEMM_OK(desc)                                                                  \
  do {                                                                                     \
    std::string _err;                                                                      \
    if (launch_gemm(desc, st, &_err) != cudaSuccess) return set_error("rgm_dit: " + _err); \
  } while (0)

static GemmDesc linear_desc(const __half* A, long long M, long long a_rows, int K, const __half* Bw, int N, int epi) {
  GemmDesc d;
  d.A = A;
  d.n_img = 1;
  d.H = 1;
  d.W = (int)M;
  d.a_rows = a_rows;
  d.C = K;
  d.lda = K;
  d.B = Bw;
  d.rows_b = N;
  d.N = N;
  d.conv = CONV_1x1;
  d.epi = epi;
  d.e.alpha = 1.f;
  return d;
}

static int dit_forward_chunk(Dit* m, void* workspace, const float* x, const float* t, const long long* y, float* out,
                             int B, int H, cudaStream_t st) {
  const int D = m->D, T = H * m->tpt;
  const long long M = (long long)B * T;
  const long long Mp = round_up(M, 256), Bp = round_up(B, 256);
  Carver c(workspace);
  __half* tok16 = c.take<__half>(Mp * m->K0);
  __half* h0_16 = c.take<__half>(Mp * 256);
  float* x32 = c.take<float>(Mp * D);
  __half* a16 = c.take<__half>(Mp * D);
  __half* q16 = c.take<__half>(Mp * D);
  __half* k16 = c.take<__half>(Mp * D);
  __half* vt16 = c.take<__half>(Mp * D + 128LL * T);  // the last head's V^T box reads up to 16 rows past the end
  __half* o16 = c.take<__half>(Mp * D);
  __half* h16 = c.take<__half>(Mp * m->mlp);
  __half* emb16 = c.take<__half>(Bp * m->freq_dim);
  __half* c1_16 = c.take<__half>(Bp * D);
  __half* sc16 = c.take<__half>(Bp * D);
  float* mod32 = c.take<float>((long long)B * m->nada);

  // token MLP
  RGM_CUDA_OK(launch_patchify(x, tok16, B, m->C, H, m->W, m->P, m->K0, st));
  {
    GemmDesc d = linear_desc(tok16, M, Mp, m->K0, m->w_x0, 256, EPI_F16);
    d.e.out = h0_16;
    d.e.ldo = 256;
    d.e.bias = m->b_x0;
    d.e.act = ACT_SILU;
    RGM_GEMM_OK(d);
  }
  {
    GemmDesc d = linear_desc(h0_16, M, Mp, 256, m->w_x2, D, EPI_F32);
    d.e.out = x32;
    d.e.ldo = D;
    d.e.bias = m->b_x2;
    RGM_GEMM_OK(d);
  }
  // conditioning vector c = t_embedder(t) + y_embedder(y); every consumer applies SiLU first, so keep SiLU(c)
  RGM_CUDA_OK(launch_timestep_embedding(t, m->t_freqs, emb16, B, m->freq_dim / 2, m->freq_dim, st));
  {
    GemmDesc d = linear_desc(emb16, B, Bp, m->freq_dim, m->w_t0, D, EPI_F16);
    d.e.out = c1_16;
    d.e.ldo = D;
    d.e.bias = m->b_t0;
    d.e.act = ACT_SILU;
    RGM_GEMM_OK(d);
  }
  {
    GemmDesc d = linear_desc(c1_16, B, Bp, D, m->w_t2, D, EPI_F16);
    d.e.out = sc16;
    d.e.ldo = D;
    d.e.bias = m->b_t2;
    d.e.act = ACT_SILU;
    if (y != nullptr && m->label_rows > 0) {
      d.e.addtab = m->ytab;
      d.e.addidx = y;
    }
    RGM_GEMM_OK(d);
  }
  {
    GemmDesc d = linear_desc(sc16, B, Bp, D, m->w_ada, m->nada, EPI_F32);
    d.e.out = mod32;
    d.e.ldo = m->nada;
    d.e.bias = m->b_ada;
    RGM_GEMM_OK(d);
  }
  const float attn_scale = 1.0f / sqrtf((float)m->dh);
  for (int i = 0; i < m->depth; ++i) {
    const float* mod = mod32 + (long long)i * 6 * D;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    RGM_CUDA_OK(launch_ln_modulate(x32, mod, mod + D, m->nada, a16, M, D, T, 1e-6f, st));
    {
      GemmDesc d = linear_desc(a16, M, Mp, D, m->w_qkv[i], 3 * D, EPI_QKV_ROPE);
      d.e.bias = m->b_qkv[i];
      d.e.q = q16;
      d.e.k = k16;
      d.e.v = vt16;
      d.e.rope_cs = m->rope_cs;
      d.e.T = T;
      d.e.heads = m->heads;
      d.e.dh = m->dh;
      d.e.dh_pad = m->dh;
      d.e.rot_dim = m->rot;
      RGM_GEMM_OK(d);
    }
    {
      std::string err;
      if (launch_attention(q16, k16, vt16, o16, B, m->heads, T, m->dh, attn_scale, st, &err) != cudaSuccess)
        return set_error("rgm_dit: " + err);
    }
    {
      GemmDesc d = linear_desc(o16, M, Mp, D, m->w_proj[i], D, EPI_GATE_RESID);
      d.e.out = x32;
      d.e.ldo = D;
      d.e.bias = m->b_proj[i];
      d.e.gate = mod + 2 * D;
      d.e.gate_ld = m->nada;
      d.e.rows_per_sample = T;
      RGM_GEMM_OK(d);
    }
    RGM_CUDA_OK(launch_ln_modulate(x32, mod + 3 * D, mod + 4 * D, m->nada, a16, M, D, T, 1e-6f, st));
    {
      GemmDesc d = linear_desc(a16, M, Mp, D, m->w_fc1[i], m->mlp, EPI_F16);
      d.e.out = h16;
      d.e.ldo = m->mlp;
      d.e.bias = m->b_fc1[i];
      d.e.act = ACT_GELU_TANH;
      RGM_GEMM_OK(d);
    }
    {
      GemmDesc d = linear_desc(h16, M, Mp, m->mlp, m->w_fc2[i], D, EPI_GATE_RESID);
      d.e.out = x32;
      d.e.ldo = D;
      d.e.bias = m->b_fc2[i];
      d.e.gate = mod + 5 * D;
      d.e.gate_ld = m->nada;
      d.e.rows_per_sample = T;
      RGM_GEMM_OK(d);
    }
  }
  {
    const float* mod = mod32 + (long long)m->depth * 6 * D;  // shift, scale
    RGM_CUDA_OK(launch_ln_modulate(x32, mod, mod + D, m->nada, a16, M, D, T, 1e-6f, st));
    GemmDesc d = linear_desc(a16, M, Mp, D, m->w_fin, m->nfin, EPI_UNPATCH);
    d.block_n = 32;
    d.e.out = out;
    d.e.bias = m->b_fin;
    d.e.tpt = m->tpt;
    d.e.c_out = m->Cout;
    d.e.latH = H;
    d.e.latW = m->W;
    d.e.n_valid = m->P * m->Cout;
    RGM_GEMM_OK(d);
  }
  return 0;
}

static size_t dit_workspace_bytes(const Dit* m, int B, int T) {
  const long long M = (long long)B * T, Mp = round_up(M, 256), Bp = round_up(B, 256);
  Carver c(nullptr);
  c.take<__half>(Mp * m->K0);
  c.take<__half>(Mp * 256);
  c.take<float>(Mp * m->D);
  for (int i = 0; i < 3; ++i) c.take<__half>(Mp * m->D);
  c.take<__half>(Mp * m->D + 128LL * T);
  c.take<__half>(Mp * m->D);
  c.take<__half>(Mp * m->mlp);
  c.take<__half>(Bp * m->freq_dim);
  c.take<__half>(Bp * m->D);
  c.take<__half>(Bp * m->D);
  c.take<float>((long long)B * m->nada);
  return c.off + 4096;
}

// Rotary table for T tokens and workspaces for a batch of B samples (chunked): everything rgm_dit_forward allocates.
static int dit_prepare(Dit* m, int B, int T, cudaStream_t st) {
  if (m->rope_dirty || m->rope_T < T) {
    if (m->rope_T < T) {
      const int nf = m->rot / 2 > 0 ? m->rot / 2 : 1;
      RGM_CUDA_OK(m->rope_buf.reserve((size_t)T * nf * sizeof(float2), st));
      m->rope_cs = static_cast<float2*>(m->rope_buf.p);
      m->rope_T = T;
    }
    if (m->rot > 0) RGM_CUDA_OK(launch_rope_table(m->rope_freqs, m->rope_cs, m->rope_T, m->rot / 2, st));
    m->rope_dirty = false;
  }
  const int chunk = m->chunk < B ? m->chunk : B;
  const int n_chunks = (B + chunk - 1) / chunk;
  const int lanes = (m->n_lanes > 1 && n_chunks > 1) ? 2 : 1;
  for (int l = 0; l < lanes; ++l) RGM_CUDA_OK(m->ws[l].reserve(dit_workspace_bytes(m, chunk, T), st));
  return 0;
}

}  // namespace rgm

using namespace rgm;

extern "C" {

int rgm_dit_create(rgm_dit** out, int depth, int hidden, int heads, int patch, int in_channels, int out_channels,
                   int label_rows, int latent_w, int mlp_hidden) {
  if (rgm_check_device()) return -1;
  if (!out) return set_error("rgm_dit_create: null out");
  if (hidden % 128 != 0 || heads <= 0 || hidden % heads != 0 || (hidden / heads) % 8 != 0 || hidden / heads > 128)
    return set_error("rgm_dit_create: hidden must be a multiple of 128 and head_dim a multiple of 8, <= 128");
  if (mlp_hidden % 128 != 0 || patch <= 0 || latent_w % patch != 0 || depth <= 0)
    return set_error("rgm_dit_create: bad mlp_hidden / patch / depth");
  Dit* m = new Dit();
  m->depth = depth;
  m->D = hidden;
  m->heads = heads;
  m->P = patch;
  m->C = in_channels;
  m->Cout = out_channels;
  m->label_rows = label_rows;
  m->W = latent_w;
  m->mlp = mlp_hidden;
  if (const char* e = getenv("RGM_DIT_CHUNK")) m->chunk = atoi(e) > 0 ? atoi(e) : m->chunk;
  if (const char* e = getenv("RGM_DIT_LANES")) m->n_lanes = atoi(e) >= 2 ? 2 : 1;
  if (dit_alloc(m) != 0) {
    delete m;
    return -1;
  }
  *out = reinterpret_cast<rgm_dit*>(m);
  return 0;
}

int rgm_dit_set_lanes(rgm_dit* h, int lanes) {
  if (!h) return set_error("rgm_dit_set_lanes: null handle");
  reinterpret_cast<Dit*>(h)->n_lanes = lanes >= 2 ? 2 : 1;
  return 0;
}

int rgm_dit_reserve(rgm_dit* h, int B, int H) {
  if (rgm_check_device()) return -1;
  if (!h) return set_error("rgm_dit_reserve: null handle");
  Dit* m = reinterpret_cast<Dit*>(h);
  if (B <= 0 || H <= 0) return 0;
  return dit_prepare(m, B, H * m->tpt, nullptr);
}

int rgm_dit_destroy(rgm_dit* h) {
  if (h) {
    cudaDeviceSynchronize();
    delete reinterpret_cast<Dit*>(h);
  }
  return 0;
}

int rgm_dit_load(rgm_dit* h, const char* key, const float* src, long long numel, void* stream) {
  if (!h || !key || !src) return set_error("rgm_dit_load: null argument");
  Dit* m = reinterpret_cast<Dit*>(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const std::string k(key);
  if (k.size() > 22 && k.compare(k.size() - 22, 22, ".attn.rotary_emb.freqs") == 0) return 0;  // alias of rotary_emb.freqs
  LoadTarget t;
  if (!dit_resolve(m, k, &t)) return 1;  // not a tensor of this path: ignored, like load_state_dict(strict=False)
  if (numel != t.numel)
    return set_error("rgm_dit_load: " + k + ": expected " + std::to_string(t.numel) + " elements, got " +
                     std::to_string(numel));
  if (t.h) {
    const long long total = t.rows * t.cdst;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    convert_pad_kernel<<<grid, 256, 0, st>>>(src, t.h, t.rows, t.csrc, t.cdst);
    g_aux_launches++;
    return check_cuda(cudaGetLastError(), "rgm_dit_load");
  }
  return check_cuda(cudaMemcpyAsync(t.f, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, st),
                    "rgm_dit_load");
}

int rgm_dit_forward(rgm_dit* h, const float* x, const float* t, const long long* y, float* out, int B, int H,
                    void* stream) {
  if (rgm_check_device()) return -1;
  if (!h || !x || !t || !out) return set_error("rgm_dit_forward: null argument");
  Dit* m = reinterpret_cast<Dit*>(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B <= 0) return 0;
  const int T = H * m->tpt;
  if (T < 64 || T > 256 || T % 64 != 0)
    return set_error("rgm_dit_forward: " + std::to_string(T) + " tokens; this build supports 64, 128, 192 or 256 (latent H "
                     "32 ... 128 in steps of 32 at patch 8; the reference's checkpoints are trained at H = 128)");
  if (dit_prepare(m, B, T, st) != 0) return -1;
  const int chunk = m->chunk < B ? m->chunk : B;
  const int n_chunks = (B + chunk - 1) / chunk;
  const int lanes = (m->n_lanes > 1 && n_chunks > 1) ? 2 : 1;
  const long long per_in = (long long)m->C * H * m->W, per_out = (long long)m->Cout * H * m->W;
  if (lanes == 1) {
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int nb = (B - b0) < chunk ? (B - b0) : chunk;
      if (dit_forward_chunk(m, m->ws[0].p, x + b0 * per_in, t + b0, y ? y + b0 : nullptr, out + b0 * per_out, nb, H,
                            st) != 0)
        return -1;
    }
    return 0;
  }
  // fork: both lane streams start after everything already enqueued on the caller's stream (inputs, rope table) ...
  if (!m->fork) RGM_CUDA_OK(cudaEventCreateWithFlags(&m->fork, cudaEventDisableTiming));
  for (int l = 0; l < 2; ++l) {
    if (!m->lane_stream[l]) RGM_CUDA_OK(cudaStreamCreateWithFlags(&m->lane_stream[l], cudaStreamNonBlocking));
    if (!m->lane_done[l]) RGM_CUDA_OK(cudaEventCreateWithFlags(&m->lane_done[l], cudaEventDisableTiming));
  }
  RGM_CUDA_OK(cudaEventRecord(m->fork, st));
  for (int l = 0; l < 2; ++l) RGM_CUDA_OK(cudaStreamWaitEvent(m->lane_stream[l], m->fork, 0));
  int ci = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++ci) {
    const int nb = (B - b0) < chunk ? (B - b0) : chunk;
    const int l = ci & 1;
    if (dit_forward_chunk(m, m->ws[l].p, x + b0 * per_in, t + b0, y ? y + b0 : nullptr, out + b0 * per_out, nb, H,
                          m->lane_stream[l]) != 0)
      return -1;
  }
  // ... and join: the caller's stream continues after both lanes
  for (int l = 0; l < 2; ++l) {
    RGM_CUDA_OK(cudaEventRecord(m->lane_done[l], m->lane_stream[l]));
    RGM_CUDA_OK(cudaStreamWaitEvent(st, m->lane_done[l], 0));
  }
  return 0;
}

}  // extern "C"
