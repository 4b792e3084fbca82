// Memory-bound helper kernels around the tensor-core GEMMs: gathers, normalisation, activation, softmax, transposes
// and the sampler's elementwise updates.  Every kernel reads/writes 16-byte vectors, consecutive lanes touching
// consecutive addresses, with grids sized from the row count (these are all HBM-bound).
#include <atomic>

#include "api_util.h"
#include <cstdio>
#include <cstdlib>

#include "aux_kernels.h"
#include "ptx.cuh"

namespace rgm {

// x * sigmoid(x) in 5 issue slots (ptx.cuh silu_f): the GroupNorm passes are ISSUE-bound, not MUFU- or HBM-bound.
__device__ __forceinline__ float swish_fast(float v) { return swish_vae(v); }

namespace {
std::atomic<unsigned long long> g_launches{0};
inline cudaError_t done() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
inline int blocks_for(long long n, int per_block, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}
}  // namespace

unsigned long long aux_launch_count() { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------------------
// DiT
// ------------------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ x, __half* __restrict__ tok, int B, int C, int H, int W,
                                int P, int kpad) {
  const int tpt = W / P;
  const long long total = (long long)B * H * tpt * kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % kpad);
    const long long row = i / kpad;
    const int part = (int)(row % tpt);
    const long long bt = row / tpt;
    const int time = (int)(bt % H);
    const int b = (int)(bt / H);
    float v = 0.f;
    if (f < C * P) {
      const int pl = f / C, ch = f - pl * C;
      v = x[(((long long)b * C + ch) * H + time) * W + part * P + pl];
    }
    tok[i] = __float2half_rn(v);
  }
}

cudaError_t launch_patchify(const float* x, __half* tok, int B, int C, int H, int W, int P, int kpad,
                            cudaStream_t s) {
  const long long total = (long long)B * H * (W / P) * kpad;
  patchify_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, tok, B, C, H, W, P, kpad);
  return done();
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                          __half* __restrict__ emb, int B, int half, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float a = t[b] * freqs[j];
  emb[(long long)b * ld + j] = __float2half_rn(cosf(a));
  emb[(long long)b * ld + half + j] = __float2half_rn(sinf(a));
}

cudaError_t launch_timestep_embedding(const float* t, const float* freqs, __half* emb, int B, int half, int ld,
                                      cudaStream_t s) {
  timestep_embedding_kernel<<<(B * half + 255) / 256, 256, 0, s>>>(t, freqs, emb, B, half, ld);
  return done();
}

__global__ void rope_table_kernel(const float* __restrict__ freqs, float2* __restrict__ cs, int T, int nfreq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * nfreq) return;
  const int pos = i / nfreq, j = i - pos * nfreq;
  const float a = (float)pos * freqs[j];
  cs[i] = make_float2(cosf(a), sinf(a));
}

cudaError_t launch_rope_table(const float* freqs, float2* cs, int T, int nfreq, cudaStream_t s) {
  rope_table_kernel<<<(T * nfreq + 255) / 256, 256, 0, s>>>(freqs, cs, T, nfreq);
  return done();
}

// One warp per row; the row (D = 128*NV floats) lives in registers between the statistics and the apply pass, so
// HBM sees exactly one fp32 read and one fp16 write per element.
template <int NV>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ shift,
                                                          const float* __restrict__ scale, int mod_ld,
                                                          __half* __restrict__ out, long long rows,
                                                          int rows_per_sample, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xr[i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
    const long long b = row / rows_per_sample;
    const float4* sh = reinterpret_cast<const float4*>(shift + b * mod_ld);
    const float4* sc = reinterpret_cast<const float4*>(scale + b * mod_ld);
    uint2* o2 = reinterpret_cast<uint2*>(out + row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 h = __ldg(sh + i * 32 + lane);
      const float4 c = __ldg(sc + i * 32 + lane);
      const float y0 = (v[i].x - mean) * rstd * (1.f + c.x) + h.x;
      const float y1 = (v[i].y - mean) * rstd * (1.f + c.y) + h.y;
      const float y2 = (v[i].z - mean) * rstd * (1.f + c.z) + h.z;
      const float y3 = (v[i].w - mean) * rstd * (1.f + c.w) + h.w;
      __half2 p0 = __floats2half2_rn(y0, y1), p1 = __floats2half2_rn(y2, y3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      o2[i * 32 + lane] = u;
    }
  }
}

// Two rows per warp.  Rows 2k and 2k+1 are adjacent in memory, so the warp streams 2*D contiguous floats: lane l owns
// the 8-element chunks l, l + 32, ... (NV of them) -- two 16-byte loads in, ONE 16-byte store out per chunk (the one-row
// kernel above stores 8 bytes per lane and keeps half as many loads in flight; it measured 3.6 TB/s = 0.55 of the copy
// peak).  Same arithmetic per element; the two rows' statistics are kept apart by the chunk's row.
template <int NV, int OCC>
__global__ void __launch_bounds__(256, OCC) ln_modulate2_kernel(const float* __restrict__ x, const float* __restrict__ shift,
                                                              const float* __restrict__ scale, int mod_ld,
                                                              __half* __restrict__ out, long long pairs,
                                                              int rows_per_sample, float eps) {
  constexpr int D = NV * 128;
  constexpr int CPR = D / 8;  // chunks per row
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pr = warp0; pr < pairs; pr += nwarps) {
    const long long row0 = 2 * pr;
    const float4* xp = reinterpret_cast<const float4*>(x + row0 * D);
    float4 v[NV][2];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i][0] = xp[(lane + 32 * i) * 2];
      v[i][1] = xp[(lane + 32 * i) * 2 + 1];
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float cs = ((v[i][0].x + v[i][0].y) + (v[i][0].z + v[i][0].w)) + ((v[i][1].x + v[i][1].y) + (v[i][1].z + v[i][1].w));
      if (lane + 32 * i >= CPR) s1 += cs;
      else s0 += cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float m0 = s0 * (1.0f / D), m1 = s1 * (1.0f / D);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const bool r1 = lane + 32 * i >= CPR;
      const float m = r1 ? m1 : m0;
      float cq = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float a = v[i][h].x - m, b = v[i][h].y - m, c = v[i][h].z - m, d = v[i][h].w - m;
        cq += (a * a + b * b) + (c * c + d * d);
      }
      if (r1) q1 += cq;
      else q0 += cq;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      q0 += __shfl_xor_sync(0xffffffffu, q0, o);
      q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    const float rs0 = rsqrtf(q0 * (1.0f / D) + eps), rs1 = rsqrtf(q1 * (1.0f / D) + eps);
    const long long b = row0 / rows_per_sample;  // both rows belong to one sample (rows_per_sample is even)
    const float4* sh = reinterpret_cast<const float4*>(shift + b * mod_ld);
    const float4* sc = reinterpret_cast<const float4*>(scale + b * mod_ld);
    uint4* o4 = reinterpret_cast<uint4*>(out + row0 * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      const bool r1 = c >= CPR;
      const float m = r1 ? m1 : m0, rstd = r1 ? rs1 : rs0;
      const int off = (r1 ? c - CPR : c) * 2;  // float4 index inside the row
      uint4 u;
      uint32_t* uw = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 hh = __ldg(sh + off + h);
        const float4 cc = __ldg(sc + off + h);
        const float y0 = (v[i][h].x - m) * rstd * (1.f + cc.x) + hh.x;
        const float y1 = (v[i][h].y - m) * rstd * (1.f + cc.y) + hh.y;
        const float y2 = (v[i][h].z - m) * rstd * (1.f + cc.z) + hh.z;
        const float y3 = (v[i][h].w - m) * rstd * (1.f + cc.w) + hh.w;
        __half2 p0 = __floats2half2_rn(y0, y1), p1 = __floats2half2_rn(y2, y3);
        uw[2 * h] = *reinterpret_cast<uint32_t*>(&p0);
        uw[2 * h + 1] = *reinterpret_cast<uint32_t*>(&p1);
      }
      o4[c] = u;
    }
  }
}

cudaError_t launch_ln_modulate(const float* x, const float* shift, const float* scale, int mod_ld, __half* out,
                               long long rows, int D, int rows_per_sample, float eps, cudaStream_t s) {
  if (D % 128 != 0) return cudaErrorInvalidValue;
  ProfScope prof("ln_modulate", 0, 0, (double)rows * D * 6.0, s);
  static const int one_row = [] {
    const char* e = getenv("RGM_LN_ONE_ROW");  // development knob: the one-row-per-warp kernel
    return e ? atoi(e) : 0;
  }();
  static const int occ2 = [] {
    const char* e = getenv("RGM_LN_OCC");  // 2 (default): 128 registers, two blocks per SM, 48 B of spills: 4.0 TB/s;
    return e ? atoi(e) == 2 : 1;           // 1: 246 registers, one block per SM: 3.6 TB/s (B200, config 2)
  }();
  if (!one_row && rows % 2 == 0 && rows_per_sample % 2 == 0 && mod_ld % 4 == 0) {
    const long long pairs = rows / 2;
    const int grid = blocks_for(pairs, 8, 148 * 8);
#define RGM_LN2(NV)                                                                                                   \
  case NV:                                                                                                            \
    if (occ2) ln_modulate2_kernel<NV, 2><<<grid, 256, 0, s>>>(x, shift, scale, mod_ld, out, pairs, rows_per_sample, eps); \
    else ln_modulate2_kernel<NV, 1><<<grid, 256, 0, s>>>(x, shift, scale, mod_ld, out, pairs, rows_per_sample, eps);      \
    break;
    switch (D / 128) {
      RGM_LN2(2) RGM_LN2(3) RGM_LN2(4) RGM_LN2(6) RGM_LN2(8) RGM_LN2(9)
      default:
        return cudaErrorInvalidValue;
    }
#undef RGM_LN2
    return done();
  }
  const int grid = blocks_for(rows, 8, 148 * 8);
#define RGM_LN(NV)                                                                                                   \
  case NV:                                                                                                           \
    ln_modulate_kernel<NV><<<grid, 256, 0, s>>>(x, shift, scale, mod_ld, out, rows, rows_per_sample, eps);           \
    break;
  switch (D / 128) {
    RGM_LN(2) RGM_LN(3) RGM_LN(4) RGM_LN(6) RGM_LN(8) RGM_LN(9)
    default:
      return cudaErrorInvalidValue;
  }
#undef RGM_LN
  return done();
}

// ------------------------------------------------------------------------------------------------------------
// VAE decoder
// ------------------------------------------------------------------------------------------------------------
// One block per VAE tile.  z tile pixel (row = latent pitch p, col = latent time tau) = lat[cand, c, k*16+tau, p]
// (gaussian_diffusion.py:1350-1353: permute(0,1,3,2), chunk along time, concatenate on batch).
__global__ void __launch_bounds__(256) vae_stem_kernel(const float* __restrict__ lat, float scale,
                                                       const float* __restrict__ pq_w, const float* __restrict__ pq_b,
                                                       const float* __restrict__ cin_w, const float* __restrict__ cin_b,
                                                       __half* __restrict__ out, int n_cand, int Hlat, int tile0,
                                                       int Cout) {
  __shared__ float zq[18][18][4];  // post_quant output with the zero halo conv_in pads with
  __shared__ float wsm[64][36];    // conv_in weights of the 64 output channels being produced
  __shared__ float bsm[64];
  const int g = tile0 + blockIdx.x;
  const int kt = g / n_cand, cand = g - kt * n_cand;
  for (int i = threadIdx.x; i < 18 * 18 * 4; i += blockDim.x) (&zq[0][0][0])[i] = 0.f;
  __syncthreads();
  {
    const int pix = threadIdx.x;  // 256 pixels
    const int r = pix >> 4, c = pix & 15;
    float z[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      z[ch] = lat[(((long long)cand * 4 + ch) * Hlat + kt * 16 + c) * 16 + r] / scale;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float a = pq_b[o];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) a += pq_w[o * 4 + ch] * z[ch];
      zq[r + 1][c + 1][o] = a;
    }
  }
  __syncthreads();
  const int pix = threadIdx.x;
  const int r = pix >> 4, c = pix & 15;
  float in[36];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) in[ch * 9 + ky * 3 + kx] = zq[r + ky][c + kx][ch];
  __half* orow = out + ((long long)blockIdx.x * 256 + pix) * Cout;
  for (int co0 = 0; co0 < Cout; co0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) wsm[i / 36][i % 36] = cin_w[(long long)co0 * 36 + i];
    if (threadIdx.x < 64) bsm[threadIdx.x] = cin_b[co0 + threadIdx.x];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < 64; j += 8) {
      uint4 pk;
      __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        float a0 = bsm[j + u], a1 = bsm[j + u + 1];
#pragma unroll
        for (int q = 0; q < 36; ++q) {
          a0 += wsm[j + u][q] * in[q];
          a1 += wsm[j + u + 1][q] * in[q];
        }
        h2[u >> 1] = __floats2half2_rn(a0, a1);
      }
      *reinterpret_cast<uint4*>(orow + co0 + j) = pk;
    }
  }
}

cudaError_t launch_vae_stem(const float* lat, float scale, const float* pq_w, const float* pq_b,
                            const float* cin_w, const float* cin_b, __half* out, int n_cand, int Hlat, int tile0,
                            int n_tiles, int Cout, cudaStream_t s) {
  if (Cout % 64 != 0 || n_tiles <= 0) return cudaErrorInvalidValue;
  ProfScope prof("vae_stem", 0, 0, (double)n_tiles * 256.0 * (Cout * 2.0 + 16.0), s);
  vae_stem_kernel<<<n_tiles, 256, 0, s>>>(lat, scale, pq_w, pq_b, cin_w, cin_b, out, n_cand, Hlat, tile0, Cout);
  return done();
}

// Encoder.conv_in: one block = 16x16 output pixels of one image; the 18x18xCin fp32 halo and all weights sit in
// shared memory, a thread owns one pixel and walks the output channels (8 at a time -> one 16-byte store).
__global__ void __launch_bounds__(256) vae_enc_stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, __half* __restrict__ out,
                                                           int Cin, int Cout) {
  extern __shared__ float sm[];
  float* halo = sm;                        // [Cin][18][18]
  float* wsm = sm + Cin * 324;             // [Cout][Cin*9]
  float* bsm = wsm + Cout * Cin * 9;       // [Cout]
  const int img = blockIdx.z, by = blockIdx.y * 16, bx = blockIdx.x * 16;
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < Cin * 324; i += blockDim.x) {
    const int c = i / 324, r = (i % 324) / 18, q = i % 18;
    const int yy = by + r - 1, xx = bx + q - 1;
    halo[i] = (yy >= 0 && yy < 128 && xx >= 0 && xx < 128) ? x[(((long long)img * Cin + c) * 128 + yy) * 128 + xx] : 0.f;
  }
  for (int i = threadIdx.x; i < Cout * K; i += blockDim.x) wsm[i] = w[i];  // torch layout [Cout][Cin][3][3]
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bsm[i] = b[i];
  __syncthreads();
  const int r = threadIdx.x >> 4, q = threadIdx.x & 15;
  float in[36];
  for (int c = 0; c < Cin; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) in[c * 9 + t] = halo[c * 324 + (r + t / 3) * 18 + q + t % 3];
  __half* orow = out + (((long long)img * 128 + by + r) * 128 + bx + q) * Cout;
  for (int co = 0; co < Cout; co += 8) {
    uint4 pk;
    __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      float a0 = bsm[co + u], a1 = bsm[co + u + 1];
      for (int k = 0; k < K; ++k) {
        a0 = fmaf(wsm[(co + u) * K + k], in[k], a0);
        a1 = fmaf(wsm[(co + u + 1) * K + k], in[k], a1);
      }
      h2[u >> 1] = __floats2half2_rn(a0, a1);
    }
    *reinterpret_cast<uint4*>(orow + co) = pk;
  }
}

cudaError_t launch_vae_enc_stem(const float* x, const float* w, const float* b, __half* out, int n, int Cin, int Cout,
                                cudaStream_t s) {
  if (Cin < 1 || Cin > 4 || Cout % 8 != 0 || n <= 0) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(Cin * 324 + Cout * Cin * 9 + Cout) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  ProfScope prof("vae_enc_stem", 0, 0, (double)n * 16384.0 * (Cout * 2.0 + Cin * 4.0), s);
  vae_enc_stem_kernel<<<dim3(8, 8, n), 256, smem, s>>>(x, w, b, out, Cin, Cout);
  return done();
}

// quant_conv: thread per (image, pixel); C <= 8 channels in, C out, fp32; output NCHW
__global__ void __launch_bounds__(256) vae_quant_kernel(const float* __restrict__ h, int ld, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ moments,
                                                        long long total, int HW, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long img = i / HW;
  const int pix = (int)(i - img * HW);
  float v[8];
  for (int c = 0; c < C; ++c) v[c] = h[i * ld + c];
  for (int o = 0; o < C; ++o) {
    float a = b[o];
    for (int c = 0; c < C; ++c) a = fmaf(w[o * C + c], v[c], a);
    moments[(img * C + o) * HW + pix] = a;
  }
}

cudaError_t launch_vae_quant(const float* h, int ld, const float* w, const float* b, float* moments, int n, int HW,
                             int C, cudaStream_t s) {
  if (C < 1 || C > 8 || ld < C) return cudaErrorInvalidValue;
  const long long total = (long long)n * HW;
  ProfScope prof("vae_quant", 0, 0, (double)total * (ld + C) * 4.0, s);
  vae_quant_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h, ld, w, b, moments, total, HW, C);
  return done();
}

// GroupNorm statistics straight from the tensor: one block per (image, 8-channel slab); used where no conv epilogue
// produced partial sums (the stem output).  Accumulates in fp32 per thread, double across the block.
__global__ void __launch_bounds__(256) gn_stats_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float2* __restrict__ ab, int HW,
                                                       int C, float eps) {
  const int img = blockIdx.y, grp = blockIdx.x;
  const int cpg = C / 32;  // channels per group: 4, 8 or 16
  const __half* base = x + (long long)img * HW * C + grp * cpg;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < HW * cpg; i += blockDim.x) {
    const int p = i / cpg, c = i - p * cpg;
    const float v = __half2float(base[(long long)p * C + c]);
    s += v;
    q += v * v;
  }
  __shared__ double ss[256], qq[256];
  ss[threadIdx.x] = s;
  qq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      qq[threadIdx.x] += qq[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < cpg) {
    const double n = (double)HW * cpg;
    const double mean = ss[0] / n;
    const double var = qq[0] / n - mean * mean;
    const float rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps));
    const int c = grp * cpg + threadIdx.x;
    const float a = rstd * gamma[c];
    ab[(long long)img * C + c] = make_float2(a, beta[c] - (float)mean * a);
  }
}

cudaError_t launch_gn_stats(const __half* x, const float* gamma, const float* beta, float2* ab, int n, int HW, int C,
                            float eps, cudaStream_t s) {
  if (C % 32 != 0) return cudaErrorInvalidValue;
  ProfScope prof("gn_stats", 0, 0, (double)n * HW * C * 2.0, s);
  gn_stats_kernel<<<dim3(32, n), 256, 0, s>>>(x, gamma, beta, ab, HW, C, eps);
  return done();
}

// Fold the conv epilogue's per-(128 rows x 4 channels) partial sums into per-(image, group) statistics in a fixed
// order (deterministic), in double, and emit the per-channel affine.  One block per image: thread t owns quad
// t % nq (consecutive threads read consecutive float2 -> coalesced) and every (256 / nq)-th slot.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ part,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float2* __restrict__ ab,
                                                          int slots_per_img, int n_par, long long par_stride, int C,
                                                          int HW_out, float eps) {
  __shared__ double ss[256], qq[256];
  __shared__ double gsum[32], gsq[32];
  const int img = blockIdx.x;
  const int nq = C / 4;          // 32, 64 or 128 quads
  const int lanes = 256 / nq;    // slot lanes
  const int qi = threadIdx.x % nq, sl0 = threadIdx.x / nq;
  const float2* p2 = reinterpret_cast<const float2*>(part);
  double s = 0.0, q = 0.0;
  for (int par = 0; par < n_par; ++par)
    for (int sl = sl0; sl < slots_per_img; sl += lanes) {
      const float2 v = p2[(par * par_stride + (long long)img * slots_per_img + sl) * nq + qi];
      s += v.x;
      q += v.y;
    }
  ss[threadIdx.x] = s;
  qq[threadIdx.x] = q;
  __syncthreads();
  const int qpg = C / 128;  // quads per group (1, 2 or 4)
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l)
      for (int k = 0; k < qpg; ++k) {
        a += ss[l * nq + threadIdx.x * qpg + k];
        b += qq[l * nq + threadIdx.x * qpg + k];
      }
    gsum[threadIdx.x] = a;
    gsq[threadIdx.x] = b;
  }
  __syncthreads();
  const int cpg = C / 32;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int grp = c / cpg;
    const double n = (double)HW_out * cpg;
    const double mean = gsum[grp] / n;
    const double var = gsq[grp] / n - mean * mean;
    const float rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps));
    const float a = rstd * gamma[c];
    ab[(long long)img * C + c] = make_float2(a, beta[c] - (float)mean * a);
  }
}

cudaError_t launch_gn_finalize(const float* part, const float* gamma, const float* beta, float2* ab, int n,
                               int slots_per_img, int n_par, long long par_stride, int C, int HW_out, float eps,
                               cudaStream_t s) {
  if (C % 128 != 0 || C > 1024) return cudaErrorInvalidValue;
  ProfScope prof("gn_finalize", 0, 0, (double)n * n_par * slots_per_img * (C / 4) * 8.0, s);
  gn_finalize_kernel<<<n, 256, 0, s>>>(part, gamma, beta, ab, slots_per_img, n_par, par_stride, C, HW_out, eps);
  return done();
}

// grid = (pixel chunks, images).  A thread owns ONE 8-channel vector position (its GroupNorm coefficients stay in
// registers) and walks the image's pixels with a stride, three 16-byte loads in flight; a warp covers 512 contiguous
// bytes per load.  One fp16 read and one fp16 write per element.
// PIPE = 1 keeps the next batch's loads in flight while this batch is normalised and stored (64 registers, 4 blocks
// per SM); PIPE = 0 is the 40-register form that fits beside a resident GEMM CTA.  Measured: 4.9 vs 4.7 TB/s, and the
// two-lane decode is faster with PIPE = 1 as well.
template <int SWISH, int PIPE>
__global__ void __launch_bounds__(256, PIPE ? 4 : 6) gn_apply_kernel(const __half* __restrict__ x,
                                                                     const float2* __restrict__ ab,
                                                                     __half* __restrict__ y, int HW, int C) {
  const int cv = C >> 3;                       // 8-channel vectors per pixel (16, 32 or 64)
  const int c8 = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;           // pixel lane within the block
  const int ppb = 256 / cv;                    // pixels per block iteration
  const int img = blockIdx.y;
  float a[8], b[8];
  {
    const float4* abp = reinterpret_cast<const float4*>(ab + (long long)img * C + c8 * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = __ldg(abp + j);  // (a0, b0, a1, b1)
      a[2 * j] = t.x;
      b[2 * j] = t.y;
      a[2 * j + 1] = t.z;
      b[2 * j + 1] = t.w;
    }
  }
  const uint4* xin = reinterpret_cast<const uint4*>(x) + (long long)img * HW * cv + c8;
  uint4* yout = reinterpret_cast<uint4*>(y) + (long long)img * HW * cv + c8;
  const int step = gridDim.x * ppb;
  auto apply = [&](const uint4& u) -> uint4 {
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      float v0 = fmaf(a[2 * j], f.x, b[2 * j]), v1 = fmaf(a[2 * j + 1], f.y, b[2 * j + 1]);
      if (SWISH) {
        v0 = swish_fast(v0);
        v1 = swish_fast(v1);
      }
      o2[j] = __floats2half2_rn(v0, v1);
    }
    return o;
  };
  // software pipeline: the loads of the next batch of three pixels are in flight while this batch is normalised and
  // stored, so the thread always has 3-6 x 16 B outstanding (without it the loads were in flight about half the time
  // and the kernel sat at 4.1 TB/s)
  int p = blockIdx.x * ppb + prow;
  if constexpr (!PIPE) {
    for (; p + 2 * step < HW; p += 3 * step) {
      uint4 u[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) u[i] = __ldcs(xin + (long long)(p + i * step) * cv);
#pragma unroll
      for (int i = 0; i < 3; ++i) yout[(long long)(p + i * step) * cv] = apply(u[i]);
    }
    for (; p < HW; p += step) yout[(long long)p * cv] = apply(__ldcs(xin + (long long)p * cv));
    return;
  }
  uint4 u[3];
  bool have = p + 2 * step < HW;
  if (have) {
#pragma unroll
    for (int i = 0; i < 3; ++i) u[i] = __ldcs(xin + (long long)(p + i * step) * cv);
  }
  while (have) {
    const int pn = p + 3 * step;
    const bool more = pn + 2 * step < HW;
    uint4 nx[3];
    if (more) {
#pragma unroll
      for (int i = 0; i < 3; ++i) nx[i] = __ldcs(xin + (long long)(pn + i * step) * cv);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) yout[(long long)(p + i * step) * cv] = apply(u[i]);
    if (more) {
#pragma unroll
      for (int i = 0; i < 3; ++i) u[i] = nx[i];
    }
    p = pn;
    have = more;
  }
  for (; p < HW; p += step) yout[(long long)p * cv] = apply(__ldcs(xin + (long long)p * cv));
}

cudaError_t launch_gn_apply(const __half* x, const float2* ab, __half* y, int n, int HW, int C, int swish,
                            cudaStream_t s) {
  if (C % 8 != 0 || 256 % (C / 8) != 0) return cudaErrorInvalidValue;
  const long long total_vec = (long long)n * HW * (C / 8);
  char pname[48];
  snprintf(pname, sizeof pname, "gn_apply HW%d C%d", HW, C);
  ProfScope prof(pname, 0, 0, (double)total_vec * 32.0, s);
  const int ppb = 256 / (C / 8);
  // enough blocks to fill the machine (~16 per SM) without making the per-thread loops shorter than one batch of 4
  int chunks = (HW + ppb * 4 - 1) / (ppb * 4);
  const int want = (148 * 16 + n - 1) / n;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  static const int pipe = [] {
    const char* e = getenv("RGM_GN_PIPE");  // 1 (default): software-pipelined loads, 64 registers, 4 blocks per SM;
    return e ? atoi(e) : 1;                 // 0: 40 registers, 6 blocks per SM (4.7 vs 4.9 TB/s at 128x128x128)
  }();
  if (swish && pipe) gn_apply_kernel<1, 1><<<dim3(chunks, n), 256, 0, s>>>(x, ab, y, HW, C);
  else if (swish) gn_apply_kernel<1, 0><<<dim3(chunks, n), 256, 0, s>>>(x, ab, y, HW, C);
  else gn_apply_kernel<0, 0><<<dim3(chunks, n), 256, 0, s>>>(x, ab, y, HW, C);
  return done();
}

// norm_out (GroupNorm apply + swish, model.py:533-535) + conv_out 3x3 C -> out_ch <= 8 (model.py:536) + assembly of
// the piano roll (gaussian_diffusion.py:1355) in one kernel.  With 3 output channels a tcgen05 tile would be > 90 %
// padding and the 9 shifted re-reads of the 128-channel input would dominate, so: one block = 16x16 output pixels; the
// 18x18 halo of the RAW input is loaded once, normalised + activated on the way into shared memory (zero for
// out-of-image pixels: the conv pads the ACTIVATED tensor); the 9-tap contraction then runs on warp-level
// mma.sync m16n8k16 (16 pixels x 8 padded output channels x 16 input channels, fp32 accumulate) with ldmatrix
// operand loads -- a CUDA-core version of the same loop measured instruction-issue bound at 80 K warp-instructions per
// block.  Pixel stride in smem is padded by 16 B so ldmatrix rows (consecutive pixels) hit distinct banks.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// conv_out weights as B fragments of mma.m16n8k16 (col-major k16 x n8) for the "one A fragment serves the three dx taps"
// contraction of vae_out_kernel: for a kernel ROW dy the B matrix is [K = input channels][N = (dx, out channel)], so
// Y_dy[q][(dx, co)] = sum_k act[q][k] * W[co][k][dy][dx] for every halo pixel q, and
// out[r][c][co] = sum_dy sum_dx Y_dy[(r + dy) * 18 + c + dx][(dx, co)].
// Two packings, built ONCE when the weights are loaded: NF = 1 -- columns n = dx (output channel 0 only: what the SCG
// rules read) -- and NF = 2 -- columns n = dx * out_ch + co (all channels, <= 16 columns).  Entry
// [dy][kc][nf][lane] = (W[k0][n], W[k0+1][n]) and (W[k0+8][n], W[k0+9][n]) as two half2, k0 = kc*16 + 2*(lane%4),
// n = nf*8 + lane/4; columns past the last real one are zero.
__global__ void vae_out_pack_kernel(const float* __restrict__ w, uint2* __restrict__ bfrag, int C, int out_ch, int NF) {
  const int kchunks = C >> 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * kchunks * NF * 32) return;
  const int l = i & 31, nf = (i >> 5) % NF, kc = (i / (32 * NF)) % kchunks, dy = i / (32 * NF * kchunks);
  const int n = nf * 8 + (l >> 2), k0 = kc * 16 + 2 * (l & 3);
  const int cols_per_dx = NF == 1 ? 1 : out_ch;
  const int dx = n / cols_per_dx, co = n - dx * cols_per_dx;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (dx < 3 && co < out_ch) {
    const float* wp = w + (long long)co * C * 9 + dy * 3 + dx;  // torch layout [out_ch][C][3][3]
    v[0] = wp[(k0) * 9];
    v[1] = wp[(k0 + 1) * 9];
    v[2] = wp[(k0 + 8) * 9];
    v[3] = wp[(k0 + 9) * 9];
  }
  __half2 lo = __floats2half2_rn(v[0], v[1]), hi = __floats2half2_rn(v[2], v[3]);
  bfrag[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// bfrag: room for both packings back to back: 3 * C/16 * 32 uint2 (NF = 1) then 3 * C/16 * 2 * 32 uint2 (NF = 2)
cudaError_t launch_vae_out_pack(const float* w, void* bfrag, int C, int out_ch, cudaStream_t s) {
  if (C % 16 != 0 || out_ch < 1 || 3 * out_ch > 16) return cudaErrorInvalidValue;
  const int n1 = 3 * (C / 16) * 32;
  vae_out_pack_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(w, static_cast<uint2*>(bfrag), C, out_ch, 1);
  vae_out_pack_kernel<<<(2 * n1 + 255) / 256, 256, 0, s>>>(w, static_cast<uint2*>(bfrag) + n1, C, out_ch, 2);
  return done();
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// norm_out (GroupNorm apply + swish, model.py:533-535) + conv_out 3x3 C -> out_ch (model.py:536) + assembly of the piano
// roll (gaussian_diffusion.py:1355) in one kernel.  One block = 16x16 output pixels of one tile.  Per SLICE of 64 input
// channels: the raw 18x18 halo arrives by cp.async (everything in flight at once), is normalised + activated in place
// (out-of-image pixels stay zero: the conv pads the ACTIVATED tensor; gn_apply_kernel's arithmetic and rounding), and is
// contracted on mma.sync m16n8k16 with the halo FLATTENED to 324 (+12 padding) pixels = 21 m16 tiles: one ldmatrix
// fragment of 16 consecutive halo pixels serves all three dx taps of a kernel row (they are columns of the B matrix) and
// all three rows dy (three B fragments) -- 4x fewer shared-memory wavefronts than nine shifted fragment loads per
// output row, which is what bounded the previous form (ncu: 84 M shared wavefronts per launch, 72 % of its cycles).
// The per-(dy, halo pixel, dx) partial products then go through shared memory once and each output pixel sums its nine.
constexpr int VO_SLICE = 64;
constexpr int VO_PSTRIDE = VO_SLICE * 2 + 16;           // bytes per halo pixel: 16 B of padding -> conflict-free ldmatrix
constexpr int VO_HALO_PIX = 18 * 18;                    // 324
constexpr int VO_TILES = (VO_HALO_PIX + 15) / 16;       // 21 m16 tiles over the flattened halo
constexpr int VO_HALO_BYTES = VO_TILES * 16 * VO_PSTRIDE;  // 48 384 (pixels 324..335 are zero padding)
constexpr int VO_KC = VO_SLICE / 16;                    // k-steps per slice
constexpr int VO_WARPS = 16;

template <int NF>
__global__ void __launch_bounds__(32 * VO_WARPS, NF == 1 ? 2 : 1) vae_out_kernel(const __half* __restrict__ x, const float2* __restrict__ ab,
                                                                  const uint2* __restrict__ bfrag_g,
                                                                  const float* __restrict__ bias, float* __restrict__ roll,
                                                                  int C, int tile0, int n_cand, int roll_len,
                                                                  int roll_ch) {
  extern __shared__ __align__(16) uint8_t osm[];
  // two slice buffers: [336][VO_PSTRIDE] halo + [3][VO_KC][NF][32 lanes] weight fragments each (buffer 0 later holds Y)
  constexpr int BF_BYTES = 3 * VO_KC * NF * 32 * 8;
  constexpr int BUF_BYTES = VO_HALO_BYTES + BF_BYTES;
  const int kchunks = C >> 4;
  const int img = blockIdx.y;
  const int by = blockIdx.x >> 3, bx = blockIdx.x & 7;  // 8 x 8 blocks of 16 x 16 pixels per 128 x 128 image
  constexpr int cv = VO_SLICE >> 3;                      // 8 chunks of 8 channels per pixel and slice
  constexpr int NT = 32 * VO_WARPS;
  const __half* xin = x + (long long)img * 128 * 128 * C;
  const uint32_t osm_s = static_cast<uint32_t>(__cvta_generic_to_shared(osm));
  constexpr int items = VO_TILES * 16 * cv;  // padding pixels included (written as zeros)
  const int c8 = threadIdx.x % cv;           // NT is a multiple of cv: a thread always handles the same 8-channel chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[2][3][NF][4];
#pragma unroll
  for (int i = 0; i < 2 * 3 * NF * 4; ++i) (&acc[0][0][0][0])[i] = 0.f;
  // ldmatrix.x4 address of this lane: matrix m = lane / 8 covers pixels 8*(m & 1) .. +7 and channels 8*(m >> 1) .. +7
  const int lm_pix = (lane & 7) + 8 * ((lane >> 3) & 1);
  const int lm_coff = (lane >> 4) * 16;  // bytes
  constexpr int bf_pieces = BF_BYTES / 16;  // 16-byte pieces of one slice's fragments
  // everything a slice reads from global memory goes in flight at once with cp.async (16 bytes each, no registers held),
  // one commit group per slice; slice s + 1 is fetched into the other buffer while slice s is normalised and contracted
  auto fetch = [&](int c0, int buf) {
    const uint32_t halo_s = osm_s + buf * BUF_BYTES;
    for (int i = threadIdx.x; i < items; i += NT) {
      const int pix = i / cv;
      const int hy = pix / 18, hx = pix - hy * 18;
      const int gy = by * 16 + hy - 1, gx = bx * 16 + hx - 1;
      const uint32_t dst = halo_s + pix * VO_PSTRIDE + c8 * 16;
      if (pix < VO_HALO_PIX && gy >= 0 && gy < 128 && gx >= 0 && gx < 128)
        cp_async_16(dst, xin + ((long long)gy * 128 + gx) * C + c0 + c8 * 8);
      else
        sts_v4(dst, make_uint4(0u, 0u, 0u, 0u));
    }
    for (int i = threadIdx.x; i < bf_pieces; i += NT) {
      constexpr int per_dy = VO_KC * NF * 16;  // pieces per kernel row of the slice
      const int dy = i / per_dy, r = i - dy * per_dy;
      cp_async_16(halo_s + VO_HALO_BYTES + i * 16,
                  reinterpret_cast<const uint4*>(bfrag_g) + ((long long)dy * kchunks + (c0 >> 4)) * NF * 16 + r);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(0, 0);
  int sl = 0;
  for (int c0 = 0; c0 < C; c0 += VO_SLICE, ++sl) {
    const int buf = sl & 1;
    const uint32_t halo_s = osm_s + buf * BUF_BYTES;
    const uint2* bfrag = reinterpret_cast<const uint2*>(osm + buf * BUF_BYTES + VO_HALO_BYTES);
    const bool more = c0 + VO_SLICE < C;
    if (more) {
      if (sl >= 1) __syncthreads();  // the contraction of slice s - 1 has read the buffer slice s + 1 goes into
      fetch(c0 + VO_SLICE, buf ^ 1);
    }
    float ga[8], gb[8];
    {
      const float4* abp = reinterpret_cast<const float4*>(ab + (long long)img * C + c0 + c8 * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(abp + j);  // (a0, b0, a1, b1)
        ga[2 * j] = t.x;
        gb[2 * j] = t.y;
        ga[2 * j + 1] = t.z;
        gb[2 * j + 1] = t.w;
      }
    }
    // this thread's copies of slice s have landed (it transforms exactly the chunks it fetched)
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int i = threadIdx.x; i < VO_HALO_PIX * cv; i += NT) {
      const int pix = i / cv;
      const int hy = pix / 18, hx = pix - hy * 18;
      const int gy = by * 16 + hy - 1, gx = bx * 16 + hx - 1;
      if (gy >= 0 && gy < 128 && gx >= 0 && gx < 128) {
        const uint32_t addr = halo_s + pix * VO_PSTRIDE + c8 * 16;
        uint4 u = lds_v4(addr);
        __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          const float v0 = swish_fast(fmaf(ga[2 * j], f.x, gb[2 * j]));
          const float v1 = swish_fast(fmaf(ga[2 * j + 1], f.y, gb[2 * j + 1]));
          h2[j] = __floats2half2_rn(v0, v1);
        }
        sts_v4(addr, u);
      }
    }
    __syncthreads();
    // contraction: warp w owns halo-pixel tiles w and w + 16 (21 tiles)
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) {
      const int tile = warp + ti * VO_WARPS;
      if (tile < VO_TILES) {
        const uint32_t a_row0 = halo_s + (tile * 16 + lm_pix) * VO_PSTRIDE + lm_coff;
#pragma unroll
        for (int kc = 0; kc < VO_KC; ++kc) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(a_row0 + kc * 32, a0, a1, a2, a3);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
              const uint2 b = bfrag[((dy * VO_KC + kc) * NF + nf) * 32 + lane];
              mma_16816(acc[ti][dy][nf], a0, a1, a2, a3, b.x, b.y);
            }
        }
      }
    }
  }
  __syncthreads();  // every warp is done with the halo: its memory now holds Y[dy][halo pixel][8 * NF columns] fp32
  float* Y = reinterpret_cast<float*>(osm);  // 3 * 336 * 8 * NF floats <= two slice buffers
  constexpr int YC = 8 * NF;
  // accumulator layout: c0,c1 = (pixel lane/4, columns 2*(lane%4) + {0,1}); c2,c3 = (pixel lane/4 + 8, same columns)
#pragma unroll
  for (int ti = 0; ti < 2; ++ti) {
    const int tile = warp + ti * VO_WARPS;
    if (tile < VO_TILES) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          float* yp = Y + ((dy * (VO_TILES * 16) + tile * 16 + (lane >> 2)) * YC) + nf * 8 + 2 * (lane & 3);
          *reinterpret_cast<float2*>(yp) = make_float2(acc[ti][dy][nf][0], acc[ti][dy][nf][1]);
          *reinterpret_cast<float2*>(yp + 8 * YC) = make_float2(acc[ti][dy][nf][2], acc[ti][dy][nf][3]);
        }
    }
  }
  __syncthreads();
  // each output pixel sums its nine partial products; consecutive threads = consecutive time columns of a roll row
  const int g = tile0 + img;  // global tile index, tile-major: g = k * n_cand + cand
  const int kt = g / n_cand, cand = g - kt * n_cand;
  const int cols_per_dx = NF == 1 ? 1 : 3;
  for (int o = threadIdx.x; o < 256 * roll_ch; o += NT) {
    const int ch = o >> 8, r = (o >> 4) & 15, c = o & 15;
    float v = bias[ch];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        v += Y[(dy * (VO_TILES * 16) + (r + dy) * 18 + c + dx) * YC + dx * cols_per_dx + ch];
    roll[(((long long)cand * roll_ch + ch) * 128 + by * 16 + r) * roll_len + kt * 128 + bx * 16 + c] = v;
  }
}

cudaError_t launch_vae_out(const __half* x, const float2* ab, const void* bfrag, const float* bias, float* roll, int n,
                           int C, int out_ch, int tile0, int n_cand, int roll_len, int roll_ch, cudaStream_t s) {
  if (C % VO_SLICE != 0 || out_ch != 3 || roll_ch < 1 || roll_ch > out_ch) return cudaErrorInvalidValue;
  ProfScope prof("vae_out(norm+swish+conv_out+roll)", 2.0 * n * 16384.0 * out_ch * 9.0 * C, 2.0 * n * 16384.0 * 8 * 9.0 * C,
                 (double)n * 16384.0 * (C * 2.0 + roll_ch * 4.0), s);
  const uint2* bf = static_cast<const uint2*>(bfrag);
  const int n1 = 3 * (C / 16) * 32;  // uint2 entries of the one-channel packing (launch_vae_out_pack)
  if (roll_ch == 1) {  // the SCG path: the rules read channel 0 only
    const size_t smem = 2 * (VO_HALO_BYTES + 3 * VO_KC * 1 * 32 * 8);  // two slice buffers; the partial products reuse them
    static SmemAttr attr;
    if (cudaError_t e = attr.ensure(vae_out_kernel<1>, smem); e != cudaSuccess) return e;
    vae_out_kernel<1><<<dim3(64, n), 32 * VO_WARPS, smem, s>>>(x, ab, bf, bias, roll, C, tile0, n_cand, roll_len, roll_ch);
  } else {
    const size_t smem = 2 * (VO_HALO_BYTES + 3 * VO_KC * 2 * 32 * 8);  // 63 KB of partial products fit the two buffers
    static SmemAttr attr;
    if (cudaError_t e = attr.ensure(vae_out_kernel<2>, smem); e != cudaSuccess) return e;
    vae_out_kernel<2><<<dim3(64, n), 32 * VO_WARPS, smem, s>>>(x, ab, bf + n1, bias, roll, C, tile0, n_cand, roll_len,
                                                              roll_ch);
  }
  return done();
}

// warp per row, cols <= 1024 and a multiple of 32
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, __half* __restrict__ y,
                                                           long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int per = cols >> 5;  // <= 32
  for (long long row = warp0; row < rows; row += nwarps) {
    const float* xr = x + row * cols;
    float v[32];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) {
        v[i] = xr[i * 32 + lane];
        m = fmaxf(m, v[i]);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) {
        v[i] = __expf(v[i] - m);
        s += v[i];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) y[row * cols + i * 32 + lane] = __float2half_rn(v[i] * inv);
  }
}

cudaError_t launch_softmax_rows(const float* x, __half* y, long long rows, int cols, cudaStream_t s) {
  if (cols % 32 != 0 || cols > 1024) return cudaErrorInvalidValue;
  ProfScope prof("softmax_rows", 0, 0, (double)rows * cols * 6.0, s);
  softmax_rows_kernel<<<blocks_for(rows, 8, 148 * 8), 256, 0, s>>>(x, y, rows, cols);
  return done();
}

__global__ void transpose_kernel(const __half* __restrict__ x, __half* __restrict__ y, int R, int C) {
  __shared__ __half tile[32][33];
  const long long base = (long long)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    tile[j][threadIdx.x] = x[base + (long long)(r0 + j) * C + c0 + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    y[base + (long long)(c0 + j) * R + r0 + threadIdx.x] = tile[threadIdx.x][j];
}

cudaError_t launch_transpose(const __half* x, __half* y, int n, int R, int C, cudaStream_t s) {
  if (R % 32 != 0 || C % 32 != 0) return cudaErrorInvalidValue;
  ProfScope prof("transpose", 0, 0, (double)n * R * C * 4.0, s);
  transpose_kernel<<<dim3(C / 32, R / 32, n), dim3(32, 8), 0, s>>>(x, y, R, C);
  return done();
}

// ------------------------------------------------------------------------------------------------------------
// sampler elementwise
// ------------------------------------------------------------------------------------------------------------
__global__ void scg_fanout_kernel(const float* __restrict__ mean, const float* __restrict__ g,
                                  const float* __restrict__ noise, float* __restrict__ cand, int N, int B,
                                  long long elems) {
  const long long per_n = (long long)B * elems;
  const long long total = per_n * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = i % per_n;
    const int b = (int)(j / elems);
    cand[i] = mean[j] + g[b] * noise[i];
  }
}

cudaError_t launch_scg_fanout(const float* mean, const float* g, const float* noise, float* cand, int N, int B,
                              long long elems, cudaStream_t s) {
  scg_fanout_kernel<<<blocks_for((long long)N * B * elems, 256), 256, 0, s>>>(mean, g, noise, cand, N, B, elems);
  return done();
}

__global__ void x0_from_eps_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                   const float* __restrict__ a, const float* __restrict__ c, float* __restrict__ x0,
                                   long long total, long long elems, int clamp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / elems);
    // same two roundings as the reference's a*x - c*eps (two products, one subtraction; no fused multiply-add)
    float v = __fsub_rn(__fmul_rn(a[b], x[i]), __fmul_rn(c[b], eps[i]));
    if (clamp) v = fminf(fmaxf(v, -1.f), 1.f);
    x0[i] = v;
  }
}

cudaError_t launch_x0_from_eps(const float* x, const float* eps, const float* a, const float* c, float* x0, int B,
                               long long elems, int clamp, cudaStream_t s) {
  const long long total = (long long)B * elems;
  x0_from_eps_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, eps, a, c, x0, total, elems, clamp);
  return done();
}

}  // namespace rgm
