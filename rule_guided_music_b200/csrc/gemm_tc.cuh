// tcgen05 implicit-GEMM kernel shared by every dense contraction on the path:
//   * DiT linears (qkv / proj / fc1 / fc2 / adaLN / embedders / final)        -- reference guided_diffusion/dit.py:256-286,324-336
//   * VAE decoder convolutions 3x3 / 1x1 / nearest-2x-upsample+3x3 (as 4 parity sub-convs) and the mid attention
//     matmuls                                                               -- reference taming/modules/diffusionmodules/model.py:38-53,78-192
//
// D[M, N] = sum_taps A_shift(tap)[M, C] * B[N, tap*C : (tap+1)*C]^T, fp16 operands, fp32 accumulation in TMEM.
// A is an NHWC fp16 tensor [n_img, H, W, C]; an M tile is a (bh x bw) pixel box of one image, loaded by TMA
// with the tap's (dy, dx) shift -- out-of-image pixels are zero-filled by TMA, which IS the conv's zero padding.
// A plain linear layer is the degenerate case n_img = 1, H = 1, W = M, one tap.
#pragma once
#include "ptx.cuh"

namespace rgm {

enum EpiKind : int {
  EPI_F16 = 0,         // out16[orow, n] = act(alpha*acc + bias[n] + addtab[addidx[row], n]) + resid16[orow, n]
  EPI_F32 = 1,         // out32[orow, n] = act(alpha*acc + bias[n] + addtab[addidx[row], n])
  EPI_GATE_RESID = 2,  // x32[row, n] += gate[row / rows_per_sample, n] * (acc + bias[n])
  EPI_QKV_ROPE = 3,    // + bias, rotary on q,k, scatter q,k to [B, heads, T, dh_pad] and v to [B, heads, dh, T] fp16
  EPI_UNPATCH = 4,     // final layer: + bias, scatter tokens back to NCHW fp32 latent
  EPI_ROLL = 5,        // conv_out: + bias, scatter VAE tiles into the piano roll [cand, ch, 128, L] fp32
};

enum ActKind : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU_TANH = 2 };

struct EpiParams {
  void* out;      // fp16 or fp32 depending on kind
  int ldo;        // row stride of out, in elements
  const float* bias;
  float alpha;
  int act;
  // EPI_F16 residual (same row mapping as out)
  const __half* resid;
  int ldr;
  // optional gathered row-vector add (label embedding): addtab[addidx[row] * N + n]
  const float* addtab;
  const long long* addidx;
  // EPI_GATE_RESID
  const float* gate;
  int gate_ld;
  int rows_per_sample;
  // EPI_QKV_ROPE
  __half* q;
  __half* k;
  __half* v;
  const float* rope_cos;  // [T, rot_dim/2]
  const float* rope_sin;
  int T, heads, dh, dh_pad, rot_dim;
  // nearest-2x upsample output mapping (EPI_F16 with up2 = 1): low-res image dims
  int up2, upH, upW;
  // EPI_UNPATCH: tokens per time step (W / patch), out channels, latent H, W
  int tpt, c_out, latH, latW, n_valid;
  // EPI_ROLL
  int tile0, n_cand, roll_len, roll_ch;
  // optional GroupNorm partial statistics of the fp16 values just stored (EPI_F16):
  // gn_part[par * num_m_tiles * 4 + row / 32][N / 4] = (sum, sumsq) over 32 rows x 4 channels
  float* gn_part;
};

struct GemmParams {
  int M, N;
  int num_m_tiles, num_n_tiles, num_par;
  int num_taps, kb_per_tap;  // kb_per_tap = C / 64
  int tiles_per_img, tiles_per_row, bh, bw;
  int b_batched;              // B's third coordinate follows the image index (batched GEMM)
  signed char tap_dy[4][9];
  signed char tap_dx[4][9];
  EpiParams epi;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr uint32_t A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr uint32_t TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
// epilogue: one thread owns one output row and 32 consecutive columns of it
// ------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int n0, int par, const uint32_t (&r)[32]) {
  const EpiParams& e = p.epi;
  const bool row_ok = row < p.M;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

  if constexpr (EPI == EPI_F16 || EPI == EPI_F32) {
    const float* addrow = nullptr;
    if (e.addtab != nullptr && row_ok) addrow = e.addtab + e.addidx[row] * (long long)p.N;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = v[i] * e.alpha;
      if (e.bias) x += __ldg(e.bias + n0 + i);
      if (addrow) x += __ldg(addrow + n0 + i);
      if (e.act == ACT_SILU) x = silu_f(x);
      else if (e.act == ACT_GELU_TANH) x = gelu_tanh_f(x);
      v[i] = x;
    }
    long long orow = row;
    if (e.up2) {
      const int hw = e.upH * e.upW;
      const int img = row / hw, rem = row - img * hw;
      const int h = rem / e.upW, w = rem - h * e.upW;
      orow = ((long long)img * (2 * e.upH) + (2 * h + (par >> 1))) * (2 * e.upW) + (2 * w + (par & 1));
    }
    if constexpr (EPI == EPI_F32) {
      if (row_ok && e.act != 99) {  // act 99: development knob, skip the store
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(e.out) + orow * e.ldo + n0);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    } else {
      if (e.resid != nullptr && row_ok) {
        const uint4* rs = reinterpret_cast<const uint4*>(e.resid + orow * e.ldr + n0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 u = __ldg(rs + i);
          const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            v[8 * i + 2 * j] += f.x;
            v[8 * i + 2 * j + 1] += f.y;
          }
        }
      }
      uint4 packed[4];
      __half2* h2 = reinterpret_cast<__half2*>(packed);
#pragma unroll
      for (int i = 0; i < 16; ++i) h2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(static_cast<__half*>(e.out) + orow * e.ldo + n0);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = packed[i];
      }
      if (e.gn_part != nullptr) {
        // GroupNorm partial statistics of exactly the values the next layer reads (fp16-rounded), without
        // atomics so the result is run-to-run deterministic: this warp's 32 rows x 32 columns become 8
        // (sum, sumsq) pairs, one per 4-channel quad; gn_finalize_kernel folds quads into groups and sums the
        // per-warp slots of an image in a fixed order.
        float s4[8], q4[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float2 f = __half22float2(h2[2 * g + i]);
            s += f.x + f.y;
            q += f.x * f.x + f.y * f.y;
          }
          s4[g] = row_ok ? s : 0.f;
          q4[g] = row_ok ? q : 0.f;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s4[g] += __shfl_xor_sync(0xffffffffu, s4[g], o);
            q4[g] += __shfl_xor_sync(0xffffffffu, q4[g], o);
          }
        }
        if ((threadIdx.x & 31) == 0) {
          const long long slot = (long long)par * (p.num_m_tiles * 4) + (row >> 5);
          float2* dst = reinterpret_cast<float2*>(e.gn_part) + slot * (p.N >> 2) + (n0 >> 2);
#pragma unroll
          for (int g = 0; g < 8; ++g) dst[g] = make_float2(s4[g], q4[g]);
        }
      }
    }
  } else if constexpr (EPI == EPI_GATE_RESID) {
    if (row_ok) {
      const float* g = e.gate + (long long)(row / e.rows_per_sample) * e.gate_ld + n0;
      float4* x = reinterpret_cast<float4*>(static_cast<float*>(e.out) + (long long)row * e.ldo + n0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 xv = x[i];
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
        const float4 bv = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + i);
        xv.x += gv.x * (v[4 * i] + bv.x);
        xv.y += gv.y * (v[4 * i + 1] + bv.y);
        xv.z += gv.z * (v[4 * i + 2] + bv.z);
        xv.w += gv.w * (v[4 * i + 3] + bv.w);
        x[i] = xv;
      }
    }
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    if (row_ok) {
      const int D = e.heads * e.dh;
      const int b = row / e.T, tok = row - b * e.T;
      const int half_rot = e.rot_dim >> 1;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int n = n0 + i;
        const int which = n / D;
        const int rem = n - which * D;
        const int head = rem / e.dh;
        const int d = rem - head * e.dh;  // even
        float x0 = v[i] + __ldg(e.bias + n);
        float x1 = v[i + 1] + __ldg(e.bias + n + 1);
        if (which < 2 && d < e.rot_dim) {
          const float c = __ldg(e.rope_cos + tok * half_rot + (d >> 1));
          const float s = __ldg(e.rope_sin + tok * half_rot + (d >> 1));
          const float y0 = x0 * c - x1 * s;
          const float y1 = x1 * c + x0 * s;
          x0 = y0;
          x1 = y1;
        }
        if (which == 2) {
          // V is stored transposed, [B, heads, dh, T]: it is the K-major B operand of the P.V MMA (attention.cu).
          // Consecutive lanes hold consecutive tokens, so each 2-byte store instruction covers 64 contiguous bytes.
          __half* vt = e.v + (((long long)b * e.heads + head) * e.dh + d) * e.T + tok;
          vt[0] = __float2half_rn(x0);
          vt[e.T] = __float2half_rn(x1);
        } else {
          __half* base = which == 0 ? e.q : e.k;
          __half2* dst =
              reinterpret_cast<__half2*>(base + (((long long)b * e.heads + head) * e.T + tok) * e.dh_pad + d);
          *dst = __floats2half2_rn(x0, x1);
        }
      }
    }
  } else if constexpr (EPI == EPI_UNPATCH) {
    if (row_ok) {
      const int tokens = e.latH * e.tpt;
      const int b = row / tokens, j = row - b * tokens;
      const int time = j / e.tpt, part = j - time * e.tpt;
      const int patch = e.latW / e.tpt;
      float* out = static_cast<float*>(e.out);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int f = n0 + i;
        if (f < e.n_valid) {
          const int pl = f / e.c_out, ch = f - pl * e.c_out;
          const int pitch = part * patch + pl;
          out[(((long long)b * e.c_out + ch) * e.latH + time) * e.latW + pitch] = v[i] + __ldg(e.bias + f);
        }
      }
    }
  } else if constexpr (EPI == EPI_ROLL) {
    if (row_ok) {
      const int img = row >> 14, pix = row & 16383;  // 128 x 128 output pixels per VAE tile
      const int h = pix >> 7, w = pix & 127;         // h = pitch, w = time within the tile
      const int g = e.tile0 + img;                   // global tile index, tile-major: g = k * n_cand + cand
      const int kt = g / e.n_cand, cand = g - kt * e.n_cand;
      float* out = static_cast<float*>(e.out);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int ch = n0 + i;
        if (ch < e.roll_ch)
          out[(((long long)cand * e.roll_ch + ch) * 128 + h) * e.roll_len + kt * 128 + w] = v[i] + __ldg(e.bias + ch);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..5 = epilogue
// persistent over tiles, double-buffered TMEM accumulator so the epilogue of tile i overlaps the
// mainloop of tile i+1.
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int total_tiles = tiles_mn * p.num_par;
  const int num_kb = p.num_taps * p.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int par = t / tiles_mn;
        const int tt = t - par * tiles_mn;
        const int m_tile = tt / p.num_n_tiles;
        const int n_tile = tt - m_tile * p.num_n_tiles;
        const int img = m_tile / p.tiles_per_img;
        const int rr = m_tile - img * p.tiles_per_img;
        const int h0 = (rr / p.tiles_per_row) * p.bh;
        const int w0 = (rr % p.tiles_per_row) * p.bw;
        const int brow = par * p.N + n_tile * BLOCK_N;
        const int bz = p.b_batched ? img : 0;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int dy = p.tap_dy[par][tap], dx = p.tap_dx[par][tap];
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            tma_load_4d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kc * GEMM_BLOCK_K, w0 + dx,
                        h0 + dy, img);
            tma_load_3d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage],
                        (tap * p.kb_per_tap + kc) * GEMM_BLOCK_K, brow, bz);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_a + stage * Cfg::A_BYTES);
          const uint64_t bdesc = umma_desc_sw128(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete → epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int par = t / tiles_mn;
      const int tt = t - par * tiles_mn;
      const int m_tile = tt / p.num_n_tiles;
      const int n_tile = tt - m_tile * p.num_n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_tile * GEMM_BLOCK_M + quad * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, row, n_tile * BLOCK_N + c * 32, par, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace rgm
