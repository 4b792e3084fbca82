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
#include <type_traits>

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
  const float2* rope_cs;  // [T, rot_dim/2] (cos, sin) of position * frequency
  int T, heads, dh, dh_pad, rot_dim;
  // nearest-2x upsample output mapping (EPI_F16 with up2 = 1): low-res image dims
  int up2, upH, upW;
  // EPI_UNPATCH: tokens per time step (W / patch), out channels, latent H, W
  int tpt, c_out, latH, latW, n_valid;
  // EPI_ROLL
  int tile0, n_cand, roll_len, roll_ch;
  // optional GroupNorm partial statistics of the fp16 values just stored (EPI_F16, feature-major kernel):
  // gn_part[par * ceil(M/128) + row / 128][N / 4] = (sum, sumsq) over 128 rows x 4 channels
  float* gn_part;
  // GroupNorm (+ swish) of this convolution's OWN output applied inside the epilogue (feature-major kernels, EPI_F16,
  // gn_epilogue_loop): when gn_sums != nullptr `out` receives swish(GroupNorm(conv(x))) and the raw tensor is never
  // written.  gn_sums: per-(image, group) accumulators [image][32][2] -- fixed-point sum / sum of squares in bits 63..8,
  // arrivals in bits 7..0 -- zeroed before the launch; gn_inv_count = 1 / (pixels x channels per group); gn_gamma /
  // gn_beta: the norm's affine [N]; gn_err: set to 1 if a wait gives up, 2 if a partial sum leaves the fixed-point range
  // (never in a healthy run)
  unsigned long long* gn_sums;
  float gn_inv_count;
  // "dual" form: `out` receives the RAW tensor as usual (residual allowed) and gn_out2 (same layout) the normalised copy
  __half* gn_out2;
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int gn_swish;
  int* gn_err;
};

struct GemmParams {
  int M, N;
  int num_m_tiles, num_n_tiles, num_par;
  int num_taps, kb_per_tap;  // kb_per_tap = C / 64
  int tiles_per_img, tiles_per_row, bh, bw;
  int b_batched;              // B's third coordinate follows the image index (batched GEMM)
  int slots_per_par;          // GroupNorm partial slots (128-row groups) per output parity = ceil(M / 128)
  int debug;                  // development aid (pair kernel): bit 0 = skip the TMA loads, bit 1 = skip the epilogue,
                              // bit 2 = epilogue reads TMEM but computes / stores nothing, bit 3 = EPI_F16 computes but does not store
  int band_n;                 // pair kernel rasterisation: feature-tile pairs per band (0 = all: feature pairs fastest)
  int par_fast;               // output parities of an upsample conv vary fastest in the tile order (else slowest)
  int in_stride;              // 1, or 2 for the stride-2 Downsample conv: input pixel = in_stride * output pixel + tap
  signed char tap_dy[4][9];
  signed char tap_dx[4][9];
  // development aid: when non-null, CTA 0 writes clock64() at pipeline events of each of its tiles, 8 slots per tile:
  // 0 producer tile start, 1 MMA accumulator free, 2 MMA first operands landed, 3 MMA last issue,
  // 4 epilogue accumulator ready, 5 epilogue done
  unsigned long long* trace;
  EpiParams epi;
};

// Tile index -> (output parity, tile within the parity).  The parity varies FASTEST: the four parity sub-convolutions of
// an upsample conv read the same input tile, so running them on neighbouring CTAs at the same time turns three of the
// four DRAM reads of the input into L2 hits (round 1 measured 1078 MB read for a 268 MB input with parity-major order).
__device__ __forceinline__ void split_parity(int t, int num_par, int tiles_mn, bool par_fast, int& par, int& tt) {
  if (num_par == 1) {
    par = 0;
    tt = t;
  } else if (par_fast) {
    par = t & (num_par - 1);  // num_par is 1 or 4
    tt = t / num_par;
  } else {
    par = t / tiles_mn;
    tt = t - par * tiles_mn;
  }
}

constexpr int GEMM_BLOCK_M = 128;   // rows of one UMMA (and of one TMA box of A)
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

// BLOCK_N: output columns per tile; MT: 128-row sub-tiles per tile (the B stage is shared by the MT UMMAs, which
// halves L2->smem operand traffic per flop for narrow outputs such as the 128-channel convolutions).
template <int BLOCK_N, int MT>
struct GemmCfg {
  static constexpr uint32_t A_BYTES = MT * GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES_EPI = GEMM_EPI_WARPS * 32 * 32 * 4;  // per-warp transpose staging
  static constexpr int STAGES = (A_BYTES + B_BYTES) >= 49152 ? 4 : ((A_BYTES + B_BYTES) >= 32768 ? 6 : 8);
  static constexpr uint32_t ACC_COLS = MT * BLOCK_N;                  // TMEM columns of one accumulator buffer
  static constexpr uint32_t TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr size_t SMEM_BYTES =
      1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + STAGE_BYTES_EPI + 256 /*barriers*/;
  static_assert(TMEM_COLS <= 512, "accumulators do not fit tensor memory");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// ------------------------------------------------------------------------------------------------
// epilogue
//
// An epilogue warp owns 32 accumulator rows (its TMEM lane quadrant) and walks its share of the columns in rounds
// of 32.  tcgen05.ld hands thread i row i of the round; most epilogues then transpose the 32x32 block through a
// warp-private, XOR-swizzled shared-memory tile so that thread i owns COLUMN n0+i instead: global loads and stores
// of a row then cover 64 (fp16) or 128 (fp32) contiguous bytes per instruction, per-column constants (bias, gate)
// are one register, and GroupNorm column sums need two shuffles.  The scatter epilogues whose destination is
// contiguous along rows (V^T, the piano roll, the latent) keep the row-per-thread orientation.
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == ACT_SILU) return silu_f(x);
  else if constexpr (ACT == ACT_GELU_TANH) return gelu_tanh_f(x);
  else return x;
}

// Fast column loops: the 32 rows of the round are all valid and their destination rows are equally spaced (`os`
// elements apart).  Every instruction issued here costs tensor-pipe time: with the epilogue switched off the same
// mainloop runs 30-40 % faster (profiles/r1_probe_bound.txt; tile time = MMA time + epilogue issue slots per SM
// sub-partition), so the loops are written for instruction count: the row stride is a template constant for the strides
// the path uses (stores and residual loads then take immediate offsets -- no address arithmetic at all), and the
// GroupNorm sums are taken from the fp32 value before rounding (no convert-back; the reference's statistics are fp32
// as well).  5 instructions per element (FFMA, F2F, STG, FADD, FFMA) instead of 10.
template <int ACT, bool RESID, int OS>
__device__ __forceinline__ void cols_f16_fast(const float (&val)[32], __half* op, const __half* rp, int os_dyn,
                                              float alpha, float bias, float& s, float& q, bool nostore) {
  const int os = OS > 0 ? OS : os_dyn;
  // The residual may alias the output (ResnetBlock shortcut written in place), so the compiler will not move a
  // residual load above an earlier store: issue all 32 loads first, then the dependent stores (measured: a
  // load -> store chain costs ~600 cycles per row).
  __half rv[32];
  if constexpr (RESID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) rv[j] = rp[j * os];
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = apply_act<ACT>(fmaf(val[j], alpha, bias));
    if constexpr (RESID) x += __half2float(rv[j]);
    if (!nostore) op[j * os] = __float2half_rn(x);  // (nostore: bound analysis only, GemmParams::debug bit 3)
    s += x;
    q = fmaf(x, x, q);
  }
}

// row-stride dispatch: the strides of the VAE levels (128 / 256 / 512 channels, x2 for the upsample scatter) and of the
// DiT MLP (4608); anything else takes the dynamic-stride instantiation (one IMAD.WIDE per element)
template <int ACT, bool RESID>
__device__ __forceinline__ void cols_f16_dispatch(const float (&val)[32], __half* op, const __half* rp, int os,
                                                  float alpha, float bias, float& s, float& q, bool nostore) {
  switch (os) {
    case 128: cols_f16_fast<ACT, RESID, 128>(val, op, rp, os, alpha, bias, s, q, nostore); break;
    case 256: cols_f16_fast<ACT, RESID, 256>(val, op, rp, os, alpha, bias, s, q, nostore); break;
    case 512: cols_f16_fast<ACT, RESID, 512>(val, op, rp, os, alpha, bias, s, q, nostore); break;
    case 1024: cols_f16_fast<ACT, RESID, 1024>(val, op, rp, os, alpha, bias, s, q, nostore); break;
    case 4608: cols_f16_fast<ACT, RESID, 4608>(val, op, rp, os, alpha, bias, s, q, nostore); break;
    default: cols_f16_fast<ACT, RESID, 0>(val, op, rp, os, alpha, bias, s, q, nostore); break;
  }
}

// Column-owner epilogue core: thread `lane` owns output feature n and the 32 accumulator rows val[0..31] of the round
// (rows row0..row0+31).  FROM_SMEM: the ragged/general path re-reads the values from the warp's staging tile
// (row-major kernel) instead of indexing the register array dynamically.
template <int EPI, bool FROM_SMEM>
__device__ __forceinline__ void epilogue_cols_core(const GemmParams& p, const float (&val)[32], uint32_t stage,
                                                   int lane, int row0, int orow_lane, bool uniform_rows, int orow0,
                                                   int orow_step, int n, int par, float& s, float& q) {
  const EpiParams& e = p.epi;
  const float bias = e.bias ? __ldg(e.bias + n) : 0.f;
  auto value = [&](int j) -> float {
    if constexpr (FROM_SMEM) return lds_f32(stage + 4u * (j * 32 + (lane ^ j)));
    else return val[j];
  };

  if constexpr (EPI == EPI_F16) {
    // s, q: running GroupNorm sums of this thread's feature over the rows it has stored (written by the caller)
    if (uniform_rows && e.addtab == nullptr && (e.resid == nullptr || e.ldr == e.ldo)) {
      __half* op = static_cast<__half*>(e.out) + (long long)orow0 * e.ldo + n;
      const int os = orow_step * e.ldo;
      const bool ns = (p.debug & 8) != 0;
      if (e.resid != nullptr) {
        const __half* rp = e.resid + (long long)orow0 * e.ldr + n;
        if (e.act == ACT_NONE) cols_f16_dispatch<ACT_NONE, true>(val, op, rp, os, e.alpha, bias, s, q, ns);
        else if (e.act == ACT_SILU) cols_f16_fast<ACT_SILU, true, 0>(val, op, rp, os, e.alpha, bias, s, q, ns);
        else cols_f16_fast<ACT_GELU_TANH, true, 0>(val, op, rp, os, e.alpha, bias, s, q, ns);
      } else {
        if (e.act == ACT_NONE) cols_f16_dispatch<ACT_NONE, false>(val, op, nullptr, os, e.alpha, bias, s, q, ns);
        else if (e.act == ACT_SILU) cols_f16_fast<ACT_SILU, false, 0>(val, op, nullptr, os, e.alpha, bias, s, q, ns);
        else cols_f16_dispatch<ACT_GELU_TANH, false>(val, op, nullptr, os, e.alpha, bias, s, q, ns);
      }
    } else {
      // general path: ragged last rows, 16-pixel-wide upsample rows, gathered row-vector add
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int orow = __shfl_sync(0xffffffffu, orow_lane, j);
        float x = value(j) * e.alpha + bias;
        if (orow >= 0) {
          if (e.addtab) x += __ldg(e.addtab + e.addidx[row0 + j] * (long long)p.N + n);
          if (e.act == ACT_SILU) x = silu_f(x);
          else if (e.act == ACT_GELU_TANH) x = gelu_tanh_f(x);
          if (e.resid) x += __half2float(e.resid[(long long)orow * e.ldr + n]);
          static_cast<__half*>(e.out)[(long long)orow * e.ldo + n] = __float2half_rn(x);
          s += x;
          q = fmaf(x, x, q);
        }
      }
    }
  } else if constexpr (EPI == EPI_F32) {
    if (uniform_rows && e.addtab == nullptr && e.act == ACT_NONE) {
      float* op = static_cast<float*>(e.out) + (long long)orow0 * e.ldo + n;
      const long long ostride = (long long)orow_step * e.ldo;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        *op = fmaf(val[j], e.alpha, bias);
        op += ostride;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int orow = __shfl_sync(0xffffffffu, orow_lane, j);
        float x = value(j) * e.alpha + bias;
        if (orow >= 0) {
          if (e.addtab) x += __ldg(e.addtab + e.addidx[row0 + j] * (long long)p.N + n);
          if (e.act == ACT_SILU) x = silu_f(x);
          else if (e.act == ACT_GELU_TANH) x = gelu_tanh_f(x);
          static_cast<float*>(e.out)[(long long)orow * e.ldo + n] = x;
        }
      }
    }
  } else if constexpr (EPI == EPI_GATE_RESID) {
    // x[row, n] += gate[sample, n] * (acc + bias[n]); the 32 rows of a round belong to one sample
    const float g = __ldg(e.gate + (long long)(row0 / e.rows_per_sample) * e.gate_ld + n);
    if (uniform_rows) {
      float* px = static_cast<float*>(e.out) + (long long)orow0 * e.ldo + n;
      auto rmw = [&](auto LD) {  // LD: compile-time row stride (immediate offsets) or 0 = e.ldo
        const int ld = decltype(LD)::value > 0 ? decltype(LD)::value : e.ldo;
        float xin[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) xin[j] = px[j * ld];
#pragma unroll
        for (int j = 0; j < 32; ++j) px[j * ld] = fmaf(g, val[j] + bias, xin[j]);
      };
      if (e.ldo == 1152) rmw(std::integral_constant<int, 1152>{});
      else rmw(std::integral_constant<int, 0>{});
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int orow = __shfl_sync(0xffffffffu, orow_lane, j);
        if (orow >= 0) {
          float* px = static_cast<float*>(e.out) + (long long)orow * e.ldo + n;
          *px = *px + g * (value(j) + bias);
        }
      }
    }
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    // q / k columns: + bias, rotary on interleaved pairs (partner column = lane ^ 1), head split
    const int D = e.heads * e.dh;
    const int which = n / D;
    const int rem = n - which * D;
    const int head = rem / e.dh;
    const int d = rem - head * e.dh;
    const bool rot = d < e.rot_dim;
    const int half_rot = e.rot_dim >> 1;
    const int b0 = row0 / e.T, tok0 = row0 - b0 * e.T;  // T is a multiple of 32: one sample per round
    __half* qp = (which == 0 ? e.q : e.k) + (((long long)b0 * e.heads + head) * e.T + tok0) * e.dh_pad + d;
    // This thread's frequency is fixed for the round, so the 32 table entries it needs are one rotation apart:
    // (cos, sin)((tok0 + j) w) = R(w)^j (cos, sin)(tok0 w).  ONE table load per round (+ the step (cos w, sin w) =
    // the table's row of token 1) and 4 FMAs per row replace 32 loads; the table (36 KB at T = 256) does not fit the
    // 28 KB of L1 left beside the operand ring, and those exposed L2 round trips made this epilogue 18 K cycles per
    // tile against 12.5 K for the MMAs (profiles/r1_trace_dit.txt).  Restarting from the table every 32 rows keeps the
    // recurrence's drift below 4e-6, under the table's own argument rounding (pos * freq in fp32).
    // Lanes outside the rotary dimensions rotate by the identity: uniform control flow.
    const int fi = rot ? (d >> 1) : 0;
    const float2 cs0 = __ldg(e.rope_cs + tok0 * half_rot + fi);
    const float2 stp = __ldg(e.rope_cs + half_rot + fi);  // token 1 (T >= 32 here)
    float c = rot ? cs0.x : 1.f, sn = rot ? cs0.y : 0.f;
    const float dc = rot ? stp.x : 1.f, ds = rot ? stp.y : 0.f;
    const float sgn = (d & 1) ? 1.f : -1.f;
    const int nrows = p.M - row0 < 32 ? p.M - row0 : 32;
    // y = x cos + sgn * partner * sin with the sign folded into the sine (ss = sgn * sin rotates like sin does when
    // the step's sine carries the same sign twice: ss' = ss dc + c (sgn ds); c' = c dc - ss (sgn ds))
    float ss = sgn * sn;
    const float dss = sgn * ds;
    auto rows = [&](auto DP, bool full) {  // DP: compile-time row pitch of q / k (immediate store offsets) or 0
      const int dp = decltype(DP)::value > 0 ? decltype(DP)::value : e.dh_pad;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = val[j] + bias;
        const float px = __shfl_xor_sync(0xffffffffu, x, 1);
        const float y = fmaf(px, ss, x * c);
        if (full || j < nrows) qp[j * dp] = __float2half_rn(y);
        const float c2 = fmaf(c, dc, -ss * dss);
        ss = fmaf(ss, dc, c * dss);
        c = c2;
      }
    };
    if (e.dh_pad == 72 && nrows == 32) rows(std::integral_constant<int, 72>{}, true);
    else rows(std::integral_constant<int, 0>{}, false);
  }
}

// Row-major kernel: tcgen05.ld hands thread i ROW i of the round; transpose the 32x32 block through the warp's
// XOR-swizzled staging tile (element (row, col) at row*32 + (col ^ row): conflict-free both ways) so that thread i
// owns column n0 + i, then run the column-owner core.
template <int EPI>
__device__ __forceinline__ void epilogue_round_cols(const GemmParams& p, uint32_t stage, const uint32_t (&r)[32],
                                                    int lane, int row0, int orow_lane, bool uniform_rows,
                                                    int orow0, int orow_step, int n0, int par) {
#pragma unroll
  for (int c = 0; c < 32; ++c) sts_f32(stage + 4u * (lane * 32 + (c ^ lane)), __uint_as_float(r[c]));
  __syncwarp();
  float val[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) val[j] = lds_f32(stage + 4u * (j * 32 + (lane ^ j)));
  float s = 0.f, q = 0.f;  // (the row-major kernel does not emit GroupNorm partials)
  epilogue_cols_core<EPI, true>(p, val, stage, lane, row0, orow_lane, uniform_rows, orow0, orow_step, n0 + lane, par, s, q);
  __syncwarp();  // the staging tile is rewritten by the next round
}

// row-per-thread epilogues: destination contiguous along rows (tokens / pixels)
template <int EPI>
__device__ __forceinline__ void epilogue_round_rows(const GemmParams& p, const uint32_t (&r)[32], int row, int n0) {
  const EpiParams& e = p.epi;
  if (row >= p.M) return;
  if constexpr (EPI == EPI_QKV_ROPE) {
    // V columns -> V^T [B, heads, dh, T]: it is the K-major B operand of the P.V MMA (attention.cu).  Consecutive
    // lanes hold consecutive tokens, so each 2-byte store instruction covers 64 contiguous bytes.
    const int D = e.heads * e.dh;
    const int b = row / e.T, tok = row - b * e.T;
    float bv[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + i);
      bv[4 * i] = t.x;
      bv[4 * i + 1] = t.y;
      bv[4 * i + 2] = t.z;
      bv[4 * i + 3] = t.w;
    }
    const int rem0 = n0 - 2 * D;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int rem = rem0 + i;
      const int head = rem / e.dh, d = rem - head * e.dh;
      e.v[(((long long)b * e.heads + head) * e.dh + d) * e.T + tok] = __float2half_rn(__uint_as_float(r[i]) + bv[i]);
    }
  } else if constexpr (EPI == EPI_UNPATCH) {
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
        out[(((long long)b * e.c_out + ch) * e.latH + time) * e.latW + part * patch + pl] =
            __uint_as_float(r[i]) + __ldg(e.bias + f);
      }
    }
  } else if constexpr (EPI == EPI_ROLL) {
    const int img = row >> 14, pix = row & 16383;  // 128 x 128 output pixels per VAE tile
    const int h = pix >> 7, w = pix & 127;         // h = pitch, w = time within the tile
    const int g = e.tile0 + img;                   // global tile index, tile-major: g = k * n_cand + cand
    const int kt = g / e.n_cand, cand = g - kt * e.n_cand;
    float* out = static_cast<float*>(e.out);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int ch = n0 + i;
      if (ch < e.roll_ch)
        out[(((long long)cand * e.roll_ch + ch) * 128 + h) * e.roll_len + kt * 128 + w] =
            __uint_as_float(r[i]) + __ldg(e.bias + ch);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..9 = epilogue.
// Persistent over tiles, double-buffered TMEM accumulator so the epilogue of tile i overlaps the mainloop of
// tile i+1.
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int MT, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, MT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t A_SUB = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  float* smem_epi = reinterpret_cast<float*>(smem_b + STAGES * Cfg::B_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_epi) + Cfg::STAGE_BYTES_EPI);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // epilogue warps split one accumulator buffer (ACC_COLS TMEM columns) into two column halves of CW columns
  constexpr int CW = (Cfg::ACC_COLS >= 64) ? Cfg::ACC_COLS / 2 : Cfg::ACC_COLS;
  constexpr int EPI_ACTIVE = (Cfg::ACC_COLS >= 64) ? GEMM_EPI_WARPS : 4;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_ACTIVE);
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

  const int num_ms = (p.num_m_tiles + MT - 1) / MT;  // M super-tiles of MT x 128 rows
  const int tiles_mn = num_ms * p.num_n_tiles;
  const int total_tiles = tiles_mn * p.num_par;
  const int num_kb = p.num_taps * p.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int par, tt;
        split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
        const int ms = tt / p.num_n_tiles;
        const int n_tile = tt - ms * p.num_n_tiles;
        const int brow = par * p.N + n_tile * BLOCK_N;
        int img[MT], h0[MT], w0[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m_tile = ms * MT + mt;  // may be == num_m_tiles for the odd tail: image index out of range -> zeros
          img[mt] = m_tile / p.tiles_per_img;
          const int rr = m_tile - img[mt] * p.tiles_per_img;
          h0[mt] = (rr / p.tiles_per_row) * p.bh * p.in_stride;
          w0[mt] = (rr % p.tiles_per_row) * p.bw * p.in_stride;
        }
        const int bz = p.b_batched ? img[0] : 0;
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 0] = clock64();
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int dy = p.tap_dy[par][tap], dx = p.tap_dx[par][tap];
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tma_load_4d(smem_a + stage * Cfg::A_BYTES + mt * A_SUB, &tmap_a, &full_bar[stage], kc * GEMM_BLOCK_K,
                          w0[mt] + dx, h0[mt] + dy, img[mt]);
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
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 2] = clock64();
          const uint64_t bdesc = umma_desc_sw128(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = umma_desc_sw128(smem_a + stage * Cfg::A_BYTES + mt * A_SUB);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              umma_f16(d_tmem + mt * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete → epilogue
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 3] = clock64();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp - 2 < EPI_ACTIVE) {
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int chalf = ew >> 2;   // which half of the accumulator buffer's columns
    const uint32_t stage_tile = smem_u32(smem_epi) + ew * 4096u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int par, tt;
      split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
      const int ms = tt / p.num_n_tiles;
      const int n_tile = tt - ms * p.num_n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / gridDim.x) * 8 + 4] = clock64();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll 1
      for (int c = 0; c < CW; c += 32) {
        const int acc_col = chalf * CW + c;
        const int mt = acc_col / BLOCK_N;
        const int col = acc_col - mt * BLOCK_N;
        const int m_tile = ms * MT + mt;
        if (m_tile >= p.num_m_tiles) continue;  // odd tail of a 2-sub-tile tile
        const int row0 = m_tile * GEMM_BLOCK_M + quad * 32;
        const int n0 = n_tile * BLOCK_N + col;
        uint32_t r[32];
        const bool tracing = p.trace && blockIdx.x == 0 && warp == 2;
        long long tq0 = 0;
        if (tracing) tq0 = clock64();
        tmem_ld_32x32(taddr + acc_col, r);
        tmem_ld_wait();
        if (tracing && lane == 0) p.trace[(t / gridDim.x) * 8 + 6] += clock64() - tq0;  // cycles inside tcgen05.ld
        if constexpr (EPI == EPI_UNPATCH || EPI == EPI_ROLL) {
          epilogue_round_rows<EPI>(p, r, row0 + lane, n0);
        } else {
          if constexpr (EPI == EPI_QKV_ROPE) {
            if (n0 >= 2 * p.epi.heads * p.epi.dh) {  // V third of the QKV output (uniform per round)
              epilogue_round_rows<EPI>(p, r, row0 + lane, n0);
              continue;
            }
          }
          // destination rows: lane i owns accumulator row row0 + i.  When all 32 rows are valid and equally spaced
          // at the destination (always, except ragged tails and 16-pixel-wide upsample rows) the column loops use
          // pointer increments; otherwise each row's destination is broadcast by shuffle.
          const int row = row0 + lane;
          int orow = row < p.M ? row : -1;
          bool uniform_rows = row0 + 32 <= p.M;
          int orow0 = row0, orow_step = 1;
          if (p.epi.up2) {
            const int hw = p.epi.upH * p.epi.upW;
            if (orow >= 0) {
              const int im = row / hw, rem = row - im * hw;
              const int h = rem / p.epi.upW, w = rem - h * p.epi.upW;
              orow = (im * (2 * p.epi.upH) + (2 * h + (par >> 1))) * (2 * p.epi.upW) + (2 * w + (par & 1));
            }
            uniform_rows = uniform_rows && (p.epi.upW % 32 == 0);
            orow0 = __shfl_sync(0xffffffffu, orow, 0);
            orow_step = 2;
          }
          epilogue_round_cols<EPI>(p, stage_tile, r, lane, row0, orow, uniform_rows, orow0, orow_step, n0, par);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / gridDim.x) * 8 + 5] = clock64();
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

// ================================================================================================================
// Feature-major ("swapped") kernel: the workhorse for every output whose feature count is a multiple of 128.
//
// The UMMA computes D^T: M = 128 output FEATURES (A operand = the weight tile, 16 KB per k-block), N = 256 output
// ROWS (B operand = the activation / im2col tile, 32 KB per k-block).  Compared with a 128-row x 128-feature tile this
// halves the shared-memory operand reads per flop (96 B/clk instead of 128 B/clk -- a 128x128 UMMA is smem-bound) and
// the L2 traffic per flop, for narrow layers (128-channel convolutions, N = 1152 / 3456 linears) as well as wide ones.
// In TMEM a lane is a feature and a column is a row, which is exactly the "thread owns a feature, walks rows" layout
// the coalesced epilogue wants: no shared-memory transposition, 64 / 128 contiguous bytes per store instruction.
// ================================================================================================================
constexpr int SW_ROWS = 256;      // output rows (pixels / tokens) per tile = UMMA N
constexpr int SW_FEATS = 128;     // output features per tile = UMMA M
constexpr int SW_STAGES = 4;
constexpr uint32_t SW_W_BYTES = SW_FEATS * GEMM_BLOCK_K * 2;  // 16 KB
constexpr uint32_t SW_X_BYTES = SW_ROWS * GEMM_BLOCK_K * 2;   // 32 KB
constexpr size_t SW_SMEM_BYTES = 1024 + SW_STAGES * (SW_W_BYTES + SW_X_BYTES) + 256;

// Epilogue of one feature-major tile for one epilogue warp: TMEM lanes quad*32.. = 32 features, TMEM columns
// rhalf*128.. = 128 of the tile's 256 rows, walked in rounds of 32 rows.
template <int EPI>
__device__ __forceinline__ void sw_epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int m_tile, int n_tile,
                                                 int par, int quad, int rhalf, int lane,
                                                 float2* stats = nullptr) {  // stats: this thread's (sum, sum of squares)
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + rhalf * 128;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;  // this thread's output feature
  float gs = 0.f, gq = 0.f;  // GroupNorm sum / sum of squares of feature n over this warp's 128 rows
#pragma unroll 1
  for (int c = 0; c < 128; c += 32) {
    const int row0 = m_tile * SW_ROWS + rhalf * 128 + c;
    if (row0 >= p.M) break;
    uint32_t r[32];
    tmem_ld_32x32(taddr + c, r);
    tmem_ld_wait();
    if (p.debug & 4) continue;  // bound analysis: TMEM reads only
    if constexpr (EPI == EPI_QKV_ROPE) {
      const int D = p.epi.heads * p.epi.dh;
      if (n_tile * SW_FEATS + quad * 32 >= 2 * D) {
        // V features -> V^T [B, heads, dh, T] (the K-major B operand of the P.V MMA): this thread holds 32
        // consecutive tokens of one (head, d) row = 64 contiguous bytes
        const EpiParams& e = p.epi;
        const int rem = n - 2 * D;
        const int head = rem / e.dh, d = rem - head * e.dh;
        const int b0 = row0 / e.T, tok0 = row0 - b0 * e.T;
        const float bias = __ldg(e.bias + n);
        uint4 pk[4];
        __half2* h2 = reinterpret_cast<__half2*>(pk);
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          h2[j >> 1] = __floats2half2_rn(__uint_as_float(r[j]) + bias, __uint_as_float(r[j + 1]) + bias);
        __half* dst = e.v + (((long long)b0 * e.heads + head) * e.dh + d) * e.T + tok0;
        if (row0 + 32 <= p.M) {
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst)[i] = pk[i];
        } else {
          const __half* hv = reinterpret_cast<const __half*>(pk);
          for (int j = 0; j < p.M - row0; ++j) dst[j] = hv[j];
        }
        continue;
      }
    }
    float val[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(r[j]);
    // destination rows (see epilogue_cols_core)
    const int row = row0 + lane;
    int orow = row < p.M ? row : -1;
    bool uniform_rows = row0 + 32 <= p.M;
    int orow0 = row0, orow_step = 1;
    if (p.epi.up2) {
      const int hw = p.epi.upH * p.epi.upW;
      if (orow >= 0) {
        const int im = row / hw, rem = row - im * hw;
        const int h = rem / p.epi.upW, w = rem - h * p.epi.upW;
        orow = (im * (2 * p.epi.upH) + (2 * h + (par >> 1))) * (2 * p.epi.upW) + (2 * w + (par & 1));
      }
      uniform_rows = uniform_rows && (p.epi.upW % 32 == 0);
      orow0 = __shfl_sync(0xffffffffu, orow, 0);
      orow_step = 2;
    }
    epilogue_cols_core<EPI, false>(p, val, 0u, lane, row0, orow, uniform_rows, orow0, orow_step, n, par, gs, gq);
  }
  if constexpr (EPI == EPI_F16) {
    const int rbase = m_tile * SW_ROWS + rhalf * 128;
    if (stats != nullptr) {
      *stats = make_float2(gs, gq);  // the caller publishes them (gn_dual_loop)
    } else if (p.epi.gn_part != nullptr && rbase < p.M) {
      // GroupNorm partials per (128 rows x 4 channels), no atomics (deterministic): lanes 4k..4k+3 hold one quad;
      // gn_finalize_kernel folds quads into groups and sums an image's slots in a fixed order
      gs += __shfl_xor_sync(0xffffffffu, gs, 1);
      gq += __shfl_xor_sync(0xffffffffu, gq, 1);
      gs += __shfl_xor_sync(0xffffffffu, gs, 2);
      gq += __shfl_xor_sync(0xffffffffu, gq, 2);
      if ((lane & 3) == 0) {
        const long long slot = (long long)par * p.slots_per_par + (rbase >> 7);
        reinterpret_cast<float2*>(p.epi.gn_part)[slot * (p.N >> 2) + (n >> 2)] = make_float2(gs, gq);
      }
    }
  }
}

// GroupNorm (+ swish) of the convolution's own output INSIDE its epilogue -- the normalise pass over the tensor
// (taming/modules/diffusionmodules/model.py:117-137: h = conv2(swish(norm2(conv1(...))))) never touches HBM.
//
// GroupNorm needs statistics over a whole image, i.e. over the tiles of up to 64 other CTAs, so the accumulator WAITS IN
// TENSOR MEMORY for them: (1) a first pass over the tile's TMEM columns takes the fp32 sums of this warp's 32 features x
// 128 rows and adds them, as fixed-point integers that also count arrivals, to the image's per-group accumulators;
// (2) the lanes poll their group's words until the image's other row tiles have contributed -- meanwhile the MMA warp is
// filling the second accumulator buffer with the next tile; (3) the totals give the affine (integer sums: deterministic
// whatever the arrival order); (4) a second pass over the same TMEM columns stores swish(a x + b) as fp16.  The raw
// convolution output is never written and never re-read: per element one 2-byte store instead of store + load + store.
//
// Progress: tiles are statically assigned (tile t -> CTA t mod grid), the host launches no more CTAs than are
// co-resident and only layers whose images span at most `grid` consecutive tiles, so two tiles of one image never sit on
// the same CTA and every wait is for tiles of the current or an earlier wave of other, running CTAs.  A wait that lasts
// ~1 s gives up and raises gn_err instead of hanging the device.
// y = swish(a2 r + b2) with the conv's alpha / bias folded into the GroupNorm affine by the caller
template <int LDO>
__device__ __forceinline__ void gn_store_cols(const uint32_t (&r)[32], __half* op, int ldo_dyn, float a2, float b2,
                                              bool swish) {
  const int ldo = LDO > 0 ? LDO : ldo_dyn;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float y = fmaf(a2, __uint_as_float(r[j]), b2);
    if (swish) y = swish_vae(y);
    op[j * ldo] = __float2half_rn(y);
  }
}

// Publish a warp's per-feature (sum, sum of squares) over its 128 rows to the image's per-group accumulators.
// Each (image, group) owns two 64-bit words: bits 63..8 accumulate the sum (sum of squares) as a 36.20 FIXED-POINT
// integer, bits 7..0 count the contributions.  One atomic add delivers a warp's partial AND its arrival, so no fence
// has to order data before a flag (a release cost ~3 K cycles per tile here), and integer addition is associative:
// the totals -- and the whole decode -- are bit-identical whatever order the CTAs arrive in.
__device__ __forceinline__ void gn_publish(const GemmParams& p, int m_tile, int n_tile, int quad, int lane, float gs,
                                           float gq) {
  const EpiParams& e = p.epi;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;
  // sums over this warp's channels of each GroupNorm group (4, 8 or 16 consecutive lanes)
  const int cpg = p.N >> 5;  // channels per group (32 groups)
  for (int o = 1; o < cpg; o <<= 1) {
    gs += __shfl_xor_sync(0xffffffffu, gs, o);
    gq += __shfl_xor_sync(0xffffffffu, gq, o);
  }
  const int img = m_tile / p.tiles_per_img;
  unsigned long long* gsum = e.gn_sums + ((long long)img * 32 + n / cpg) * 2;
  if ((lane & (cpg - 1)) == 0) {
    // range of the 36.20 format with up to 255 contributions: a warp's partial must stay below 2^27 (a group whose
    // values have an rms above ~250 over a whole image); beyond it the launch is flagged instead of wrapping silently
    if (fabsf(gs) > 1.0e8f || gq > 1.0e8f) *e.gn_err = 2;
    atomicAdd(gsum, ((unsigned long long)__double2ll_rn((double)gs * 1048576.0) << 8) + 1ull);
    atomicAdd(gsum + 1, ((unsigned long long)__double2ll_rn((double)gq * 1048576.0) << 8) + 1ull);
  }
}

// first pass: statistics of this warp's 32 features x 128 rows, published to the image's accumulators
__device__ __forceinline__ void gn_pass1(const GemmParams& p, uint32_t tmem_acc, int m_tile, int n_tile, int quad,
                                         int rhalf, int lane, unsigned long long* tr) {  // tr: trace slots 4, 6
  const EpiParams& e = p.epi;
  if (tr && lane == 0) tr[4] = clock64();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + rhalf * 128;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;  // this thread's output feature
  const float bias = e.bias ? __ldg(e.bias + n) : 0.f;
  // Sums of the RAW accumulators r (two instructions per element); the statistics of x = alpha r + bias follow in closed
  // form: sum x = alpha S + 128 bias, sum x^2 = alpha^2 Q + 2 alpha bias S + 128 bias^2.  (Every instruction of an
  // epilogue is paid in clock on this power-bound part.)  The host guarantees M % 256 == 0: every row is valid.
  // The load of the next 32 columns is in flight while these 32 are summed.
  float rs = 0.f, rq = 0.f;
  uint32_t ra[32], rb[32];
  tmem_ld_32x32(taddr, ra);
#pragma unroll
  for (int c = 0; c < 128; c += 64) {
    tmem_ld_wait();
    tmem_ld_32x32(taddr + c + 32, rb);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float r = __uint_as_float(ra[j]);
      rs += r;
      rq = fmaf(r, r, rq);
    }
    tmem_ld_wait();
    if (c + 64 < 128) tmem_ld_32x32(taddr + c + 64, ra);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float r = __uint_as_float(rb[j]);
      rs += r;
      rq = fmaf(r, r, rq);
    }
  }
  const float ab2 = e.alpha * bias;
  const float gs = fmaf(e.alpha, rs, 128.f * bias);
  const float gq = fmaf(e.alpha * e.alpha, rq, fmaf(2.f * ab2, rs, 128.f * bias * bias));
  gn_publish(p, m_tile, n_tile, quad, lane, gs, gq);
  if (tr && lane == 0) tr[6] = clock64();  // statistics published
}

// a group's complete totals -> the affine y = a x + b of one of its channels
__device__ __forceinline__ void gn_affine(const EpiParams& e, unsigned long long w0, unsigned long long w1, float gamma,
                                          float beta, float& a, float& b) {
  const double s = (double)((long long)w0 >> 8) * (1.0 / 1048576.0);
  const double q = (double)((long long)w1 >> 8) * (1.0 / 1048576.0);
  const double inv = (double)e.gn_inv_count;  // 1 / (pixels x channels per group)
  const double mean = s * inv;
  const double var = q * inv - mean * mean;
  const float rstd = rsqrtf(fmaxf((float)var, 0.f) + e.gn_eps);
  a = rstd * gamma;
  b = beta - (float)mean * a;
}

// whether the image's other row tiles have contributed to this lane's group (a word is complete when its count reaches
// the image's row tiles x 2, at most 128 < 256); warp-uniform result, the words in w0 / w1
__device__ __forceinline__ bool gn_ready(const GemmParams& p, int m_tile, int n_tile, int quad, int lane,
                                         unsigned long long& w0, unsigned long long& w1) {
  const EpiParams& e = p.epi;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;
  const int img = m_tile / p.tiles_per_img;
  const unsigned spi = (unsigned)(p.tiles_per_img * 2 * p.num_par);  // contributions per (image, group)
  const unsigned long long* gsum = e.gn_sums + ((long long)img * 32 + n / (p.N >> 5)) * 2;
  w0 = ld_relaxed_gpu_u64(gsum);
  w1 = ld_relaxed_gpu_u64(gsum + 1);
  const bool done = (((unsigned)w0 & 255u) == spi && ((unsigned)w1 & 255u) == spi) || (p.debug & 16);
  return __all_sync(0xffffffffu, done);
}

// second pass: normalise + activate + store from the same TMEM columns, given the group's complete totals
__device__ __forceinline__ void gn_pass2(const GemmParams& p, uint32_t tmem_acc, int m_tile, int n_tile, int quad,
                                         int rhalf, int lane, unsigned long long w0, unsigned long long w1,
                                         unsigned long long* tr) {  // tr: trace slots 7, 5
  const EpiParams& e = p.epi;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + rhalf * 128;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;
  const float bias = e.bias ? __ldg(e.bias + n) : 0.f;
  const float gamma = __ldg(e.gn_gamma + n), beta = __ldg(e.gn_beta + n);
  const int rbase = m_tile * SW_ROWS + rhalf * 128;
  if (tr && lane == 0) tr[7] = clock64();  // the image's statistics are complete
  float a, b;
  gn_affine(e, w0, w1, gamma, beta, a, b);
  __half* op = static_cast<__half*>(e.out) + (long long)rbase * e.ldo + n;
  const bool sw = e.gn_swish != 0;
  const float a2 = a * e.alpha, b2 = fmaf(a, bias, b);  // a (alpha r + bias) + b
  auto store = [&](const uint32_t (&r)[32], int c) {
    __half* oc = op + (long long)c * e.ldo;
    switch (e.ldo) {
      case 128: gn_store_cols<128>(r, oc, e.ldo, a2, b2, sw); break;
      case 256: gn_store_cols<256>(r, oc, e.ldo, a2, b2, sw); break;
      case 512: gn_store_cols<512>(r, oc, e.ldo, a2, b2, sw); break;
      default: gn_store_cols<0>(r, oc, e.ldo, a2, b2, sw); break;
    }
  };
  uint32_t ra[32], rb[32];  // the next 32 columns are in flight while these are normalised and stored
  tmem_ld_32x32(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < 128; c += 64) {
    tmem_ld_wait();
    tmem_ld_32x32(taddr + c + 32, rb);
    store(ra, c);
    tmem_ld_wait();
    if (c + 64 < 128) tmem_ld_32x32(taddr + c + 64, ra);
    store(rb, c + 32);
  }
  if (tr && lane == 0) tr[5] = clock64();
}

// The epilogue warps' tile loop of a GroupNorm-in-epilogue launch.  While a tile waits for its image's statistics the
// warps are idle (3-5 K cycles of other CTAs' skew), so the wait is OPPORTUNISTICALLY filled with the first pass of the
// next tile when that tile's accumulator is already complete (the epilogue-bound case: first pass 2 K + second pass 8 K
// cycles against the 12 K-cycle mainloop of the K = 1152 convolutions); when the MMAs are the bottleneck the next
// accumulator is not ready and nothing is reordered.  Tile k of this CTA lives in accumulator buffer k & 1.
// coords(t, m_tile, n_tile) maps a tile index; release(acc) frees a buffer.
template <class Coords, class Release>
__device__ __forceinline__ void gn_epilogue_loop(const GemmParams& p, uint32_t tmem_base, uint64_t* tmem_full, int first,
                                                 int stride, int total, int quad, int rhalf, int lane, bool tracer,
                                                 Coords coords, Release release) {
  int k = 0, m_tile = 0, n_tile = 0;
  int t = first;
  if (t >= total) return;
  unsigned long long* trace = (tracer && p.trace) ? p.trace : nullptr;
  coords(t, m_tile, n_tile);
  mbar_wait(&tmem_full[0], 0);
  tc_fence_after();
  gn_pass1(p, tmem_base, m_tile, n_tile, quad, rhalf, lane, trace);
  for (;;) {
    const int tn = t + stride, kn = k + 1;
    int m_next = 0, n_next = 0;
    bool next_done = tn >= total;  // nothing to pull forward after the last tile
    if (!next_done) coords(tn, m_next, n_next);
    unsigned long long w0, w1;
    const long long t0 = clock64();
    while (!gn_ready(p, m_tile, n_tile, quad, lane, w0, w1)) {
      if (!next_done) {
        const uint32_t ok = mbar_test(&tmem_full[kn & 1], (kn >> 1) & 1);
        if (__all_sync(0xffffffffu, ok != 0)) {
          tc_fence_after();
          gn_pass1(p, tmem_base + (kn & 1) * SW_ROWS, m_next, n_next, quad, rhalf, lane, trace ? trace + kn * 8 : nullptr);
          next_done = true;
          continue;
        }
      }
      if (clock64() - t0 > (1LL << 31)) {  // ~1 s: give up instead of hanging the device
        *p.epi.gn_err = 1;
        break;
      }
    }
    gn_pass2(p, tmem_base + (k & 1) * SW_ROWS, m_tile, n_tile, quad, rhalf, lane, w0, w1, trace ? trace + k * 8 : nullptr);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release(k & 1);
    if (tn >= total) break;
    if (!next_done) {
      mbar_wait(&tmem_full[kn & 1], (kn >> 1) & 1);
      tc_fence_after();
      gn_pass1(p, tmem_base + (kn & 1) * SW_ROWS, m_next, n_next, quad, rhalf, lane, trace ? trace + kn * 8 : nullptr);
    }
    t = tn;
    m_tile = m_next;
    n_tile = n_next;
    ++k;
  }
}

// ---- "dual" form: the RAW output is needed too (a ResnetBlock's output feeds the next block's shortcut as well as its
// norm1, model.py:117-137), so the tile is stored by the ordinary epilogue (residual, fp16 rounding, statistics of the
// fp32 values), its accumulator is released at once, and the normalised copy is written one tile LATER: by then the
// image's other tiles have contributed (no idle wait), and the warp re-reads its own 32 features x 128 rows -- still
// in L2 -- normalises them and stores them to gn_out2.  Compared with the separate pass this saves the DRAM read of
// the raw tensor and a launch, and never holds tensor memory.
// all 128 rows of the warp: the 32 loads of the next round are issued before this round is normalised and stored (an L2
// round trip per round would otherwise be exposed four times per tile)
template <int LDO>
__device__ __forceinline__ void gn_dual_cols(const __half* ip, __half* op, int ldo_dyn, long long round_stride, float a,
                                             float b, bool swish) {
  const int ldo = LDO > 0 ? LDO : ldo_dyn;
  __half v[2][32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[0][j] = __ldcg(ip + j * ldo);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c + 1 < 4) {
      const __half* in = ip + (c + 1) * round_stride;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[(c + 1) & 1][j] = __ldcg(in + j * ldo);
    }
    __half* oc = op + c * round_stride;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float y = fmaf(a, __half2float(v[c & 1][j]), b);
      if (swish) y = swish_vae(y);
      oc[j * ldo] = __float2half_rn(y);
    }
  }
}

// one round of 32 rows (the upsample conv's scattered rows)
__device__ __forceinline__ void gn_dual_round(const __half* ip, __half* op, int ldo, float a, float b, bool swish) {
  __half v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __ldcg(ip + j * ldo);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float y = fmaf(a, __half2float(v[j]), b);
    if (swish) y = swish_vae(y);
    op[j * ldo] = __float2half_rn(y);
  }
}

__device__ __forceinline__ void gn_dual_pass2(const GemmParams& p, int m_tile, int n_tile, int par, int quad, int rhalf,
                                              int lane) {
  const EpiParams& e = p.epi;
  const int n = n_tile * SW_FEATS + quad * 32 + lane;
  const float gamma = __ldg(e.gn_gamma + n), beta = __ldg(e.gn_beta + n);
  unsigned long long w0, w1;
  const long long t0 = clock64();
  while (!gn_ready(p, m_tile, n_tile, quad, lane, w0, w1)) {
    if (clock64() - t0 > (1LL << 31)) {  // ~1 s: give up instead of hanging the device
      *e.gn_err = 1;
      break;
    }
  }
  float a, b;
  gn_affine(e, w0, w1, gamma, beta, a, b);
  const bool sw = e.gn_swish != 0;
  if (e.up2) {
    // upsample conv: the 32 low-resolution pixels of a round (one image row segment: upW % 32 == 0) land on every
    // second pixel of output row 2h + (par >> 1), starting at column 2w + (par & 1) (sw_epilogue_tile's mapping)
    const int hw = e.upH * e.upW;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      const int row0 = m_tile * SW_ROWS + rhalf * 128 + c;
      const int im = row0 / hw, rem = row0 - im * hw;
      const int h = rem / e.upW, w = rem - h * e.upW;
      const long long orow0 = ((long long)im * (2 * e.upH) + (2 * h + (par >> 1))) * (2 * e.upW) + (2 * w + (par & 1));
      const long long o = orow0 * e.ldo + n;
      gn_dual_round(static_cast<const __half*>(e.out) + o, e.gn_out2 + o, 2 * e.ldo, a, b, sw);
    }
    return;
  }
  const long long off = (long long)(m_tile * SW_ROWS + rhalf * 128) * e.ldo + n;
  const __half* ip = static_cast<const __half*>(e.out) + off;
  __half* op = e.gn_out2 + off;
  const long long rs = 32LL * e.ldo;
  switch (e.ldo) {
    case 128: gn_dual_cols<128>(ip, op, e.ldo, rs, a, b, sw); break;
    case 256: gn_dual_cols<256>(ip, op, e.ldo, rs, a, b, sw); break;
    case 512: gn_dual_cols<512>(ip, op, e.ldo, rs, a, b, sw); break;
    default: gn_dual_cols<0>(ip, op, e.ldo, rs, a, b, sw); break;
  }
}

template <class Coords, class Release>
__device__ __forceinline__ void gn_dual_loop(const GemmParams& p, uint32_t tmem_base, uint64_t* tmem_full, int first,
                                             int stride, int total, int quad, int rhalf, int lane, Coords coords,
                                             Release release) {
  int pm = -1, pn = 0, pp = 0;  // the tile whose normalised copy is still owed
  int k = 0;
  for (int t = first; t < total; t += stride, ++k) {
    int m_tile, n_tile, par;
    coords(t, m_tile, n_tile, par);
    mbar_wait(&tmem_full[k & 1], (k >> 1) & 1);
    tc_fence_after();
    float2 st = make_float2(0.f, 0.f);
    sw_epilogue_tile<EPI_F16>(p, tmem_base + (k & 1) * SW_ROWS, m_tile, n_tile, par, quad, rhalf, lane, &st);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release(k & 1);
    gn_publish(p, m_tile, n_tile, quad, lane, st.x, st.y);
    if (pm >= 0) gn_dual_pass2(p, pm, pn, pp, quad, rhalf, lane);
    pm = m_tile;
    pn = n_tile;
    pp = par;
  }
  if (pm >= 0) gn_dual_pass2(p, pm, pn, pp, quad, rhalf, lane);
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sw_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + SW_STAGES * SW_W_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + SW_STAGES * SW_X_BYTES);
  uint64_t* empty_bar = full_bar + SW_STAGES;
  uint64_t* tmem_full = empty_bar + SW_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < SW_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // p.num_m_tiles counts 256-row tiles, p.num_n_tiles 128-feature tiles; feature tiles vary fastest so CTAs that run
  // together share the activation tile in L2
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int total_tiles = tiles_mn * p.num_par;
  const int num_kb = p.num_taps * p.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int par, tt;
        split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
        const int m_tile = tt / p.num_n_tiles;
        const int n_tile = tt - m_tile * p.num_n_tiles;
        const int img = m_tile / p.tiles_per_img;
        const int rr = m_tile - img * p.tiles_per_img;
        const int h0 = (rr / p.tiles_per_row) * p.bh * p.in_stride;
        const int w0 = (rr % p.tiles_per_row) * p.bw * p.in_stride;
        const int wrow = par * p.N + n_tile * SW_FEATS;
        const int bz = p.b_batched ? img : 0;
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 0] = clock64();
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int dy = p.tap_dy[par][tap], dx = p.tap_dx[par][tap];
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], SW_W_BYTES + SW_X_BYTES);
            tma_load_4d(smem_x + stage * SW_X_BYTES, &tmap_x, &full_bar[stage], kc * GEMM_BLOCK_K, w0 + dx, h0 + dy,
                        img);
            tma_load_3d(smem_w + stage * SW_W_BYTES, &tmap_w, &full_bar[stage],
                        (tap * p.kb_per_tap + kc) * GEMM_BLOCK_K, wrow, bz);
            if (++stage == SW_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(SW_FEATS, SW_ROWS);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * SW_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 2] = clock64();
          const uint64_t wdesc = umma_desc_sw128(smem_w + stage * SW_W_BYTES);
          const uint64_t xdesc = umma_desc_sw128(smem_x + stage * SW_X_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16(d_tmem, wdesc + 2 * k, xdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == SW_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (p.trace && blockIdx.x == 0) p.trace[(t / gridDim.x) * 8 + 3] = clock64();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quadrant = 32 features
    const int rhalf = ew >> 2;   // which 128 of the tile's 256 rows
    int acc = 0;
    uint32_t acc_phase = 0;
    if (EPI == EPI_F16 && p.epi.gn_sums != nullptr && p.epi.gn_out2 != nullptr) {
      gn_dual_loop(
          p, tmem_base, tmem_full, blockIdx.x, gridDim.x, total_tiles, quad, rhalf, lane,
          [&](int t, int& m_tile, int& n_tile, int& par) {
            int tt;
            split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
            m_tile = tt / p.num_n_tiles;
            n_tile = tt - m_tile * p.num_n_tiles;
          },
          [&](int a) { mbar_arrive(&tmem_empty[a]); });
    } else if (EPI == EPI_F16 && p.epi.gn_sums != nullptr) {  // (num_par == 1 in this mode)
      gn_epilogue_loop(
          p, tmem_base, tmem_full, blockIdx.x, gridDim.x, total_tiles, quad, rhalf, lane, blockIdx.x == 0 && warp == 2,
          [&](int t, int& m_tile, int& n_tile) {
            m_tile = t / p.num_n_tiles;
            n_tile = t - m_tile * p.num_n_tiles;
          },
          [&](int a) { mbar_arrive(&tmem_empty[a]); });
    } else
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int par, tt;
      split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
      const int m_tile = tt / p.num_n_tiles;
      const int n_tile = tt - m_tile * p.num_n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / gridDim.x) * 8 + 4] = clock64();
      sw_epilogue_tile<EPI>(p, tmem_base + acc * SW_ROWS, m_tile, n_tile, par, quad, rhalf, lane);
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / gridDim.x) * 8 + 5] = clock64();
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
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================================
// CTA-pair variant of the feature-major kernel: a 2-CTA cluster computes 256 features x 256 rows with
// tcgen05.mma.cta_group::2.  CTA r of the pair owns features [128 r, 128 r + 128) of the pair tile (its half of the M
// dimension: its own weight tile, its own TMEM lanes, its own epilogue) and loads rows [128 r, 128 r + 128) of the
// activation tile (its half of the shared B operand).  Per CTA and k-block that is 16 KB + 16 KB instead of
// 16 KB + 32 KB: a third less L2 -> shared-memory traffic and six pipeline stages instead of four in the same
// 192 KB, which is what the short-K (K = 1152) GEMMs were starved of (ncu: tensor pipe 63 % active, L2 41 % busy).
// ================================================================================================================
constexpr int SW2_STAGES = 6;
constexpr uint32_t SW2_W_BYTES = SW_FEATS * GEMM_BLOCK_K * 2;       // 16 KB: this CTA's 128 features
constexpr uint32_t SW2_X_BYTES = (SW_ROWS / 2) * GEMM_BLOCK_K * 2;  // 16 KB: this CTA's 128 of the 256 rows
constexpr size_t SW2_SMEM_BYTES = 1024 + SW2_STAGES * (SW2_W_BYTES + SW2_X_BYTES) + 256;

// Tile order of the pair kernel.  Pairs that run at the same time work on consecutive tile indices; with feature pairs
// fastest (band = all) the ~74 concurrent pairs share few row tiles and every feature slice, with a band of `gn`
// feature pairs they form a (74 / gn) x gn patch: each activation tile is read by gn pairs at once and each weight slice
// by 74 / gn.  tt -> (row tile, feature pair).
__device__ __forceinline__ void sw2_tile_coords(int tt, int num_m, int pairs_n, int gn, int& m_tile, int& n_pair) {
  if (gn <= 0 || gn >= pairs_n) {
    m_tile = tt / pairs_n;
    n_pair = tt - m_tile * pairs_n;
    return;
  }
  const int full = pairs_n / gn;
  const int band_tiles = num_m * gn;
  if (tt < full * band_tiles) {
    const int band = tt / band_tiles;
    const int r = tt - band * band_tiles;
    m_tile = r / gn;
    n_pair = band * gn + (r - m_tile * gn);
  } else {
    const int r = tt - full * band_tiles;
    const int g2 = pairs_n - full * gn;
    m_tile = r / g2;
    n_pair = full * gn + (r - m_tile * g2);
  }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_sw2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + SW2_STAGES * SW2_W_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + SW2_STAGES * SW2_X_BYTES);  // used in the leader only
  uint64_t* empty_bar = full_bar + SW2_STAGES;
  uint64_t* tmem_full = empty_bar + SW2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;                                                   // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < SW2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * GEMM_EPI_WARPS);  // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair tiles: (row tile, PAIR of feature tiles); an odd last feature tile is paired with out-of-range weight rows
  // (TMA zero-fill) whose epilogue is skipped
  const int pairs_n = (p.num_n_tiles + 1) >> 1;
  const int tiles_mn = p.num_m_tiles * pairs_n;
  const int total_tiles = tiles_mn * p.num_par;
  const int num_kb = p.num_taps * p.kb_per_tap;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int half_h = p.bh >> 1;  // images: this CTA's rows of the pixel box; 0 for linear layers (H == 1)

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < total_tiles; t += num_pairs) {
        int par, tt;
        split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
        int m_tile, n_pair;
        sw2_tile_coords(tt, p.num_m_tiles, pairs_n, p.band_n, m_tile, n_pair);
        const int n_tile = 2 * n_pair + (int)rank;
        const int img = m_tile / p.tiles_per_img;
        const int rr = m_tile - img * p.tiles_per_img;
        int h0 = (rr / p.tiles_per_row) * p.bh;
        int w0 = (rr % p.tiles_per_row) * p.bw;
        if (half_h > 0) h0 += (int)rank * half_h;   // lower / upper half of the pixel box
        else w0 += (int)rank * (SW_ROWS / 2);       // linear layer: second 128 rows
        const int wrow = par * p.N + n_tile * SW_FEATS;
        const int bz = p.b_batched ? img : 0;
        if (p.trace && blockIdx.x == 0) p.trace[(t / num_pairs) * 8 + 0] = clock64();
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int dy = p.tap_dy[par][tap], dx = p.tap_dx[par][tap];
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (p.debug & 1) {  // bound analysis: MMAs on stale shared memory, no operand traffic
              if (rank == 0) mbar_arrive(&full_bar[stage]);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (SW2_W_BYTES + SW2_X_BYTES));
              tma_load_4d_pair(smem_x + stage * SW2_X_BYTES, &tmap_x, &full_bar[stage], kc * GEMM_BLOCK_K, w0 + dx,
                               h0 + dy, img);
              tma_load_3d_pair(smem_w + stage * SW2_W_BYTES, &tmap_w, &full_bar[stage],
                               (tap * p.kb_per_tap + kc) * GEMM_BLOCK_K, wrow, bz);
            }
            if (++stage == SW2_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * SW_FEATS, SW_ROWS);  // M = 256 over the pair, N = 256
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair_id; t < total_tiles; t += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.trace && blockIdx.x == 0) p.trace[(t / num_pairs) * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * SW_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && p.trace && blockIdx.x == 0) p.trace[(t / num_pairs) * 8 + 2] = clock64();
          const uint64_t wdesc = umma_desc_sw128(smem_w + stage * SW2_W_BYTES);
          const uint64_t xdesc = umma_desc_sw128(smem_x + stage * SW2_X_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16_pair(d_tmem, wdesc + 2 * k, xdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);  // frees this stage in BOTH CTAs
          if (++stage == SW2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tmem_full[acc]);  // accumulators of both CTAs complete
        if (p.trace && blockIdx.x == 0) p.trace[(t / num_pairs) * 8 + 3] = clock64();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int rhalf = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (EPI == EPI_F16 && p.epi.gn_sums != nullptr && p.epi.gn_out2 != nullptr) {
      gn_dual_loop(
          p, tmem_base, tmem_full, pair_id, num_pairs, total_tiles, quad, rhalf, lane,
          [&](int t, int& m_tile, int& n_tile, int& par) {
            int tt, n_pair;
            split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
            sw2_tile_coords(tt, p.num_m_tiles, pairs_n, p.band_n, m_tile, n_pair);
            n_tile = 2 * n_pair + (int)rank;
          },
          [&](int a) {
            if (rank == 0) mbar_arrive(&tmem_empty[a]);
            else mbar_arrive_remote(&tmem_empty[a], 0);
          });
    } else if (EPI == EPI_F16 && p.epi.gn_sums != nullptr) {  // (num_par == 1, an even number of feature tiles in this mode)
      gn_epilogue_loop(
          p, tmem_base, tmem_full, pair_id, num_pairs, total_tiles, quad, rhalf, lane, blockIdx.x == 0 && warp == 2,
          [&](int t, int& m_tile, int& n_tile) {
            int n_pair;
            sw2_tile_coords(t, p.num_m_tiles, pairs_n, p.band_n, m_tile, n_pair);
            n_tile = 2 * n_pair + (int)rank;
          },
          [&](int a) {
            if (rank == 0) mbar_arrive(&tmem_empty[a]);
            else mbar_arrive_remote(&tmem_empty[a], 0);
          });
    } else
    for (int t = pair_id; t < total_tiles; t += num_pairs) {
      int par, tt;
      split_parity(t, p.num_par, tiles_mn, p.par_fast != 0, par, tt);
      int m_tile, n_pair;
      sw2_tile_coords(tt, p.num_m_tiles, pairs_n, p.band_n, m_tile, n_pair);
      const int n_tile = 2 * n_pair + (int)rank;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / num_pairs) * 8 + 4] = clock64();
      if (n_tile < p.num_n_tiles && !(p.debug & 2)) {
        sw_epilogue_tile<EPI>(p, tmem_base + acc * SW_ROWS, m_tile, n_tile, par, quad, rhalf, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[(t / num_pairs) * 8 + 5] = clock64();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_remote(&tmem_empty[acc], 0);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace rgm
