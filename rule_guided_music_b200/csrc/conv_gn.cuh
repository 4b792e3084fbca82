// 3x3 convolution with the GroupNorm apply + swish of its INPUT fused into the operand path (reference
// taming/modules/diffusionmodules/model.py:117-137: h = conv(swish(norm(x)))), for the 128-pixel-wide levels of the
// VAE (H = W = 128, 128 output features): the stand-alone normalise pass (one fp16 read + one fp16 write of the whole
// activation per GroupNorm, 96 ms of a 1.26 s step in round 1) disappears.
//
//   * one tile = 2 image rows x 128 pixels x 128 features.  Per 64-channel k-block TMA lands ONE raw halo tile
//     [4 rows][130 pixels][64 ch] (rows y0-1 .. y0+2, columns -1 .. 128; out-of-image pixels are TMA zero fill = the
//     convolution's zero padding) with the 128-byte swizzle, pixel-linear with pitch 130;
//   * four TRANSFORM warps normalise + activate it in place, shared -> shared: a thread owns one 8-channel chunk (its
//     2 x 8 GroupNorm coefficients live in registers, exactly gn_apply_kernel's arithmetic and rounding) and leaves
//     the out-of-image pixels zero (the reference pads the ACTIVATED tensor);
//   * the nine taps are nine SHIFTED VIEWS of that one buffer: tap (dy, dx) of output row j is the K-major operand that
//     starts at halo pixel (j + dy + 1) * 130 + dx + 1 -- a descriptor whose start address is advanced by whole 128-byte
//     rows (the hardware swizzle is a function of the absolute shared-memory address: profiles/r2_probe_shift_desc.txt),
//     one tcgen05.mma of N = 128 pixels per image row, M = 128 features (feature-major like gemm_sw_kernel, so the
//     epilogue -- bias, residual, fp16 store, GroupNorm partials of the OUTPUT -- is shared);
//   * the activation is read from L2 once per k-block instead of nine times, the weights stream through a 5-stage ring.
// Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2..9 epilogue, 10..13 transform.
#pragma once
#include "gemm_tc.cuh"

namespace rgm {

constexpr int CG_W = 128;                      // image width = pixels per MMA
constexpr int CG_ROWS = 2;                     // output rows per tile
constexpr int CG_HALO_W = CG_W + 2;            // 130
constexpr int CG_HALO_ROWS = CG_ROWS + 2;      // 4
constexpr int CG_HALO_PIX = CG_HALO_W * CG_HALO_ROWS;             // 520
constexpr uint32_t CG_HALO_BYTES = CG_HALO_PIX * 128;             // 66 560 = 65 KB (keeps 1024-byte alignment)
constexpr int CG_WSTAGES = 5;
constexpr uint32_t CG_W_BYTES = SW_FEATS * GEMM_BLOCK_K * 2;      // 16 KB: 128 features x 64 channels of one tap
constexpr int CG_TRANSFORM_WARPS = 4;
constexpr int CG_THREADS = 64 + 32 * GEMM_EPI_WARPS + 32 * CG_TRANSFORM_WARPS;  // 448
constexpr size_t CG_SMEM_BYTES = 1024 + 2 * CG_HALO_BYTES + CG_WSTAGES * CG_W_BYTES + 256;

__device__ __forceinline__ uint64_t umma_desc_sw128_addr(uint32_t a) {
  uint64_t d = static_cast<uint64_t>((a & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct ConvGnParams {
  const float2* in_ab;  // GroupNorm affine of the INPUT tensor: [image][Cin] (a, b), y = swish(a x + b)
  int H;                // image rows (multiple of 2); the width is CG_W
  int kb;               // Cin / 64
};

__global__ void __launch_bounds__(CG_THREADS, 1)
conv_gn_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ GemmParams p, const __grid_constant__ ConvGnParams cg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_halo = smem;
  uint8_t* smem_w = smem + 2 * CG_HALO_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + CG_WSTAGES * CG_W_BYTES);
  uint64_t* halo_full = bars;            // [2]    TMA landed the raw halo
  uint64_t* halo_ready = bars + 2;       // [2][2] transform done: halo rows 0-1 / rows 2-3
  uint64_t* halo_empty = bars + 6;       // [2]    the k-block's MMAs have read the buffer
  uint64_t* w_full = bars + 8;           // [CG_WSTAGES]
  uint64_t* w_empty = w_full + CG_WSTAGES;
  uint64_t* tmem_full = w_empty + CG_WSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_ready[2 * i], 32 * CG_TRANSFORM_WARPS);
      mbar_init(&halo_ready[2 * i + 1], 32 * CG_TRANSFORM_WARPS);
      mbar_init(&halo_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS);
    }
    for (int i = 0; i < CG_WSTAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
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

  const int tiles_per_img = cg.H / CG_ROWS;
  const int total_tiles = p.num_m_tiles;  // images x tiles_per_img
  const int KB = cg.kb;
  // development aid: CTA 0 writes clock64() at pipeline events, 32 slots per tile (tools/gpu_trace_conv_gn.py):
  // 0-1 halo TMA issued (k-block 0, 1); 2-4 / 5-7 transform start, rows 0-1 done, rows 2-3 done (k-block 0 / 1);
  // 8 accumulator free; 9-11 / 12-14 MMA: rows 0-1 ready, rows 2-3 ready, last issue (k-block 0 / 1); 16 epilogue
  // accumulator ready, 17 epilogue done.  p.debug bit 4: the transform warps skip the arithmetic (bound analysis).
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
#define CG_TRACE(slot) \
  if (tracing) p.trace[(t / gridDim.x) * 32 + (slot)] = clock64()

  if (warp == 0) {
    // Two independent TMA streams on two lanes of the producer warp (independent thread scheduling): the halo stream is
    // gated by the k-block buffers (halo_empty), the weight stream by its ring (w_empty).  On one thread the halo of
    // k-block g+1 would queue behind the last weight tiles of g, which wait for the MMAs of g to free ring slots --
    // and the transform of g+1 could not start until g was half done.
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int img = t / tiles_per_img;
        const int y0 = (t - img * tiles_per_img) * CG_ROWS;
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int buf = g & 1;
          mbar_wait(&halo_empty[buf], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&halo_full[buf], CG_HALO_BYTES);
          tma_load_4d(smem_halo + buf * CG_HALO_BYTES, &tmap_x, &halo_full[buf], kb * GEMM_BLOCK_K, -1, y0 - 1, img);
          if (kb < 2) CG_TRACE(kb);
        }
      }
    } else if (lane == 1) {
      int ws = 0;
      uint32_t wph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&w_empty[ws], wph ^ 1);
            mbar_arrive_expect_tx(&w_full[ws], CG_W_BYTES);
            tma_load_3d(smem_w + ws * CG_W_BYTES, &tmap_w, &w_full[ws], (tap * KB + kb) * GEMM_BLOCK_K, 0, 0);
            if (++ws == CG_WSTAGES) {
              ws = 0;
              wph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(SW_FEATS, CG_W);
      int ws = 0, acc = 0;
      uint32_t wph = 0, acc_phase = 0, g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        CG_TRACE(8);
        const uint32_t d_tmem = tmem_base + acc * (CG_ROWS * CG_W);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int buf = g & 1;
          const uint32_t hph = (g >> 1) & 1;
          const uint32_t hbase = smem_u32(smem_halo + buf * CG_HALO_BYTES);
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            // halo rows 0-1 serve dy = -1; output row 1 at dy = 0 is the first to read halo row 2
            if (tap == 0) {
              mbar_wait(&halo_ready[2 * buf], hph);
              tc_fence_after();
              if (kb < 2) CG_TRACE(9 + 3 * kb);
            } else if (tap == 3) {
              mbar_wait(&halo_ready[2 * buf + 1], hph);
              tc_fence_after();
              if (kb < 2) CG_TRACE(10 + 3 * kb);
            }
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            const uint64_t wdesc = umma_desc_sw128(smem_w + ws * CG_W_BYTES);
#pragma unroll
            for (int j = 0; j < CG_ROWS; ++j) {
              const uint64_t xdesc = umma_desc_sw128_addr(hbase + ((j + dy + 1) * CG_HALO_W + dx + 1) * 128);
#pragma unroll
              for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                umma_f16(d_tmem + j * CG_W, wdesc + 2 * k, xdesc + 2 * k, idesc, (kb | tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(&w_empty[ws]);
            if (++ws == CG_WSTAGES) {
              ws = 0;
              wph ^= 1;
            }
          }
          umma_commit(&halo_empty[buf]);
          if (kb < 2) CG_TRACE(11 + 3 * kb);
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 2 + GEMM_EPI_WARPS) {
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quadrant = 32 features
    const int rhalf = ew >> 2;   // which image row of the tile (128 of its 256 GEMM rows)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 && lane == 0) CG_TRACE(16);
      if (!(p.debug & 2)) sw_epilogue_tile<EPI_F16>(p, tmem_base + acc * (CG_ROWS * CG_W), t, 0, 0, quad, rhalf, lane);
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) CG_TRACE(17);
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // transform warps: thread = (pixel lane, logical 16-byte chunk); 16 pixels x 8 chunks per pass
    const int tt = threadIdx.x - 32 * (2 + GEMM_EPI_WARPS);
    const int chunk = tt & 7;
    const int prow = tt >> 3;
    const int Cin = KB * GEMM_BLOCK_K;
    uint32_t g = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int img = t / tiles_per_img;
      const int y0 = (t - img * tiles_per_img) * CG_ROWS;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int buf = g & 1;
        float a[8], b[8];
        {
          const float4* abp =
              reinterpret_cast<const float4*>(cg.in_ab + (long long)img * Cin + kb * GEMM_BLOCK_K + chunk * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = __ldg(abp + j);  // (a0, b0, a1, b1)
            a[2 * j] = v.x;
            b[2 * j] = v.y;
            a[2 * j + 1] = v.z;
            b[2 * j + 1] = v.w;
          }
        }
        mbar_wait(&halo_full[buf], (g >> 1) & 1);
        if (tt == 0 && kb < 2) CG_TRACE(2 + 3 * kb);
        const uint32_t base = smem_u32(smem_halo + buf * CG_HALO_BYTES);
        int hr = 0, hx = prow;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int pend = (half + 1) * 2 * CG_HALO_W;  // pixels of halo rows 0-1, then rows 2-3
#pragma unroll 2
          for (int px = half * 2 * CG_HALO_W + ((prow - half * 2 * CG_HALO_W) & 15); px < pend; px += 16) {
            // (hr, hx) of px, kept incrementally
            const int gy = y0 - 1 + hr, gx = hx - 1;
            if (static_cast<unsigned>(gy) < static_cast<unsigned>(cg.H) && static_cast<unsigned>(gx) < CG_W &&
                !(p.debug & 16)) {
              const uint32_t addr = base + px * 128 + ((chunk ^ (px & 7)) << 4);
              uint4 u = lds_v4(addr);
              __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                const float v0 = swish_vae(fmaf(a[2 * j], f.x, b[2 * j]));
                const float v1 = swish_vae(fmaf(a[2 * j + 1], f.y, b[2 * j + 1]));
                h2[j] = __floats2half2_rn(v0, v1);
              }
              sts_v4(addr, u);
            }
            hx += 16;
            if (hx >= CG_HALO_W) {
              hx -= CG_HALO_W;
              ++hr;
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's (async-proxy) reads
          mbar_arrive(&halo_ready[2 * buf + half]);
          if (tt == 0 && kb < 2) CG_TRACE(3 + 3 * kb + half);
        }
      }
    }
  }

#undef CG_TRACE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace rgm
