// Multi-head self-attention of the DiT block on tcgen05 (reference guided_diffusion/dit.py:263-288, the
// F.scaled_dot_product_attention branch): out = softmax(q k^T * scale) v per (sample, head), no mask.
//
// One persistent CTA per SM walks over (sample, head) pairs.  Per pair the whole K [T, dh] and V^T [dh, T] sit in
// shared memory (TMA, 128-byte swizzle), S = Q K^T for each 128-query tile is accumulated in TMEM (T <= 256 fp32
// columns per tile, two tiles = all 512 columns), four softmax warps (one thread per query row = one TMEM lane) read
// S, write un-normalised P = exp2((s - max) * scale*log2e) as fp16 into a swizzled K-major smem tile, the MMA thread
// runs O = P V into the TMEM columns S occupied, and the softmax warps scale O by 1/rowsum on the way out.
// The two query tiles are software-pipelined: the tensor core computes S1 and P0 V while the softmax warps work on
// tile 0 / tile 1, so the softmax (128 x T exponentials per tile) overlaps the MMAs.
// EIGHT softmax warps: warps w and w + 4 share a TMEM lane quadrant (the same 32 query rows) and each takes half of the
// columns; they exchange their row maxima and row sums through shared memory under a 64-thread named barrier.  With
// one warp per scheduler the softmax was a chain of exposed tcgen05.ld / MUFU / st.shared latencies (49 ms per step
// against an 11 ms MUFU floor); two warps per scheduler overlap each other's latencies and halve the per-warp work.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "api_util.h"
#include "aux_kernels.h"
#include "ptx.cuh"

namespace rgm {

namespace {

constexpr int ATT_SOFTMAX_WARPS = 8;
constexpr int ATT_THREADS = 32 + 32 * ATT_SOFTMAX_WARPS;  // warp 0: TMA + MMA issue (+ TMEM owner); warps 1..8: softmax / epilogue

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct AttnParams {
  int n_pairs;   // B * heads
  int heads, T, dh;
  int nkb;       // 64-wide k-blocks of the head dimension (1 or 2)
  int ksteps;    // ceil(dh / 16) MMA K steps for S
  int npv;       // dh rounded up to 16: N of the P V MMA
  int n_tiles;   // ceil(T / 128) query tiles; T is a multiple of 64 (the last tile may be half full)
  float scale_log2e;
  __half* out;   // [B*T, heads*dh]
  // development aid: CTA 0 writes clock64() at pipeline events of each of its pairs, 16 slots per pair
  // (tools/gpu_trace_attention.py): MMA thread 0 pair start (TMEM free), 1 operands A landed, 2 S tiles issued,
  // 3 P0 seen, 4 PV0 issued, 5 P1 seen, 6 PV1 issued; softmax warp 1: 8 S0 ready, 9 max pass 0 done, 10 P0 written,
  // 11 S1 ready, 12 P1 written, 13 O0 ready, 14 epilogue done
  unsigned long long* trace;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int T = p.T;
  const uint32_t q_bytes = p.nkb * 16384u;           // per query tile: nkb blocks of [128 rows][128 B]
  // one k-block of K: [n_tiles * 128 rows][128 B] -- K arrives in 128-row boxes, so the buffer holds whole boxes; rows
  // past T (the next pair's, or zero fill at the end of the tensor) are never read: the S MMA has N = T
  const uint32_t k_blk = (uint32_t)p.n_tiles * 16384u;
  const uint32_t v_blk = (uint32_t)p.npv * 128u;     // one 64-token block of V^T: [npv rows][128 B]
  uint8_t* sQ = smem;                                // query tile 0 (tile 1 is staged in the P buffer, see below)
  uint8_t* sK = sQ + q_bytes;
  uint8_t* sV = sK + p.nkb * k_blk;
  uint8_t* sP = sV + (T / 64) * v_blk;               // [T/64 blocks][128 rows][128 B]
  // Query tile 1 lives at the start of the P buffer until S1 = Q1 K^T has completed: P0 is not written before that
  // (the softmax warps wait on bar_s[1] first), so 227 KB of shared memory hold Q0, K, V^T and P.
  uint8_t* sQ1 = sP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (T / 64) * 16384u);
  uint64_t* bar_load = bars;        // group A landed: Q tile 0 + K (prefetched for the next pair as soon as S is done)
  uint64_t* bar_load_b = bars + 9;  // group B landed: Q tile 1 (staged in the P buffer) + V^T
  uint64_t* bar_s = bars + 1;       // [2] S tile complete
  uint64_t* bar_p = bars + 3;       // [2] P tile written (128 arrivals)
  uint64_t* bar_o = bars + 5;       // [2] O tile complete
  uint64_t* bar_done = bars + 7;    // epilogue has drained TMEM (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* red = reinterpret_cast<float*>(bars + 12);  // [column half][row]: partner exchange of row max, then row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(bar_load, 1);
    mbar_init(bar_load_b, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 32 * ATT_SOFTMAX_WARPS);
      mbar_init(&bar_o[i], 1);
    }
    mbar_init(bar_done, 32 * ATT_SOFTMAX_WARPS);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc_s = umma_idesc_f16(128, T);
  const uint32_t idesc_o = umma_idesc_f16(128, p.npv);
  const uint32_t bytes_a = q_bytes + p.nkb * k_blk;
  const uint32_t bytes_b = (p.n_tiles - 1) * q_bytes + (T / 64) * v_blk;
  // Operand loads of a pair arrive at HBM speed (170 KB per pair at 16 x 72: ~6 K cycles per SM when all SMs load at
  // once), so they are issued ahead: Q0 + K of the NEXT pair as soon as this pair's S MMAs have read them, Q1 + V^T as
  // soon as this pair's P V MMAs are done with the P buffer and V^T.
  auto issue_a = [&](int pr) {
    mbar_arrive_expect_tx(bar_load, bytes_a);
    const int row0 = pr * T;
    for (int kb = 0; kb < p.nkb; ++kb) {
      tma_load_2d(sQ + kb * 16384, &map_q, bar_load, kb * 64, row0);
      for (int j = 0; j < p.n_tiles; ++j)
        tma_load_2d(sK + kb * k_blk + j * 16384, &map_k, bar_load, kb * 64, row0 + j * 128);
    }
  };
  auto issue_b = [&](int pr) {
    mbar_arrive_expect_tx(bar_load_b, bytes_b);
    const int row0 = pr * T;
    if (p.n_tiles > 1)
      for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sQ1 + kb * 16384, &map_q, bar_load_b, kb * 64, row0 + 128);
    for (int tb = 0; tb < T / 64; ++tb) tma_load_2d(sV + tb * v_blk, &map_v, bar_load_b, tb * 64, pr * p.dh);
  };
  if (warp == 0 && lane == 0 && (int)blockIdx.x < p.n_pairs) {
    issue_a(blockIdx.x);
    issue_b(blockIdx.x);
  }

  uint32_t it = 0;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
#define ATT_TRACE(slot) \
  if (tracing) p.trace[it * 16 + (slot)] = clock64()
  for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x, ++it) {
    const uint32_t ph = it & 1;
    if (warp == 0) {
      if (lane == 0) {
        // the epilogue of the previous pair has drained TMEM (bar_done): the S / O columns are free
        const int next = pair + gridDim.x;
        if (it > 0) mbar_wait(bar_done, ph ^ 1);
        tc_fence_after();
        ATT_TRACE(0);
        mbar_wait(bar_load, ph);
        tc_fence_after();
        ATT_TRACE(1);
        // S tiles (tile 1 needs Q1 from group B)
        for (int m = 0; m < p.n_tiles; ++m) {
          if (m == 1) {
            mbar_wait(bar_load_b, ph);
            tc_fence_after();
          }
          for (int ks = 0; ks < p.ksteps; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            umma_f16(tmem_base + m * 256, umma_desc_sw128((m ? sQ1 : sQ) + kb * 16384) + 2 * kk,
                     umma_desc_sw128(sK + kb * k_blk) + 2 * kk, idesc_s, ks != 0);
          }
          umma_commit(&bar_s[m]);
        }
        ATT_TRACE(2);
        // Q0 and K have been read once the last S tile is complete: prefetch the next pair's
        if (next < p.n_pairs) {
          mbar_wait(&bar_s[p.n_tiles - 1], ph);
          issue_a(next);
        }
        if (p.n_tiles == 1) {  // V^T (group B) has not been waited for yet
          mbar_wait(bar_load_b, ph);
          tc_fence_after();
        }
        // O tiles: P (A operand, K = tokens) x V^T (B operand, [npv rows][tokens])
        for (int m = 0; m < p.n_tiles; ++m) {
          mbar_wait(&bar_p[m], ph);
          tc_fence_after();
          ATT_TRACE(3 + 2 * m);
          for (int ks = 0; ks < T / 16; ++ks) {
            const int tb = ks >> 2, kk = ks & 3;
            umma_f16(tmem_base + m * 256, umma_desc_sw128(sP + tb * 16384) + 2 * kk,
                     umma_desc_sw128(sV + tb * v_blk) + 2 * kk, idesc_o, ks != 0);
          }
          umma_commit(&bar_o[m]);
          ATT_TRACE(4 + 2 * m);
        }
        // the P buffer (where Q1 is staged) and V^T are free once the last P V tile is complete
        if (next < p.n_pairs) {
          mbar_wait(&bar_o[p.n_tiles - 1], ph);
          issue_b(next);
        }
      }
      __syncwarp();
    } else {
      const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
      const int half = (warp - 1) >> 2;          // which half of the columns (warps w and w + 4 share the rows)
      const int r = quad * 32 + lane;            // query row within the tile
      const int cw = T >> 1;                     // columns per warp
      const int c_lo = half * cw;
      float inv_sum[2] = {0.f, 0.f};
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m >= p.n_tiles) break;
        mbar_wait(&bar_s[m], ph);
        tc_fence_after();
        if (threadIdx.x == 32) ATT_TRACE(8 + 3 * m);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + m * 256;
        float mx = -INFINITY;
        for (int c = c_lo; c < c_lo + cw; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        red[half * 128 + r] = mx;
        named_bar_sync(1 + quad, 64);
        mx = fmaxf(mx, red[(half ^ 1) * 128 + r]);
        named_bar_sync(1 + quad, 64);  // both maxima read: the slots are free for the sums
        // the single P buffer is read by the previous tile's P V MMAs: wait for them before overwriting it
        if (m > 0) mbar_wait(&bar_o[m - 1], ph);
        else if (p.n_tiles > 1) mbar_wait(&bar_s[1], ph);  // Q1 is staged in the P buffer until S1 is done
        if (threadIdx.x == 32 && m == 0) ATT_TRACE(9);
        const float mneg = -mx * p.scale_log2e;
        float sum = 0.f;
        for (int c = c_lo; c < c_lo + cw; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
          uint4 pk[4];
          __half2* h2 = reinterpret_cast<__half2*>(pk);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2e, mneg));
            const float e1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, mneg));
            sum += e0 + e1;
            h2[i >> 1] = __floats2half2_rn(e0, e1);
          }
          // columns c..c+31 = 16-byte chunks (c%64)/8 .. +3 of row r in token block c/64, 128-byte swizzle
          const uint32_t rowp = smem_u32(sP) + (c >> 6) * 16384 + r * 128;
          const int ch0 = (c & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) sts_v4(rowp + (((ch0 + j) ^ (r & 7)) << 4), pk[j]);
        }
        red[half * 128 + r] = sum;
        tc_fence_before();
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&bar_p[m]);
        if (threadIdx.x == 32) ATT_TRACE(10 + 2 * m);
        named_bar_sync(1 + quad, 64);
        inv_sum[m] = 1.0f / (sum + red[(half ^ 1) * 128 + r]);
        named_bar_sync(1 + quad, 64);  // both sums read before the next tile's maxima overwrite the slots
      }
      // epilogue: O / rowsum -> out[b*T + row, head*dh + d]; the 16-column groups of O alternate between the two warps
      const int b = pair / p.heads, head = pair - b * p.heads;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m >= p.n_tiles) break;
        mbar_wait(&bar_o[m], ph);
        tc_fence_after();
        if (threadIdx.x == 32 && m == 0) ATT_TRACE(13);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + m * 256;
        __half* orow = p.out + ((long long)b * T + m * 128 + r) * (p.heads * p.dh) + head * p.dh;
        const bool row_ok = m * 128 + r < T;  // T = 64 or 192: the last query tile is half full
        for (int c = half * 16; c < p.npv; c += 32) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c, v);
          tmem_ld_wait();
          if (!row_ok) continue;
          uint4 pk[2];
          __half2* h2 = reinterpret_cast<__half2*>(pk);
#pragma unroll
          for (int i = 0; i < 16; i += 2)
            h2[i >> 1] = __floats2half2_rn(__uint_as_float(v[i]) * inv_sum[m], __uint_as_float(v[i + 1]) * inv_sum[m]);
          if (c + 8 <= p.dh) *reinterpret_cast<uint4*>(orow + c) = pk[0];
          if (c + 16 <= p.dh) *reinterpret_cast<uint4*>(orow + c + 8) = pk[1];
        }
      }
      tc_fence_before();
      mbar_arrive(bar_done);
      if (threadIdx.x == 32) ATT_TRACE(14);
    }
  }
#undef ATT_TRACE

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

bool map2d(CUtensorMap* m, const void* ptr, unsigned long long cols, unsigned long long rows, unsigned box_c,
           unsigned box_r) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_c, box_r};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

std::atomic<unsigned long long> g_att_launches{0};

}  // namespace

unsigned long long attention_launch_count() { return g_att_launches.load(); }

cudaError_t launch_attention(const __half* q, const __half* k, const __half* vt, __half* out, int B, int heads, int T,
                             int dh, float scale, cudaStream_t s, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = m;
    return cudaErrorInvalidValue;
  };
  if (T < 64 || T > 256 || T % 64 != 0) return fail("attention: T must be 64, 128, 192 or 256 tokens");
  if (dh % 8 != 0 || dh > 128 || dh < 16) return fail("attention: head_dim must be a multiple of 8 in [16, 128]");
  AttnParams p;
  p.n_pairs = B * heads;
  p.heads = heads;
  p.T = T;
  p.dh = dh;
  p.nkb = (dh + 63) / 64;
  p.ksteps = (dh + 15) / 16;
  p.npv = ((dh + 15) / 16) * 16;
  p.n_tiles = (T + 127) / 128;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = out;
  p.trace = nullptr;
  if (const char* tp = getenv("RGM_DEBUG_TRACE_PTR")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0));
  CUtensorMap mq, mk, mv;
  const unsigned long long rows = (unsigned long long)B * heads * T;
  if (!map2d(&mq, q, dh, rows, 64, 128) || !map2d(&mk, k, dh, rows, 64, 128) ||
      !map2d(&mv, vt, T, (unsigned long long)B * heads * dh, 64, p.npv))
    return fail("attention: cuTensorMapEncodeTiled failed");
  const size_t p_bytes = (size_t)(T / 64) * 16384 > (size_t)(p.n_tiles - 1) * p.nkb * 16384
                             ? (size_t)(T / 64) * 16384 : (size_t)p.nkb * 16384;  // P blocks; query tile 1 is staged there
  const size_t smem = 1024 + p.nkb * 16384 + (size_t)p.nkb * p.n_tiles * 16384 + (size_t)(T / 64) * p.npv * 128 + p_bytes +
                      128 + 2 * 128 * sizeof(float);  // (barriers + slots: 96 B of the 128)
  static SmemAttr attr;
  if (cudaError_t e = attr.ensure(attention_kernel, smem); e != cudaSuccess) {
    if (err) *err = std::string("attention: cudaFuncSetAttribute(shared memory): ") + cudaGetErrorString(e);
    return e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_pairs < sms ? p.n_pairs : sms;
  const double fl = 4.0 * (double)B * heads * T * T * dh;
  ProfScope prof("attention", fl, 4.0 * (double)B * heads * T * T * p.npv, 0.0, s);
  attention_kernel<<<grid, ATT_THREADS, smem, s>>>(mq, mk, mv, p);
  g_att_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t st = cudaGetLastError();
  if (st != cudaSuccess && err) *err = std::string("attention launch: ") + cudaGetErrorString(st);
  return st;
}

}  // namespace rgm
