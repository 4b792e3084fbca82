// Rule evaluators and SCG candidate selection as warp-shuffle reductions over the decoded piano roll.
// Reference: music_rule_guidance/music_rules.py (piano_like :23-26, total_pitch_class_histogram :29-43,
// note_density :46-83), rule_maps.py:17-26 (losses), guided_diffusion/gaussian_diffusion.py:531-554 (weighted sum,
// first-max argmax over the N candidates, gather of the winner).
//
// Like the reference, the evaluators WRITE THROUGH channel 0 of the roll (pitch mask, -0.95 threshold), so a later
// rule sees what an earlier rule left behind; thresholds and counts are exact integer work, the histogram is fp32.
#include <atomic>

#include "api_util.h"
#include "aux_kernels.h"

namespace rgm {

namespace {
std::atomic<unsigned long long> g_rule_launches{0};
inline cudaError_t done() {
  g_rule_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
constexpr int MIN_PIANO = 21, MAX_PIANO = 108;
}  // namespace

unsigned long long rules_launch_count() { return g_rule_launches.load(); }

// grid = n candidates, 256 threads: warp w sums pitches w, w+8, ... over time; 12 threads fold pitch mod 12.
__global__ void __launch_bounds__(256) pitch_hist_kernel(float* __restrict__ roll, float* __restrict__ hist, int ch,
                                                         int L) {
  __shared__ float per_pitch[132];
  const int cand = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = roll + (long long)cand * ch * 128 * L;  // channel 0
  if (threadIdx.x < 4) per_pitch[128 + threadIdx.x] = 0.f;
  for (int pitch = warp; pitch < 128; pitch += 8) {
    float4* row = reinterpret_cast<float4*>(base + (long long)pitch * L);
    float s = 0.f;
    if (pitch < MIN_PIANO || pitch > MAX_PIANO) {
      for (int i = lane; i < L / 4; i += 32) row[i] = make_float4(-1.f, -1.f, -1.f, -1.f);  // piano_like, in place
    } else {
      for (int i = lane; i < L / 4; i += 32) {
        const float4 v = row[i];
        s += ((v.x + 1.f) * 0.5f + (v.y + 1.f) * 0.5f) + ((v.z + 1.f) * 0.5f + (v.w + 1.f) * 0.5f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) per_pitch[pitch] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float h = 0.f;
    if (lane < 12)
      for (int k = 0; k < 11; ++k) h += per_pitch[12 * k + lane];
    float tot = h;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane < 12) hist[cand * 12 + lane] = h / (tot + 1e-12f);
  }
}

cudaError_t launch_rule_pitch_hist(float* roll, float* hist, int n, int ch, int L, cudaStream_t s) {
  if (L % 4 != 0 || n <= 0) return cudaErrorInvalidValue;
  ProfScope prof("rule_pitch_hist", 0, 0, (double)n * 128.0 * L * 4.0, s);
  pitch_hist_kernel<<<n, 256, 0, s>>>(roll, hist, ch, L);
  return done();
}

// grid = (L/128 column blocks, n candidates), 128 threads = 128 consecutive time columns (coalesced rows).
// A thread walks the 128 pitches of its column: piano mask and the -0.95 threshold are written back in place,
// notes are counted, and an onset is a note whose left neighbour (zero-padded at column -1) is not a note.
__global__ void __launch_bounds__(128) note_density_kernel(float* __restrict__ roll, float* __restrict__ out, int ch,
                                                           int L, int interval, float hscale) {
  const int cand = blockIdx.y;
  const int col = blockIdx.x * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float* base = roll + (long long)cand * ch * 128 * L;
  int count = 0, onset = 0;
  for (int pitch = 0; pitch < 128; ++pitch) {
    float* row = base + (long long)pitch * L;
    const bool in_range = pitch >= MIN_PIANO && pitch <= MAX_PIANO;
    float v = row[col];
    float left = (lane == 0 && col > 0) ? row[col - 1] : 0.f;
    if (!in_range) v = -1.f;
    // `pr[pr < -0.95] = -1`, then (pr+1)/2 >= 0.01: for v >= -0.95f, (v+1)/2 >= 0.02499 >= 0.01, so note <=> v >= -0.95f
    const bool note = in_range && (v >= -0.95f);
    if (!note) v = -1.f;
    row[col] = v;
    const unsigned notes = __ballot_sync(0xffffffffu, note);
    bool left_note;
    if (lane == 0) left_note = (col > 0) && in_range && (left >= -0.95f);
    else left_note = (notes >> (lane - 1)) & 1u;
    count += note ? 1 : 0;
    onset |= (note && !left_note) ? 1 : 0;
  }
  // per-window sums: windows are `interval` columns wide, interval a power of two in [1, 128] or a multiple of 128
  __shared__ int s_cnt[128], s_on[128];
  s_cnt[threadIdx.x] = count;
  s_on[threadIdx.x] = onset;
  __syncthreads();
  const int nwin = L / interval;
  if (interval >= 128) {
    // one (partial) window per block: reduce 128 columns, accumulate with integer atomics when the window spans blocks
    if (threadIdx.x == 0) {
      int c = 0, o = 0;
      for (int i = 0; i < 128; ++i) {
        c += s_cnt[i];
        o += s_on[i];
      }
      const int w = col / interval;
      int* acc = reinterpret_cast<int*>(out);  // integer accumulators, converted by the finalize kernel
      atomicAdd(acc + ((long long)cand * 2 * nwin + w), c);
      atomicAdd(acc + ((long long)cand * 2 * nwin + nwin + w), o);
    }
  } else {
    const int wins = 128 / interval;
    if (threadIdx.x < wins) {
      int c = 0, o = 0;
      for (int i = 0; i < interval; ++i) {
        c += s_cnt[threadIdx.x * interval + i];
        o += s_on[threadIdx.x * interval + i];
      }
      const int w = blockIdx.x * wins + threadIdx.x;
      int* acc = reinterpret_cast<int*>(out);
      acc[(long long)cand * 2 * nwin + w] = c;
      acc[(long long)cand * 2 * nwin + nwin + w] = o;
    }
  }
}

// integer window sums -> the reference's fp32 values: vertical = count / interval (mean), horizontal = onsets / hscale
__global__ void note_density_finalize_kernel(float* __restrict__ out, long long total, int nwin, int interval,
                                             float hscale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = reinterpret_cast<int*>(out)[i];
  const bool vertical = (i % (2 * nwin)) < nwin;
  out[i] = vertical ? (float)v / (float)interval : (float)v / hscale;
}

cudaError_t launch_rule_note_density(float* roll, float* out, int n, int ch, int L, int interval, float hscale,
                                     cudaStream_t s) {
  if (L % 128 != 0 || n <= 0 || interval <= 0 || L % interval != 0) return cudaErrorInvalidValue;
  if (interval < 128 && (128 % interval) != 0) return cudaErrorInvalidValue;
  if (interval >= 128 && (interval % 128) != 0) return cudaErrorInvalidValue;
  const int nwin = L / interval;
  const long long total = (long long)n * 2 * nwin;
  ProfScope prof("rule_note_density", 0, 0, (double)n * 128.0 * L * 8.0, s);
  cudaError_t e = cudaMemsetAsync(out, 0, total * sizeof(float), s);
  if (e != cudaSuccess) return e;
  note_density_kernel<<<dim3(L / 128, n), 128, 0, s>>>(roll, out, ch, L, interval, hscale);
  g_rule_launches.fetch_add(1, std::memory_order_relaxed);
  note_density_finalize_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(out, total, nwin, interval, hscale);
  return done();
}

// total[i] += weight * -(loss_i); one warp per candidate, sequential-in-K partial sums folded by shuffles
__global__ void rule_loss_accum_kernel(const float* __restrict__ gen, const float* __restrict__ target,
                                       float* __restrict__ total, int n, int B, int K, int kind, float weight) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* g = gen + (long long)i * K;
  const float* t = target + (long long)(i % B) * K;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) {
    if (kind == 0) {
      const float d = g[k] - t[k];
      s += d * d;
    } else {
      s += (g[k] != t[k]) ? 1.f : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const float loss = s / (float)K;
    total[i] = total[i] + (-loss) * weight;
  }
}

cudaError_t launch_rule_loss_accum(const float* gen, const float* target, float* total, int n, int B, int K, int kind,
                                   float weight, cudaStream_t s) {
  rule_loss_accum_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(gen, target, total, n, B, K, kind, weight);
  return done();
}

// One block per sample: warp 0 finds the first maximal candidate (ties -> lowest index, like torch.argmax), then the
// block copies the winner.  total [N, B] (candidate-major, like total_log_prob.view(N, -1)), cand [N, B, elems].
__global__ void __launch_bounds__(256) scg_select_kernel(const float* __restrict__ total,
                                                         const float* __restrict__ cand, float* __restrict__ out,
                                                         long long* __restrict__ idx, int N, int B, long long elems) {
  __shared__ int s_best;
  const int b = blockIdx.x;
  if (threadIdx.x < 32) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    bool any_nan = false;
    for (int n = threadIdx.x; n < N; n += 32) {
      const float v = total[(long long)n * B + b];
      // torch.argmax treats NaN as the maximum (first NaN wins)
      if (v != v) {
        if (!any_nan) {
          any_nan = true;
          bi = n;
        }
      } else if (!any_nan && (v > best || (v == best && n < bi))) {
        best = v;
        bi = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int on = __shfl_xor_sync(0xffffffffu, (int)any_nan, o);
      if (on && any_nan) {
        bi = oi < bi ? oi : bi;
      } else if (on) {
        any_nan = true;
        bi = oi;
      } else if (!any_nan && (ob > best || (ob == best && oi < bi))) {
        best = ob;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) {
      s_best = bi;
      idx[b] = bi;
    }
  }
  __syncthreads();
  const float* srcf = cand + ((long long)s_best * B + b) * elems;
  float* dstf = out + (long long)b * elems;
  if ((elems & 3) == 0) {
    const float4* src = reinterpret_cast<const float4*>(srcf);
    float4* dst = reinterpret_cast<float4*>(dstf);
    for (long long i = threadIdx.x; i < elems / 4; i += blockDim.x) dst[i] = src[i];
  } else {
    for (long long i = threadIdx.x; i < elems; i += blockDim.x) dstf[i] = srcf[i];
  }
}

cudaError_t launch_scg_select(const float* total, const float* cand, float* out, long long* idx, int N, int B,
                              long long elems, cudaStream_t s) {
  if (N <= 0 || B <= 0) return cudaErrorInvalidValue;
  if ((elems % 4) == 0 && ((reinterpret_cast<uintptr_t>(cand) | reinterpret_cast<uintptr_t>(out)) & 15))
    return cudaErrorInvalidValue;
  scg_select_kernel<<<B, 256, 0, s>>>(total, cand, out, idx, N, B, elems);
  return done();
}

}  // namespace rgm
