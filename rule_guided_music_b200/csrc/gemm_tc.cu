// Host launcher for the tcgen05 implicit-GEMM kernel: builds the TMA tensor maps and picks the tile shape.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "api_util.h"
#include "conv_gn.cuh"
#include "gemm_host.h"

namespace rgm {

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  unsigned long long d[4];
  unsigned long long s[3];
  unsigned box[4];
  unsigned estr[4];
  int rank;
  bool operator==(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
};

std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// fp16 tensor map, 128-byte swizzle, zero fill out of bounds. dims/strides innermost first; strides in bytes
// for dims 1..rank-1.
bool make_map(CUtensorMap* out, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box, std::string* err, const cuuint32_t* elem_strides = nullptr) {
  MapKey key;
  std::memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) {
    key.d[i] = dims[i];
    key.box[i] = box[i];
    key.estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i) key.s[i - 1] = strides[i - 1];
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return true;
    }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    if (err) *err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
               (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
               box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
      *err = buf;
    }
    return false;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return true;
}

std::atomic<unsigned long long> g_launches{0};

template <int BN, int MT, int EPI>
cudaError_t launch_inst(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int grid,
                        cudaStream_t stream) {
  static SmemAttr attr;
  auto kern = gemm_tc_kernel<BN, MT, EPI>;
  if (cudaError_t e = attr.ensure(kern, GemmCfg<BN, MT>::SMEM_BYTES); e != cudaSuccess) return e;
  kern<<<grid, GEMM_THREADS, GemmCfg<BN, MT>::SMEM_BYTES, stream>>>(ma, mb, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <int EPI>
cudaError_t launch_sw(const CUtensorMap& mx, const CUtensorMap& mw, const GemmParams& p, int grid,
                      cudaStream_t stream) {
  static SmemAttr attr;
  auto kern = gemm_sw_kernel<EPI>;
  if (cudaError_t e = attr.ensure(kern, SW_SMEM_BYTES); e != cudaSuccess) return e;
  kern<<<grid, GEMM_THREADS, SW_SMEM_BYTES, stream>>>(mx, mw, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <int EPI>
cudaError_t launch_sw2(const CUtensorMap& mx, const CUtensorMap& mw, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  static SmemAttr attr;
  auto kern = gemm_sw2_kernel<EPI>;
  if (cudaError_t e = attr.ensure(kern, SW2_SMEM_BYTES); e != cudaSuccess) return e;
  kern<<<grid, GEMM_THREADS, SW2_SMEM_BYTES, stream>>>(mx, mw, p);  // 2-CTA clusters (__cluster_dims__)
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

bool pair_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RGM_GEMM_PAIR");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// CTAs (single-CTA kernel) / CTA pairs (pair kernel) of the EPI_F16 feature-major kernels that are co-resident on the
// current device: the GroupNorm-in-epilogue convolutions wait on each other's tiles, so their grid must not exceed it
// (a pair needs two free SMs of one TPC; a GPC with an odd SM count leaves one SM without a partner).
int resident_units(bool pair) {
  static std::atomic<int> cache[64][2];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 63;
  int v = cache[dev][pair ? 1 : 0].load(std::memory_order_relaxed);
  if (v > 0) return v;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (pair) {
    auto kern = gemm_sw2_kernel<EPI_F16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW2_SMEM_BYTES);
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (sms / 2), 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SW2_SMEM_BYTES;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = sms / 2 - 8;  // conservative: one pair lost per GPC
    }
    v = n < sms / 2 ? n : sms / 2;
  } else {
    v = sms;  // the single-CTA kernel launches at most one CTA per SM, and one always fits (198 KB of shared memory)
  }
  if (v < 1) v = 1;
  if (dev != 63) cache[dev][pair ? 1 : 0].store(v, std::memory_order_relaxed);
  return v;
}

}  // namespace

int gemm_resident_units(bool pair) { return resident_units(pair); }

unsigned long long gemm_launch_count() { return g_launches.load(); }

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

cudaError_t launch_gemm(const GemmDesc& d, cudaStream_t stream, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = m;
    return cudaErrorInvalidValue;
  };
  if (d.C % GEMM_BLOCK_K != 0) return fail("gemm: C must be a multiple of 64");
  if (d.lda % 8 != 0 || d.lda < d.C) return fail("gemm: lda must be >= C and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(d.A) & 15) || (reinterpret_cast<uintptr_t>(d.B) & 15))
    return fail("gemm: operands must be 16-byte aligned");

  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  const bool col_epi = d.epi == EPI_F16 || d.epi == EPI_F32 || d.epi == EPI_GATE_RESID || d.epi == EPI_QKV_ROPE;
  // feature-major kernel (128 features x 256 rows per tile) for every feature count that is a multiple of 128;
  // the row-major kernel with 32-column tiles serves the tiny outputs (final layer, conv_out) and odd widths
  const bool sw = col_epi && d.N % SW_FEATS == 0 && d.block_n != 32;
  const int rows_per_tile = sw ? SW_ROWS : GEMM_BLOCK_M;
  // output grid: the input grid, except for the stride-2 Downsample conv
  const bool down = d.conv == CONV_DOWN2;
  if (down && (d.H % 2 != 0 || d.W % 2 != 0 || d.H < 2)) return fail("gemm: stride-2 conv needs even image dims");
  const int oH = down ? d.H / 2 : d.H, oW = down ? d.W / 2 : d.W;
  // a linear layer (H == 1) uses boxes of rows_per_tile rows; rows past the end of A are zero-filled by TMA and
  // masked in the epilogue.  Images use full-width boxes of bh rows.
  const int bw = (oH == 1 || oW >= rows_per_tile) ? rows_per_tile : oW;
  if (rows_per_tile % bw != 0) return fail("gemm: image width must divide the row tile");
  const int bh = rows_per_tile / bw;
  if (bh > 1 && oH % bh != 0) return fail("gemm: H not a multiple of the tile height");
  if (oH > 1 && oW > rows_per_tile) return fail("gemm: images wider than the row tile are not supported");
  if (d.n_img > 1 && oH == 1 && oW % rows_per_tile != 0)
    return fail("gemm: batched rows must be a multiple of the row tile");
  if (down && (2 * bw > 256 || 2 * bh > 256)) return fail("gemm: stride-2 box exceeds the TMA box limit");
  p.bw = bw;
  p.bh = bh;
  p.in_stride = down ? 2 : 1;
  p.tiles_per_row = (oW + bw - 1) / bw;
  p.tiles_per_img = p.tiles_per_row * (oH / bh > 0 ? oH / bh : 1);
  p.M = d.n_img * oH * oW;
  p.N = d.N;
  p.num_m_tiles = d.n_img * p.tiles_per_img;
  p.slots_per_par = (p.M + 127) / 128;
  p.kb_per_tap = d.C / GEMM_BLOCK_K;
  p.b_batched = d.b_batch > 1 ? 1 : 0;
  p.num_par = 1;
  if (d.conv == CONV_1x1) {
    p.num_taps = 1;
  } else if (d.conv == CONV_3x3) {
    p.num_taps = 9;
    for (int t = 0; t < 9; ++t) {
      p.tap_dy[0][t] = (signed char)(t / 3 - 1);
      p.tap_dx[0][t] = (signed char)(t % 3 - 1);
    }
  } else if (d.conv == CONV_DOWN2) {
    // F.pad(x, (0,1,0,1)) + conv(stride 2, padding 0): output (h, w) reads input (2h + ky, 2w + kx), ky, kx in 0..2;
    // the padded last row / column is the TMA out-of-bounds zero fill
    p.num_taps = 9;
    for (int t = 0; t < 9; ++t) {
      p.tap_dy[0][t] = (signed char)(t / 3);
      p.tap_dx[0][t] = (signed char)(t % 3);
    }
  } else if (d.conv == CONV_UP2) {
    p.num_taps = 4;
    p.num_par = 4;
    for (int par = 0; par < 4; ++par)
      for (int t = 0; t < 4; ++t) {
        p.tap_dy[par][t] = (signed char)((t >> 1) + (par >> 1) - 1);
        p.tap_dx[par][t] = (signed char)((t & 1) + (par & 1) - 1);
      }
  } else {
    return fail("gemm: unknown conv kind");
  }
  p.epi = d.e;
  p.trace = d.trace;
  p.par_fast = 1;
  if (const char* pf = getenv("RGM_PAR_FAST")) p.par_fast = atoi(pf);  // A/B knob: 0 = parity-major tile order
  if (const char* bn = getenv("RGM_GEMM_BAND")) p.band_n = atoi(bn);  // experiment knobs, read per launch
  if (const char* dbg = getenv("RGM_GEMM_DEBUG")) p.debug = atoi(dbg);
  if (const char* tp = getenv("RGM_DEBUG_TRACE_PTR")) {  // development aid: trace every launch with a given epilogue
    const char* te = getenv("RGM_DEBUG_TRACE_EPI");
    if (te && atoi(te) == d.epi && (!getenv("RGM_DEBUG_TRACE_N") || atoi(getenv("RGM_DEBUG_TRACE_N")) == d.N))
      p.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0));
  }

  const int bn = sw ? SW_FEATS : 32;
  if (d.N % bn != 0) return fail("gemm: N must be a multiple of 32");
  if (!sw && d.epi != EPI_F16 && d.epi != EPI_F32 && d.epi != EPI_UNPATCH && d.epi != EPI_ROLL)
    return fail("gemm: this epilogue needs N to be a multiple of 128");
  if (d.rows_b < p.num_par * d.N) return fail("gemm: B has fewer rows than num_par * N");
  if (!sw && d.e.gn_part != nullptr) return fail("gemm: GroupNorm partials need a feature count that is a multiple of 128");
  p.num_n_tiles = d.N / bn;
  if (d.epi == EPI_QKV_ROPE && (d.e.T % 32 != 0 || (d.e.heads * d.e.dh) % 32 != 0))
    return fail("gemm: the QKV epilogue needs T and hidden to be multiples of 32");
  if (d.epi == EPI_GATE_RESID && d.e.rows_per_sample % 32 != 0)
    return fail("gemm: the gate/residual epilogue needs rows_per_sample to be a multiple of 32");
  if ((long long)p.M * (d.conv == CONV_UP2 ? 4 : 1) >= (1LL << 31)) return fail("gemm: more than 2^31 output rows");

  CUtensorMap ma, mb;
  {
    cuuint64_t wdim = (cuuint64_t)d.W;
    if (d.H == 1 && d.n_img == 1 && d.a_rows > d.W) wdim = (cuuint64_t)d.a_rows;
    cuuint64_t dims[4] = {(cuuint64_t)d.C, wdim, (cuuint64_t)d.H, (cuuint64_t)d.n_img};
    cuuint64_t strides[3] = {(cuuint64_t)d.lda * 2, wdim * d.lda * 2, (cuuint64_t)d.H * wdim * d.lda * 2};
    // with element strides s, TMA loads ceil(box / s) elements per dimension: box = s * (elements wanted)
    const cuuint32_t es = down ? 2 : 1;
    cuuint32_t box[4] = {GEMM_BLOCK_K, (cuuint32_t)bw * es, (cuuint32_t)bh * es, 1};
    cuuint32_t estr[4] = {1, es, es, 1};
    if (!make_map(&ma, d.A, 4, dims, strides, box, err, estr)) return cudaErrorInvalidValue;
  }
  {
    const long long K = (long long)p.num_taps * d.C;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)d.rows_b, (cuuint64_t)d.b_batch};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * d.rows_b * 2};
    cuuint32_t box[3] = {GEMM_BLOCK_K, (cuuint32_t)bn, 1};
    if (!make_map(&mb, d.B, 3, dims, strides, box, err)) return cudaErrorInvalidValue;
  }

  // CTA pairs (gemm_sw2_kernel) when there are at least two feature tiles and enough pair tiles to fill the machine
  const long long pair_tiles = (long long)p.num_m_tiles * ((p.num_n_tiles + 1) / 2) * p.num_par;
  const bool pair =
      sw && !down && pair_mode_enabled() && p.num_n_tiles >= 2 && pair_tiles >= device_sm_count() / 2;
  CUtensorMap ma2;
  if (pair) {
    // this CTA's half of the 256-row activation box: half the image rows, or 128 of the 256 linear rows
    cuuint64_t wdim = (cuuint64_t)d.W;
    if (d.H == 1 && d.n_img == 1 && d.a_rows > d.W) wdim = (cuuint64_t)d.a_rows;
    cuuint64_t dims[4] = {(cuuint64_t)d.C, wdim, (cuuint64_t)d.H, (cuuint64_t)d.n_img};
    cuuint64_t strides[3] = {(cuuint64_t)d.lda * 2, wdim * d.lda * 2, (cuuint64_t)d.H * wdim * d.lda * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, (cuuint32_t)(bh > 1 ? bw : bw / 2), (cuuint32_t)(bh > 1 ? bh / 2 : 1), 1};
    if (!make_map(&ma2, d.A, 4, dims, strides, box, err)) return cudaErrorInvalidValue;
  }
  const long long total = pair ? pair_tiles : (long long)p.num_m_tiles * p.num_n_tiles * p.num_par;
  int grid = (int)(total < device_sm_count() ? total : device_sm_count());
  if (pair) {
    const int max_pairs = device_sm_count() / 2;
    grid = 2 * (int)(total < max_pairs ? total : max_pairs);
  }
  if (grid <= 0) return cudaSuccess;
  const bool gn_fuse = d.e.gn_sums != nullptr;
  if (gn_fuse) {
    // GroupNorm inside the epilogue (gn_epilogue_loop): whole 256-row tiles of one image, statically assigned tiles,
    // every CTA of the grid resident, an image's tiles within one grid-stride of each other
    const bool up_ok = d.e.up2 && d.e.gn_out2 != nullptr && d.e.upW % 32 == 0 && p.par_fast && d.e.resid == nullptr;
    if (!sw || d.epi != EPI_F16 || (p.num_par != 1 && !up_ok) || down ||
        (d.e.resid != nullptr && d.e.gn_out2 == nullptr) || (d.e.resid != nullptr && d.e.ldr != d.e.ldo) ||
        d.e.addtab != nullptr || (d.e.up2 && !up_ok) ||
        d.e.act != ACT_NONE || d.e.gn_sums == nullptr || d.e.gn_gamma == nullptr || d.e.gn_beta == nullptr ||
        d.e.gn_err == nullptr || d.b_batch > 1)
      return fail("gemm: this layer cannot normalise its output in the epilogue");
    if ((oH * oW) % SW_ROWS != 0 || p.M % SW_ROWS != 0 || (d.N != 128 && d.N != 256 && d.N != 512))
      return fail("gemm: GroupNorm in the epilogue needs images of a multiple of 256 pixels and 128 / 256 / 512 features");
    const int units = resident_units(pair);
    if (pair) grid = 2 * (int)(total < units ? total : units);
    else grid = (int)(total < units ? total : units);
    const int span = p.num_par * p.tiles_per_img * (pair ? (p.num_n_tiles + 1) / 2 : p.num_n_tiles);
    if (span > (pair ? grid / 2 : grid)) return fail("gemm: an image spans more tiles than there are resident CTAs");
    if (p.tiles_per_img * 2 * p.num_par > 255) return fail("gemm: more than 255 statistics contributions per image");
    // RGM_GN_ALIGN=1 rounds the grid down to whole images per wave (no image straddles two waves of the grid).  Measured
    // (profiles/r2_trace_conv_norm.txt): not worth the idle CTAs -- 0.599 vs 0.597 ms at 128 -> 128 @ 128x128 (128 of
    // 148 CTAs), 0.499 vs 0.474 ms at 256 -> 256 @ 64x64 (64 of 74 pairs) -- so it is off.
    static const int align = [] {
      const char* e = getenv("RGM_GN_ALIGN");
      return e ? atoi(e) : 0;
    }();
    if (align) {
      int g = pair ? grid / 2 : grid;
      g = (g / span) * span;
      grid = pair ? 2 * g : g;
    }
    p.band_n = 0;
    p.epi.gn_inv_count = 1.0f / ((float)(p.num_par * oH * oW) * (float)(d.N / 32));
  }

  // profiling label: conv kind, K, N and epilogue identify the layer family; flops_alg counts the reference's
  // arithmetic (a 3x3 conv on the upsampled image for CONV_UP2), flops_exec what this kernel executes
  char pname[96];
  const double Kexec = (double)p.num_taps * d.C;
  const double rows = (double)p.M * p.num_par;
  const double f_exec = 2.0 * rows * d.N * Kexec;
  const double f_alg = d.conv == CONV_UP2 ? 2.0 * rows * d.N * 9.0 * d.C : f_exec;
  if (g_prof_on.load(std::memory_order_relaxed))
    snprintf(pname, sizeof pname, "gemm_tc conv%d H%d K%d N%d epi%d %s%s", d.conv, d.H, (int)Kexec, d.N, d.epi,
             pair ? "f256xr256 pair" : (sw ? "f128xr256" : "r128xf32"), gn_fuse ? (d.e.gn_out2 != nullptr ? " +raw+norm" : " +norm") : "");
  ProfScope prof(pname, f_alg, f_exec, 0.0, stream);

  cudaError_t st = cudaErrorInvalidValue;
  if (pair) {
    switch (d.epi) {
      case EPI_F16: st = launch_sw2<EPI_F16>(ma2, mb, p, grid, stream); break;
      case EPI_F32: st = launch_sw2<EPI_F32>(ma2, mb, p, grid, stream); break;
      case EPI_GATE_RESID: st = launch_sw2<EPI_GATE_RESID>(ma2, mb, p, grid, stream); break;
      default: st = launch_sw2<EPI_QKV_ROPE>(ma2, mb, p, grid, stream); break;
    }
  } else if (sw) {
    switch (d.epi) {
      case EPI_F16: st = launch_sw<EPI_F16>(ma, mb, p, grid, stream); break;
      case EPI_F32: st = launch_sw<EPI_F32>(ma, mb, p, grid, stream); break;
      case EPI_GATE_RESID: st = launch_sw<EPI_GATE_RESID>(ma, mb, p, grid, stream); break;
      default: st = launch_sw<EPI_QKV_ROPE>(ma, mb, p, grid, stream); break;
    }
  } else {
    switch (d.epi) {
      case EPI_F16: st = launch_inst<32, 1, EPI_F16>(ma, mb, p, grid, stream); break;
      case EPI_F32: st = launch_inst<32, 1, EPI_F32>(ma, mb, p, grid, stream); break;
      case EPI_UNPATCH: st = launch_inst<32, 1, EPI_UNPATCH>(ma, mb, p, grid, stream); break;
      default: st = launch_inst<32, 1, EPI_ROLL>(ma, mb, p, grid, stream); break;
    }
  }
  if (st != cudaSuccess && err) *err = std::string("gemm launch: ") + cudaGetErrorString(st);
  return st;
}

// Layers whose epilogue can normalise their own output (gn_epilogue_loop); the same geometry as launch_gemm.
bool gemm_gn_fuse_supported(const GemmDesc& d) {
  const bool up = d.conv == CONV_UP2;  // (dual form only: the caller passes EpiParams::gn_out2)
  if (d.conv != CONV_3x3 && d.conv != CONV_1x1 && !up) return false;
  if (up) {
    const char* pf = getenv("RGM_PAR_FAST");
    if (d.W % 32 != 0 || (pf && atoi(pf) == 0)) return false;
  }
  // 32 groups of 4, 8 or 16 channels: whole channel quads per group, groups inside one warp's 32 features
  if (d.epi != EPI_F16 || (d.N != 128 && d.N != 256 && d.N != 512) || d.b_batch > 1 || d.block_n == 32) return false;
  if (d.e.addtab != nullptr || d.e.act != ACT_NONE) return false;  // (a residual needs the dual form)
  if (d.H < 2 || d.W > SW_ROWS || SW_ROWS % d.W != 0) return false;
  const long long HW = (long long)d.H * d.W;
  if (HW % SW_ROWS != 0) return false;
  const int tiles_per_img = (int)(HW / SW_ROWS), n_tiles = d.N / SW_FEATS;
  const int npar = up ? 4 : 1;
  const long long num_m = (long long)d.n_img * tiles_per_img;
  const long long pair_tiles = num_m * ((n_tiles + 1) / 2) * npar;
  const bool pair = pair_mode_enabled() && n_tiles >= 2 && pair_tiles >= device_sm_count() / 2;
  const long long total = pair ? pair_tiles : num_m * n_tiles * npar;
  const int units = resident_units(pair);
  const long long g = total < units ? total : units;
  const int span = npar * tiles_per_img * (pair ? (n_tiles + 1) / 2 : n_tiles);
  return span <= g && tiles_per_img * 2 * npar <= 255;
}

// shape test: layers the fused kernel can run
bool conv_gn_shape_ok(const GemmDesc& d) {
  return d.conv == CONV_3x3 && d.epi == EPI_F16 && d.W == CG_W && d.H >= 2 && d.H % CG_ROWS == 0 && d.N == SW_FEATS &&
         d.C % GEMM_BLOCK_K == 0 && d.lda == d.C && d.b_batch <= 1 && d.e.up2 == 0 && d.e.addtab == nullptr;
}

// policy: whether the VAE uses it.  OFF by default: measured on B200 (profiles/README.md, round 2) one launch of the fused
// kernel takes 0.93 ms at 128 -> 128 @ 128x128 against 0.57 + 0.22 ms for convolution + GroupNorm pass -- an N = 128
// tcgen05.mma with both operands in shared memory is shared-memory-bound (107 instead of 64 cycles per 128x128x16) and
// the in-place transform (11 K cycles per k-block on four warps) does not hide under it.  RGM_CONV_GN=1 switches it on.
bool conv_gn_supported(const GemmDesc& d) {
  static const int on = [] {
    const char* e = getenv("RGM_CONV_GN");
    return (e && atoi(e) == 1) ? 1 : 0;
  }();
  return on && conv_gn_shape_ok(d);
}

cudaError_t launch_conv_gn(const GemmDesc& d, const float2* in_ab, cudaStream_t stream, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = m;
    return cudaErrorInvalidValue;
  };
  if (!conv_gn_shape_ok(d) || in_ab == nullptr) return fail("conv_gn: unsupported layer shape");
  if ((reinterpret_cast<uintptr_t>(d.A) & 15) || (reinterpret_cast<uintptr_t>(d.B) & 15))
    return fail("conv_gn: operands must be 16-byte aligned");
  if (d.rows_b < d.N) return fail("conv_gn: weight rows");
  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  p.M = d.n_img * d.H * d.W;
  p.N = d.N;
  p.num_m_tiles = d.n_img * (d.H / CG_ROWS);
  p.num_n_tiles = 1;
  p.num_par = 1;
  p.num_taps = 9;
  p.kb_per_tap = d.C / GEMM_BLOCK_K;
  p.slots_per_par = (p.M + 127) / 128;
  p.in_stride = 1;
  p.epi = d.e;
  p.trace = d.trace;
  if (const char* dbg = getenv("RGM_GEMM_DEBUG")) p.debug = atoi(dbg);
  ConvGnParams cg;
  cg.in_ab = in_ab;
  cg.H = d.H;
  cg.kb = d.C / GEMM_BLOCK_K;
  CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.C, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.n_img};
    cuuint64_t strides[3] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.W * d.lda * 2, (cuuint64_t)d.H * d.W * d.lda * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, CG_HALO_W, CG_HALO_ROWS, 1};
    if (!make_map(&mx, d.A, 4, dims, strides, box, err)) return cudaErrorInvalidValue;
  }
  {
    const long long K = 9LL * d.C;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)d.rows_b, 1};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * d.rows_b * 2};
    cuuint32_t box[3] = {GEMM_BLOCK_K, SW_FEATS, 1};
    if (!make_map(&mw, d.B, 3, dims, strides, box, err)) return cudaErrorInvalidValue;
  }
  int grid = p.num_m_tiles < device_sm_count() ? p.num_m_tiles : device_sm_count();
  if (grid <= 0) return cudaSuccess;
  char pname[96];
  const double fl = 2.0 * (double)p.M * d.N * 9.0 * d.C;
  if (g_prof_on.load(std::memory_order_relaxed))
    snprintf(pname, sizeof pname, "conv_gn conv1 H%d K%d N%d epi0 f128xr2x128 norm+swish fused", d.H, 9 * d.C, d.N);
  ProfScope prof(pname, fl, fl, 0.0, stream);
  static SmemAttr attr;
  if (cudaError_t e = attr.ensure(conv_gn_kernel, CG_SMEM_BYTES); e != cudaSuccess) return e;
  conv_gn_kernel<<<grid, CG_THREADS, CG_SMEM_BYTES, stream>>>(mx, mw, p, cg);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t st = cudaGetLastError();
  if (st != cudaSuccess && err) *err = std::string("conv_gn launch: ") + cudaGetErrorString(st);
  return st;
}

}  // namespace rgm
