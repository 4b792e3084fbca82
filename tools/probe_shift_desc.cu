// Development probe (GPU): does a tcgen05 shared-memory descriptor with the 128-byte swizzle accept a start address
// that is shifted by whole 128-byte ROWS inside a densely stored pixel-linear tile?
//
// The fused GroupNorm+conv kernel wants to land ONE halo tile of an image in shared memory ([pixels][64 channels],
// TMA SWIZZLE_128B) and feed the 9 taps of a 3x3 convolution as nine shifted views of it: tap (dy, dx) of output pixels
// p .. p+127 is the operand whose row r is halo pixel p + r + dy*pitch + dx, i.e. the same buffer with the start address
// advanced by (dy*pitch + dx) * 128 bytes.  CUTLASS only ever builds descriptors on 1024-byte-aligned atoms (plus
// K-advances inside a row), so this probe measures what the hardware does for row shifts, with the descriptor's
// base-offset field (bits 49-51) either 0 or (start >> 7) & 7, for the shifted tile used as the B operand (N side,
// feature-major kernels) and as the A operand (M side).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rule_guided_music_b200/csrc tools/probe_shift_desc.cu
//        -o tools/bin/probe_shift_desc -lcuda      (built by tools/build_probes.sh; run on the GPU box)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptx.cuh"

using namespace rgm;

constexpr int PIX = 384;     // pixels in the tile (3 TMA boxes of 128 rows)
constexpr int FEAT = 128;    // weight rows
constexpr int KC = 64;       // channels = one 128-byte swizzle row
constexpr int NSHIFT = 16;
__constant__ int c_shifts[NSHIFT] = {0, 1, 2, 3, 5, 7, 8, 9, 13, 16, 64, 129, 130, 131, 255, 256};

__device__ __forceinline__ uint64_t desc_sw128_shift(uint32_t addr, int base_off_mode) {
  uint64_t d = static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  if (base_off_mode) d |= static_cast<uint64_t>((addr >> 7) & 7u) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// out[mode][side][shift][128][128] fp32: mode = base-offset field off/on, side 0 = shifted tile is B (D = W . Xs^T,
// rows = features), side 1 = shifted tile is A (D = Xs . W^T, rows = pixels)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                    // [PIX][128 B]
  uint8_t* sW = smem + PIX * 128;        // [FEAT][128 B]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sW + FEAT * 128);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, (PIX + FEAT) * 128);
    for (int i = 0; i < PIX / 128; ++i) tma_load_2d(sX + i * 128 * 128, &map_x, bar_load, 0, i * 128);
    tma_load_2d(sW, &map_w, bar_load, 0, 0);
  }
  mbar_wait(bar_load, 0);
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_f16(128, 128);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int side = 0; side < 2; ++side)
      for (int si = 0; si < NSHIFT; ++si) {
        if (threadIdx.x == 0) {
          const uint32_t xs = smem_u32(sX) + c_shifts[si] * 128;
          const uint32_t ws = smem_u32(sW);
          for (int k = 0; k < KC / 16; ++k) {
            const uint64_t dx = desc_sw128_shift(xs, mode) + 2 * k;
            const uint64_t dw = desc_sw128_shift(ws, 0) + 2 * k;
            umma_f16(tmem, side == 0 ? dw : dx, side == 0 ? dx : dw, idesc, k != 0);
          }
          umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        float* o = out + ((((size_t)mode * 2 + side) * NSHIFT + si) * 128 + warp * 32 + lane) * 128;
        for (int c = 0; c < 128; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
          tmem_ld_wait();
          for (int j = 0; j < 32; ++j) o[c + j] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
      }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool map2d(PFN_encodeTiled enc, CUtensorMap* m, void* ptr, unsigned long long rows) {
  cuuint64_t dims[2] = {KC, rows};
  cuuint64_t strides[1] = {KC * 2};
  cuuint32_t box[2] = {KC, 128};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fp);
  std::vector<__half> hx((size_t)PIX * KC), hw((size_t)FEAT * KC);
  std::vector<float> fx(hx.size()), fw(hw.size());
  unsigned s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((s >> 9) & 0xFFFF) / 65536.0f - 0.5f;
  };
  for (size_t i = 0; i < hx.size(); ++i) {
    hx[i] = __float2half_rn(rnd());
    fx[i] = __half2float(hx[i]);
  }
  for (size_t i = 0; i < hw.size(); ++i) {
    hw[i] = __float2half_rn(rnd());
    fw[i] = __half2float(hw[i]);
  }
  __half *dx, *dw;
  float* dout;
  const size_t nout = (size_t)2 * 2 * NSHIFT * 128 * 128;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dout, nout * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, nout * 4);
  CUtensorMap mx, mw;
  if (!map2d(enc, &mx, dx, PIX) || !map2d(enc, &mw, dw, FEAT)) {
    printf("tensor map failed\n");
    return 1;
  }
  const size_t smem = 1024 + (PIX + FEAT) * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(mx, mw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> out(nout);
  cudaMemcpy(out.data(), dout, nout * 4, cudaMemcpyDeviceToHost);
  const int shifts[NSHIFT] = {0, 1, 2, 3, 5, 7, 8, 9, 13, 16, 64, 129, 130, 131, 255, 256};
  printf("max |D - ref| for shifted SW128 descriptors (rows of 128 B), 128x128x64 fp16 MMA; ~1e-6 = correct\n");
  printf("%-28s", "shift (rows):");
  for (int si = 0; si < NSHIFT; ++si) printf("%9d", shifts[si]);
  printf("\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int side = 0; side < 2; ++side) {
      printf("%-28s", mode == 0 ? (side == 0 ? "base_offset=0, shifted B" : "base_offset=0, shifted A")
                                : (side == 0 ? "base_offset=row&7, shifted B" : "base_offset=row&7, shifted A"));
      for (int si = 0; si < NSHIFT; ++si) {
        double worst = 0.0;
        const float* o = out.data() + (((size_t)mode * 2 + side) * NSHIFT + si) * 128 * 128;
        for (int r = 0; r < 128; ++r)
          for (int c = 0; c < 128; ++c) {
            // side 0: D[feature r][pixel c]; side 1: D[pixel r][feature c]
            const int f = side == 0 ? r : c, p = (side == 0 ? c : r) + shifts[si];
            double ref = 0.0;
            if (p < PIX)
              for (int k = 0; k < KC; ++k) ref += (double)fw[(size_t)f * KC + k] * fx[(size_t)p * KC + k];
            else
              continue;
            const double d = fabs(o[r * 128 + c] - ref);
            if (d > worst) worst = d;
          }
        printf("%9.1e", worst);
      }
      printf("\n");
    }
  return 0;
}
