// Host-side description of one implicit-GEMM launch (see gemm_tc.cuh).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>

#include "gemm_tc.cuh"

namespace rgm {

enum ConvKind : int {
  CONV_1x1 = 0,   // also: plain linear layer
  CONV_3x3 = 1,   // 9 taps, zero padding 1
  CONV_UP2 = 2,   // nearest-2x upsample followed by 3x3 conv, run as 4 output-parity 2x2 sub-convs
  CONV_DOWN2 = 3, // Downsample (model.py:55-75): pad right/bottom by one, 3x3 conv with stride 2; H, W are the INPUT
                  // dims (even), the output is H/2 x W/2; loaded by TMA boxes with element strides (1, 2, 2, 1)
};

struct GemmDesc {
  // A operand: fp16 [n_img, H, W, C] with pixel stride lda (elements, >= C, multiple of 8)
  const __half* A = nullptr;
  int n_img = 1, H = 1, W = 1, C = 0, lda = 0;
  long long a_rows = 0;  // linear layers (H == 1, n_img == 1): rows actually allocated behind A (>= W); 0 = W
  // B operand: fp16 [b_batch][rows_b][K]; K = taps * C; rows_b = num_par * N (N padded to the N tile)
  const __half* B = nullptr;
  int rows_b = 0, b_batch = 1;
  int N = 0;          // output columns (multiple of the N tile)
  int conv = CONV_1x1;
  int epi = EPI_F16;
  int block_n = 0;    // 0 = choose
  unsigned long long* trace = nullptr;  // development aid, see GemmParams::trace
  EpiParams e{};
};

// Launch on `stream`. Returns cudaSuccess or the failing status; message in *err if given.
cudaError_t launch_gemm(const GemmDesc& d, cudaStream_t stream, std::string* err = nullptr);

// 3x3 convolution (CONV_3x3, EPI_F16) whose INPUT is normalised + activated on the way into the operand stage
// (conv_gn.cuh): d.A is the RAW tensor [n_img, H, 128, C], in_ab its GroupNorm affine [n_img][C].  Needs W == 128,
// H even, N == 128 output features.  conv_gn_shape_ok tells whether a layer qualifies, conv_gn_supported whether the VAE
// should use it (policy: off unless RGM_CONV_GN=1, see gemm_tc.cu).
bool conv_gn_shape_ok(const GemmDesc& d);
bool conv_gn_supported(const GemmDesc& d);
cudaError_t launch_conv_gn(const GemmDesc& d, const float2* in_ab, cudaStream_t stream, std::string* err = nullptr);

// Whether launch_gemm can apply GroupNorm (+ swish) to this layer's own output inside the epilogue (EpiParams::gn_sums
// etc., gemm_tc.cuh gn_epilogue_loop): 3x3 / 1x1 convolution, 128 / 256 / 512 features, images of a multiple of 256
// pixels spanning no more tiles than there are co-resident CTAs.  d.e needs no output / statistics pointers for the test.
bool gemm_gn_fuse_supported(const GemmDesc& d);
// Scratch such a launch needs, zeroed before it: per image 32 groups x 2 accumulator words (EpiParams::gn_sums).
inline size_t gn_scratch_bytes(int n_img) { return (size_t)n_img * 512; }

// number of kernels this translation unit has launched since process start (bench.py's gpu_launches claim)
unsigned long long gemm_launch_count();

int device_sm_count();

// CTAs (pair = false) or CTA pairs (pair = true) of the feature-major EPI_F16 kernels co-resident on the current device
int gemm_resident_units(bool pair);

// fp32 [Cout,Cin,kh,kw] -> packed fp16 rows (capi_core.cu)
cudaError_t launch_pack_conv_weight(const float* w32, __half* w16, int Cout, int Cin, int cout_pad, int cin_pad,
                                    int kind, cudaStream_t st);

}  // namespace rgm
