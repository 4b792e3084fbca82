/*
 * rgm_b200 -- C ABI of the B200-native (sm_100a) rule-guided sampling hot path.
 *
 * The reference (yjhuangcd/rule-guided-music) is pure Python/PyTorch and has no FFI of its own; its drop-in boundary
 * for this path is four Python callables.  Every entry point below names the reference
 * callable(s) it replaces.  The Python mirror of the reference API (rule_guided_music_b200/) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*), nothing synchronises
 *   - return 0 on success, negative on error; rgm_last_error() returns the message for the calling thread
 *   - handles own their packed weights and workspace; the workspace grows on first use of a larger batch
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails with an error
 */
#ifndef RGM_B200_H_
#define RGM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------------ */
const char* rgm_last_error(void);
int rgm_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
unsigned long long rgm_launch_count(void);
/* 0 when the current device is sm_100 (B200); negative with an error message otherwise */
int rgm_check_device(void);

/* ---- building blocks (exposed for the parity tests) ---------------------------------------------------------- */
/* out32[M,N] = A16[M,K] . B16[N,K]^T + bias[N]      (torch.nn.functional.linear; reference dit.py:256,286,324-326)
 * block_n: 0 = choose, else 32 / 128 / 256 */
int rgm_gemm_f16(const void* a16, const void* b16, const float* bias, float* out32, int M, int N, int K, int block_n,
                 void* stream);
/* NHWC fp16 convolution with fp32 accumulation           (torch.nn.Conv2d; reference model.py:38-53,78-137)
 * kind: 0 = 1x1, 1 = 3x3 pad 1, 2 = nearest-2x upsample + 3x3 pad 1 (weights packed by rgm_pack_conv_weight)
 * x16 [n,H,W,Cin], out16 [n,H',W',Cout], optional resid16 like out16; gn_part may be NULL */
int rgm_conv_f16(const void* x16, const void* w16_packed, const float* bias, const void* resid16, void* out16,
                 int n_img, int H, int W, int Cin, int Cout, int kind, int block_n, float* gn_part, void* stream);
/* weight fp32 [Cout,Cin,kh,kw] (torch layout) -> packed fp16 rows for rgm_conv_f16; cin_pad >= Cin (multiple of 64),
 * cout_pad >= Cout. Output size: kind 0: cout_pad*cin_pad; kind 1: cout_pad*9*cin_pad; kind 2: 4*cout_pad*4*cin_pad */
int rgm_pack_conv_weight(const float* w32, void* w16_packed, int Cout, int Cin, int cout_pad, int cin_pad, int kind,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGM_B200_H_ */
