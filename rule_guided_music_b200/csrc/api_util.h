// Shared helpers for the C ABI translation units: thread-local error string, launch accounting and the optional
// per-launch device timer behind rgm_prof_* (CUDA events on the launching stream; off by default).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

namespace rgm {
extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_aux_launches;
int set_error(const std::string& m);
int check_cuda(cudaError_t e, const char* what);

extern std::atomic<int> g_prof_on;
void prof_open(const char* name, double flops_alg, double flops_exec, double bytes, cudaStream_t st);
void prof_close(cudaStream_t st);

// Times everything enqueued on `st` during its lifetime when profiling is on; free otherwise.
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(const char* name, double flops_alg, double flops_exec, double bytes, cudaStream_t s)
      : st(s), on(g_prof_on.load(std::memory_order_relaxed) != 0) {
    if (on) prof_open(name, flops_alg, flops_exec, bytes, st);
  }
  ~ProfScope() {
    if (on) prof_close(st);
  }
};
}  // namespace rgm

namespace rgm {
// Device buffer that only ever grows.  A buffer that a CUDA graph may have captured is never freed while its handle
// lives: growing allocates a new block and RETIRES the old one (freed with the owner), so a graph captured against the
// old block keeps replaying into valid memory.  No synchronisation, no cudaFree on the hot path.  Growing while the
// stream is being captured is refused with a message (cudaMalloc is illegal under a global-mode capture): callers
// pre-size with rgm_dit_reserve / rgm_vae_reserve or run the shape once eagerly first.
struct GrowBuf {
  void* p = nullptr;
  size_t bytes = 0;
  std::vector<void*> retired;
  unsigned generation = 0;  // bumps on every growth (diagnostics / graph keys)
  cudaError_t reserve(size_t need, cudaStream_t st = nullptr) {
    if (need <= bytes) return cudaSuccess;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (st != nullptr && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
      return cudaErrorStreamCaptureUnsupported;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, need);
    if (e != cudaSuccess) return e;
    if (p) retired.push_back(p);
    p = q;
    bytes = need;
    ++generation;
    return cudaSuccess;
  }
  ~GrowBuf() {
    if (p) cudaFree(p);
    for (void* r : retired) cudaFree(r);
  }
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: remember, per device, the largest size a
// kernel has been opted into (one process may drive several GPUs).
struct SmemAttr {
  std::atomic<size_t> set[64];
  SmemAttr() {
    for (auto& s : set) s.store(0);
  }
  template <typename K>
  cudaError_t ensure(K kern, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 63;
    if (bytes <= set[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev != 63) set[dev].store(bytes, std::memory_order_relaxed);
    return e;
  }
};
}  // namespace rgm

#define RGM_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return rgm::check_cuda(_e, #expr); \
  } while (0)
