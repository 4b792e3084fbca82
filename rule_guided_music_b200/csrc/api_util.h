// Shared helpers for the C ABI translation units: thread-local error string, launch accounting and the optional
// per-launch device timer behind rgm_prof_* (CUDA events on the launching stream; off by default).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

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

#define RGM_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return rgm::check_cuda(_e, #expr); \
  } while (0)
