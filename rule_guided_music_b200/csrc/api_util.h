// Shared helpers for the C ABI translation units: thread-local error string and launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

namespace rgm {
extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_aux_launches;
int set_error(const std::string& m);
int check_cuda(cudaError_t e, const char* what);
}  // namespace rgm

#define RGM_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return rgm::check_cuda(_e, #expr); \
  } while (0)
