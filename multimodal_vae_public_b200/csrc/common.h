// Shared host-side helpers of libmvae_b200: error reporting across the C ABI and launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvae {
// Records a formatted message retrievable through mvae_last_error(); returns `code`.
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
}  // namespace mvae

#define MVAE_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::mvae::set_error(MVAE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
  } while (0)
