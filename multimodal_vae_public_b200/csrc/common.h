// Shared host-side helpers of libmvae_b200: error reporting across the C ABI and launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace mvae {
// Records a formatted message retrievable through mvae_last_error(); returns `code`.
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

// ---- Programmatic dependent launch (PDL).  A training step is ~18 short kernels in one stream (one CUDA graph): without
// PDL every kernel node starts only after its predecessor has drained and been retired (~2-4 us of launch latency and
// prologue per node -- barrier init, TMEM allocation, tensor-map prefetch for the GEMM).  Kernels that begin with
// pdl_prologue() (or, for the GEMM, pdl_trigger() at entry and pdl_wait() after its prologue) and are launched through
// launch_pdl() may be scheduled while the predecessor is still running: `griddepcontrol.launch_dependents` lets the NEXT
// kernel's CTAs become resident early, `griddepcontrol.wait` blocks until every prerequisite grid has completed and its
// memory is visible, so nothing is read or written early.  Both instructions are no-ops in a normally launched kernel,
// and a normally launched kernel after a PDL kernel keeps the full stream-order dependency, so the two kinds mix freely.
// The attribute is OPT-IN (MVAE_PDL=1): measured neutral to slightly negative at the BASELINE batch sizes (common.cu).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif
}  // namespace mvae

#define MVAE_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::mvae::set_error(MVAE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
  } while (0)
