// Micro-benchmark: issue cadence and delivered bandwidth of cp.async.bulk.tensor.2d (TMA) loads of GEMM operand tiles,
// launched with and without a thread-block cluster.  Open question from round 1: the CTA-pair
// GEMM's producer needs ~950 cycles per k-block for two TMA instructions even while its ring is empty, the 1-CTA kernel
// ~550.  One thread per CTA issues, per "k-block", an A box {32 floats x 128 rows} and a B box {32 x rows_b} into a
// ring of `stages` slots, waits for the slot's previous fill (mbarrier) before reuse, and records clock64 after every
// issue.  No consumer: the only limits are the TMA unit, L2 and shared-memory write bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multimodal_vae_public_b200/csrc tma_issue.cu -o tma_issue -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace mvae;

constexpr int kStages = 6;

template <bool kCluster>
__global__ void __launch_bounds__(128, 1) issue_kernel(const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b, int rows_b, int kblocks,
                                                       int m_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&map_a); ptx::prefetch_tmap(&map_b);
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full_bar[s], 1);
    ptx::fence_mbar_init();
  }
  if (kCluster) ptx::cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) {
    const int stage_bytes = 16384 + rows_b * 128;
    const int m0 = (blockIdx.x % m_tiles) * 128;
    long long t_prev = clock64(), sum = 0, mx = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % kStages;
      if (kb >= kStages) ptx::mbar_wait(&full_bar[s], ((kb / kStages) - 1) & 1);   // previous fill of this slot landed
      ptx::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
      ptx::tma_load_2d(smem + s * 32768, &map_a, &full_bar[s], (kb * 32) % 512, m0);
      ptx::tma_load_2d(smem + s * 32768 + 16384, &map_b, &full_bar[s], (kb * 32) % 512, 0);
      const long long t = clock64();
      if (kb >= 2 * kStages) { sum += t - t_prev; mx = (t - t_prev) > mx ? (t - t_prev) : mx; }
      t_prev = t;
    }
    for (int s = 0; s < kStages; ++s) {          // drain
      const int last = ((kblocks - 1 - s) / kStages) * kStages + s;
      if (last >= 0 && last < kblocks) ptx::mbar_wait(&full_bar[s], (last / kStages) & 1);
    }
    if (blockIdx.x == 0) { out[0] = sum / (kblocks - 2 * kStages); out[1] = mx; }
  }
  if (kCluster) ptx::cluster_sync_all(); else __syncthreads();
}

static CUtensorMap make_map(const float* base, int inner, int rows, int box_rows) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return m;
}

template <bool kCluster>
void run(const char* name, const float* a, const float* b, int rows_b, long long* out) {
  const int M = 8192, K = 512, kblocks = 600, grid = 148;
  CUtensorMap ma = make_map(a, K, M, 128), mb = make_map(b, K, 512, rows_b);
  const int smem = kStages * 32768 + 1024;
  auto k = issue_kernel<kCluster>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kCluster ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, ma, mb, rows_b, kblocks, M / 128, out);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = double(grid) * kblocks * (16384.0 + rows_b * 128.0);
    if (rep) printf("%-34s B rows %3d: %5lld cycles per k-block issue (max %6lld), %.2f TB/s delivered  (%s %s)\n", name,
                    rows_b, h[0], h[1], bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
}

int main() {
  float *a, *b; long long* out;
  cudaMalloc(&a, 8192 * 512 * 4); cudaMalloc(&b, 512 * 512 * 4); cudaMalloc(&out, 64);
  cudaMemset(a, 0, 8192 * 512 * 4); cudaMemset(b, 0, 512 * 512 * 4);
  for (int rows_b : {128, 64}) {
    run<false>("no cluster (1-CTA kernel's loads)", a, b, rows_b, out);
    run<true>("cluster of 2 (pair kernel's loads)", a, b, rows_b, out);
  }
  return 0;
}
