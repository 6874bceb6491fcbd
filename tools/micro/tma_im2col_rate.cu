// Micro-benchmark: how fast does the TMA unit deliver an implicit-GEMM A operand?  One thread per CTA (148 CTAs) issues, per
// k-block, the A box of a 128-pixel tile x 32 channels and a tiled B box {32 x rows_b}, into a 6-slot ring with no
// consumer (a slot is re-used as soon as its previous fill has landed), walking tiles and k-blocks in the order the GEMM
// does (tile = blockIdx + i * gridDim; k-block = (tap, 32-channel block)).  Forms of the A load:
//   tiled    cp.async.bulk.tensor.2d over a DENSE [M][K] matrix (what a materialised im2col matrix costs to read)
//   im2col   cp.async.bulk.tensor.4d ... im2col over the NHWC activation (what the implicit GEMM issues)
// Question (round 2, profiles/r02_notes.md section 6): the sub-pixel GEMMs in plain TF32 -- no splitter, 128-cycle MMAs --
// still need ~940 cycles per k-block; is that the im2col-mode TMA rate (~6 cycles per 128-byte pixel row)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multimodal_vae_public_b200/csrc tma_im2col_rate.cu -o tma_im2col_rate -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "ptx.cuh"
using namespace mvae;

constexpr int kStages = 6;

struct Geom {          // im2col walk of the activation
  int OW, OHW;         // output grid of the view
  int lower, stride, taps, cblocks;
};

template <bool kIm2col>
__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ CUtensorMap map_a,
                                                      const __grid_constant__ CUtensorMap map_b, Geom g, int rows_b,
                                                      int kblocks, int tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&map_a); ptx::prefetch_tmap(&map_b);
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full_bar[s], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int stage_bytes = 16384 + rows_b * 128;
    long long t0 = clock64();
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = t * 128;
      const int n = m0 / g.OHW, rem = m0 - n * g.OHW, oh = rem / g.OW, ow = rem - oh * g.OW;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % kStages;
        if (it >= kStages) ptx::mbar_wait(&full_bar[s], ((it / kStages) - 1) & 1);
        ptx::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        if (kIm2col) {
          const int tap = kb / g.cblocks, c0 = (kb - tap * g.cblocks) * 32;
          const int th = tap / g.taps, tw = tap - th * g.taps;
          ptx::tma_load_im2col_4d(smem + s * 32768, &map_a, &full_bar[s], c0, g.lower + ow * g.stride, g.lower + oh * g.stride, n,
                                  static_cast<uint16_t>(tw), static_cast<uint16_t>(th));
        } else {
          ptx::tma_load_2d(smem + s * 32768, &map_a, &full_bar[s], kb * 32, m0);
        }
        ptx::tma_load_2d(smem + s * 32768 + 16384, &map_b, &full_bar[s], kb * 32, 0);
      }
    }
    for (int s = 0; s < kStages; ++s) {          // drain
      const int last = ((it - 1 - s) / kStages) * kStages + s;
      if (it - 1 - s >= 0) ptx::mbar_wait(&full_bar[s], (last / kStages) & 1);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = (t1 - t0) / (it > 0 ? it : 1); out[1] = it; }
  }
  __syncthreads();
}

static void* entry(const char* name) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  return p;
}
static CUtensorMap tiled_map(const float* base, long long inner, long long rows, int box_rows) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)entry("cuTensorMapEncodeTiled"))(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, es,
                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("tiled encode failed %d\n", (int)r);
  return m;
}
static CUtensorMap im2col_map(const float* base, int N, int H, int W, int C, int lower, int upper, int stride) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lo[2] = {lower, lower}, up[2] = {upper, upper};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = ((Enc)entry("cuTensorMapEncodeIm2col"))(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstr, lo, up, 32,
                                                        128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("im2col encode failed %d\n", (int)r);
  return m;
}

template <bool kIm2col>
static void run(const char* name, CUtensorMap ma, CUtensorMap mb, Geom g, int rows_b, int kblocks, int tiles, long long* out) {
  const int smem = kStages * 32768 + 1024;
  auto k = rate_kernel<kIm2col>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, 128, smem>>>(ma, mb, g, rows_b, kblocks, tiles, out);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = double(tiles) * kblocks * (16384.0 + rows_b * 128.0);
    if (rep) printf("%-58s B rows %3d: %5lld cycles per k-block (CTA 0, %lld k-blocks), %7.1f us, %.2f TB/s into smem  (%s)\n", name,
                    rows_b, h[0], h[1], ms * 1e3, bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(e2));
  }
}

int main() {
  // FashionMNIST decoder ConvT1 as sub-pixel problem: x [8192,7,7,128], 2x2 stride-1 view, 16 k-blocks; M = 401408 -> 3136 tiles
  const int N = 8192, H = 7, W = 7, C = 128;
  const long long M = (long long)N * H * W;
  float *x, *dense, *b; long long* out;
  cudaMalloc(&x, M * C * 4); cudaMalloc(&dense, M * 512 * 4); cudaMalloc(&b, 1024 * 512 * 4); cudaMalloc(&out, 64);
  cudaMemset(x, 0, M * C * 4); cudaMemset(dense, 0, M * 512 * 4); cudaMemset(b, 0, 1024 * 512 * 4);
  const int tiles = (int)(M / 128);
  Geom g2{7, 49, -1, 1, 2, 4};
  for (int rows_b : {64, 128}) {
    CUtensorMap mb = tiled_map(b, 512, 1024, rows_b);
    run<false>("tiled 2-D A over a dense [401408][512] matrix (822 MB)", tiled_map(dense, 512, M, 128), mb, g2, rows_b, 16, tiles, out);
    run<false>("tiled 2-D A over a dense [401408][128] matrix (205 MB), x4", tiled_map(x, 128, M, 128), mb, Geom{7, 49, 0, 1, 1, 4}, rows_b, 4, tiles, out);
    run<true>("im2col A, 2x2 stride-1 view of [8192,7,7,128] (205 MB)", im2col_map(x, N, H, W, C, -1, -1, 1), mb, g2, rows_b, 16, tiles, out);
  }
  // FashionMNIST conv2 forward: x [4096,14,14,64], k4 s2 p1 -> 7x7, K = 1024 = 32 k-blocks, M = 200704 -> 1568 tiles
  {
    CUtensorMap mb = tiled_map(b, 1024, 128, 128);
    Geom g4{7, 49, -1, 2, 4, 2};
    run<true>("im2col A, k4 s2 p1 view of [4096,14,14,64] (205 MB)", im2col_map(x, 4096, 14, 14, 64, -1, -2, 2), mb, g4, 128, 32, 1568, out);
  }
  return 0;
}
