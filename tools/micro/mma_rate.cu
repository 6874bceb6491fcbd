// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128 per CTA, N=128, K=8) on one SM / one CTA pair.
// Forms: SS (A,B from smem), TS (A from TMEM), cta_group::1 and ::2.  Operands are whatever is in smem (zeros).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multimodal_vae_public_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mvae;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}

// mode bits: 1 = the real 3x operand pattern (a_lo*b, a_hi*b_lo, a_hi*b over 4 k-slices, 4 rotating stages)
//            2 = warps 2,3 stream shared memory (LDS.128 + STS.128 over a 32 KiB window) while the MMAs run
//            4 = warps 2,3 issue tcgen05.st (32 lanes x 64 columns each) in a loop while the MMAs run
template <bool kPair, bool kTS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int n_mma, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); tmem_base_smem[1] = 0; }
  if (warp == 0) {
    if (kPair) { ptx::tmem_alloc2(&tmem_base_smem[0], 512); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(&tmem_base_smem[0], 512); ptx::tmem_relinquish(); }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem[0];
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t M = kPair ? 256 : 128;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((M >> 4) << 24);
    const uint32_t sa = ptx::smem_u32(smem), sb = sa + 16384;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        int ks = i & 3;
        uint32_t boff = 0, aoff = 0;
        if (mode & 1) {
          const int prod = i % 3, kk = (i / 3) & 3, stage = (i / 12) & 3;
          ks = kk;
          boff = stage * 24576 + (prod == 1 ? 8192 : 0);     // b_lo tile for the middle product
          aoff = stage * 64 + (prod == 0 ? 32 : 0);          // a_lo for the first product
        }
        const uint64_t db = make_desc(sb + boff + ks * 32, 16, 1024, 2);
        if (kTS) {
          if (kPair) ptx::mma_tf32_ts2(tmem_base, tmem_base + 256 + aoff + ks * 8, db, idesc, 1u);
          else ptx::mma_tf32_ts(tmem_base, tmem_base + 256 + aoff + ks * 8, db, idesc, 1u);
        } else {
          const uint64_t da = make_desc(sa + ks * 32, 16, 1024, 2);
          ptx::mma_tf32_ss(tmem_base, da, db, idesc, 1u);
        }
      }
      if (kPair) ptx::mma_commit2(&bar); else ptx::mma_commit(&bar);
      const long long t1 = clock64();
      ptx::mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2 * it] = t1 - t0; out[2 * it + 1] = t2 - t0; }
    }
    *reinterpret_cast<volatile int*>(&tmem_base_smem[1]) = 1;
  }
  if (warp >= 2 && (mode & 6)) {
    // background traffic until the MMA thread is done (it sets `done` in smem)
    volatile int* done = reinterpret_cast<volatile int*>(&tmem_base_smem[1]);
    float4* win = reinterpret_cast<float4*>(smem + 131072);
    uint32_t regs[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) regs[j] = j;
    const uint32_t ta = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + 128;
    long long n = 0;
    while (!*done) {
      if (mode & 2) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = win[(lane + 64 * j + 32 * (warp - 2)) & 2047];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[(lane + 64 * j + 32 * (warp - 2) + 1024) & 2047] = v[j];
      }
      if (mode & 4) {
        ptx::tmem_st_32x32(ta, regs);
        ptx::tmem_st_32x32(ta + 32, regs);
        ptx::tmem_st_wait();
      }
      ++n;
    }
    if (blockIdx.x == 0 && lane == 0) out[32 + warp] = n;
  }
  if (kPair && rank == 1 && warp == 1 && lane == 0) {   // the peer's barrier also receives the multicast commits
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) { ptx::mbar_wait(&bar, phase); phase ^= 1; }
    *reinterpret_cast<volatile int*>(&tmem_base_smem[1]) = 1;
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <bool kPair, bool kTS>
void run(const char* name, int grid, int mode) {
  long long* out; cudaMalloc(&out, 64 * sizeof(long long)); cudaMemset(out, 0, 64 * sizeof(long long));
  const int smem = 200 * 1024;
  auto k = rate_kernel<kPair, kTS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n_mma : {768}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kPair ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, out, 4, n_mma, mode);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[40]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-24s mode=%d grid=%3d n_mma=%4d  issue %lld cyc, complete %lld cyc -> %.1f cyc/MMA  traffic iters %lld/%lld = %.0f B/clk smem (%s %s)\n",
           name, mode, grid, n_mma, h[6], h[7], double(h[7]) / n_mma, h[34], h[35],
           (mode & 2) ? double(h[34] + h[35]) * 4 * 8192.0 / 4 / double(h[1] + h[3] + h[5] + h[7]) * 1.0 : 0.0,
           cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  cudaFree(out);
}

int main() {
  for (int mode : {0, 1, 3, 5, 7}) {
    run<false, false>("cta1 SS (A,B smem)", 148, mode);
    run<false, true>("cta1 TS (A tmem)", 148, mode);
    run<true, true>("cta2 TS (A tmem, pair)", 148, mode);
  }
  return 0;
}
