// Probe of TMA im2col-mode loads (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d...im2col) on sm_100a, the building
// block of the implicit-GEMM convolution: does one instruction deliver `pixels` consecutive OUTPUT pixels (crossing row
// and image boundaries) x `channels` input channels of filter tap (kh, kw), zero-filled at the padding, in the same
// shared-memory layout as a tiled {channels, pixels} box?  Checked against a CPU im2col for
//   (a) K-major operand tiles: 128 pixels x 32 channels, SWIZZLE_128B                  (conv forward / dgrad-as-conv A)
//   (b) MN-major operand boxes: 32 pixels x 32 channels, SWIZZLE_128B_ATOM_32B          (wgrad operands)
// for k4 s2 p1 (OW = 7: tiles straddle rows and images) and for the 2x2 / stride-1 sub-pixel taps of a transposed conv.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 im2col_probe.cu -o im2col_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, int c0, int w0, int h0, int n0, int offw, int offh,
                             int bytes, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(w0), "r"(h0), "r"(n0),
          "h"(static_cast<uint16_t>(offw)), "h"(static_cast<uint16_t>(offh))
        : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (clock64() - t0 > (1LL << 31)) { out[0] = -12345.f; return; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<const float*>(smem)[i];
}

static EncIm2col get_enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q);
  return reinterpret_cast<EncIm2col>(p);
}

// x: [N][H][W][C] floats.  Returns #mismatches of one probe.
static int run_case(const char* name, const float* dx, const std::vector<float>& hx, int N, int H, int W, int C, int lower,
                    int upper, int stride, int pixels, int channels, CUtensorMapSwizzle swz, bool mn_major, int m0, int c0,
                    int kh, int kw, int OW, int OH) {
  EncIm2col enc = get_enc();
  if (!enc) { printf("%s: cuTensorMapEncodeIm2col not available\n", name); return -1; }
  CUtensorMap map;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lo[2] = {lower, lower}, up[2] = {upper, upper};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dx), gdim, gstr, lo, up, channels, pixels, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%-46s encode failed (%d)\n", name, (int)r); return -1; }
  const int bytes = pixels * channels * 4;
  float* dout; cudaMalloc(&dout, bytes); cudaMemset(dout, 0xff, bytes);
  const int n = m0 / (OH * OW), rem = m0 % (OH * OW), p = rem / OW, q = rem % OW;
  probe_kernel<<<1, 128, bytes + 1024>>>(map, c0, lower + q * stride, lower + p * stride, n, kw, kh, bytes, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-46s kernel failed: %s\n", name, cudaGetErrorString(e)); return -1; }
  std::vector<float> got(bytes / 4);
  cudaMemcpy(got.data(), dout, bytes, cudaMemcpyDeviceToHost);
  cudaFree(dout);
  if (got[0] == -12345.f) { printf("%-46s TIMEOUT (barrier never completed)\n", name); return -1; }
  int bad = 0;
  for (int pi = 0; pi < pixels; ++pi) {
    const int m = m0 + pi;
    const int nn = m / (OH * OW), rr = m % (OH * OW), pp = rr / OW, qq = rr % OW;
    const int iy = lower + pp * stride + kh, ix = lower + qq * stride + kw;
    for (int ci = 0; ci < channels; ++ci) {
      float ref = 0.f;
      if (nn < N && iy >= 0 && iy < H && ix >= 0 && ix < W) ref = hx[(((size_t)nn * H + iy) * W + ix) * C + c0 + ci];
      size_t idx;
      if (!mn_major) {   // row = pixel (128 B), 16-byte chunks XOR (row & 7)   [SWIZZLE_128B]
        idx = (size_t)pi * 32 + ((((ci >> 2) ^ (pi & 7)) << 2) | (ci & 3));
      } else {           // row = pixel (k), 32 m-values per row (128 B), 32-byte chunks XOR (row & 3)  [SWIZZLE_128B_ATOM_32B]
        idx = (size_t)pi * 32 + ((((ci >> 3) ^ (pi & 3)) << 3) | (ci & 7));
      }
      if (got[idx] != ref) { if (bad < 4) printf("   mismatch pixel %d ch %d: got %g want %g\n", pi, ci, got[idx], ref); ++bad; }
    }
  }
  printf("%-46s m0=%6d tap(%d,%d) c0=%3d : %s (%d mismatches of %d)\n", name, m0, kh, kw, c0, bad ? "WRONG" : "ok", bad,
         pixels * channels);
  return bad;
}

int main() {
  const int N = 5, H = 14, W = 14, C = 64;
  std::vector<float> hx((size_t)N * H * W * C);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u) % 100003) / 1000.0f + 1.0f;
  float* dx; cudaMalloc(&dx, hx.size() * 4); cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  int bad = 0;
  // k4 s2 p1: lower = -pad = -1, upper = pad - (k - 1) = -2, traversal stride 2, OH = OW = 7
  for (int m0 : {0, 128, 200}) {
    bad += run_case("k4s2p1 K-major 128x32 SWIZZLE_128B", dx, hx, N, H, W, C, -1, -2, 2, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, m0, 32, 0, 0, 7, 7) != 0;
    bad += run_case("k4s2p1 K-major 128x32 SWIZZLE_128B", dx, hx, N, H, W, C, -1, -2, 2, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, m0, 0, 3, 2, 7, 7) != 0;
    bad += run_case("k4s2p1 MN-major 32x32 SWIZZLE_128B_ATOM_32B", dx, hx, N, H, W, C, -1, -2, 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, true, m0 + 32, 32, 1, 3, 7, 7) != 0;
  }
  // sub-pixel taps of a transposed conv (2x2, stride 1) on the same tensor: parity 0: lower = upper = -1; parity 1: 0 / 0
  bad += run_case("k2s1 parity0 K-major 128x32", dx, hx, N, H, W, C, -1, -1, 1, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, 190, 0, 1, 0, 14, 14) != 0;
  bad += run_case("k2s1 parity1 K-major 128x32", dx, hx, N, H, W, C, 0, 0, 1, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, 190, 32, 1, 1, 14, 14) != 0;
  bad += run_case("k2s1 parity1 MN-major 32x32", dx, hx, N, H, W, C, 0, 0, 1, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, true, 64, 0, 0, 1, 14, 14) != 0;
  // strided row gather (1x1 filter, stride 2, lower = parity): rows of one sub-pixel class of a [N,14,14,64] tensor
  bad += run_case("k1s2 class(1,1) MN-major 32x32", dx, hx, N, H, W, C, 1, 0, 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, true, 40, 0, 0, 0, 7, 7) != 0;
  bad += run_case("k1s2 class(0,0) K-major 128x32", dx, hx, N, H, W, C, 0, -1, 2, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, 100, 32, 0, 0, 7, 7) != 0;
  // tail tile: pixels beyond the last image must come back as zeros
  bad += run_case("k4s2p1 K-major tail tile", dx, hx, N, H, W, C, -1, -2, 2, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B, false, 5 * 49 - 60, 0, 2, 1, 7, 7) != 0;
  printf("im2col probe: %d failing cases\n", bad);
  return bad != 0;
}
