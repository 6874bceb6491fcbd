// Label / attribute encoder evaluated on its V-row table (sm_100a).
//
// The reference's TextEncoder (mnist/model.py:108-125, fashionmnist/model.py:124-146) is Embedding(V=10, 512) -> Swish ->
// Linear(512,512) -> Swish -> two heads Linear(512, L): its input takes only V distinct values, so the whole network is
// a function of the CLASS.  The reference (and round 1 of this implementation) evaluates it on all B rows -- two
// [B,512]x[512,512]-class GEMMs forward, four backward, plus a gather and a segmented sum.  Here it runs once per class:
//
//   forward   h1 = swish(emb)            [V, D]
//             a2 = h1 W2^T + b2 ; h2 = swish(a2)                       (stage 1: one warp per output column)
//             tab = h2 W3^T + b3         [V, N3]  (N3 = 2 L: mu | logvar) (stage 2)
//             the PoE kernels read row text[b] of `tab` (mvae_poe_fwd_g: gather index per expert)
//   backward  dtab[v] = sum over {b : text[b] = v} of d(enc)[b]        (red.add inside mvae_poe_bwd_g)
//             dW3 += dtab^T h2 ; db3 += sum_v dtab ; dA2 = (dtab W3) * swish'(a2)          (stage 1)
//             dW2 += dA2^T h1 ; db2 += sum_v dA2 ; d emb += (dA2 W2) * swish'(emb)         (stage 2)
//
// Exact (same sums, regrouped by class; fp32 FMA chains instead of 3xTF32 tensor-core products) and ~100x less work
// at B = 4096.  Everything is tiny (W2 = 1 MB is the largest operand): plain CUDA cores, coalesced 128-bit weight reads,
// V <= 16 rows held in shared memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

constexpr int kMaxV = 16;

__device__ __forceinline__ float sigmoid_x(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_x(float x) { return x * sigmoid_x(x); }
__device__ __forceinline__ float dswish_x(float x) {
  const float s = sigmoid_x(x);
  return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// These kernels are LATENCY-bound (a few hundred KB of operands on 148 SMs): every loop that touches global memory
// issues all of its independent loads before the first use (a first version with one load in flight per iteration took
// 42 us forward / 138 us backward; the data volume is worth ~3 us).

// Stage the V x D operand in shared memory, optionally through Swish: up to 8 independent 128-bit loads per thread in
// flight (V * D <= 8192 floats with 256 threads).
template <bool kSwish>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, float* dst, int n) {
  const int n4 = n >> 2;   // n % 4 == 0 (D % 4 == 0)
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < n4) v[i] = __ldg(reinterpret_cast<const float4*>(src) + idx);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < n4) {
      float4 x = v[i];
      if (kSwish) { x.x = swish_x(x.x); x.y = swish_x(x.y); x.z = swish_x(x.z); x.w = swish_x(x.w); }
      reinterpret_cast<float4*>(dst)[idx] = x;
    }
  }
  for (int idx = threadIdx.x + 8 * blockDim.x; idx < n4; idx += blockDim.x) {   // (larger tables: not latency-critical)
    float4 x = __ldg(reinterpret_cast<const float4*>(src) + idx);
    if (kSwish) { x.x = swish_x(x.x); x.y = swish_x(x.y); x.z = swish_x(x.z); x.w = swish_x(x.w); }
    reinterpret_cast<float4*>(dst)[idx] = x;
  }
}

// out[v][n] = sum_k act(in[v][k]) * W[n][k] + b[n]  for all v < V; one warp per output column n.
//   kSwishIn : in = raw table (embedding), the operand is swish(in)  (forward stage 1); otherwise in is used as is
//   out2     : optional swish(out)
// Shared memory: the V x D operand (V*D floats).
template <bool kSwishIn>
__global__ void __launch_bounds__(256) rows_linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out,
                                                          float* __restrict__ out2, int V, int D, int N) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_in[];   // [V][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  // this warp's weight row first (independent of the staging below): D / 128 loads in flight, D <= 1024 on the fast path
  float4 w[8];
  const float* wrow = W + static_cast<int64_t>(n < N ? n : 0) * D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane * 4 + i * 128;
    w[i] = (k < D) ? __ldg(reinterpret_cast<const float4*>(wrow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float b = (bias && n < N) ? __ldg(bias + n) : 0.f;
  stage_rows<kSwishIn>(in, s_in, V * D);
  __syncthreads();
  if (n >= N) return;
  float acc[kMaxV];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) acc[v] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane * 4 + i * 128;
    if (k < D) {
#pragma unroll
      for (int v = 0; v < kMaxV; ++v) {
        if (v < V) {
          const float4 x = *reinterpret_cast<const float4*>(s_in + v * D + k);
          acc[v] += w[i].x * x.x + w[i].y * x.y + w[i].z * x.z + w[i].w * x.w;
        }
      }
    }
  }
  for (int k = lane * 4 + 1024; k < D; k += 128) {    // D > 1024: plain loop
    const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      if (v < V) {
        const float4 x = *reinterpret_cast<const float4*>(s_in + v * D + k);
        acc[v] += wv.x * x.x + wv.y * x.y + wv.z * x.z + wv.w * x.w;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) {
    if (v < V) {
      const float s = warp_sum_f(acc[v]) + b;
      if (lane == 0) {
        out[static_cast<int64_t>(v) * N + n] = s;
        if (out2) out2[static_cast<int64_t>(v) * N + n] = swish_x(s);
      }
    }
  }
}

// Backward of one table layer  y[v][n] = sum_k x[v][k] W[n][k] + b[n]  given dy [V][N]:
//   blocks [0, gw)          : dW[n][k] += sum_v dy[v][n] x[v][k] ; db[n] += sum_v dy[v][n]    (kWRows rows n per block)
//   blocks [gw, gw + gx)    : dx[v][k] = (sum_n dy[v][n] W[n][k]) * swish'(pre[v][k])  -- an n-chunk per block, so the
//                             results are ADDED (red.add) into dx, which the caller zero-initialises; the factor
//                             swish'(pre) distributes over the chunks.  kSwishX: x = swish(pre) is recomputed from `pre`
//                             (first layer: pre = the embedding table itself), otherwise x is read from `xin`.
// Shared memory: dy [V][N] and (for the dW blocks) x [V][D].
constexpr int kWRows = 4;       // dW rows per block
constexpr int kNChunk = 32;     // n-range of one dx block
template <bool kSwishX>
__global__ void __launch_bounds__(256) rows_linear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xin,
                                                              const float* __restrict__ pre, const float* __restrict__ W,
                                                              float* __restrict__ dW, float* __restrict__ db,
                                                              float* __restrict__ dx, int V, int D, int N, int gw) {
  pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  float* s_dy = smem;             // [V][N]
  float* s_x = smem + V * N;      // [V][D]   (dW blocks only)
  if (static_cast<int>(blockIdx.x) < gw) {
    const int n0 = blockIdx.x * kWRows;
    const int d4 = D >> 2;
    // this thread's dW elements (read-modify-write; this block owns rows [n0, n0 + kWRows)): loads first
    constexpr int kMaxPer = 4;            // kWRows * D / 4 / 256 float4 per thread for D <= 1024
    float4 old[kMaxPer];
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int idx = threadIdx.x + i * blockDim.x;
      const int n = n0 + idx / d4;
      if (idx < kWRows * d4 && n < N)
        old[i] = *reinterpret_cast<const float4*>(dW + static_cast<int64_t>(n) * D + (idx % d4) * 4);
    }
    stage_rows<false>(dy, s_dy, V * N);
    stage_rows<kSwishX>(kSwishX ? pre : xin, s_x, V * D);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int idx = threadIdx.x + i * blockDim.x;
      const int n = n0 + idx / d4, k = (idx % d4) * 4;
      if (idx < kWRows * d4 && n < N) {
        float4 g = old[i];
        for (int v = 0; v < V; ++v) {
          const float d = s_dy[v * N + n];
          const float4 x = *reinterpret_cast<const float4*>(s_x + v * D + k);
          g.x += d * x.x; g.y += d * x.y; g.z += d * x.z; g.w += d * x.w;
        }
        *reinterpret_cast<float4*>(dW + static_cast<int64_t>(n) * D + k) = g;
      }
    }
    for (int idx = threadIdx.x + kMaxPer * blockDim.x; idx < kWRows * d4; idx += blockDim.x) {   // D > 1024
      const int n = n0 + idx / d4, k = (idx % d4) * 4;
      if (n >= N) continue;
      float4* p = reinterpret_cast<float4*>(dW + static_cast<int64_t>(n) * D + k);
      float4 g = *p;
      for (int v = 0; v < V; ++v) {
        const float d = s_dy[v * N + n];
        const float4 x = *reinterpret_cast<const float4*>(s_x + v * D + k);
        g.x += d * x.x; g.y += d * x.y; g.z += d * x.z; g.w += d * x.w;
      }
      *p = g;
    }
    if (db != nullptr && threadIdx.x < kWRows && n0 + threadIdx.x < N) {
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += s_dy[v * N + n0 + threadIdx.x];
      db[n0 + threadIdx.x] += s;
    }
    return;
  }
  if (dx == nullptr) return;
  // dx blocks: blockIdx - gw = chunk * kblocks + kb ; each thread owns one column k (coalesced W reads along k)
  const int kblocks = (D + blockDim.x - 1) / blockDim.x;
  const int rel = blockIdx.x - gw;
  const int chunk = rel / kblocks, kb = rel - chunk * kblocks;
  const int k = kb * blockDim.x + threadIdx.x;
  const int nb = chunk * kNChunk;
  float w[kNChunk];
#pragma unroll
  for (int j = 0; j < kNChunk; ++j)   // the whole chunk's weights in flight at once
    w[j] = (k < D && nb + j < N) ? __ldg(W + static_cast<int64_t>(nb + j) * D + k) : 0.f;
  float prev[kMaxV];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) prev[v] = (v < V && k < D) ? __ldg(pre + static_cast<int64_t>(v) * D + k) : 0.f;
  stage_rows<false>(dy, s_dy, V * N);
  __syncthreads();
  if (k >= D) return;
  float acc[kMaxV];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) acc[v] = 0.f;
#pragma unroll
  for (int j = 0; j < kNChunk; ++j) {
    if (nb + j < N) {
#pragma unroll
      for (int v = 0; v < kMaxV; ++v)
        if (v < V) acc[v] += s_dy[v * N + nb + j] * w[j];
    }
  }
#pragma unroll
  for (int v = 0; v < kMaxV; ++v)
    if (v < V) atomicAdd(dx + static_cast<int64_t>(v) * D + k, acc[v] * dswish_x(prev[v]));
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_label_table_fwd(const float* emb, const float* w2, const float* b2, const float* w3, const float* b3,
                                    float* a2, float* h2, float* tab, int V, int D, int N3, void* stream) {
  if (!emb || !w2 || !w3 || !a2 || !h2 || !tab || V < 1 || V > kMaxV || D < 4 || (D & 3) || N3 < 1)
    return set_error(MVAE_ERR_BAD_ARG, "label_table_fwd: bad arguments (V <= %d, D %% 4 == 0)", kMaxV);
  if ((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w3)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "label_table_fwd: weights must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(V) * D * sizeof(float);
  if (smem > 48 * 1024 || V * D > 8 * 256 * 4)
    return set_error(MVAE_ERR_UNSUPPORTED, "label_table_fwd: V*D too large (<= 8192 floats)");
  launch_pdl(rows_linear_kernel<true>, dim3((D + 7) / 8), dim3(256), smem, st, emb, w2, b2, a2, h2, V, D, D);
  launch_pdl(rows_linear_kernel<false>, dim3((N3 + 7) / 8), dim3(256), smem, st, h2, w3, b3, tab, static_cast<float*>(nullptr), V, D, N3);
  count_launch(2);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_label_table_bwd(const float* emb, const float* w2, const float* w3, const float* a2, const float* h2,
                                    const float* dtab, float* d_a2, float* d_emb, float* dw2, float* db2, float* dw3,
                                    float* db3, int V, int D, int N3, void* stream) {
  if (!emb || !w2 || !w3 || !a2 || !h2 || !dtab || !d_a2 || !d_emb || !dw2 || !dw3 || V < 1 || V > kMaxV || D < 4 ||
      (D & 3) || N3 < 1)
    return set_error(MVAE_ERR_BAD_ARG, "label_table_bwd: bad arguments (V <= %d, D %% 4 == 0)", kMaxV);
  if ((reinterpret_cast<uintptr_t>(dw2) | reinterpret_cast<uintptr_t>(dw3)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "label_table_bwd: weight gradients must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int kblocks = (D + 255) / 256;
  // stage 1: heads.  d_a2 is an accumulation target (the n-chunks add into it): zero-initialised by the caller.
  {
    const int gw = (N3 + kWRows - 1) / kWRows, gx = ((N3 + kNChunk - 1) / kNChunk) * kblocks;
    const size_t smem = (static_cast<size_t>(V) * N3 + static_cast<size_t>(V) * D) * sizeof(float);
    if (smem > 48 * 1024) return set_error(MVAE_ERR_UNSUPPORTED, "label_table_bwd: V*(D+N3) too large for shared memory");
    launch_pdl(rows_linear_bwd_kernel<false>, dim3(gw + gx), dim3(256), smem, st, dtab, h2, a2, w3, dw3, db3, d_a2, V, D, N3, gw);
  }
  // stage 2: the hidden layer; its input is swish(embedding), the embedding gradient accumulates into d_emb (the caller's
  // gradient buffer, zeroed with the rest of the bucket)
  {
    const int gw = (D + kWRows - 1) / kWRows, gx = ((D + kNChunk - 1) / kNChunk) * kblocks;
    const size_t smem = 2 * static_cast<size_t>(V) * D * sizeof(float);
    if (smem > 48 * 1024) return set_error(MVAE_ERR_UNSUPPORTED, "label_table_bwd: V*D too large for shared memory");
    launch_pdl(rows_linear_bwd_kernel<true>, dim3(gw + gx), dim3(256), smem, st, d_a2, static_cast<const float*>(nullptr), emb, w2, dw2, db2, d_emb, V, D, D, gw);
  }
  count_launch(2);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
