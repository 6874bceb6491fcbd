// Train-mode BatchNorm (2d and 1d) + Swish, Dropout and NCHW->NHWC staging for the CelebA-flavour MVAE
// (celeba/model.py:76-92,113-126,145-153,172-183).  Activations are [rows, C] (NHWC flattened); the rows of a
// launch are split into S equal SEGMENTS (one per stacked pass): statistics are per (segment, channel) exactly as the
// reference computes them per model() call, running statistics are updated once per call in the reference's call
// order.  All kernels are bandwidth-bound column reductions / element-wise passes with float4 access along C.
//
//   forward : bn_stats (sum, sum^2 -> double) -> bn_finalize (mean, invstd, running stats) -> bn_apply (+Swish)
//   backward: bn_bwd_reduce (sum da, sum da*xhat per segment) -> bn_bwd_apply (dx; dgamma/dbeta by the finalizer)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

constexpr int kRowsPerBlock = 128;
constexpr int kMaxSeg = 32;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// grid = (S * chunks_per_seg, ceil(C/32)); block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int64_t ld, int seg_rows, int C,
                                                       int chunks_per_seg, double* __restrict__ acc /*[S][C][2]*/) {
  __shared__ float s1[8][33], s2[8][33];
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int seg = blockIdx.x / chunks_per_seg, chunk = blockIdx.x % chunks_per_seg;
  const int r0 = chunk * kRowsPerBlock, r1 = min(seg_rows, r0 + kRowsPerBlock);
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float* base = x + (static_cast<int64_t>(seg) * seg_rows) * ld + c;
    for (int r = r0 + rl; r < r1; r += 8) {
      const float v = base[static_cast<int64_t>(r) * ld];
      a += v; b += v * v;
    }
  }
  s1[rl][threadIdx.x & 31] = a; s2[rl][threadIdx.x & 31] = b;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += s1[i][threadIdx.x & 31]; tb += s2[i][threadIdx.x & 31]; }
    atomicAdd(acc + (static_cast<int64_t>(seg) * C + c) * 2, ta);
    atomicAdd(acc + (static_cast<int64_t>(seg) * C + c) * 2 + 1, tb);
  }
}

struct BnOrder {
  int n;
  int seg[kMaxSeg];
};

// one thread per channel: finalise the S segments, then apply the running-stat updates in the given call order
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int S, int C, int seg_rows, float eps, float momentum,
                                   float* mean /*[S][C]*/, float* invstd /*[S][C]*/, float* running_mean,
                                   float* running_var, BnOrder order) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float var_unb[kMaxSeg], mu[kMaxSeg];
  const double n = static_cast<double>(seg_rows);
  for (int s = 0; s < S; ++s) {
    const double m = acc[(static_cast<int64_t>(s) * C + c) * 2] / n;
    double v = acc[(static_cast<int64_t>(s) * C + c) * 2 + 1] / n - m * m;
    v = v < 0.0 ? 0.0 : v;
    mean[s * C + c] = static_cast<float>(m);
    invstd[s * C + c] = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
    mu[s] = static_cast<float>(m);
    var_unb[s] = static_cast<float>(seg_rows > 1 ? v * n / (n - 1.0) : v);
  }
  if (running_mean != nullptr) {
    float rm = running_mean[c], rv = running_var[c];
    for (int i = 0; i < order.n; ++i) {
      const int s = order.seg[i];
      rm = (1.0f - momentum) * rm + momentum * mu[s];
      rv = (1.0f - momentum) * rv + momentum * var_unb[s];
    }
    running_mean[c] = rm; running_var[c] = rv;
  }
}

// eval mode: every segment uses the running statistics
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int S,
                                     int C, float eps, float* mean, float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int s = 0; s < S; ++s) {
    mean[s * C + c] = running_mean[c];
    invstd[s * C + c] = rsqrtf(running_var[c] + eps);
  }
}

// y = gamma * (x - mean) * invstd + beta ; h = swish(y) (or y if !act).  float4 along C.
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, int64_t ldx, float* h, int64_t ldh,
                                                       int rows, int seg_rows, int C4, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int act) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(rows) * C4) return;
  const int r = static_cast<int>(gid / C4);
  const int c = static_cast<int>(gid - static_cast<int64_t>(r) * C4) * 4;
  const int s = r / seg_rows;
  const int C = C4 * 4;
  const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * ldx + c);
  const float4 m = *reinterpret_cast<const float4*>(mean + s * C + c);
  const float4 is = *reinterpret_cast<const float4*>(invstd + s * C + c);
  const float4 g = *reinterpret_cast<const float4*>(gamma + c);
  const float4 b = *reinterpret_cast<const float4*>(beta + c);
  float y[4] = {g.x * ((xv.x - m.x) * is.x) + b.x, g.y * ((xv.y - m.y) * is.y) + b.y, g.z * ((xv.z - m.z) * is.z) + b.z,
                g.w * ((xv.w - m.w) * is.w) + b.w};
  if (act) {
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q] = y[q] * sigmoid_f(y[q]);
  }
  *reinterpret_cast<float4*>(h + static_cast<int64_t>(r) * ldh + c) = make_float4(y[0], y[1], y[2], y[3]);
}

// da = dh * swish'(a), a = gamma*xhat + beta ; per (segment, channel): acc2 += {sum da, sum da*xhat}
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ dh, int64_t lddh, int seg_rows, int C,
                                                            int chunks_per_seg, int seg0, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int act, double* __restrict__ acc2 /*[S][C][2]*/) {
  __shared__ float s1[8][33], s2[8][33];
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int seg = seg0 + blockIdx.x / chunks_per_seg, chunk = blockIdx.x % chunks_per_seg;
  const int r0 = chunk * kRowsPerBlock, r1 = min(seg_rows, r0 + kRowsPerBlock);
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float m = mean[seg * C + c], is = invstd[seg * C + c], g = gamma[c], b = beta[c];
    const int64_t row0 = static_cast<int64_t>(seg) * seg_rows;
    for (int r = r0 + rl; r < r1; r += 8) {
      const float xh = (x[(row0 + r) * ldx + c] - m) * is;
      float d = dh[(row0 + r) * lddh + c];
      if (act) {
        const float a = g * xh + b;
        const float sg = sigmoid_f(a);
        d *= sg * (1.0f + a * (1.0f - sg));
      }
      a1 += d; a2 += d * xh;
    }
  }
  s1[rl][threadIdx.x & 31] = a1; s2[rl][threadIdx.x & 31] = a2;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += s1[i][threadIdx.x & 31]; tb += s2[i][threadIdx.x & 31]; }
    atomicAdd(acc2 + (static_cast<int64_t>(seg) * C + c) * 2, ta);
    atomicAdd(acc2 + (static_cast<int64_t>(seg) * C + c) * 2 + 1, tb);
  }
}

// dgamma[c] += sum_seg sum(da*xhat), dbeta[c] += sum_seg sum(da) over the live segments [seg0, seg0+nseg)
__global__ void bn_bwd_params_kernel(const double* __restrict__ acc2, int C, int seg0, int nseg, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, b = 0.0;
  for (int s = seg0; s < seg0 + nseg; ++s) {
    b += acc2[(static_cast<int64_t>(s) * C + c) * 2];
    g += acc2[(static_cast<int64_t>(s) * C + c) * 2 + 1];
  }
  dgamma[c] += static_cast<float>(g);
  dbeta[c] += static_cast<float>(b);
}

// dx = gamma*invstd * (da - mean_seg(da) - xhat * mean_seg(da*xhat)); rows [seg0*seg_rows, (seg0+nseg)*seg_rows)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, int64_t ldx,
                                                           const float* __restrict__ dh, int64_t lddh, float* dx,
                                                           int64_t lddx, int row_begin, int rows, int seg_rows, int C4,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int act, const double* __restrict__ acc2, int batch_stats) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(rows) * C4) return;
  const int r = row_begin + static_cast<int>(gid / C4);
  const int c = static_cast<int>(gid % C4) * 4;
  const int s = r / seg_rows;
  const int C = C4 * 4;
  const float inv_n = 1.0f / static_cast<float>(seg_rows);
  const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * ldx + c);
  const float4 dv = *reinterpret_cast<const float4*>(dh + static_cast<int64_t>(r) * lddh + c);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
  const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
  float out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float m = mean[s * C + c + q], is = invstd[s * C + c + q], g = gamma[c + q], b = beta[c + q];
    const float xh = (xs[q] - m) * is;
    float d = ds[q];
    if (act) {
      const float a = g * xh + b;
      const float sg = sigmoid_f(a);
      d *= sg * (1.0f + a * (1.0f - sg));
    }
    if (batch_stats) {
      const float m1 = static_cast<float>(acc2[(static_cast<int64_t>(s) * C + c + q) * 2]) * inv_n;
      const float m2 = static_cast<float>(acc2[(static_cast<int64_t>(s) * C + c + q) * 2 + 1]) * inv_n;
      out[q] = g * is * (d - m1 - xh * m2);
    } else {  // eval mode: fixed (running) statistics -> BatchNorm is an affine map
      out[q] = g * is * d;
    }
  }
  *reinterpret_cast<float4*>(dx + static_cast<int64_t>(r) * lddx + c) = make_float4(out[0], out[1], out[2], out[3]);
}

// ---------------------------------------------------------------- dropout: y[r,:] = x[r % x_rows,:] * mask / (1-p)
__device__ __forceinline__ uint32_t hash32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__global__ void dropout_kernel(const float* __restrict__ x, int x_rows, float* y, float* mask, const float* mask_in,
                               int64_t n, int D, float p, uint32_t seed, const int32_t* step_dev) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = i / D;
  const int c = static_cast<int>(i - r * D);
  float keep;
  if (mask_in != nullptr) keep = mask_in[i];
  else {
    const uint32_t h = hash32(static_cast<uint32_t>(i), seed, step_dev ? static_cast<uint32_t>(*step_dev) : 0u);
    keep = (static_cast<float>(h >> 8) * (1.0f / 16777216.0f)) >= p ? 1.0f : 0.0f;
  }
  if (mask != nullptr) mask[i] = keep;
  y[i] = x[(r % x_rows) * D + c] * keep * (1.0f / (1.0f - p));
}
// dx[r0,:] = sum over the stacked copies of dy * mask / (1-p)
__global__ void dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ mask, float* dx, int x_rows,
                                   int copies, int D, float p) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(x_rows) * D) return;
  float s = 0.f;
  for (int k = 0; k < copies; ++k) {
    const int64_t j = static_cast<int64_t>(k) * x_rows * D + i;
    s += dy[j] * mask[j];
  }
  dx[i] = s * (1.0f / (1.0f - p));
}

// ---------------------------------------------------------------- NCHW -> NHWC
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* y, int B, int C, int HW) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // index in NHWC
  if (i >= static_cast<int64_t>(B) * C * HW) return;
  const int c = static_cast<int>(i % C);
  const int64_t t = i / C;
  const int p = static_cast<int>(t % HW);
  const int b = static_cast<int>(t / HW);
  y[i] = x[(static_cast<int64_t>(b) * C + c) * HW + p];
}

}  // namespace
}  // namespace mvae

using namespace mvae;

#define ST(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" int mvae_bn_stats(const float* x, int64_t ldx, int S, int seg_rows, int C, double* acc, void* stream) {
  if (!x || !acc || S < 1 || S > kMaxSeg || seg_rows < 1 || C < 1) return set_error(MVAE_ERR_BAD_ARG, "bn_stats: bad args");
  const int chunks = (seg_rows + kRowsPerBlock - 1) / kRowsPerBlock;
  MVAE_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * S * C, ST(stream)));
  dim3 grid(S * chunks, (C + 31) / 32);
  bn_stats_kernel<<<grid, 256, 0, ST(stream)>>>(x, ldx, seg_rows, C, chunks, acc);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_bn_finalize(const double* acc, int S, int seg_rows, int C, float eps, float momentum, float* mean,
                                float* invstd, float* running_mean, float* running_var, const int32_t* update_order,
                                int n_updates, void* stream) {
  if (!acc || !mean || !invstd || S < 1 || S > kMaxSeg || n_updates < 0 || n_updates > kMaxSeg)
    return set_error(MVAE_ERR_BAD_ARG, "bn_finalize: bad args");
  BnOrder o;
  o.n = running_mean ? n_updates : 0;
  for (int i = 0; i < o.n; ++i) {
    if (update_order[i] < 0 || update_order[i] >= S) return set_error(MVAE_ERR_BAD_ARG, "bn_finalize: bad update order");
    o.seg[i] = update_order[i];
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(acc, S, C, seg_rows, eps, momentum, mean, invstd, running_mean,
                                                              running_var, o);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_bn_eval_stats(const float* running_mean, const float* running_var, int S, int C, float eps, float* mean,
                                  float* invstd, void* stream) {
  if (!running_mean || !running_var || !mean || !invstd || S < 1 || S > kMaxSeg) return set_error(MVAE_ERR_BAD_ARG, "bn_eval_stats: bad args");
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(running_mean, running_var, S, C, eps, mean, invstd);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_bn_apply(const float* x, int64_t ldx, float* h, int64_t ldh, int rows, int seg_rows, int C,
                             const float* mean, const float* invstd, const float* gamma, const float* beta, int swish_act,
                             void* stream) {
  if (!x || !h || !mean || !invstd || !gamma || !beta || rows < 1 || (C & 3) || (ldx & 3) || (ldh & 3))
    return set_error(MVAE_ERR_BAD_ARG, "bn_apply: bad args (C, ld multiples of 4)");
  const int64_t n = static_cast<int64_t>(rows) * (C / 4);
  bn_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(x, ldx, h, ldh, rows, seg_rows, C / 4, mean,
                                                                                  invstd, gamma, beta, swish_act);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_bn_bwd(const float* x, int64_t ldx, const float* dh, int64_t lddh, float* dx, int64_t lddx, int S,
                           int seg_rows, int C, int seg0, int nseg, const float* mean, const float* invstd,
                           const float* gamma, const float* beta, int swish_act, int batch_stats, double* acc2,
                           float* dgamma, float* dbeta, void* stream) {
  if (!x || !dh || !dx || !acc2 || !dgamma || !dbeta || S < 1 || S > kMaxSeg || seg0 < 0 || nseg < 1 || seg0 + nseg > S ||
      (C & 3) || (ldx & 3) || (lddh & 3) || (lddx & 3))
    return set_error(MVAE_ERR_BAD_ARG, "bn_bwd: bad args");
  const int chunks = (seg_rows + kRowsPerBlock - 1) / kRowsPerBlock;
  MVAE_CUDA_CHECK(cudaMemsetAsync(acc2, 0, sizeof(double) * 2 * S * C, ST(stream)));
  dim3 grid(nseg * chunks, (C + 31) / 32);
  bn_bwd_reduce_kernel<<<grid, 256, 0, ST(stream)>>>(x, ldx, dh, lddh, seg_rows, C, chunks, seg0, mean, invstd, gamma, beta,
                                                     swish_act, acc2);
  bn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(acc2, C, seg0, nseg, dgamma, dbeta);
  const int rows = nseg * seg_rows;
  const int64_t n = static_cast<int64_t>(rows) * (C / 4);
  bn_bwd_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(
      x, ldx, dh, lddh, dx, lddx, seg0 * seg_rows, rows, seg_rows, C / 4, mean, invstd, gamma, beta, swish_act, acc2,
      batch_stats);
  count_launch(3);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_dropout_fwd(const float* x, int x_rows, float* y, float* mask_out, const float* mask_in, int copies,
                                int D, float p, uint64_t seed, const int32_t* step_dev, void* stream) {
  if (!x || !y || x_rows < 1 || copies < 1 || D < 1 || p < 0.f || p >= 1.f) return set_error(MVAE_ERR_BAD_ARG, "dropout_fwd: bad args");
  const int64_t n = static_cast<int64_t>(copies) * x_rows * D;
  dropout_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(x, x_rows, y, mask_out, mask_in, n, D, p,
                                                                                 static_cast<uint32_t>(seed), step_dev);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_dropout_bwd(const float* dy, const float* mask, float* dx, int x_rows, int copies, int D, float p,
                                void* stream) {
  if (!dy || !mask || !dx || x_rows < 1 || copies < 1) return set_error(MVAE_ERR_BAD_ARG, "dropout_bwd: bad args");
  const int64_t n = static_cast<int64_t>(x_rows) * D;
  dropout_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(dy, mask, dx, x_rows, copies, D, p);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_nchw_to_nhwc(const float* x, float* y, int B, int C, int HW, void* stream) {
  if (!x || !y || B < 1 || C < 1 || HW < 1) return set_error(MVAE_ERR_BAD_ARG, "nchw_to_nhwc: bad args");
  const int64_t n = static_cast<int64_t>(B) * C * HW;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(x, y, B, C, HW);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
