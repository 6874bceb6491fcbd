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

constexpr int kMaxSeg = 32;

// MUFU sigmoid (ex2.approx + rcp.approx): the IEEE expf + division version made these kernels instruction-bound.
__device__ __forceinline__ float sigmoid_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// Rows of a segment handled by one block of the column reductions.  The first version used 128 rows: one activation
// tensor of the CelebA decoder ([3B*32*32, 32] at B = 1024: 3.1 M rows) then became 24 K blocks x 64 double atomics on
// 192 addresses.  Sized here so that the launch has ~6 blocks per SM, each streaming >= 256 rows with 128-bit loads.
int rows_per_block(int S, int seg_rows, int C) {
  const int col_groups = (C + 31) / 32;
  int sms = mvae_device_sm_count();
  if (sms <= 0) sms = 148;
  const int64_t target = 6ll * sms;                                  // blocks in the whole launch
  int64_t chunks = target / (static_cast<int64_t>(S) * col_groups);  // chunks per segment
  if (chunks < 1) chunks = 1;
  int64_t rows = (seg_rows + chunks - 1) / chunks;
  if (rows < 256) rows = 256;
  rows = (rows + 31) / 32 * 32;
  return static_cast<int>(rows);
}

// Block = 8 column groups (float4) x 32 row lanes over a 32-column slab.  Shared reduction of two partial sums per
// column over the 32 row lanes; thread c < 32 issues the two double atomics of its column.
__device__ __forceinline__ void block_colsum2(float (&a)[4], float (&b)[4], int cg, int rl, int c0, int C, double* acc_seg) {
  __shared__ float s1[32][33], s2[32][33];
#pragma unroll
  for (int q = 0; q < 4; ++q) { s1[rl][cg * 4 + q] = a[q]; s2[rl][cg * 4 + q] = b[q]; }
  __syncthreads();
  if (threadIdx.x < 32 && c0 + threadIdx.x < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) { ta += s1[i][threadIdx.x]; tb += s2[i][threadIdx.x]; }
    atomicAdd(acc_seg + (c0 + threadIdx.x) * 2, ta);
    atomicAdd(acc_seg + (c0 + threadIdx.x) * 2 + 1, tb);
  }
}

// grid = (S * chunks_per_seg, ceil(C/32)); block = 8 float4 column groups x 32 row lanes (C % 4 == 0)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int64_t ld, int seg_rows, int C,
                                                       int chunks_per_seg, int rows_per_chunk,
                                                       double* __restrict__ acc /*[S][C][2]*/) {
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * 32, c = c0 + cg * 4;
  const int seg = blockIdx.x / chunks_per_seg, chunk = blockIdx.x % chunks_per_seg;
  const int r0 = chunk * rows_per_chunk, r1 = min(seg_rows, r0 + rows_per_chunk);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const float* base = x + (static_cast<int64_t>(seg) * seg_rows) * ld + c;
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + static_cast<int64_t>(r) * ld));
      a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
      b[0] += v.x * v.x; b[1] += v.y * v.y; b[2] += v.z * v.z; b[3] += v.w * v.w;
    }
  }
  block_colsum2(a, b, cg, rl, c0, C, acc + static_cast<int64_t>(seg) * C * 2);
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
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (host checks the range)
  if (gid >= static_cast<unsigned>(rows) * static_cast<unsigned>(C4)) return;
  const int r = static_cast<int>(gid / static_cast<unsigned>(C4));
  const int c = static_cast<int>(gid - static_cast<unsigned>(r) * static_cast<unsigned>(C4)) * 4;
  const int s = static_cast<int>(static_cast<unsigned>(r) / static_cast<unsigned>(seg_rows));
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
                                                            int chunks_per_seg, int rows_per_chunk, int seg0,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int act, double* __restrict__ acc2 /*[S][C][2]*/) {
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * 32, c = c0 + cg * 4;
  const int seg = seg0 + blockIdx.x / chunks_per_seg, chunk = blockIdx.x % chunks_per_seg;
  const int r0 = chunk * rows_per_chunk, r1 = min(seg_rows, r0 + rows_per_chunk);
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const float4 m4 = *reinterpret_cast<const float4*>(mean + seg * C + c);
    const float4 i4 = *reinterpret_cast<const float4*>(invstd + seg * C + c);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b4 = *reinterpret_cast<const float4*>(beta + c);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, is[4] = {i4.x, i4.y, i4.z, i4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
    const int64_t row0 = static_cast<int64_t>(seg) * seg_rows;
#pragma unroll 2
    for (int r = r0 + rl; r < r1; r += 32) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * ldx + c));
      const float4 dv = __ldg(reinterpret_cast<const float4*>(dh + (row0 + r) * lddh + c));
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float d[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float xh = (xs[q] - m[q]) * is[q];
        if (act) {
          const float a = g[q] * xh + bt[q];
          const float sg = sigmoid_f(a);
          d[q] *= sg * (1.0f + a * (1.0f - sg));
        }
        a1[q] += d[q]; a2[q] += d[q] * xh;
      }
    }
  }
  block_colsum2(a1, a2, cg, rl, c0, C, acc2 + static_cast<int64_t>(seg) * C * 2);
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
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (host checks the range)
  if (gid >= static_cast<unsigned>(rows) * static_cast<unsigned>(C4)) return;
  const unsigned rq = gid / static_cast<unsigned>(C4);
  const int r = row_begin + static_cast<int>(rq);
  const int c = static_cast<int>(gid - rq * static_cast<unsigned>(C4)) * 4;
  const int s = static_cast<int>(static_cast<unsigned>(r) / static_cast<unsigned>(seg_rows));
  const int C = C4 * 4;
  const float inv_n = 1.0f / static_cast<float>(seg_rows);
  const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * ldx + c);
  const float4 dv = *reinterpret_cast<const float4*>(dh + static_cast<int64_t>(r) * lddh + c);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
  const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
  const float4 m4 = *reinterpret_cast<const float4*>(mean + s * C + c);
  const float4 i4 = *reinterpret_cast<const float4*>(invstd + s * C + c);
  const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
  const float4 b4 = *reinterpret_cast<const float4*>(beta + c);
  const float ms[4] = {m4.x, m4.y, m4.z, m4.w}, iss[4] = {i4.x, i4.y, i4.z, i4.w};
  const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
  float out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float m = ms[q], is = iss[q], g = gs[q], b = bs[q];
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

// ---------------------------------------------------------------- device-resident dataset -> batch
// out[b, :] = data[idx[b], :] / 255 (torchvision ToTensor of a uint8 image, mnist/train.py:160), labels_out[b] =
// labels[idx[b]]: one thread per 4 bytes of a row (uchar4 -> float4)
__global__ void __launch_bounds__(256) gather_batch_u8_kernel(const uint8_t* __restrict__ data, int64_t row_bytes,
                                                              const int64_t* __restrict__ labels,
                                                              const int64_t* __restrict__ idx, int B, float* out,
                                                              int64_t ld_out, int64_t* labels_out) {
  const unsigned q4 = static_cast<unsigned>(row_bytes >> 2);
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= static_cast<unsigned>(B) * q4) return;
  const unsigned b = gid / q4, j = gid - b * q4;
  const int64_t row = idx[b];
  const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(data + row * row_bytes) + j);
  *reinterpret_cast<float4*>(out + static_cast<int64_t>(b) * ld_out + 4 * j) =
      make_float4(static_cast<float>(v.x) / 255.0f, static_cast<float>(v.y) / 255.0f, static_cast<float>(v.z) / 255.0f,
                  static_cast<float>(v.w) / 255.0f);
  if (j == 0 && labels_out != nullptr) labels_out[b] = labels[row];
}

}  // namespace
}  // namespace mvae

using namespace mvae;

#define ST(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" int mvae_bn_stats(const float* x, int64_t ldx, int S, int seg_rows, int C, double* acc, void* stream) {
  if (!x || !acc || S < 1 || S > kMaxSeg || seg_rows < 1 || C < 1 || (C & 3) || (ldx & 3) ||
      (reinterpret_cast<uintptr_t>(x) & 15))
    return set_error(MVAE_ERR_BAD_ARG, "bn_stats: bad args (C, ldx multiples of 4, x 16-byte aligned)");
  const int rpb = rows_per_block(S, seg_rows, C);
  const int chunks = (seg_rows + rpb - 1) / rpb;
  MVAE_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * S * C, ST(stream)));
  dim3 grid(S * chunks, (C + 31) / 32);
  bn_stats_kernel<<<grid, 256, 0, ST(stream)>>>(x, ldx, seg_rows, C, chunks, rpb, acc);
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
  const int rpb = rows_per_block(nseg, seg_rows, C);
  const int chunks = (seg_rows + rpb - 1) / rpb;
  MVAE_CUDA_CHECK(cudaMemsetAsync(acc2, 0, sizeof(double) * 2 * S * C, ST(stream)));
  dim3 grid(nseg * chunks, (C + 31) / 32);
  bn_bwd_reduce_kernel<<<grid, 256, 0, ST(stream)>>>(x, ldx, dh, lddh, seg_rows, C, chunks, rpb, seg0, mean, invstd, gamma,
                                                     beta, swish_act, acc2);
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

extern "C" int mvae_gather_batch_u8(const uint8_t* data, int64_t row_bytes, const int64_t* labels, const int64_t* idx, int B,
                                    float* out, int64_t ld_out, int64_t* labels_out, void* stream) {
  if (!data || !idx || !out || B < 1 || row_bytes < 4 || (row_bytes & 3) || (ld_out & 3) || ld_out < row_bytes ||
      (labels_out != nullptr && labels == nullptr) ||
      ((reinterpret_cast<uintptr_t>(data) & 3) | (reinterpret_cast<uintptr_t>(out) & 15)) != 0)
    return set_error(MVAE_ERR_BAD_ARG, "gather_batch_u8: bad args (row_bytes %% 4 == 0, out 16-byte aligned, ld_out %% 4 == 0)");
  const int64_t n = static_cast<int64_t>(B) * (row_bytes / 4);
  if (n >= (1ll << 32) - 256) return set_error(MVAE_ERR_UNSUPPORTED, "gather_batch_u8: batch too large for one call");
  gather_batch_u8_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(data, row_bytes, labels, idx, B, out,
                                                                                         ld_out, labels_out);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
