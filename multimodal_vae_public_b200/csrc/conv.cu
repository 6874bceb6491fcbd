// 4x4 / stride 2 / pad 1 convolution plumbing for the conv MVAE flavours (fashionmnist/model.py:79-82,112-114):
// NHWC activations, the contraction itself runs on the tcgen05 GEMM (gemm_tcgen05.cu); these two bandwidth-bound
// kernels move data between image layout and GEMM-operand layout.
//
//   Conv2d(Cin->Cout,4,2,1)           y = im2col(x) * W^T            cols [B*OH*OW, 16*Cin]  (k = (kh*4+kw)*Cin + ci)
//   ConvTranspose2d(Cin->Cout,4,2,1)  cols = x * Wt^T ; y = col2im(cols)   cols [B*IH*IW, 16*Cout]
//   and their adjoints: d/dx Conv = col2im(dcols), d/dcols ConvT = im2col(dy).
// col2im is written as a GATHER (each output pixel sums its <= 4 contributing taps): no atomics, deterministic.
// Optional fusions: Swish of the gathered sum (decoder forward) or multiplication by Swish'(aux) (encoder backward).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

// MUFU sigmoid (ex2.approx + rcp.approx, relative error ~(2 + |x|) * 2^-23): the IEEE expf + division version made
// these bandwidth-bound kernels instruction-bound (same finding as the BCE kernel, profiles/r01_notes.md).
__device__ __forceinline__ float sigmoid_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// one thread = one float4 of channels (or one scalar when C < 4) of one (output pixel, tap)
template <int VEC>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, float* __restrict__ cols, int B, int H,
                                                     int W, int C, int64_t ld_cols, int stride, int pad) {
  // 32-bit index arithmetic (the host checks that the element count fits): the original 64-bit div/mod chain cost more
  // instructions than the 16-byte copy it addresses and held the kernel at ~1.6 TB/s
  const unsigned OH = (H + 2 * pad - 4) / stride + 1, OW = (W + 2 * pad - 4) / stride + 1;
  const unsigned cv = C / VEC;
  const unsigned total = static_cast<unsigned>(B) * OH * OW * 16u * cv;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  unsigned r = gid / cv;
  const int c = static_cast<int>(gid - r * cv) * VEC;
  const int tap = static_cast<int>(r & 15u);
  r >>= 4;                       // output pixel index m = (b*OH + oh)*OW + ow
  const unsigned r2 = r / OW;
  const int ow = static_cast<int>(r - r2 * OW);
  const unsigned bq = r2 / OH;
  const int oh = static_cast<int>(r2 - bq * OH);
  const int b = static_cast<int>(bq);
  const int kh = tap >> 2, kw = tap & 3;
  const int ih = stride * oh - pad + kh, iw = stride * ow - pad + kw;
  float* dst = cols + static_cast<int64_t>(r) * ld_cols + tap * C + c;
  const bool in = (ih >= 0 && ih < H && iw >= 0 && iw < W);
  const float* src = x + ((static_cast<int64_t>(b) * H + ih) * W + iw) * C + c;
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(dst) = in ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    *dst = in ? __ldg(src) : 0.f;
  }
}

// Channel counts that are not a multiple of 4 (the RGB layers, C = 3): one thread still WRITES one float4 of the
// [pixels, 16*C] matrix (rows are 16*C floats, always a multiple of 4) and gathers its four scalars; the scalar-store
// version ran at 0.7 TB/s.
__global__ void __launch_bounds__(256) im2col_anyc_kernel(const float* __restrict__ x, float* __restrict__ cols, int B, int H,
                                                          int W, int C, int64_t ld_cols, int stride, int pad) {
  const unsigned OH = (H + 2 * pad - 4) / stride + 1, OW = (W + 2 * pad - 4) / stride + 1;
  const unsigned rv = 4u * C;                          // float4 per row
  const unsigned total = static_cast<unsigned>(B) * OH * OW * rv;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const unsigned r = gid / rv;                         // output pixel index m = (b*OH + oh)*OW + ow
  const unsigned j = gid - r * rv;
  const unsigned r2 = r / OW;
  const int ow = static_cast<int>(r - r2 * OW);
  const unsigned bq = r2 / OH;
  const int oh = static_cast<int>(r2 - bq * OH);
  const float* xb = x + static_cast<int64_t>(bq) * H * W * C;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const unsigned e = 4u * j + q;
    const unsigned tap = e / static_cast<unsigned>(C);
    const int c = static_cast<int>(e - tap * C);
    const int ih = stride * oh - pad + static_cast<int>(tap >> 2), iw = stride * ow - pad + static_cast<int>(tap & 3);
    v[q] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + (static_cast<int64_t>(ih) * W + iw) * C + c) : 0.f;
  }
  *reinterpret_cast<float4*>(cols + static_cast<int64_t>(r) * ld_cols + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
}

// out[b, oh, ow, c] = sum over taps (kh,kw) with oh = 2*ih - 1 + kh, ow = 2*iw - 1 + kw of cols[(b,ih,iw), (kh,kw,c)]
template <int VEC>
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ cols, int64_t ld_cols, float* out,
                                                     float* out_act, const float* __restrict__ aux, int B, int IH,
                                                     int IW, int C, int stride, int pad) {
  const unsigned OH = (IH - 1) * stride - 2 * pad + 4, OW = (IW - 1) * stride - 2 * pad + 4;
  const unsigned cv = C / VEC;
  const unsigned total = static_cast<unsigned>(B) * OH * OW * cv;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const unsigned r = gid / cv;   // output pixel
  const int c = static_cast<int>(gid - r * cv) * VEC;
  const unsigned r2 = r / OW;
  const int ow = static_cast<int>(r - r2 * OW);
  const unsigned bq = r2 / OH;
  const int oh = static_cast<int>(r2 - bq * OH);
  const int b = static_cast<int>(bq);
  float acc[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int th = oh + pad - kh;               // = stride * ih
    if (th < 0 || (stride == 2 ? (th & 1) : (th % stride)) != 0) continue;
    const int ih = stride == 2 ? (th >> 1) : (th / stride);
    if (ih >= IH) continue;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int tw = ow + pad - kw;
      if (tw < 0 || (stride == 2 ? (tw & 1) : (tw % stride)) != 0) continue;
      const int iw = stride == 2 ? (tw >> 1) : (tw / stride);
      if (iw >= IW) continue;
      const float* src = cols + ((static_cast<int64_t>(b) * IH + ih) * IW + iw) * ld_cols + (kh * 4 + kw) * C + c;
      if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      } else {
        acc[0] += __ldg(src);
      }
    }
  }
  const int64_t o = static_cast<int64_t>(r) * C + c;
  if (aux != nullptr) {  // backward through the Swish that produced this layer's input: multiply by Swish'(aux)
    float a[VEC];
    if constexpr (VEC == 4) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(aux + o));
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    } else {
      a[0] = aux[o];
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float s = sigmoid_f(a[q]);
      acc[q] *= s * (1.0f + a[q] * (1.0f - s));
    }
  }
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (out_act != nullptr)
      *reinterpret_cast<float4*>(out_act + o) = make_float4(acc[0] * sigmoid_f(acc[0]), acc[1] * sigmoid_f(acc[1]),
                                                            acc[2] * sigmoid_f(acc[2]), acc[3] * sigmoid_f(acc[3]));
  } else {
    out[o] = acc[0];
    if (out_act != nullptr) out_act[o] = acc[0] * sigmoid_f(acc[0]);
  }
}

// C <= 4 and not a multiple of 4 (RGB): one thread = one output PIXEL (all its channels), instead of one thread per
// scalar -- a third of the threads and of the index arithmetic, adjacent loads per tap.
__global__ void __launch_bounds__(256) col2im_smallc_kernel(const float* __restrict__ cols, int64_t ld_cols, float* out,
                                                            float* out_act, const float* __restrict__ aux, int B, int IH,
                                                            int IW, int C, int stride, int pad) {
  const unsigned OH = (IH - 1) * stride - 2 * pad + 4, OW = (IW - 1) * stride - 2 * pad + 4;
  const unsigned total = static_cast<unsigned>(B) * OH * OW;
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;   // output pixel
  if (r >= total) return;
  const unsigned r2 = r / OW;
  const int ow = static_cast<int>(r - r2 * OW);
  const unsigned bq = r2 / OH;
  const int oh = static_cast<int>(r2 - bq * OH);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int th = oh + pad - kh;
    if (th < 0 || (stride == 2 ? (th & 1) : (th % stride)) != 0) continue;
    const int ih = stride == 2 ? (th >> 1) : (th / stride);
    if (ih >= IH) continue;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int tw = ow + pad - kw;
      if (tw < 0 || (stride == 2 ? (tw & 1) : (tw % stride)) != 0) continue;
      const int iw = stride == 2 ? (tw >> 1) : (tw / stride);
      if (iw >= IW) continue;
      const float* src = cols + ((static_cast<int64_t>(bq) * IH + ih) * IW + iw) * ld_cols + (kh * 4 + kw) * C;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) acc[c] += __ldg(src + c);
    }
  }
  const int64_t o = static_cast<int64_t>(r) * C;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= C) break;
    float v = acc[c];
    if (aux != nullptr) {
      const float a = aux[o + c];
      const float sg = sigmoid_f(a);
      v *= sg * (1.0f + a * (1.0f - sg));
    }
    out[o + c] = v;
    if (out_act != nullptr) out_act[o + c] = v * sigmoid_f(v);
  }
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_im2col_k4(const float* x, float* cols, int64_t ld_cols, int B, int H, int W, int C, int stride,
                              int pad, void* stream) {
  if (!x || !cols || B < 1 || H < 1 || W < 1 || C < 1 || ld_cols < 16 * C || stride < 1 || pad < 0 ||
      H + 2 * pad < 4 || W + 2 * pad < 4)
    return set_error(MVAE_ERR_BAD_ARG, "im2col: bad arguments (ld_cols >= 16*C)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool v4 = (C % 4 == 0) && (ld_cols % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(cols)) & 15) == 0;
  const int OH = (H + 2 * pad - 4) / stride + 1, OW = (W + 2 * pad - 4) / stride + 1;
  const int64_t total = static_cast<int64_t>(B) * OH * OW * 16 * (v4 ? C / 4 : C);
  if (total >= (1ll << 32) - 256) return set_error(MVAE_ERR_UNSUPPORTED, "im2col: more than 2^32 elements in one call");
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  const bool out4 = (ld_cols % 4 == 0) && (reinterpret_cast<uintptr_t>(cols) & 15) == 0;
  if (v4) im2col_kernel<4><<<blocks, 256, 0, st>>>(x, cols, B, H, W, C, ld_cols, stride, pad);
  else if (out4) im2col_anyc_kernel<<<static_cast<unsigned>((total / 4 + 255) / 256), 256, 0, st>>>(x, cols, B, H, W, C, ld_cols,
                                                                                                 stride, pad);
  else    im2col_kernel<1><<<blocks, 256, 0, st>>>(x, cols, B, H, W, C, ld_cols, stride, pad);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_im2col_k4s2p1(const float* x, float* cols, int64_t ld_cols, int B, int H, int W, int C,
                                  void* stream) {
  if ((H & 1) || (W & 1)) return set_error(MVAE_ERR_BAD_ARG, "im2col_k4s2p1: even H, W required");
  return mvae_im2col_k4(x, cols, ld_cols, B, H, W, C, 2, 1, stream);
}

extern "C" int mvae_col2im_k4(const float* cols, int64_t ld_cols, float* out, float* out_act, const float* aux, int B,
                              int IH, int IW, int C, int stride, int pad, void* stream) {
  if (!cols || !out || B < 1 || IH < 1 || IW < 1 || C < 1 || ld_cols < 16 * C || stride < 1 || pad < 0)
    return set_error(MVAE_ERR_BAD_ARG, "col2im: bad arguments (ld_cols >= 16*C)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool v4 = (C % 4 == 0) && (ld_cols % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(cols) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_act) |
                    reinterpret_cast<uintptr_t>(aux)) & 15) == 0;
  const int OH = (IH - 1) * stride - 2 * pad + 4, OW = (IW - 1) * stride - 2 * pad + 4;
  const int64_t total = static_cast<int64_t>(B) * OH * OW * (v4 ? C / 4 : C);
  if (total >= (1ll << 32) - 256) return set_error(MVAE_ERR_UNSUPPORTED, "col2im: more than 2^32 elements in one call");
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (v4) col2im_kernel<4><<<blocks, 256, 0, st>>>(cols, ld_cols, out, out_act, aux, B, IH, IW, C, stride, pad);
  else if (C <= 4)
    col2im_smallc_kernel<<<static_cast<unsigned>((total / C + 255) / 256), 256, 0, st>>>(cols, ld_cols, out, out_act, aux, B, IH,
                                                                                       IW, C, stride, pad);
  else    col2im_kernel<1><<<blocks, 256, 0, st>>>(cols, ld_cols, out, out_act, aux, B, IH, IW, C, stride, pad);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_col2im_k4s2p1(const float* cols, int64_t ld_cols, float* out, float* out_act, const float* aux,
                                  int B, int IH, int IW, int C, void* stream) {
  return mvae_col2im_k4(cols, ld_cols, out, out_act, aux, B, IH, IW, C, 2, 1, stream);
}
