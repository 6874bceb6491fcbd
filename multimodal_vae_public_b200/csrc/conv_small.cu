// Direct 4x4 / stride-2 / pad-1 convolutions whose image side has 1 or 3 channels (sm_100a, CUDA cores, HBM-bound):
//   Conv2d(CIN -> Cout)           first encoder layer   fashionmnist/model.py:79 (1 -> 64), celeba/model.py:77 (3 -> 32)
//   ConvTranspose2d(Cin -> COUT)  last decoder layer    fashionmnist/model.py:114 (64 -> 1), celeba/model.py:126 (32 -> 3)
// and their weight / data gradients.
//
// Why not the tensor-core GEMM: with K = 16 (or 48) the contraction is 1-2 k-blocks deep, the im2col / cols matrices are
// 16x the image, and the epilogue -- not the MMA -- sets the pace (round 1: 3-27 TFLOP/s on these layers; together
// 1.2 of the 7.2 ms FashionMNIST step, profiles/r02_times_fashion_b4096_v0.txt).  The work is 16 * CIN FMAs per output
// element: far below the FMA roofline, so these kernels stream the wide (Cout- or Cin-channel) activation exactly once
// with 128-bit accesses and keep everything else (image taps, weights, the per-pixel "cols" vector) in registers /
// shared memory.  No im2col, no cols, no col2im buffers.
//
// Layouts (as in conv.cu): activations NHWC; conv weight Wc [Cout][(kh,kw,ci)], transposed-conv weight Wt [(kh,kw,co)][Cin].
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

__device__ __forceinline__ float sigmoid_m(float x) {   // MUFU sigmoid: relative error ~(2 + |x|) * 2^-23 (as in conv.cu)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float swish_m(float x) { return x * sigmoid_m(x); }
__device__ __forceinline__ float dswish_m(float x) {
  const float s = sigmoid_m(x);
  return s * (1.0f + x * (1.0f - s));
}

// Stage the filter taps of `npx` consecutive output pixels [p0, p0 + npx) in shared memory: s_tap[pixel][K = 16 CIN], k =
// (kh, kw, ci), zero at the padding and beyond the last pixel.  4 tasks per pixel (one per kh): the (image, row, column)
// decomposition and the bounds tests are done ONCE per 4 CIN taps instead of once per multiply (the first version
// recomputed them per FMA operand and was instruction-bound: 155 M warp instructions for 411 MB of output).
template <int CIN>
__device__ __forceinline__ void stage_taps(const float* __restrict__ x, float* s_tap, int64_t p0, int npx, int64_t npix,
                                           int H, int W, int OH, int OW) {
  constexpr int K = 16 * CIN;
  for (int task = threadIdx.x; task < npx * 4; task += blockDim.x) {
    const int pi = task >> 2, kh = task & 3;
    const int64_t p = p0 + pi;
    float v[4 * CIN];
#pragma unroll
    for (int j = 0; j < 4 * CIN; ++j) v[j] = 0.f;
    if (p < npix) {
      const int b = static_cast<int>(p / (OH * OW));
      const int r = static_cast<int>(p - static_cast<int64_t>(b) * OH * OW);
      const int oh = r / OW, ow = r - oh * OW;
      const int iy = 2 * oh - 1 + kh;
      if (iy >= 0 && iy < H) {
        const float* row = x + (static_cast<int64_t>(b) * H + iy) * W * CIN;
        const int ix0 = 2 * ow - 1;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int ix = ix0 + kw;
          if (ix >= 0 && ix < W) {
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) v[kw * CIN + ci] = __ldg(row + ix * CIN + ci);
          }
        }
      }
    }
    float* dst = s_tap + pi * K + kh * 4 * CIN;
#pragma unroll
    for (int j = 0; j < 4 * CIN; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Conv2d(CIN -> Cout), forward: a[p][co] = sum_{kh,kw,ci} x[b, 2oh-1+kh, 2ow-1+kw, ci] * Wc[co][(kh,kw,ci)], h = swish(a)
// A block walks chunks of kPix * (256 / (Cout/4)) output pixels: taps staged in shared memory, then one thread = 4
// consecutive output channels of kPix pixels (every weight float4 is reused for kPix pixels; taps and weights come from
// shared memory as 128-bit broadcasts); the Cout/4 threads of a pixel are adjacent lanes, so the 128-bit stores cover
// whole rows of a / h.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPix = 4;
template <int CIN>
__global__ void __launch_bounds__(256) conv_cin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wc,
                                                           float* __restrict__ a, float* __restrict__ h, int B, int H, int W,
                                                           int Cout) {
  constexpr int K = 16 * CIN;
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                                  // [K][Cout] (transposed: the threads of a pixel read consecutive co)
  float* s_tap = smem + K * Cout;                     // [chunk pixels][K]
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int co = i / K, k = i - co * K;
    s_w[k * Cout + co] = wc[i];
  }
  const int OH = H >> 1, OW = W >> 1;
  const int cg = Cout >> 2;                           // threads per pixel
  const int ppb = blockDim.x / cg;                    // pixels per pass
  const int chunk = ppb * kPix;
  const int64_t npix = static_cast<int64_t>(B) * OH * OW;
  const int c4 = (threadIdx.x % cg) * 4;
  const int pl = threadIdx.x / cg;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * chunk; base < npix; base += static_cast<int64_t>(gridDim.x) * chunk) {
    __syncthreads();                                  // s_w ready / previous chunk's taps consumed
    stage_taps<CIN>(x, s_tap, base, chunk, npix, H, W, OH, OW);
    __syncthreads();
    float acc[kPix][4];
#pragma unroll
    for (int j = 0; j < kPix; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
#pragma unroll 4
    for (int k4 = 0; k4 < K / 4; ++k4) {
      float4 w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(s_w + (4 * k4 + i) * Cout + c4);
#pragma unroll
      for (int j = 0; j < kPix; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(s_tap + (pl + j * ppb) * K + 4 * k4);
        acc[j][0] += t.x * w[0].x + t.y * w[1].x + t.z * w[2].x + t.w * w[3].x;
        acc[j][1] += t.x * w[0].y + t.y * w[1].y + t.z * w[2].y + t.w * w[3].y;
        acc[j][2] += t.x * w[0].z + t.y * w[1].z + t.z * w[2].z + t.w * w[3].z;
        acc[j][3] += t.x * w[0].w + t.y * w[1].w + t.z * w[2].w + t.w * w[3].w;
      }
    }
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const int64_t p = base + pl + static_cast<int64_t>(j) * ppb;
      if (p >= npix) continue;
      __stcs(reinterpret_cast<float4*>(a + p * Cout + c4), make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
      __stcs(reinterpret_cast<float4*>(h + p * Cout + c4),
             make_float4(swish_m(acc[j][0]), swish_m(acc[j][1]), swish_m(acc[j][2]), swish_m(acc[j][3])));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Conv2d(CIN -> Cout), weight gradient: dWc[co][k] += sum_p da[p][co] * xcol[p][k]   (k = (kh,kw,ci); no data gradient:
// the image is an input).  A thread owns a 4 (co) x 4 (k) patch of dWc; the block walks a contiguous pixel range in chunks
// of kWgChunk pixels whose taps are staged in shared memory; the pixel lanes of the block take alternate pixels.  Per
// pixel a thread issues one 128-bit load of da (coalesced over the co threads, read exactly once), one 128-bit shared
// load of taps and 16 FMAs.  Partial patches are summed over the lanes in shared memory and added to dWc with one
// red.add.v4 per patch row per block.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWgChunk = 256;
template <int CIN>
__global__ void __launch_bounds__(256) conv_cin_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ da,
                                                             float* __restrict__ dwc, int B, int H, int W, int Cout,
                                                             int pix_per_block) {
  constexpr int K = 16 * CIN, KG = K / 4;            // k groups of 4 consecutive k
  extern __shared__ __align__(16) float smem[];
  float* s_tap = smem;                                // [kWgChunk][K]
  float* s_red = smem + kWgChunk * K;                 // [lanes][Cout * K] partial patches
  const int OH = H >> 1, OW = W >> 1;
  const int cg = Cout >> 2;
  const int tpl = cg * KG;                           // threads per pixel lane
  const int lanes = blockDim.x / tpl;
  const int lane_id = threadIdx.x / tpl, t = threadIdx.x - lane_id * tpl;
  const bool active = lane_id < lanes;
  const int c4 = (t % cg) * 4, kq = t / cg;          // my patch: co [c4, c4+4) x k [4 kq, 4 kq + 4)
  const int64_t npix = static_cast<int64_t>(B) * OH * OW;
  const int64_t p_begin = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p_end = p_begin + pix_per_block < npix ? p_begin + pix_per_block : npix;
  float acc[4][4];                                   // [k][co]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
  for (int64_t c0 = p_begin; c0 < p_end; c0 += kWgChunk) {
    const int n = static_cast<int>(p_end - c0 < kWgChunk ? p_end - c0 : kWgChunk);
    __syncthreads();
    stage_taps<CIN>(x, s_tap, c0, n, npix, H, W, OH, OW);
    __syncthreads();
    if (active) {
#pragma unroll 8
      for (int pi = lane_id; pi < n; pi += lanes) {
        const float4 d4 = __ldcs(reinterpret_cast<const float4*>(da + (c0 + pi) * Cout + c4));
        const float4 x4 = *reinterpret_cast<const float4*>(s_tap + pi * K + 4 * kq);
        acc[0][0] += x4.x * d4.x; acc[0][1] += x4.x * d4.y; acc[0][2] += x4.x * d4.z; acc[0][3] += x4.x * d4.w;
        acc[1][0] += x4.y * d4.x; acc[1][1] += x4.y * d4.y; acc[1][2] += x4.y * d4.z; acc[1][3] += x4.y * d4.w;
        acc[2][0] += x4.z * d4.x; acc[2][1] += x4.z * d4.y; acc[2][2] += x4.z * d4.z; acc[2][3] += x4.z * d4.w;
        acc[3][0] += x4.w * d4.x; acc[3][1] += x4.w * d4.y; acc[3][2] += x4.w * d4.z; acc[3][3] += x4.w * d4.w;
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)   // s_red[lane][co][k]
#pragma unroll
      for (int q = 0; q < 4; ++q) s_red[(lane_id * Cout + c4 + q) * K + 4 * kq + i] = acc[i][q];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < Cout * KG; idx += blockDim.x) {   // one float4 of dWc[co][4 kq ..]
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < lanes; ++l) {
      const float4 v = *reinterpret_cast<const float4*>(s_red + static_cast<size_t>(l) * Cout * K + idx * 4);
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dwc + idx * 4), "f"(s4.x), "f"(s4.y), "f"(s4.z), "f"(s4.w)
                 : "memory");
  }
}

// Cooperative, fully coalesced copy of `rows` activation rows of Cin floats (global, contiguous) into shared memory with
// a row stride of Cin + 4 floats (conflict-free 128-bit row reads by one thread per row), up to 8 loads per thread in
// flight.  rows_valid < rows: the remaining rows are zero-filled.
__device__ __forceinline__ void stage_rows_padded(const float* __restrict__ src, float* dst, int rows, int rows_valid, int Cin,
                                                  int ld) {
  const int c4n = Cin >> 2, ld4 = ld >> 2;
  const int total = rows * c4n;
  for (int i0 = 0; i0 < total; i0 += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x + threadIdx.x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total && i / c4n < rows_valid) v[u] = __ldcs(reinterpret_cast<const float4*>(src) + i);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x + threadIdx.x;
      if (i < total) {
        const int r = i / c4n, c = i - r * c4n;
        reinterpret_cast<float4*>(dst)[r * ld4 + c] = v[u];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// ConvTranspose2d(Cin -> COUT), forward: out[b, oy, ox, co] = sum over the <= 2 x 2 taps (iy, kh), (ix, kw) with
// oy = 2 iy - 1 + kh, ox = 2 ix - 1 + kw of  sum_ci hin[b, iy, ix, ci] * Wt[(kh,kw,co)][ci].
// A block takes `tr` input rows of one image (+ one halo row on each side): phase 1, one thread per input pixel, the
// pixel's Cin channels streamed once (128-bit) against the 16 * COUT weight rows held in shared memory -> the pixel's
// "cols" vector in shared memory; phase 2 gathers the <= 4 contributions of every output pixel of the tile and writes
// whole output rows.  The wide activation is read once (the halo rows twice); nothing else touches HBM.
// (Measured, FashionMNIST B = 4096: 235 us, L1-bound (every lane streams its own 256-byte row).  A variant that staged the
// rows coalesced through shared memory and register-blocked over pixel pairs issued 45 % more instructions and took
// 304 us; the im2col-free tensor-core route it replaces took 322 us.  profiles/r02_conv_small_ncu.txt)
// ------------------------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256) convT_cout_fwd_kernel(const float* __restrict__ hin, const float* __restrict__ wt,
                                                             float* __restrict__ out, int B, int IH, int IW, int Cin,
                                                             int tr, int tiles_per_img) {
  constexpr int KC = 16 * COUT;
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                                  // [KC][Cin]
  float* s_cols = smem + KC * Cin;                    // [(tr + 2) * IW][KC + 1]  (+1: conflict-free column gathers)
  for (int i = threadIdx.x; i < KC * Cin / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
  const int c4n = Cin >> 2;
  const int OW = 2 * IW;
  for (int tile = blockIdx.x; tile < B * tiles_per_img; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int row0 = (tile - b * tiles_per_img) * tr;          // first owned input row
    const int rows = min(tr, IH - row0);
    const int lo = row0 - 1;                                    // first staged input row (may be -1: zero)
    const int npx = (rows + 2) * IW;
    __syncthreads();                                            // s_w ready / previous tile's s_cols consumed
    // ---- phase 1: cols of the staged pixels
    for (int px = threadIdx.x; px < npx; px += blockDim.x) {
      const int ry = px / IW, ix = px - ry * IW;
      const int iy = lo + ry;
      float cols[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) cols[k] = 0.f;
      if (iy >= 0 && iy < IH) {
        const float4* src = reinterpret_cast<const float4*>(hin + ((static_cast<int64_t>(b) * IH + iy) * IW + ix) * Cin);
        for (int c = 0; c < c4n; c += 4) {                     // 4 x 128-bit loads in flight
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = (c + u < c4n) ? __ldcs(src + c + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c + u < c4n) {
#pragma unroll
              for (int k = 0; k < KC; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * Cin + (c + u) * 4);   // warp-uniform: broadcast
                cols[k] += v[u].x * w4.x + v[u].y * w4.y + v[u].z * w4.z + v[u].w * w4.w;
              }
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) s_cols[px * (KC + 1) + k] = cols[k];
    }
    __syncthreads();
    // ---- phase 2: output rows [2 row0, 2 (row0 + rows)) of image b
    const int nout = 2 * rows * OW * COUT;
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
      const int oyl = o / (OW * COUT);
      const int rem = o - oyl * (OW * COUT);
      const int ox = rem / COUT, co = rem - ox * COUT;
      const int oy = 2 * row0 + oyl;
      float s = 0.f;
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const int kh = ((oy + 1) & 1) + 2 * a2;                 // kh = (oy + 1) mod 2, + 2
        const int iy = (oy + 1 - kh) >> 1;                      // exact: oy + 1 - kh is even
        if (iy < 0 || iy >= IH) continue;
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const int kw = ((ox + 1) & 1) + 2 * b2;
          const int ix = (ox + 1 - kw) >> 1;
          if (ix < 0 || ix >= IW) continue;
          s += s_cols[((iy - lo) * IW + ix) * (KC + 1) + (kh * 4 + kw) * COUT + co];
        }
      }
      out[((static_cast<int64_t>(b) * 2 * IH + oy) * OW) * COUT + rem] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// ConvTranspose2d(Cin -> COUT), backward (given dout = d loss / d out, [B, 2IH, 2IW, COUT]):
//   dcols[p][(kh,kw,co)] = dout[b, 2iy-1+kh, 2ix-1+kw, co]               (a gather: never materialised in HBM)
//   d hin[p][ci] = (sum_k dcols[p][k] Wt[k][ci]) * swish'(ain[p][ci])    (ain = pre-activation of the layer below)
//   dWt[k][ci]  += sum_p dcols[p][k] * hin[p][ci]
// Persistent blocks walk tiles of `tr` input rows.  The tile's ain rows are copied to shared memory coalesced; phase 1 (one
// thread per (pixel, half of the channels)) gathers dcols, overwrites the staged ain rows with d hin IN PLACE, and the
// block then streams them out coalesced; phase 2 (a thread owns a 4 (k) x 4 (ci) patch of dWt, pixel lanes in parallel,
// hin read coalesced from global memory exactly once) accumulates the weight gradient in REGISTERS across all tiles of
// the block; one red.add.v4 per patch row per block at the end.
// ------------------------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256, COUT == 1 ? 3 : 2) convT_cout_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ hin,
                                                             const float* __restrict__ ain, const float* __restrict__ wt,
                                                             float* __restrict__ dhin, float* __restrict__ dwt, int B, int IH,
                                                             int IW, int Cin, int tr, int tiles_per_img) {
  constexpr int KC = 16 * COUT, KG = KC / 4;
  extern __shared__ __align__(16) float smem[];
  const int c4n = Cin >> 2, ldh = Cin + 8;            // (+8: two threads per row, alternate float4s: conflict-free)
  const int npx_max = tr * IW;
  const int OW = 2 * IW, OH = 2 * IH;
  float* s_w = smem;                                  // [KC][Cin]
  float* s_dc = s_w + KC * Cin;                       // [KC][npx_max]  dcols of the tile, k-major
  float* s_do = s_dc + KC * npx_max;                  // [(2 tr + 2) * OW * COUT]  dout rows 2 row0 - 1 .. 2 (row0 + rows)
  float* s_a = s_do + (((2 * tr + 2) * OW * COUT + 3) & ~3);   // [npx_max][Cin + 4]  ain rows -> d hin rows (in place)
  float* s_red = s_a;                                 // (reused at the very end) [lanes][KC * Cin]
  for (int i = threadIdx.x; i < KC * Cin / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
  const int tpl = c4n * KG;                           // threads per pixel lane in phase 2
  const int lanes = blockDim.x / tpl;
  const int lane_id = threadIdx.x / tpl, t2 = threadIdx.x - lane_id * tpl;
  const int ci4 = (t2 % c4n) * 4, kq = t2 / c4n;      // my dWt patch: k [4 kq, +4) x ci [ci4, +4)
  float wacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) wacc[i][q] = 0.f;
  for (int tile = blockIdx.x; tile < B * tiles_per_img; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int row0 = (tile - b * tiles_per_img) * tr;
    const int rows = min(tr, IH - row0);
    const int npx = rows * IW;
    const int oy_lo = 2 * row0 - 1;                   // first staged dout row
    const int nrows_o = 2 * rows + 2;
    const int64_t p0 = (static_cast<int64_t>(b) * IH + row0) * IW;   // first pixel (row of hin / ain / dhin) of the tile
    __syncthreads();                                  // previous tile fully consumed
    for (int i = threadIdx.x; i < nrows_o * OW * COUT; i += blockDim.x) {
      const int ry = i / (OW * COUT);
      const int oy = oy_lo + ry;
      s_do[i] = (oy >= 0 && oy < OH) ? __ldg(dout + (static_cast<int64_t>(b) * OH + oy) * OW * COUT + (i - ry * OW * COUT)) : 0.f;
    }
    stage_rows_padded(ain + p0 * Cin, s_a, npx, npx, Cin, ldh);
    __syncthreads();
    // ---- phase 1: task = (pixel, half of the channels): dcols gather (+ publish to s_dc), d hin in place
    for (int task = threadIdx.x; task < 2 * npx; task += blockDim.x) {
      const int px = task >> 1, chalf = task & 1;
      const int ryl = px / IW, ix = px - ryl * IW;
      float dc[KC];
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int ox = 2 * ix - 1 + kw;
          const int ry = 2 * ryl + kh;                // staged row of oy = 2 iy - 1 + kh
#pragma unroll
          for (int co = 0; co < COUT; ++co)
            dc[(kh * 4 + kw) * COUT + co] = (ox >= 0 && ox < OW) ? s_do[(ry * OW + ox) * COUT + co] : 0.f;
        }
      }
      if (chalf == 0) {
#pragma unroll
        for (int k = 0; k < KC; ++k) s_dc[k * npx_max + px] = dc[k];
      }
      float* arow = s_a + px * ldh;
      for (int c = chalf; c < c4n; c += 2) {          // alternate float4s of the row
        const float4 av = *reinterpret_cast<const float4*>(arow + 4 * c);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * Cin + 4 * c);     // (2 addresses per warp)
          g.x += dc[k] * w4.x; g.y += dc[k] * w4.y; g.z += dc[k] * w4.z; g.w += dc[k] * w4.w;
        }
        g.x *= dswish_m(av.x); g.y *= dswish_m(av.y); g.z *= dswish_m(av.z); g.w *= dswish_m(av.w);
        *reinterpret_cast<float4*>(arow + 4 * c) = g;
      }
    }
    __syncthreads();
    // ---- d hin rows out, coalesced
    {
      const int total = npx * c4n, ld4 = ldh >> 2;
      float4* dst = reinterpret_cast<float4*>(dhin + p0 * Cin);
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / c4n, c = i - r * c4n;
        __stcs(dst + i, reinterpret_cast<const float4*>(s_a)[r * ld4 + c]);
      }
    }
    // ---- phase 2: dWt patch += sum over the tile's pixels of dcols[p][k] * hin[p][ci]
    if (lane_id < lanes) {
      const float* hb = hin + p0 * Cin + ci4;
#pragma unroll 4
      for (int px = lane_id; px < npx; px += lanes) {
        const float4 h4 = __ldcs(reinterpret_cast<const float4*>(hb + static_cast<int64_t>(px) * Cin));
        const float4 d4 = make_float4(s_dc[(4 * kq) * npx_max + px], s_dc[(4 * kq + 1) * npx_max + px],
                                      s_dc[(4 * kq + 2) * npx_max + px], s_dc[(4 * kq + 3) * npx_max + px]);
        wacc[0][0] += d4.x * h4.x; wacc[0][1] += d4.x * h4.y; wacc[0][2] += d4.x * h4.z; wacc[0][3] += d4.x * h4.w;
        wacc[1][0] += d4.y * h4.x; wacc[1][1] += d4.y * h4.y; wacc[1][2] += d4.y * h4.z; wacc[1][3] += d4.y * h4.w;
        wacc[2][0] += d4.z * h4.x; wacc[2][1] += d4.z * h4.y; wacc[2][2] += d4.z * h4.z; wacc[2][3] += d4.z * h4.w;
        wacc[3][0] += d4.w * h4.x; wacc[3][1] += d4.w * h4.y; wacc[3][2] += d4.w * h4.z; wacc[3][3] += d4.w * h4.w;
      }
    }
  }
  // ---- reduce the lanes' patches and add to dWt
  __syncthreads();
  if (lane_id < lanes) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(s_red + (static_cast<size_t>(lane_id) * KC + 4 * kq + i) * Cin + ci4) =
          make_float4(wacc[i][0], wacc[i][1], wacc[i][2], wacc[i][3]);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < KC * c4n; idx += blockDim.x) {
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < lanes; ++l) {
      const float4 v = *reinterpret_cast<const float4*>(s_red + static_cast<size_t>(l) * KC * Cin + idx * 4);
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dwt + idx * 4), "f"(s4.x), "f"(s4.y), "f"(s4.z), "f"(s4.w)
                 : "memory");
  }
}

int sm_count() {
  const int n = mvae_device_sm_count();
  return n > 0 ? n : 148;
}

template <typename F>
int set_smem(F fn, size_t bytes, const char* who) {
  if (bytes > 200 * 1024) return set_error(MVAE_ERR_UNSUPPORTED, "%s: tile needs %zu bytes of shared memory", who, bytes);
  if (bytes > 48 * 1024) MVAE_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return MVAE_OK;
}

// input rows per tile of the transposed-conv kernels: whole images when they are small (14 x 14), else ~256 pixels
int tile_rows(int IH, int IW) {
  if (IH * IW <= 256) return IH;
  int tr = 256 / IW;
  return tr < 1 ? 1 : tr;
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_conv_k4s2p1_cin_fwd(const float* x, const float* wc, float* a, float* h, int B, int H, int W, int Cin,
                                        int Cout, void* stream) {
  if (!x || !wc || !a || !h || B < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || (Cin != 1 && Cin != 3) || Cout < 4 ||
      (Cout & 3) || 256 % (Cout / 4) != 0)
    return set_error(MVAE_ERR_BAD_ARG, "conv_k4s2p1_cin_fwd: need Cin in {1,3}, Cout %% 4 == 0 dividing 1024, even H, W");
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(h)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "conv_k4s2p1_cin_fwd: outputs must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t npix = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  const int ppb = 256 / (Cout / 4) * kPix;
  int64_t blocks = (npix + ppb - 1) / ppb;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const size_t smem = static_cast<size_t>(16) * Cin * (Cout + ppb) * sizeof(float);    // weights + one chunk of taps
  if (Cin == 1) conv_cin_fwd_kernel<1><<<static_cast<unsigned>(blocks), 256, smem, st>>>(x, wc, a, h, B, H, W, Cout);
  else conv_cin_fwd_kernel<3><<<static_cast<unsigned>(blocks), 256, smem, st>>>(x, wc, a, h, B, H, W, Cout);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_conv_k4s2p1_cin_wgrad(const float* x, const float* da, float* dwc, int B, int H, int W, int Cin, int Cout,
                                          void* stream) {
  if (!x || !da || !dwc || B < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || (Cin != 1 && Cin != 3) || Cout < 4 || (Cout & 3))
    return set_error(MVAE_ERR_BAD_ARG, "conv_k4s2p1_cin_wgrad: need Cin in {1,3}, Cout %% 4 == 0, even H, W");
  const int tpl = (Cout / 4) * (16 * Cin / 4);
  if (tpl > 256) return set_error(MVAE_ERR_UNSUPPORTED, "conv_k4s2p1_cin_wgrad: Cout * Cin too large (Cout * Cin <= 64)");
  if ((reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(dwc)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "conv_k4s2p1_cin_wgrad: da / dwc must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t npix = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  int64_t blocks = static_cast<int64_t>(sm_count()) * 8;
  int64_t per = (npix + blocks - 1) / blocks;
  if (per < 64) per = 64;
  blocks = (npix + per - 1) / per;
  const int lanes = 256 / tpl;
  const size_t smem = (static_cast<size_t>(lanes) * Cout + kWgChunk) * 16 * Cin * sizeof(float);
  int rc;
  if (Cin == 1) {
    if ((rc = set_smem(conv_cin_wgrad_kernel<1>, smem, "conv_k4s2p1_cin_wgrad"))) return rc;
    conv_cin_wgrad_kernel<1><<<static_cast<unsigned>(blocks), 256, smem, st>>>(x, da, dwc, B, H, W, Cout, static_cast<int>(per));
  } else {
    if ((rc = set_smem(conv_cin_wgrad_kernel<3>, smem, "conv_k4s2p1_cin_wgrad"))) return rc;
    conv_cin_wgrad_kernel<3><<<static_cast<unsigned>(blocks), 256, smem, st>>>(x, da, dwc, B, H, W, Cout, static_cast<int>(per));
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_convt_k4s2p1_cout_fwd(const float* hin, const float* wt, float* out, int B, int IH, int IW, int Cin,
                                          int Cout, void* stream) {
  if (!hin || !wt || !out || B < 1 || IH < 1 || IW < 1 || (Cout != 1 && Cout != 3) || Cin < 4 || (Cin & 3))
    return set_error(MVAE_ERR_BAD_ARG, "convT_k4s2p1_cout_fwd: need Cout in {1,3}, Cin %% 4 == 0");
  if ((reinterpret_cast<uintptr_t>(hin) | reinterpret_cast<uintptr_t>(wt)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "convT_k4s2p1_cout_fwd: hin / wt must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tr = tile_rows(IH, IW);
  const int tiles_per_img = (IH + tr - 1) / tr;
  const int KC = 16 * Cout;
  const size_t smem = (static_cast<size_t>(KC) * Cin + static_cast<size_t>(tr + 2) * IW * (KC + 1)) * sizeof(float);
  int64_t blocks = static_cast<int64_t>(B) * tiles_per_img;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  int rc;
  if (Cout == 1) {
    if ((rc = set_smem(convT_cout_fwd_kernel<1>, smem, "convT_k4s2p1_cout_fwd"))) return rc;
    convT_cout_fwd_kernel<1><<<static_cast<unsigned>(blocks), 256, smem, st>>>(hin, wt, out, B, IH, IW, Cin, tr, tiles_per_img);
  } else {
    if ((rc = set_smem(convT_cout_fwd_kernel<3>, smem, "convT_k4s2p1_cout_fwd"))) return rc;
    convT_cout_fwd_kernel<3><<<static_cast<unsigned>(blocks), 256, smem, st>>>(hin, wt, out, B, IH, IW, Cin, tr, tiles_per_img);
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_convt_k4s2p1_cout_bwd(const float* dout, const float* hin, const float* ain, const float* wt, float* dhin,
                                          float* dwt, int B, int IH, int IW, int Cin, int Cout, void* stream) {
  if (!dout || !hin || !ain || !wt || !dhin || !dwt || B < 1 || IH < 1 || IW < 1 || (Cout != 1 && Cout != 3) || Cin < 4 ||
      (Cin & 3))
    return set_error(MVAE_ERR_BAD_ARG, "convT_k4s2p1_cout_bwd: need Cout in {1,3}, Cin %% 4 == 0");
  const int KC = 16 * Cout;
  const int tpl = (Cin / 4) * (KC / 4);
  if (tpl > 256) return set_error(MVAE_ERR_UNSUPPORTED, "convT_k4s2p1_cout_bwd: Cin * Cout too large (Cin * Cout <= 64)");
  if ((reinterpret_cast<uintptr_t>(hin) | reinterpret_cast<uintptr_t>(ain) | reinterpret_cast<uintptr_t>(wt) |
       reinterpret_cast<uintptr_t>(dhin) | reinterpret_cast<uintptr_t>(dwt)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "convT_k4s2p1_cout_bwd: activations / weights must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tr = tile_rows(IH, IW);
  const int tiles_per_img = (IH + tr - 1) / tr;
  const int lanes = 256 / tpl;
  const size_t s_do = (static_cast<size_t>(2 * tr + 2) * 2 * IW * Cout + 3) & ~size_t(3);      // staged dout rows
  size_t tail = static_cast<size_t>(tr) * IW * (Cin + 8);                       // staged ain rows -> d hin rows ...
  const size_t red = static_cast<size_t>(lanes) * KC * Cin;                   // ... reused for the final lane reduction
  if (red > tail) tail = red;
  const size_t smem = (static_cast<size_t>(KC) * Cin + static_cast<size_t>(tr) * IW * KC + s_do + tail) * sizeof(float);
  int64_t blocks = static_cast<int64_t>(B) * tiles_per_img;
  int rc;
  if (Cout == 1) { if ((rc = set_smem(convT_cout_bwd_kernel<1>, smem, "convT_k4s2p1_cout_bwd"))) return rc; }
  else { if ((rc = set_smem(convT_cout_bwd_kernel<3>, smem, "convT_k4s2p1_cout_bwd"))) return rc; }
  int per_sm = 1;                                                              // persistent: register-resident dWt patches,
  if (Cout == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, convT_cout_bwd_kernel<1>, 256, smem);   // one wave
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, convT_cout_bwd_kernel<3>, 256, smem);
  const int64_t cap = static_cast<int64_t>(sm_count()) * (per_sm < 1 ? 1 : per_sm);
  if (blocks > cap) blocks = cap;
  if (Cout == 1) {
    convT_cout_bwd_kernel<1><<<static_cast<unsigned>(blocks), 256, smem, st>>>(dout, hin, ain, wt, dhin, dwt, B, IH, IW, Cin, tr,
                                                                               tiles_per_img);
  } else {
    if ((rc = set_smem(convT_cout_bwd_kernel<3>, smem, "convT_k4s2p1_cout_bwd"))) return rc;
    convT_cout_bwd_kernel<3><<<static_cast<unsigned>(blocks), 256, smem, st>>>(dout, hin, ain, wt, dhin, dwt, B, IH, IW, Cin, tr,
                                                                               tiles_per_img);
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
