// Bandwidth-bound fused kernels of the MVAE step (sm_100a): product-of-experts + reparametrise + KL
// (forward and backward, any number of modality subsets in one pass over the expert outputs),
// BCE-with-logits / cross-entropy reconstruction loss + analytic gradient, bias-gradient column sums,
// Embedding+Swish, flat Adam.  All HBM traffic is 128-bit vectorised and coalesced where the layout
// allows; reductions are warp-shuffle -> shared -> one double atomic per block.
//
// Reference sites replaced: mnist/model.py:29-35,46-64,156-163,172-185 (variant A),
// celeba/model.py:200-207 (variant B), mnist/train.py:20-94 (elbo_loss, BCE, CE), :168,219 (Adam).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

constexpr int kMaxExperts = 24;
constexpr int kMaxPasses = 32;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// One element of BCE-with-logits: loss = max(x,0) - x*t + log(1 + exp(-|x|)), sig = sigmoid(x).
// MUFU-based exp/log/reciprocal (ex2, lg2, rcp: absolute error < 5e-7 on the log term, 2 ulp on the sigmoid) keep the
// kernel on the HBM roofline instead of the FMA/ALU pipes; the loss SUM is accumulated in double by the callers.
__device__ __forceinline__ void bce_point(float x, float t, float& loss, float& sig) {
  const float e = __expf(-fabsf(x));
  const float d = 1.0f + e;
  const float inv = __fdividef(1.0f, d);
  loss = fmaxf(x, 0.f) - x * t + __logf(d);
  sig = x >= 0.f ? inv : e * inv;
}
__device__ __forceinline__ float dswish_f(float x) {
  const float s = sigmoid_f(x);
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum -> one atomicAdd(double) by thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ void block_atomic_add(double v, double* dst, double* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double s = lane < nw ? scratch[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(dst, s);
  }
}

// Segmented variant: every thread contributes (v, seg); acc[seg] += v.  Fast path when the whole block
// (or at least the whole warp) sits in one segment, per-thread atomics otherwise (segment boundaries).
__device__ __forceinline__ void block_atomic_add_seg(double v, int seg, double* acc, double* scratch, int* seg_smem) {
  if (threadIdx.x == 0) *seg_smem = seg;
  __syncthreads();
  const int seg0 = *seg_smem;
  const bool uniform = __syncthreads_and(seg == seg0 || v == 0.0);
  if (uniform) {
    block_atomic_add(v, acc + seg0, scratch);
  } else {
    const int wseg = __shfl_sync(0xffffffffu, seg, 0);
    if (__all_sync(0xffffffffu, seg == wseg || v == 0.0)) {
      v = warp_sum(v);
      if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(acc + wseg, v);
    } else if (v != 0.0) {
      atomicAdd(acc + seg, v);
    }
  }
}

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t ctr, uint32_t step, float (&n)[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), step, 0x4d564145u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float u0 = (static_cast<float>(c[0]) + 0.5f) * 2.3283064365386963e-10f;
  const float u1 = (static_cast<float>(c[1]) + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = (static_cast<float>(c[2]) + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = (static_cast<float>(c[3]) + 0.5f) * 2.3283064365386963e-10f;
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1; n[3] = r1 * s1;
}

// ---------------------------------------------------------------- PoE + reparam + KL
struct PoeArgs {
  const float* mu_e[kMaxExperts];
  const float* lv_e[kMaxExperts];
  float* dmu_e[kMaxExperts];
  float* dlv_e[kMaxExperts];
  uint32_t masks[kMaxPasses];
  int64_t ld_e, ldd_e, ldz;
  int E, P, B, L;
  int variant, training;
  const float* noise;
  float* noise_out;
  uint64_t seed, offset;
  const int32_t* step_dev;
  float* z;
  float* mu_out;
  float* lv_out;
  double* kl_acc;
  const float* dz;
  const float* dmu_up;   // optional upstream gradient w.r.t. the fused mu output      [P*B, L]
  const float* dlv_up;   // optional upstream gradient w.r.t. the fused logvar output  [P*B, L]
  float kl_scale;
  const float* kl_scale_dev;
  int no_prior;          // 1: the implicit N(0,1) prior expert is NOT part of the product
  // optional per-expert row index (label-table experts, csrc/label_table.cu): expert e's row for sample b is
  // gidx[e][b] of a V-row table instead of row b; backward then red.add's the sample's gradient into that table row
  const int64_t* gidx[kMaxExperts];
};

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
  using type = float4;
};
template <>
struct VecT<1> {
  using type = float;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}

// ---- MUFU forms for the bandwidth-bound fast path (ex2 / lg2 / rcp / sqrt: <= 2 ulp each; the outputs are compared
// against the fp64 oracle at rtol 2e-5, tests/test_kernels_gpu.py).  The IEEE expf / logf / division forms cost ~10-20
// instructions each and made these kernels instruction-bound: 1.9 TB/s at roofline size (profiles/r02_bench_v0_*).
// The .ftz PTX forms are used directly: without -ftz the __expf / __logf / sqrtf intrinsics carry a denormal guard (scale,
// select, sometimes a slow-path call) around every MUFU, which doubled the instruction count of these kernels
// (profiles/r02_poe_raw_key_metrics.txt: 1120 instructions per 4-latent item, issue slots 66 % busy, DRAM 37 %).
__device__ __forceinline__ float fex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float flg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fsqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fsin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fcos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fexp(float x) { return fex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float flog(float x) { return flg2(x) * 0.6931471805599453f; }

// Four N(0,1) draws from one Philox4x32-10 block, Box-Muller with MUFU log / sin / cos (|error| ~ 1e-6 on the draws:
// irrelevant for noise; the stream stays a pure function of (seed, counter, step)).
__device__ __forceinline__ void normal4_fast(uint64_t seed, uint64_t ctr, uint32_t step, float (&n)[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), step, 0x4d564145u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float u0 = (static_cast<float>(c[0]) + 0.5f) * 2.3283064365386963e-10f;
  const float u1 = (static_cast<float>(c[1]) + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = (static_cast<float>(c[2]) + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = (static_cast<float>(c[3]) + 0.5f) * 2.3283064365386963e-10f;
  const float r0 = fsqrt(-1.3862943611198906f * flg2(u0)), r1 = fsqrt(-1.3862943611198906f * flg2(u2));   // -2 ln u
  const float t0 = 6.283185307179586f * u1 - 3.141592653589793f;        // argument in [-pi, pi): MUFU range, full circle
  const float t1 = 6.283185307179586f * u3 - 3.141592653589793f;
  n[0] = r0 * fcos(t0); n[1] = r0 * fsin(t0); n[2] = r1 * fcos(t1); n[3] = r1 * fsin(t1);
}

constexpr int kFastPasses = 4;   // passes whose KL partial sums a thread keeps in registers on the fast path

// Fast path (E <= EMAX <= 4 experts, P <= 4 passes, L % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive latent
// dims of one sample, grid-stride over the samples, 128-bit loads / stores, MUFU math, KL partial sums in registers and
// ONE double atomic per (block, pass) at the very end -- with one atomic per 256 threads per pass the same-address
// atomics serialised in L2 at roofline size (16 K blocks x 3 passes on 3 addresses).
template <int EMAX>
__global__ void __launch_bounds__(256, 3) poe_fwd_fast_kernel(const __grid_constant__ PoeArgs a) {
  pdl_prologue();
  __shared__ double scratch[32];
  const int l4n = a.L >> 2;
  const int64_t total = static_cast<int64_t>(a.B) * l4n;
  const float e1 = 1e-8f;
  const float e2 = a.variant == 0 ? 1e-8f : 0.0f;
  const float T0 = a.no_prior ? 0.0f : 1.0f / ((1.0f + e1) + e2);  // prior expert: mu = 0, logvar = 0
  const uint32_t step = (a.training && a.noise == nullptr && a.step_dev) ? static_cast<uint32_t>(__ldg(a.step_dev)) : 0u;
  float klacc[kFastPasses] = {0.f, 0.f, 0.f, 0.f};
  // (b, l4) walk the grid-stride sequence incrementally: one 64-bit division per thread, none per item
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int db = static_cast<int>(stride / l4n), dl = static_cast<int>(stride - static_cast<int64_t>(db) * l4n);
  int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int b = static_cast<int>(gid / l4n), l4 = static_cast<int>(gid - static_cast<int64_t>(b) * l4n);
  for (; gid < total; gid += stride, b += db, l4 += dl) {
    if (l4 >= l4n) { l4 -= l4n; ++b; }
    const int l = l4 * 4;
    float T[EMAX][4], M[EMAX][4];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < a.E) {
        const int64_t row = a.gidx[e] ? __ldg(a.gidx[e] + b) : b;
        float mu[4], lv[4];
        load_vec<4>(a.mu_e[e] + row * a.ld_e + l, mu);
        load_vec<4>(a.lv_e[e] + row * a.ld_e + l, lv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          T[e][q] = frcp((fexp(lv[q]) + e1) + e2);
          M[e][q] = mu[q] * T[e][q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { T[e][q] = 0.f; M[e][q] = 0.f; }
      }
    }
#pragma unroll
    for (int p = 0; p < kFastPasses; ++p) {
      if (p < a.P) {
        const uint32_t mask = a.masks[p];
        float S[4], N[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { S[q] = T0; N[q] = 0.0f; }
#pragma unroll
        for (int e = 0; e < EMAX; ++e) {
          if ((mask >> e) & 1u) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { S[q] += T[e][q]; N[q] += M[e][q]; }
          }
        }
        const int64_t row = static_cast<int64_t>(p) * a.B + b;
        float nz[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.training) {
          if (a.noise != nullptr) {
            load_vec<4>(a.noise + row * a.L + l, nz);
          } else {
            normal4_fast(a.seed, a.offset + static_cast<uint64_t>(row) * l4n + (l >> 2), step, nz);
            store_vec<4>(a.noise_out + row * a.L + l, nz);
          }
        }
        float z[4], mu[4], lv[4];
        float klf = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float pv = frcp(S[q]);
          mu[q] = N[q] * pv;
          const float ev = pv + e2;                 // = exp(logvar) of the fused posterior
          lv[q] = flog(ev);
          z[q] = a.training ? nz[q] * fsqrt(ev) + mu[q] : mu[q];        // exp(0.5 logvar) = sqrt(exp(logvar))
          klf += 1.0f + lv[q] - mu[q] * mu[q] - ev;
        }
        klacc[p] += -0.5f * klf;
        store_vec<4>(a.z + row * a.ldz + l, z);
        if (a.mu_out) store_vec<4>(a.mu_out + row * a.L + l, mu);
        if (a.lv_out) store_vec<4>(a.lv_out + row * a.L + l, lv);
      }
    }
  }
  if (a.kl_acc != nullptr) {
#pragma unroll
    for (int p = 0; p < kFastPasses; ++p)
      if (p < a.P) block_atomic_add(static_cast<double>(klacc[p]), a.kl_acc + p, scratch);
  }
}

template <int EMAX>
__global__ void __launch_bounds__(256, 3) poe_bwd_fast_kernel(const __grid_constant__ PoeArgs a) {
  pdl_prologue();
  const int l4n = a.L >> 2;
  const int64_t total = static_cast<int64_t>(a.B) * l4n;
  const float e1 = 1e-8f;
  const float e2 = a.variant == 0 ? 1e-8f : 0.0f;
  const float kls = a.kl_scale * (a.kl_scale_dev ? __ldg(a.kl_scale_dev) : 1.0f);
  const float T0 = a.no_prior ? 0.0f : 1.0f / ((1.0f + e1) + e2);
  // (b, l4) walk the grid-stride sequence incrementally: one 64-bit division per thread, none per item
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int db = static_cast<int>(stride / l4n), dl = static_cast<int>(stride - static_cast<int64_t>(db) * l4n);
  int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int b = static_cast<int>(gid / l4n), l4 = static_cast<int>(gid - static_cast<int64_t>(b) * l4n);
  for (; gid < total; gid += stride, b += db, l4 += dl) {
    if (l4 >= l4n) { l4 -= l4n; ++b; }
    const int l = l4 * 4;
    float T[EMAX][4], MU[EMAX][4], EX[EMAX][4], dMU[EMAX][4], dLV[EMAX][4];
    int64_t erow[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { T[e][q] = 0.f; MU[e][q] = 0.f; EX[e][q] = 0.f; dMU[e][q] = 0.f; dLV[e][q] = 0.f; }
      erow[e] = b;
      if (e < a.E) {
        if (a.gidx[e]) erow[e] = __ldg(a.gidx[e] + b);
        float lv[4];
        load_vec<4>(a.mu_e[e] + erow[e] * a.ld_e + l, MU[e]);
        load_vec<4>(a.lv_e[e] + erow[e] * a.ld_e + l, lv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          EX[e][q] = fexp(lv[q]);
          T[e][q] = frcp((EX[e][q] + e1) + e2);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kFastPasses; ++p) {
      if (p < a.P) {
        const uint32_t mask = a.masks[p];
        float S[4], N[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { S[q] = T0; N[q] = 0.f; }
#pragma unroll
        for (int e = 0; e < EMAX; ++e) {
          if ((mask >> e) & 1u) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { S[q] += T[e][q]; N[q] += MU[e][q] * T[e][q]; }
          }
        }
        const int64_t row = static_cast<int64_t>(p) * a.B + b;
        float dz[4], nz[4] = {0.f, 0.f, 0.f, 0.f}, gmu_up[4] = {0.f, 0.f, 0.f, 0.f}, glv_up[4] = {0.f, 0.f, 0.f, 0.f};
        load_vec<4>(a.dz + row * a.ldz + l, dz);
        if (a.training) load_vec<4>(a.noise + row * a.L + l, nz);
        if (a.dmu_up) load_vec<4>(a.dmu_up + row * a.L + l, gmu_up);
        if (a.dlv_up) load_vec<4>(a.dlv_up + row * a.L + l, glv_up);
        float g_mu[4], g_lvS[4], mu[4], invS[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          invS[q] = frcp(S[q]);
          mu[q] = N[q] * invS[q];
          const float pv = invS[q];
          const float ev = pv + e2;                                 // exp(logvar)
          g_mu[q] = dz[q] + gmu_up[q] + kls * mu[q];
          float g_lv = glv_up[q] + kls * 0.5f * (ev - 1.0f);
          if (a.training) g_lv += dz[q] * nz[q] * 0.5f * fsqrt(ev);
          g_lvS[q] = g_lv * (-(pv * pv) * frcp(ev));                // d logvar / d S
        }
#pragma unroll
        for (int e = 0; e < EMAX; ++e) {
          if ((mask >> e) & 1u) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dMU[e][q] += g_mu[q] * T[e][q] * invS[q];
              const float dT = g_mu[q] * (MU[e][q] - mu[q]) * invS[q] + g_lvS[q];
              dLV[e][q] += dT * (-(T[e][q] * T[e][q]) * EX[e][q]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < a.E) {
        float* pm = a.dmu_e[e] + erow[e] * a.ldd_e + l;
        float* pl = a.dlv_e[e] + erow[e] * a.ldd_e + l;
        if (a.gidx[e]) {   // table-backed expert: all samples of a class add into the class's row
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pm), "f"(dMU[e][0]), "f"(dMU[e][1]),
                       "f"(dMU[e][2]), "f"(dMU[e][3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pl), "f"(dLV[e][0]), "f"(dLV[e][1]),
                       "f"(dLV[e][2]), "f"(dLV[e][3]) : "memory");
        } else {
          store_vec<4>(pm, dMU[e]);
          store_vec<4>(pl, dLV[e]);
        }
      }
    }
  }
}

// One thread = VEC consecutive latent dims of one sample.  Expert precisions are computed once and
// reused by every pass.
template <int VEC, int EMAX>
__global__ void __launch_bounds__(256) poe_fwd_kernel(const __grid_constant__ PoeArgs a) {
  pdl_prologue();
  __shared__ double scratch[32];
  const int lv_per_row = a.L / VEC;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = gid < static_cast<int64_t>(a.B) * lv_per_row;
  const int b = active ? static_cast<int>(gid / lv_per_row) : 0;
  const int l = active ? static_cast<int>(gid - static_cast<int64_t>(b) * lv_per_row) * VEC : 0;
  const float e1 = 1e-8f;
  const float e2 = a.variant == 0 ? 1e-8f : 0.0f;
  float T[EMAX][VEC], M[EMAX][VEC];
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (e < a.E && active) {
      float mu[VEC], lv[VEC];
      const int64_t erow = a.gidx[e] ? a.gidx[e][b] : b;
      load_vec<VEC>(a.mu_e[e] + erow * a.ld_e + l, mu);
      load_vec<VEC>(a.lv_e[e] + erow * a.ld_e + l, lv);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const float var = expf(lv[q]) + e1;
        T[e][q] = 1.0f / (var + e2);
        M[e][q] = mu[q] * T[e][q];
      }
    } else {
#pragma unroll
      for (int q = 0; q < VEC; ++q) { T[e][q] = 0.f; M[e][q] = 0.f; }
    }
  }
  const float T0 = a.no_prior ? 0.0f : 1.0f / ((1.0f + e1) + e2);  // prior expert: mu = 0, logvar = 0
  for (int p = 0; p < a.P; ++p) {
    const uint32_t mask = a.masks[p];
    float z[VEC], mu[VEC], lv[VEC];
    double kl = 0.0;
    if (active) {
      float S[VEC], N[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) { S[q] = T0; N[q] = 0.0f * T0; }
#pragma unroll
      for (int e = 0; e < EMAX; ++e) {
        if ((mask >> e) & 1u) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) { S[q] += T[e][q]; N[q] += M[e][q]; }
        }
      }
      const int64_t row = static_cast<int64_t>(p) * a.B + b;
      float nz[VEC];
      if (a.training) {
        if (a.noise != nullptr) {
          load_vec<VEC>(a.noise + row * a.L + l, nz);
        } else {
          float n4[4];
          normal4(a.seed, a.offset + static_cast<uint64_t>(row) * a.L / VEC + l / VEC,
                  a.step_dev ? static_cast<uint32_t>(__ldg(a.step_dev)) : 0u, n4);
#pragma unroll
          for (int q = 0; q < VEC; ++q) nz[q] = n4[q];
          store_vec<VEC>(a.noise_out + row * a.L + l, nz);
        }
      }
      float klf = 0.f;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        mu[q] = N[q] / S[q];
        const float pv = 1.0f / S[q];
        lv[q] = logf(pv + e2);
        z[q] = a.training ? nz[q] * expf(0.5f * lv[q]) + mu[q] : mu[q];
        klf += 1.0f + lv[q] - mu[q] * mu[q] - expf(lv[q]);
      }
      kl = -0.5 * static_cast<double>(klf);
      store_vec<VEC>(a.z + row * a.ldz + l, z);
      if (a.mu_out) store_vec<VEC>(a.mu_out + row * a.L + l, mu);
      if (a.lv_out) store_vec<VEC>(a.lv_out + row * a.L + l, lv);
    }
    if (a.kl_acc != nullptr) block_atomic_add(kl, a.kl_acc + p, scratch);
  }
}

template <int VEC, int EMAX>
__global__ void __launch_bounds__(256) poe_bwd_kernel(const __grid_constant__ PoeArgs a) {
  pdl_prologue();
  const int lv_per_row = a.L / VEC;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(a.B) * lv_per_row) return;
  const int b = static_cast<int>(gid / lv_per_row);
  const int l = static_cast<int>(gid - static_cast<int64_t>(b) * lv_per_row) * VEC;
  const float e1 = 1e-8f;
  const float e2 = a.variant == 0 ? 1e-8f : 0.0f;
  const float kls = a.kl_scale * (a.kl_scale_dev ? __ldg(a.kl_scale_dev) : 1.0f);
  float T[EMAX][VEC], MU[EMAX][VEC], EX[EMAX][VEC], dMU[EMAX][VEC], dLV[EMAX][VEC];
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) { T[e][q] = 0.f; MU[e][q] = 0.f; EX[e][q] = 0.f; dMU[e][q] = 0.f; dLV[e][q] = 0.f; }
    if (e < a.E) {
      float lv[VEC];
      const int64_t erow = a.gidx[e] ? a.gidx[e][b] : b;
      load_vec<VEC>(a.mu_e[e] + erow * a.ld_e + l, MU[e]);
      load_vec<VEC>(a.lv_e[e] + erow * a.ld_e + l, lv);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        EX[e][q] = expf(lv[q]);
        T[e][q] = 1.0f / ((EX[e][q] + e1) + e2);
      }
    }
  }
  const float T0 = a.no_prior ? 0.0f : 1.0f / ((1.0f + e1) + e2);
  for (int p = 0; p < a.P; ++p) {
    const uint32_t mask = a.masks[p];
    float S[VEC], N[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) { S[q] = T0; N[q] = 0.f; }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if ((mask >> e) & 1u) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) { S[q] += T[e][q]; N[q] += MU[e][q] * T[e][q]; }
      }
    }
    const int64_t row = static_cast<int64_t>(p) * a.B + b;
    float dz[VEC], nz[VEC], gmu_up[VEC], glv_up[VEC];
    load_vec<VEC>(a.dz + row * a.ldz + l, dz);
    if (a.training) load_vec<VEC>(a.noise + row * a.L + l, nz);
#pragma unroll
    for (int q = 0; q < VEC; ++q) { gmu_up[q] = 0.f; glv_up[q] = 0.f; }
    if (a.dmu_up) load_vec<VEC>(a.dmu_up + row * a.L + l, gmu_up);
    if (a.dlv_up) load_vec<VEC>(a.dlv_up + row * a.L + l, glv_up);
    float g_mu[VEC], g_lvS[VEC], mu[VEC], invS[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      invS[q] = 1.0f / S[q];
      mu[q] = N[q] * invS[q];
      const float pv = invS[q];
      const float lv = logf(pv + e2);
      g_mu[q] = dz[q] + gmu_up[q] + kls * mu[q];
      float g_lv = glv_up[q] + kls * 0.5f * (expf(lv) - 1.0f);
      if (a.training) g_lv += dz[q] * nz[q] * 0.5f * expf(0.5f * lv);
      g_lvS[q] = g_lv * (-(pv * pv) / (pv + e2));  // d logvar / d S
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if ((mask >> e) & 1u) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          dMU[e][q] += g_mu[q] * T[e][q] * invS[q];
          const float dT = g_mu[q] * (MU[e][q] - mu[q]) * invS[q] + g_lvS[q];
          dLV[e][q] += dT * (-(T[e][q] * T[e][q]) * EX[e][q]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (e < a.E) {
      if (a.gidx[e]) {
        const int64_t erow = a.gidx[e][b];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          atomicAdd(a.dmu_e[e] + erow * a.ldd_e + l + q, dMU[e][q]);
          atomicAdd(a.dlv_e[e] + erow * a.ldd_e + l + q, dLV[e][q]);
        }
      } else {
        store_vec<VEC>(a.dmu_e[e] + static_cast<int64_t>(b) * a.ldd_e + l, dMU[e]);
        store_vec<VEC>(a.dlv_e[e] + static_cast<int64_t>(b) * a.ldd_e + l, dLV[e]);
      }
    }
  }
}

// ---------------------------------------------------------------- stand-alone reparametrise (module API)
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                          const float* noise, float* noise_out, uint64_t seed,
                                                          uint64_t offset, float* z, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e;
  if (noise != nullptr) {
    e = noise[i];
  } else {
    float n4[4];
    normal4(seed, offset + static_cast<uint64_t>(i), 0u, n4);
    e = n4[0];
    noise_out[i] = e;
  }
  z[i] = e * expf(0.5f * lv[i]) + mu[i];
}
__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ lv, const float* __restrict__ noise,
                                                          const float* __restrict__ dz, float* dlv, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dlv[i] = dz[i] * noise[i] * 0.5f * expf(0.5f * lv[i]);
}

// ---------------------------------------------------------------- stand-alone KL(q || N(0,1)): sum + grad
__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv, float* dmu,
                                                 float* dlv, int64_t n, float scale, double* acc) {
  __shared__ double scratch[32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  double kl = 0.0;
  if (i < n) {
    const float m = mu[i], l = lv[i], e = expf(l);
    kl = -0.5 * static_cast<double>(1.0f + l - m * m - e);
    if (dmu) dmu[i] = scale * m;
    if (dlv) dlv[i] = scale * 0.5f * (e - 1.0f);
  }
  if (acc != nullptr) block_atomic_add(kl, acc, scratch);
}

// ---------------------------------------------------------------- BCE with logits: loss + grad
// Each block streams `slabs` consecutive 16 KiB slabs (4 x float4 per thread in flight), keeps its loss partial
// in a register and issues ONE double atomic per (block, segment): with one slab per block the same-address
// atomics serialised in L2 and capped the kernel at 43 % of HBM bandwidth at roofline size.  32-bit index math
// keeps the kernel at 4 blocks/SM.
template <int kBceUnroll>
__global__ void __launch_bounds__(256, kBceUnroll == 4 ? 4 : 2) bce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ t,
                                                     int ldt, int t_rows, float* dx, int lddx, int R, int D4,
                                                     float scale, double* loss_acc, int seg_rows, float* loss_elem,
                                                     int ldl, int slabs) {
  pdl_prologue();
  __shared__ double scratch[32];
  __shared__ int seg_smem;
  const unsigned n4 = static_cast<unsigned>(R) * static_cast<unsigned>(D4);
  const unsigned slab = 256u * kBceUnroll;
  double acc = 0.0;
  int cur_seg = -1;
  for (int sidx = 0; sidx < slabs; ++sidx) {
    const unsigned base = (blockIdx.x * static_cast<unsigned>(slabs) + sidx) * slab + threadIdx.x;
    if (base - threadIdx.x >= n4) break;
    float4 xv[kBceUnroll], tv[kBceUnroll];
    unsigned rr[kBceUnroll], cc[kBceUnroll];
    bool ok[kBceUnroll];
#pragma unroll
    for (int u = 0; u < kBceUnroll; ++u) {
      const unsigned i = base + u * 256u;
      ok[u] = i < n4;
      if (ok[u]) {
        const unsigned r = i / static_cast<unsigned>(D4);
        const unsigned c = (i - r * D4) * 4u;
        rr[u] = r; cc[u] = c;
        xv[u] = __ldcs(reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + c));
        tv[u] = __ldg(reinterpret_cast<const float4*>(t + static_cast<size_t>(r % static_cast<unsigned>(t_rows)) * ldt + c));
      }
    }
#pragma unroll
    for (int u = 0; u < kBceUnroll; ++u) {
      if (!ok[u]) continue;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      const float ts[4] = {tv[u].x, tv[u].y, tv[u].z, tv[u].w};
      float g[4], le[4];
      float lsum = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s;
        bce_point(xs[q], ts[q], le[q], s);
        lsum += le[q];
        g[q] = scale * (s - ts[q]);
      }
      const int seg = static_cast<int>(rr[u] / static_cast<unsigned>(seg_rows));
      if (seg != cur_seg) {  // rows are visited in increasing order: flush the finished segment (rare)
        if (cur_seg >= 0 && loss_acc != nullptr && acc != 0.0) atomicAdd(loss_acc + cur_seg, acc);
        cur_seg = seg;
        acc = 0.0;
      }
      acc += static_cast<double>(lsum);
      if (dx != nullptr)
        __stcs(reinterpret_cast<float4*>(dx + static_cast<size_t>(rr[u]) * lddx + cc[u]), make_float4(g[0], g[1], g[2], g[3]));
      if (loss_elem != nullptr)
        __stcs(reinterpret_cast<float4*>(loss_elem + static_cast<size_t>(rr[u]) * ldl + cc[u]),
               make_float4(le[0], le[1], le[2], le[3]));
    }
  }
  if (loss_acc != nullptr) block_atomic_add_seg(acc, cur_seg < 0 ? 0 : cur_seg, loss_acc, scratch, &seg_smem);
}

// Stacked-pass variant: x holds `copies` (<= 4) passes of t_rows rows each that share ONE target tensor (the image is
// the target of both the joint and the image-only pass).  Each thread loads its target float4 once and applies it to
// all copies: 4 + 4/copies + 4 bytes of traffic per element instead of 12, one loss accumulator per copy.
template <int COPIES>
__global__ void __launch_bounds__(256, 4) bce_stacked_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ t,
                                                             int ldt, int t_rows, float* dx, int lddx, int D4, float scale,
                                                             double* loss_acc) {
  pdl_prologue();
  __shared__ double scratch[32];
  const unsigned n4 = static_cast<unsigned>(t_rows) * static_cast<unsigned>(D4);
  double acc[COPIES];
#pragma unroll
  for (int k = 0; k < COPIES; ++k) acc[k] = 0.0;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < n4; i += gridDim.x * 256u) {
    const unsigned r = i / static_cast<unsigned>(D4);
    const unsigned c = (i - r * D4) * 4u;
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t + static_cast<size_t>(r) * ldt + c));
    float4 xv[COPIES];
#pragma unroll
    for (int k = 0; k < COPIES; ++k)
      xv[k] = __ldcs(reinterpret_cast<const float4*>(x + (static_cast<size_t>(k) * t_rows + r) * ldx + c));
    const float ts[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
    for (int k = 0; k < COPIES; ++k) {
      const float xs[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w};
      float g[4];
      float lsum = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float le, s;
        bce_point(xs[q], ts[q], le, s);
        lsum += le;
        g[q] = scale * (s - ts[q]);
      }
      acc[k] += static_cast<double>(lsum);
      if (dx != nullptr)
        __stcs(reinterpret_cast<float4*>(dx + (static_cast<size_t>(k) * t_rows + r) * lddx + c),
               make_float4(g[0], g[1], g[2], g[3]));
    }
  }
  if (loss_acc != nullptr) {
#pragma unroll
    for (int k = 0; k < COPIES; ++k) block_atomic_add(acc[k], loss_acc + k, scratch);
  }
}

// BCE for narrow rows (D not a multiple of 4, e.g. the 18 CelebA attributes): one thread per row.
__global__ void __launch_bounds__(256) bce_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ t,
                                                       int64_t ldt, int t_rows, float* dx, int64_t lddx, int R, int D,
                                                       float scale, double* loss_acc, int seg_rows, float* loss_elem,
                                                       int64_t ldl) {
  __shared__ double scratch[32];
  __shared__ int seg_smem;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (r < R) {
    const float* xr = x + static_cast<int64_t>(r) * ldx;
    const float* tr = t + static_cast<int64_t>(r % t_rows) * ldt;
    for (int k = 0; k < D; ++k) {
      const float xq = xr[k], tq = tr[k];
      const float e = expf(-fabsf(xq));
      const float inv = 1.0f / (1.0f + e);
      const float le = fmaxf(xq, 0.f) - xq * tq + logf(1.0f + e);
      loss += static_cast<double>(le);
      if (loss_elem) loss_elem[static_cast<int64_t>(r) * ldl + k] = le;
      if (dx) dx[static_cast<int64_t>(r) * lddx + k] = scale * ((xq >= 0.f ? inv : e * inv) - tq);
    }
  }
  if (loss_acc != nullptr) block_atomic_add_seg(loss, (r < R ? r : R - 1) / seg_rows, loss_acc, scratch, &seg_smem);
}

// ---------------------------------------------------------------- cross entropy (K small): one thread per row
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ target,
                                                 int t_rows, float* dx, int64_t lddx, int R, int K, float scale,
                                                 double* loss_acc, int seg_rows, float* loss_rows, int64_t ldl) {
  pdl_prologue();
  __shared__ double scratch[32];
  __shared__ int seg_smem;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (r < R) {
    const float* xr = x + static_cast<int64_t>(r) * ldx;
    const int tg = static_cast<int>(target[r % t_rows]);
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, xr[k] + 1e-6f);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf((xr[k] + 1e-6f) - mx);
    const float lse = mx + logf(se);
    loss = -static_cast<double>((xr[tg] + 1e-6f) - lse);
    if (loss_rows != nullptr) {
      float* lr = loss_rows + static_cast<int64_t>(r) * ldl;
      for (int k = 0; k < K; ++k) lr[k] = (k == tg) ? -((xr[k] + 1e-6f) - lse) : 0.0f;
    }
    if (dx != nullptr) {
      float* dr = dx + static_cast<int64_t>(r) * lddx;
      for (int k = 0; k < K; ++k) {
        const float pk = expf((xr[k] + 1e-6f) - lse);
        dr[k] = scale * (pk - (k == tg ? 1.0f : 0.0f));
      }
    }
  }
  if (loss_acc != nullptr)
    block_atomic_add_seg(loss, (r < R ? r : R - 1) / seg_rows, loss_acc, scratch, &seg_smem);
}

// ---------------------------------------------------------------- column sums (bias gradients)
// block = 32 columns x 8 row lanes; grid.x = column blocks, grid.y = row chunks of kColsumRows.
constexpr int kColsumRows = 256;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, int64_t ld, float* db, int M, int N) {
  pdl_prologue();
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * kColsumRows;
  const int r1 = min(M, r0 + kColsumRows);
  float s = 0.f;
  if (c < N)
    for (int r = r0 + rl; r < r1; r += 8) s += dy[static_cast<int64_t>(r) * ld + c];
  part[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x & 31];
    atomicAdd(db + c, tot);
  }
}

// ---------------------------------------------------------------- swish
__global__ void swish_fwd_kernel(const float* __restrict__ x, float* y, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    *reinterpret_cast<float4*>(y + i) =
        make_float4(v.x * sigmoid_f(v.x), v.y * sigmoid_f(v.y), v.z * sigmoid_f(v.z), v.w * sigmoid_f(v.w));
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = x[j] * sigmoid_f(x[j]);
  }
}
__global__ void swish_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* dx, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    const float4 g = *reinterpret_cast<const float4*>(dy + i);
    *reinterpret_cast<float4*>(dx + i) =
        make_float4(g.x * dswish_f(v.x), g.y * dswish_f(v.y), g.z * dswish_f(v.z), g.w * dswish_f(v.w));
  } else {
    for (int64_t j = i; j < n; ++j) dx[j] = dy[j] * dswish_f(x[j]);
  }
}

// ---------------------------------------------------------------- embedding + swish
// MUFU sigmoid (ex2.approx + rcp.approx, relative error ~(2 + |x|) * 2^-23) for the bandwidth-bound kernels
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__global__ void emb_swish_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, float* a,
                                     float* h, int B, int D4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * D4) return;
  const int b = static_cast<int>(i / D4);
  const int c = static_cast<int>(i - static_cast<int64_t>(b) * D4);
  const float4 v = __ldg(reinterpret_cast<const float4*>(table) + idx[b] * D4 + c);
  if (a != nullptr) reinterpret_cast<float4*>(a)[i] = v;
  reinterpret_cast<float4*>(h)[i] = make_float4(v.x * sigmoid_fast(v.x), v.y * sigmoid_fast(v.y), v.z * sigmoid_fast(v.z),
                                                v.w * sigmoid_fast(v.w));
}
// Embedding backward through Swish = segmented row sum by class (only V rows receive gradient).
// grid = (ceil(D/128), row chunks of kEmbRows); block = 128 threads (one column each); per-class partial sums
// live in shared memory, one atomic per (class, column) per block.
constexpr int kEmbRows = 32;   // 128 rows per block left 128 blocks x 4 warps on the machine: 25 us of exposed DRAM latency
constexpr int kEmbMaxV = 32;
__global__ void __launch_bounds__(128) emb_swish_bwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                                            const float* __restrict__ dh, int64_t lddh, float* dtable,
                                                            int B, int D, int V) {
  __shared__ float acc[kEmbMaxV][128];
  __shared__ int sidx[kEmbRows];
  const int d = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * kEmbRows, r1 = min(B, r0 + kEmbRows);
  for (int v = 0; v < V; ++v) acc[v][threadIdx.x] = 0.f;
  for (int r = r0 + threadIdx.x; r < r1; r += 128) sidx[r - r0] = static_cast<int>(idx[r]);
  __syncthreads();
  if (d < D) {
    // all loads of a group of 8 rows in flight before the shared-memory accumulation
    for (int rb = r0; rb < r1; rb += 8) {
      float v8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v8[j] = (rb + j < r1) ? __ldg(dh + static_cast<int64_t>(rb + j) * lddh + d) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (rb + j < r1) acc[sidx[rb + j - r0]][threadIdx.x] += v8[j];
    }
    for (int v = 0; v < V; ++v) {
      const float s = acc[v][threadIdx.x];
      if (s != 0.f) atomicAdd(dtable + static_cast<int64_t>(v) * D + d, s * dswish_f(table[static_cast<int64_t>(v) * D + d]));
    }
  }
}

// ---------------------------------------------------------------- Adam
__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* __restrict__ g, float* m, float* v, int64_t n,
                                                   float lr, const float* lr_mult_dev, float beta1, float beta2,
                                                   float eps, float grad_scale, const int32_t* step_count) {
  pdl_prologue();
  // bias corrections: one double-precision pow per BLOCK (thread 0), broadcast through shared memory
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  if (threadIdx.x == 0) {
    const int t = *step_count + 1;
    const float lr_eff = lr * (lr_mult_dev ? __ldg(lr_mult_dev) : 1.0f);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), t);
    s_step_size = static_cast<float>(lr_eff / bc1);
    s_inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float step_size = s_step_size;
  const float inv_sqrt_bc2 = s_inv_sqrt_bc2;
  if (i + 3 < n) {
    float4 P = *reinterpret_cast<float4*>(p + i);
    const float4 G = *reinterpret_cast<const float4*>(g + i);
    float4 Mv = *reinterpret_cast<float4*>(m + i);
    float4 Vv = *reinterpret_cast<float4*>(v + i);
    float* pp = reinterpret_cast<float*>(&P);
    const float* gg = reinterpret_cast<const float*>(&G);
    float* mm = reinterpret_cast<float*>(&Mv);
    float* vv = reinterpret_cast<float*>(&Vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gq = gg[q] * grad_scale;
      mm[q] = beta1 * mm[q] + (1.0f - beta1) * gq;
      vv[q] = beta2 * vv[q] + (1.0f - beta2) * gq * gq;
      const float denom = sqrtf(vv[q]) * inv_sqrt_bc2 + eps;
      pp[q] -= step_size * (mm[q] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = P;
    *reinterpret_cast<float4*>(m + i) = Mv;
    *reinterpret_cast<float4*>(v + i) = Vv;
  } else {
    for (int64_t j = i; j < n; ++j) {
      const float gq = g[j] * grad_scale;
      m[j] = beta1 * m[j] + (1.0f - beta1) * gq;
      v[j] = beta2 * v[j] + (1.0f - beta2) * gq * gq;
      const float denom = sqrtf(v[j]) * inv_sqrt_bc2 + eps;
      p[j] -= step_size * (m[j] / denom);
    }
  }
}
__global__ void step_inc_kernel(int32_t* step_count) {
  pdl_prologue();
  *step_count += 1;
}

__global__ void elbo_finalize_kernel(const double* recon_img, const double* recon_txt, const double* kl, int P,
                                     float lambda_image, float lambda_text, float beta, const float* beta_dev,
                                     float inv_batch, float* out) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double b = static_cast<double>(beta) * (beta_dev ? static_cast<double>(*beta_dev) : 1.0);
  double tot = 0.0;
  for (int p = 0; p < P; ++p) {
    double v = 0.0;
    if (recon_img) v += lambda_image * recon_img[p];
    if (recon_txt) v += lambda_text * recon_txt[p];
    if (kl) v += b * kl[p];
    v *= inv_batch;
    out[1 + p] = static_cast<float>(v);
    tot += v;
  }
  out[0] = static_cast<float>(tot);
}

int fill_poe_args(PoeArgs& a, const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                  const uint32_t* masks, int P, int B, int L, int variant, int training,
                  const int64_t* const* gather_idx) {
  if (E < 1 || E > kMaxExperts) return set_error(MVAE_ERR_BAD_ARG, "poe: E=%d out of [1,%d]", E, kMaxExperts);
  if (P < 1 || P > kMaxPasses) return set_error(MVAE_ERR_BAD_ARG, "poe: P=%d out of [1,%d]", P, kMaxPasses);
  if (B < 1 || L < 1 || !mu_e || !lv_e || !masks) return set_error(MVAE_ERR_BAD_ARG, "poe: bad B/L/pointers");
  if (variant < 0 || variant > 3)
    return set_error(MVAE_ERR_BAD_ARG, "poe: variant must be 0 (A) or 1 (B), optionally | MVAE_POE_NO_PRIOR");
  for (int e = 0; e < E; ++e) {
    if (!mu_e[e] || !lv_e[e]) return set_error(MVAE_ERR_BAD_ARG, "poe: expert %d has a NULL pointer", e);
    a.mu_e[e] = mu_e[e]; a.lv_e[e] = lv_e[e];
    a.gidx[e] = gather_idx ? gather_idx[e] : nullptr;
  }
  for (int p = 0; p < P; ++p) {
    if (E < 32 && (masks[p] >> E) != 0) return set_error(MVAE_ERR_BAD_ARG, "poe: pass %d references an expert >= E", p);
    a.masks[p] = masks[p];
  }
  a.ld_e = ld_e; a.E = E; a.P = P; a.B = B; a.L = L; a.variant = variant & 1; a.no_prior = (variant >> 1) & 1;
  a.training = training;
  return MVAE_OK;
}

// grid of the grid-stride fast kernels: enough blocks to fill the machine (8 x 256 threads per SM), never more than the work
unsigned fast_grid(int64_t n_threads) {
  const int sms = mvae_device_sm_count() > 0 ? mvae_device_sm_count() : 148;
  int64_t blocks = (n_threads + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms) * 8;
  return static_cast<unsigned>(blocks < cap ? blocks : cap);
}

bool vec4_ok(const PoeArgs& a, bool bwd) {
  if (a.E > 4 || (a.L & 3) || (a.ld_e & 3) || (a.ldz & 3)) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int e = 0; e < a.E; ++e) {
    if (!al(a.mu_e[e]) || !al(a.lv_e[e])) return false;
    if (bwd && (!al(a.dmu_e[e]) || !al(a.dlv_e[e]) || (a.ldd_e & 3))) return false;
  }
  if (!al(a.noise) || !al(a.noise_out) || !al(a.z) || !al(a.mu_out) || !al(a.lv_out) || !al(a.dz)) return false;
  if (!al(a.dmu_up) || !al(a.dlv_up)) return false;
  return true;
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_poe_fwd(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                            const uint32_t* pass_masks, int P, int B, int L, int variant, int training,
                            const float* noise, float* noise_out, uint64_t seed, uint64_t offset,
                            const int32_t* step_dev, float* z, int64_t ldz, float* mu_out, float* lv_out,
                            double* kl_acc, void* stream) {
  return mvae_poe_fwd_g(mu_e, lv_e, ld_e, E, nullptr, pass_masks, P, B, L, variant, training, noise, noise_out, seed, offset,
                        step_dev, z, ldz, mu_out, lv_out, kl_acc, stream);
}

extern "C" int mvae_poe_fwd_g(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                              const int64_t* const* gather_idx, const uint32_t* pass_masks, int P, int B, int L,
                              int variant, int training, const float* noise, float* noise_out, uint64_t seed,
                              uint64_t offset, const int32_t* step_dev, float* z, int64_t ldz, float* mu_out,
                              float* lv_out, double* kl_acc, void* stream) {
  PoeArgs a = {};
  int rc = fill_poe_args(a, mu_e, lv_e, ld_e, E, pass_masks, P, B, L, variant, training, gather_idx);
  if (rc) return rc;
  if (!z) return set_error(MVAE_ERR_BAD_ARG, "poe_fwd: z is NULL");
  if (training && !noise && !noise_out)
    return set_error(MVAE_ERR_BAD_ARG, "poe_fwd: training without noise needs noise_out for the Philox draws");
  a.noise = noise; a.noise_out = noise_out; a.seed = seed; a.offset = offset; a.step_dev = step_dev;
  a.z = z; a.ldz = ldz; a.mu_out = mu_out; a.lv_out = lv_out; a.kl_acc = kl_acc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (vec4_ok(a, false) && P <= kFastPasses && !getenv("MVAE_POE_SLOW")) {
    const int64_t n = static_cast<int64_t>(B) * (L / 4);
    if (E <= 2) launch_pdl(poe_fwd_fast_kernel<2>, dim3(fast_grid(n)), dim3(256), 0, st, a);
    else launch_pdl(poe_fwd_fast_kernel<4>, dim3(fast_grid(n)), dim3(256), 0, st, a);
  } else if (vec4_ok(a, false)) {
    const int64_t n = static_cast<int64_t>(B) * (L / 4);
    launch_pdl(poe_fwd_kernel<4, 4>, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, a);
  } else {
    const int64_t n = static_cast<int64_t>(B) * L;
    launch_pdl(poe_fwd_kernel<1, kMaxExperts>, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, a);
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_poe_bwd(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                            const uint32_t* pass_masks, int P, int B, int L, int variant, int training,
                            const float* noise, const float* dz, int64_t lddz, const float* dmu_up,
                            const float* dlv_up, float kl_scale, const float* kl_scale_dev, float* const* dmu_e,
                            float* const* dlv_e, int64_t ldd_e, void* stream) {
  return mvae_poe_bwd_g(mu_e, lv_e, ld_e, E, nullptr, pass_masks, P, B, L, variant, training, noise, dz, lddz, dmu_up, dlv_up,
                        kl_scale, kl_scale_dev, dmu_e, dlv_e, ldd_e, stream);
}

extern "C" int mvae_poe_bwd_g(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                              const int64_t* const* gather_idx, const uint32_t* pass_masks, int P, int B, int L,
                              int variant, int training, const float* noise, const float* dz, int64_t lddz,
                              const float* dmu_up, const float* dlv_up, float kl_scale, const float* kl_scale_dev,
                              float* const* dmu_e, float* const* dlv_e, int64_t ldd_e, void* stream) {
  PoeArgs a = {};
  int rc = fill_poe_args(a, mu_e, lv_e, ld_e, E, pass_masks, P, B, L, variant, training, gather_idx);
  if (rc) return rc;
  if (!dz || !dmu_e || !dlv_e) return set_error(MVAE_ERR_BAD_ARG, "poe_bwd: NULL dz/dmu_e/dlv_e");
  if (training && !noise) return set_error(MVAE_ERR_BAD_ARG, "poe_bwd: training needs the forward noise");
  for (int e = 0; e < E; ++e) {
    if (!dmu_e[e] || !dlv_e[e]) return set_error(MVAE_ERR_BAD_ARG, "poe_bwd: expert %d grad pointer is NULL", e);
    a.dmu_e[e] = dmu_e[e]; a.dlv_e[e] = dlv_e[e];
  }
  a.noise = noise; a.dz = dz; a.ldz = lddz; a.kl_scale = kl_scale; a.kl_scale_dev = kl_scale_dev; a.ldd_e = ldd_e;
  a.dmu_up = dmu_up; a.dlv_up = dlv_up;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (vec4_ok(a, true) && P <= kFastPasses && !getenv("MVAE_POE_SLOW")) {
    const int64_t n = static_cast<int64_t>(B) * (L / 4);
    if (E <= 2) launch_pdl(poe_bwd_fast_kernel<2>, dim3(fast_grid(n)), dim3(256), 0, st, a);
    else launch_pdl(poe_bwd_fast_kernel<4>, dim3(fast_grid(n)), dim3(256), 0, st, a);
  } else if (vec4_ok(a, true)) {
    const int64_t n = static_cast<int64_t>(B) * (L / 4);
    launch_pdl(poe_bwd_kernel<4, 4>, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, a);
  } else {
    const int64_t n = static_cast<int64_t>(B) * L;
    launch_pdl(poe_bwd_kernel<1, kMaxExperts>, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, a);
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_reparam_fwd(const float* mu, const float* logvar, const float* noise, float* noise_out,
                                uint64_t seed, uint64_t offset, float* z, int64_t n, void* stream) {
  if (!mu || !logvar || !z || n < 1 || (!noise && !noise_out)) return set_error(MVAE_ERR_BAD_ARG, "reparam_fwd: bad args");
  reparam_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mu, logvar, noise, noise_out, seed, offset, z, n);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_reparam_bwd(const float* logvar, const float* noise, const float* dz, float* dlogvar, int64_t n,
                                void* stream) {
  if (!logvar || !noise || !dz || !dlogvar || n < 1) return set_error(MVAE_ERR_BAD_ARG, "reparam_bwd: bad args");
  reparam_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logvar, noise, dz, dlogvar, n);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_kl_fwd_bwd(const float* mu, const float* logvar, float* dmu, float* dlogvar, int64_t n,
                               float scale, double* kl_acc, void* stream) {
  if (!mu || !logvar || n < 1) return set_error(MVAE_ERR_BAD_ARG, "kl: bad args");
  kl_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mu, logvar, dmu, dlogvar, n, scale, kl_acc);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_bce_logits_fwd_bwd(const float* x, int64_t ldx, const float* t, int64_t ldt, int t_rows,
                                       float* dx, int64_t lddx, int R, int D, float scale, double* loss_acc,
                                       int seg_rows, float* loss_elem, int64_t ldl, void* stream) {
  if (!x || !t || R < 1 || D < 1 || t_rows < 1) return set_error(MVAE_ERR_BAD_ARG, "bce: bad pointers/shape");
  if (seg_rows < 1) seg_rows = R;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((D & 3) || (ldx & 3) || (ldt & 3) || (dx && (lddx & 3)) || !al(x) || !al(t) || !al(dx) ||
      (loss_elem && ((ldl & 3) || !al(loss_elem)))) {
    if (D > 4096) return set_error(MVAE_ERR_UNSUPPORTED, "bce: wide rows need D, ld multiples of 4 and 16B aligned pointers");
    bce_rows_kernel<<<(R + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, t, ldt, t_rows, dx, lddx, R, D, scale, loss_acc, seg_rows, loss_elem, ldl);
    count_launch();
    MVAE_CUDA_CHECK(cudaGetLastError());
    return MVAE_OK;
  }
  // stacked passes sharing one target (seg_rows == t_rows, R = copies * t_rows): read the target once
  if (R % t_rows == 0 && R / t_rows >= 2 && R / t_rows <= 4 && seg_rows == t_rows && loss_elem == nullptr &&
      static_cast<int64_t>(R) * (D / 4) < (int64_t(1) << 31) && ldx < (int64_t(1) << 31) && !getenv("MVAE_BCE_NO_STACK")) {
    const int copies = R / t_rows;
    const int64_t n4t = static_cast<int64_t>(t_rows) * (D / 4);
    const int sms = mvae_device_sm_count() > 0 ? mvae_device_sm_count() : 148;
    int64_t blocks = (n4t + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sms) * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned gb = static_cast<unsigned>(blocks);
    if (copies == 2)
      launch_pdl(bce_stacked_kernel<2>, dim3(gb), dim3(256), 0, st, x, (int)ldx, t, (int)ldt, t_rows, dx, (int)lddx, D / 4, scale, loss_acc);
    else if (copies == 3)
      launch_pdl(bce_stacked_kernel<3>, dim3(gb), dim3(256), 0, st, x, (int)ldx, t, (int)ldt, t_rows, dx, (int)lddx, D / 4, scale, loss_acc);
    else
      launch_pdl(bce_stacked_kernel<4>, dim3(gb), dim3(256), 0, st, x, (int)ldx, t, (int)ldt, t_rows, dx, (int)lddx, D / 4, scale, loss_acc);
    count_launch();
    MVAE_CUDA_CHECK(cudaGetLastError());
    return MVAE_OK;
  }
  const int64_t n4 = static_cast<int64_t>(R) * (D / 4);
  if (n4 >= (int64_t(1) << 31) || ldx >= (int64_t(1) << 31) || ldt >= (int64_t(1) << 31) || lddx >= (int64_t(1) << 31) ||
      ldl >= (int64_t(1) << 31))
    return set_error(MVAE_ERR_UNSUPPORTED, "bce: tensor too large for 32-bit indexing");
  static const int unroll = getenv("MVAE_BCE_UNROLL") ? atoi(getenv("MVAE_BCE_UNROLL")) : 4;
  const int per_block = 256 * (unroll == 8 ? 8 : 4);
  const int64_t nslabs = (n4 + per_block - 1) / per_block;
  const int sms = mvae_device_sm_count() > 0 ? mvae_device_sm_count() : 148;
  int slabs = static_cast<int>(nslabs / (static_cast<int64_t>(sms) * 8));   // aim for >= 8 blocks per SM
  slabs = slabs < 1 ? 1 : (slabs > 16 ? 16 : slabs);
  const int64_t blocks = (nslabs + slabs - 1) / slabs;
  if (unroll == 8)
    launch_pdl(bce_kernel<8>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
        x, static_cast<int>(ldx), t, static_cast<int>(ldt), t_rows, dx, static_cast<int>(lddx), R, D / 4, scale, loss_acc,
        seg_rows, loss_elem, static_cast<int>(ldl), slabs);
  else
    launch_pdl(bce_kernel<4>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
        x, static_cast<int>(ldx), t, static_cast<int>(ldt), t_rows, dx, static_cast<int>(lddx), R, D / 4, scale, loss_acc,
        seg_rows, loss_elem, static_cast<int>(ldl), slabs);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_ce_fwd_bwd(const float* x, int64_t ldx, const int64_t* target, int t_rows, float* dx,
                               int64_t lddx, int R, int K, float scale, double* loss_acc, int seg_rows,
                               float* loss_rows, int64_t ldl, void* stream) {
  if (!x || !target || R < 1 || K < 1 || t_rows < 1) return set_error(MVAE_ERR_BAD_ARG, "ce: bad pointers/shape");
  if (seg_rows < 1) seg_rows = R;
  launch_pdl(ce_kernel, dim3((R + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, ldx, target, t_rows, dx, lddx, R,
                                                                                  K, scale, loss_acc, seg_rows, loss_rows, ldl);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_colsum_accumulate(const float* dy, int64_t lddy, float* db, int M, int N, void* stream) {
  if (!dy || !db || M < 1 || N < 1) return set_error(MVAE_ERR_BAD_ARG, "colsum: bad pointers/shape");
  dim3 grid((N + 31) / 32, (M + kColsumRows - 1) / kColsumRows);
  launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), dy, lddy, db, M, N);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_swish_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (!x || !y || n < 1) return set_error(MVAE_ERR_BAD_ARG, "swish_fwd: bad args");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)
    return set_error(MVAE_ERR_UNSUPPORTED, "swish_fwd: pointers must be 16B aligned");
  const int64_t n4 = (n + 3) / 4;
  swish_fwd_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_swish_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
  if (!x || !dy || !dx || n < 1) return set_error(MVAE_ERR_BAD_ARG, "swish_bwd: bad args");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15)
    return set_error(MVAE_ERR_UNSUPPORTED, "swish_bwd: pointers must be 16B aligned");
  const int64_t n4 = (n + 3) / 4;
  swish_bwd_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, dy, dx, n);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_embedding_swish_fwd(const float* table, const int64_t* idx, float* a, float* h, int B, int D,
                                        int V, void* stream) {
  if (!table || !idx || !h || B < 1 || D < 4 || (D & 3) || V < 1)
    return set_error(MVAE_ERR_BAD_ARG, "embedding_swish_fwd: bad args (D must be a multiple of 4)");
  const int64_t n = static_cast<int64_t>(B) * (D / 4);
  emb_swish_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, idx, a, h, B, D / 4);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
extern "C" int mvae_embedding_swish_bwd(const float* table, const int64_t* idx, const float* dh, int64_t lddh,
                                        float* dtable, int B, int D, int V, void* stream) {
  if (!table || !idx || !dh || !dtable || B < 1 || D < 1 || V < 1)
    return set_error(MVAE_ERR_BAD_ARG, "embedding_swish_bwd: bad args");
  if (V > kEmbMaxV) return set_error(MVAE_ERR_UNSUPPORTED, "embedding_swish_bwd: V=%d > %d", V, kEmbMaxV);
  dim3 grid((D + 127) / 128, (B + kEmbRows - 1) / kEmbRows);
  emb_swish_bwd_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table, idx, dh, lddh, dtable, B, D, V);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                              const float* lr_mult_dev, float beta1, float beta2, float eps, float grad_scale,
                              int32_t* step_count, void* stream) {
  if (!p || !g || !m || !v || !step_count || n < 1) return set_error(MVAE_ERR_BAD_ARG, "adam: bad args");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return set_error(MVAE_ERR_UNSUPPORTED, "adam: buffers must be 16B aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n4 = (n + 3) / 4;
  launch_pdl(adam_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, st, p, g, m, v, n, lr, lr_mult_dev, beta1, beta2, eps,
                                                                        grad_scale, step_count);
  launch_pdl(step_inc_kernel, dim3(1), dim3(1), 0, st, step_count);
  count_launch(2);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

namespace mvae {
namespace {
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  float4 l;
  l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo[i] = l;
}
}  // namespace
}  // namespace mvae

extern "C" int mvae_split_lo(const float* x, float* lo, int64_t n, void* stream) {
  if (!x || !lo || n < 4 || (n & 3) || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15))
    return set_error(MVAE_ERR_BAD_ARG, "split_lo: n %% 4 == 0 and 16-byte aligned buffers required");
  const int64_t n4 = n / 4;
  split_lo_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(lo), n4);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

extern "C" int mvae_elbo_finalize(const double* recon_img, const double* recon_txt, const double* kl, int P,
                                  float lambda_image, float lambda_text, float beta, const float* beta_dev,
                                  float inv_batch, float* out, void* stream) {
  if (!out || P < 1 || P > kMaxPasses) return set_error(MVAE_ERR_BAD_ARG, "elbo_finalize: bad args");
  launch_pdl(elbo_finalize_kernel, dim3(1), dim3(32), 0, reinterpret_cast<cudaStream_t>(stream), recon_img, recon_txt, kl, P, lambda_image,
                                                                              lambda_text, beta, beta_dev, inv_batch, out);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
