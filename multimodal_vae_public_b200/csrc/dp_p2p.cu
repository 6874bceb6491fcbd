// Data-parallel exchange step as ONE kernel per rank over NVLink peer memory:
//   gradient reduce-scatter (peer loads) -> Adam on this rank's 1/N slice -> parameter all-gather (peer stores).
//
// Replaces ncclAllReduce(gradient bucket) + the flat Adam kernel of the data-parallel step (the reference has no
// multi-GPU path at all: mnist/train.py:196-219 is one process; the exchange step is SURVEY.md section 8e).  Every rank
// launches this kernel on its own stream after its backward pass; the ranks meet inside the kernel through flags that
// live in peer-mapped memory.
//
//   barrier 1  every rank tells every peer "my gradients are final" (one system-scope store per peer) and waits for
//              the same word from every peer (local polling, the peers write into MY flag array);
//   slice      rank r owns floats [r*chunk, (r+1)*chunk): g = sum over peers of their gradient slice (128-bit peer
//              loads), Adam with this rank's m / v slice, the new parameter values are stored into EVERY rank's
//              parameter buffer (128-bit peer stores) -- each rank moves n*(N-1)/N floats in and out, the NVLink-optimal
//              volume, and does 1/N of the optimizer arithmetic;
//   tail       the T loss scalars that ride behind the gradients are summed by every rank for itself;
//   barrier 2  the last block of every rank tells every peer "my stores into your parameters are done" and waits for
//              all peers before the kernel ends: when the kernel is over, the local parameters are complete and nobody
//              reads this rank's gradients any more (the next step zeroes them).
//
// Flags carry an explicit 1-based exchange number (the "epoch") kept in the flag array itself (word 2N+2, advanced by the
// kernel), so a captured CUDA graph replays unchanged and the numbering survives the caller resetting its Adam step
// counter.  The protocol keeps the ranks in lockstep by construction (nobody can finish exchange e before every peer has
// entered it), so a flag that ends a wait must EQUAL the local epoch; anything else means the ranks did not launch the
// same sequence of exchanges (e.g. a rank ran an extra warm-up step) and is reported instead of being papered over.
// Failure behaviour (sticky error word 2N: 1 = a peer did not arrive within ~9 s, 2 = a peer is at a different epoch):
// the rank that sees it skips its Adam update and all its parameter stores, every later launch does the same, and the
// host raises at its next synchronisation point (trainer.check_device_errors) -- no silent training on partial sums.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"
#include "common.h"

namespace mvae {
namespace {

constexpr int kMaxWorld = 16;
constexpr long long kSpinTimeout = 1LL << 34;   // cycles (~9 s)

struct P2PArgs {
  float* grads[kMaxWorld];        // peer-mapped gradient buckets (n + tail floats), rank order
  float* params[kMaxWorld];       // peer-mapped parameter buckets (n floats)
  uint32_t* flags[kMaxWorld];     // peer-mapped flag arrays: [0, N) barrier 1, [N, 2N) barrier 2, [2N] error, [2N+1] done
                                  // counter, [2N+2] launches so far (the flag epoch)
  float* m;
  float* v;
  float* tail_out;
  const float* lr_mult_dev;
  int32_t* step_count;
  int64_t n;
  int64_t chunk;                  // floats per rank slice (multiple of 4)
  int tail, rank, world;
  float lr, beta1, beta2, eps;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer gradient loads: read-once, straight from the owner's memory
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// A rank that gives up tells its peers so (both of its flag slots in every peer's array get an impossible exchange
// number): they fail fast with "mismatch" instead of each waiting for its own timeout.
constexpr uint32_t kPoison = 0xFFFFFFFFu;
__device__ __forceinline__ void poison_peers(const P2PArgs& a) {
  for (int p = 0; p < a.world; ++p) {
    if (p == a.rank) continue;
    st_release_sys(a.flags[p] + a.rank, kPoison);
    st_release_sys(a.flags[p] + a.world + a.rank, kPoison);
  }
}

// Wait until every peer wrote `epoch` into my flag words [base, base + N).  Returns 0, or the error code that was also
// recorded in the sticky error word: 1 = timeout, 2 = a peer's flag carries a different exchange number.
__device__ __forceinline__ uint32_t wait_all(const P2PArgs& a, int base, uint32_t epoch) {
  uint32_t* mine = a.flags[a.rank];
  const long long t0 = clock64();
  for (int p = 0; p < a.world; ++p) {
    uint32_t v;
    while ((v = ld_acquire_sys(mine + base + p)) < epoch) {
      __nanosleep(100);
      if (*reinterpret_cast<volatile uint32_t*>(mine + 2 * a.world) != 0u) return 1u;   // another block already gave up
      if (clock64() - t0 > kSpinTimeout) {
        if (atomicCAS(mine + 2 * a.world, 0u, 1u) == 0u) poison_peers(a);
        return 1u;
      }
    }
    if (v != epoch) {
      if (atomicCAS(mine + 2 * a.world, 0u, 2u) == 0u) poison_peers(a);
      return 2u;
    }
  }
  return 0u;
}

__global__ void __launch_bounds__(256) allreduce_adam_p2p_kernel(const P2PArgs a) {
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  __shared__ int s_last;
  __shared__ uint32_t s_err;
  const int N = a.world;
  const uint32_t epoch = a.flags[a.rank][2 * N + 2] + 1u;               // same on every rank: all launch in lockstep
  const int adam_t = *a.step_count + 1;
  const uint32_t sticky = *reinterpret_cast<volatile uint32_t*>(a.flags[a.rank] + 2 * N);   // an earlier exchange failed
  // ---- barrier 1: my gradients are final (this kernel is stream-ordered after my backward pass)
  if (blockIdx.x == 0 && threadIdx.x < N) st_release_sys(a.flags[threadIdx.x] + a.rank, epoch);
  if (threadIdx.x == 0) {
    const float lr_eff = a.lr * (a.lr_mult_dev ? __ldg(a.lr_mult_dev) : 1.0f);
    const double bc1 = 1.0 - pow(static_cast<double>(a.beta1), adam_t);
    const double bc2 = 1.0 - pow(static_cast<double>(a.beta2), adam_t);
    s_step_size = static_cast<float>(lr_eff / bc1);
    s_inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
    s_err = sticky != 0u ? sticky : wait_all(a, 0, epoch);          // every block polls its own (local) copy of the flags
  }
  __syncthreads();
  const float step_size = s_step_size, inv_sqrt_bc2 = s_inv_sqrt_bc2;
  const bool ok = s_err == 0u;
  // ---- my slice: reduce over peers, Adam, broadcast (skipped entirely after a failed rendezvous)
  const int64_t begin = a.chunk * a.rank;
  const int64_t end = !ok ? begin : (begin + a.chunk < a.n ? begin + a.chunk : a.n);
  for (int64_t i = begin + (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < end;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
    // (the arena pads every tensor to a multiple of 4 floats and n itself is a multiple of 4)
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = 0; p < N; ++p) {
      const float4 t = ld_peer(a.grads[p] + i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 P = *reinterpret_cast<const float4*>(a.params[a.rank] + i);
    float4 M = *reinterpret_cast<const float4*>(a.m + i);
    float4 V = *reinterpret_cast<const float4*>(a.v + i);
    float* pp = &P.x; float* mm = &M.x; float* vv = &V.x; const float* gg = &g.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mm[q] = a.beta1 * mm[q] + (1.0f - a.beta1) * gg[q];
      vv[q] = a.beta2 * vv[q] + (1.0f - a.beta2) * gg[q] * gg[q];
      const float denom = sqrtf(vv[q]) * inv_sqrt_bc2 + a.eps;
      pp[q] -= step_size * (mm[q] / denom);
    }
    *reinterpret_cast<float4*>(a.m + i) = M;
    *reinterpret_cast<float4*>(a.v + i) = V;
    for (int p = 0; p < N; ++p) st_peer(a.params[p] + i, P);
  }
  // ---- loss scalars behind the gradients: every rank sums them for itself
  if (ok && blockIdx.x == 0 && threadIdx.x < a.tail) {
    float s = 0.f;
    for (int p = 0; p < N; ++p) {
      float t;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(t) : "l"(a.grads[p] + a.n + threadIdx.x) : "memory");
      s += t;
    }
    a.tail_out[threadIdx.x] = s;
  }
  // ---- barrier 2: my peer stores are done; the last block of this rank tells the peers and waits for theirs
  __threadfence_system();
  __syncthreads();
  uint32_t* mine = a.flags[a.rank];
  if (threadIdx.x == 0) s_last = atomicAdd(mine + 2 * N + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < N) st_release_sys(a.flags[threadIdx.x] + N + a.rank, epoch);
    if (threadIdx.x == 0) {
      const uint32_t err2 = sticky != 0u ? sticky : wait_all(a, N, epoch);
      mine[2 * N + 1] = 0u;                       // done counter ready for the next launch
      mine[2 * N + 2] = epoch;                    // (every block read both counters at its start)
      if (err2 == 0u && *reinterpret_cast<volatile uint32_t*>(mine + 2 * N) == 0u) *a.step_count = adam_t;
      __threadfence_system();
    }
  }
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_allreduce_adam_p2p(float* const* grad_ptrs, float* const* param_ptrs, uint32_t* const* flag_ptrs,
                                       float* m, float* v, int64_t n, int tail, float* tail_out, int rank, int world,
                                       float lr, const float* lr_mult_dev, float beta1, float beta2, float eps,
                                       int32_t* step_count, void* stream) {
  if (!grad_ptrs || !param_ptrs || !flag_ptrs || !m || !v || !step_count || n < 4 || (n & 3) || world < 1 ||
      world > kMaxWorld || rank < 0 || rank >= world || tail < 0 || tail > 32 || (tail > 0 && !tail_out))
    return set_error(MVAE_ERR_BAD_ARG, "allreduce_adam_p2p: bad arguments (n %% 4 == 0, world <= %d, tail <= 32)", kMaxWorld);
  P2PArgs a;
  for (int p = 0; p < world; ++p) {
    if (!grad_ptrs[p] || !param_ptrs[p] || !flag_ptrs[p] ||
        ((reinterpret_cast<uintptr_t>(grad_ptrs[p]) | reinterpret_cast<uintptr_t>(param_ptrs[p])) & 15))
      return set_error(MVAE_ERR_BAD_ARG, "allreduce_adam_p2p: peer pointer %d missing or not 16-byte aligned", p);
    a.grads[p] = grad_ptrs[p]; a.params[p] = param_ptrs[p]; a.flags[p] = flag_ptrs[p];
  }
  if ((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
    return set_error(MVAE_ERR_BAD_ARG, "allreduce_adam_p2p: m / v must be 16-byte aligned");
  a.m = m; a.v = v; a.tail_out = tail_out; a.lr_mult_dev = lr_mult_dev; a.step_count = step_count;
  a.n = n; a.tail = tail; a.rank = rank; a.world = world;
  a.chunk = ((n / 4 + world - 1) / world) * 4;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  int sms = mvae_device_sm_count();
  if (sms <= 0) return set_error(MVAE_ERR_CUDA, "no CUDA device");
  // every block spins on the flags, so the whole grid must be resident at once: 2 blocks of 256 threads per SM
  const int64_t work = (a.chunk / 4 + 255) / 256;
  int grid = 2 * sms;
  if (work < grid) grid = static_cast<int>(work < 1 ? 1 : work);
  allreduce_adam_p2p_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
