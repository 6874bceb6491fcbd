// Persistent, warp-specialised tcgen05 GEMM for sm_100a (B200).
//
//   C[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue), fp32 in HBM, tf32 tensor-core products, fp32
//   accumulation in TMEM.  A and B may each be K-major or MN-major in memory, so the same kernel serves
//   nn.Linear forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without any transposed copies.
//
// Structure (one CTA per SM, static round-robin tile scheduler over a small batch of problems):
//   warp 0      : TMA producer   -- cp.async.bulk.tensor 128B-swizzled boxes into a smem ring
//   warp 1      : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N<=128, K=8 per instr)
//   warps 2..9  : epilogue       -- tcgen05.ld TMEM -> registers -> smem transpose -> bias / Swish / dSwish /
//                                   bias-gradient column sums -> coalesced global stores (or red.add)
//   warps 10..13: operand split  -- (3xTF32 mode only) A -> hi|lo in TENSOR MEMORY (tcgen05.st), B -> lo tile in
//                                   smem (raw B is the hi operand: kind::tf32 truncates), then 3 MMAs/product:
//                                   lo*hi + hi*lo + hi*hi with A read from TMEM (fp32-class accuracy)
// Pipelines: smem full/ready/empty mbarrier ring; double-buffered TMEM accumulators (2 x 128 columns)
// so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Replaces: torch addmm/mm + F.sigmoid*x of the reference's Linear/Swish stacks
// (mnist/model.py:75-78,81-84,95-98,101-105,117-119,122-125,136-139,142-146,166-169) and their autograd.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mvae_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace mvae {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N_MAX = 128;
constexpr int BLOCK_K = 32;                                 // 32 fp32 = 128 B = one swizzle row
constexpr int UMMA_K = 8;                                   // tf32: 32 B per MMA along K
constexpr int OPERAND_BYTES = BLOCK_M * BLOCK_K * 4;        // 16 KiB per operand per stage
constexpr int ACC_COLS = 256;                               // 2 accumulator stages x 128 fp32 columns
constexpr int A_STAGE_COLS = 64;                            // 3xTF32: A operand hi (32 cols) | lo (32 cols) per stage
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_SPLIT_WARPS = 4;
constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * 32 * 4;  // one swizzled 32x32 fp32 staging tile per epilogue warp
constexpr int MAX_PROBLEMS = MVAE_GEMM_MAX_CHAIN;             // problem slots of one launch (batch uses the first 4)
// Chain workspace (int32, caller-owned, zero-initialised once; the kernel leaves the counters zeroed again):
//   ws[0] = CTAs that finished, ws[1] = sticky error flag (a dependency wait timed out), ws[2 + i] = completion
//   counter of row block i (all problems of the chain, concatenated).
constexpr int WS_DONE = 0, WS_ERR = 1, WS_CTR0 = MVAE_GEMM_CHAIN_WS_HEADER;
constexpr long long DEP_TIMEOUT_CYCLES = 1LL << 30;           // ~0.5 s: a broken chain fails loudly instead of hanging

template <bool kSplit, bool kPair = false>
struct Cfg {
  // tf32     : [A][B]                        32 KiB x 6 stages
  // 3x       : [A raw][B raw][B lo]          48 KiB x 4 stages; the A operand is re-staged as hi|lo in TENSOR MEMORY
  // 3x, pair : [A raw][B raw half][B lo half] 32 KiB x 4 stages: a CTA pair (cta_group::2) works on a 256 x block_n tile,
  //            each CTA holds its own 128 rows of A and HALF of the B tile; the tensor cores of the two SMs exchange
  //            the B halves, so every SM reads (and splits) half as many B bytes from its shared memory per product
  static constexpr int kBBytes = kPair ? OPERAND_BYTES / 2 : OPERAND_BYTES;           // B tile bytes held by one CTA
  static constexpr int kStageBytes = kSplit ? OPERAND_BYTES + 2 * kBBytes : 2 * OPERAND_BYTES;
  // Shared-memory ring depth.  The 3x modes also stage A in tensor memory (kTmemStages slots of 64 columns next to the
  // 2 x 128 accumulator columns = all 512).  In pair mode the smem ring is DEEPER than the TMEM ring: the TMA runs up to
  // 6 k-blocks ahead (covers the L2 -> smem latency), the splitters up to 4 ahead of the tensor core; measured with
  // 4/4 the pair main loop was round-trip-latency bound at ~1300 cycles per k-block instead of the tensor pipe's 768.
  static constexpr int kStages = kSplit ? (kPair ? 6 : 4) : 6;
  static constexpr int kTmemStages = 4;
  static constexpr int kTmemCols = kSplit ? 512 : 256;
  static constexpr int kThreads = 32 * (2 + NUM_EPI_WARPS + (kSplit ? NUM_SPLIT_WARPS : 0));
  static constexpr int kSmemBytes = kStages * kStageBytes + EPI_STAGE_BYTES + 1024;  // + 1024 B alignment slack
  static constexpr int kTileM = kPair ? 2 * BLOCK_M : BLOCK_M;   // rows of one scheduled tile
};

struct alignas(64) GemmProblem {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_b_lo;   // 3xTF32 with a pre-split B (weights): the lo tile arrives by TMA instead of being computed in the loop
  int has_b_lo;
  float* C;
  const float* bias;
  const float* aux;
  float* out2;
  float* colsum;          // optional [N]: += column sums of the stored C tile (bias gradient)
  int64_t ldc, ldaux, ldout2;
  int M, N, K;
  int block_n;            // MMA N (multiple of 16, <= 128)
  int tiles_m, tiles_n;   // output tiles
  int split_k;            // k-range splits
  int num_kblocks;        // ceil(K / BLOCK_K)
  int kblocks_per_split;
  int tile_begin;         // first global tile index of this problem
  int a_mn, b_mn;         // operand majorness
  // shared-memory matrix-descriptor parameters per operand (bytes; layout = UMMA LayoutType code)
  uint32_t a_lbo, a_sbo, a_kstep, a_layout;
  uint32_t b_lbo, b_sbo, b_kstep, b_layout;
  int epilogue;
  int atomic;             // accumulate with red.add instead of st
  // ---- chain mode (mvae_gemm_chain): the A operand of this problem is the C/out2 of problem `dep` of the same launch
  int dep;                // producer problem index or -1
  int dep_ctr_base;       // first counter of the producer's row blocks (+ the row-block offset of A inside it)
  int dep_target;         // counter value of a complete producer row block: 8 epilogue warps x tiles_n x split_k
  int ctr_base;           // first counter of this problem's row blocks
  int row_blocks;         // ceil(M / 128): counters of this problem
  int dep_row_blocks;     // counters of the producer from dep_ctr_base on (row blocks beyond do not exist)
  // ---- fused split-K ("last arriver"): the k range of an output tile is split over split_k CTAs whose partial
  // accumulators are red.add'ed into a zeroed fp32 scratch matrix; per (tile, epilogue warp) an arrival counter tells the
  // last of them that its 32-row slab is complete -- it reads the sums back, applies the real epilogue (bias / Swish /
  // Swish' / column sums), stores C / out2, re-zeroes the scratch and publishes the row block to the chain.
  // ---- implicit-GEMM convolution operands (mvae_conv_view): the operand is fetched with TMA im2col-mode loads from the
  // NHWC activation itself.  Per operand: cblocks = C / 32 (k-blocks, or 32-wide boxes, per filter tap; 0 = plain matrix),
  // output raster OW / OH*OW, position of the first filter tap of output pixel (p, q): (lower + p*stride, lower + q*stride).
  // Divisions by OW, OH*OW, C, cblocks, taps_w run on the single TMA-producer thread once per k-block: they use
  // host-computed reciprocals (x / d = (x * mul) >> 40, exact for x < 2^28, d < 2^12).
  struct ConvGeom { int cblocks, C, OW, OHW, lower_h, lower_w, stride, taps_w;
                    unsigned long long m_OW, m_OHW, m_C, m_cb, m_tw; } a_cv, b_cv;
  // ---- tap-split B (sub-pixel transposed convolutions): the K axis of the problem is (tap slot t, 32 r), the matching
  // B rows (K-major) / columns (MN-major) of slot t start at bt_table[t] * bt_mn; bt_cb = k-blocks per slot (0 = off)
  int bt_cb, bt_mn;
  unsigned long long bt_m_cb;
  unsigned char bt_table[16];
  // ---- output row map (the same): GEMM row (n, j, i) of an IH x IW grid is stored at NHWC pixel (n, sy j + py, sx i + px)
  // of an OH x OW image: row' = (n OH + s j + py) OW + s i + px.  Applies to C, out2 and aux.  rm_IW = 0: identity.
  int rm_IW, rm_IHW, rm_OW, rm_OHW, rm_s, rm_py, rm_px;
  unsigned long long rm_m_IW, rm_m_IHW;
  float* split_ws;        // [M][ldp] scratch or nullptr (plain split-K: partial sums red.add'ed straight into C)
  int64_t ldp;
  int tctr_base;          // first arrival counter of this problem (tiles_m * tiles_n * 8 counters)
};

// TMA base coordinates (w, h, n) of output pixel `pix` (raster index over n, p, q) of an implicit conv operand
struct ConvPos { int w, h, n; };
__device__ __forceinline__ int fdiv(int x, unsigned long long mul) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(x)) * mul) >> 40);
}
__device__ __forceinline__ ConvPos conv_pos(const GemmProblem::ConvGeom& g, int pix) {
  ConvPos r;
  r.n = fdiv(pix, g.m_OHW);
  const int rem = pix - r.n * g.OHW;
  const int pp = fdiv(rem, g.m_OW);
  r.h = g.lower_h + pp * g.stride;
  r.w = g.lower_w + (rem - pp * g.OW) * g.stride;
  return r;
}
// (channel offset, tap column, tap row) of index `idx` = (tap, channel) of an implicit operand's (tap, channel) axis
struct ConvTap { int c0; uint16_t tw, th; };
__device__ __forceinline__ ConvTap conv_tap(const GemmProblem::ConvGeom& g, int idx) {
  ConvTap t;
  const int tap = fdiv(idx, g.m_C);
  t.c0 = idx - tap * g.C;
  const int th = fdiv(tap, g.m_tw);
  t.th = static_cast<uint16_t>(th);
  t.tw = static_cast<uint16_t>(tap - th * g.taps_w);
  return t;
}

struct GemmBatch {
  GemmProblem p[MAX_PROBLEMS];
  int num_problems;
  int total_tiles;
  int* ws;          // chain workspace or nullptr (independent problems: no signalling, no waiting)
  int num_ctrs;
  long long* dbg;   // optional timeline buffer (MVAE_DBG_TIMELINE): [block < 8][role < 6][64] clock64 stamps
  int dbg_flags;    // MVAE_DBG_EPI: 1 = skip global stores, 2 = skip sigmoid math, 4 = skip smem transpose
  long long* tilelog;  // optional per-tile log (MVAE_DBG_TILELOG): [cta][TILELOG_MAX][4] = {tile, t_begin, t_deps_ready,
                       // t_epilogue_done} in globaltimer ns -- a Gantt chart of a (chained) launch
};

constexpr int TILELOG_MAX = 64;
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tilelog(const GemmBatch& b, int slot, int field, long long v) {
  if (b.tilelog != nullptr && slot < TILELOG_MAX) b.tilelog[(static_cast<long long>(blockIdx.x) * TILELOG_MAX + slot) * 4 + field] = v;
}

__device__ __forceinline__ void dbg_stamp(const GemmBatch& b, int role, int& n) {
  if (b.dbg != nullptr && blockIdx.x < 8 && n < 64) {
    b.dbg[(blockIdx.x * 6 + role) * 64 + n] = clock64();
    ++n;
  }
}

struct TileInfo {
  int prob, m_blk, n_blk, kb_begin, kb_end;
};

__device__ __forceinline__ TileInfo decode_tile(const GemmBatch& b, int t) {
  TileInfo ti;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < MAX_PROBLEMS; ++i)
    if (i < b.num_problems && t >= b.p[i].tile_begin) pi = i;
  const GemmProblem& p = b.p[pi];
  int local = t - p.tile_begin;
  int per_split = p.tiles_m * p.tiles_n;
  int split = local / per_split;
  int rem = local - split * per_split;
  ti.prob = pi;
  ti.m_blk = rem / p.tiles_n;
  ti.n_blk = rem - ti.m_blk * p.tiles_n;
  ti.kb_begin = split * p.kblocks_per_split;
  int e = ti.kb_begin + p.kblocks_per_split;
  ti.kb_end = e < p.num_kblocks ? e : p.num_kblocks;
  return ti;
}

// Shared-memory matrix descriptor (sm_100 "version 1").
//  K-major  (layout 2 = SWIZZLE_128B, TMA SWIZZLE_128B): rows of 128 B (32 fp32 of K); 8-row swizzle
//           atoms stacked at SBO = 1024 B; LBO unused; +32 B start address per K=8 slice.
//  MN-major (layout 1 = SWIZZLE_128B_BASE32B, TMA SWIZZLE_128B_ATOM_32B -- the only MN-major layout the
//           tensor core accepts for 32-bit operands): rows of 128 B (32 fp32 of M/N), one row per k;
//           swizzle atoms of 4 k-rows (512 B) stacked at SBO = 512 B; next 32-wide MN chunk (next TMA
//           box) at LBO = 4096 B; +1024 B start address per K=8 slice.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}

// ---- chain mode: row-block completion counters in global memory (release/acquire at gpu scope)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Wait until row block counter `ctr` reaches `target` (the producer tiles covering those rows have stored their
// results).  Bounded: after DEP_TIMEOUT_CYCLES the sticky error flag is raised and every later wait returns at once,
// so a scheduling bug shows up as a wrong result + error status, never as a hung GPU.
__device__ __forceinline__ void wait_row_block(const int* ctr, int target, int* ws) {
  if (ld_acquire_gpu(ctr) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(64);
    if (*reinterpret_cast<volatile int*>(ws + WS_ERR) != 0) return;
    if (clock64() - t0 > DEP_TIMEOUT_CYCLES) {
      atomicExch(ws + WS_ERR, 1);
      return;
    }
  }
}

// Epilogue sigmoid: ex2.approx + rcp.approx (5 instructions; max relative error ~ (2 + |x|) * 2^-23).  The epilogue
// runs on 8 warps next to a tensor-core main loop of ~9 K cycles per tile, so instructions per element are the
// budget: the IEEE expf / correctly rounded reciprocal version cost ~45 instructions per element and made the
// epilogue (not the MMA) the critical path (profiles/r01_epilogue_ablation.txt).
__device__ __forceinline__ float sigmoidf_acc(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));   // exp(-x); 0 / +inf at the ends
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));                   // rcp(+inf) = 0
  return r;
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Per-tile epilogue parameters, copied into registers once per tile (the problem table lives in kernel-parameter
// space and is indexed dynamically; re-reading it inside the element loops is slow).
struct EpiParams {
  float* C;
  const float* bias;
  const float* aux;
  float* out2;
  float* colsum;
  int64_t ldc, ldaux, ldout2;
  int M, N;
  int epilogue, atomic;
};

// Output row map of the sub-pixel transposed-convolution problems (GemmProblem::rm_*): only the general epilogue path of
// the kExtra kernel instantiation uses it.
struct RowMap {
  int IW, IHW, OW, OHW, s, py, px;
  unsigned long long m_IW, m_IHW;
};
__device__ __forceinline__ int64_t map_row(const RowMap& m, int row) {
  if (m.IW == 0) return row;
  const int n = static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(row)) * m.m_IHW) >> 40);
  const int rem = row - n * m.IHW;
  const int j = static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(rem)) * m.m_IW) >> 40);
  const int i = rem - j * m.IW;
  return static_cast<int64_t>(n) * m.OHW + static_cast<int64_t>(m.s * j + m.py) * m.OW + m.s * i + m.px;
}

template <bool kMap = false>
__device__ __forceinline__ float4 load_aux4(const EpiParams& e, int row, int col, int nvalid, const RowMap* rm = nullptr) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < e.M && nvalid > 0) {
    const float* arow = e.aux + (kMap ? map_row(*rm, row) : static_cast<int64_t>(row)) * e.ldaux + col;
    if (nvalid == 4) a = *reinterpret_cast<const float4*>(arow);
    else {
      a.x = arow[0];
      if (nvalid > 1) a.y = arow[1];
      if (nvalid > 2) a.z = arow[2];
    }
  }
  return a;
}

// One 32x32 accumulator chunk of one epilogue warp.  `st4` is the warp's 4 KiB staging tile (32 rows x 8 float4,
// float4 index XOR-swizzled with (row & 7): conflict-free for the row-wise writes and the transposed reads).
// The row loop is deliberately NOT unrolled: the fully unrolled version was instruction-fetch bound (the epilogue
// is executed once per tile, ~25 KiB of straight-line code per pass, see profiles/r01_epilogue_ablation.txt).
template <bool kMap = false>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& e, float4* st4, const uint32_t (&r)[32], int ncols,
                                               int lane, int row_base, int col0, const float (&bv)[4], float4 a_next,
                                               int dbg_flags, const RowMap* rm = nullptr) {
  const int sub = lane >> 3, cq = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (4 * j < ncols)
      st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  const int col = col0 + 4 * cq;
  int nvalid = e.N - col;                       // valid columns of this lane's float4 (0..4)
  nvalid = (4 * cq < ncols) ? (nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid)) : 0;
  const bool dsw = e.epilogue == MVAE_EPI_MUL_DSWISH;
  const bool bsw = e.epilogue == MVAE_EPI_BIAS_SWISH;
  const bool want_cs = e.colsum != nullptr;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + sub;
    const int row = row_base + rl;
    const float4 a_cur = a_next;
    if (dsw && i < 7) a_next = load_aux4<kMap>(e, row + 4, col, nvalid, rm);   // one row group ahead
    if (nvalid == 0 || row >= e.M) continue;
    const int64_t orow = kMap ? map_row(*rm, row) : static_cast<int64_t>(row);
    const float4 s4 = st4[rl * 8 + (cq ^ (rl & 7))];
    float v[4] = {s4.x + bv[0], s4.y + bv[1], s4.z + bv[2], s4.w + bv[3]};
    if (dsw && !(dbg_flags & 2)) {
      const float a[4] = {a_cur.x, a_cur.y, a_cur.z, a_cur.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sg = sigmoidf_acc(a[q]);
        v[q] *= sg * (1.0f + a[q] * (1.0f - sg));
      }
    }
    if (want_cs) {
#pragma unroll
      for (int q = 0; q < 4; ++q) cs[q] += (q < nvalid) ? v[q] : 0.f;
    }
    float* cptr = e.C + orow * e.ldc + col;
    if (dbg_flags & 1) {
      if (v[0] == 123.456f) *cptr = v[1] + v[2] + v[3];
    } else if (e.atomic) {
      if (nvalid == 4) ptx::red_add_v4_f32(cptr, v[0], v[1], v[2], v[3]);
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nvalid) ptx::red_add_f32(cptr + q, v[q]);
      }
    } else {
      if (nvalid == 4) *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nvalid) cptr[q] = v[q];
      }
    }
    if (bsw) {
      float* hptr = e.out2 + orow * e.ldout2 + col;
      float h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = (dbg_flags & 2) ? v[q] * 0.5f : v[q] * sigmoidf_acc(v[q]);
      if (dbg_flags & 1) {
        if (h[0] == 123.456f) *hptr = h[1] + h[2] + h[3];
      } else if (nvalid == 4) *reinterpret_cast<float4*>(hptr) = make_float4(h[0], h[1], h[2], h[3]);
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nvalid) hptr[q] = h[q];
      }
    }
  }
  if (want_cs) {
    // bias gradient: column sums of the stored tile (32 rows of this warp) -> one red.add per column
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 8);
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 16);
    }
    if (sub == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < nvalid) ptx::red_add_f32(e.colsum + col + q, cs[q]);
    }
  }
  __syncwarp();  // the staging tile is rewritten by the next chunk
}

// Interior-tile fast path of the above: the 32x32 chunk lies completely inside C (all rows < M, all columns < N), so
// there are no bounds predicates; the epilogue kind / atomic / column-sum choices are template parameters; row pointers
// advance by a constant stride; the staging tile is addressed in the shared window (ld/st.shared, not generic); and the
// aux rows of the whole chunk (8 x 128-bit per lane) were requested before the accumulator was ready.  ~3x fewer
// instructions per element than the general path: the epilogue warps share their schedulers with the operand
// splitters, so every instruction saved here is an issue slot for the main loop, and the tail of a launch (the last
// tile's epilogue, which nothing overlaps) shrinks with it.
// kMap: the 8 rows of a lane go to the row-mapped pixels roff[i] (sub-pixel transposed convolutions) instead of to
// row_base + sub + 4 i.
template <int kEpi, bool kAtomic, bool kColsum, bool kMap = false>
__device__ __forceinline__ void epilogue_chunk_fast(const EpiParams& e, uint32_t st_addr, const uint32_t (&r)[32],
                                                    int lane, int row_base, int col0, const float (&bv)[4],
                                                    const float4 (&aux)[8], const int* roff = nullptr) {
  const int sub = lane >> 3, cq = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts128(st_addr + ((lane * 8 + (j ^ (lane & 7))) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  const int col = col0 + 4 * cq;
  float* cptr = e.C + static_cast<int64_t>(row_base + sub) * e.ldc + col;
  float* hptr = kEpi == MVAE_EPI_BIAS_SWISH ? e.out2 + static_cast<int64_t>(row_base + sub) * e.ldout2 + col : nullptr;
  const int64_t cstep = 4 * e.ldc, hstep = 4 * e.ldout2;
  // row rl = 4 i + sub of the staging tile: float4 index rl * 8 + (cq ^ (rl & 7)), (rl & 7) = 4 (i & 1) + sub
  const uint32_t rd0 = st_addr + (((sub * 8) + (cq ^ sub)) << 4);
  const uint32_t rd1 = st_addr + ((((4 + sub) * 8) + (cq ^ (4 + sub))) << 4);
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 s4 = lds128(((i & 1) ? rd1 : rd0) + ((i >> 1) << 10));   // 8 rows = 1024 B further per pair of i
    float v[4] = {s4.x + bv[0], s4.y + bv[1], s4.z + bv[2], s4.w + bv[3]};
    if (kEpi == MVAE_EPI_MUL_DSWISH) {
      const float a[4] = {aux[i].x, aux[i].y, aux[i].z, aux[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sg = sigmoidf_acc(a[q]);
        v[q] *= sg * (1.0f + a[q] * (1.0f - sg));
      }
    }
    if (kColsum) {
#pragma unroll
      for (int q = 0; q < 4; ++q) cs[q] += v[q];
    }
    if (kMap) {
      cptr = e.C + static_cast<int64_t>(roff[i]) * e.ldc + col;
      if (kEpi == MVAE_EPI_BIAS_SWISH) hptr = e.out2 + static_cast<int64_t>(roff[i]) * e.ldout2 + col;
    }
    if (kAtomic) ptx::red_add_v4_f32(cptr, v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
    if (kEpi == MVAE_EPI_BIAS_SWISH) {
      *reinterpret_cast<float4*>(hptr) = make_float4(v[0] * sigmoidf_acc(v[0]), v[1] * sigmoidf_acc(v[1]),
                                                     v[2] * sigmoidf_acc(v[2]), v[3] * sigmoidf_acc(v[3]));
      hptr += hstep;
    }
    cptr += cstep;
  }
  if (kColsum) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 8);
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 16);
    }
    if (sub == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) ptx::red_add_f32(e.colsum + col + q, cs[q]);
    }
  }
  __syncwarp();  // the staging tile is rewritten by the next chunk
}

// 16-column tail chunk of an interior tile with the plain STORE epilogue (N = 48 outputs of the C = 3 conv layers: a
// 32-column chunk + this one).  Staging rows are 64 B; float4 index row * 4 + (chunk ^ ((row >> 1) & 3)) keeps both the
// row-wise writes and the 8-rows-x-64-B transposed reads conflict-free.
template <bool kAtomic>
__device__ __forceinline__ void epilogue_chunk_fast16(const EpiParams& e, uint32_t st_addr, const uint32_t (&r)[32],
                                                      int lane, int row_base, int col0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(st_addr + ((lane * 4 + (j ^ ((lane >> 1) & 3))) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  const int sub = lane >> 2, cq = lane & 3;
  const int col = col0 + 4 * cq;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bv[q] = __ldg(e.bias + col + q);
  }
  float* cptr = e.C + static_cast<int64_t>(row_base + sub) * e.ldc + col;
  const int64_t cstep = 8 * e.ldc;
  const uint32_t rd = st_addr + (((sub * 4) + (cq ^ ((sub >> 1) & 3))) << 4);   // row 8 i + sub: + 512 B per i
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 s4 = lds128(rd + (i << 9));
    const float v[4] = {s4.x + bv[0], s4.y + bv[1], s4.z + bv[2], s4.w + bv[3]};
    if (kAtomic) ptx::red_add_v4_f32(cptr, v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
    cptr += cstep;
  }
  __syncwarp();  // the staging tile is rewritten by the next chunk
}

// Fused split-K, phase B (last arriver of a (tile, epilogue warp) slab): 32 rows x `ncols` columns of the summed partial
// accumulators are read back from the scratch matrix (L2: the partials were red.add'ed there by other SMs), re-zeroed,
// and pushed through the real epilogue.  Lane mapping as in epilogue_chunk: sub = lane / 8 rows of a group of 4,
// cq = lane % 8 -> 4 consecutive columns; all 8 row loads of a lane are in flight together.  N % 4 == 0 (checked on host).
__device__ __forceinline__ float4 ldcg128(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void epilogue_from_scratch(const EpiParams& e, float* P, int64_t ldp, int lane, int row_base,
                                                      int col0, int ncols) {
  const int sub = lane >> 3, cq = lane & 7;
  const int col = col0 + 4 * cq;
  const bool col_ok = 4 * cq < ncols && col < e.N;
  float4 acc[8], aux[8];
  const bool dsw = e.epilogue == MVAE_EPI_MUL_DSWISH;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + 4 * i + sub;
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    aux[i] = acc[i];
    if (col_ok && row < e.M) {
      acc[i] = ldcg128(P + static_cast<int64_t>(row) * ldp + col);
      if (dsw) aux[i] = *reinterpret_cast<const float4*>(e.aux + static_cast<int64_t>(row) * e.ldaux + col);
    }
  }
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (col_ok && e.bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bv[q] = __ldg(e.bias + col + q);
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + 4 * i + sub;
    if (!(col_ok && row < e.M)) continue;
    *reinterpret_cast<float4*>(P + static_cast<int64_t>(row) * ldp + col) = make_float4(0.f, 0.f, 0.f, 0.f);   // ready for the next launch
    float v[4] = {acc[i].x + bv[0], acc[i].y + bv[1], acc[i].z + bv[2], acc[i].w + bv[3]};
    if (dsw) {
      const float a[4] = {aux[i].x, aux[i].y, aux[i].z, aux[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sg = sigmoidf_acc(a[q]);
        v[q] *= sg * (1.0f + a[q] * (1.0f - sg));
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[q] += v[q];
    *reinterpret_cast<float4*>(e.C + static_cast<int64_t>(row) * e.ldc + col) = make_float4(v[0], v[1], v[2], v[3]);
    if (e.epilogue == MVAE_EPI_BIAS_SWISH)
      *reinterpret_cast<float4*>(e.out2 + static_cast<int64_t>(row) * e.ldout2 + col) =
          make_float4(v[0] * sigmoidf_acc(v[0]), v[1] * sigmoidf_acc(v[1]), v[2] * sigmoidf_acc(v[2]), v[3] * sigmoidf_acc(v[3]));
  }
  if (e.colsum != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 8);
      cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 16);
    }
    if (sub == 0 && col_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) ptx::red_add_f32(e.colsum + col + q, cs[q]);
    }
  }
}

// kExtra: bit 0 = the launch contains fused split-K problems, bit 1 = row-mapped problems.  Separate instantiations: the
// extra code next to the main epilogue costs registers (100-500 bytes of spills at the 128-register budget of the 14-warp
// CTA), which launches without such problems -- all the large plain ones -- must not pay.
template <bool kSplit, bool kPair, int kExtra = 0>
__device__ __forceinline__ void gemm_body(const GemmBatch& batch) {
  constexpr bool kFused = (kExtra & 1) != 0;     // fused split-K (last-arriver epilogue)
  constexpr bool kRowMap = (kExtra & 2) != 0;    // row-mapped epilogue (sub-pixel transposed convolutions)
  using C = Cfg<kSplit, kPair>;
  static_assert(!kPair || kSplit, "the CTA-pair variant exists for the 3xTF32 mode only");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t ready_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  pdl_trigger();   // the next kernel of the stream may be scheduled as soon as every CTA of this one is resident (common.h)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // pair mode: rank 0 = leader (issues the MMAs, owns ready / tmem_empty barriers), rank 1 = peer
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const int tile0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < batch.num_problems; ++i) {
      ptx::prefetch_tmap(&batch.p[i].map_a);
      ptx::prefetch_tmap(&batch.p[i].map_b);
      if (batch.p[i].has_b_lo) ptx::prefetch_tmap(&batch.p[i].map_b_lo);
    }
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&ready_bar[s], (kPair ? 2 : 1) * NUM_SPLIT_WARPS * 32);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], (kPair ? 2 : 1) * NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) {
      ptx::tmem_alloc2(&tmem_base_smem, C::kTmemCols);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(&tmem_base_smem, C::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all();   // also: the peer's barriers are initialised before anyone arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();      // everything above ran beside the predecessor's tail; from here on its results are complete and visible
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int dn = 0;
      int tl = 0;
      int seen[6] = {0, 0, 0, 0, 0, 0};   // dependency counters of the coming tile, read one tile ahead
      bool peeked = false;
      // row blocks [rb0, rb1] of the producer that hold the A rows of tile `ti` (K-major A: my row block; MN-major A,
      // i.e. wgrad: the tile's k range)
      auto dep_range = [&](const TileInfo& ti, const GemmProblem& p, int my_rb, const int*& ctr, int& rb0, int& rb1) {
        rb0 = rb1 = my_rb;
        if (p.a_mn) {
          const int k_end = ti.kb_end * BLOCK_K < p.K ? ti.kb_end * BLOCK_K : p.K;
          rb0 = (ti.kb_begin * BLOCK_K) / BLOCK_M;
          rb1 = (k_end - 1) / BLOCK_M;
        }
        if (rb1 >= p.dep_row_blocks) rb1 = p.dep_row_blocks - 1;  // (pair mode: the peer's rows may lie beyond M)
        ctr = batch.ws + WS_CTR0 + p.dep_ctr_base;
      };
      dbg_stamp(batch, 0, dn);
      for (int t = tile0; t < batch.total_tiles; t += tile_step, ++tl) {
        const TileInfo ti = decode_tile(batch, t);
        const GemmProblem& p = batch.p[ti.prob];
        const int block_n = p.block_n, a_mn = p.a_mn, b_mn = p.b_mn;
        if (batch.tilelog != nullptr) { tilelog(batch, tl, 0, t); tilelog(batch, tl, 1, globaltimer_ns()); }
        const CUtensorMap* map_a = &p.map_a;
        const CUtensorMap* map_b = &p.map_b;
        const int my_rb = kPair ? 2 * ti.m_blk + static_cast<int>(rank) : ti.m_blk;   // my 128-row block of C / A
        const int m0 = my_rb * BLOCK_M;
        const int n_mine = kPair ? block_n >> 1 : block_n;          // B rows this CTA stages
        const int n0 = ti.n_blk * block_n + (kPair ? static_cast<int>(rank) * n_mine : 0);
        const uint32_t a_bytes = OPERAND_BYTES;
        const uint32_t b_bytes = static_cast<uint32_t>(n_mine) * BLOCK_K * 4;
        if (batch.ws != nullptr && p.dep >= 0) {
          // chain mode: the rows of A this tile reads are produced by earlier tiles of this same launch.
          // Fast check first: the producers are usually long done, and the acquire loads of the slow path cost ~0.5 us
          // EACH (measured 1.1 us per tile for one row block, 2.1 us for the four of a wgrad tile -- the TMA pipeline ran
          // dry at every tile boundary).  Relaxed gpu-scope loads read the counters at L2, all in flight together, and
          // they were already issued while the PREVIOUS tile's loads were being enqueued (peek below); the operands
          // themselves are then fetched by TMA from L2 (never through this SM's L1).
          const int* ctr; int rb0, rb1;
          dep_range(ti, p, my_rb, ctr, rb0, rb1);
          if (!peeked) {
#pragma unroll
            for (int j = 0; j < 6; ++j)   // unconditional loads (index clamped): nothing consumes the value until the check
              seen[j] = ld_relaxed_gpu(ctr + (rb0 + j <= rb1 ? rb0 + j : rb1));
          }
          bool all_done = rb1 - rb0 < 6;
#pragma unroll
          for (int j = 0; j < 6; ++j) all_done = all_done && seen[j] >= p.dep_target;
          // (No proxy fence on this side in the fast path: the writers ordered their generic-proxy stores against the async
          // proxy BEFORE releasing the counter -- the same division of labour as st.shared -> fence.proxy.async -> barrier
          // -> TMA store -- and a fence here cost another ~0.4 us per tile.)
          if (!all_done) {
            for (int rb = rb0; rb <= rb1; ++rb) wait_row_block(ctr + rb, p.dep_target, batch.ws);
            fence_proxy_async_all();
          }
        }
        // peek at the NEXT tile's dependencies now; the values are looked at one tile later (counters only grow)
        peeked = false;
        if (batch.ws != nullptr && t + tile_step < batch.total_tiles) {
          const TileInfo tn = decode_tile(batch, t + tile_step);
          const GemmProblem& pn = batch.p[tn.prob];
          if (pn.dep >= 0) {
            const int* ctr; int rb0, rb1;
            dep_range(tn, pn, kPair ? 2 * tn.m_blk + static_cast<int>(rank) : tn.m_blk, ctr, rb0, rb1);
#pragma unroll
            for (int j = 0; j < 6; ++j)   // unconditional loads (index clamped): nothing consumes the value until the check
              seen[j] = ld_relaxed_gpu(ctr + (rb0 + j <= rb1 ? rb0 + j : rb1));
            peeked = true;
          }
        }
        if (batch.tilelog != nullptr) tilelog(batch, tl, 2, globaltimer_ns());
        int acw = 0, ach = 0, acn = 0;     // implicit K-major A: base position of the tile's first output pixel
        ConvTap abox[BLOCK_M / 32], bbox[BLOCK_N_MAX / 32];   // implicit MN-major operands: (tap, channels) of the tile's 32-wide boxes
        if (p.a_cv.cblocks != 0) {
          if (!a_mn) {
            const ConvPos tp = conv_pos(p.a_cv, m0);
            acw = tp.w; ach = tp.h; acn = tp.n;
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 32; ++j) abox[j] = conv_tap(p.a_cv, m0 + 32 * j);
          }
        }
        if (p.b_cv.cblocks != 0) {
#pragma unroll
          for (int j = 0; j < BLOCK_N_MAX / 32; ++j) bbox[j] = conv_tap(p.b_cv, n0 + 32 * j);
        }
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + OPERAND_BYTES;
          const bool b_lo_tma = kSplit && p.has_b_lo;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], a_bytes + (b_lo_tma ? 2 * b_bytes : b_bytes));
          const int k0 = kb * BLOCK_K;
          if (p.a_cv.cblocks == 0) {
            if (!a_mn) {
              ptx::tma_load_2d(sa, map_a, &full_bar[stage], k0, m0);  // box {32 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 32; ++j)                    // 4 boxes {32 m, 32 k-rows}
                ptx::tma_load_2d(sa + j * 4096, map_a, &full_bar[stage], m0 + 32 * j, k0);
            }
          } else if (!a_mn) {
            // implicit K-major A: 128 output pixels from (acw, ach, acn) x the 32 channels [c0, c0 + 32) of filter tap kb / cblocks
            const int tap = fdiv(kb, p.a_cv.m_cb), c0 = (kb - tap * p.a_cv.cblocks) * 32;
            const int th = fdiv(tap, p.a_cv.m_tw), tw = tap - th * p.a_cv.taps_w;
            ptx::tma_load_im2col_4d(sa, map_a, &full_bar[stage], c0, acw, ach, acn, static_cast<uint16_t>(tw),
                                    static_cast<uint16_t>(th));
          } else {
            // implicit MN-major A (weight-gradient form): the reduction runs over the pixels [k0, k0 + 32); the tile's 128 m
            // values are (tap, channel) pairs, 32 channels of one tap per box (abox: computed once per tile)
            const ConvPos kp = conv_pos(p.a_cv, k0);
#pragma unroll
            for (int j = 0; j < BLOCK_M / 32; ++j)
              ptx::tma_load_im2col_4d(sa + j * 4096, map_a, &full_bar[stage], abox[j].c0, kp.w, kp.h, kp.n, abox[j].tw, abox[j].th);
          }
          if (p.b_cv.cblocks == 0) {
            // tap-split B: k-block kb = (tap slot t, r); the slot's rows (K-major) / columns (MN-major) start at
            // bt_table[t] * bt_mn and the k coordinate restarts at 32 r
            int bk = k0, bn = n0;
            if (p.bt_cb != 0) {
              const int t = fdiv(kb, p.bt_m_cb);
              bk = (kb - t * p.bt_cb) * BLOCK_K;
              bn = n0 + static_cast<int>(p.bt_table[t]) * p.bt_mn;
            }
            if (!b_mn) {
              ptx::tma_load_2d(sb, map_b, &full_bar[stage], bk, bn);  // box {32 k, block_n rows}
              if (b_lo_tma) ptx::tma_load_2d(sb + C::kBBytes, &p.map_b_lo, &full_bar[stage], bk, bn);
            } else {
              for (int j = 0; j < n_mine / 32; ++j)
                ptx::tma_load_2d(sb + j * 4096, map_b, &full_bar[stage], bn + 32 * j, bk);
              if (b_lo_tma)
                for (int j = 0; j < n_mine / 32; ++j)
                  ptx::tma_load_2d(sb + C::kBBytes + j * 4096, &p.map_b_lo, &full_bar[stage], bn + 32 * j, bk);
            }
          } else {
            // implicit MN-major B (Conv2d weight gradient: dW = dy^T im2col(x)): boxes of 32 pixels x 32 channels of one tap
            const ConvPos kp = conv_pos(p.b_cv, k0);
#pragma unroll
            for (int j = 0; j < BLOCK_N_MAX / 32; ++j)
              if (j < n_mine / 32)
                ptx::tma_load_im2col_4d(sb + j * 4096, map_b, &full_bar[stage], bbox[j].c0, kp.w, kp.h, kp.n, bbox[j].tw, bbox[j].th);
          }
          dbg_stamp(batch, 0, dn);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (pair mode: the leader CTA's)
    // The whole warp runs the loop in uniform control flow and ONE elected lane issues: tcgen05.mma / commit take their
    // operands from uniform registers, and inside an `if (lane == 0)` region the compiler cannot prove uniformity -- it
    // wrapped every MMA in an ELECT + 6x R2UR + branch sequence (~12 instructions).  Computed warp-uniformly the
    // descriptors live in uniform registers and an MMA is a couple of uniform adds.
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int slot = 0;        // TMEM A staging slot of the current k-block (3x modes)
      int iter = 0;
      int dn = 0;
      if (lane == 0) dbg_stamp(batch, 1, dn);
      for (int t = tile0; t < batch.total_tiles; t += tile_step, ++iter) {
        const TileInfo ti = decode_tile(batch, t);
        const GemmProblem& p = batch.p[ti.prob];
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        if (kPair) ptx::mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        else ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N_MAX;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kSplit ? 0 : p.a_mn) << 15) |
                               (static_cast<uint32_t>(p.b_mn) << 16) |
                               (static_cast<uint32_t>(p.block_n >> 3) << 17) | (static_cast<uint32_t>(C::kTileM >> 4) << 24);
        // One thread issues every MMA of this SM (pair), so ITS instruction stream is the tensor pipe's feed rate: with
        // the descriptors rebuilt from the problem table per MMA the loop issued one MMA per ~110 cycles (measured:
        // tools/micro/mma_rate.cu, profiles/r01_notes.md) against the pipe's 64.  The descriptors of one tile differ
        // only in the start-address field (bits [0,14), units of 16 B, no carry: smem < 256 KiB), so everything else
        // is built once per tile and each MMA costs one 64-bit add.
        const uint32_t smem0 = ptx::smem_u32(smem);
        const uint64_t da0 = make_desc(smem0, p.a_lbo, p.a_sbo, p.a_layout);                   // stage 0, k-slice 0
        const uint64_t db0 = make_desc(smem0 + OPERAND_BYTES, p.b_lbo, p.b_sbo, p.b_layout);
        const uint64_t a_k16 = p.a_kstep >> 4, b_k16 = p.b_kstep >> 4;
        const uint32_t a_tm0 = tmem_base + ACC_COLS;
        uint32_t accumulate = 0;
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          if (kPair) ptx::mbar_wait_cluster(&ready_bar[stage], phase);   // both CTAs' splitters arrive here
          else ptx::mbar_wait(kSplit ? &ready_bar[stage] : &full_bar[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) dbg_stamp(batch, 1, dn);
          const uint64_t st16 = static_cast<uint64_t>(stage * (C::kStageBytes >> 4));
          const uint64_t db = db0 + st16;
          if (kSplit) {
            // A (hi | lo) comes from tensor memory, B raw (= hi by hardware truncation) and B lo from smem
            const uint64_t db_lo = db + (C::kBBytes >> 4);
            const uint32_t a_hi = a_tm0 + slot * A_STAGE_COLS;
            const uint32_t a_lo = a_hi + 32;
            if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
              const uint32_t acc0 = ks == 0 ? accumulate : 1u;
              if (kPair) {
                ptx::mma_tf32_ts2(d_tmem, a_lo + ks * UMMA_K, db + ks * b_k16, idesc, acc0);   // lo*hi
                ptx::mma_tf32_ts2(d_tmem, a_hi + ks * UMMA_K, db_lo + ks * b_k16, idesc, 1u);  // hi*lo
                ptx::mma_tf32_ts2(d_tmem, a_hi + ks * UMMA_K, db + ks * b_k16, idesc, 1u);     // hi*hi
              } else {
                ptx::mma_tf32_ts(d_tmem, a_lo + ks * UMMA_K, db + ks * b_k16, idesc, acc0);    // lo*hi
                ptx::mma_tf32_ts(d_tmem, a_hi + ks * UMMA_K, db_lo + ks * b_k16, idesc, 1u);   // hi*lo
                ptx::mma_tf32_ts(d_tmem, a_hi + ks * UMMA_K, db + ks * b_k16, idesc, 1u);      // hi*hi
              }
            }
            if (kPair) ptx::mma_commit2(&empty_bar[stage]);  // frees the slot in BOTH CTAs when these MMAs retire
            else ptx::mma_commit(&empty_bar[stage]);         // frees the smem slot when these MMAs retire
            }
          } else {
            const uint64_t da = da0 + st16;
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks)
                ptx::mma_tf32_ss(d_tmem, da + ks * a_k16, db + ks * b_k16, idesc, ks == 0 ? accumulate : 1u);
              ptx::mma_commit(&empty_bar[stage]);         // frees the smem slot when these MMAs retire
            }
          }
          accumulate = 1u;
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          if (++slot == C::kTmemStages) slot = 0;
        }
        if (ptx::elect_one()) {
          if (kPair) ptx::mma_commit2(&tmem_full_bar[acc]);  // accumulator complete -> both CTAs' epilogues
          else ptx::mma_commit(&tmem_full_bar[acc]);         // accumulator complete -> epilogue
        }
      }
    }
  } else if (warp < 2 + NUM_EPI_WARPS) {
    // ===================================================== epilogue warps (8: two per TMEM lane quarter)
    // TMEM hands each lane one accumulator ROW (32 consecutive columns of a chunk).  The lanes park their rows
    // in a padded smem tile and re-read it transposed so that one warp instruction covers 4 rows x 128 B of
    // C / aux / out2 (full sectors) instead of 32 rows x 16 B.  The two warps of a quarter take alternate
    // 32-column chunks.  Bias and the first aux slab are fetched BEFORE waiting for the accumulator.
    const int ew = warp - 2;
    const int quarter = warp & 3;       // TMEM lanes [32*quarter, +32) are accessible to this warp
    const int half = ew >> 2;           // chunks half, half+2
    const int sub = lane >> 3;          // row within a group of 4
    const int c4 = (lane & 7) * 4;      // 4 consecutive columns of the 32-column chunk
    float4* stage_buf = reinterpret_cast<float4*>(smem + C::kStages * C::kStageBytes) + ew * (32 * 8);
    int iter = 0;
    int dn = 0, dn3 = 0;
    for (int t = tile0; t < batch.total_tiles; t += tile_step, ++iter) {
      const TileInfo ti = decode_tile(batch, t);
      const GemmProblem& p = batch.p[ti.prob];
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int n0 = ti.n_blk * p.block_n;
      const int my_rb = kPair ? 2 * ti.m_blk + static_cast<int>(rank) : ti.m_blk;
      const int row_base = my_rb * BLOCK_M + quarter * 32;
      const int block_n = p.block_n;
      const int nchunks = (block_n + 31) >> 5;
      const int last_c = (half + 2 < nchunks) ? half + 2 : half;
      EpiParams e;
      e.C = p.C; e.bias = p.bias; e.aux = p.aux; e.out2 = p.out2; e.colsum = p.colsum;
      e.ldc = p.ldc; e.ldaux = p.ldaux; e.ldout2 = p.ldout2; e.M = p.M; e.N = p.N;
      e.epilogue = p.epilogue; e.atomic = p.atomic;
      // fused split-K: phase A below adds the raw partial accumulator into the scratch matrix with the plain atomic
      // STORE epilogue; the real epilogue parameters are kept in `ef` for the last arriver (phase B)
      // (the real parameters are re-read from the problem table by the last arriver: keeping a second copy live across
      // phase A cost ~100 bytes of spills in every launch, fused or not)
      const bool fused = kFused && p.split_ws != nullptr;
      if (fused) {
        e.C = p.split_ws; e.ldc = p.ldp; e.bias = nullptr; e.aux = nullptr; e.out2 = nullptr; e.colsum = nullptr;
        e.epilogue = MVAE_EPI_STORE; e.atomic = 1;
      }
      const bool dsw = e.epilogue == MVAE_EPI_MUL_DSWISH;
      // output row map (sub-pixel transposed convolutions; kFused instantiation only): such problems take the general path
      RowMap rm;
      rm.IW = 0;
      if (kRowMap && p.rm_IW != 0) {
        rm.IW = p.rm_IW; rm.IHW = p.rm_IHW; rm.OW = p.rm_OW; rm.OHW = p.rm_OHW; rm.s = p.rm_s; rm.py = p.rm_py; rm.px = p.rm_px;
        rm.m_IW = p.rm_m_IW; rm.m_IHW = p.rm_m_IHW;
      }
      const bool mapped = kRowMap && rm.IW != 0;
      // interior tiles (the common case) take the specialised epilogue; edge tiles / unusual combinations the general one
      const bool rows_inside0 = row_base + 32 <= e.M && batch.dbg_flags == 0 &&
                                !(e.colsum != nullptr && e.epilogue == MVAE_EPI_BIAS_SWISH) &&
                                !(e.colsum != nullptr && e.atomic);
      const bool rows_inside = rows_inside0 && !mapped;
      const bool mapped_fast = kRowMap && mapped && rows_inside0 && !e.atomic && e.colsum == nullptr;
      int roff[8];                          // mapped rows of this lane (interior tiles of row-mapped problems)
      if (kRowMap && mapped_fast) {
#pragma unroll
        for (int i = 0; i < 8; ++i) roff[i] = static_cast<int>(map_row(rm, row_base + sub + 4 * i));
      }
      // ---- prefetch (independent of the accumulator): bias of my columns, aux rows of my chunk
      float bv[4];
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 auxv[8];
      auto prefetch = [&](int c) {
        const int col = n0 + 32 * c + c4;
        int nvalid = e.N - col;
        nvalid = (c4 < block_n - 32 * c) ? (nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid)) : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) bv[q] = (e.bias != nullptr && q < nvalid) ? __ldg(e.bias + col + q) : 0.f;
        if (dsw) {
          if (rows_inside && n0 + 32 * c + 32 <= e.N && block_n - 32 * c >= 32) {
            const float* ap = e.aux + static_cast<int64_t>(row_base + sub) * e.ldaux + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) auxv[i] = *reinterpret_cast<const float4*>(ap + static_cast<int64_t>(4 * i) * e.ldaux);
          } else if (kRowMap && mapped_fast && n0 + 32 * c + 32 <= e.N && block_n - 32 * c >= 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              auxv[i] = *reinterpret_cast<const float4*>(e.aux + static_cast<int64_t>(roff[i]) * e.ldaux + col);
          } else if (mapped) {
            a0 = load_aux4<true>(e, row_base + sub, col, nvalid, &rm);
          } else {
            a0 = load_aux4(e, row_base + sub, col, nvalid);
          }
        }
      };
      if (half < nchunks) prefetch(half);
      if (ew == 0 && lane == 0) dbg_stamp(batch, 2, dn);
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      if (ew == 0 && lane == 0) dbg_stamp(batch, 2, dn);
      int* const my_ctr = (batch.ws != nullptr && my_rb < p.row_blocks) ? batch.ws + WS_CTR0 + p.ctr_base + my_rb : nullptr;
      // pair mode: the accumulator-free barrier lives in the leader CTA and counts both CTAs' epilogue warps
      const uint32_t tmem_empty_addr = kPair ? ptx::mapa_shared(&tmem_empty_bar[acc], 0) : 0u;
      if (half >= nchunks) {  // nothing to do for this warp on a narrow tile: release immediately
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) ptx::mbar_arrive_cluster(tmem_empty_addr);
          else ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
        bool publish = true;
        if (fused) {   // (this warp role has no columns on a narrow tile, but only ONE of the split CTAs may publish for it)
          int last = 0;
          int* const tctr = batch.ws + WS_CTR0 + p.tctr_base + (ti.m_blk * p.tiles_n + ti.n_blk) * NUM_EPI_WARPS + ew;
          if (lane == 0) last = atomicAdd(tctr, 1) == p.split_k - 1;
          publish = __shfl_sync(0xffffffffu, last, 0) != 0;
        }
        if (publish && my_ctr != nullptr && lane == 0) red_release_gpu_add(my_ctr, 1);
        continue;
      }
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N_MAX;
#pragma unroll 1
      for (int c = half; c < nchunks; c += 2) {
        const int c0 = 32 * c;
        if (c != half) prefetch(c);
        uint32_t r[32];
        const int ncols = (block_n - c0) >= 32 ? 32 : 16;
        if (ncols == 32) ptx::tmem_ld_32x32(taddr_row + c0, r);
        else             ptx::tmem_ld_32x16(taddr_row + c0, r);
        ptx::tmem_ld_wait();
        if (c == last_c) {
          // all TMEM reads of this warp for this accumulator are done: hand it back to the MMA warp early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair) ptx::mbar_arrive_cluster(tmem_empty_addr);
            else ptx::mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        if (ew == 0 && lane == 0) dbg_stamp(batch, 3, dn3);
        if (rows_inside && ncols == 32 && n0 + c0 + 32 <= e.N) {
          const uint32_t st_addr = ptx::smem_u32(stage_buf);
          if (e.epilogue == MVAE_EPI_BIAS_SWISH)
            epilogue_chunk_fast<MVAE_EPI_BIAS_SWISH, false, false>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
          else if (dsw && e.colsum != nullptr)
            epilogue_chunk_fast<MVAE_EPI_MUL_DSWISH, false, true>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
          else if (dsw)
            epilogue_chunk_fast<MVAE_EPI_MUL_DSWISH, false, false>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
          else if (e.atomic)
            epilogue_chunk_fast<MVAE_EPI_STORE, true, false>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
          else if (e.colsum != nullptr)
            epilogue_chunk_fast<MVAE_EPI_STORE, false, true>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
          else
            epilogue_chunk_fast<MVAE_EPI_STORE, false, false>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv);
        } else if (kRowMap && mapped_fast && ncols == 32 && n0 + c0 + 32 <= e.N) {
          const uint32_t st_addr = ptx::smem_u32(stage_buf);
          if (e.epilogue == MVAE_EPI_BIAS_SWISH)
            epilogue_chunk_fast<MVAE_EPI_BIAS_SWISH, false, false, true>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv, roff);
          else if (dsw)
            epilogue_chunk_fast<MVAE_EPI_MUL_DSWISH, false, false, true>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv, roff);
          else
            epilogue_chunk_fast<MVAE_EPI_STORE, false, false, true>(e, st_addr, r, lane, row_base, n0 + c0, bv, auxv, roff);
        } else if (rows_inside && ncols == 16 && n0 + c0 + 16 <= e.N && e.epilogue == MVAE_EPI_STORE &&
                   e.colsum == nullptr) {
          const uint32_t st_addr = ptx::smem_u32(stage_buf);
          if (e.atomic) epilogue_chunk_fast16<true>(e, st_addr, r, lane, row_base, n0 + c0);
          else epilogue_chunk_fast16<false>(e, st_addr, r, lane, row_base, n0 + c0);
        } else {
          if (mapped) epilogue_chunk<true>(e, stage_buf, r, ncols, lane, row_base, n0 + c0, bv, a0, batch.dbg_flags, &rm);
          else epilogue_chunk(e, stage_buf, r, ncols, lane, row_base, n0 + c0, bv, a0, batch.dbg_flags);
        }
        if (ew == 0 && lane == 0) dbg_stamp(batch, 3, dn3);
      }
      if (fused) {
        // every lane's partial sums are on their way to L2; lane 0 (after the warp converged) orders them before the
        // arrival count.  The split CTA that counts last owns the finished slab: phase B.
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          int* const tctr = batch.ws + WS_CTR0 + p.tctr_base + (ti.m_blk * p.tiles_n + ti.n_blk) * NUM_EPI_WARPS + ew;
          __threadfence();
          last = atomicAdd(tctr, 1) == p.split_k - 1;
          __threadfence();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) continue;                       // (nothing to publish: the slab is not complete yet)
        e.C = p.C; e.bias = p.bias; e.aux = p.aux; e.out2 = p.out2; e.colsum = p.colsum;
        e.ldc = p.ldc; e.epilogue = p.epilogue; e.atomic = 0;
#pragma unroll 1
        for (int c = half; c < nchunks; c += 2)
          epilogue_from_scratch(e, p.split_ws, p.ldp, lane, row_base, n0 + 32 * c, (block_n - 32 * c) >= 32 ? 32 : 16);
      }
      if (my_ctr != nullptr) {
        // chain mode: publish this warp's share of the tile.  Every lane orders its own stores against the async
        // proxy (consumers read them through TMA), the warp converges, lane 0 releases at gpu scope.
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(my_ctr, 1);   // release at gpu scope: covers the lanes' stores (syncwarp above)
      }
      if (ew == 0 && lane == 0) dbg_stamp(batch, 2, dn);
      if (batch.tilelog != nullptr && ew == 0 && lane == 0) tilelog(batch, iter, 3, globaltimer_ns());
    }
  } else if (kSplit) {
    // ===================================================== operand splitters (3xTF32 only), warps 10..13
    // A: each thread owns one of the 128 tile rows; it gathers the row's 32 k-values from the swizzled smem tile,
    //    splits x = hi + lo (hi = rna_tf32(x), exact remainder lo) and stores both halves to TENSOR MEMORY
    //    (lane = row, 32 + 32 columns), from where the MMA reads its A operand -- this halves the tensor core's
    //    shared-memory traffic.  B: raw fp32 stays in smem and IS the hi operand (kind::tf32 truncates its
    //    operands: measured), only lo = x - trunc_tf32(x) is written to a second smem tile.
    const int tid = threadIdx.x - 32 * (2 + NUM_EPI_WARPS);  // 0..127
    const int quarter = warp & 3;                              // TMEM lanes [32*quarter, +32) are writable by this warp
    const int row = quarter * 32 + lane;                       // tile row owned by this thread
    int stage = 0;
    uint32_t phase = 0;
    int slot = 0;                       // TMEM A staging slot
    int old_stage = 0;                  // smem stage (and its phase) of the k-block that used this TMEM slot last
    uint32_t old_phase = 0;
    int kcount = 0;
    int dn4 = 0;
    for (int t = tile0; t < batch.total_tiles; t += tile_step) {
      const TileInfo ti = decode_tile(batch, t);
      const int a_mn = batch.p[ti.prob].a_mn;
      const bool b_presplit = batch.p[ti.prob].has_b_lo != 0;   // the lo tile of B came in by TMA (pre-split weights)
      for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
        if (tid == 0) dbg_stamp(batch, 4, dn4);
        ptx::mbar_wait(&full_bar[stage], phase);
        if (C::kStages != C::kTmemStages && kcount >= C::kTmemStages) {
          // the TMEM slot is free once the MMAs of the k-block kTmemStages back have retired: that is the completion of
          // ITS smem stage's empty barrier (the producer may not have refilled that stage yet, so full_bar says nothing)
          ptx::mbar_wait(&empty_bar[old_stage], old_phase);
          ptx::tc_fence_after();
        }
        if (tid == 0) dbg_stamp(batch, 4, dn4);
        // (shared-window ld/st: through generic pointers these compiled to LD.E / ST.E, the slower generic path)
        const uint32_t sa = ptx::smem_u32(smem) + stage * C::kStageBytes;
        uint32_t hi[32], lo[32];
        if (!a_mn) {
          // K-major tile: row r at r*128 B, 16-byte chunks XOR-swizzled with (r & 7)
          const uint32_t rp = sa + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = lds128(rp + ((j ^ (row & 7)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float h = ptx::round_tf32(xs[q]);
              hi[4 * j + q] = __float_as_uint(h);
              lo[4 * j + q] = __float_as_uint(xs[q] - h);
            }
          }
        } else {
          // MN-major tile: 4 boxes of [32 k-rows][32 m]; box = quarter; 32-byte chunks XOR-swizzled with (k & 3)
          const uint32_t bp = sa + quarter * 4096 + ((lane & 7) << 2);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = lds32(bp + k * 128 + ((((lane >> 3) ^ (k & 3))) << 5));
            const float h = ptx::round_tf32(x);
            hi[k] = __float_as_uint(h);
            lo[k] = __float_as_uint(x - h);
          }
        }
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ACC_COLS + slot * A_STAGE_COLS;
        ptx::tmem_st_32x32(ta, hi);
        ptx::tmem_st_32x32(ta + 32, lo);
        if (tid == 0) dbg_stamp(batch, 4, dn4);
        // B: lo tile only -- unless the caller supplied it pre-split (weights: x - trunc_tf32(x) is computed once per step
        // by mvae_split_lo instead of once per tile and k-block here; this B pass was ~300 of the splitters' ~1,050 cycles
        // per k-block, and the splitters are the critical path of the 3xTF32 main loop)
        if (!b_presplit) {
        const uint32_t braw = sa + OPERAND_BYTES + (tid << 4);
        const uint32_t blo = braw + C::kBBytes;
        // all loads first, then all stores: interleaved, every load would wait for the previous store (the compiler
        // cannot prove that braw and blo do not alias), which serialised 8 shared-memory round trips per k-block
        constexpr int kBIters = C::kBBytes / 16 / (NUM_SPLIT_WARPS * 32);
        float4 bx[kBIters];
#pragma unroll
        for (int i = 0; i < kBIters; ++i) bx[i] = lds128(braw + i * (NUM_SPLIT_WARPS * 32 * 16));
#pragma unroll
        for (int i = 0; i < kBIters; ++i) {
          float4 l;
          l.x = bx[i].x - __uint_as_float(__float_as_uint(bx[i].x) & 0xFFFFE000u);
          l.y = bx[i].y - __uint_as_float(__float_as_uint(bx[i].y) & 0xFFFFE000u);
          l.z = bx[i].z - __uint_as_float(__float_as_uint(bx[i].z) & 0xFFFFE000u);
          l.w = bx[i].w - __uint_as_float(__float_as_uint(bx[i].w) & 0xFFFFE000u);
          sts128(blo + i * (NUM_SPLIT_WARPS * 32 * 16), l.x, l.y, l.z, l.w);
        }
        }
        if (tid == 0) dbg_stamp(batch, 4, dn4);
        ptx::tmem_st_wait();
        ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        if (tid == 0) dbg_stamp(batch, 4, dn4);
        ptx::tc_fence_before();
        if (kPair) ptx::mbar_arrive_cluster(ptx::mapa_shared(&ready_bar[stage], 0));   // the leader's barrier
        else ptx::mbar_arrive(&ready_bar[stage]);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        if (++slot == C::kTmemStages) slot = 0;
        if (++kcount > C::kTmemStages) {
          if (++old_stage == C::kStages) { old_stage = 0; old_phase ^= 1; }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all();   // the peer's smem / TMEM / barriers stay valid until the leader is done too
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc2(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
  if (batch.ws != nullptr) {
    // the last CTA to get here zeroes the counters, so the workspace is ready for the next launch on this stream
    __shared__ int is_last;
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = atomicAdd(batch.ws + WS_DONE, 1) == static_cast<int>(gridDim.x) - 1;
      __threadfence();
    }
    __syncthreads();
    if (is_last) {
      for (int i = threadIdx.x; i < batch.num_ctrs; i += blockDim.x) batch.ws[WS_CTR0 + i] = 0;
      if (threadIdx.x == 0) batch.ws[WS_DONE] = 0;
    }
  }
}

template <bool kSplit, int kExtra = 0>
__global__ void __launch_bounds__(Cfg<kSplit>::kThreads, 1) gemm_kernel(const __grid_constant__ GemmBatch batch) {
  gemm_body<kSplit, false, kExtra>(batch);
}

// CTA-pair variant (3xTF32): launched as clusters of two CTAs (adjacent SMs of a TPC), tcgen05 cta_group::2.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<true, true>::kThreads, 1)
    gemm_pair_kernel(const __grid_constant__ GemmBatch batch) {
  gemm_body<true, true>(batch);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2-D fp32 tensor map over a row-major [rows][inner] array with row stride ld (elements).
int make_map(CUtensorMap* m, const float* base, int64_t inner, int64_t rows, int64_t ld, int box_inner,
             int box_rows, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(MVAE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0)
    return set_error(MVAE_ERR_BAD_ARG, "GEMM operand must be 16-byte aligned with ld %% 4 == 0 (ptr=%p ld=%lld)",
                     (const void*)base, (long long)ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MVAE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MVAE_OK;
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeIm2col get_encode_im2col() {
  static PFN_encodeIm2col fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeIm2col>(p);
  return fn;
}

// im2col-mode map over an NHWC fp32 tensor: `pixels` output pixels x 32 channels per load
int make_im2col_map(CUtensorMap* m, const float* base, const mvae_conv_view& v, int pixels, CUtensorMapSwizzle swizzle) {
  PFN_encodeIm2col enc = get_encode_im2col();
  if (!enc) return set_error(MVAE_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
  if (reinterpret_cast<uintptr_t>(base) & 15) return set_error(MVAE_ERR_BAD_ARG, "conv operand must be 16-byte aligned");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(v.W), static_cast<cuuint64_t>(v.H),
                        static_cast<cuuint64_t>(v.N)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(v.C) * 4, static_cast<cuuint64_t>(v.W) * v.C * 4,
                        static_cast<cuuint64_t>(v.H) * v.W * v.C * 4};
  int lo[2] = {v.lower_w, v.lower_h}, up[2] = {v.upper_w, v.upper_h};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(v.stride), static_cast<cuuint32_t>(v.stride), 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, lo, up, 32,
                   static_cast<cuuint32_t>(pixels), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MVAE_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d)", (int)r);
  return MVAE_OK;
}

// geometry of a conv view; returns the number of output pixels (0 on inconsistent parameters)
int64_t conv_geom(const mvae_conv_view& v, GemmProblem::ConvGeom* g) {
  if (v.C <= 0 || (v.C & 31) || v.N < 1 || v.H < 1 || v.W < 1 || v.stride < 1 || v.taps_h < 1 || v.taps_w < 1) return 0;
  const int oh = (v.H + v.upper_h - v.lower_h - 1) / v.stride + 1, ow = (v.W + v.upper_w - v.lower_w - 1) / v.stride + 1;
  if (oh < 1 || ow < 1) return 0;
  g->cblocks = v.C / 32; g->C = v.C; g->OW = ow; g->OHW = oh * ow;
  g->lower_h = v.lower_h; g->lower_w = v.lower_w; g->stride = v.stride; g->taps_w = v.taps_w;
  auto recip = [](int d) { return ((1ULL << 40) + static_cast<unsigned long long>(d) - 1) / static_cast<unsigned long long>(d); };
  if (ow >= 4096 || v.C >= 4096 || v.taps_w >= 4096 || static_cast<int64_t>(oh) * ow >= 4096 * 16) return 0;
  g->m_OW = recip(ow); g->m_OHW = recip(oh * ow); g->m_C = recip(v.C); g->m_cb = recip(v.C / 32); g->m_tw = recip(v.taps_w);
  if (static_cast<int64_t>(v.N) * oh * ow >= (1LL << 28)) return 0;     // (exactness range of the reciprocal divisions)
  return static_cast<int64_t>(v.N) * oh * ow;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Debug knobs (environment) for the MN-major operand encoding; unset in normal operation.
struct MnEncoding {
  uint32_t lbo = 4096, sbo = 512, kstep = 1024, layout = 1;
  CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};
MnEncoding mn_encoding() {
  MnEncoding e;
  if (const char* v = getenv("MVAE_DBG_MN_LBO")) e.lbo = static_cast<uint32_t>(atoi(v));
  if (const char* v = getenv("MVAE_DBG_MN_SBO")) e.sbo = static_cast<uint32_t>(atoi(v));
  if (const char* v = getenv("MVAE_DBG_MN_KSTEP")) e.kstep = static_cast<uint32_t>(atoi(v));
  if (const char* v = getenv("MVAE_DBG_MN_LAYOUT")) e.layout = static_cast<uint32_t>(atoi(v));
  if (const char* v = getenv("MVAE_DBG_MN_SWIZZLE")) e.swizzle = static_cast<CUtensorMapSwizzle>(atoi(v));
  return e;
}

}  // namespace
}  // namespace mvae

using namespace mvae;

namespace mvae {
namespace {

// Shared host path of mvae_gemm_batch (independent problems) and mvae_gemm_chain (problems whose A operand is produced
// by an earlier problem of the same launch; deps != nullptr, ws != nullptr).
// MVAE_PAIR=1 routes 3xTF32 launches to the CTA-pair kernel (default off until it is the measured winner).
bool use_pair_kernel(int precision) {
  static int cached = -1;
  if (cached < 0) {
    const char* v = getenv("MVAE_PAIR");
    cached = (v != nullptr && atoi(v) != 0) ? 1 : 0;
  }
  return cached == 1 && precision == MVAE_PREC_3XTF32;
}

// Clusters of two CTAs that can be resident at once (one CTA per SM; <= 74 on a B200).  Chain mode needs every CTA of
// the launch resident, so the grid never exceeds this.
int max_pair_clusters() {
  static int cached = 0;
  if (cached > 0) return cached;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * 74, 1, 1);
  cfg.blockDim = dim3(Cfg<true, true>::kThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg<true, true>::kSmemBytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_pair_kernel, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = mvae_device_sm_count() / 2 - 4;   // conservative guess
  }
  cached = n;
  return cached;
}

int launch_problems(const char* who, const mvae_gemm_desc* descs, const int32_t* deps, int n, int max_n, int32_t* ws,
                    int64_t ws_ints, int precision, void* stream) {
  if (descs == nullptr || n < 1 || n > max_n) return set_error(MVAE_ERR_BAD_ARG, "%s: n must be in [1,%d]", who, max_n);
  if (precision != MVAE_PREC_TF32 && precision != MVAE_PREC_3XTF32)
    return set_error(MVAE_ERR_BAD_ARG, "%s: unknown precision %d", who, precision);
  GemmBatch batch;   // ~7 KiB of launch parameters (copied by value at launch)
  memset(&batch, 0, sizeof(batch));
  int tiles = 0, ctrs = 0;
  bool any_fused = false, any_rowmap = false;
  const bool pair = use_pair_kernel(precision);
  const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
  const MnEncoding mn = mn_encoding();
  for (int i = 0; i < n; ++i) {
    const mvae_gemm_desc& d = descs[i];
    GemmProblem& p = batch.p[i];
    if (d.M < 1 || d.N < 1 || d.K < 1 || !d.A || !d.B || !d.C)
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: bad shape/pointer (M=%d N=%d K=%d)", who, i, d.M, d.N, d.K);
    if (d.epilogue == MVAE_EPI_BIAS_SWISH && !d.out2)
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: BIAS_SWISH needs out2 (bias may be NULL: bias=False convs)", who, i);
    if (d.epilogue == MVAE_EPI_MUL_DSWISH && !d.aux)
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: MUL_DSWISH needs aux", who, i);
    if (d.epilogue < 0 || d.epilogue > MVAE_EPI_MUL_DSWISH)
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: unknown epilogue %d", who, i, d.epilogue);
    const int split = d.split_k < 1 ? 1 : d.split_k;
    // fused split-K: partial sums meet in the scratch matrix d.split_ws, the last arriver runs the real epilogue
    const bool fused = split > 1 && !d.accumulate && d.split_ws != nullptr;
    const bool atomic = !fused && (split > 1 || d.accumulate);
    if (atomic && (d.bias || d.epilogue != MVAE_EPI_STORE || d.colsum))
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: split_k > 1 with a bias / activation / colsum epilogue needs split_ws "
                       "(fused split-K); accumulate only with the plain STORE epilogue", who, i);
    if (fused && (ws == nullptr || (d.N & 3) || (reinterpret_cast<uintptr_t>(d.split_ws) & 15) || pair))
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: fused split-K needs mvae_gemm_chain (counter workspace), N %% 4 == 0 and a "
                       "16-byte aligned split_ws", who, i);
    if ((d.ldc & 3) || (reinterpret_cast<uintptr_t>(d.C) & 15))
      return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: C must be 16B aligned with ldc %% 4 == 0", who, i);
    // MMA N: multiple of 16; MN-major B is fetched in 32-wide boxes.
    // Pair mode: each CTA stages half of the B tile, so the half must itself be a whole number of boxes / swizzle atoms.
    int block_n = d.N >= BLOCK_N_MAX ? BLOCK_N_MAX : round_up(d.N, d.b_mn_major ? (pair ? 64 : 32) : 16);
    p.block_n = block_n;
    p.M = d.M; p.N = d.N; p.K = d.K;
    p.tiles_m = (d.M + tile_m - 1) / tile_m;
    p.row_blocks = (d.M + BLOCK_M - 1) / BLOCK_M;
    p.tiles_n = (d.N + block_n - 1) / block_n;
    p.num_kblocks = (d.K + BLOCK_K - 1) / BLOCK_K;
    int s = split > p.num_kblocks ? p.num_kblocks : split;
    p.kblocks_per_split = (p.num_kblocks + s - 1) / s;
    s = (p.num_kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;  // no empty splits
    p.split_k = s;
    p.tile_begin = tiles;
    tiles += p.tiles_m * p.tiles_n * s;
    p.ctr_base = ctrs;
    ctrs += p.row_blocks;
    p.split_ws = nullptr; p.ldp = 0; p.tctr_base = 0;
    if (fused && s > 1) {
      any_fused = true;
      p.split_ws = d.split_ws;
      p.ldp = (d.N + 3) / 4 * 4;
      p.tctr_base = ctrs;
      ctrs += p.tiles_m * p.tiles_n * NUM_EPI_WARPS;
    }
    p.a_mn = d.a_mn_major ? 1 : 0;
    p.b_mn = d.b_mn_major ? 1 : 0;
    p.epilogue = d.epilogue;
    p.atomic = atomic ? 1 : 0;
    p.C = d.C; p.bias = d.bias; p.aux = d.aux; p.out2 = d.out2; p.colsum = d.colsum;
    p.ldc = d.ldc; p.ldaux = d.ldaux; p.ldout2 = d.ldout2;
    p.a_lbo = p.a_mn ? mn.lbo : 16u;  p.a_sbo = p.a_mn ? mn.sbo : 1024u;
    p.a_kstep = p.a_mn ? mn.kstep : 32u;  p.a_layout = p.a_mn ? mn.layout : 2u;
    p.b_lbo = p.b_mn ? mn.lbo : 16u;  p.b_sbo = p.b_mn ? mn.sbo : 1024u;
    p.b_kstep = p.b_mn ? mn.kstep : 32u;  p.b_layout = p.b_mn ? mn.layout : 2u;
    p.dep = -1;
    if (deps != nullptr && deps[i] >= 0) {
      // A must be (a row-block aligned slice of) the C or out2 matrix of an EARLIER problem of this chain: stored rows
      // of A are rows of the producer for K-major A, and the reduction index for MN-major A (wgrad).
      const int q = deps[i];
      if (q >= i) return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: dependency %d must be an earlier problem", who, i, q);
      const mvae_gemm_desc& dq = descs[q];
      const GemmProblem& pq = batch.p[q];
      int64_t row_off = -1;
      const float* bases[2] = {dq.C, dq.out2};
      const int64_t lds[2] = {dq.ldc, dq.ldout2};
      for (int b = 0; b < 2 && row_off < 0; ++b) {
        if (bases[b] == nullptr || lds[b] != d.lda || d.A < bases[b]) continue;
        const int64_t delta = d.A - bases[b];
        if (delta % lds[b] == 0 && delta / lds[b] < dq.M) row_off = delta / lds[b];
      }
      const int a_rows = p.a_mn ? d.K : d.M;   // stored rows of A
      if (row_off < 0 || row_off % BLOCK_M != 0 || row_off + a_rows > dq.M)
        return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: A is not a %d-row aligned slice of the output of problem %d", who, i,
                         BLOCK_M, q);
      p.dep = q;
      p.dep_ctr_base = pq.ctr_base + static_cast<int>(row_off / BLOCK_M);
      p.dep_row_blocks = pq.row_blocks - static_cast<int>(row_off / BLOCK_M);
      // (fused split-K producers publish once per (tile, warp): only the last-arriving split CTA does)
      p.dep_target = NUM_EPI_WARPS * pq.tiles_n * (pq.split_ws != nullptr ? 1 : pq.split_k);
    }
    int rc;
    memset(&p.a_cv, 0, sizeof(p.a_cv)); memset(&p.b_cv, 0, sizeof(p.b_cv));
    if (d.a_view.C > 0) {
      // implicit-GEMM A: rows (K-major) or the reduction index (MN-major) are the output pixels of the view
      const int64_t pixels = conv_geom(d.a_view, &p.a_cv);
      const int64_t kk = static_cast<int64_t>(d.a_view.taps_h) * d.a_view.taps_w * d.a_view.C;
      if (pixels == 0 || pair || p.dep >= 0 || (!p.a_mn && (pixels != d.M || kk != d.K)) || (p.a_mn && (pixels != d.K || kk != d.M)))
        return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: a_view does not match M/K (pixels %lld, taps*C %lld; C %% 32 == 0; no "
                         "chain dependency, no pair kernel)", who, i, (long long)pixels, (long long)kk);
      rc = make_im2col_map(&p.map_a, d.A, d.a_view, p.a_mn ? 32 : BLOCK_M, p.a_mn ? mn.swizzle : CU_TENSOR_MAP_SWIZZLE_128B);
    } else if (!p.a_mn) rc = make_map(&p.map_a, d.A, d.K, d.M, d.lda, BLOCK_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B);
    else         rc = make_map(&p.map_a, d.A, d.M, d.K, d.lda, 32, BLOCK_K, mn.swizzle);
    if (rc) return rc;
    if (d.b_view.C > 0) {
      const int64_t pixels = conv_geom(d.b_view, &p.b_cv);
      const int64_t kk = static_cast<int64_t>(d.b_view.taps_h) * d.b_view.taps_w * d.b_view.C;
      if (pixels == 0 || pair || !p.b_mn || pixels != d.K || kk != d.N || (d.N % 32) != 0)
        return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: b_view must be MN-major with pixels == K and taps*C == N", who, i);
      rc = make_im2col_map(&p.map_b, d.B, d.b_view, 32, mn.swizzle);
    } else if (!p.b_mn) rc = make_map(&p.map_b, d.B, d.K, d.N, d.ldb, BLOCK_K, pair ? block_n / 2 : block_n, CU_TENSOR_MAP_SWIZZLE_128B);
    else         rc = make_map(&p.map_b, d.B, d.N, d.K, d.ldb, 32, BLOCK_K, mn.swizzle);
    if (rc) return rc;
    p.bt_cb = 0; p.bt_mn = 0; p.bt_m_cb = 0; memset(p.bt_table, 0, sizeof(p.bt_table));
    p.rm_IW = 0;
    if (d.b_tap_slots > 0) {
      // K = slots * b_tap_k; B is a [rows][cols] matrix in which slot t occupies rows (K-major) / columns (MN-major)
      // [b_tap_table[t] * b_tap_mn, + N) and k in [0, b_tap_k): the maps made above cover it if their extents say so
      if (d.b_tap_slots > 16 || d.b_tap_k < 32 || (d.b_tap_k & 31) || static_cast<int64_t>(d.b_tap_slots) * d.b_tap_k != d.K ||
          d.b_tap_mn < d.N || d.b_view.C > 0 || pair)
        return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: tap-split B needs slots <= 16, b_tap_k %% 32 == 0, slots * b_tap_k == K, "
                         "b_tap_mn >= N", who, i);
      int max_t = 0;
      for (int t = 0; t < d.b_tap_slots; ++t) {
        p.bt_table[t] = static_cast<unsigned char>(d.b_tap_table[t]);
        if (d.b_tap_table[t] > max_t) max_t = d.b_tap_table[t];
      }
      p.bt_cb = d.b_tap_k / 32; p.bt_mn = d.b_tap_mn;
      p.bt_m_cb = ((1ULL << 40) + static_cast<unsigned long long>(p.bt_cb) - 1) / static_cast<unsigned long long>(p.bt_cb);
      // re-encode the B map(s) over the whole tapped matrix: (max_t + 1) * b_tap_mn rows/cols, b_tap_k deep
      const int64_t mn_extent = static_cast<int64_t>(max_t + 1) * d.b_tap_mn;
      if (!p.b_mn) rc = make_map(&p.map_b, d.B, d.b_tap_k, mn_extent, d.ldb, BLOCK_K, block_n, CU_TENSOR_MAP_SWIZZLE_128B);
      else         rc = make_map(&p.map_b, d.B, mn_extent, d.b_tap_k, d.ldb, 32, BLOCK_K, mn.swizzle);
      if (rc) return rc;
    }
    if (d.rowmap_IW > 0) {
      const int ihw = d.rowmap_IH * d.rowmap_IW;
      if (d.rowmap_IH < 1 || d.rowmap_s < 1 || d.M % ihw != 0 || ihw >= 4096 * 16 || d.rowmap_IW >= 4096 || d.M >= (1 << 28) ||
          d.rowmap_py < 0 || d.rowmap_py >= d.rowmap_s || d.rowmap_px < 0 || d.rowmap_px >= d.rowmap_s || atomic || fused || d.colsum)
        return set_error(MVAE_ERR_BAD_ARG, "%s[%d]: bad output row map (M must be a whole number of IH x IW grids; plain "
                         "store / activation epilogues only)", who, i);
      p.rm_IW = d.rowmap_IW; p.rm_IHW = ihw; p.rm_s = d.rowmap_s; p.rm_OW = d.rowmap_s * d.rowmap_IW;
      p.rm_OHW = d.rowmap_s * d.rowmap_s * ihw; p.rm_py = d.rowmap_py; p.rm_px = d.rowmap_px;
      p.rm_m_IW = ((1ULL << 40) + static_cast<unsigned long long>(p.rm_IW) - 1) / static_cast<unsigned long long>(p.rm_IW);
      p.rm_m_IHW = ((1ULL << 40) + static_cast<unsigned long long>(ihw) - 1) / static_cast<unsigned long long>(ihw);
      any_rowmap = true;
    }
    p.has_b_lo = 0;
    if (d.B_lo != nullptr && precision == MVAE_PREC_3XTF32 && !pair && d.b_view.C == 0) {
      int64_t bk_ext = d.K, bn_ext = d.N;
      if (p.bt_cb != 0) {
        int max_t = 0;
        for (int t = 0; t < d.b_tap_slots; ++t) max_t = d.b_tap_table[t] > max_t ? d.b_tap_table[t] : max_t;
        bk_ext = d.b_tap_k; bn_ext = static_cast<int64_t>(max_t + 1) * d.b_tap_mn;
      }
      if (!p.b_mn) rc = make_map(&p.map_b_lo, d.B_lo, bk_ext, bn_ext, d.ldb, BLOCK_K, block_n, CU_TENSOR_MAP_SWIZZLE_128B);
      else         rc = make_map(&p.map_b_lo, d.B_lo, bn_ext, bk_ext, d.ldb, 32, BLOCK_K, mn.swizzle);
      if (rc) return rc;
      p.has_b_lo = 1;
    }
  }
  batch.num_problems = n;
  batch.total_tiles = tiles;
  if (ws != nullptr) {
    if (static_cast<int64_t>(WS_CTR0) + ctrs > ws_ints)
      return set_error(MVAE_ERR_BAD_ARG, "%s: workspace too small (%lld ints, need %d)", who, (long long)ws_ints,
                       WS_CTR0 + ctrs);
    batch.ws = ws;
    batch.num_ctrs = ctrs;
  }
  if (const char* v = getenv("MVAE_DBG_TIMELINE")) batch.dbg = reinterpret_cast<long long*>(strtoull(v, nullptr, 0));
  if (const char* v = getenv("MVAE_DBG_EPI")) batch.dbg_flags = atoi(v);
  if (const char* v = getenv("MVAE_DBG_TILELOG")) batch.tilelog = reinterpret_cast<long long*>(strtoull(v, nullptr, 0));
  const int sms = mvae_device_sm_count();
  if (sms <= 0) return set_error(MVAE_ERR_CUDA, "no CUDA device");
  // chain mode needs every CTA resident at once (a waiting tile's producers must be running): one CTA per SM
  const int grid = tiles < sms ? tiles : sms;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set[3] = {false, false, false};
  if (pair) {
    if (!attr_set[2]) {
      MVAE_CUDA_CHECK(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<true, true>::kSmemBytes));
      attr_set[2] = true;
    }
    const int clusters = tiles < max_pair_clusters() ? tiles : max_pair_clusters();
    gemm_pair_kernel<<<2 * clusters, Cfg<true, true>::kThreads, Cfg<true, true>::kSmemBytes, st>>>(batch);
  } else {
    const int extra = (any_fused ? 1 : 0) | (any_rowmap ? 2 : 0);
    const bool split3 = precision == MVAE_PREC_3XTF32;
    const int threads = split3 ? Cfg<true>::kThreads : Cfg<false>::kThreads;
    const int smem = split3 ? Cfg<true>::kSmemBytes : Cfg<false>::kSmemBytes;
    void (*kern)(const GemmBatch) = nullptr;
    if (split3) {
      kern = extra == 0 ? gemm_kernel<true, 0> : extra == 1 ? gemm_kernel<true, 1> : extra == 2 ? gemm_kernel<true, 2> : gemm_kernel<true, 3>;
    } else {
      kern = extra == 0 ? gemm_kernel<false, 0> : extra == 1 ? gemm_kernel<false, 1> : extra == 2 ? gemm_kernel<false, 2> : gemm_kernel<false, 3>;
    }
    static bool attr_done[2][4] = {{false, false, false, false}, {false, false, false, false}};
    if (!attr_done[split3 ? 1 : 0][extra]) {
      MVAE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_done[split3 ? 1 : 0][extra] = true;
    }
    // programmatic dependent launch: this kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch) and its
    // launch latency overlap the tail of the preceding kernel; gemm_body waits (griddepcontrol.wait) before it touches
    // global memory (common.h)
    launch_pdl(kern, dim3(grid), dim3(threads), static_cast<size_t>(smem), st, batch);
  }
  count_launch();
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

}  // namespace
}  // namespace mvae

extern "C" int mvae_gemm_batch(const mvae_gemm_desc* descs, int n, int precision, void* stream) {
  return launch_problems("mvae_gemm_batch", descs, nullptr, n, MVAE_GEMM_MAX_BATCH, nullptr, 0, precision, stream);
}

extern "C" int mvae_gemm_chain(const mvae_gemm_desc* descs, const int32_t* deps, int n, int32_t* ws, int64_t ws_ints,
                               int precision, void* stream) {
  if (deps == nullptr || ws == nullptr)
    return set_error(MVAE_ERR_BAD_ARG, "mvae_gemm_chain: deps and ws must be non-NULL");
  return launch_problems("mvae_gemm_chain", descs, deps, n, MVAE_GEMM_MAX_CHAIN, ws, ws_ints, precision, stream);
}

extern "C" int mvae_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                               int64_t ldy, float* h, int64_t ldh, int M, int N, int K, int precision,
                               void* stream) {
  mvae_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.A = x; d.lda = ldx; d.B = w; d.ldb = ldw; d.M = M; d.N = N; d.K = K;
  d.C = y; d.ldc = ldy; d.bias = bias; d.out2 = h; d.ldout2 = ldh;
  d.epilogue = h ? MVAE_EPI_BIAS_SWISH : MVAE_EPI_STORE;
  d.split_k = 1;
  return mvae_gemm_batch(&d, 1, precision, stream);
}

extern "C" int mvae_linear_dgrad(const float* dy, int64_t lddy, const float* w, int64_t ldw, const float* a_prev,
                                 int64_t lda_prev, float* dx, int64_t lddx, int M, int N, int K, int accumulate,
                                 int precision, void* stream) {
  // dx[M,K] = dy[M,N] * W[N,K]:  A = dy (K-major over N), B[n'=K][k'=N] = W^T -> stored [N][K] => MN-major.
  mvae_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.A = dy; d.lda = lddy; d.B = w; d.ldb = ldw; d.b_mn_major = 1;
  d.M = M; d.N = K; d.K = N;
  d.C = dx; d.ldc = lddx; d.aux = a_prev; d.ldaux = lda_prev;
  d.epilogue = a_prev ? MVAE_EPI_MUL_DSWISH : MVAE_EPI_STORE;
  d.split_k = 1; d.accumulate = accumulate;
  return mvae_gemm_batch(&d, 1, precision, stream);
}

extern "C" int mvae_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, int64_t lddw,
                                 int M, int N, int K, int split_k, int precision, void* stream) {
  // dw[N,K] += dy[M,N]^T * x[M,K]:  A[m'=N][k'=M] = dy^T (MN-major), B[n'=K][k'=M] = x^T (MN-major).
  mvae_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.A = dy; d.lda = lddy; d.a_mn_major = 1; d.B = x; d.ldb = ldx; d.b_mn_major = 1;
  d.M = N; d.N = K; d.K = M;
  d.C = dw; d.ldc = lddw;
  d.epilogue = MVAE_EPI_STORE;
  d.split_k = split_k < 1 ? 1 : split_k; d.accumulate = 1;
  return mvae_gemm_batch(&d, 1, precision, stream);
}
