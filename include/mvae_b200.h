/* mvae_b200 -- C ABI of the B200-native MVAE training-step kernels (libmvae_b200.so).
 *
 * Drop-in boundary for the hot path of mhw32/multimodal-vae-public: the reference has no FFI of its
 * own (pure Python on torch.nn / ATen), so every entry point below names the reference Python
 * site(s) it replaces (file:line relative to the upstream repo).  All pointers are DEVICE pointers
 * to fp32 (unless stated) owned by the caller; nothing is allocated behind the caller's back; every
 * call only enqueues work on `stream` (cudaStream_t passed as void*) and is capturable in a CUDA
 * graph.  Return value: 0 = ok, negative = error (see mvae_last_error()).  No C++ exceptions
 * cross this boundary.
 *
 * Layout conventions: matrices are row-major with an explicit leading dimension `ld*` in ELEMENTS
 * (ld % 4 == 0 and 16-byte aligned base pointers are required wherever a tensor is read by the
 * tensor-core GEMM through TMA).
 */
#ifndef MVAE_B200_H_
#define MVAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVAE_OK 0
#define MVAE_ERR_BAD_ARG (-1)
#define MVAE_ERR_CUDA (-2)
#define MVAE_ERR_UNSUPPORTED (-3)

/* GEMM arithmetic modes (tcgen05.mma kind::tf32, fp32 accumulate in TMEM). */
#define MVAE_PREC_TF32 0    /* one MMA per product: 10-bit mantissa operands                         */
#define MVAE_PREC_3XTF32 1  /* hi/lo operand split in-kernel, 3 MMAs per product: fp32-class result  */

/* Epilogues of mvae_gemm. */
#define MVAE_EPI_STORE 0       /* C = acc (+ bias[n])                                               */
#define MVAE_EPI_BIAS_SWISH 1  /* C = acc + bias[n] (pre-activation), out2 = C * sigmoid(C)         */
#define MVAE_EPI_MUL_DSWISH 2  /* C = acc * swish'(aux[m,n])   (aux = saved pre-activation)         */

#define MVAE_POE_VARIANT_A 0
#define MVAE_POE_VARIANT_B 1
#define MVAE_POE_NO_PRIOR 2

int mvae_version(void);
const char* mvae_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t mvae_launch_count(void);
int mvae_device_sm_count(void);

/* Implicit-GEMM convolution operand: instead of a matrix, operand A (or B) of mvae_gemm_desc may be the im2col VIEW of an
 * NHWC activation tensor [N, H, W, C] -- rows = output pixels (n, p, q) in raster order, columns k = (th, tw, c) =
 * filter tap x input channel -- fetched tile by tile with TMA im2col-mode loads (zero fill at the padding), never
 * materialised in HBM.  Replaces the im2col buffers in front of the conv GEMMs (Conv2d forward / weight gradient,
 * ConvTranspose2d data / weight gradient: fashionmnist/model.py:80-82,112-113; celeba/model.py:78-87,117-125).
 *   output pixel (p, q) reads input (lower_h + p*stride + th, lower_w + q*stride + tw);  k4 s2 p1: lower = -1, taps 4x4
 *   OH = (H + upper_h - lower_h - 1) / stride + 1 with upper = pad - (taps - 1)    (k4 s2 p1: upper = -2)
 *   C % 32 == 0; C == 0 means "plain matrix operand".  The operand pointer (A / B) is the tensor's base address.
 * As K-major A: M = N*OH*OW, K = taps_h*taps_w*C.  As MN-major operand (weight gradients: a_mn_major / b_mn_major = 1):
 * the reduction runs over the pixels, the M (or N) index is (th, tw, c).                                              */
typedef struct {
  int32_t N, H, W, C;
  int32_t lower_h, lower_w, upper_h, upper_w;
  int32_t stride;
  int32_t taps_h, taps_w;
} mvae_conv_view;

/* One GEMM problem  C[M,N] = A[M,K] * B[N,K]^T  with a fused epilogue.
 *   a_mn_major = 0: A stored [M][K] (K contiguous, lda = row stride)
 *   a_mn_major = 1: A stored [K][M] (M contiguous, lda = row stride)   -- same for B with N.
 * Replaces torch addmm/mm inside nn.Linear forward and its autograd dgrad/wgrad:
 *   mnist/model.py:75-78,95-98,117-119,136-139 (15 Linear sites), F.sigmoid*x Swish at :166-169.   */
typedef struct {
  const float* A; int64_t lda; int32_t a_mn_major;
  const float* B; int64_t ldb; int32_t b_mn_major;
  int32_t M, N, K;
  float* C; int64_t ldc;
  const float* bias;              /* [N] or NULL                                                    */
  const float* aux; int64_t ldaux;/* [M,N] operand of the epilogue or NULL                          */
  float* out2; int64_t ldout2;    /* second [M,N] output or NULL                                    */
  float* colsum;                  /* optional [N]: colsum[n] += sum_m C[m,n] of the stored values   */
                                  /* (bias gradient fused into the dgrad that produces dA)          */
  int32_t epilogue;               /* MVAE_EPI_*                                                     */
  int32_t split_k;                /* >= 1; > 1 => partial products are atomically added into C      */
  int32_t accumulate;             /* 1 => C += result (atomic red.add), 0 => C = result             */
  float* split_ws;                /* optional [M][ceil4(N)] fp32 scratch, zero-initialised ONCE by the caller (the kernel     */
                                  /* leaves it zeroed): with split_k > 1 and accumulate = 0 the k range of every output tile  */
                                  /* is split over split_k CTAs, their partial sums meet here and the LAST arriver applies    */
                                  /* the full epilogue (bias / Swish / Swish' / colsum) -- for layers with fewer output tiles  */
                                  /* than SMs (small per-GPU batches, K = 6272 classifier layers).  mvae_gemm_chain only;     */
                                  /* N % 4 == 0; one scratch per problem of a launch.                                          */
  const float* B_lo;              /* optional, MVAE_PREC_3XTF32 only: B - trunc_tf32(B), same shape / ldb as B (mvae_split_lo).  */
                                  /* B = weights are constant over a step, so their low halves are computed once per step      */
                                  /* and fetched by TMA next to B instead of being recomputed per tile in the GEMM main loop.  */
  mvae_conv_view a_view;          /* a_view.C > 0: A is the im2col view of the NHWC tensor at `A` (lda ignored)                */
  mvae_conv_view b_view;          /* b_view.C > 0: B likewise (MN-major only: the weight-gradient form)                       */
  /* Sub-pixel form of a stride-s transposed convolution (ConvTranspose2d forward, Conv2d data gradient: no cols buffer, no     */
  /* col2im pass): the s*s output parity classes are separate problems whose A is a (k/s x k/s)-tap stride-1 view, whose B      */
  /* picks the filter taps of the class out of the full weight matrix, and whose rows land on the class's output pixels.        */
  /*   tap-split B : K = b_tap_slots * b_tap_k; slot t of the reduction uses B rows (K-major) / columns (MN-major)              */
  /*                 [b_tap_table[t] * b_tap_mn, + N) and k in [0, b_tap_k).  b_tap_slots = 0: plain B.                         */
  /*   row map     : GEMM row (n, j, i) of rowmap_IH x rowmap_IW grids is stored at pixel (n, s j + py, s i + px) of an          */
  /*                 (s IH) x (s IW) NHWC image -- C, out2 and aux alike.  rowmap_IW = 0: identity.                              */
  int32_t b_tap_slots, b_tap_k, b_tap_mn;
  int32_t b_tap_table[16];
  int32_t rowmap_IH, rowmap_IW, rowmap_s, rowmap_py, rowmap_px;
} mvae_gemm_desc;

/* Launch up to MVAE_GEMM_MAX_BATCH independent problems in ONE persistent kernel. */
#define MVAE_GEMM_MAX_BATCH 4
int mvae_gemm_batch(const mvae_gemm_desc* descs, int n, int precision, void* stream);

/* Launch up to MVAE_GEMM_MAX_CHAIN problems in ONE persistent kernel where problem i may consume, as its A operand,
 * the C (or out2) matrix written by an earlier problem deps[i] < i of the same launch (deps[i] = -1: independent).
 * This is how a whole Linear+Swish stack -- mnist/model.py:81-84 (encoder), :101-105 (decoder) -- or its autograd chain
 * runs as one launch: tiles are scheduled problem after problem, a tile's TMA producer waits until the row blocks of A
 * it reads have been stored (per-row-block completion counters, release/acquire at gpu scope), so the tail of layer l
 * overlaps the head of layer l+1 and no launch boundary separates them.
 *   K-major A  : rows [m0, m0+128) of A  <- the producer's row block(s) holding those rows
 *   MN-major A : (wgrad) the k range of the tile <- the producer's row blocks covering that range
 * A must start at a multiple of 128 rows inside the producer's output and share its leading dimension.  Only the A
 * operand may be produced inside the chain, and an output that a later problem of the chain still reads must not be
 * overwritten by another problem of the same chain (no launch boundary orders those accesses any more).
 *   ws : int32 workspace of ws_ints >= MVAE_GEMM_CHAIN_WS_HEADER + sum_i ceil(M_i/128) elements, zero-initialised ONCE
 *        by the caller; the kernel leaves the counters zeroed.  ws[1] is a sticky error flag: non-zero after a
 *        dependency wait timed out (~0.5 s; results are then invalid).
 * The launch uses one CTA per SM (all CTAs co-resident), which the wait-for-producer scheme relies on. */
#define MVAE_GEMM_MAX_CHAIN 16
#define MVAE_GEMM_CHAIN_WS_HEADER 2
int mvae_gemm_chain(const mvae_gemm_desc* descs, const int32_t* deps, int n, int32_t* ws, int64_t ws_ints,
                    int precision, void* stream);

/* lo[i] = x[i] - trunc_tf32(x[i]) (low 13 mantissa bits cleared): the low halves of 3xTF32 operands, for mvae_gemm_desc.B_lo.
 * n % 4 == 0, 16-byte aligned. */
int mvae_split_lo(const float* x, float* lo, int64_t n, void* stream);

/* y = x W^T + b (optionally also h = swish(y)):  nn.Linear.forward + Swish, mnist/model.py:81-84. */
int mvae_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                    int64_t ldy, float* h, int64_t ldh, int M, int N, int K, int precision, void* stream);
/* dx = dy W  (optionally * swish'(a_prev)):  autograd of nn.Linear w.r.t. its input. */
int mvae_linear_dgrad(const float* dy, int64_t lddy, const float* w, int64_t ldw, const float* a_prev,
                      int64_t lda_prev, float* dx, int64_t lddx, int M, int N, int K, int accumulate,
                      int precision, void* stream);
/* dw += dy^T x :  autograd of nn.Linear w.r.t. weight (accumulates; caller zeroes grads once per
 * step like optimizer.zero_grad(), mnist/train.py:197). */
int mvae_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, int64_t lddw, int M,
                      int N, int K, int split_k, int precision, void* stream);
/* db[n] += sum_m dy[m,n] : bias gradient. */
int mvae_colsum_accumulate(const float* dy, int64_t lddy, float* db, int M, int N, void* stream);

/* Swish forward/backward as stand-alone element-wise ops (mnist/model.py:166-169). */
int mvae_swish_fwd(const float* x, float* y, int64_t n, void* stream);
int mvae_swish_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream);

/* nn.Embedding(V,D) gather + Swish (mnist/model.py:116,123): a[b,:] = table[idx[b],:], h = swish(a). */
int mvae_embedding_swish_fwd(const float* table, const int64_t* idx, float* a, float* h, int B, int D, int V,
                             void* stream);
/* dtable[v,:] += sum_{b: idx[b]==v} dh[b,:] * swish'(table[v,:])  (Embedding backward through Swish). */
int mvae_embedding_swish_bwd(const float* table, const int64_t* idx, const float* dh, int64_t lddh, float* dtable,
                             int B, int D, int V, void* stream);

/* Fused ProductOfExperts + reparametrize + KL for P "passes" (modality subsets) over E encoder
 * experts; the N(0,1) prior expert (prior_expert(), mnist/model.py:172-185) is implicit.
 *   mu_e[e], lv_e[e] : [B, L] with row stride ld_e  (e < E <= 24)
 *   pass_masks[p]    : bit e set => expert e present in pass p      (p < P <= 32)
 *   z                : [P*B, L] row stride ldz, pass p at rows [p*B, (p+1)*B)
 *   noise            : [P*B, L] N(0,1) draws or NULL.  NULL with training=1 => Philox(seed, offset)
 *                      noise generated in-kernel and written to noise_out (must be non-NULL).
 *   step_dev         : optional device int32 mixed into the Philox counter (so a replayed CUDA graph
 *                      draws fresh noise every step); pass the Adam step counter.
 *   training=0       : z = mu (MVAE.eval(), mnist/model.py:34-35)
 *   mu_out/lv_out    : optional [P*B, L] fused posterior parameters (MVAE.infer outputs)
 *   kl_acc           : double[P]; kl_acc[p] += sum_b KL_b (un-weighted, un-averaged), or NULL
 *   variant          : 0 = "A" (mnist/model.py:156-163), 1 = "B" (celeba/model.py:200-207);
 *                      | MVAE_POE_NO_PRIOR: the listed experts are the whole product (used by the
 *                      stand-alone ProductOfExperts.forward(mu, logvar) module call)
 * Replaces MVAE.infer's torch.cat chain + ProductOfExperts.forward + MVAE.reparametrize + the KLD
 * line of elbo_loss (mnist/model.py:29-35,46-64,156-163; mnist/train.py:56).                       */
int mvae_poe_fwd(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E, const uint32_t* pass_masks,
                 int P, int B, int L, int variant, int training, const float* noise, float* noise_out,
                 uint64_t seed, uint64_t offset, const int32_t* step_dev, float* z, int64_t ldz, float* mu_out,
                 float* lv_out, double* kl_acc, void* stream);
/* Backward of the above w.r.t. every expert output, including d(beta/B * sum KL):
 *   dz : [P*B, L] row stride lddz ;  noise: the draws used in forward (NULL iff training=0)
 *   dmu_up, dlv_up : optional [P*B, L] upstream gradients w.r.t. the fused mu / logvar outputs
 *   dmu_e[e], dlv_e[e] : [B, L] row stride ldd_e, OVERWRITTEN with the sum over passes
 *   kl_scale = annealing_factor / B_global  (mnist/train.py:57: mean over the batch); if
 *   kl_scale_dev != NULL the effective scale is kl_scale * (*kl_scale_dev) read on the device, so a
 *   captured CUDA graph can follow the per-step KL annealing schedule (mnist/train.py:180-186).   */
int mvae_poe_bwd(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E, const uint32_t* pass_masks,
                 int P, int B, int L, int variant, int training, const float* noise, const float* dz, int64_t lddz,
                 const float* dmu_up, const float* dlv_up, float kl_scale, const float* kl_scale_dev,
                 float* const* dmu_e, float* const* dlv_e, int64_t ldd_e, void* stream);

/* The same two kernels with an optional ROW INDEX per expert (gather_idx[e] = int64 [B] or NULL; gather_idx itself may be
 * NULL): expert e is then a V-row table -- the label / attribute encoders of mnist/model.py:108-125 only ever see V
 * distinct inputs, see mvae_label_table_fwd -- whose row gather_idx[e][b] is sample b's (mu | logvar); in backward the
 * sample's gradient is ADDED (red.add) into row gather_idx[e][b] of dmu_e[e] / dlv_e[e], which the caller zero-initialises
 * (this is the segmented sum by class that Embedding's autograd performs, mnist/model.py:116).                      */
int mvae_poe_fwd_g(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                   const int64_t* const* gather_idx, const uint32_t* pass_masks, int P, int B, int L, int variant,
                   int training, const float* noise, float* noise_out, uint64_t seed, uint64_t offset,
                   const int32_t* step_dev, float* z, int64_t ldz, float* mu_out, float* lv_out, double* kl_acc,
                   void* stream);
int mvae_poe_bwd_g(const float* const* mu_e, const float* const* lv_e, int64_t ld_e, int E,
                   const int64_t* const* gather_idx, const uint32_t* pass_masks, int P, int B, int L, int variant,
                   int training, const float* noise, const float* dz, int64_t lddz, const float* dmu_up,
                   const float* dlv_up, float kl_scale, const float* kl_scale_dev, float* const* dmu_e,
                   float* const* dlv_e, int64_t ldd_e, void* stream);

/* Label encoder evaluated once per CLASS (V <= 16 rows) instead of once per sample: TextEncoder of mnist/model.py:108-125
 * and fashionmnist/model.py:124-146 = Embedding(V, D) -> Swish -> Linear(D, D) -> Swish -> heads Linear(D, N3 = 2L).
 *   fwd: a2 = swish(emb) w2^T + b2, h2 = swish(a2), tab = h2 w3^T + b3      (emb [V,D], w2 [D,D], w3 [N3,D], tab [V,N3])
 *   bwd: given dtab [V,N3] (= class-wise sums of the per-sample gradients, accumulated by mvae_poe_bwd_g):
 *        dw3 += dtab^T h2, db3 += colsum(dtab), dw2 += dA2^T swish(emb), db2 += colsum(dA2), d_emb += (dA2 w2) * swish'(emb)
 *        with dA2 = (dtab w3) * swish'(a2) ACCUMULATED into the scratch d_a2 [V,D], which the caller zero-initialises before
 *        every call.  Gradients ACCUMULATE too (zero them once per step like optimizer.zero_grad(), mnist/train.py:197).  Exact regrouping of the per-sample sums; plain fp32 FMA. */
int mvae_label_table_fwd(const float* emb, const float* w2, const float* b2, const float* w3, const float* b3, float* a2,
                         float* h2, float* tab, int V, int D, int N3, void* stream);
int mvae_label_table_bwd(const float* emb, const float* w2, const float* w3, const float* a2, const float* h2,
                         const float* dtab, float* d_a2, float* d_emb, float* dw2, float* db2, float* dw3, float* db3,
                         int V, int D, int N3, void* stream);

/* Stand-alone MVAE.reparametrize (mnist/model.py:29-35, training branch): z = noise*exp(0.5*logvar) + mu.
 * noise == NULL => Philox(seed, offset) draws written to noise_out.  Backward: dmu = dz (identity),
 * dlogvar = dz * noise * 0.5 * exp(0.5*logvar). */
int mvae_reparam_fwd(const float* mu, const float* logvar, const float* noise, float* noise_out, uint64_t seed,
                     uint64_t offset, float* z, int64_t n, void* stream);
int mvae_reparam_bwd(const float* logvar, const float* noise, const float* dz, float* dlogvar, int64_t n,
                     void* stream);

/* KL(q(z|.) || N(0,1)) summed over all n = B*L elements (mnist/train.py:56) and its gradient:
 *   kl_acc[0] += -0.5 * sum(1 + lv - mu^2 - exp(lv));  dmu = scale*mu;  dlogvar = scale*0.5*(exp(lv)-1). */
int mvae_kl_fwd_bwd(const float* mu, const float* logvar, float* dmu, float* dlogvar, int64_t n, float scale,
                    double* kl_acc, void* stream);

/* Fused binary_cross_entropy_with_logits + row/batch sum + analytic gradient
 * (mnist/train.py:62-74 and the torch.sum(dim=1)/torch.mean of :43-45,57):
 *   x [R, D] logits (row stride ldx), target row r = t[(r % t_rows), :] (row stride ldt)
 *   dx = scale * (sigmoid(x) - t)  written to dx (row stride lddx; may alias x)
 *   loss_acc[r / seg_rows] += sum of BCE over row r (un-scaled), double accumulators, or NULL
 *   (seg_rows <= 0: one accumulator for all rows; seg_rows = B keeps the passes separate)
 *   loss_elem: optional [R, D] (row stride ldl) element-wise BCE values (the function's own return value) */
int mvae_bce_logits_fwd_bwd(const float* x, int64_t ldx, const float* t, int64_t ldt, int t_rows, float* dx,
                            int64_t lddx, int R, int D, float scale, double* loss_acc, int seg_rows,
                            float* loss_elem, int64_t ldl, void* stream);
/* Fused cross_entropy(input, target, eps=1e-6) row loss + gradient (mnist/train.py:77-94):
 *   x [R, K] logits, target[r % t_rows] int64 class index; dx = scale*(softmax(x+eps) - onehot).
 *   loss_rows: optional [R, K] (row stride ldl) = -onehot * log_softmax, the function's own return value.  */
int mvae_ce_fwd_bwd(const float* x, int64_t ldx, const int64_t* target, int t_rows, float* dx, int64_t lddx, int R,
                    int K, float scale, double* loss_acc, int seg_rows, float* loss_rows, int64_t ldl, void* stream);

/* 4x4 / stride 2 / pad 1 convolution data movement for the conv flavours (fashionmnist/model.py:79-82,112-114;
 * celeba/model.py:77-87,117-126), NHWC activations; the contraction runs on mvae_gemm_batch.
 *   im2col : x [B,H,W,C] -> cols [B*(H/2)*(W/2), 16*C] (row stride ld_cols), column (kh*4+kw)*C + c
 *            = operand of Conv2d forward / ConvTranspose2d backward-data.
 *   col2im : cols [B*IH*IW, 16*C] -> out [B,2IH,2IW,C] (gather-sum of the <= 4 contributing taps)
 *            = ConvTranspose2d forward tail / Conv2d backward-data tail.  Optional fusions:
 *            out_act = swish(out) (decoder forward), out *= swish'(aux) with aux [B,2IH,2IW,C] (encoder backward). */
int mvae_im2col_k4s2p1(const float* x, float* cols, int64_t ld_cols, int B, int H, int W, int C, void* stream);
int mvae_col2im_k4s2p1(const float* cols, int64_t ld_cols, float* out, float* out_act, const float* aux, int B, int IH,
                       int IW, int C, void* stream);
/* General 4x4 variants (stride, pad): CelebA also uses k4 s1 p0 (8x8 <-> 5x5, celeba/model.py:85,117).
 *   im2col: OH = (H + 2 pad - 4)/stride + 1 ;  col2im: OH = (IH - 1) stride - 2 pad + 4.                     */
int mvae_im2col_k4(const float* x, float* cols, int64_t ld_cols, int B, int H, int W, int C, int stride, int pad,
                   void* stream);
int mvae_col2im_k4(const float* cols, int64_t ld_cols, float* out, float* out_act, const float* aux, int B, int IH, int IW,
                   int C, int stride, int pad, void* stream);

/* Direct (no im2col / cols buffers, CUDA cores, HBM-bound) 4x4 / stride 2 / pad 1 convolutions on the IMAGE side of the
 * nets, where one side has 1 or 3 channels and a tensor-core GEMM would be 1-2 k-blocks deep:
 *   conv_k4s2p1_cin_fwd   : Conv2d(Cin in {1,3} -> Cout) + Swish, fashionmnist/model.py:79 (1->64), celeba/model.py:77 (3->32)
 *                           x [B,H,W,Cin] NHWC, wc [Cout][(kh,kw,ci)] -> a = conv(x) (pre-activation), h = swish(a): [B*H/2*W/2, Cout]
 *   conv_k4s2p1_cin_wgrad : dwc[co][(kh,kw,ci)] += sum_pixels da[p][co] * x-tap   (the image needs no data gradient)
 *   convT_k4s2p1_cout_fwd : ConvTranspose2d(Cin -> Cout in {1,3}), fashionmnist/model.py:114 (64->1), celeba/model.py:126 (32->3)
 *                           hin [B,IH,IW,Cin], wt [(kh,kw,co)][Cin] -> out [B,2IH,2IW,Cout] (logits)
 *   convT_k4s2p1_cout_bwd : given dout [B,2IH,2IW,Cout]: dhin = (ConvT^T dout) * swish'(ain)  (ain = pre-activation that
 *                           produced hin = swish(ain)), dwt += sum_pixels dcols^T hin                                    */
int mvae_conv_k4s2p1_cin_fwd(const float* x, const float* wc, float* a, float* h, int B, int H, int W, int Cin, int Cout,
                             void* stream);
int mvae_conv_k4s2p1_cin_wgrad(const float* x, const float* da, float* dwc, int B, int H, int W, int Cin, int Cout,
                               void* stream);
int mvae_convt_k4s2p1_cout_fwd(const float* hin, const float* wt, float* out, int B, int IH, int IW, int Cin, int Cout,
                               void* stream);
int mvae_convt_k4s2p1_cout_bwd(const float* dout, const float* hin, const float* ain, const float* wt, float* dhin, float* dwt,
                               int B, int IH, int IW, int Cin, int Cout, void* stream);

/* Train-mode BatchNorm2d/1d (+ Swish) over [rows, C] activations split into S equal row segments (one per stacked
 * model() call), celeba/model.py:80,83,86,118,121,124,149,152,176,179,182.  eps = 1e-5, momentum = 0.1 (nn defaults).
 *   bn_stats    : acc[S][C][2] (double) = per-segment sum / sum of squares (zeroed by the call)
 *   bn_finalize : mean/invstd [S][C]; running stats updated once per entry of update_order (segment indices in the
 *                 reference's call order; a segment may appear twice when the reference evaluates a net twice)
 *   bn_eval_stats: mean/invstd from the running statistics (model.eval())
 *   bn_apply    : h = [swish](gamma * (x - mean) * invstd + beta)
 *   bn_bwd      : for the live segments [seg0, seg0+nseg): dx (through Swish and BN), dgamma += , dbeta += ;
 *                 batch_stats = 0 differentiates the eval-mode (fixed statistics, affine) BatchNorm instead       */
int mvae_bn_stats(const float* x, int64_t ldx, int S, int seg_rows, int C, double* acc, void* stream);
int mvae_bn_finalize(const double* acc, int S, int seg_rows, int C, float eps, float momentum, float* mean, float* invstd,
                     float* running_mean, float* running_var, const int32_t* update_order, int n_updates, void* stream);
int mvae_bn_eval_stats(const float* running_mean, const float* running_var, int S, int C, float eps, float* mean,
                       float* invstd, void* stream);
int mvae_bn_apply(const float* x, int64_t ldx, float* h, int64_t ldh, int rows, int seg_rows, int C, const float* mean,
                  const float* invstd, const float* gamma, const float* beta, int swish_act, void* stream);
int mvae_bn_bwd(const float* x, int64_t ldx, const float* dh, int64_t lddh, float* dx, int64_t lddx, int S, int seg_rows,
                int C, int seg0, int nseg, const float* mean, const float* invstd, const float* gamma, const float* beta,
                int swish_act, int batch_stats, double* acc2, float* dgamma, float* dbeta, void* stream);

/* nn.Dropout(p) (celeba/model.py:91) on `copies` stacked calls over the same input x [x_rows, D]:
 *   y[k*x_rows + r] = x[r] * mask / (1 - p), fresh mask per copy (hash of (index, seed, *step_dev)) or mask_in.
 *   backward: dx[r] = sum_k dy[k*x_rows + r] * mask / (1 - p).                                                      */
int mvae_dropout_fwd(const float* x, int x_rows, float* y, float* mask_out, const float* mask_in, int copies, int D,
                     float p, uint64_t seed, const int32_t* step_dev, void* stream);
int mvae_dropout_bwd(const float* dy, const float* mask, float* dx, int x_rows, int copies, int D, float p, void* stream);

/* NCHW -> NHWC staging of the input image batch (the kernels work on NHWC). */
int mvae_nchw_to_nhwc(const float* x, float* y, int B, int C, int HW, void* stream);

/* Device-resident input pipeline: the uint8 dataset stays in HBM; one launch builds a batch.
 *   out[b, :] = data[idx[b], :] / 255   (transforms.ToTensor of the uint8 image, mnist/train.py:159-160)
 *   labels_out[b] = labels[idx[b]]      (may be NULL together with labels)
 * idx: int64 [B] row indices, e.g. a slice of a device-side permutation (DataLoader(shuffle=True), mnist/train.py:161).
 * Replaces the host DataLoader + per-step H2D copy of mnist/train.py:188-193.  row_bytes % 4 == 0.                   */
int mvae_gather_batch_u8(const uint8_t* data, int64_t row_bytes, const int64_t* labels, const int64_t* idx, int B,
                         float* out, int64_t ld_out, int64_t* labels_out, void* stream);

/* Fused flat Adam over one contiguous parameter bucket (torch.optim.Adam defaults, mnist/train.py:168,219):
 *   g is first multiplied by grad_scale (1/world_size after a sum-allreduce).
 *   lr_mult_dev: optional device float multiplying lr.
 *   step_count: device int32, incremented by the kernel AFTER use (1-based step = *step_count + 1). */
int mvae_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, const float* lr_mult_dev,
                   float beta1, float beta2, float eps, float grad_scale, int32_t* step_count, void* stream);

/* The data-parallel exchange step as ONE kernel per rank over NVLink peer memory (the reference
 * is single-process, mnist/train.py:196-219): gradient reduce-scatter by peer loads -> Adam on this rank's 1/world slice
 * -> parameter all-gather by peer stores, with the two rendezvous points done through flags in peer-mapped memory.
 * Replaces ncclAllReduce(gradient bucket) + mvae_adam_flat.
 *   grad_ptrs[world]  : HOST array of peer-mapped device pointers, rank order; each bucket holds n gradients followed by
 *                       `tail` extra floats (the loss scalars), all produced by the caller's backward pass on `stream`
 *   param_ptrs[world] : peer-mapped parameter buckets (n floats); every rank ends up with identical updated values
 *   flag_ptrs[world]  : peer-mapped uint32 arrays of >= 2*world + 3 words, zero-initialised once on every rank BEFORE any
 *                       rank's first call ([0,world) "gradients final", [world,2world) "stores done", [2world] sticky
 *                       error, [2world+1] scratch, [2world+2] launch counter = flag epoch)
 *   m, v              : this rank's Adam moments, full length n (only the rank's slice is touched)
 *   tail_out          : [tail] sums over ranks of the tail floats (each rank computes them for itself)
 *   step_count        : device int32 Adam step counter; read by the kernel, incremented at its end (the launch can sit
 *                       in a CUDA graph that is replayed every step)
 * Every rank of the group must enqueue the SAME sequence of these calls (it spins, bounded at ~9 s, until the peers arrive).
 * The flags carry an explicit exchange number: a peer that is at a different exchange, or that does not arrive in time,
 * sets the sticky error word flag_ptrs[rank][2*world] (1 = timeout, 2 = exchange-number mismatch); the rank then skips its
 * Adam update and every parameter store of this and all later launches -- the host must check the word. */
int mvae_allreduce_adam_p2p(float* const* grad_ptrs, float* const* param_ptrs, uint32_t* const* flag_ptrs, float* m,
                            float* v, int64_t n, int tail, float* tail_out, int rank, int world, float lr,
                            const float* lr_mult_dev, float beta1, float beta2, float eps, int32_t* step_count,
                            void* stream);

/* elbo = sum_p ( recon_img[p] * lambda_image + recon_txt[p] * lambda_text + beta * kl[p] ) / B
 * from the double accumulators above -> float32 out[0] = total, out[1..P] per pass
 * (mnist/train.py:57-58,214); beta_dev: optional device float multiplying beta.                    */
int mvae_elbo_finalize(const double* recon_img, const double* recon_txt, const double* kl, int P, float lambda_image,
                       float lambda_text, float beta, const float* beta_dev, float inv_batch, float* out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVAE_B200_H_ */
