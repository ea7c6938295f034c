/*
 * cris_b200.h — C ABI of libcris_b200.so, the B200 (sm_100a) kernel library behind
 * cris.pytorch_b200 (the drop-in for the reference's `model.segmenter.CRIS` hot path).
 *
 * The reference (DerrickWang005/CRIS.pytorch) has NO native/FFI layer: its hot path is
 * a plain nn.Module calling ATen (model/segmenter.py:29-62, model/clip.py, model/layers.py).
 * Every entry point below therefore replaces an ATen call site of the reference; the
 * file:line each one stands in for is cited next to it.  See INTEGRATION.md for the
 * ctypes binding (cris/pytorch_b200/_lib.py) a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless noted
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is retained
 *   - every function launches on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately; 0 = success, negative = error (message via cris_last_error())
 *   - re-entrant; CUDA-graph capturable (no host syncs, no allocations)
 *   - activations are bf16 "padded NHWC": [N, H+2, W+2, C] with a zero border, viewed as a
 *     row-major matrix [N*(H+2)*(W+2), C]; token tensors are [rows, C]
 */
#ifndef CRIS_B200_H_
#define CRIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ---------------------------------------------------------------- */
const char* cris_last_error(void);          /* thread-local message of the last failure   */
int  cris_abi_version(void);                /* bumped on any signature change             */
int  cris_device_check(void);               /* 0 iff current device is sm_100 (B200)      */
/* debug: device buffer of 32*16 int64 receiving per-tile clock64 stamps of CTA 0 of every following GEMM (NULL = off) */
void cris_debug_set_trace(void* dev_buf);
uint64_t cris_launch_count(void);           /* kernels launched by this library so far    */
void cris_add_launch_count(uint64_t n);     /* a replayed CUDA graph adds the launches it contains */

/* ---- the GEMM / implicit-GEMM-conv core --------------------------------------------- */
/* RELU / QUICKGELU act on (acc*alpha + bias) before the residual is added; RELU_POST after it:
 * relu(acc + bias + resid) = eval-mode conv + BatchNorm (folded) + identity + ReLU (model/clip.py:44-57) */
enum { CRIS_ACT_NONE = 0, CRIS_ACT_RELU = 1, CRIS_ACT_QUICKGELU = 2, CRIS_ACT_RELU_POST = 3 };
enum { CRIS_TAP_NONE = 0, CRIS_TAP_ACCUM = 1, CRIS_TAP_WGRAD = 2 };

/*
 * D[b][m][n] (+)= alpha * sum_k A[b][m][k] * B[b][n][k]   (bf16 x bf16 -> fp32 accumulate)
 *
 * Replaces: nn.Conv2d 1x1/3x3 (model/clip.py:17-25,165-182; model/layers.py:8-11),
 * nn.Linear / F.linear inside nn.MultiheadAttention and F.multi_head_attention_forward
 * (model/clip.py:119-139,246-260; model/layers.py:202-212,233-245), torch.bmm inside MHA,
 * and their autograd backward (dgrad / wgrad).
 *
 * Operand storage:
 *   a_mn = 0: A is K-major, element (m,k) at A[m*lda + k]     a_mn = 1: MN-major, A[k*lda + m]
 *   b_mn = 0: B is K-major, element (n,k) at B[n*ldb + k]     b_mn = 1: MN-major, B[k*ldb + n]
 *   lda/ldb/ldd/ldr in elements, must make row pitches multiples of 16 bytes; bases 16B aligned.
 * Tap modes (3x3 convolution over the padded-NHWC row matrix):
 *   CRIS_TAP_ACCUM : for t < taps: A rows shifted by tap_off[t]; B k-offset t*b_tap_k (K-major)
 *                    or n-offset t*b_tap_n (MN-major); all taps accumulate into one D tile.
 *   CRIS_TAP_WGRAD : one output slab per tap: B (MN-major) k-rows shifted by tap_off[t];
 *                    D column offset t*d_tap_n.
 *   Out-of-range rows/columns read as zero (TMA zero fill).
 * splits > 1 divides the K loop over `splits` CTAs that atomically add into fp32 D
 * (requires d_fp32 = 1, accumulate = 1; D pre-zeroed by the caller unless accumulating).
 * splits = 0 with accumulate = 1 lets the library choose tile width and split count together so that the work
 * units fill whole waves of its persistent grid.
 * Epilogue order: acc*alpha -> +bias[n] -> act -> +resid -> row mask -> store / atomic add
 *   -> optional per-column (sum, sumsq) partials of the stored values atomically added into
 *      colstats[(m/128) % 64][2][N] (fp32, zeroed by the caller) — the BatchNorm batch statistics.
 * Row mask: if mask_wp > 0, row r is a padded-NHWC pixel (h,w) = ((r % (mask_hp*mask_wp)) / mask_wp,
 *   r % mask_wp); border rows (h==0, h==hp-1, w==0, w==wp-1) are written as zero.
 */
typedef struct cris_gemm_args {
  const void* A; int64_t lda; int64_t strideA;   /* bf16 */
  const void* B; int64_t ldb; int64_t strideB;   /* bf16 */
  void*       D; int64_t ldd; int64_t strideD;
  int32_t M, N, K, batch;
  int32_t a_mn, b_mn;
  int32_t d_fp32;            /* 0: D is bf16, 1: D is fp32 */
  int32_t accumulate;        /* 1: atomically add into fp32 D */
  int32_t tap_mode, taps;
  int32_t tap_off[9];
  int32_t b_tap_k, b_tap_n, d_tap_n;
  int32_t splits;
  float   alpha;
  const float* bias;         /* [N] fp32 or NULL */
  int32_t act;
  const void* resid; int64_t ldr; int64_t strideR; int32_t resid_fp32;
  int32_t mask_hp, mask_wp;
  float*  colstats;          /* [min(64, ceil(M/128))][2][N] fp32, pre-zeroed, or NULL */
  int32_t a_rows, b_rows;    /* physical row extents of A/B for OOB zero fill (0 = derive) */
  /* two-level batch: batch index z -> (outer = z / batch_inner, inner = z % batch_inner); the outer level
   * uses strideA/B/D/R above, the inner level the strides below (attention heads: inner = head, stride 64) */
  int32_t batch_inner;       /* 0/1 = single-level batch */
  int64_t strideA2, strideB2, strideD2, strideR2;
  int32_t d_col_stride;      /* accumulate mode only: element stride between D columns (0/1 = dense);
                                lets a 3x3 wgrad land directly in the OIHW fp32 gradient */
} cris_gemm_args;

int cris_gemm(const cris_gemm_args* args, void* stream);
/* host-only: the tile width (32/64/128/256 output columns per 128-row tile) and split-K count cris_gemm would use
 * for these arguments on the current device (148 SMs assumed when no device is present) */
int cris_gemm_plan(const cris_gemm_args* args, int* tile_n, int* splits);
int cris_gemm_args_size(void);         /* sizeof(cris_gemm_args): lets a binding verify its struct mirror */
int cris_gemm_args_last_offset(void);  /* offsetof(cris_gemm_args, d_col_stride) */

/* ---- column reductions / BatchNorm ------------------------------------------------------ */
/*
 * Column reduction of a [rows, C] matrix, ACCUMULATED (fp32 atomics) into partials[min(n_blocks, 64)][2][C], which the
 * caller zero-fills; block i adds into row i % 64 (the finalize kernels below then read at most 64 rows):
 *   mode 0: (sum x, sum x^2)                       batch statistics (nn.BatchNorm2d training,
 *                                                  model/clip.py:18-26,171-183; model/layers.py:8-16,262)
 *   mode 1: (sum dz, sum dz*xhat), dz = dy*(y>0)   BatchNorm backward (batch_norm_backward_reduce); with y == NULL the
 *                                                  ReLU mask is recomputed as x*scale[c]+shift[c] > 0 (no residual)
 *   mode 2: (sum dy, -)                            bias gradients of nn.Linear / Conv2d(bias=True)
 *   mode 3: (sum dy, sum dy*xhat), per-ROW mean/rstd   LayerNorm gamma/beta gradients
 * hp/wp > 0 skips the zero border rows of a padded-NHWC matrix.
 */
int cris_col_reduce(int mode, const void* a, int64_t lda, int a_fp32, const void* a2, int64_t lda2, const void* y,
                    int64_t ldy, const void* x, int64_t ldx, int x_fp32, const float* mean, const float* rstd,
                    const float* scale, const float* shift, int64_t rows, int C, int relu, int hp, int wp,
                    float* partials, int n_blocks, void* stream);
/* fused: partials -> sums -> scale/shift/mean/invstd + running-stat update (single-rank BatchNorm forward) */
int cris_bn_finalize_fwd(const float* partials, int n_tiles, int C, float* sums, double count, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                         float* scale, float* shift, float* mean, float* invstd, void* stream);
/* fused: partials -> sums[2][C] (optional) and two parameter-gradient vectors g0 = sum0, g1 = sum1 (optional) */
int cris_stats_finalize_bwd(const float* partials, int n_tiles, int C, float* sums, float* g0, float* g1, void* stream);
/* sums[2][C] = sum over tiles of partials[n_tiles][2][C] (GEMM-epilogue or col_reduce partials) */
int cris_bn_reduce_partials(const float* partials, int n_tiles, int C, float* sums, void* stream);
/*
 * (sum, sumsq, count) -> scale/shift (+ mean, invstd) and the running-stat update (momentum, unbiased
 * variance) — torch batch_norm_gather_stats_with_counts semantics; under SyncBN the caller all-reduces
 * `sums` across ranks first and passes the global count.  training = 0 uses the running statistics.
 */
/* eval mode: the coefficients of EVERY BatchNorm in one launch.  table_dev: DEVICE array of records {const float* gamma,
 * beta, running_mean, running_var; float* out (4C floats: scale|shift|mean|invstd); int32 C; int32 c0} with c0 the
 * exclusive prefix sum of ceil(C/128); n_blocks = the total (cris_bn_eval_entry_bytes() per record) */
int cris_bn_eval_entry_bytes(void);
int cris_bn_coeffs_multi(const void* table_dev, int n_entries, int n_blocks, float eps, void* stream);
int cris_bn_coeffs(const float* sums, double count, const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                   int C, int training, void* stream);
/* BatchNorm backward statistics of a layer with residual + ReLU, and the ReLU-masked gradient in the same pass:
 * dzm = dy * (y > 0) (zero on border rows) is written to `dzm` — it is the gradient of the residual branch — while
 * (sum dzm, sum dzm*xhat) accumulate into `partials` as cris_col_reduce mode 1 does.  cris_bn_bwd_apply(dzm, y = NULL,
 * relu = 0, dres = NULL) then finishes the layer without re-reading dy and y.  Streaming kernel only: power-of-two C in
 * [8, 2048], rows >= 2048, 16-byte aligned operands (anything else is an error; use the two-call path). */
int cris_bn_bwd_reduce_masked(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x, int64_t ldx,
                              const float* mean, const float* invstd, int64_t rows, int C, int hp, int wp, void* dzm,
                              int64_t lddzm, float* partials, int n_blocks, void* stream);
/* y = mask(relu?(x*scale[c] + shift[c] (+ resid))) over a padded-NHWC row matrix [rows, C] (bf16) */
int cris_bn_apply(const void* x, int64_t ldx, const float* scale, const float* shift, const void* resid, int64_t ldr,
                  void* y, int64_t ldy, int64_t rows, int C, int relu, int hp, int wp, void* stream);
/* dx = gamma*invstd*(dz - sum_dz/count - xhat*sum_dzxhat/count); optional dres (+)= dz (residual branch) */
int cris_bn_bwd_apply(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x, int64_t ldx,
                      const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const float* sums, double count, void* dx, int64_t lddx, void* dres, int64_t lddres,
                      int dres_accumulate, int64_t rows, int C, int relu, int hp, int wp, void* stream);

/* ---- LayerNorm (nn.LayerNorm; model/clip.py:226-231, model/layers.py:199-216) --------------- */
/* y = LN(x); optional y2 = y + add[row % add_period] (bf16) — the "+ positional encoding" copy for q/k */
int cris_layernorm_fwd(const void* x, int x_fp32, int64_t ldx, const float* gamma, const float* beta, const float* add,
                       int64_t ldadd, int add_period, void* y, int y_fp32, int64_t ldy, void* y2, int64_t ldy2,
                       float* mean, float* rstd, int64_t rows, int C, float eps, void* stream);
/* dx (+)= LN backward of (dy + dy2); dgamma/dbeta (fp32, caller-zeroed or running totals) are ACCUMULATED with
 * atomics in the same pass (NULL pair = skip); dx == NULL computes the parameter gradients only */
int cris_layernorm_bwd(const void* dy, int dy_fp32, int64_t lddy, const void* dy2, int64_t lddy2, const void* x,
                       int x_fp32, int64_t ldx, const float* gamma, const float* mean, const float* rstd, void* dx,
                       int dx_fp32, int64_t lddx, int dx_accumulate, float* dgamma, float* dbeta, int64_t rows, int C,
                       void* stream);

/* ---- spatial ops on padded NHWC (nn.AvgPool2d clip.py:23,35,184; F.interpolate bilinear layers.py:54-56,293,304;
 *      f5*state layers.py:290; reshape/permute glue clip.py:113-118,140, layers.py:166,179; CoordConv
 *      layers.py:30-39; stem conv1 clip.py:165-170) ------------------------------------------------ */
int cris_avgpool2_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int N, int H, int W, int C, void* stream);
int cris_avgpool2_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int accumulate, int N, int H, int W, int C,
                      void* stream);
int cris_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int N, int H, int W, int C, void* stream);
int cris_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int accumulate, int N, int H, int W,
                        int C, void* stream);
int cris_mul_bcast(const void* x, int64_t ldx, const void* s, int64_t lds, void* y, int64_t ldy, int64_t rows,
                   int rows_per_image, int C, void* stream);
int cris_mul_bcast_bwd_s(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* ds, int n_images,
                         int rows_per_image, int C, void* stream);
int cris_padded_to_tokens(const void* x, int64_t ldx, const float* add, int64_t ldadd, void* tok, int tok_fp32,
                          int64_t ldt, int N, int H, int W, int C, void* stream);
int cris_tokens_to_padded(const void* tok, int tok_fp32, int64_t ldt, void* y, int64_t ldy, int N, int H, int W, int C,
                          void* stream);
int cris_coord_fill(void* buf, int64_t ld, int c0, int N, int H, int W, void* stream);
/* fp32 NCHW image -> bf16 im2col patches [N, H/2+2, W/2+2, 32] of the 3x3/s2 stem conv (27 taps + 5 zeros) */
int cris_stem_im2col(const float* img, void* out, int N, int Hin, int Win, void* stream);
int cris_stem_conv1_fwd(const float* img, const float* w, void* z, int64_t ldz, int N, int Hin, int Win, int Cout,
                        void* stream);
int cris_stem_conv1_wgrad(const float* img, const void* dz, int64_t lddz, float* dw, int N, int Hin, int Win, int Cout,
                          void* stream);

/* ---- token ops (softmax/dropout inside MHA clip.py:119-139,255-260, layers.py:235,240-243; nn.Embedding
 *      clip.py:440-443; EOT gather clip.py:451-452; residual dropout layers.py:237,245,249; QuickGELU clip.py:234-236) */
int cris_softmax_fwd(const void* S, void* P, void* Pd, int64_t ld, int64_t batch_stride, int nb, int Lq, int Lk,
                     int heads, const int64_t* kpm_word, int causal, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* stream);
int cris_softmax_bwd(const void* P, void* dP, int64_t ld, int64_t batch_stride, int nb, int Lq, int Lk, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* stream);
int cris_embed_fwd(const int64_t* word, const float* table, const float* pos, float* x, int B, int L, int C,
                   void* stream);
int cris_embed_bwd(const int64_t* word, const float* dx, float* dtable, float* dpos, int B, int L, int C,
                   void* stream);
int cris_eot_gather(const int64_t* word, const void* x, int x_fp32, int64_t ldx, void* out, int64_t ldo, int B, int L,
                    int C, void* stream);
int cris_eot_scatter(const int64_t* word, const void* dout, int d_fp32, int64_t ldd, void* dx, int dx_fp32,
                     int64_t lddx, int B, int L, int C, void* stream);
/* op 0: a(+b)  1: a+dropout(b)  2: dropout(a)  3: quickgelu(a)  4: b*quickgelu'(a)  5: (a>0)?b:0
 * dropout masks are a pure function of (seed + *seed_dev, element index): seed_dev (device, may be NULL) lets a
 * captured CUDA graph draw fresh masks on every replay. */
int cris_elementwise(int op, const void* a, int a_fp32, int64_t lda, const void* b, int b_fp32, int64_t ldb, void* out,
                     int out_fp32, int64_t ldo, int64_t rows, int C, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* stream);
int cris_pack_conv_weight(const float* w, void* out, int Cout, int Cin, int taps, int cin_pad, void* stream);
/* wgrad accumulator [Cout][taps][cin_pad] fp32 (what the tap-mode wgrad GEMM reduces into with unit stride) ->
 * the reference's OIHW gradient gw[co][ci][t] (autograd of nn.Conv2d, model/clip.py:17-25) */
int cris_unpack_conv_wgrad(const float* acc, float* gw, int Cout, int Cin, int taps, int cin_pad, void* stream);
int cris_pack_matrix(const float* w, void* out, int64_t rows, int cols, int ld, void* stream);
/* all weight copies in ONE launch: table_dev = DEVICE array of n_entries records {const float* src; void* dst;
 * int64 rows; int32 cols; int32 ld; int32 taps; int32 pad; int64 chunk0; const float* row_scale (NULL = none)}
 * (cris_pack_entry_bytes() each): taps == 1 is
 * cris_pack_matrix(src, dst, rows, cols, ld), taps > 1 is cris_pack_conv_weight(src, dst, rows, cols, taps, ld);
 * chunk0 = exclusive prefix sum of ceil(rows*taps*ld / cris_pack_chunk_elems()), n_chunks the total */
int cris_pack_entry_bytes(void);
int cris_pack_chunk_elems(void);
int cris_pack_multi(const void* table_dev, int n_entries, long long n_chunks, void* stream);
/* the same bf16 kernel layouts with every output row (= output channel) multiplied by row_scale[row] first:
 * eval-mode BatchNorm folded into the convolution, w'[co] = w[co] * gamma[co]/sqrt(running_var[co]+eps)
 * (model/layers.py:8-11, model/clip.py:17-25 in model.eval()) */
int cris_pack_conv_weight_scaled(const float* w, const float* row_scale, void* out, int Cout, int Cin, int taps, int cin_pad,
                                 void* stream);
int cris_pack_matrix_scaled(const float* w, const float* row_scale, void* out, int64_t rows, int cols, int ld, void* stream);
int cris_batch_reduce(const void* in, int in_fp32, int64_t ldin, float* out, int64_t ldo, int B, int T, int C,
                      int accumulate, void* stream);
int cris_small_matmul(const float* R, const float* X, float* out, int M, int K, int C, int transpose_r, int accumulate,
                      void* stream);

/* ---- text-to-pixel head + loss (Projector grouped conv model/layers.py:71-84; nearest mask resize + BCE
 *      model/segmenter.py:56-59) ------------------------------------------------------------------------ */
/* metric_counts (optional, u32 [B][2], zeroed by the caller): per-sample {sum(o & g), sum(o | g)} with
 * o = sigmoid(pred) >= metric_thr, g = target != 0 — the counts trainMetricGPU (utils/misc.py:114-129) reduces, taken
 * while the logits are still in registers */
int cris_dynconv_bce_fwd(const void* x, int64_t ldx, const float* t, int64_t ldt, const float* mask, int Hm, int Wm,
                         float* pred, float* mask_out, float* loss_sum, unsigned* metric_counts, float metric_thr, int B,
                         int H, int W, int C, void* stream);
int cris_dynconv_bce_bwd(const void* x, int64_t ldx, const float* t, int64_t ldt, const float* pred,
                         const float* target, const float* g, float* dl, void* dx, int64_t lddx, float* dt,
                         int64_t lddt, int B, int H, int W, int C, void* stream);

/* ---- cross-GPU statistics exchange over NVLink peer memory (the collective inside torch.nn.SyncBatchNorm,
 *      train.py:97-98: forward all-gather of batch statistics, backward all-reduce of sum_dy / sum_dy_xmu).
 *      Each rank creates ONE buffer, ships its 64-byte CUDA-IPC handle to the other ranks of the node (any side
 *      channel; the Python surface uses torch.distributed's object all-gather), opens the peers' handles, and then
 *      cris_peer_allreduce_f32 sums an fp32 vector across ranks in ONE kernel (flag handshake + peer loads),
 *      CUDA-graph replayable.  `slot` identifies the exchange site: every rank must issue the same slot sequence. */
#define CRIS_PEER_MAX_WORLD 8
#define CRIS_PEER_MAX_SLOTS 1024
#define CRIS_PEER_SLOT_FLOATS 4096
#define CRIS_PEER_HANDLE_BYTES 64
size_t cris_peer_buffer_bytes(void);
int cris_peer_buffer_create(void** dev_ptr, unsigned char* handle_out /* [CRIS_PEER_HANDLE_BYTES] */);
int cris_peer_buffer_open(const unsigned char* handle, void** dev_ptr);
int cris_peer_buffer_close(void* dev_ptr, int owned);
/* peer_ptrs: HOST array of `world` device pointers (index = rank; entry `rank` is the own buffer).  out may alias in.
 * timeout_s <= 0 -> 120 s; a peer that never arrives makes the kernel print and trap instead of hanging. */
int cris_peer_allreduce_f32(void* const* peer_ptrs, int world, int rank, int slot, const float* in, float* out, int n,
                            double timeout_s, void* stream);
/* One SyncBatchNorm exchange site in ONE kernel: local partials [n_tiles][2][C] -> this rank's (sum, sumsq) ->
 * pushed into every peer's memory over NVLink -> rank-ordered global sums ->
 *   fwd: scale/shift/mean/invstd + running-statistics update from the GLOBAL statistics (global_count = elements per
 *        channel over all ranks) — torch.nn.SyncBatchNorm.forward (train.py:97-98)
 *   bwd: grad_beta / grad_gamma = LOCAL (sum dz, sum dz*xhat) (DDP averages them), sums_out[2C] = GLOBAL sums for
 *        cris_bn_bwd_apply — SyncBatchNorm.backward's all-reduce.
 * `site` is the exchange-site index of the pass (every rank issues the same sequence); 2*C <= CRIS_PEER_SLOT_FLOATS. */
int cris_peer_bn_sync_fwd(void* const* peer_ptrs, int world, int rank, int site, const float* partials, int n_tiles, int C,
                          double global_count, const float* gamma, const float* beta, float eps, float momentum,
                          float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                          double timeout_s, void* stream);
int cris_peer_bn_sync_bwd(void* const* peer_ptrs, int world, int rank, int site, const float* partials, int n_tiles, int C,
                          float* sums_out, float* grad_beta, float* grad_gamma, double timeout_s, void* stream);

/* ---- fused multi-head attention for 64-wide heads (csrc/attention.cu): O = dropout(softmax(alpha Q K^T)) V per
 *      (image, head), scores kept in tensor memory / shared memory (never written to HBM) — the core of
 *      F.multi_head_attention_forward as called at model/layers.py:233-237 (decoder self attention, 676 keys,
 *      attention dropout) and model/clip.py:119-139 (AttentionPool2d).  q [B*Lq, heads*64], k / v [B*Lk, heads*64]
 *      bf16 with their own pitches (column slices of packed projections); lse [B*heads*Lq] fp32 receives the row
 *      log2-sum-exp for the backward.  No masks: causal / key-padded attention (17 keys) stays on the
 *      cris_gemm + cris_softmax path.  Dropout masks are the same pure function of (seed + *seed_dev, element index
 *      ((b*heads+h)*Lq + q)*round8(Lk) + k) that cris_softmax_fwd uses.
 *      backward: d_scratch [B*heads*Lq] fp32 workspace; dq_acc fp32 [B*Lq, heads*64] must be ZEROED by the caller
 *      (the key blocks of one image/head add into it); dk, dv bf16 [B*Lk, heads*64] are overwritten. */
int cris_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                       float* lse, int B, int heads, int Lq, int Lk, float alpha, float p_drop, uint64_t seed,
                       const uint64_t* seed_dev, void* stream);
int cris_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                       int64_t ldo, const void* d_o, int64_t lddo, const float* lse, float* d_scratch, float* dq_acc,
                       int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int heads, int Lq, int Lk,
                       float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream);

/* ---- evaluation post-processing (SURVEY 8f "next" row 2; engine/engine.py:101-124,172-190): logits -> IoU against the
 *      ground truth of the ORIGINAL photo.  cris_postproc_upsample: prob_up[b] = bicubic(sigmoid(logits[b]), align_corners
 *      = True) (F.interpolate semantics), fp32 [B,H,W] -> [B,OH,OW].  cris_postproc_warp_iou: per sample the
 *      cv2.warpAffine(prob_up, m, (w,h), INTER_CUBIC, borderValue 0) of the reference (fixed-point 1/32-pixel
 *      coordinates, float cubic table), `> thr`, and counts[b] = {sum(pred & gt), sum(pred | gt)} (u64, zeroed by the
 *      caller).  samples_dev: DEVICE array of B records {double m[6]; int h, w; int64 off} (cris_postproc_sample_bytes()
 *      each; m is the matrix the reference passes to warpAffine, off the sample's offset into the packed uint8 buffers
 *      gt / pred_out); pred_out (optional) receives the binary prediction; max_pixels = max over samples of h*w. */
int cris_postproc_sample_bytes(void);
int cris_postproc_upsample(const float* logits, float* prob_up, int B, int H, int W, int OH, int OW, void* stream);
int cris_postproc_warp_iou(const float* prob_up, int B, int SH, int SW, const void* samples_dev, const uint8_t* gt,
                           uint8_t* pred_out, float thr, unsigned long long* counts, long long max_pixels, void* stream);

/* ---- input transform (SURVEY 8f "next" row 3; utils/dataset.py:148-163,210-221): decoded 8-bit photos -> the model's
 *      input.  Per sample cv2.warpAffine(photo RGB u8 [h,w,3], m, (OW,OH), INTER_CUBIC, borderValue = border3) with
 *      OpenCV's 8-bit fixed-point arithmetic (bit-exact), then ((v / 255) - mean) / std in fp32 -> img_out fp32
 *      [B,3,OH,OW]; optionally cv2.warpAffine(mask u8 [h,w], m, INTER_LINEAR, border 0) / 255 -> mask_out fp32 [B,OH,OW]
 *      (samples without a mask give zeros).  images / masks: DEVICE byte buffers holding every sample back to back;
 *      samples_dev: DEVICE array of B records {double m[6]; int h, w; int64 img_off; int64 mask_off (-1 = none)}
 *      (cris_feeder_sample_bytes() each).  border3 / mean3 / std3 are HOST arrays of three values. */
int cris_feeder_sample_bytes(void);
int cris_feeder_letterbox(const uint8_t* images, const uint8_t* masks, const void* samples_dev, int B, int OH, int OW,
                          const double* border3, const float* mean3, const float* std3, float* img_out, float* mask_out,
                          void* stream);

/* ---- optimizer step (SURVEY 8f "next" row: torch.optim.Adam driven by GradScaler, train.py:105-111,
 *      engine/engine.py:52-57).  One launch updates every tensor of a parameter group:
 *      g' = g / *grad_scale (+ weight_decay * p); m, v moments; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).
 *      table_dev: DEVICE array of n_tensors entries {float* p; const float* g; float* m; float* v; int64 n;
 *      int64 chunk0} where chunk0 is the exclusive prefix sum of ceil(n / cris_adam_chunk_elems()) and n_chunks the
 *      total.  `step` is the step count after this update (bias correction).  *found_inf != 0 skips the update
 *      (GradScaler semantics); either pointer may be NULL. */
int cris_adam_table_entry_bytes(void);
int cris_adam_chunk_elems(void);
int cris_adam_step(const void* table_dev, int n_tensors, long long n_chunks, double lr, double beta1, double beta2,
                   double eps, double weight_decay, double step, const float* grad_scale, const float* found_inf,
                   void* stream);

/* ---- EXPERIMENTAL: 3x3 convolution for 32/64-channel layers through one "halo" tile per output tile (stem conv2 /
 *      conv3, layer1 3x3 convs; model/clip.py:17-25,165-182).  Same result as cris_gemm in CRIS_TAP_ACCUM mode with
 *      the border mask: z[r][co] = sum_tap sum_ci x[r + off_tap][ci] * w[co][tap][ci] on interior rows, 0 on border
 *      rows; x, z padded NHWC bf16; w_packed = cris_pack_conv_weight layout [Cout][9][cin_pad]; optional BatchNorm
 *      column statistics accumulated like cris_gemm's colstats (caller-zeroed [min(64, ceil(rows/128))][2][Cout]). */
/* out[ci][t'][co] = w[co][ci][8 - t'] (bf16 [Cin][9][cout_pad]): the weights with which the data gradient of a 3x3
 * convolution is itself a forward 3x3 convolution of dz, dx = cris_conv3x3_halo(dz, out) with Cin/Cout swapped */
int cris_pack_conv_weight_dgrad(const float* w, void* out, int Cout, int Cin, int cout_pad, void* stream);
int cris_conv3x3_halo(const void* x, int64_t ldx, const void* w_packed, int64_t ldw, int cin_pad, void* z, int64_t ldz,
                      float* colstats, int N, int H, int W, int Cin, int Cout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRIS_B200_H_ */
