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
/* 0 = tcgen05 GEMM (product path), 1 = SIMT reference GEMM (differential testing only) */
void cris_set_gemm_impl(int impl);
int  cris_get_gemm_impl(void);
uint64_t cris_launch_count(void);           /* kernels launched by this library so far    */

/* ---- the GEMM / implicit-GEMM-conv core --------------------------------------------- */
enum { CRIS_ACT_NONE = 0, CRIS_ACT_RELU = 1, CRIS_ACT_QUICKGELU = 2 };
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
 * Epilogue order: acc*alpha -> +bias[n] -> act -> +resid -> row mask -> store / atomic add
 *   -> optional per-column (sum, sumsq) partials of the stored values into
 *      colstats[m_tile][2][N] (fp32), m_tile = m/128 — the BatchNorm batch statistics.
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
  float*  colstats;          /* [ceil(M/128)][2][N] fp32 or NULL */
  int32_t a_rows, b_rows;    /* physical row extents of A/B for OOB zero fill (0 = derive) */
} cris_gemm_args;

int cris_gemm(const cris_gemm_args* args, void* stream);

/* ---- normalisation -------------------------------------------------------------------- */
/*
 * BatchNorm (training: batch statistics; replaces nn.BatchNorm2d/1d + SyncBatchNorm,
 * model/clip.py:18-26,171-183, model/layers.py:8-16,262; torch/nn/modules/_functions.py).
 * cris_bn_finalize reduces colstats partials -> local (sum, sumsq) in stats[2][C];
 * after the optional cross-rank exchange the caller passes the global sums here again via
 * cris_bn_coeffs to get scale/shift and update running stats (momentum, unbiased var).
 */
int cris_bn_reduce_partials(const float* partials, int n_tiles, int C, float* sums /*[2][C]*/,
                            void* stream);
int cris_bn_coeffs(const float* sums /*[2][C]*/, double count, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var,
                   float* scale, float* shift, float* mean, float* invstd, int C, int training,
                   void* stream);
/* y = mask(act(x*scale[c] + shift[c] (+ resid)))   over a padded-NHWC row matrix [rows, C] */
int cris_bn_apply(const void* x, int64_t ldx, const float* scale, const float* shift,
                  const void* resid, int64_t ldr, void* y, int64_t ldy, int64_t rows, int C,
                  int relu, int hp, int wp, void* stream);
/* backward: dz = dy * (y > 0) if relu; sums[0][c] = sum dz, sums[1][c] = sum dz * xhat */
int cris_bn_bwd_reduce(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x,
                       int64_t ldx, const float* mean, const float* invstd, int64_t rows, int C,
                       int relu, float* partials /*[n_blocks][2][C]*/, int n_blocks, void* stream);
/* dx = gamma*invstd*(dz - sum_dz/count - xhat*sum_dzxhat/count); dres = dz (optional) */
int cris_bn_bwd_apply(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x,
                      int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                      const float* sums, double count, void* dx, int64_t lddx, void* dres,
                      int64_t lddres, int64_t rows, int C, int relu, int hp, int wp, void* stream);

/* LayerNorm over the last dim (nn.LayerNorm, model/clip.py:226-231; model/layers.py:199-216) */
int cris_layernorm_fwd(const void* x, int x_fp32, int64_t ldx, const float* gamma, const float* beta,
                       const void* add, int64_t ldadd, int add_period, void* y, int y_fp32, int64_t ldy,
                       void* y2, int64_t ldy2, float* mean, float* rstd, int64_t rows, int C, float eps,
                       void* stream);
int cris_layernorm_bwd(const void* dy, int dy_fp32, int64_t lddy, const void* dy2, int64_t lddy2,
                       const void* x, int x_fp32, int64_t ldx, const float* gamma, const float* mean,
                       const float* rstd, void* dx, int dx_fp32, int64_t lddx, int dx_accumulate,
                       float* dgamma_partials, float* dbeta_partials, int n_blocks, int64_t rows, int C,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRIS_B200_H_ */
