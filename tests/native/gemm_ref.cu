// gemm_ref.cu — SIMT restatement of cris_gemm (same argument semantics, one thread per output
// element).  TEST INFRASTRUCTURE: it is compiled only into tests/native/gemm_selftest (never into
// libcris_b200.so) and exists for differential testing of the tcgen05 kernel on the GPU.
#include "../../cris/pytorch_b200/csrc/common.cuh"

namespace cris {

struct RefArgs {
  cris_gemm_args a;
  long long a_rows, b_rows;
};

__device__ __forceinline__ float ref_act(float v, int act) {
  if (act == CRIS_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CRIS_ACT_QUICKGELU) return v / (1.f + __expf(-1.702f * v));
  return v;
}

__global__ void gemm_ref_kernel(const RefArgs r) {
  const cris_gemm_args& a = r.a;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  int z = blockIdx.z;
  const int taps_z = (a.tap_mode == CRIS_TAP_WGRAD) ? a.taps : 1;
  const int ztap = z % taps_z;
  const int batch = z / taps_z;
  if (m >= a.M || n >= a.N) return;
  const int bin = a.batch_inner > 1 ? a.batch_inner : 1;
  const long long b_in = batch % bin, b_out = batch / bin;
  const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(a.A) + b_out * a.strideA + b_in * a.strideA2;
  const __nv_bfloat16* B = reinterpret_cast<const __nv_bfloat16*>(a.B) + b_out * a.strideB + b_in * a.strideB2;
  const int ntl = (a.tap_mode == CRIS_TAP_ACCUM) ? a.taps : 1;
  float acc = 0.f;
  for (int tl = 0; tl < ntl; ++tl) {
    const int t = (a.tap_mode == CRIS_TAP_ACCUM) ? tl : ztap;
    const int a_row_off = (a.tap_mode == CRIS_TAP_ACCUM) ? a.tap_off[t] : 0;
    int b_k_off = 0, b_n_off = 0;
    if (a.tap_mode == CRIS_TAP_ACCUM) {
      b_k_off = t * a.b_tap_k;
      b_n_off = t * a.b_tap_n;
    } else if (a.tap_mode == CRIS_TAP_WGRAD) {
      b_k_off = a.tap_off[t];
    }
    for (int k = 0; k < a.K; ++k) {
      float av = 0.f, bv = 0.f;
      if (!a.a_mn) {
        const long long row = (long long)m + a_row_off;
        if (row >= 0 && row < r.a_rows) av = bf2f(A[row * a.lda + k]);
      } else {
        if (k < r.a_rows) av = bf2f(A[(long long)k * a.lda + m]);
      }
      if (!a.b_mn) {
        bv = bf2f(B[(long long)(n + b_n_off) * a.ldb + k + b_k_off]);
      } else {
        const long long kr = (long long)k + b_k_off;
        if (kr >= 0 && kr < r.b_rows) bv = bf2f(B[kr * a.ldb + n + b_n_off]);
      }
      acc = fmaf(av, bv, acc);
    }
  }
  float v = acc * a.alpha;
  if (a.bias) v += a.bias[n];
  v = ref_act(v, a.act);
  const int dcol = n + ((a.tap_mode == CRIS_TAP_WGRAD) ? ztap * a.d_tap_n : 0);
  if (a.resid) {
    const long long ri = b_out * a.strideR + b_in * a.strideR2 + (long long)m * a.ldr + dcol;
    v += a.resid_fp32 ? reinterpret_cast<const float*>(a.resid)[ri]
                      : bf2f(reinterpret_cast<const __nv_bfloat16*>(a.resid)[ri]);
  }
  if (!interior_row(m, a.mask_hp, a.mask_wp)) v = 0.f;
  long long di = b_out * a.strideD + b_in * a.strideD2 + (long long)m * a.ldd + dcol;
  if (a.accumulate && a.d_col_stride > 1)
    di = b_out * a.strideD + b_in * a.strideD2 + (long long)m * a.ldd + (long long)n * a.d_col_stride +
         ((a.tap_mode == CRIS_TAP_WGRAD) ? ztap * a.d_tap_n : 0);
  if (a.d_fp32) {
    float* D = reinterpret_cast<float*>(a.D);
    if (a.accumulate) atomicAdd(D + di, v);
    else D[di] = v;
  } else {
    reinterpret_cast<__nv_bfloat16*>(a.D)[di] = f2bf(v);
  }
}

// partials[m_tile % 64][2][N] (+)= column sums of the stored D values of a 128-row tile
__global__ void gemm_ref_colstats(const cris_gemm_args a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int mt = blockIdx.y;
  if (n >= a.N) return;
  float s0 = 0.f, s1 = 0.f;
  for (int i = 0; i < 128; ++i) {
    const long long m = (long long)mt * 128 + i;
    if (m >= a.M) break;
    const long long di = m * a.ldd + n;
    const float x = a.d_fp32 ? reinterpret_cast<const float*>(a.D)[di]
                             : bf2f(reinterpret_cast<const __nv_bfloat16*>(a.D)[di]);
    s0 += x;
    s1 += x * x;
  }
  atomicAdd(a.colstats + (size_t)(mt & 63) * 2 * a.N + n, s0);
  atomicAdd(a.colstats + (size_t)(mt & 63) * 2 * a.N + a.N + n, s1);
}

static int gemm_ref_launch(const cris_gemm_args* a, cudaStream_t stream) {
  RefArgs r;
  r.a = *a;
  r.a_rows = a->a_rows > 0 ? a->a_rows : (a->a_mn ? a->K : a->M);
  r.b_rows = a->b_rows > 0 ? a->b_rows : (a->b_mn ? a->K : a->N);
  const int taps_z = (a->tap_mode == CRIS_TAP_WGRAD) ? a->taps : 1;
  dim3 block(32, 8);
  dim3 grid((a->N + 31) / 32, (a->M + 7) / 8, a->batch * taps_z);
  gemm_ref_kernel<<<grid, block, 0, stream>>>(r);
  if (cudaGetLastError() != cudaSuccess) return -3;
  if (a->colstats) {
    dim3 g2((a->N + 127) / 128, (a->M + 127) / 128);
    gemm_ref_colstats<<<g2, 128, 0, stream>>>(*a);
    if (cudaGetLastError() != cudaSuccess) return -3;
  }
  return 0;
}

}  // namespace cris

extern "C" int cris_ref_gemm(const cris_gemm_args* args, void* stream) {
  return cris::gemm_ref_launch(args, reinterpret_cast<cudaStream_t>(stream));
}
