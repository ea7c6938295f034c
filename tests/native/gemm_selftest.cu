// gemm_selftest.cu — differential test of the tcgen05 GEMM core against the SIMT restatement
// (gemm_ref.cu) and, for small cases, a double-precision host loop.  Pure CUDA, no Python:
//   gemm_selftest list            -> number of cases
//   gemm_selftest <i>             -> run case i (exit 0 = pass)
//   gemm_selftest perf            -> timing of a few product-sized shapes
// Each case is run in its own process by tests/native/run_selftest.sh so that a trapped kernel
// cannot poison the following cases.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cris_b200.h"

extern "C" int cris_ref_gemm(const cris_gemm_args* args, void* stream);  // tests/native/gemm_ref.cu (test-only SIMT restatement)

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(3);                                                                     \
    }                                                                              \
  } while (0)

struct Case {
  std::string name;
  int M, N, K, batch, a_mn, b_mn;
  int tap_mode, taps, hp, wp;  // hp/wp: padded geometry used for tap offsets and the row mask
  int splits, d_fp32, accumulate;
  int bias, act, resid, resid_fp32, mask, colstats;
  float alpha;
  int cpu_check;
  int heads;  // > 0: two-level batch (outer = batch/heads images, inner = heads, head dim 64 slices of a [L, heads*64] matrix)
};

static std::vector<Case> make_cases() {
  std::vector<Case> c;
  auto add = [&](const char* name, int M, int N, int K, int batch, int a_mn, int b_mn) {
    Case x{};
    x.name = name; x.M = M; x.N = N; x.K = K; x.batch = batch; x.a_mn = a_mn; x.b_mn = b_mn;
    x.tap_mode = 0; x.taps = 1; x.splits = 1; x.alpha = 1.f;
    c.push_back(x);
    return &c.back();
  };
  // plain NT (both K-major)
  add("nt_128x128x64", 128, 128, 64, 1, 0, 0)->cpu_check = 1;
  add("nt_128x128x256", 128, 128, 256, 1, 0, 0)->cpu_check = 1;
  add("nt_300x200x136", 300, 200, 136, 1, 0, 0)->cpu_check = 1;
  add("nt_1088x1536x512", 1088, 1536, 512, 1, 0, 0);
  add("nt_fp32out_700x512x2048", 700, 512, 2048, 1, 0, 0)->d_fp32 = 1;
  { auto x = add("nt_bias_relu_640x256x512", 640, 256, 512, 1, 0, 0); x->bias = 1; x->act = 1; }
  { auto x = add("nt_bias_qgelu_resid_bf16", 384, 192, 128, 1, 0, 0); x->bias = 1; x->act = 2; x->resid = 1; }
  { auto x = add("nt_resid_fp32_out_fp32", 384, 512, 2048, 1, 0, 0); x->bias = 1; x->resid = 1; x->resid_fp32 = 1; x->d_fp32 = 1; }
  { auto x = add("nt_alpha_N17", 676, 17, 64, 3, 0, 0); x->alpha = 0.125f; x->cpu_check = 1; }
  add("nt_N64_tile", 500, 64, 320, 1, 0, 0);
  add("nt_N32_tile", 500, 32, 192, 1, 0, 0);
  add("nt_K32_N32", 1000, 32, 32, 1, 0, 0)->cpu_check = 1;
  add("nt_K32_N64", 1000, 64, 32, 1, 0, 0);
  add("nt_batched_676x676x64", 676, 676, 64, 5, 0, 0);
  add("nt_batched_17x17x64", 17, 17, 64, 7, 0, 0)->cpu_check = 1;
  // persistent scheduling: more tiles than SMs (several tiles per CTA, both TMEM accumulator buffers in flight)
  add("nt_many_tiles_20000x512x256", 20000, 512, 256, 1, 0, 0);
  { auto x = add("nt_many_tiles_bias_relu_resid", 30000, 192, 64, 1, 0, 0); x->bias = 1; x->act = 1; x->resid = 1; }
  { auto x = add("conv_fwd_many_tiles_stats", 40 * 30 * 30, 256, 128, 1, 0, 0); x->tap_mode = 1; x->taps = 9; x->hp = 30; x->wp = 30; x->mask = 1; x->colstats = 1; }
  { auto x = add("conv_wgrad_many_tiles", 256, 384, 40 * 30 * 30, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 30; x->wp = 30; x->d_fp32 = 1; x->accumulate = 1; x->splits = 6; }
  add("nt_batched_many_676x676x64", 676, 676, 64, 40, 0, 0);
  // two-level batch: per-head slices of packed [B, L, heads*64] projections (attention)
  { auto x = add("heads_qk_100x100x64", 100, 100, 64, 8, 0, 0); x->heads = 4; x->cpu_check = 1; x->alpha = 0.125f; }
  { auto x = add("heads_pv_100x64x100", 100, 64, 100, 8, 0, 1); x->heads = 4; x->cpu_check = 1; }
  { auto x = add("heads_dk_100x64x100", 100, 64, 100, 8, 1, 1); x->heads = 4; x->cpu_check = 1; }
  // A K-major, B MN-major (dgrad / P.V / dS.K)
  add("nn_128x128x64", 128, 128, 64, 1, 0, 1)->cpu_check = 1;
  add("nn_300x200x136", 300, 200, 136, 1, 0, 1)->cpu_check = 1;
  add("nn_pv_676x64x676_b", 676, 64, 676, 4, 0, 1);
  add("nn_pv_676x64x17_b", 676, 64, 17, 4, 0, 1)->cpu_check = 1;
  add("nn_dgrad_4000x512x2048", 4000, 512, 2048, 1, 0, 1);
  // A MN-major, B MN-major (wgrad / dK / dV)
  add("tn_128x128x64", 128, 128, 64, 1, 1, 1)->cpu_check = 1;
  add("tn_300x200x136", 300, 200, 136, 1, 1, 1)->cpu_check = 1;
  { auto x = add("tn_wgrad_lin_splitk", 512, 2048, 5000, 1, 1, 1); x->d_fp32 = 1; x->accumulate = 1; x->splits = 5; }
  add("tn_dk_676x64x676_b", 676, 64, 676, 3, 1, 1);
  add("tn_17x64x676_b", 17, 64, 676, 3, 1, 1);
  // conv: 9 taps accumulate, K-major (forward) on a padded 10x12 grid, 3 images
  { auto x = add("conv_fwd_c64_c128", 3 * 10 * 12, 128, 64, 1, 0, 0); x->tap_mode = 1; x->taps = 9; x->hp = 10; x->wp = 12; x->mask = 1; x->cpu_check = 1; }
  { auto x = add("conv_fwd_c256_c256_stats", 6 * 28 * 28, 256, 256, 1, 0, 0); x->tap_mode = 1; x->taps = 9; x->hp = 28; x->wp = 28; x->mask = 1; x->colstats = 1; }
  { auto x = add("conv_fwd_c32_c32_k32", 2 * 30 * 30, 32, 32, 1, 0, 0); x->tap_mode = 1; x->taps = 9; x->hp = 30; x->wp = 30; x->mask = 1; x->colstats = 1; }
  { auto x = add("conv_fwd_c32_c64_k32", 2 * 30 * 30, 64, 32, 1, 0, 0); x->tap_mode = 1; x->taps = 9; x->hp = 30; x->wp = 30; x->mask = 1; }
  { auto x = add("conv1x1_stats_mask", 4 * 15 * 15, 192, 320, 1, 0, 0); x->hp = 15; x->wp = 15; x->mask = 1; x->colstats = 1; x->cpu_check = 1; }
  // conv dgrad: taps accumulate with MN-major B (weights [Cout][tap][Cin])
  { auto x = add("conv_dgrad_c128_c64", 3 * 10 * 12, 64, 128, 1, 0, 1); x->tap_mode = 1; x->taps = 9; x->hp = 10; x->wp = 12; x->mask = 1; x->cpu_check = 1; }
  { auto x = add("conv_dgrad_c256_c256", 6 * 28 * 28, 256, 256, 1, 0, 1); x->tap_mode = 1; x->taps = 9; x->hp = 28; x->wp = 28; x->mask = 1; }
  // conv wgrad: one slab per tap, split-K, fp32 atomics
  { auto x = add("conv_wgrad_c128_c64", 128, 64, 3 * 10 * 12, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 10; x->wp = 12; x->d_fp32 = 1; x->accumulate = 1; x->splits = 2; x->cpu_check = 1; }
  { auto x = add("conv_wgrad_c256_c256", 256, 256, 6 * 28 * 28, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 28; x->wp = 28; x->d_fp32 = 1; x->accumulate = 1; x->splits = 4; }
  { auto x = add("conv_wgrad_auto_split", 256, 512, 8 * 26 * 26, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 26; x->wp = 26; x->d_fp32 = 1; x->accumulate = 1; x->splits = 0; }
  { auto x = add("tn_wgrad_lin_auto_split", 512, 640, 9000, 1, 1, 1); x->d_fp32 = 1; x->accumulate = 1; x->splits = 0; }
  // TMA-reduction epilogue: partial last chunk (N % 32 != 0 -> scalar atomics for that chunk only), N % 4 != 0, M < 128
  { auto x = add("conv_wgrad_c128_c72_partial_chunk", 128, 72, 3 * 10 * 12, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 10; x->wp = 12; x->d_fp32 = 1; x->accumulate = 1; x->splits = 2; x->cpu_check = 1; }
  { auto x = add("tn_wgrad_lin_N514", 200, 514, 3000, 1, 1, 1); x->d_fp32 = 1; x->accumulate = 1; x->splits = 3; x->cpu_check = 1; }
  { auto x = add("conv_wgrad_c64_c64_auto", 64, 64, 16 * 30 * 30, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 30; x->wp = 30; x->d_fp32 = 1; x->accumulate = 1; x->splits = 0; }
  { auto x = add("conv_wgrad_c32_c32", 32, 32, 2 * 30 * 30, 1, 1, 1); x->tap_mode = 2; x->taps = 9; x->hp = 30; x->wp = 30; x->d_fp32 = 1; x->accumulate = 1; x->splits = 3; }
  return c;
}

static uint32_t rng_state = 12345u;
static float frand() {  // [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static int round8(int x) { return (x + 7) / 8 * 8; }

struct Buffers {
  std::vector<__nv_bfloat16> hA, hB, hR16;
  std::vector<float> hbias, hR32;
  __nv_bfloat16 *dA = nullptr, *dB = nullptr;
  void* dR = nullptr;
  float* dbias = nullptr;
  void *dD_tc = nullptr, *dD_ref = nullptr;
  float *dS_tc = nullptr, *dS_ref = nullptr;
  size_t d_elems = 0, s_elems = 0;
  cris_gemm_args args{};
};

static void setup(const Case& cs, Buffers& b) {
  cris_gemm_args& a = b.args;
  memset(&a, 0, sizeof(a));
  const int taps = cs.tap_mode ? cs.taps : 1;
  a.M = cs.M; a.N = cs.N; a.K = cs.K; a.batch = cs.batch; a.a_mn = cs.a_mn; a.b_mn = cs.b_mn;
  a.tap_mode = cs.tap_mode; a.taps = taps; a.splits = cs.splits; a.alpha = cs.alpha;
  a.d_fp32 = cs.d_fp32; a.accumulate = cs.accumulate; a.act = cs.act;
  if (cs.tap_mode) {
    int t = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) a.tap_off[t++] = dy * cs.wp + dx;
  }
  // ---- A ----
  long long a_rows = cs.a_mn ? cs.K : cs.M;
  long long a_cols = cs.a_mn ? cs.M : cs.K;
  a.lda = round8((int)a_cols) + 8;
  a.strideA = a_rows * a.lda;
  b.hA.resize((size_t)a.strideA * cs.batch);
  for (auto& v : b.hA) v = __float2bfloat16(frand());
  // ---- B ----
  long long b_rows, b_cols;
  if (!cs.b_mn) { b_rows = cs.N; b_cols = (long long)cs.K * (cs.tap_mode == 1 ? taps : 1); a.b_tap_k = cs.K; }
  else { b_rows = cs.K; b_cols = (long long)cs.N * (cs.tap_mode == 1 ? taps : 1); a.b_tap_n = cs.N; }
  if (cs.tap_mode != 1) { a.b_tap_k = 0; a.b_tap_n = 0; }
  a.ldb = round8((int)b_cols) + (cs.tap_mode ? 0 : 8);
  a.strideB = b_rows * a.ldb;
  b.hB.resize((size_t)a.strideB * cs.batch);
  for (auto& v : b.hB) v = __float2bfloat16(frand() * 0.5f);
  // For conv modes the activation operand must carry a zero border (as the product keeps it).
  if (cs.tap_mode == 1 || cs.mask) {
    for (long long r = 0; r < cs.M && !cs.a_mn; ++r) {
      int rr = (int)(r % (cs.hp * cs.wp)); int h = rr / cs.wp, w = rr % cs.wp;
      if (h == 0 || h == cs.hp - 1 || w == 0 || w == cs.wp - 1)
        for (int k = 0; k < a.lda; ++k) b.hA[r * a.lda + k] = __float2bfloat16(0.f);
    }
  }
  if (cs.tap_mode == 2) {  // both operands are [pixel rows][channels]; zero their border rows
    for (long long r = 0; r < cs.K; ++r) {
      int rr = (int)(r % (cs.hp * cs.wp)); int h = rr / cs.wp, w = rr % cs.wp;
      if (h == 0 || h == cs.hp - 1 || w == 0 || w == cs.wp - 1) {
        for (int k = 0; k < a.lda; ++k) b.hA[r * a.lda + k] = __float2bfloat16(0.f);
        for (int k = 0; k < a.ldb; ++k) b.hB[r * a.ldb + k] = __float2bfloat16(0.f);
      }
    }
  }
  // ---- D ----
  const int d_cols = cs.N * (cs.tap_mode == 2 ? taps : 1);
  a.d_tap_n = cs.N;
  a.ldd = round8(d_cols) + (cs.tap_mode == 2 ? 0 : 8);
  a.strideD = (long long)cs.M * a.ldd;
  if (cs.heads > 0) {
    // operands that are "per-head 64-wide column slices" get inner stride 64; score-like operands stay dense
    const int H = cs.heads, Bo = cs.batch / H;
    a.batch_inner = H;
    auto sliced = [&](bool mn, int inner_dim) { return inner_dim == 64; };
    // A: K-major -> inner dim K; MN-major -> inner dim M
    const int a_inner = cs.a_mn ? cs.M : cs.K, a_outer = cs.a_mn ? cs.K : cs.M;
    if (sliced(cs.a_mn, a_inner)) { a.lda = H * 64; a.strideA2 = 64; a.strideA = (long long)a_outer * a.lda; b.hA.assign((size_t)a.strideA * Bo, __float2bfloat16(0.f)); }
    else { a.strideA2 = a.strideA; a.strideA = a.strideA2 * H; }
    const int b_inner = cs.b_mn ? cs.N : cs.K, b_outer = cs.b_mn ? cs.K : cs.N;
    if (sliced(cs.b_mn, b_inner)) { a.ldb = H * 64; a.strideB2 = 64; a.strideB = (long long)b_outer * a.ldb; b.hB.assign((size_t)a.strideB * Bo, __float2bfloat16(0.f)); }
    else { a.strideB2 = a.strideB; a.strideB = a.strideB2 * H; }
    for (auto& v : b.hA) v = __float2bfloat16(frand());
    for (auto& v : b.hB) v = __float2bfloat16(frand() * 0.5f);
    if (cs.N == 64) { a.ldd = H * 64; a.strideD2 = 64; a.strideD = (long long)cs.M * a.ldd; b.d_elems = (size_t)a.strideD * Bo; }
    else { a.strideD2 = a.strideD; a.strideD = a.strideD2 * H; b.d_elems = (size_t)a.strideD * Bo; }
  } else
  b.d_elems = (size_t)a.strideD * cs.batch;
  const size_t dbytes = b.d_elems * (cs.d_fp32 ? 4 : 2);
  CK(cudaMalloc(&b.dD_tc, dbytes)); CK(cudaMalloc(&b.dD_ref, dbytes));
  CK(cudaMemset(b.dD_tc, 0, dbytes)); CK(cudaMemset(b.dD_ref, 0, dbytes));
  CK(cudaMalloc(&b.dA, b.hA.size() * 2)); CK(cudaMalloc(&b.dB, b.hB.size() * 2));
  CK(cudaMemcpy(b.dA, b.hA.data(), b.hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b.dB, b.hB.data(), b.hB.size() * 2, cudaMemcpyHostToDevice));
  a.A = b.dA; a.B = b.dB;
  if (cs.bias) {
    b.hbias.resize(cs.N);
    for (auto& v : b.hbias) v = frand();
    CK(cudaMalloc(&b.dbias, cs.N * 4));
    CK(cudaMemcpy(b.dbias, b.hbias.data(), cs.N * 4, cudaMemcpyHostToDevice));
    a.bias = b.dbias;
  }
  if (cs.resid) {
    a.ldr = a.ldd; a.strideR = a.strideD; a.resid_fp32 = cs.resid_fp32;
    if (cs.resid_fp32) {
      b.hR32.resize(b.d_elems);
      for (auto& v : b.hR32) v = frand();
      CK(cudaMalloc(&b.dR, b.d_elems * 4));
      CK(cudaMemcpy(b.dR, b.hR32.data(), b.d_elems * 4, cudaMemcpyHostToDevice));
    } else {
      b.hR16.resize(b.d_elems);
      for (auto& v : b.hR16) v = __float2bfloat16(frand());
      CK(cudaMalloc(&b.dR, b.d_elems * 2));
      CK(cudaMemcpy(b.dR, b.hR16.data(), b.d_elems * 2, cudaMemcpyHostToDevice));
    }
    a.resid = b.dR;
  }
  if (cs.mask) { a.mask_hp = cs.hp; a.mask_wp = cs.wp; }
  if (cs.colstats) {
    b.s_elems = (size_t)((cs.M + 127) / 128) * 2 * cs.N;
    CK(cudaMalloc(&b.dS_tc, b.s_elems * 4)); CK(cudaMalloc(&b.dS_ref, b.s_elems * 4));
    CK(cudaMemset(b.dS_tc, 0, b.s_elems * 4)); CK(cudaMemset(b.dS_ref, 0, b.s_elems * 4));
  }
}

static double host_value(const Case& cs, const Buffers& b, int batch, int m, int n, int ztap) {
  const cris_gemm_args& a = b.args;
  const int bin = a.batch_inner > 1 ? a.batch_inner : 1;
  const size_t bi = batch % bin, bo = batch / bin;
  const __nv_bfloat16* A = b.hA.data() + bo * a.strideA + bi * a.strideA2;
  const __nv_bfloat16* B = b.hB.data() + bo * a.strideB + bi * a.strideB2;
  const long long a_rows = cs.a_mn ? cs.K : cs.M, b_rows = cs.b_mn ? cs.K : cs.N;
  double acc = 0;
  const int ntl = cs.tap_mode == 1 ? a.taps : 1;
  for (int tl = 0; tl < ntl; ++tl) {
    const int t = cs.tap_mode == 1 ? tl : ztap;
    const int aoff = cs.tap_mode == 1 ? a.tap_off[t] : 0;
    int bk = 0, bn = 0;
    if (cs.tap_mode == 1) { bk = t * a.b_tap_k; bn = t * a.b_tap_n; }
    if (cs.tap_mode == 2) bk = a.tap_off[t];
    for (int k = 0; k < cs.K; ++k) {
      double av = 0, bv = 0;
      if (!cs.a_mn) { long long r = (long long)m + aoff; if (r >= 0 && r < a_rows) av = __bfloat162float(A[r * a.lda + k]); }
      else av = __bfloat162float(A[(long long)k * a.lda + m]);
      if (!cs.b_mn) bv = __bfloat162float(B[(long long)(n + bn) * a.ldb + k + bk]);
      else { long long kr = (long long)k + bk; if (kr >= 0 && kr < b_rows) bv = __bfloat162float(B[kr * a.ldb + n + bn]); }
      acc += av * bv;
    }
  }
  double v = acc * cs.alpha;
  if (cs.bias) v += b.hbias[n];
  if (cs.act == 1) v = v > 0 ? v : 0;
  if (cs.act == 2) v = v / (1.0 + exp(-1.702 * v));
  const int dcol = n + (cs.tap_mode == 2 ? ztap * a.d_tap_n : 0);
  if (cs.resid) {
    size_t ri = (size_t)batch * a.strideR + (size_t)m * a.ldr + dcol;
    v += cs.resid_fp32 ? b.hR32[ri] : __bfloat162float(b.hR16[ri]);
  }
  if (cs.mask) {
    int rr = m % (cs.hp * cs.wp); int h = rr / cs.wp, w = rr % cs.wp;
    if (h == 0 || h == cs.hp - 1 || w == 0 || w == cs.wp - 1) v = 0;
  }
  return v;
}

static int run_case(const Case& cs) {
  Buffers b;
  setup(cs, b);
  cris_gemm_args a = b.args;
  // tcgen05 path
  a.D = b.dD_tc; a.colstats = b.dS_tc;
  if (cris_gemm(&a, nullptr) != 0) { printf("  tc launch error: %s\n", cris_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  tc kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  // SIMT restatement
  a.D = b.dD_ref; a.colstats = b.dS_ref;
  if (cris_ref_gemm(&a, nullptr) != 0) { printf("  ref launch error\n"); return 1; }
  CK(cudaDeviceSynchronize());

  const size_t n = b.d_elems;
  std::vector<float> tc(n), rf(n);
  if (cs.d_fp32) {
    CK(cudaMemcpy(tc.data(), b.dD_tc, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rf.data(), b.dD_ref, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> t16(n), r16(n);
    CK(cudaMemcpy(t16.data(), b.dD_tc, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r16.data(), b.dD_ref, n * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) { tc[i] = __bfloat162float(t16[i]); rf[i] = __bfloat162float(r16[i]); }
  }
  double max_ref = 0, max_diff = 0; size_t worst = 0; size_t nan_cnt = 0;
  for (size_t i = 0; i < n; ++i) {
    if (isnan(tc[i]) || isinf(tc[i])) { ++nan_cnt; continue; }
    max_ref = fmax(max_ref, fabs(rf[i]));
    double d = fabs((double)tc[i] - rf[i]);
    if (d > max_diff) { max_diff = d; worst = i; }
  }
  const double tol = (cs.d_fp32 ? 2e-4 : 1.0 / 96) * fmax(max_ref, 1e-3);
  int fail = 0;
  if (nan_cnt) { printf("  %zu non-finite outputs\n", nan_cnt); fail = 1; }
  if (max_diff > tol) {
    printf("  tc vs ref: max|diff| %.5g > tol %.5g at flat index %zu (row %lld col %lld) tc=%g ref=%g\n", max_diff, tol,
           worst, (long long)((worst % (size_t)b.args.strideD) / b.args.ldd), (long long)(worst % b.args.ldd), tc[worst], rf[worst]);
    // print a small error map over 32x32 blocks to help localise layout bugs
    const int bm = (cs.M + 31) / 32, bn = (int)((b.args.ldd + 31) / 32);
    printf("  error map (rows/32 x cols/32, '#' = block has mismatch), batch 0:\n");
    for (int i = 0; i < bm && i < 24; ++i) {
      printf("   ");
      for (int j = 0; j < bn && j < 64; ++j) {
        bool bad = false;
        for (int r = i * 32; r < (i + 1) * 32 && r < cs.M && !bad; ++r)
          for (int c2 = j * 32; c2 < (j + 1) * 32 && c2 < b.args.ldd; ++c2) {
            size_t idx = (size_t)r * b.args.ldd + c2;
            if (fabs((double)tc[idx] - rf[idx]) > tol) { bad = true; break; }
          }
        printf("%c", bad ? '#' : '.');
      }
      printf("\n");
    }
    fail = 1;
  }
  if (cs.colstats) {
    std::vector<float> st(b.s_elems), sr(b.s_elems);
    CK(cudaMemcpy(st.data(), b.dS_tc, b.s_elems * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sr.data(), b.dS_ref, b.s_elems * 4, cudaMemcpyDeviceToHost));
    double ms = 0, md = 0;
    for (size_t i = 0; i < b.s_elems; ++i) { ms = fmax(ms, fabs(sr[i])); md = fmax(md, fabs((double)st[i] - sr[i])); }
    if (md > 2e-2 * fmax(ms, 1e-3)) { printf("  colstats: max|diff| %.5g (scale %.5g)\n", md, ms); fail = 1; }
    else printf("  colstats ok: max|diff| %.3g (scale %.3g)\n", md, ms);
  }
  if (cs.cpu_check) {  // validate the SIMT restatement itself on a sample of outputs
    double md = 0, mr = 0;
    rng_state = 777;
    const int taps_z = cs.tap_mode == 2 ? cs.taps : 1;
    for (int s = 0; s < 400; ++s) {
      int bt = (int)((frand() * 0.5f + 0.5f) * cs.batch) % cs.batch;
      int m = (int)((frand() * 0.5f + 0.5f) * cs.M) % cs.M;
      int nn = (int)((frand() * 0.5f + 0.5f) * cs.N) % cs.N;
      int zt = (int)((frand() * 0.5f + 0.5f) * taps_z) % taps_z;
      double hv = host_value(cs, b, bt, m, nn, zt);
      const int bin2 = b.args.batch_inner > 1 ? b.args.batch_inner : 1;
      size_t di = (size_t)(bt / bin2) * b.args.strideD + (size_t)(bt % bin2) * b.args.strideD2 + (size_t)m * b.args.ldd + nn + (cs.tap_mode == 2 ? zt * b.args.d_tap_n : 0);
      md = fmax(md, fabs(hv - rf[di])); mr = fmax(mr, fabs(hv));
    }
    const double htol = (cs.d_fp32 ? 1e-3 : 1.0 / 96) * fmax(mr, 1e-3);
    if (md > htol) { printf("  ref vs host: max|diff| %.5g > %.5g\n", md, htol); fail = 1; }
    else printf("  ref vs host ok: max|diff| %.3g (scale %.3g)\n", md, mr);
  }
  printf("%s %s  M=%d N=%d K=%d b=%d amn=%d bmn=%d tap=%d  max|diff|=%.4g scale=%.4g\n", fail ? "FAIL" : "PASS",
         cs.name.c_str(), cs.M, cs.N, cs.K, cs.batch, cs.a_mn, cs.b_mn, cs.tap_mode, max_diff, max_ref);
  return fail;
}

static void perf(int only = -1) {
  struct P { const char* name; int M, N, K, taps, hp, wp, a_mn, b_mn; };  // a_mn && b_mn && taps > 1: tap-mode wgrad
  const P ps[] = {
      {"lin 43264x512x512", 43264, 512, 512, 1, 0, 0, 0, 0},
      {"lin 43264x2048x512", 43264, 2048, 512, 1, 0, 0, 0, 0},
      {"gemm 8192^3", 8192, 8192, 8192, 1, 0, 0, 0, 0},
      {"conv3x3 64x(106x106) 512->256", 64 * 106 * 106, 256, 512, 9, 106, 106, 0, 0},
      {"conv3x3 64x(54x54) 512->512", 64 * 54 * 54, 512, 512, 9, 54, 54, 0, 0},
      {"conv3x3 64x(106x106) 64->64", 64 * 106 * 106, 64, 64, 9, 106, 106, 0, 0},
      {"conv1x1 64x(106x106) 64->256", 64 * 106 * 106, 256, 64, 1, 106, 106, 0, 0},
      {"conv1x1 64x(106x106) 256->64", 64 * 106 * 106, 64, 256, 1, 106, 106, 0, 0},
      {"dgrad 43264x512x2048 (B MN)", 43264, 512, 2048, 1, 0, 0, 0, 1},
      {"wgrad 512x512x43264 (A,B MN)", 512, 512, 43264, 1, 0, 0, 1, 1},
      {"wgrad 256x512x719104 (A,B MN)", 256, 512, 719104, 1, 0, 0, 1, 1},
      {"wgrad3x3 256x512 64x(106x106)", 256, 512, 64 * 106 * 106, 9, 106, 106, 1, 1},
      {"wgrad3x3 512x512 64x(28x28)", 512, 512, 64 * 28 * 28, 9, 28, 28, 1, 1},
      {"wgrad3x3 64x64 64x(106x106)", 64, 64, 64 * 106 * 106, 9, 106, 106, 1, 1},
      {"wgrad1x1 2048x512 64x(15x15)", 2048, 512, 64 * 15 * 15, 1, 0, 0, 1, 1},
      {"dgrad3x3 64x(106x106) 256->512", 64 * 106 * 106, 512, 256, 9, 106, 106, 0, 1},
      {"wgrad3x3 32x32 64x(210x210)", 32, 32, 64 * 210 * 210, 9, 210, 210, 1, 1},
      {"wgrad3x3 64x32 64x(210x210)", 64, 32, 64 * 210 * 210, 9, 210, 210, 1, 1},
  };
  int pi = -1;
  for (const P& q : ps) {
    ++pi;
    if (only >= 0 && pi != only) continue;
    cris_gemm_args a; memset(&a, 0, sizeof(a));
    a.M = q.M; a.N = q.N; a.K = q.K; a.batch = 1; a.alpha = 1.f; a.splits = 1; a.a_mn = q.a_mn; a.b_mn = q.b_mn;
    int wg = (q.a_mn && q.b_mn);
    a.taps = q.taps; a.tap_mode = q.taps > 1 ? (wg ? 2 : 1) : 0;
    if (q.taps > 1) {
      int t = 0; for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) a.tap_off[t++] = dy * q.wp + dx;
      if (wg) a.d_tap_n = q.N; else if (q.b_mn) a.b_tap_n = q.N; else a.b_tap_k = q.K;
    }
    if (q.hp && !wg) { a.mask_hp = q.hp; a.mask_wp = q.wp; }
    long long a_rows = q.a_mn ? q.K : q.M, a_cols = q.a_mn ? q.M : q.K;
    long long b_rows = q.b_mn ? q.K : q.N, b_cols = (q.b_mn ? (long long)q.N * (wg ? 1 : q.taps) : (long long)q.K * q.taps);
    const int dtaps = (wg && q.taps > 1) ? q.taps : 1;
    a.lda = a_cols; a.ldb = b_cols; a.ldd = (long long)q.N * dtaps;
    void *A, *B, *D;
    CK(cudaMalloc(&A, a_rows * a_cols * 2)); CK(cudaMalloc(&B, b_rows * b_cols * 2)); CK(cudaMalloc(&D, (size_t)q.M * q.N * dtaps * (wg ? 4 : 2)));
    CK(cudaMemset(A, 0x3c, a_rows * a_cols * 2)); CK(cudaMemset(B, 0x3c, b_rows * b_cols * 2)); CK(cudaMemset(D, 0, (size_t)q.M * q.N * dtaps * (wg ? 4 : 2)));
    a.A = A; a.B = B; a.D = D;
    if (wg) { a.d_fp32 = 1; a.accumulate = 1; a.splits = 0; }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) if (cris_gemm(&a, nullptr)) { printf("launch error %s\n", cris_last_error()); return; }
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) cris_gemm(&a, nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    double fl = 2.0 * q.M * q.N * (double)q.K * q.taps;
    printf("PERF %-36s %8.3f ms  %8.1f TFLOP/s\n", q.name, ms, fl / ms * 1e-9);
    if (getenv("CRIS_B200_TRACE")) {
      long long* tr; CK(cudaMalloc(&tr, 64 * 16 * 8)); CK(cudaMemset(tr, 0, 64 * 16 * 8));
      cris_debug_set_trace(tr);
      cris_gemm(&a, nullptr); CK(cudaDeviceSynchronize());
      cris_debug_set_trace(nullptr);
      long long h[64 * 16]; CK(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
      const long long t0 = h[0];
      printf("  epilogue warp 0 timeline (SM cycles): tile | slot:cycle (0 = tile start, 14 = tile end)\n");
      for (int j = 0; j < 10; ++j) {
        printf("  %3d |", j);
        for (int k2 = 0; k2 < 15; ++k2) if (h[j * 16 + k2]) printf(" %d:%lld", k2, h[j * 16 + k2] - t0);
        printf("\n");
      }
      cudaFree(tr);
    }
    cudaFree(A); cudaFree(B); cudaFree(D);
  }
}

int main(int argc, char** argv) {
  auto cases = make_cases();
  if (argc < 2 || !strcmp(argv[1], "list")) { printf("%zu\n", cases.size()); return 0; }
  if (cris_device_check() != 0) { printf("device check failed: %s\n", cris_last_error()); return 2; }
  if (!strcmp(argv[1], "perf")) { perf(argc > 2 ? atoi(argv[2]) : -1); return 0; }
  int i = atoi(argv[1]);
  if (i < 0 || i >= (int)cases.size()) return 2;
  return run_case(cases[i]);
}
