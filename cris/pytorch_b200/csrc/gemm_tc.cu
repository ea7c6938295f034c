// gemm_tc.cu — the tcgen05 GEMM / implicit-GEMM-convolution core of libcris_b200.
//
//   D[b][m][n] (+)= alpha * sum_k A[b][m][k] * B[b][n][k]        bf16 x bf16 -> fp32 (TMEM)
//
// One CTA = one 128 x BN output tile.  Warp roles (192 threads):
//   warps 0-3 : epilogue  (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
//   warp  4   : TMA producer (cp.async.bulk.tensor into a STAGES-deep smem ring, mbarrier tx)
//   warp  5   : TMEM allocator + single-thread tcgen05.mma issuer (tcgen05.commit frees slots)
// Two CTAs are co-resident per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
//
// 3x3 convolution is an implicit GEMM over the zero-bordered ("padded NHWC") row matrix: tap
// (dy,dx) is the same A matrix shifted by dy*(W+2)+dx rows, so a conv is 9 x (K/BK) k-blocks
// accumulated in one TMEM tile, every load a plain 2-D TMA box (out-of-range rows zero-fill).
// Reference ops this replaces: nn.Conv2d / nn.Linear / MHA projections / bmm and their
// backward (model/clip.py:17-25,119-139,165-182,246-260; model/layers.py:8-16,202-212).
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace cris {

struct GemmKArgs {
  int M, N, K;
  int nkb;  // k-blocks per tap
  int tap_mode, taps, splits, taps_z;
  int tap_off[9];
  int b_tap_k, b_tap_n, d_tap_n;
  void* D;
  long long ldd, strideD, strideD2;
  int batch_inner;
  int d_fp32, accumulate;
  float alpha;
  const float* bias;
  int act;
  const void* resid;
  long long ldr, strideR, strideR2;
  int resid_fp32;
  int mask_hp, mask_wp;
  float* colstats;
  int d_col_stride;
};

constexpr int BM = 128;
constexpr int GEMM_THREADS = 192;

template <int BN, int BK>
struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = 98304 / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CRIS_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CRIS_ACT_QUICKGELU) return v / (1.f + __expf(-1.702f * v));
  return v;
}

template <int BN, int BK, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ GemmKArgs p) {
  using Cfg = TileCfg<BN, BK>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  __shared__ float s_stats[4][2][BN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  int z = blockIdx.z;
  const int split = z % p.splits;
  z /= p.splits;
  const int ztap = z % p.taps_z;
  const int batch = z / p.taps_z;
  const int b_in = batch % p.batch_inner, b_out = batch / p.batch_inner;

  // k-block schedule of this CTA
  const int kb_per_split = (p.nkb + p.splits - 1) / p.splits;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(p.nkb, kb_begin + kb_per_split);
  const int kb_cnt = max(0, kb_end - kb_begin);
  const int n_taps_loop = (p.tap_mode == CRIS_TAP_ACCUM) ? p.taps : 1;
  const int total_iters = n_taps_loop * kb_cnt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 5) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int t = (p.tap_mode == CRIS_TAP_ACCUM) ? (it / kb_cnt) : ztap;
        const int kb = kb_begin + (it % kb_cnt);
        const int k = kb * BK;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u, 100 + s);
        ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const int a_row_off = (p.tap_mode == CRIS_TAP_ACCUM) ? p.tap_off[t] : 0;
        int b_k_off = 0, b_n_off = 0;
        if (p.tap_mode == CRIS_TAP_ACCUM) {
          b_k_off = t * p.b_tap_k;
          b_n_off = t * p.b_tap_n;
        } else if (p.tap_mode == CRIS_TAP_WGRAD) {
          b_k_off = p.tap_off[t];
        }
        if constexpr (!A_MN) {
          ptx::tma_load_4d(sa, &tmA, &full_bar[s], k, m0 + a_row_off, b_in, b_out);
        } else {
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            ptx::tma_load_4d(sa + i * (BK * 128), &tmA, &full_bar[s], m0 + 64 * i, k, b_in, b_out);
        }
        if constexpr (!B_MN) {
          ptx::tma_load_4d(sb, &tmB, &full_bar[s], k + b_k_off, n0 + b_n_off, b_in, b_out);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            ptx::tma_load_4d(sb + j * (BK * 128), &tmB, &full_bar[s], n0 + b_n_off + 64 * j, k + b_k_off,
                             b_in, b_out);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t k_layout = (BK == 64) ? ptx::kLayoutSW128 : ptx::kLayoutSW64;
      constexpr uint32_t k_sbo = (BK == 64) ? 1024u : 512u;
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        ptx::mbar_wait(&full_bar[s], ph, 200 + s);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t adesc = A_MN ? ptx::make_smem_desc(sa, BK * 128, 1024, ptx::kLayoutSW128)
                                    : ptx::make_smem_desc(sa, 16, k_sbo, k_layout);
        const uint64_t bdesc = B_MN ? ptx::make_smem_desc(sb, BK * 128, 1024, ptx::kLayoutSW128)
                                    : ptx::make_smem_desc(sb, 16, k_sbo, k_layout);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          // advance 16 k-elements: K-major = 32 B inside the swizzled row; MN-major = 16 rows of 128 B
          const uint64_t a_adv = (uint64_t)((A_MN ? kk * 2048 : kk * 32) >> 4);
          const uint64_t b_adv = (uint64_t)((B_MN ? kk * 2048 : kk * 32) >> 4);
          ptx::umma_bf16(tmem_base, adesc + a_adv, bdesc + b_adv, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      ptx::umma_commit(tmem_full);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 0-3) =====================
    const long long row = (long long)m0 + warp * 32 + lane;
    const bool row_in = row < p.M;
    const bool row_valid = row_in && interior_row(row, p.mask_hp, p.mask_wp);
    const int dcol0 = n0 + ((p.tap_mode == CRIS_TAP_WGRAD) ? ztap * p.d_tap_n : 0);
    uint8_t* Dbase = reinterpret_cast<uint8_t*>(p.D);
    const long long drow = (long long)b_out * p.strideD + (long long)b_in * p.strideD2 + row * p.ldd;
    const long long rrow = (long long)b_out * p.strideR + (long long)b_in * p.strideR2 + row * p.ldr;
    if (total_iters > 0) {
      ptx::mbar_wait(tmem_full, 0, 300);
      ptx::tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      if (total_iters > 0) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const int ncol0 = n0 + c * 32;  // logical column (bias / N bound)
      if (ncol0 >= p.N) {
        if (p.colstats != nullptr) {
          s_stats[warp][0][c * 32 + lane] = 0.f;
          s_stats[warp][1][c * 32 + lane] = 0.f;
        }
        continue;
      }
      if (p.bias != nullptr) {
        const float* bp = p.bias + ncol0;
        if (ncol0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(bp + j));
            v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ncol0 + j < p.N) v[j] += __ldg(bp + j);
        }
      }
      if (p.act != CRIS_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
      }
      if (p.resid != nullptr && row_in) {
        if (p.resid_fp32) {
          const float* rp = reinterpret_cast<const float*>(p.resid) + rrow + dcol0 + c * 32;
          if (ncol0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = *reinterpret_cast<const float4*>(rp + j);
              v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol0 + j < p.N) v[j] += rp[j];
          }
        } else {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.resid) + rrow + dcol0 + c * 32;
          if (ncol0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
              const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c2 = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
              v[j] += a.x; v[j + 1] += a.y; v[j + 2] += b.x; v[j + 3] += b.y;
              v[j + 4] += c2.x; v[j + 5] += c2.y; v[j + 6] += d.x; v[j + 7] += d.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol0 + j < p.N) v[j] += bf2f(rp[j]);
          }
        }
      }
      if (!row_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      // ---- store ----
      if (row_in) {
        if (p.d_fp32) {
          float* dp = reinterpret_cast<float*>(Dbase) + drow + dcol0 + c * 32;
          if (p.accumulate) {
            float* ap = reinterpret_cast<float*>(Dbase) + drow +
                        ((p.tap_mode == CRIS_TAP_WGRAD) ? ztap * p.d_tap_n : 0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol0 + j < p.N) atomicAdd(ap + (long long)(ncol0 + j) * p.d_col_stride, v[j]);
          } else if (ncol0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol0 + j < p.N) dp[j] = v[j];
          }
        } else {
          __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(Dbase) + drow + dcol0 + c * 32;
          if (ncol0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 q;
              q.x = pack_bf16x2(v[j], v[j + 1]);
              q.y = pack_bf16x2(v[j + 2], v[j + 3]);
              q.z = pack_bf16x2(v[j + 4], v[j + 5]);
              q.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dp + j) = q;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol0 + j < p.N) dp[j] = f2bf(v[j]);
          }
        }
      }
      // ---- per-column batch statistics of the STORED (rounded) values ----
      if (p.colstats != nullptr) {
        float a[32], q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = p.d_fp32 ? v[j] : bf2f(f2bf(v[j]));
          if (!(ncol0 + j < p.N)) x = 0.f;
          a[j] = x;
          q[j] = x * x;
        }
        // transposing butterfly: after the 5 steps lane L holds the 32-lane total of column L
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool up = (lane & s) != 0;
#pragma unroll
          for (int i = 0; i < s; ++i) {
            const float send_a = up ? a[i] : a[i + s];
            const float keep_a = up ? a[i + s] : a[i];
            a[i] = keep_a + __shfl_xor_sync(0xffffffffu, send_a, s);
            const float send_q = up ? q[i] : q[i + s];
            const float keep_q = up ? q[i + s] : q[i];
            q[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, s);
          }
        }
        s_stats[warp][0][c * 32 + lane] = a[0];
        s_stats[warp][1][c * 32 + lane] = q[0];
      }
    }
    if (p.colstats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = threadIdx.x; j < BN; j += 128) {
        if (n0 + j < p.N) {
          const float s0 = s_stats[0][0][j] + s_stats[1][0][j] + s_stats[2][0][j] + s_stats[3][0][j];
          const float s1 = s_stats[0][1][j] + s_stats[1][1][j] + s_stats[2][1][j] + s_stats[3][1][j];
          float* dst = p.colstats + (size_t)blockIdx.y * 2 * p.N;
          dst[n0 + j] = s0;
          dst[p.N + n0 + j] = s1;
        }
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =========================== host side ===========================================
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

// operand stored as [outer_extent][inner_extent] bf16 rows with pitch ld (elements) and batch stride
static int make_tmap(CUtensorMap* tm, const void* base, long long inner_extent, long long outer_extent,
                     long long ld, long long batch, long long batch_stride, long long batch_in,
                     long long batch_in_stride, int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  auto fn = get_encode_fn();
  CRIS_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CRIS_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base not 16B aligned");
  CRIS_CHECK_ARG((ld * 2) % 16 == 0, "GEMM operand pitch %lld elements is not a multiple of 16 bytes", ld);
  if (batch_in <= 1) {
    batch_in = 1;
    batch_in_stride = ld * (outer_extent > 0 ? outer_extent : 1);
  }
  const long long batch_out = batch / batch_in;
  if (batch_out <= 1) batch_stride = batch_in_stride * batch_in;
  CRIS_CHECK_ARG((batch_stride * 2) % 16 == 0 && (batch_in_stride * 2) % 16 == 0,
                 "GEMM batch strides must be multiples of 16 bytes");
  cuuint64_t dims[4] = {(cuuint64_t)inner_extent, (cuuint64_t)outer_extent, (cuuint64_t)batch_in,
                        (cuuint64_t)(batch_out > 0 ? batch_out : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), (cuuint64_t)(batch_in_stride * 2), (cuuint64_t)(batch_stride * 2)};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CRIS_CHECK_ARG(r == CUDA_SUCCESS,
                 "cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld ld=%lld batch=%lld box=%dx%d", (int)r,
                 inner_extent, outer_extent, ld, batch, box_inner, box_outer);
  return 0;
}

template <int BN, int BK, bool A_MN, bool B_MN>
static int launch_tc(const cris_gemm_args* a, const GemmKArgs& k, cudaStream_t stream) {
  using Cfg = TileCfg<BN, BK>;
  CUtensorMap tmA, tmB;
  const long long bin = a->batch_inner > 1 ? a->batch_inner : 1;
  const long long a_rows = a->a_rows > 0 ? a->a_rows : (A_MN ? a->K : a->M);
  const long long b_rows = a->b_rows > 0 ? a->b_rows : (B_MN ? a->K : a->N);
  int rc;
  if (A_MN)
    rc = make_tmap(&tmA, a->A, a->M, a_rows, a->lda, a->batch, a->strideA, bin, a->strideA2, 64, BK,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = make_tmap(&tmA, a->A, a->K, a_rows, a->lda, a->batch, a->strideA, bin, a->strideA2, BK, BM,
                   BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  // B inner extent: K-major = total K over all taps; MN-major = total N over all taps
  long long b_inner;
  if (B_MN)
    b_inner = (a->tap_mode == CRIS_TAP_ACCUM && a->b_tap_n > 0) ? (long long)(a->taps - 1) * a->b_tap_n + a->N : a->N;
  else
    b_inner = (a->tap_mode == CRIS_TAP_ACCUM && a->b_tap_k > 0) ? (long long)(a->taps - 1) * a->b_tap_k + a->K : a->K;
  if (B_MN)
    rc = make_tmap(&tmB, a->B, b_inner, b_rows, a->ldb, a->batch, a->strideB, bin, a->strideB2, 64, BK,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = make_tmap(&tmB, a->B, b_inner, b_rows, a->ldb, a->batch, a->strideB, bin, a->strideB2, BK, BN,
                   BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  GemmKArgs kk = k;
  kk.nkb = (a->K + BK - 1) / BK;
  auto kern = gemm_tc_kernel<BN, BK, A_MN, B_MN>;
  static bool attr_done = false;  // per template instantiation
  if (!attr_done) {
    CRIS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid((a->N + BN - 1) / BN, (a->M + BM - 1) / BM, a->batch * kk.taps_z * kk.splits);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, kk);
  CRIS_LAUNCH_OK();
  return 0;
}

int gemm_ref_launch(const cris_gemm_args* a, cudaStream_t stream);  // gemm_ref.cu

static std::atomic<int> g_gemm_impl{0};

int gemm_dispatch(const cris_gemm_args* a, cudaStream_t stream) {
  CRIS_CHECK_ARG(a != nullptr, "null gemm args");
  CRIS_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->batch >= 1, "bad GEMM shape M=%d N=%d K=%d batch=%d", a->M,
                 a->N, a->K, a->batch);
  CRIS_CHECK_ARG(a->tap_mode == CRIS_TAP_NONE || (a->taps >= 1 && a->taps <= 9), "bad tap count %d", a->taps);
  CRIS_CHECK_ARG(a->splits >= 1, "splits must be >= 1");
  CRIS_CHECK_ARG(a->splits == 1 || (a->d_fp32 && a->accumulate), "split-K needs fp32 atomic accumulation");
  CRIS_CHECK_ARG(!a->accumulate || a->d_fp32, "accumulate needs fp32 D");
  CRIS_CHECK_ARG(a->d_col_stride <= 1 || a->accumulate, "d_col_stride needs accumulate mode");
  CRIS_CHECK_ARG(a->tap_mode != CRIS_TAP_WGRAD || (a->a_mn && a->b_mn), "wgrad tap mode needs MN-major A and B");
  CRIS_CHECK_ARG(a->tap_mode == CRIS_TAP_NONE || a->batch == 1, "tap modes are unbatched");
  CRIS_CHECK_ARG(a->batch_inner <= 1 || a->batch % a->batch_inner == 0, "batch must be a multiple of batch_inner");
  if (g_gemm_impl.load() == 1) return gemm_ref_launch(a, stream);

  GemmKArgs k;
  k.M = a->M; k.N = a->N; k.K = a->K; k.nkb = 0;
  k.tap_mode = a->tap_mode;
  k.taps = a->tap_mode == CRIS_TAP_NONE ? 1 : a->taps;
  k.splits = a->splits;
  k.taps_z = a->tap_mode == CRIS_TAP_WGRAD ? a->taps : 1;
  for (int i = 0; i < 9; ++i) k.tap_off[i] = a->tap_off[i];
  k.b_tap_k = a->b_tap_k; k.b_tap_n = a->b_tap_n; k.d_tap_n = a->d_tap_n;
  k.D = a->D; k.ldd = a->ldd; k.strideD = a->strideD; k.strideD2 = a->strideD2;
  k.batch_inner = a->batch_inner > 1 ? a->batch_inner : 1;
  k.strideR2 = a->strideR2;
  k.d_fp32 = a->d_fp32; k.accumulate = a->accumulate;
  k.alpha = a->alpha; k.bias = a->bias; k.act = a->act;
  k.resid = a->resid; k.ldr = a->ldr; k.strideR = a->strideR; k.resid_fp32 = a->resid_fp32;
  k.mask_hp = a->mask_hp; k.mask_wp = a->mask_wp; k.colstats = a->colstats;
  k.d_col_stride = a->d_col_stride > 0 ? a->d_col_stride : 1;

  const bool amn = a->a_mn != 0, bmn = a->b_mn != 0;
  const bool k32 = (a->K <= 32) && !amn && !bmn;  // stem convs: 32 input channels per tap
  if (k32) {
    if (a->N <= 32) return launch_tc<32, 32, false, false>(a, k, stream);
    if (a->N <= 64) return launch_tc<64, 32, false, false>(a, k, stream);
    return launch_tc<128, 32, false, false>(a, k, stream);
  }
  if (!amn && !bmn) {
    if (a->N <= 32) return launch_tc<32, 64, false, false>(a, k, stream);
    if (a->N <= 64) return launch_tc<64, 64, false, false>(a, k, stream);
    return launch_tc<128, 64, false, false>(a, k, stream);
  }
  if (!amn && bmn) {
    if (a->N <= 64) return launch_tc<64, 64, false, true>(a, k, stream);
    return launch_tc<128, 64, false, true>(a, k, stream);
  }
  if (amn && bmn) {
    if (a->N <= 64) return launch_tc<64, 64, true, true>(a, k, stream);
    return launch_tc<128, 64, true, true>(a, k, stream);
  }
  set_error("GEMM operand majorness a_mn=1,b_mn=0 is not instantiated");
  return -1;
}

}  // namespace cris

extern "C" {
int cris_gemm(const cris_gemm_args* args, void* stream) {
  return cris::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}
void cris_set_gemm_impl(int impl) { cris::g_gemm_impl.store(impl); }
int cris_get_gemm_impl(void) { return cris::g_gemm_impl.load(); }
}
