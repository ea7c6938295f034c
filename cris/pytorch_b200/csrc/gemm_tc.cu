// gemm_tc.cu — the tcgen05 GEMM / implicit-GEMM-convolution core of libcris_b200.
//
//   D[b][m][n] (+)= alpha * sum_k A[b][m][k] * B[b][n][k]        bf16 x bf16 -> fp32 (TMEM)
//
// Persistent: one CTA per SM walks 128 x BN output tiles (BN = 32..256).  Warp roles (320 threads):
//   warps 0-7 : epilogue  (tcgen05.ld TMEM -> registers -> fused epilogue -> global), 2 per SM sub-partition
//   warp  8   : TMA producer (cp.async.bulk.tensor into a STAGES-deep ~196 KB smem ring, mbarrier tx)
//   warp  9   : TMEM allocator + single-thread tcgen05.mma issuer (tcgen05.commit frees slots)
// The fp32 accumulator is double buffered in TMEM (2 x BN columns): the epilogue of tile j runs while
// the MMAs of tile j+1 are issued, and the smem ring never drains between tiles.
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
  int tma_store;     // D is written with TMA tile stores (bf16/fp32, dense, 16 B aligned pitch)
  unsigned flags;    // F_* bits
  long long* trace;  // debug: clock64 stamps of CTA 0 / thread 0 (16 per tile), or null
};

constexpr int BM = 128;
constexpr int EPI_WARPS = 8;                       // 2 per SM sub-partition: a lone warp cannot hide ALU latency
constexpr int GEMM_THREADS = (EPI_WARPS + 2) * 32;  // + TMA producer warp + MMA issuer warp

template <int BN, int BK>
struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;   // one persistent CTA per SM
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;               // one accumulator buffer
  static constexpr int TMEM_COLS = 2 * ACC_COLS;                   // double-buffered accumulator
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CRIS_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CRIS_ACT_QUICKGELU) return v / (1.f + __expf(-1.702f * v));
  return v;
}


// ---- epilogue of one 32-column chunk (one output row per lane) -------------------------------------------
// The hot loop must stay compact and contiguous: with every option inlined behind runtime flags one chunk
// iteration jumped across ~58 KB of SASS and was instruction-fetch bound (ncu: stall_no_inst /
// branch_resolving, ~860 cycles per 100 executed instructions).  EPI specialises the common cases at compile
// time; everything rare (partial / unaligned chunks, QuickGELU, fp32 residuals) takes the cold generic branch.
enum { EPI_PLAIN = 0, EPI_STATS = 1, EPI_RESID = 2, EPI_ACCUM = 3 };
// runtime option bits, evaluated on the HOST and kept in one ordinary register: testing kernel-parameter fields
// inside the chunk loop costs a constant-bank load -> uniform predicate -> branch chain (~50-100 cycles each with
// only two warps per scheduler to hide it; ncu: stall_short_sb on UISETP after LDCU)
enum { F_BIAS = 1, F_RELU = 2, F_FAST = 4, F_TMA = 8, F_DFP32 = 16, F_STATS = 32, F_WGRAD = 64, F_ALPHA = 128,
       F_RELU_POST = 256 /* ReLU after the residual add: relu(acc + bias + resid), folded eval-mode BatchNorm */,
       F_WG3 = 512 /* 3x3 wgrad with <= 64 input channels: one unit = the three taps of one kernel ROW (dx = -1, 0, +1).
                      They read the same x rows shifted by one, so ONE (BK + 2)-row box serves all three (descriptor start
                      shifted by one 128-byte row per tap) and the dz tile is loaded once instead of three times: the L2
                      traffic of these L2-bound layers drops from 18 to ~6 operand passes.  Three 64-column accumulators. */ };

struct ChunkCtx {
  long long drow, rrow;  // element offsets of this lane's row in D / resid (batch offsets included)
  int ncol0;             // logical first column of the chunk (bias index / N bound)
  int dcol;              // first D column of the chunk (tap offset included)
  int ztap;
  bool row_in, row_valid;
};

// per-column (sum, sumsq) over the 32 lanes: transposing butterfly, lane L ends with column L
__device__ __forceinline__ void chunk_colstats(const GemmKArgs& p, const float (&v)[32], int ncol0, int lane,
                                               float* st0, float* st1) {
  float a[32], q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = p.d_fp32 ? v[j] : bf2f(f2bf(v[j]));
    if (!(ncol0 + j < p.N)) x = 0.f;
    a[j] = x;
    q[j] = x * x;
  }
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
  st0[lane] = a[0];
  st1[lane] = q[0];
}

// generic path: any option, any alignment (cold)
__device__ __forceinline__ void chunk_generic(const GemmKArgs& p, float (&v)[32], const ChunkCtx& cx, int lane,
                                              float* st0, float* st1) {
  const int ncol0 = cx.ncol0;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (ncol0 + j < p.N) v[j] += __ldg(p.bias + ncol0 + j);
  }
  if (p.act != CRIS_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
  }
  if (p.resid != nullptr && cx.row_in) {
    if (p.resid_fp32) {
      const float* rp = reinterpret_cast<const float*>(p.resid) + cx.rrow + cx.dcol;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j < p.N) v[j] += rp[j];
    } else {
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.resid) + cx.rrow + cx.dcol;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j < p.N) v[j] += bf2f(rp[j]);
    }
  }
  if (p.act == CRIS_ACT_RELU_POST) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (!cx.row_valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (cx.row_in) {
    if (p.d_fp32) {
      float* base = reinterpret_cast<float*>(p.D) + cx.drow;
      if (p.accumulate) {
        float* ap = base + ((p.tap_mode == CRIS_TAP_WGRAD) ? cx.ztap * p.d_tap_n : 0);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ncol0 + j < p.N) atomicAdd(ap + (long long)(ncol0 + j) * p.d_col_stride, v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ncol0 + j < p.N) base[cx.dcol + j] = v[j];
      }
    } else {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + cx.drow + cx.dcol;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j < p.N) dp[j] = f2bf(v[j]);
    }
  }
  if (p.colstats != nullptr) chunk_colstats(p, v, ncol0, lane, st0, st1);
}

// fast path: full chunk, 16-byte aligned rows, act in {none, relu}, bf16 residual
template <int EPI>
__device__ __forceinline__ void chunk_fast(const GemmKArgs& p, float (&v)[32], const ChunkCtx& cx, int lane,
                                           float* st0, float* st1, const CUtensorMap* tmD, uint8_t* stg, int trow,
                                           int b_in, int b_out, unsigned flags, const uint4 (&rq)[4], bool rq_ok) {
  if (flags & F_BIAS) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + cx.ncol0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 q = __ldg(bp + j);
      v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
    }
  }
  if (flags & F_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if constexpr (EPI == EPI_RESID) {
    // the residual chunk was fetched one chunk ahead (see the epilogue loop): its DRAM latency overlaps the
    // previous chunk's TMEM read / convert / store instead of stalling this one
    if (rq_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q = rq[j];
        const float2 e0 = unpack_bf16x2(q.x), e1 = unpack_bf16x2(q.y), e2 = unpack_bf16x2(q.z), e3 = unpack_bf16x2(q.w);
        v[8 * j] += e0.x; v[8 * j + 1] += e0.y; v[8 * j + 2] += e1.x; v[8 * j + 3] += e1.y;
        v[8 * j + 4] += e2.x; v[8 * j + 5] += e2.y; v[8 * j + 6] += e3.x; v[8 * j + 7] += e3.y;
      }
    }
  }
  if (flags & F_RELU_POST) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (!cx.row_valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if constexpr (EPI == EPI_ACCUM) {
    if (flags & F_TMA) {
      // split-K / wgrad accumulation as TMA reductions: registers (one row per lane) -> 64-byte-swizzled smem box
      // [32 rows x 16 fp32] -> cp.reduce.async.bulk.tensor (.add): the L2 adds whole lines; rows >= M are clipped
      // by the tensor map.  (Was: 32 scalar atomicAdd per lane per chunk, one L2 operation per element.)
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lane == 0) ptx::bulk_wait_read0();  // the previous box has been read out of this buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ sw) << 4)) =
              make_float4(v[16 * h + 4 * j], v[16 * h + 4 * j + 1], v[16 * h + 4 * j + 2], v[16 * h + 4 * j + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_reduce_add_4d(tmD, stg, cx.dcol + 16 * h, trow, b_in, b_out);
          ptx::bulk_commit();
        }
      }
    } else if (cx.row_in) {
      float* ap = reinterpret_cast<float*>(p.D) + cx.drow + ((flags & F_WGRAD) ? cx.ztap * p.d_tap_n : 0) +
                  (long long)cx.ncol0 * p.d_col_stride;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(ap + (long long)j * p.d_col_stride, v[j]);
    }
  } else if (flags & F_TMA) {
    // registers (one row per lane) -> 64-byte-swizzled smem box [32 rows x 64 B] -> one TMA tile store:
    // coalesced, asynchronous, clipped at M / N by the tensor map; the LSU never sees 32 scattered rows
    const int sw = (lane >> 1) & 3;
    if (flags & F_DFP32) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lane == 0) ptx::bulk_wait_read0();  // the previous box has been read out of this buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ sw) << 4)) =
              make_float4(v[16 * h + 4 * j], v[16 * h + 4 * j + 1], v[16 * h + 4 * j + 2], v[16 * h + 4 * j + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_4d(tmD, stg, cx.dcol + 16 * h, trow, b_in, b_out);
          ptx::bulk_commit();
        }
      }
    } else {
      if (lane == 0) ptx::bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
        q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ sw) << 4)) = q;
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_4d(tmD, stg, cx.dcol, trow, b_in, b_out);
        ptx::bulk_commit();
      }
    }
  } else {
    if (cx.row_in) {
      if (flags & F_DFP32) {
        float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.D) + cx.drow + cx.dcol);
#pragma unroll
        for (int j = 0; j < 8; ++j) dp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.D) + cx.drow + cx.dcol);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
          q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
          q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          dp[j] = q;
        }
      }
    }
  }
  if constexpr (EPI == EPI_STATS) {
    if ((flags & F_TMA) && !(flags & F_DFP32)) {
      // column statistics straight from the bf16 staging box (the rounded values that are stored): lane L sums
      // column L over the 32 rows with conflict-free 2-byte shared loads instead of a 62-shuffle butterfly
      const int jc = lane >> 3, e = (lane & 7) * 2;
      float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;
#pragma unroll
      for (int r = 0; r < 32; r += 2) {
        const float x0 = bf2f(*reinterpret_cast<const __nv_bfloat16*>(stg + r * 64 + ((jc ^ ((r >> 1) & 3)) << 4) + e));
        const float x1 = bf2f(*reinterpret_cast<const __nv_bfloat16*>(stg + (r + 1) * 64 + ((jc ^ ((r >> 1) & 3)) << 4) + e));
        sa += x0; qa = fmaf(x0, x0, qa);
        sb += x1; qb = fmaf(x1, x1, qb);
      }
      st0[lane] = sa + sb;
      st1[lane] = qa + qb;
    } else {
      chunk_colstats(p, v, cx.ncol0, lane, st0, st1);
    }
  }
}

struct TileCoord {
  int n0, m0, split, ztap, b_in, b_out, kb_begin, kb_cnt, total_iters;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKArgs& p, int tile, int tiles_n, int tiles_m, int BNv) {
  TileCoord t;
  const int nt = tile % tiles_n;
  int rest = tile / tiles_n;
  const int mt = rest % tiles_m;
  int z = rest / tiles_m;
  t.n0 = nt * BNv;
  t.m0 = mt * BM;
  // tap fastest: the 9 taps of one K-range run on neighbouring CTAs at the same time, so the wgrad operands
  // (the same rows, shifted) are fetched from HBM once and re-read from L2
  t.ztap = z % p.taps_z;
  z /= p.taps_z;
  t.split = z % p.splits;
  const int batch = z / p.splits;
  t.b_in = batch % p.batch_inner;
  t.b_out = batch / p.batch_inner;
  const int kb_per_split = (p.nkb + p.splits - 1) / p.splits;
  t.kb_begin = t.split * kb_per_split;
  const int kb_end = min(p.nkb, t.kb_begin + kb_per_split);
  t.kb_cnt = max(0, kb_end - t.kb_begin);
  t.total_iters = ((p.tap_mode == CRIS_TAP_ACCUM) ? p.taps : 1) * t.kb_cnt;
  return t;
}

// F_WG3 stage layout (instance <256, 64, MN, MN, ACCUM> only): the generic 48 KB stage would hold 4 stages of which 24 KB
// are used — too little data in flight per SM (measured: 0.49 ms where the traffic allows ~0.2).  WG3 re-partitions the
// same shared memory: [A: 64 k-rows x 128 B per 64-row M block][B: 66 rows x 128 B, padded to 9 KB].  M <= 64 (all
// small-channel layers): one A block per stage, 10 stages, and the never-loaded upper M block of the M = 128 MMA is ONE
// zeroed 8 KB block behind the stages that every stage's descriptor reaches through its LBO.  M > 64: 2 blocks, 7 stages.
struct Wg3Layout {
  int stages, stage_bytes, a_blocks;
  uint32_t zero_off;  // byte offset of the shared zero block (a_blocks == 1)
};
__device__ __forceinline__ Wg3Layout wg3_layout(int M) {
  Wg3Layout l;
  if (M <= 64) { l.stages = 10; l.stage_bytes = 8192 + 9216; l.a_blocks = 1; l.zero_off = 10 * (8192 + 9216); }
  else { l.stages = 7; l.stage_bytes = 16384 + 9216; l.a_blocks = 2; l.zero_off = 0; }
  return l;
}
constexpr int WG3_MAX_STAGES = 10;

// Persistent kernel: grid = min(#tiles, #SMs); CTA c walks tiles c, c+grid, ... (n fastest, so CTAs running
// together share A rows in L2).  The smem ring keeps rolling across tiles; the TMEM accumulator is double
// buffered so the epilogue of tile j overlaps the MMAs of tile j+1.
#define CRIS_TRACE(j, slot)                                                                            \
  do {                                                                                                 \
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (j) < 32) p.trace[(j) * 16 + (slot)] = clock64(); \
  } while (0)

template <int BN, int BK, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmD, const __grid_constant__ GemmKArgs p, int tiles_n,
                   int tiles_m, int total_tiles) {
  using Cfg = TileCfg<BN, BK>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  __shared__ float s_stats[2][4][2][BN];  // double-buffered per tile: one named barrier per tile suffices
  __shared__ __align__(1024) uint8_t s_stage[EPI_WARPS][2048];  // per-warp 32 x 64 B box for TMA tile stores

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  constexpr bool kWg3Inst = (BN == 256 && A_MN && B_MN && EPI == EPI_ACCUM);
  if constexpr (kWg3Inst) {
    static_assert(STAGES <= WG3_MAX_STAGES && (2 * WG3_MAX_STAGES + 4) * 8 + 4 <= 512, "barrier area");
    if (p.flags & F_WG3) {  // up to 10 stages: the barrier arrays are re-laid out inside the same 512-byte area
      empty_bar = full_bar + WG3_MAX_STAGES;
      tmem_full = empty_bar + WG3_MAX_STAGES;
      tmem_empty = tmem_full + 2;
      tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    }
  }
  if (threadIdx.x == 0) {
    const int nbar = (kWg3Inst && (p.flags & F_WG3)) ? WG3_MAX_STAGES : STAGES;
    for (int s = 0; s < nbar; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (p.tma_store) ptx::prefetch_tmap(&tmD);
  }
  if (warp == EPI_WARPS + 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if constexpr (kWg3Inst) {
    if ((p.flags & F_WG3) && p.M <= 64) {  // the shared zero block: rows 64..127 of every stage's M = 128 A tile
      uint4* z = reinterpret_cast<uint4*>(smem + wg3_layout(p.M).zero_off);
      for (int i = threadIdx.x; i < BK * 128 / 16; i += GEMM_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
      ptx::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EPI_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile, tiles_n, tiles_m, BN);
        for (int i = 0; i < tc.total_iters; ++i, ++it) {
          if constexpr (kWg3Inst) {
            if (p.flags & F_WG3) {
              const Wg3Layout wl = wg3_layout(p.M);
              const int s3 = it % wl.stages;
              const uint32_t ph3 = (uint32_t)(it / wl.stages) & 1u;
              const int k3 = (tc.kb_begin + i) * BK;
              ptx::mbar_wait(&empty_bar[s3], ph3 ^ 1u, 100 + s3);
              uint8_t* sa3 = smem + s3 * wl.stage_bytes;
              ptx::mbar_arrive_expect_tx(&full_bar[s3], wl.a_blocks * (BK * 128) + (BK + 2) * 128);
              for (int q = 0; q < wl.a_blocks; ++q)
                ptx::tma_load_4d(sa3 + q * (BK * 128), &tmA, &full_bar[s3], tc.m0 + 64 * q, k3, tc.b_in, tc.b_out);
              // rows k + off(dy, dx = -1) .. + BK + 1 of x: the window of the three taps of kernel row `ztap`
              ptx::tma_load_4d(sa3 + wl.a_blocks * (BK * 128), &tmB, &full_bar[s3], 0, k3 + p.tap_off[3 * tc.ztap],
                               tc.b_in, tc.b_out);
              continue;
            }
          }
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          const int t = (p.tap_mode == CRIS_TAP_ACCUM) ? (i / tc.kb_cnt) : tc.ztap;
          const int k = (tc.kb_begin + (i % tc.kb_cnt)) * BK;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u, 100 + s);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          const int a_row_off = (p.tap_mode == CRIS_TAP_ACCUM) ? p.tap_off[t] : 0;
          int b_k_off = 0, b_n_off = 0;
          if (p.tap_mode == CRIS_TAP_ACCUM) {
            b_k_off = t * p.b_tap_k;
            b_n_off = t * p.b_tap_n;
          } else if (p.tap_mode == CRIS_TAP_WGRAD) {
            b_k_off = p.tap_off[t];
          }
          if constexpr (!A_MN) {
            ptx::tma_load_4d(sa, &tmA, &full_bar[s], k, tc.m0 + a_row_off, tc.b_in, tc.b_out);
          } else {
#pragma unroll
            for (int q = 0; q < BM / 64; ++q)
              ptx::tma_load_4d(sa + q * (BK * 128), &tmA, &full_bar[s], tc.m0 + 64 * q, k, tc.b_in, tc.b_out);
          }
          if constexpr (!B_MN) {
            ptx::tma_load_4d(sb, &tmB, &full_bar[s], k + b_k_off, tc.n0 + b_n_off, tc.b_in, tc.b_out);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_4d(sb + j * (BK * 128), &tmB, &full_bar[s], tc.n0 + b_n_off + 64 * j, k + b_k_off,
                               tc.b_in, tc.b_out);
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t k_layout = (BK == 64) ? ptx::kLayoutSW128 : ptx::kLayoutSW64;
      constexpr uint32_t k_sbo = (BK == 64) ? 1024u : 512u;
      int it = 0, acc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile, tiles_n, tiles_m, BN);
        if (tc.total_iters == 0) continue;
        const int buf = acc & 1;
        ptx::mbar_wait(&tmem_empty[buf], (((uint32_t)acc >> 1) & 1u) ^ 1u, 400 + buf);  // epilogue drained it
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS);
        for (int i = 0; i < tc.total_iters; ++i, ++it) {
          if constexpr (kWg3Inst) {
            if (p.flags & F_WG3) {
              const Wg3Layout wl = wg3_layout(p.M);
              const int s3 = it % wl.stages;
              const uint32_t ph3 = (uint32_t)(it / wl.stages) & 1u;
              ptx::mbar_wait(&full_bar[s3], ph3, 200 + s3);
              ptx::tc_fence_after();
              const uint32_t sa3 = ptx::smem_u32(smem + s3 * wl.stage_bytes);
              const uint32_t sb3 = sa3 + wl.a_blocks * (BK * 128);
              // second M block: the next 8 KB of the stage, or the shared zero block (its LBO differs per stage)
              const uint32_t lbo = wl.a_blocks == 2 ? (uint32_t)(BK * 128) : ptx::smem_u32(smem + wl.zero_off) - sa3;
              const uint64_t adesc3 = ptx::make_smem_desc(sa3, lbo, 1024, ptx::kLayoutSW128);
              // ONE N = 192 MMA per 16 k-rows covers the three taps: an MN-major B operand is made of 64-column blocks
              // LBO bytes apart, and LBO = 128 B = one x row, so block j is the window shifted by j rows (dx = j - 1).
              // dz (A) is read from shared memory once instead of three times; accumulator columns 64 j .. 64 j + 63 = tap j.
              const uint64_t bdesc3 = ptx::make_smem_desc(sb3, 128, 1024, ptx::kLayoutSW128);
              constexpr uint32_t idesc3 = ptx::make_idesc_bf16(BM, 192, 1, 1);
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                ptx::umma_bf16(tacc, adesc3 + (uint64_t)((kk * 2048) >> 4), bdesc3 + (uint64_t)((kk * 2048) >> 4), idesc3,
                               (i > 0 || kk > 0) ? 1u : 0u);
              ptx::umma_commit(&empty_bar[s3]);
              continue;
            }
          }
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
            ptx::umma_bf16(tacc, adesc + a_adv, bdesc + b_adv, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        ptx::umma_commit(&tmem_full[buf]);  // accumulator of this tile complete
        ++acc;
      }
    }
  } else {
    // ===================== epilogue (warps 0-7) =====================
    // warp w reads TMEM lanes 32*(w%4) .. +31 (rows of the tile) and handles the 32-column chunks c with
    // c % 2 == w / 4
    const int wq = warp & 3, chalf = warp >> 2;
    constexpr int NCHUNK = BN / 32;
    const int my_last = (chalf < NCHUNK) ? (((NCHUNK - 1 - chalf) / 2) * 2 + chalf) : -1;
    int acc = 0;
    unsigned flags = p.flags;
    float alpha = p.alpha;
    int ncols = p.N;
    asm volatile("" : "+r"(flags), "+f"(alpha), "+r"(ncols));  // pin in ordinary registers (see F_* above)
    int tcount = 0;
    // residual operand of the NEXT chunk this warp will process (EPI_RESID fast path), fetched one chunk ahead
    uint4 rnext[4] = {};
    bool rnext_ok = false;
    auto fetch_resid = [&](const TileCoord& q, int c2) {
      rnext_ok = false;
      if constexpr (EPI == EPI_RESID) {
        const long long row2 = (long long)q.m0 + wq * 32 + lane;
        if ((flags & F_FAST) && row2 < p.M && q.n0 + c2 * 32 + 32 <= ncols) {
          const uint4* rp = reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(p.resid) + (long long)q.b_out * p.strideR +
              (long long)q.b_in * p.strideR2 + row2 * p.ldr + q.n0 + c2 * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rnext[j] = rp[j];
          rnext_ok = true;
        }
      }
    };
    if constexpr (EPI == EPI_RESID) {
      if ((int)blockIdx.x < total_tiles && chalf < NCHUNK)
        fetch_resid(decode_tile(p, blockIdx.x, tiles_n, tiles_m, BN), chalf);
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const TileCoord tc = decode_tile(p, tile, tiles_n, tiles_m, BN);
      const int n0 = tc.n0, ztap = tc.ztap;
      const long long row = (long long)tc.m0 + wq * 32 + lane;
      const bool row_in = row < p.M;
      const bool row_valid = row_in && interior_row(row, p.mask_hp, p.mask_wp);
      const int dcol0 = n0 + ((p.tap_mode == CRIS_TAP_WGRAD) ? ztap * p.d_tap_n : 0);
      const long long drow = (long long)tc.b_out * p.strideD + (long long)tc.b_in * p.strideD2 + row * p.ldd;
      const long long rrow = (long long)tc.b_out * p.strideR + (long long)tc.b_in * p.strideR2 + row * p.ldr;
      const bool has_acc = tc.total_iters > 0;
      const int buf = acc & 1;
      if (has_acc) {
        ptx::mbar_wait(&tmem_full[buf], ((uint32_t)acc >> 1) & 1u, 300 + buf);
        ptx::tc_fence_after();
      }
      const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS) + ((uint32_t)(wq * 32) << 16);
      if (has_acc && my_last < 0) {  // nothing to read for this warp (BN = 32): release the buffer right away
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
      }
      CRIS_TRACE(acc, 0);
#pragma unroll 1
      for (int c = chalf; c < NCHUNK; c += 2) {
        bool wg3 = false;
        if constexpr (BN == 256 && A_MN && B_MN && EPI == EPI_ACCUM) {
          wg3 = (flags & F_WG3) != 0;
          if (wg3 && c >= 6) {  // only three 64-column accumulators are live
            if (c == my_last && has_acc) {
              ptx::tc_fence_before();
              if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
            }
            continue;
          }
        }
        uint4 rcur[4] = {};
        bool rcur_ok = false;
        if constexpr (EPI == EPI_RESID) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
          rcur_ok = rnext_ok;
          if (c + 2 < NCHUNK) {
            fetch_resid(tc, c + 2);
          } else if (tile + (int)gridDim.x < total_tiles) {
            fetch_resid(decode_tile(p, tile + gridDim.x, tiles_n, tiles_m, BN), chalf);
          } else {
            rnext_ok = false;
          }
        }
        float v[32];
        if (has_acc) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tacc + (uint32_t)(c * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          if (flags & F_ALPHA) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (c == my_last && has_acc) {
          // all TMEM reads of this tile are done: hand the accumulator buffer back to the MMA warp
          ptx::tc_fence_before();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
        }
        ChunkCtx cx;
        cx.ncol0 = n0 + c * 32;
        cx.dcol = dcol0 + c * 32;
        cx.drow = drow; cx.rrow = rrow; cx.ztap = ztap; cx.row_in = row_in; cx.row_valid = row_valid;
        if (wg3) {  // accumulator c / 2 is tap 3 * ztap + c / 2; its two chunks are input channels 0-31 / 32-63
          cx.ztap = 3 * ztap + (c >> 1);
          cx.ncol0 = (c & 1) * 32;
          cx.dcol = cx.ztap * p.d_tap_n + cx.ncol0;
        }
        float* st0 = &s_stats[tcount & 1][wq][0][c * 32];
        float* st1 = &s_stats[tcount & 1][wq][1][c * 32];
        if (cx.ncol0 >= ncols) {
          if (flags & F_STATS) { st0[lane] = 0.f; st1[lane] = 0.f; }
          continue;
        }
        if ((flags & F_FAST) && (cx.ncol0 + 32 <= ncols || ((flags & F_TMA) && EPI == EPI_PLAIN && !(flags & F_BIAS))))
          chunk_fast<EPI>(p, v, cx, lane, st0, st1, &tmD, s_stage[warp], tc.m0 + wq * 32, tc.b_in, tc.b_out, flags, rcur,
                          rcur_ok);
        else chunk_generic(p, v, cx, lane, st0, st1);
      }
      CRIS_TRACE(acc, 14);
      if (p.colstats != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int j = threadIdx.x; j < BN; j += EPI_WARPS * 32) {
          if (n0 + j < p.N) {
            const float(*ss)[2][BN] = s_stats[tcount & 1];
            const float s0 = ss[0][0][j] + ss[1][0][j] + ss[2][0][j] + ss[3][0][j];
            const float s1 = ss[0][1][j] + ss[1][1][j] + ss[2][1][j] + ss[3][1][j];
            // 64 partial rows (m_tile % 64): spreads the atomics, keeps the follow-up reduction tiny
            float* dst = p.colstats + (size_t)((tc.m0 / BM) & 63) * 2 * p.N;
            atomicAdd(dst + n0 + j, s0);
            atomicAdd(dst + p.N + n0 + j, s1);
          }
        }
        // no second barrier: the next tile fills the other s_stats buffer, and this buffer is only rewritten two
        // tiles later, i.e. after every warp passed the next tile's barrier (hence finished these reads)
      }
      if (has_acc) ++acc;
    }
    if (lane == 0) ptx::bulk_wait0();  // every TMA tile store of this warp has been performed
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
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
                     long long batch_in_stride, int box_inner, int box_outer, CUtensorMapSwizzle swz, int esz = 2) {
  auto fn = get_encode_fn();
  CRIS_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CRIS_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base not 16B aligned");
  CRIS_CHECK_ARG((ld * esz) % 16 == 0, "GEMM operand pitch %lld elements is not a multiple of 16 bytes", ld);
  if (batch_in <= 1) {
    batch_in = 1;
    batch_in_stride = ld * (outer_extent > 0 ? outer_extent : 1);
  }
  const long long batch_out = batch / batch_in;
  if (batch_out <= 1) batch_stride = batch_in_stride * batch_in;
  CRIS_CHECK_ARG((batch_stride * esz) % 16 == 0 && (batch_in_stride * esz) % 16 == 0,
                 "GEMM batch strides must be multiples of 16 bytes");
  cuuint64_t dims[4] = {(cuuint64_t)inner_extent, (cuuint64_t)outer_extent, (cuuint64_t)batch_in,
                        (cuuint64_t)(batch_out > 0 ? batch_out : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * esz), (cuuint64_t)(batch_in_stride * esz), (cuuint64_t)(batch_stride * esz)};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CRIS_CHECK_ARG(r == CUDA_SUCCESS,
                 "cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld ld=%lld batch=%lld box=%dx%d", (int)r,
                 inner_extent, outer_extent, ld, batch, box_inner, box_outer);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// tile width: the widest BN whose tile count still fills the machine (wide tiles halve the A re-reads)
static int pick_bn(const cris_gemm_args* a, int max_bn) {
  const long long z = (long long)a->batch * (a->tap_mode == CRIS_TAP_WGRAD ? a->taps : 1) * a->splits;
  const long long tm = (a->M + BM - 1) / BM;
  for (int bn = max_bn; bn > 64; bn >>= 1) {
    if (a->N < bn / 2 + 1) continue;
    const long long tiles = tm * ((a->N + bn - 1) / bn) * z;
    if (tiles >= num_sms() || bn == 128) {
      if (a->N > bn / 2) return bn;
    }
  }
  return a->N <= 32 ? 32 : 64;
}

// split-K plan of an fp32-accumulating GEMM (wgrad): tile width and split count are chosen together so that the work
// units fill whole waves of the persistent grid (two waves: the epilogue atomics of one unit overlap the next
// unit's main loop).  The widest tile that keeps >= 90% of its waves busy wins (wide tiles re-read A less).
static void plan_split_k(const cris_gemm_args* a, int* bn_out, int* splits_out) {
  const long long z = (long long)a->batch * (a->tap_mode == CRIS_TAP_WGRAD ? a->taps : 1);
  const long long tm = (a->M + BM - 1) / BM;
  const int nkb = (a->K + 63) / 64;
  const int sms = num_sms();
  double best = -1.0;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (bn > 64 && a->N <= bn / 2) continue;
    const long long tiles = tm * ((a->N + bn - 1) / bn) * z;
    long long s = tiles >= 2 * sms ? 1 : (2 * sms) / tiles;
    if (s > nkb / 4) s = nkb / 4;
    if (s < 1) s = 1;
    const long long units = tiles * s, waves = (units + sms - 1) / sms;
    const double eff = (double)units / (double)(waves * sms);
    if (eff > best) {
      best = eff;
      *bn_out = bn;
      *splits_out = (int)s;
    }
    if (eff >= 0.9) break;
  }
}

template <int BN, int BK, bool A_MN, bool B_MN, int EPI>
static int launch_tc_epi(const cris_gemm_args* a, const GemmKArgs& k, cudaStream_t stream, unsigned extra_flags = 0) {
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
    rc = make_tmap(&tmB, a->B, b_inner, b_rows, a->ldb, a->batch, a->strideB, bin, a->strideB2, 64,
                   (extra_flags & F_WG3) ? BK + 2 : BK, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = make_tmap(&tmB, a->B, b_inner, b_rows, a->ldb, a->batch, a->strideB, bin, a->strideB2, BK, BN,
                   BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  GemmKArgs kk = k;
  kk.nkb = (a->K + BK - 1) / BK;
  // D tile stores through TMA whenever D is a dense 16-byte-aligned bf16/fp32 matrix (not the atomic paths)
  CUtensorMap tmD = tmA;
  const int desz = a->d_fp32 ? 4 : 2;
  kk.tma_store = 0;
  const bool d_dense = (reinterpret_cast<uintptr_t>(a->D) & 15) == 0 && (a->ldd * desz) % 16 == 0 &&
                       (a->batch <= 1 || ((a->strideD * desz) % 16 == 0 && (bin <= 1 || (a->strideD2 * desz) % 16 == 0)));
  if (EPI != EPI_ACCUM && !a->accumulate && d_dense) {
    rc = make_tmap(&tmD, a->D, a->N, a->M, a->ldd, a->batch, a->strideD, bin, a->strideD2, 64 / desz, 32,
                   CU_TENSOR_MAP_SWIZZLE_64B, desz);
    if (rc) return rc;
    kk.tma_store = 1;
  }
  if (EPI == EPI_ACCUM && a->accumulate && a->d_fp32 && d_dense && a->d_col_stride <= 1 &&
      (a->tap_mode != CRIS_TAP_WGRAD || (a->d_tap_n % 4) == 0)) {
    // fp32 accumulation through TMA reductions; wgrad taps are column blocks d_tap_n apart in the same rows
    const long long inner = a->tap_mode == CRIS_TAP_WGRAD ? (long long)(a->taps - 1) * a->d_tap_n + a->N : a->N;
    rc = make_tmap(&tmD, a->D, inner, a->M, a->ldd, a->batch, a->strideD, bin, a->strideD2, 16, 32,
                   CU_TENSOR_MAP_SWIZZLE_64B, 4);
    if (rc) return rc;
    kk.tma_store = 1;
  }
  auto kern = gemm_tc_kernel<BN, BK, A_MN, B_MN, EPI>;
  CRIS_SET_SMEM_ONCE(kern, Cfg::SMEM_BYTES);  // per template instantiation and device, thread-safe
  const int tiles_n = (a->N + BN - 1) / BN, tiles_m = (a->M + BM - 1) / BM;
  const long long total = (long long)tiles_n * tiles_m * a->batch * kk.taps_z * kk.splits;
  CRIS_CHECK_ARG(total < (1ll << 31), "GEMM has too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  {
    const int desz2 = a->d_fp32 ? 4 : 2;
    const bool d_al = (reinterpret_cast<uintptr_t>(a->D) & 15) == 0 && (a->ldd * desz2) % 16 == 0 &&
                      (a->strideD * desz2) % 16 == 0 && (a->strideD2 * desz2) % 16 == 0;
    const bool r_al = a->resid == nullptr || (!a->resid_fp32 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0 &&
                                               (a->ldr * 2) % 16 == 0 && (a->strideR * 2) % 16 == 0 &&
                                               (a->strideR2 * 2) % 16 == 0);
    const bool b_al = a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
    const bool epi_match = ((EPI == EPI_ACCUM) ? (a->d_fp32 && a->accumulate) : !a->accumulate) &&
                           ((EPI == EPI_STATS) == (a->colstats != nullptr)) && ((EPI == EPI_RESID) == (a->resid != nullptr));
    const bool fast = d_al && r_al && b_al && epi_match && (a->act == CRIS_ACT_NONE || a->act == CRIS_ACT_RELU || a->act == CRIS_ACT_RELU_POST) &&
                      (EPI == EPI_ACCUM || a->d_col_stride <= 1);
    kk.flags = (a->bias ? F_BIAS : 0) | (a->act == CRIS_ACT_RELU ? F_RELU : 0) |
               (a->act == CRIS_ACT_RELU_POST ? F_RELU_POST : 0) | (fast ? F_FAST : 0) |
               (kk.tma_store ? F_TMA : 0) | (a->d_fp32 ? F_DFP32 : 0) | (a->colstats ? F_STATS : 0) |
               (a->tap_mode == CRIS_TAP_WGRAD ? F_WGRAD : 0) | (a->alpha != 1.0f ? F_ALPHA : 0) | extra_flags;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmD, kk, tiles_n, tiles_m, (int)total);
  CRIS_LAUNCH_OK();
  return 0;
}

// epilogue specialisation: wgrad-style fp32 atomics / BN statistics / bf16 residual / plain
template <int BN, int BK, bool A_MN, bool B_MN>
static int launch_tc(const cris_gemm_args* a, const GemmKArgs& k, cudaStream_t stream) {
  if (a->accumulate) return launch_tc_epi<BN, BK, A_MN, B_MN, EPI_ACCUM>(a, k, stream);
  if constexpr (!A_MN) {
    if (a->colstats != nullptr && a->resid == nullptr) return launch_tc_epi<BN, BK, A_MN, B_MN, EPI_STATS>(a, k, stream);
  }
  if constexpr (BK == 64) {
    if (a->resid != nullptr && a->colstats == nullptr) return launch_tc_epi<BN, BK, A_MN, B_MN, EPI_RESID>(a, k, stream);
  }
  return launch_tc_epi<BN, BK, A_MN, B_MN, EPI_PLAIN>(a, k, stream);
}

static long long* g_trace = nullptr;  // debug timeline buffer (cris_debug_set_trace)

static bool wg3_enabled() {  // CRIS_B200_WGRAD3=0 selects the tap-per-unit wgrad for small-channel 3x3 layers
  static const bool on = [] {
    const char* e = getenv("CRIS_B200_WGRAD3");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

int gemm_dispatch(const cris_gemm_args* a_in, cudaStream_t stream) {
  CRIS_CHECK_ARG(a_in != nullptr, "null gemm args");
  cris_gemm_args planned = *a_in;
  int planned_bn = 0;
  if (planned.splits == 0 && planned.accumulate && planned.d_fp32 && planned.M > 0 && planned.N > 0 && planned.K > 0 &&
      planned.batch >= 1)
    plan_split_k(&planned, &planned_bn, &planned.splits);
  const cris_gemm_args* a = &planned;
  CRIS_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->batch >= 1, "bad GEMM shape M=%d N=%d K=%d batch=%d", a->M,
                 a->N, a->K, a->batch);
  CRIS_CHECK_ARG(a->tap_mode == CRIS_TAP_NONE || (a->taps >= 1 && a->taps <= 9), "bad tap count %d", a->taps);
  CRIS_CHECK_ARG(a->splits >= 1, "splits must be >= 1 (0 = automatic, fp32-accumulating GEMMs only)");
  CRIS_CHECK_ARG(a->splits == 1 || (a->d_fp32 && a->accumulate), "split-K needs fp32 atomic accumulation");
  CRIS_CHECK_ARG(!a->accumulate || a->d_fp32, "accumulate needs fp32 D");
  CRIS_CHECK_ARG(a->d_col_stride <= 1 || a->accumulate, "d_col_stride needs accumulate mode");
  CRIS_CHECK_ARG(a->tap_mode != CRIS_TAP_WGRAD || (a->a_mn && a->b_mn), "wgrad tap mode needs MN-major A and B");
  CRIS_CHECK_ARG(a->tap_mode == CRIS_TAP_NONE || a->batch == 1, "tap modes are unbatched");
  CRIS_CHECK_ARG(a->batch_inner <= 1 || a->batch % a->batch_inner == 0, "batch must be a multiple of batch_inner");

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
  k.trace = g_trace;

  const bool amn = a->a_mn != 0, bmn = a->b_mn != 0;
  if (a->tap_mode == CRIS_TAP_WGRAD && a->taps == 9 && amn && bmn && a->accumulate && a->d_fp32 && a->N <= 64 &&
      a->batch == 1 && a->d_col_stride <= 1 && wg3_enabled()) {
    bool rows_ok = true;  // tap index = 3 * (dy + 1) + (dx + 1): the taps of one kernel row are consecutive x rows
    for (int g = 0; g < 3; ++g)
      rows_ok = rows_ok && a->tap_off[3 * g + 1] == a->tap_off[3 * g] + 1 && a->tap_off[3 * g + 2] == a->tap_off[3 * g] + 2;
    if (rows_ok) {
      k.taps_z = 3;
      if (a_in->splits == 0) {  // automatic plan: two waves of (M tiles x 3 kernel rows x splits) units
        const long long tiles = (long long)((a->M + BM - 1) / BM) * 3;
        const int nkb = (a->K + 63) / 64;
        long long sp = tiles >= 2 * num_sms() ? 1 : (2 * num_sms()) / tiles;
        if (sp > nkb / 4) sp = nkb / 4;
        k.splits = sp < 1 ? 1 : (int)sp;
      }
      return launch_tc_epi<256, 64, true, true, EPI_ACCUM>(a, k, stream, F_WG3);
    }
  }
  const bool k32 = (a->K <= 32) && !amn && !bmn;  // stem convs: 32 input channels per tap
  if (k32) {
    if (a->N <= 32) return launch_tc<32, 32, false, false>(a, k, stream);
    if (a->N <= 64) return launch_tc<64, 32, false, false>(a, k, stream);
    return launch_tc<128, 32, false, false>(a, k, stream);
  }
  const int bn = planned_bn ? planned_bn : pick_bn(a, 256);
  if (!amn && !bmn) {
    if (bn == 32) return launch_tc<32, 64, false, false>(a, k, stream);
    if (bn == 64) return launch_tc<64, 64, false, false>(a, k, stream);
    if (bn == 128) return launch_tc<128, 64, false, false>(a, k, stream);
    return launch_tc<256, 64, false, false>(a, k, stream);
  }
  if (!amn && bmn) {
    if (bn <= 64) return launch_tc<64, 64, false, true>(a, k, stream);
    if (bn == 128) return launch_tc<128, 64, false, true>(a, k, stream);
    return launch_tc<256, 64, false, true>(a, k, stream);
  }
  if (amn && bmn) {
    if (bn <= 64) return launch_tc<64, 64, true, true>(a, k, stream);
    if (bn == 128) return launch_tc<128, 64, true, true>(a, k, stream);
    return launch_tc<256, 64, true, true>(a, k, stream);
  }
  set_error("GEMM operand majorness a_mn=1,b_mn=0 is not instantiated");
  return -1;
}

}  // namespace cris

extern "C" {
int cris_gemm(const cris_gemm_args* args, void* stream) {
  return cris::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}
int cris_gemm_plan(const cris_gemm_args* args, int* tile_n, int* splits) {
  CRIS_CHECK_ARG(args != nullptr && tile_n != nullptr && splits != nullptr, "cris_gemm_plan: null argument");
  CRIS_CHECK_ARG(args->M > 0 && args->N > 0 && args->K > 0 && args->batch >= 1, "cris_gemm_plan: bad shape");
  if (args->splits == 0 && args->accumulate && args->d_fp32) {
    cris::plan_split_k(args, tile_n, splits);
  } else {
    *tile_n = ((args->K <= 32) && !args->a_mn && !args->b_mn) ? (args->N <= 32 ? 32 : args->N <= 64 ? 64 : 128)
                                                               : cris::pick_bn(args, 256);
    *splits = args->splits > 0 ? args->splits : 1;
  }
  return 0;
}
void cris_debug_set_trace(void* dev_buf) { cris::g_trace = reinterpret_cast<long long*>(dev_buf); }
}
