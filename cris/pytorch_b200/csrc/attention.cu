// attention.cu — fused multi-head attention for 64-wide heads on tcgen05 (forward and backward).
//
//   O = dropout(softmax(alpha * Q K^T)) V      per (image, head);  Q [B*Lq, heads*64], K / V [B*Lk, heads*64] bf16
//
// Reference: the core of F.multi_head_attention_forward as called by the decoder's self attention
// (model/layers.py:233-237, 676 x 676 keys, attention dropout 0.1) and by AttentionPool2d (model/clip.py:119-139,
// 169 x 169).  The previous path materialised S, P, dropout(P) and dP in HBM (468 MB each per decoder layer at
// B = 64) and ran softmax as separate passes; here the scores never leave the SM:
//
// forward  (one CTA per image x head x 128-query tile, 2 CTAs / SM):
//   pass 1: S_j = Q K_j^T (tcgen05.mma, fp32 in TMEM) for every 128-key block -> exact row maxima
//   pass 2: S_j again -> p = exp2(t - max), row sums, dropout (row hash x 16-bit hash per column pair, vec.cuh drop_keep_rc), P_j (bf16) written to
//           shared memory in the canonical 128B-swizzled K-major layout -> O += P_j V_j (tcgen05.mma) ;
//           epilogue O / (rowsum * (1 - p_drop)) -> bf16, and the row's log2-sum-exp for the backward.
//   (two passes instead of an online rescale: QK^T is 1/3 of the forward flops and the tensor pipe idles anyway)
// backward (one CTA per image x head x 128-key block, loops over the query tiles):
//   S = Q_i K^T, dP = dO_i V^T (TMEM) -> P = exp2(t - lse), Pd = dropout(P), dS = alpha * P (dropout'(dP) - D)
//   -> shared memory -> dV += Pd^T dO_i, dK += dS^T Q_i (accumulated in TMEM over the query tiles),
//   dQ_i = dS K (TMEM -> fp32 vector atomics into the dQ accumulator, 6 key blocks add into one row).
//   The P / dS tiles are written once as [query][key] and read by the tensor core both K-major (dQ) and MN-major
//   (the transposed products dV, dK): no transposition pass.
// Warp roles: NSW softmax + epilogue warps (forward 8, two CTAs per SM; backward 16, one CTA per SM) — warp w owns
// TMEM lanes 32*(w%4).. (one row per thread) and the column slice w/4 of every 128-column block, so every scheduler
// has several softmax warps (the phase is instruction-issue bound: exp2 + dropout hash + bf16 pack per score; the
// hash is shared by column pairs and interior tiles skip the edge masks); then one TMA producer warp and one TMEM
// allocator + single-thread MMA issuer warp.
#include <cudaTypedefs.h>

#include <mutex>

#include "vec.cuh"
#include "ptx.cuh"

namespace cris {

constexpr int AT_THREADS = 320;       // forward: 8 softmax warps + producer + MMA
constexpr int AT_SM_THREADS = 256;
constexpr int ATB_SW = 16;            // backward: 16 softmax warps (4 column quarters)
constexpr int ATB_SM_THREADS = ATB_SW * 32;
constexpr int ATB_THREADS = ATB_SM_THREADS + 64;
constexpr int AT_TILE = 16384;  // 128 rows x 64 bf16 = one 128B-swizzled box

struct AttnArgs {
  int B, heads, Lq, Lk, LkPad;
  float scale_log2;  // alpha * log2(e): t = s * scale_log2 lives in the log2 domain
  float alpha;
  uint32_t drop_thresh;  // 16-bit threshold (drop_thresh16), 0 = no dropout
  float inv_keep;    // 1 / (1 - p_drop)
  uint64_t seed;
  const uint64_t* seed_dev;
  // forward
  __nv_bfloat16* O; long long ldo;
  float* lse;        // [B*heads*Lq] log2-domain log-sum-exp of t
  // backward
  const float* Dsum; // [B*heads*Lq] rowsum(dO * O)
  float* dQacc; long long lddq;          // fp32 [B*Lq, heads*64], zeroed by the caller
  __nv_bfloat16* dK; long long lddk;
  __nv_bfloat16* dV; long long lddv;
};

__device__ __forceinline__ float fast_exp2(float x) {  // x <= 0 here; MUFU.EX2, ~2 ulp
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int row, int k0, const float* v) {
  // 8 consecutive keys k0..k0+7 of row `row` into a [128 x 128] bf16 tile stored as two 128B-swizzled [128 x 64] halves
  const int half = k0 >> 6, chunk = (k0 & 63) >> 3;
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + half * AT_TILE + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
}

// K-major operand descriptor of a [128 x 64] swizzled box, advanced by `k16` steps of 16 k-elements
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, int k16) {
  return ptx::make_smem_desc(saddr + (uint32_t)k16 * 32u, 16, 1024, ptx::kLayoutSW128);
}
// K-major operand spanning two boxes (K = 128): step k16 in 0..7
__device__ __forceinline__ uint64_t desc_kmajor2(uint32_t saddr, int k16) {
  return ptx::make_smem_desc(saddr + (uint32_t)(k16 >> 2) * AT_TILE + (uint32_t)(k16 & 3) * 32u, 16, 1024, ptx::kLayoutSW128);
}
// MN-major operand: rows of the box(es) are the K index (128 rows), 64-element MN blocks `AT_TILE` bytes apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, int k16) {
  return ptx::make_smem_desc(saddr + (uint32_t)k16 * 2048u, AT_TILE, 1024, ptx::kLayoutSW128);
}

// ================================================================================================================
// forward
// ================================================================================================================
struct FwdSmem {
  static constexpr int Q = 0, KV = AT_TILE, P = 4 * AT_TILE, BAR = 6 * AT_TILE, XCH = 6 * AT_TILE + 256,
                       BYTES = 6 * AT_TILE + 256 + 1024 /* row-statistics exchange [2][128] */ + 1024 /* align */;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;       // 256 arrivals
  uint64_t* p_full = bars + 9;        // 256 arrivals
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* xch = reinterpret_cast<float*>(smem + FwdSmem::XCH);  // [2 column halves][128 rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Lk + 127) / 128;

  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 3; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, AT_SM_THREADS);
    ptx::mbar_init(p_full, AT_SM_THREADS);
    ptx::mbar_init(p_empty, 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); }
  if (warp == 9) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, AT_TILE);
      ptx::tma_load_4d(smem + FwdSmem::Q, &tmQ, q_full, 0, q0, h, b);
      const int total = 3 * nkv;  // pass 1: K_0..K_{n-1};  pass 2: K_0, V_0, K_1, V_1, ...
      for (int t = 0; t < total; ++t) {
        const int s = t % 3;
        ptx::mbar_wait(&kv_empty[s], (((uint32_t)(t / 3)) & 1u) ^ 1u, 800 + s);
        ptx::mbar_arrive_expect_tx(&kv_full[s], AT_TILE);
        const bool is_v = t >= nkv && ((t - nkv) & 1);
        const int j = t < nkv ? t : (t - nkv) >> 1;
        ptx::tma_load_4d(smem + FwdSmem::KV + s * AT_TILE, is_v ? &tmV : &tmK, &kv_full[s], 0, j * 128, h, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idS = ptx::make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idO = ptx::make_idesc_bf16(128, 64, 0, 1);
      const uint32_t aQ = ptx::smem_u32(smem + FwdSmem::Q), aP = ptx::smem_u32(smem + FwdSmem::P);
      ptx::mbar_wait(q_full, 0, 810);
      int t = 0, ns = 0;  // ring position, S tiles issued so far
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nkv; ++j) {
          int s = t % 3;
          ptx::mbar_wait(&kv_full[s], ((uint32_t)(t / 3)) & 1u, 820 + s);
          if (ns > 0) ptx::mbar_wait(s_empty, ((uint32_t)(ns - 1)) & 1u, 830);
          ptx::tc_fence_after();
          const uint32_t aK = ptx::smem_u32(smem + FwdSmem::KV + s * AT_TILE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tS, desc_kmajor(aQ, kk), desc_kmajor(aK, kk), idS, kk > 0 ? 1u : 0u);
          ptx::umma_commit(&kv_empty[s]);
          ptx::umma_commit(s_full);
          ++t; ++ns;
          if (pass == 1) {
            s = t % 3;
            ptx::mbar_wait(&kv_full[s], ((uint32_t)(t / 3)) & 1u, 840 + s);
            ptx::mbar_wait(p_full, ((uint32_t)j) & 1u, 850);
            ptx::tc_fence_after();
            const uint32_t aV = ptx::smem_u32(smem + FwdSmem::KV + s * AT_TILE);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              ptx::umma_bf16(tO, desc_kmajor2(aP, kk), desc_mnmajor(aV, kk), idO, (j > 0 || kk > 0) ? 1u : 0u);
            ptx::umma_commit(&kv_empty[s]);
            ptx::umma_commit(p_empty);
            ++t;
          }
        }
      }
      ptx::umma_commit(o_full);
    }
  } else {
    // ------------------------------ softmax + epilogue ------------------------------
    // thread = (query row r, column half ch): TMEM lanes 32*(warp%4).., columns ch*64.. of every 128-column block
    const int quad = warp & 3, ch = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int q = q0 + r;
    uint64_t seed = p.seed;
    if (p.seed_dev != nullptr) seed += *p.seed_dev;
    const uint32_t rh = drop_row_hash(seed, (uint64_t)(b * p.heads + h) * (uint64_t)p.Lq + (uint64_t)q);
    uint8_t* sP = smem + FwdSmem::P;
    float m = -INFINITY;
    int ns = 0;
    for (int j = 0; j < nkv; ++j, ++ns) {       // pass 1: row maximum (this thread's column half)
      ptx::mbar_wait(s_full, ((uint32_t)ns) & 1u, 860);
      ptx::tc_fence_after();
      const int valid = p.Lk - j * 128 - ch * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tS + lane_off + (uint32_t)(ch * 64 + c * 32), v);
        ptx::tmem_ld_wait();
        if (valid >= 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_empty);
    }
    xch[ch * 128 + r] = m;
    softmax_bar();
    m = fmaxf(xch[r], xch[128 + r]);
    softmax_bar();                              // both halves have read the maxima before xch is reused for the sums
    const float mt = m * p.scale_log2;          // alpha > 0: max commutes with the scaling
    float l = 0.f;
    for (int j = 0; j < nkv; ++j, ++ns) {       // pass 2: probabilities -> P tile
      ptx::mbar_wait(s_full, ((uint32_t)ns) & 1u, 870);
      if (j > 0) ptx::mbar_wait(p_empty, ((uint32_t)(j - 1)) & 1u, 880);
      ptx::tc_fence_after();
      const int valid = p.Lk - j * 128 - ch * 64;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(tS + lane_off + (uint32_t)(ch * 64 + c * 16), v);
        ptx::tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, -mt));
        if (valid < 64) {                       // ragged last key block only
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (!(c * 16 + i < valid)) e[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) l += e[i];
        if (p.drop_thresh != 0u) {
          const uint32_t pair0 = (uint32_t)(j * 128 + ch * 64 + c * 16) >> 1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t hb = drop_pair_bits(rh, pair0 + (uint32_t)i);
            if ((hb & 0xffffu) < p.drop_thresh) e[2 * i] = 0.f;
            if ((hb >> 16) < p.drop_thresh) e[2 * i + 1] = 0.f;
          }
        }
        st_tile_chunk(sP, r, ch * 64 + c * 16, &e[0]);
        st_tile_chunk(sP, r, ch * 64 + c * 16 + 8, &e[8]);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_empty);
      ptx::mbar_arrive(p_full);
    }
    xch[ch * 128 + r] = l;
    softmax_bar();
    l = xch[r] + xch[128 + r];
    // epilogue: this warp's 32 of the 64 output columns
    ptx::mbar_wait(o_full, 0, 890);
    ptx::tc_fence_after();
    const float inv = p.inv_keep / l;
    const bool row_ok = q < p.Lq;
    if (row_ok && ch == 0 && p.lse != nullptr) p.lse[(size_t)(b * p.heads + h) * p.Lq + q] = mt + log2f(l);
    __nv_bfloat16* op = p.O + ((long long)b * p.Lq + q) * p.ldo + h * 64 + ch * 32;
    {
      uint32_t v[32];
      // tcgen05.ld is warp-collective (.sync.aligned): every lane executes it, only valid rows store
      ptx::tmem_ld_32x32(tO + lane_off + (uint32_t)(ch * 32), v);
      ptx::tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[g8 * 8 + i]) * inv;
          st8(op + g8 * 8, o);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem);
  }
}

// ================================================================================================================
// backward
// ================================================================================================================
struct BwdSmem {
  static constexpr int K = 0, V = AT_TILE, QD = 2 * AT_TILE /* 2 stages x (Q, dO) */, P = 6 * AT_TILE, DS = 8 * AT_TILE,
                       BAR = 10 * AT_TILE, BYTES = 10 * AT_TILE + 256 + 1024;
};

// D[bh][q] = sum_d dO[q][d] * O[q][d]   (one warp per row, 64 channels = 2 per lane)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ O, long long ldo, const __nv_bfloat16* __restrict__ dO,
                                     long long lddo, float* __restrict__ D, int B, int heads, int Lq) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * heads * Lq;
  if (wid >= total) return;
  const int q = (int)(wid % Lq);
  const int bh = (int)(wid / Lq);
  const int b = bh / heads, h = bh % heads;
  const long long row = (long long)b * Lq + q;
  const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(O + row * ldo + h * 64 + lane * 2));
  const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dO + row * lddo + h * 64 + lane * 2));
  const float s = warp_sum(o.x * d.x + o.y * d.y);
  if (lane == 0) D[wid] = s;
}

__global__ void __launch_bounds__(ATB_THREADS, 1)
    attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnArgs p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BAR);
  uint64_t* kv_full = bars;           // K and V of this block (1 arrival, 2 tiles)
  uint64_t* qd_full = bars + 1;       // [2]
  uint64_t* qd_empty = bars + 3;      // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* s_empty = bars + 6;       // 256
  uint64_t* pds_full = bars + 7;      // 256
  uint64_t* pds_empty = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;     // 256
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nq = (p.Lq + 127) / 128;

  if (threadIdx.x == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&qd_full[s], 1); ptx::mbar_init(&qd_empty[s], 1); }
    ptx::mbar_init(sdp_full, 1);
    ptx::mbar_init(s_empty, ATB_SM_THREADS);
    ptx::mbar_init(pds_full, ATB_SM_THREADS);
    ptx::mbar_init(pds_empty, 1);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dq_empty, ATB_SM_THREADS);
    ptx::fence_barrier_init();
  }
  if (warp == ATB_SW && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmdO);
  }
  if (warp == ATB_SW + 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == ATB_SW) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * AT_TILE);
      ptx::tma_load_4d(smem + BwdSmem::K, &tmK, kv_full, 0, k0, h, b);
      ptx::tma_load_4d(smem + BwdSmem::V, &tmV, kv_full, 0, k0, h, b);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        ptx::mbar_wait(&qd_empty[s], (((uint32_t)(i >> 1)) & 1u) ^ 1u, 900 + s);
        ptx::mbar_arrive_expect_tx(&qd_full[s], 2 * AT_TILE);
        ptx::tma_load_4d(smem + BwdSmem::QD + s * 2 * AT_TILE, &tmQ, &qd_full[s], 0, i * 128, h, b);
        ptx::tma_load_4d(smem + BwdSmem::QD + s * 2 * AT_TILE + AT_TILE, &tmdO, &qd_full[s], 0, i * 128, h, b);
      }
    }
  } else if (warp == ATB_SW + 1) {
    if (lane == 0) {
      constexpr uint32_t idS = ptx::make_idesc_bf16(128, 128, 0, 0);   // Q K^T, dO V^T
      constexpr uint32_t idT = ptx::make_idesc_bf16(128, 64, 1, 1);    // Pd^T dO, dS^T Q
      constexpr uint32_t idQ = ptx::make_idesc_bf16(128, 64, 0, 1);    // dS K
      const uint32_t aK = ptx::smem_u32(smem + BwdSmem::K), aV = ptx::smem_u32(smem + BwdSmem::V);
      const uint32_t aP = ptx::smem_u32(smem + BwdSmem::P), aDS = ptx::smem_u32(smem + BwdSmem::DS);
      ptx::mbar_wait(kv_full, 0, 910);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        const uint32_t aQ = ptx::smem_u32(smem + BwdSmem::QD + s * 2 * AT_TILE), aDO = aQ + AT_TILE;
        ptx::mbar_wait(&qd_full[s], ((uint32_t)(i >> 1)) & 1u, 920 + s);
        if (i > 0) ptx::mbar_wait(s_empty, ((uint32_t)(i - 1)) & 1u, 930);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tS, desc_kmajor(aQ, kk), desc_kmajor(aK, kk), idS, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tdP, desc_kmajor(aDO, kk), desc_kmajor(aV, kk), idS, kk > 0 ? 1u : 0u);
        ptx::umma_commit(sdp_full);
        ptx::mbar_wait(pds_full, ((uint32_t)i) & 1u, 940);
        if (i > 0) ptx::mbar_wait(dq_empty, ((uint32_t)(i - 1)) & 1u, 950);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dV[key][d] += sum_q Pd[q][key] dO[q][d]
          ptx::umma_bf16(tdV, desc_mnmajor(aP, kk), desc_mnmajor(aDO, kk), idT, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dK[key][d] += sum_q dS[q][key] Q[q][d]
          ptx::umma_bf16(tdK, desc_mnmajor(aDS, kk), desc_mnmajor(aQ, kk), idT, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dQ[q][d] = sum_key dS[q][key] K[key][d]
          ptx::umma_bf16(tdQ, desc_kmajor2(aDS, kk), desc_mnmajor(aK, kk), idQ, kk > 0 ? 1u : 0u);
        ptx::umma_commit(&qd_empty[s]);
        ptx::umma_commit(pds_empty);
        ptx::umma_commit(dq_full);
      }
    }
  } else {
    // thread = (query row r, column quarter cq): 32 of the 128 key columns; for the dK/dV epilogue r is the key row
    const int quad = warp & 3, cq = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint64_t seed = p.seed;
    if (p.seed_dev != nullptr) seed += *p.seed_dev;
    uint8_t* sP = smem + BwdSmem::P;
    uint8_t* sDS = smem + BwdSmem::DS;
    const int valid = p.Lk - k0 - cq * 32;  // keys of this thread's column quarter that exist
    const size_t bh = (size_t)(b * p.heads + h);
    for (int i = 0; i < nq; ++i) {
      const int q = i * 128 + r;
      const bool row_ok = q < p.Lq;
      const bool interior = (i * 128 + 128 <= p.Lq) && (k0 + 128 <= p.Lk);  // CTA-uniform: no edge masks needed
      const float lse = row_ok ? p.lse[bh * p.Lq + q] : 0.f;
      const float Dq = row_ok ? p.Dsum[bh * p.Lq + q] : 0.f;
      const uint32_t rh = drop_row_hash(seed, bh * (uint64_t)p.Lq + (uint64_t)q);
      ptx::mbar_wait(sdp_full, ((uint32_t)i) & 1u, 960);
      if (i > 0) ptx::mbar_wait(pds_empty, ((uint32_t)(i - 1)) & 1u, 970);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[16], dv[16];
        ptx::tmem_ld_32x16(tS + lane_off + (uint32_t)(cq * 32 + c * 16), sv);
        ptx::tmem_ld_32x16(tdP + lane_off + (uint32_t)(cq * 32 + c * 16), dv);
        ptx::tmem_ld_wait();
        float pd[16], ds[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pd[j] = fast_exp2(fmaf(__uint_as_float(sv[j]), p.scale_log2, -lse));
        if (!interior) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (!(row_ok && (c * 16 + j < valid))) pd[j] = 0.f;
        }
        if (p.drop_thresh != 0u) {
          const uint32_t pair0 = (uint32_t)(k0 + cq * 32 + c * 16) >> 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t hb = drop_pair_bits(rh, pair0 + (uint32_t)j);
            const bool k0_ = (hb & 0xffffu) >= p.drop_thresh, k1_ = (hb >> 16) >= p.drop_thresh;
            // ds = alpha * P * (dropout'(dP) - D);  pd = dropout(P)
            ds[2 * j] = pd[2 * j] * ((k0_ ? __uint_as_float(dv[2 * j]) * p.inv_keep : 0.f) - Dq) * p.alpha;
            ds[2 * j + 1] = pd[2 * j + 1] * ((k1_ ? __uint_as_float(dv[2 * j + 1]) * p.inv_keep : 0.f) - Dq) * p.alpha;
            pd[2 * j] = k0_ ? pd[2 * j] * p.inv_keep : 0.f;
            pd[2 * j + 1] = k1_ ? pd[2 * j + 1] * p.inv_keep : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) ds[j] = pd[j] * (__uint_as_float(dv[j]) - Dq) * p.alpha;
        }
        st_tile_chunk(sP, r, cq * 32 + c * 16, &pd[0]);
        st_tile_chunk(sP, r, cq * 32 + c * 16 + 8, &pd[8]);
        st_tile_chunk(sDS, r, cq * 32 + c * 16, &ds[0]);
        st_tile_chunk(sDS, r, cq * 32 + c * 16 + 8, &ds[8]);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_empty);
      ptx::mbar_arrive(pds_full);
      // dQ of this tile (this warp's 16 of the 64 columns): TMEM -> fp32 vector atomics (the other key blocks of this
      // image/head add to the same rows)
      ptx::mbar_wait(dq_full, ((uint32_t)i) & 1u, 980);
      ptx::tc_fence_after();
      float* dq = p.dQacc + ((long long)b * p.Lq + q) * p.lddq + h * 64 + cq * 16;
      {
        uint32_t v[16];
        ptx::tmem_ld_32x16(tdQ + lane_off + (uint32_t)(cq * 16), v);
        ptx::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            atomicAdd(reinterpret_cast<float4*>(dq + g4 * 4),
                      make_float4(__uint_as_float(v[g4 * 4]), __uint_as_float(v[g4 * 4 + 1]),
                                  __uint_as_float(v[g4 * 4 + 2]), __uint_as_float(v[g4 * 4 + 3])));
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(dq_empty);
    }
    // dK, dV of this key block (every MMA has completed: the last dq_full commit covers them)
    ptx::tc_fence_after();
    const int key = k0 + r;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* dst = (which == 0 ? p.dV + ((long long)b * p.Lk + key) * p.lddv : p.dK + ((long long)b * p.Lk + key) * p.lddk) +
                           h * 64 + cq * 16;
      const uint32_t tsrc = which == 0 ? tdV : tdK;
      uint32_t v[16];
      ptx::tmem_ld_32x16(tsrc + lane_off + (uint32_t)(cq * 16), v);
      ptx::tmem_ld_wait();
      if (key < p.Lk) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[g8 * 8 + j]);
          st8(dst + g8 * 8, o);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == ATB_SW + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// ================================================================================================================
// host
// ================================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 at_encode_fn() {
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

// [B, L, heads, 64] view of a [B*L, ld] bf16 matrix (column slice of a packed projection): box = 128 rows x 64 columns
static int at_tmap(CUtensorMap* tm, const void* base, long long ld, int B, int L, int heads) {
  auto fn = at_encode_fn();
  CRIS_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CRIS_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, "attention operand not 16B aligned");
  cuuint64_t dims[4] = {64, (cuuint64_t)L, (cuuint64_t)heads, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), 128, (cuuint64_t)((long long)L * ld * 2)};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CRIS_CHECK_ARG(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d) L=%d ld=%lld", (int)r, L, ld);
  return 0;
}

static int fill_args(AttnArgs& a, int B, int heads, int Lq, int Lk, float alpha, float p_drop, uint64_t seed,
                     const uint64_t* seed_dev) {
  CRIS_CHECK_ARG(B >= 1 && heads >= 1 && Lq >= 1 && Lk >= 1, "attention: bad shape B=%d heads=%d Lq=%d Lk=%d", B, heads, Lq, Lk);
  CRIS_CHECK_ARG(alpha > 0.f && p_drop >= 0.f && p_drop < 1.f, "attention: alpha=%g p_drop=%g", alpha, p_drop);
  CRIS_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
  a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.LkPad = (Lk + 7) / 8 * 8;
  a.alpha = alpha;
  a.scale_log2 = alpha * 1.4426950408889634f;
  a.drop_thresh = p_drop > 0.f ? drop_thresh16(p_drop) : 0u;
  a.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  a.seed = seed; a.seed_dev = seed_dev;
  return 0;
}

}  // namespace cris

using namespace cris;

extern "C" {

int cris_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                       float* lse, int B, int heads, int Lq, int Lk, float alpha, float p_drop, uint64_t seed,
                       const uint64_t* seed_dev, void* stream) {
  AttnArgs a{};
  if (int rc = fill_args(a, B, heads, Lq, Lk, alpha, p_drop, seed, seed_dev)) return rc;
  CRIS_CHECK_ARG(q && k && v && o, "attention_fwd: null argument");
  CRIS_CHECK_ARG((reinterpret_cast<uintptr_t>(o) & 15) == 0 && (ldo * 2) % 16 == 0, "attention_fwd: output not 16B aligned");
  CUtensorMap tq, tk, tv;
  if (int rc = at_tmap(&tq, q, ldq, B, Lq, heads)) return rc;
  if (int rc = at_tmap(&tk, k, ldk, B, Lk, heads)) return rc;
  if (int rc = at_tmap(&tv, v, ldv, B, Lk, heads)) return rc;
  a.O = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo; a.lse = lse;
  CRIS_SET_SMEM_ONCE(attn_fwd_kernel, FwdSmem::BYTES);
  dim3 grid((Lq + 127) / 128, heads, B);
  attn_fwd_kernel<<<grid, AT_THREADS, FwdSmem::BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, a);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                       int64_t ldo, const void* d_o, int64_t lddo, const float* lse, float* d_scratch, float* dq_acc,
                       int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int heads, int Lq, int Lk,
                       float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  AttnArgs a{};
  if (int rc = fill_args(a, B, heads, Lq, Lk, alpha, p_drop, seed, seed_dev)) return rc;
  CRIS_CHECK_ARG(q && k && v && o && d_o && lse && d_scratch && dq_acc && dk && dv, "attention_bwd: null argument");
  CRIS_CHECK_ARG((reinterpret_cast<uintptr_t>(dq_acc) & 15) == 0 && (lddq * 4) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(dk) & 15) == 0 && (lddk * 2) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(dv) & 15) == 0 && (lddv * 2) % 16 == 0,
                 "attention_bwd: gradient buffers not 16B aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tq, tk, tv, tdo;
  if (int rc = at_tmap(&tq, q, ldq, B, Lq, heads)) return rc;
  if (int rc = at_tmap(&tk, k, ldk, B, Lk, heads)) return rc;
  if (int rc = at_tmap(&tv, v, ldv, B, Lk, heads)) return rc;
  if (int rc = at_tmap(&tdo, d_o, lddo, B, Lq, heads)) return rc;
  const long long rows = (long long)B * heads * Lq;
  attn_bwd_prep_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(o), ldo,
                                                                           reinterpret_cast<const __nv_bfloat16*>(d_o), lddo,
                                                                           d_scratch, B, heads, Lq);
  CRIS_LAUNCH_OK();
  a.lse = const_cast<float*>(lse); a.Dsum = d_scratch;
  a.dQacc = dq_acc; a.lddq = lddq;
  a.dK = reinterpret_cast<__nv_bfloat16*>(dk); a.lddk = lddk;
  a.dV = reinterpret_cast<__nv_bfloat16*>(dv); a.lddv = lddv;
  CRIS_SET_SMEM_ONCE(attn_bwd_kernel, BwdSmem::BYTES);
  dim3 grid((Lk + 127) / 128, heads, B);
  attn_bwd_kernel<<<grid, ATB_THREADS, BwdSmem::BYTES, s>>>(tq, tk, tv, tdo, a);
  CRIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
