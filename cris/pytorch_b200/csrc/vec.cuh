// vec.cuh — 8-wide bf16 / fp32 vector access helpers for the HBM-bound elementwise kernels.
#pragma once
#include "common.cuh"

namespace cris {

// 8 consecutive bf16 (16-byte aligned) <-> 8 floats
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* v) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ void ld8f(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// dtype-flagged access: fp32 != 0 -> float storage, else bf16
__device__ __forceinline__ void ld8x(const void* base, long long idx, int fp32, float* v) {
  if (fp32) ld8f(reinterpret_cast<const float*>(base) + idx, v);
  else ld8(reinterpret_cast<const __nv_bfloat16*>(base) + idx, v);
}
__device__ __forceinline__ void st8x(void* base, long long idx, int fp32, const float* v) {
  if (fp32) st8f(reinterpret_cast<float*>(base) + idx, v);
  else st8(reinterpret_cast<__nv_bfloat16*>(base) + idx, v);
}
__device__ __forceinline__ void ld4x(const void* base, long long idx, int fp32, float* v) {
  if (fp32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
}
__device__ __forceinline__ void st4x(void* base, long long idx, int fp32, const float* v) {
  if (fp32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint2 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = q;
  }
}

// counter-based dropout RNG: keep iff hash(seed, index) >= p * 2^32.  The mask is a pure function of
// (seed, element index) so backward regenerates it instead of storing it.  The hash is murmur3's 32-bit finaliser
// over a 32-bit fold of (index, seed): ~10 integer instructions per element (the seed terms are loop invariant).
// The 64-bit splitmix variant used before cost ~25 and made the fused attention's softmax phase issue-bound.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t mix32(uint64_t x) {  // kept for callers that want one 32-bit word per 64-bit key
  return fmix32((uint32_t)x * 0x9E3779B1u + (uint32_t)(x >> 32) * 0x7FEB352Du);
}
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t idx, uint32_t thresh) {
  const uint32_t s = (uint32_t)seed * 0x27D4EB2Fu + (uint32_t)(seed >> 32) * 0x846CA68Bu;
  return fmix32((uint32_t)idx * 0x9E3779B1u + (uint32_t)(idx >> 32) * 0x7FEB352Du + s) >= thresh;
}
// Row/column form for 2-D score matrices (attention dropout): one hash per row, then one fmix32 per column PAIR —
// each column uses 16 bits of it (keep iff bits >= p * 2^16; p = 0.1 -> 6554 / 65536).  ~5 integer instructions per
// element instead of ~10.  cris_softmax_* and the fused attention kernels share this, so they draw the same masks.
__device__ __forceinline__ uint32_t drop_row_hash(uint64_t seed, uint64_t row) {
  const uint32_t s = (uint32_t)seed * 0x27D4EB2Fu + (uint32_t)(seed >> 32) * 0x846CA68Bu;
  return fmix32((uint32_t)row * 0x9E3779B1u + (uint32_t)(row >> 32) * 0x7FEB352Du + s);
}
__device__ __forceinline__ uint32_t drop_pair_bits(uint32_t row_hash, uint32_t col_pair) {  // col_pair = col >> 1
  return fmix32(row_hash + col_pair * 0x9E3779B1u);
}
__device__ __forceinline__ bool drop_keep_rc(uint32_t row_hash, uint32_t col, uint32_t thresh16) {
  const uint32_t h = drop_pair_bits(row_hash, col >> 1);
  return ((col & 1u) ? (h >> 16) : (h & 0xffffu)) >= thresh16;
}
__host__ __device__ inline uint32_t drop_thresh16(float p) {
  const double t = (double)p * 65536.0 + 0.5;
  return t >= 65535.0 ? 65535u : (uint32_t)t;
}
__host__ __device__ inline uint32_t drop_thresh(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
}

inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace cris
