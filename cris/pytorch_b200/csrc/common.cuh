// common.cuh — shared device/host helpers for libcris_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../../include/cris_b200.h"

namespace cris {

// ---- host-side error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define CRIS_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::cris::set_error(__VA_ARGS__);        \
      return -1;                             \
    }                                        \
  } while (0)

#define CRIS_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::cris::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                      \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

#define CRIS_LAUNCH_OK()                                                                \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::cris::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),     \
                        __FILE__, __LINE__);                                            \
      return -3;                                                                        \
    }                                                                                   \
    ::cris::count_launch();                                                             \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device).  The guard is an atomic bit set:
// the forward thread and autograd-engine threads may reach the same call site concurrently (setting the attribute
// twice is harmless, a torn plain bool is not defined behaviour), and the attribute is per-device state.
#define CRIS_SET_SMEM_ONCE(kern, bytes)                                                              \
  do {                                                                                               \
    static std::atomic<unsigned long long> _done{0};                                                 \
    int _dev = 0;                                                                                    \
    cudaGetDevice(&_dev);                                                                            \
    const unsigned long long _bit = 1ull << (_dev & 63);                                             \
    if (!(_done.load(std::memory_order_acquire) & _bit)) {                                           \
      CRIS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));  \
      _done.fetch_or(_bit, std::memory_order_release);                                               \
    }                                                                                                \
  } while (0)

// ---- small device helpers ---------------------------------------------------------------
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// padded-NHWC row -> is it an interior pixel?
__device__ __forceinline__ bool interior_row(long long r, int hp, int wp) {
  if (wp <= 0) return true;
  // row indices fit 32 bits in every use (B*(H+2)*(W+2) < 2^31): 32-bit division, not the 64-bit subroutine
  const unsigned rr = (unsigned)r % (unsigned)(hp * wp);
  const unsigned h = rr / (unsigned)wp, w = rr - h * (unsigned)wp;
  return (h >= 1u) && (h <= (unsigned)(hp - 2)) && (w >= 1u) && (w <= (unsigned)(wp - 2));
}

// n / d for 0 <= n < 2^31 and a divisor fixed at launch (1 <= d < 2^31): one 32x32->64 multiply and a shift instead
// of the ~20-instruction (32-bit) or ~100-instruction (64-bit) division sequences.  m = ceil(2^s / d),
// s = 31 + ceil(log2 d); exact for the stated range (error term < 2^-ceil(log2 d) <= 1/d).
struct FastDiv {
  unsigned m, s, d;
  FastDiv() : m(0x80000000u), s(31), d(1) {}
  explicit FastDiv(unsigned d_) : d(d_ ? d_ : 1) {
    unsigned sh = 0;
    while ((1ull << sh) < d) ++sh;
    s = 31 + sh;
    m = (unsigned)(((1ull << s) + d - 1) / d);
  }
  __device__ __forceinline__ unsigned div(unsigned n) const {
    return (unsigned)(((unsigned long long)n * m) >> s);
  }
};

// (h, w) of a padded-NHWC row index, advanced incrementally: kernels that walk rows with a constant stride test
// "interior pixel?" without the two integer divisions interior_row() costs per call
struct RowWalker {
  int h, w, hp, wp;
  __device__ __forceinline__ void init(long long row, int hp_, int wp_) {
    hp = hp_; wp = wp_; h = 0; w = 0;
    if (wp > 0) {
      const unsigned rr = (unsigned)row % (unsigned)(hp * wp);
      h = (int)(rr / (unsigned)wp);
      w = (int)(rr - (unsigned)h * (unsigned)wp);
    }
  }
  __device__ __forceinline__ bool interior() const {
    return wp <= 0 || (h >= 1 && h <= hp - 2 && w >= 1 && w <= wp - 2);
  }
  __device__ __forceinline__ void advance(int step) {
    if (wp > 0) {
      w += step;
      while (w >= wp) {
        w -= wp;
        if (++h == hp) h = 0;
      }
    }
  }
};

}  // namespace cris
