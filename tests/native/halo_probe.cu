// halo_probe.cu — experiment for the planned "halo tile" loader of the small-channel 3x3 convolutions
// (DESIGN.md §9): can ONE swizzled TMA tile of padded-NHWC rows serve all nine taps, each tap being a UMMA
// shared-memory descriptor whose start address is shifted by `d` rows?  The open question is how tcgen05 derives
// the swizzle phase when the descriptor start is not aligned to the swizzle repeat (512 B for SWIZZLE_64B,
// 1024 B for SWIZZLE_128B): from the absolute shared-memory address, or relative to the start address plus the
// descriptor's 3-bit "matrix base offset" field (bits 49-51).  The probe computes
//     D_d[m][n] = sum_k A[m + d][k] * B[n][k]      (M = 128, N = 32, K = channels)
// for several d with both encodings and reports which one reproduces the host result.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include tests/native/halo_probe.cu
//        -o tests/native/halo_probe -lcuda        Run: tests/native/halo_probe   (needs a B200)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../cris/pytorch_b200/csrc/ptx.cuh"

using namespace cris;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);                \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int kRowsBox = 256;  // rows of the halo tile (128 output rows + up to 128 rows of shift)
constexpr int kN = 32;
constexpr int kMaxShifts = 8;

struct ProbeArgs {
  int C;             // channels per row: 32 (64-byte rows, SWIZZLE_64B) or 64 (128-byte rows, SWIZZLE_128B)
  int n_shifts;
  int shift[kMaxShifts];
  int mode;          // 0: base-offset field = 0; 1: base-offset = (start >> 7) & 7
  float* D;          // [n_shifts][128][32]
};

__global__ void __launch_bounds__(128) halo_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB, ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int row_bytes = p.C * 2;
  uint8_t* sA = smem;                                  // kRowsBox x row_bytes
  uint8_t* sB = smem + kRowsBox * 128;                 // kN x row_bytes (1024-aligned: 256*128 = 32 KB)
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sB + kN * 128);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<32>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar_load, (uint32_t)(kRowsBox * row_bytes + kN * row_bytes));
    ptx::tma_load_4d(sA, &tmA, bar_load, 0, 0, 0, 0);
    ptx::tma_load_4d(sB, &tmB, bar_load, 0, 0, 0, 0);
  }
  ptx::mbar_wait(bar_load, 0, 1);
  const uint64_t layout = p.C == 32 ? ptx::kLayoutSW64 : ptx::kLayoutSW128;
  const uint32_t sbo = 8u * (uint32_t)row_bytes;  // 8-row groups are contiguous: the tile is linear in the row index
  constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kN, 0, 0);
  for (int s = 0; s < p.n_shifts; ++s) {
    if (threadIdx.x == 0) {
      const uint32_t a0 = ptx::smem_u32(sA) + (uint32_t)p.shift[s] * (uint32_t)row_bytes;
      const uint32_t b0 = ptx::smem_u32(sB);
      for (int kk = 0; kk < p.C / 16; ++kk) {
        const uint32_t a_addr = a0 + (uint32_t)kk * 32u;
        uint64_t adesc = ptx::make_smem_desc(a_addr, 16, sbo, layout);
        if (p.mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t bdesc = ptx::make_smem_desc(b0 + (uint32_t)kk * 32u, 16, sbo, layout);
        ptx::umma_bf16(tmem, adesc, bdesc, idesc, kk > 0 ? 1u : 0u);
      }
      ptx::umma_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, (uint32_t)(s & 1), 2);
    ptx::tc_fence_after();
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    float* out = p.D + ((size_t)s * 128 + warp * 32 + lane) * kN;
    for (int j = 0; j < kN; ++j) out[j] = __uint_as_float(r[j]);
    ptx::tc_fence_before();
    __syncthreads();  // every warp has drained the accumulator before the next shift overwrites it
    ptx::tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<32>(tmem);
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) {
    printf("cuTensorMapEncodeTiled unavailable\n");
    exit(2);
  }
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
}

static CUtensorMap make_map(const void* base, int C, int rows, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)rows, 1, 1};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * rows, (cuuint64_t)C * 2 * rows};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(2);
  }
  return tm;
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    printf("halo_probe needs sm_100 (found sm_%d%d)\n", prop.major, prop.minor);
    return 2;
  }
  const int shifts[kMaxShifts] = {0, 1, 2, 3, 8, 13, 107, 127};
  int verdict = 0;
  for (int C : {32, 64}) {
    std::vector<__nv_bfloat16> hA((size_t)kRowsBox * C), hB((size_t)kN * C);
    uint32_t st = 12345u + (uint32_t)C;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (float)((int)((st >> 16) % 9) - 4); };  // exact in bf16
    for (auto& v : hA) v = __float2bfloat16(rnd());
    for (auto& v : hB) v = __float2bfloat16(rnd());
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, (size_t)kMaxShifts * 128 * kN * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    const CUtensorMap tmA = make_map(dA, C, kRowsBox, kRowsBox), tmB = make_map(dB, C, kN, kN);
    const int smem = kRowsBox * 128 + kN * 128 + 64 + 1024;
    CK(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 2; ++mode) {
      ProbeArgs p;
      p.C = C; p.n_shifts = kMaxShifts; p.mode = mode; p.D = dD;
      memcpy(p.shift, shifts, sizeof(shifts));
      CK(cudaMemset(dD, 0xff, (size_t)kMaxShifts * 128 * kN * 4));
      halo_probe_kernel<<<1, 128, smem>>>(tmA, tmB, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("C=%d mode=%d: kernel failed: %s\n", C, mode, cudaGetErrorString(e));
        return 3;
      }
      std::vector<float> hD((size_t)kMaxShifts * 128 * kN);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      printf("C=%d (%s) base-offset mode %d:", C, C == 32 ? "SW64" : "SW128", mode);
      bool all = true;
      for (int s = 0; s < kMaxShifts; ++s) {
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < kN; ++n) {
            float ref = 0.f;
            for (int k = 0; k < C; ++k)
              ref += __bfloat162float(hA[(size_t)(m + shifts[s]) * C + k]) * __bfloat162float(hB[(size_t)n * C + k]);
            if (hD[((size_t)s * 128 + m) * kN + n] != ref) ++bad;
          }
        printf("  d=%d:%s", shifts[s], bad ? "MISMATCH" : "ok");
        all = all && bad == 0;
      }
      printf("  => %s\n", all ? "ALL SHIFTS OK" : "not usable as is");
      if (all) verdict |= 1 << (mode + (C == 64 ? 2 : 0));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf("HALO_PROBE verdict bits (SW64 mode0, SW64 mode1, SW128 mode0, SW128 mode1): %d%d%d%d\n", verdict & 1,
         (verdict >> 1) & 1, (verdict >> 2) & 1, (verdict >> 3) & 1);
  return 0;
}
