// conv_halo.cu — 3x3 convolution (stride 1, pad 1) for SMALL channel counts (Cin, Cout in {32, 64}) on padded-NHWC
// bf16 activations: the stem convs conv2/conv3 and the layer1 3x3 convs (reference: model/clip.py:17-25,165-182).
//
// EXPERIMENTAL (round-2 work, selected only with CRIS_B200_HALO_CONV=1; gemm_tc.cu remains the default path).
//
// Why a second kernel: gemm_tc.cu runs a 3x3 conv as nine row-shifted TMA loads per output tile.  With 32-64
// channels a 128-row tile is only 8-16 KB per tap, so every activation row is pulled out of L2 nine times and the
// layer is L2-bandwidth bound (stem convs: ~2 GB of L2 reads for 0.36 GB of HBM traffic, 10x off the HBM roofline).
// Here ONE "halo" tile of TM + 2(W+3) consecutive padded rows is loaded per TM = 128*SUB output rows, and the nine
// taps are nine UMMA descriptors whose start address is shifted by tap_off rows inside that tile.  That is legal
// because tcgen05 derives the swizzle phase from the absolute shared-memory address (tests/native/halo_probe.cu,
// measured on B200: every shift exact with the descriptor's base-offset field left 0).  The 9*Cout*Cin weights
// (<= 72 KB) are loaded once per CTA and stay resident.
//
// Roles (320 threads, one persistent CTA per SM): warps 0-7 epilogue (tcgen05.ld -> border mask -> bf16 -> 64-128 B
// row stores, optional BatchNorm column statistics), warp 8 TMA producer, warp 9 MMA issuer.  The fp32
// accumulators of the SUB sub-tiles live in TMEM, double-buffered across tiles.
#include "vec.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

#include <mutex>

namespace cris {

constexpr int HC_THREADS = 320;
constexpr int HC_BOX_ROWS = 64;  // rows per TMA box of the halo tile

struct HaloArgs {
  __nv_bfloat16* z;
  long long ldz;
  float* colstats;       // [min(64, ceil(rows/128))][2][Cout] or nullptr
  long long rows;        // N * (H+2) * (W+2)
  int hp, wp;            // padded image height / width
  int halo;              // wp + 1 rows on each side
  int ra_rows;           // rows of one A stage (multiple of HC_BOX_ROWS)
  int stages;
  int n_tiles;
  int cin_pad;           // column pitch of one tap inside the packed weight matrix
};

template <int CIN, int COUT, int SUB>
__global__ void __launch_bounds__(HC_THREADS, 1)
    conv_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ HaloArgs p) {
  constexpr int ROWB = CIN * 2;                       // bytes per activation row in shared memory
  constexpr int NACC = COUT < 32 ? 32 : COUT;         // TMEM columns of one sub-tile accumulator
  constexpr int TMEM_COLS = 2 * SUB * NACC;           // double-buffered across tiles (power of two, <= 512)
  constexpr int NCH = COUT / 32;                      // 32-column chunks per sub-tile
  constexpr int TM = 128 * SUB;
  constexpr uint64_t LAYOUT = CIN == 32 ? ptx::kLayoutSW64 : ptx::kLayoutSW128;
  constexpr uint32_t SBO = 8u * ROWB;
  constexpr int W_BYTES = 9 * COUT * ROWB;
  static_assert(CIN == 32 || CIN == 64, "halo conv: Cin is 32 or 64");
  static_assert(COUT == 32 || COUT == 64, "halo conv: Cout is 32 or 64");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32 && TMEM_COLS <= 512, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                           // [9][COUT][ROWB]
  uint8_t* sA = smem + ((W_BYTES + 1023) & ~1023);              // [stages][ra_rows][ROWB]
  const int a_stage_bytes = p.ra_rows * ROWB;                   // multiple of 4096
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + (size_t)p.stages * a_stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full = empty_bar + 8;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 8);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 9) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_bar, W_BYTES);
      for (int t = 0; t < 9; ++t) ptx::tma_load_4d(sW + t * COUT * ROWB, &tmW, w_bar, t * p.cin_pad, 0, 0, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u, 100 + s);
        ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)a_stage_bytes);
        const long long r0 = (long long)tile * TM - p.halo;  // may be negative: TMA zero-fills out-of-range rows
        uint8_t* dst = sA + (size_t)s * a_stage_bytes;
        for (int j = 0; j < p.ra_rows / HC_BOX_ROWS; ++j)
          ptx::tma_load_4d(dst + (size_t)j * HC_BOX_ROWS * ROWB, &tmX, &full_bar[s], 0, (int)(r0 + j * HC_BOX_ROWS), 0, 0);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, COUT, 0, 0);
      ptx::mbar_wait(w_bar, 0, 90);
      ptx::tc_fence_after();
      const uint32_t w_base = ptx::smem_u32(sW);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const int buf = it & 1;
        ptx::mbar_wait(&tmem_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u, 200 + buf);
        ptx::mbar_wait(&full_bar[s], ph, 110 + s);
        ptx::tc_fence_after();
        const uint32_t a_base = ptx::smem_u32(sA + (size_t)s * a_stage_bytes);
#pragma unroll 1
        for (int sub = 0; sub < SUB; ++sub) {
          const uint32_t tacc = tmem_base + (uint32_t)((buf * SUB + sub) * NACC);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            const int off = (t / 3 - 1) * p.wp + (t % 3 - 1);
            const uint32_t a_row = a_base + (uint32_t)((sub * 128 + p.halo + off) * ROWB);
            const uint32_t b_row = w_base + (uint32_t)(t * COUT * ROWB);
#pragma unroll
            for (int kk = 0; kk < CIN / 16; ++kk) {
              const uint64_t adesc = ptx::make_smem_desc(a_row + kk * 32, 16, SBO, LAYOUT);
              const uint64_t bdesc = ptx::make_smem_desc(b_row + kk * 32, 16, SBO, LAYOUT);
              ptx::umma_bf16(tacc, adesc, bdesc, idesc, (t > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        ptx::umma_commit(&empty_bar[s]);    // the halo tile may be overwritten once these MMAs have read it
        ptx::umma_commit(&tmem_full[buf]);  // accumulators of this tile complete
      }
    }
  } else {
    // ===================== epilogue (warps 0-7) =====================
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter, and which half of the (sub, chunk) pairs
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      ptx::mbar_wait(&tmem_full[buf], ((uint32_t)it >> 1) & 1u, 300 + buf);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int pair = half; pair < SUB * NCH; pair += 2) {
        const int sub = pair / NCH, c = pair - sub * NCH;
        const long long row = (long long)tile * TM + sub * 128 + q * 32 + lane;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (uint32_t)((buf * SUB + sub) * NACC + c * 32) + ((uint32_t)(q * 32) << 16), r);
        ptx::tmem_ld_wait();
        const bool in = row < p.rows;
        const bool valid = in && interior_row(row, p.hp, p.wp);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? __uint_as_float(r[j]) : 0.f;
        if (in) {
          uint4* dp = reinterpret_cast<uint4*>(p.z + row * p.ldz + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
            o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            dp[j] = o;
          }
        }
        if (p.colstats != nullptr && (long long)tile * TM + sub * 128 < p.rows) {  // (partial rows exist per 128-row block)
          // per-column (sum, sumsq) of the stored (bf16-rounded) values over the warp's 32 rows: transposing
          // butterfly, lane L ends with column L; one atomic pair per lane into the tile's partial row
          float a[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xr = bf2f(f2bf(v[j]));
            a[j] = xr;
            s2[j] = xr * xr;
          }
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < sft; ++i) {
              const float send_a = up ? a[i] : a[i + sft];
              const float keep_a = up ? a[i + sft] : a[i];
              a[i] = keep_a + __shfl_xor_sync(0xffffffffu, send_a, sft);
              const float send_q = up ? s2[i] : s2[i + sft];
              const float keep_q = up ? s2[i + sft] : s2[i];
              s2[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, sft);
            }
          }
          const long long mt = (long long)tile * SUB + sub;  // index of this 128-row block
          float* dst = p.colstats + (size_t)(mt & 63) * 2 * COUT + c * 32 + lane;
          atomicAdd(dst, a[0]);
          atomicAdd(dst + COUT, s2[0]);
        }
      }
      // all TMEM reads of this tile are done: hand the accumulator buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// =========================== host side ===========================================
static PFN_cuTensorMapEncodeTiled_v12000 halo_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

static int halo_tmap(CUtensorMap* tm, const void* base, long long cols, long long rows, long long ld, int box_cols,
                     int box_rows, CUtensorMapSwizzle swz) {
  auto fn = halo_encode_fn();
  CRIS_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), (cuuint64_t)(ld * 2 * rows), (cuuint64_t)(ld * 2 * rows)};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CRIS_CHECK_ARG(r == CUDA_SUCCESS, "halo conv: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int CIN, int COUT, int SUB>
static int launch_halo(const CUtensorMap& tmX, const CUtensorMap& tmW, HaloArgs p, int smem_bytes, cudaStream_t stream) {
  auto kern = conv_halo_kernel<CIN, COUT, SUB>;
  CRIS_SET_SMEM_ONCE(kern, 227 * 1024);  // per instantiation and device
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  kern<<<grid, HC_THREADS, smem_bytes, stream>>>(tmX, tmW, p);
  CRIS_LAUNCH_OK();
  return 0;
}

template <int CIN, int COUT>
static int plan_halo(const void* x, int64_t ldx, const void* w, int64_t ldw, HaloArgs p, cudaStream_t stream) {
  constexpr int ROWB = CIN * 2;
  constexpr int W_BYTES = ((9 * COUT * ROWB + 1023) / 1024) * 1024;
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - 512 /*barriers*/ - W_BYTES;
  // widest tile (fewest halo re-reads) that still leaves room for two stages
  int sub = 0, ra = 0, stages = 0;
  for (int s : {4, 2, 1}) {
    const int rows = ((128 * s + 2 * p.halo + HC_BOX_ROWS - 1) / HC_BOX_ROWS) * HC_BOX_ROWS;
    const int st = budget / (rows * ROWB);
    if (st >= 2) {
      sub = s; ra = rows; stages = st > 4 ? 4 : st;
      break;
    }
  }
  CRIS_CHECK_ARG(sub > 0, "halo conv: image width %d needs a halo tile that does not fit in shared memory", p.wp - 2);
  p.ra_rows = ra;
  p.stages = stages;
  p.n_tiles = (int)((p.rows + 128 * sub - 1) / (128 * sub));
  CUtensorMap tmX, tmW;
  const CUtensorMapSwizzle swz = CIN == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = halo_tmap(&tmX, x, CIN, p.rows, ldx, CIN, HC_BOX_ROWS, swz);
  if (rc) return rc;
  rc = halo_tmap(&tmW, w, ldw, COUT, ldw, CIN, COUT, swz);
  if (rc) return rc;
  const int smem = 1024 + W_BYTES + stages * ra * ROWB + 512;
  if (sub == 4) return launch_halo<CIN, COUT, 4>(tmX, tmW, p, smem, stream);
  if (sub == 2) return launch_halo<CIN, COUT, 2>(tmX, tmW, p, smem, stream);
  return launch_halo<CIN, COUT, 1>(tmX, tmW, p, smem, stream);
}

}  // namespace cris

using namespace cris;

extern "C" {

int cris_conv3x3_halo(const void* x, int64_t ldx, const void* w_packed, int64_t ldw, int cin_pad, void* z, int64_t ldz,
                      float* colstats, int N, int H, int W, int Cin, int Cout, void* stream) {
  CRIS_CHECK_ARG(x && w_packed && z, "conv3x3_halo: null argument");
  CRIS_CHECK_ARG((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64), "conv3x3_halo: Cin=%d Cout=%d (32 or 64 only)", Cin,
                 Cout);
  CRIS_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(z)) & 15) == 0 &&
                     (ldx * 2) % 16 == 0 && (ldw * 2) % 16 == 0 && (ldz * 2) % 16 == 0 && ldz >= Cout && ldx >= Cin,
                 "conv3x3_halo: operands must be 16-byte aligned with 16-byte pitches");
  CRIS_CHECK_ARG(cin_pad >= Cin && ldw >= 9 * (int64_t)cin_pad, "conv3x3_halo: packed weight pitch");
  HaloArgs p;
  p.z = reinterpret_cast<__nv_bfloat16*>(z);
  p.ldz = ldz;
  p.colstats = colstats;
  p.hp = H + 2;
  p.wp = W + 2;
  p.rows = (long long)N * p.hp * p.wp;
  CRIS_CHECK_ARG(p.rows < (1ll << 31), "conv3x3_halo: too many rows");
  p.halo = p.wp + 1;
  p.cin_pad = cin_pad;
  p.ra_rows = p.stages = p.n_tiles = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (Cin == 32 && Cout == 32) return plan_halo<32, 32>(x, ldx, w_packed, ldw, p, s);
  if (Cin == 32 && Cout == 64) return plan_halo<32, 64>(x, ldx, w_packed, ldw, p, s);
  if (Cin == 64 && Cout == 32) return plan_halo<64, 32>(x, ldx, w_packed, ldw, p, s);
  return plan_halo<64, 64>(x, ldx, w_packed, ldw, p, s);
}

}  // extern "C"
