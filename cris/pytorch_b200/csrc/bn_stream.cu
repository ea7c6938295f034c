// bn_stream.cu — the three HBM-bound BatchNorm passes as shared-memory-staged streaming kernels.
//
//   bn_apply_stream      y  = mask(relu?(z*scale + shift (+ resid)))                      (forward apply)
//   bn_reduce_stream     (sum dz, sum dz*xhat) or (sum x, sum x^2) per channel            (backward / forward statistics)
//   bn_bwd_apply_stream  dx = gamma*invstd*(dz - s0/n - xhat*s1/n), dres (+)= dz          (backward apply)
//
// Reference call sites: nn.BatchNorm2d / SyncBatchNorm forward and backward (model/clip.py:18-26,171-183;
// model/layers.py:8-16,262; train.py:97-98).  Same arithmetic as the register-streaming kernels of norm.cu; what
// changes is the data movement.  Those kernels keep their loads in registers (128 registers -> 2 blocks / SM ->
// ~49 KB in flight per SM) and ran at 3.0-4.3 TB/s.  Here one warp issues bulk asynchronous copies
// (cp.async.bulk global -> shared, completion on an mbarrier) of whole row tiles, STAGES tiles ahead: bytes in
// flight are decoupled from registers (2 blocks x 3 tiles x up to 24 KB per SM), the 256 consumer threads read the
// tile with conflict-free 16-byte shared loads, and results leave through coalesced 16-byte stores.
// Every operand is a [rows, C] bf16 row matrix with its own pitch (dense, or a column slice of a concat buffer:
// then the tile is fetched row by row).  C must be a power of two >= 8 (every BatchNorm of the model).
#include "vec.cuh"
#include "ptx.cuh"

namespace cris {

constexpr int BS_THREADS = 256;
constexpr int BS_STAGES = 4;       // 4 stages x 3 operands x 8 KB = 96 KB per block, two blocks per SM (measured: three
                                   // blocks with 3 stages is slower, profiles/r02_bn_bench.txt)
constexpr int BS_TILE_BYTES = 8192;   // per operand and stage
constexpr int BS_MAX_OPS = 3;

struct BsOperand {
  const __nv_bfloat16* p;
  long long ld;
};

struct BsArgs {
  BsOperand in[BS_MAX_OPS];
  int n_in;
  long long rows;
  int C, logG;       // G = C / 8 = 1 << logG  (16-byte groups per row)
  int TR;            // rows per tile (multiple of 256 / G when G < 256)
  int hp, wp;        // padded-NHWC geometry (0 = every row is interior)
  FastDiv dHW, dW;
  int relu;
  // forward apply
  const float* scale; const float* shift;
  __nv_bfloat16* y; long long ldy;
  // reductions
  int mode;          // 0: (x, x^2)   1: (dz, dz*xhat)
  int mask_from_y;   // 1: operand 2 is y (ReLU mask = y > 0); 0: mask recomputed as z*scale+shift > 0
  const float* mean; const float* invstd;
  float* partials;
  int n_part;        // partial rows the caller zeroed: block b accumulates into row b % n_part
  // backward apply
  const float* gamma; const float* beta; const float* sums; float inv_count;
  __nv_bfloat16* dx; long long lddx;
  __nv_bfloat16* dres; long long lddres; int dres_accumulate;
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes),
                 "r"(ptx::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool bs_interior(const BsArgs& p, long long row) {
  if (p.wp <= 0) return true;
  const unsigned r = (unsigned)row;
  const unsigned rr = r - p.dHW.div(r) * (unsigned)(p.hp * p.wp);
  const unsigned h = p.dW.div(rr), w = rr - h * (unsigned)p.wp;
  return (h >= 1u) && (h <= (unsigned)(p.hp - 2)) && (w >= 1u) && (w <= (unsigned)(p.wp - 2));
}

// warp 0: fetch tile `tile` of every input operand into stage `s`
__device__ __forceinline__ void bs_issue(const BsArgs& p, uint8_t* smem, uint64_t* full, int s, long long tile, int lane) {
  const long long r0 = tile * p.TR;
  const int nr = (int)min((long long)p.TR, p.rows - r0);
  const uint32_t row_bytes = (uint32_t)p.C * 2u;
  if (lane == 0) ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)nr * row_bytes * (uint32_t)p.n_in);
  __syncwarp();
  for (int o = 0; o < p.n_in; ++o) {
    uint8_t* dst = smem + ((size_t)s * BS_MAX_OPS + o) * BS_TILE_BYTES;
    const __nv_bfloat16* src = p.in[o].p + r0 * p.in[o].ld;
    if (p.in[o].ld == p.C) {
      if (lane == 0) bulk_g2s(dst, src, (uint32_t)nr * row_bytes, &full[s]);
    } else {
      for (int r = lane; r < nr; r += 32) bulk_g2s(dst + (size_t)r * row_bytes, src + (long long)r * p.in[o].ld, row_bytes, &full[s]);
    }
  }
}

__device__ __forceinline__ void ld8s(const uint8_t* tile, int item, float* v) {
  const uint4 q = *reinterpret_cast<const uint4*>(tile + (size_t)item * 16);
  float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// KIND 0: forward apply   1: reduction   2: backward apply
template <int KIND>
__global__ void __launch_bounds__(BS_THREADS, 2) bn_stream_kernel(const BsArgs p) {
  extern __shared__ __align__(128) uint8_t bs_smem[];
  __shared__ uint64_t full[BS_STAGES];
  // the cross-row-lane reduction of KIND 1 reuses the (by then idle) tile buffers: no static shared memory
  float* red = reinterpret_cast<float*>(bs_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = 1 << p.logG;
  const long long ntiles = (p.rows + p.TR - 1) / p.TR;
  if (tid == 0) {
    for (int s = 0; s < BS_STAGES; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // this thread's fixed 8-channel group (256 % G == 0: the group does not change from item to item)
  const int g = tid & (G - 1);
  const int c = g * 8;
  const int rlane = tid >> p.logG;          // first row of the tile this thread touches
  const int rstep = BS_THREADS >> p.logG;   // rows between two of its items (G <= 256)
  float c0[8], c1[8], c2[8], c3[8];         // per-channel coefficients, meaning depends on KIND
  if (KIND == 0) {
    ld8f(p.scale + c, c0);
    ld8f(p.shift + c, c1);
  } else if (KIND == 1) {
    if (p.mode == 1) {
      ld8f(p.mean + c, c0);
      if (p.relu && !p.mask_from_y) { ld8f(p.scale + c, c1); ld8f(p.shift + c, c2); }
    }
  } else {
    // dx = A*dz + B*x + K ; ReLU mask (no residual) = fma(x, A, shf) > 0      (A = gamma*invstd)
    float mu[8], is[8], ga[8], be[8], s0[8], s1[8];
    ld8f(p.mean + c, mu); ld8f(p.invstd + c, is); ld8f(p.gamma + c, ga); ld8f(p.beta + c, be);
    ld8f(p.sums + c, s0); ld8f(p.sums + p.C + c, s1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float A = ga[k] * is[k];
      const float B = -A * is[k] * s1[k] * p.inv_count;
      c0[k] = A;
      c1[k] = B;
      c2[k] = -A * s0[k] * p.inv_count - B * mu[k];
      c3[k] = be[k] - mu[k] * A;
    }
  }
  float acc0[8], acc1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc0[k] = acc1[k] = 0.f;

  // prologue: STAGES-1 tiles in flight
  if (warp == 0) {
    for (int j = 0; j < BS_STAGES - 1; ++j) {
      const long long t = (long long)blockIdx.x + (long long)j * gridDim.x;
      if (t < ntiles) bs_issue(p, bs_smem, full, j, t, lane);
    }
  }
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % BS_STAGES;
    if (warp == 0) {
      const long long t = tile + (long long)(BS_STAGES - 1) * gridDim.x;
      if (t < ntiles) bs_issue(p, bs_smem, full, (it + BS_STAGES - 1) % BS_STAGES, t, lane);
    }
    ptx::mbar_wait(&full[s], (uint32_t)(it / BS_STAGES) & 1u, 700 + s);
    const uint8_t* t0 = bs_smem + ((size_t)s * BS_MAX_OPS + 0) * BS_TILE_BYTES;
    const uint8_t* t1 = t0 + BS_TILE_BYTES;
    const uint8_t* t2 = t1 + BS_TILE_BYTES;
    const long long r0 = tile * p.TR;
    const int nr = (int)min((long long)p.TR, p.rows - r0);
#pragma unroll 2
    for (int r = rlane; r < nr; r += rstep) {
      const long long row = r0 + r;
      const int item = (r << p.logG) + g;
      const bool in = bs_interior(p, row);
      if (KIND == 0) {
        float v[8];
        if (in) {
          ld8s(t0, item, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], c0[k], c1[k]);
          if (p.n_in > 1) {
            float rv[8];
            ld8s(t1, item, rv);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += rv[k];
          }
          if (p.relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = 0.f;
        }
        st8(p.y + row * p.ldy + c, v);
      } else if (KIND == 1) {
        if (!in) {
          if (p.dx != nullptr) {  // masked-gradient output: border rows are zero
            const float zz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            st8(p.dx + row * p.lddx + c, zz);
          }
          continue;
        }
        float a[8];
        ld8s(t0, item, a);
        if (p.mode == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { acc0[k] += a[k]; acc1[k] = fmaf(a[k], a[k], acc1[k]); }
        } else {
          float xv[8];
          ld8s(t1, item, xv);
          if (p.relu) {
            if (p.mask_from_y) {
              float yv[8];
              ld8s(t2, item, yv);
#pragma unroll
              for (int k = 0; k < 8; ++k) if (!(yv[k] > 0.f)) a[k] = 0.f;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) if (!(fmaf(xv[k], c1[k], c2[k]) > 0.f)) a[k] = 0.f;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) { acc0[k] += a[k]; acc1[k] = fmaf(a[k], xv[k] - c0[k], acc1[k]); }
          // optional: the ReLU-masked gradient dz = dy * (y > 0) leaves with the reduction — it IS the gradient of the
          // residual branch, and the apply pass then reads it instead of (dy, y): one tensor pass less per layer
          if (p.dx != nullptr) st8(p.dx + row * p.lddx + c, a);
        }
      } else {
        float o[8], dz[8];
        if (in) {
          float xv[8];
          ld8s(t0, item, dz);
          ld8s(t1, item, xv);
          if (p.relu) {
            if (p.mask_from_y) {
              float yv[8];
              ld8s(t2, item, yv);
#pragma unroll
              for (int k = 0; k < 8; ++k) if (!(yv[k] > 0.f)) dz[k] = 0.f;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) if (!(fmaf(xv[k], c0[k], c3[k]) > 0.f)) dz[k] = 0.f;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(c0[k], dz[k], fmaf(c1[k], xv[k], c2[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = dz[k] = 0.f;
        }
        st8(p.dx + row * p.lddx + c, o);
        if (p.dres != nullptr) {
          if (p.dres_accumulate) {
            float old[8];
            ld8(p.dres + row * p.lddres + c, old);
#pragma unroll
            for (int k = 0; k < 8; ++k) dz[k] += old[k];
          }
          st8(p.dres + row * p.lddres + c, dz);
        }
      }
    }
    __syncthreads();  // every thread is done with stage s: warp 0 may refill it in the next iteration
  }

  if (KIND == 1) {
    if (p.mode == 1) {  // the per-channel 1/std factor of xhat is applied once, not per element
      float rs[8];
      ld8f(p.invstd + c, rs);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc1[k] *= rs[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[tid * 16 + k] = acc0[k]; red[tid * 16 + 8 + k] = acc1[k]; }
    __syncthreads();
    if (tid < G) {
      for (int rr = 1; rr < rstep; ++rr) {
        const int o = (rr * G + tid) * 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc0[k] += red[o + k]; acc1[k] += red[o + 8 + k]; }
      }
      // n_part (<= 64) partial rows shared by all blocks (vector fp32 atomics into a caller-zeroed buffer)
      float* dst = p.partials + (size_t)(blockIdx.x % p.n_part) * 2 * p.C;
      atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(acc0[0], acc0[1], acc0[2], acc0[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + c + 4), make_float4(acc0[4], acc0[5], acc0[6], acc0[7]));
      atomicAdd(reinterpret_cast<float4*>(dst + p.C + c), make_float4(acc1[0], acc1[1], acc1[2], acc1[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + p.C + c + 4), make_float4(acc1[4], acc1[5], acc1[6], acc1[7]));
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
static int stream_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Can the streaming kernels take this problem?  (power-of-two C in [8, 2048], 16-byte aligned operands, < 2^31 rows)
bool bn_stream_ok(long long rows, int C, int hp, int wp) {
  if (C < 8 || C > 2048 || (C & (C - 1)) != 0) return false;
  if (rows < 2048 || rows >= (1ll << 31)) return false;
  if (hp < 0 || wp < 0) return false;
  const char* e = getenv("CRIS_B200_BN_STREAM");  // read per call (tests compare both paths); default on
  return !(e != nullptr && e[0] == '0');
}

static bool op_ok(const void* p, long long ld, int C) {
  return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld >= C && (ld * 2) % 16 == 0;
}

template <int KIND>
static int launch_stream(BsArgs& a, cudaStream_t s) {
  int logG = 0;
  while ((8 << logG) < a.C) ++logG;
  a.logG = logG;
  const int G = 1 << logG;
  const int rstep = BS_THREADS / G > 0 ? BS_THREADS / G : 1;
  int TR = BS_TILE_BYTES / (a.C * 2);
  if (TR < 1) TR = 1;
  if (TR < rstep) TR = rstep;                // C*2*rstep = 4096 <= tile bytes: always fits
  a.TR = TR;
  a.dHW = FastDiv((unsigned)(a.hp > 0 && a.wp > 0 ? a.hp * a.wp : 1));
  a.dW = FastDiv((unsigned)(a.wp > 0 ? a.wp : 1));
  for (int o = 0; o < a.n_in; ++o)
    CRIS_CHECK_ARG(op_ok(a.in[o].p, a.in[o].ld, a.C), "bn_stream: operand %d is not a 16-byte aligned [rows, %d] matrix", o, a.C);
  const int smem = BS_STAGES * BS_MAX_OPS * BS_TILE_BYTES;
  CRIS_SET_SMEM_ONCE(bn_stream_kernel<KIND>, smem);
  const long long ntiles = (a.rows + TR - 1) / TR;
  const long long want = 2ll * stream_sms();
  const int grid = (int)(ntiles < want ? ntiles : want);
  bn_stream_kernel<KIND><<<grid, BS_THREADS, smem, s>>>(a);
  CRIS_LAUNCH_OK();
  return 0;
}

int bn_apply_stream(const void* x, long long ldx, const float* scale, const float* shift, const void* resid, long long ldr,
                    void* y, long long ldy, long long rows, int C, int relu, int hp, int wp, cudaStream_t s) {
  BsArgs a{};
  a.in[0] = {reinterpret_cast<const __nv_bfloat16*>(x), ldx};
  a.n_in = 1;
  if (resid != nullptr) { a.in[1] = {reinterpret_cast<const __nv_bfloat16*>(resid), ldr}; a.n_in = 2; }
  a.rows = rows; a.C = C; a.hp = hp; a.wp = wp; a.relu = relu;
  a.scale = scale; a.shift = shift;
  a.y = reinterpret_cast<__nv_bfloat16*>(y); a.ldy = ldy;
  CRIS_CHECK_ARG(op_ok(y, ldy, C), "bn_apply: output is not a 16-byte aligned [rows, %d] matrix", C);
  return launch_stream<0>(a, s);
}

int bn_reduce_stream(int mode, const void* a0, long long lda, const void* y, long long ldy, const void* x, long long ldx,
                     const float* mean, const float* invstd, const float* scale, const float* shift, long long rows, int C,
                     int relu, int hp, int wp, float* partials, int n_part, cudaStream_t s, void* dzm_out = nullptr,
                     long long lddzm = 0) {
  BsArgs a{};
  a.in[0] = {reinterpret_cast<const __nv_bfloat16*>(a0), lda};
  a.n_in = 1;
  a.mode = mode;
  if (mode == 1) {
    a.in[1] = {reinterpret_cast<const __nv_bfloat16*>(x), ldx};
    a.n_in = 2;
    if (relu && y != nullptr) { a.in[2] = {reinterpret_cast<const __nv_bfloat16*>(y), ldy}; a.n_in = 3; a.mask_from_y = 1; }
  }
  a.rows = rows; a.C = C; a.hp = hp; a.wp = wp; a.relu = relu;
  a.mean = mean; a.invstd = invstd; a.scale = scale; a.shift = shift; a.partials = partials;
  a.dx = reinterpret_cast<__nv_bfloat16*>(dzm_out); a.lddx = lddzm;
  CRIS_CHECK_ARG(dzm_out == nullptr || (mode == 1 && op_ok(dzm_out, lddzm, C)), "bn_reduce: masked-gradient output misaligned");
  a.n_part = n_part < 1 ? 1 : (n_part > 64 ? 64 : n_part);
  return launch_stream<1>(a, s);
}

int bn_bwd_apply_stream(const void* dy, long long lddy, const void* y, long long ldy, const void* x, long long ldx,
                        const float* mean, const float* invstd, const float* gamma, const float* beta, const float* sums,
                        double count, void* dx, long long lddx, void* dres, long long lddres, int dres_accumulate,
                        long long rows, int C, int relu, int hp, int wp, cudaStream_t s) {
  BsArgs a{};
  a.in[0] = {reinterpret_cast<const __nv_bfloat16*>(dy), lddy};
  a.in[1] = {reinterpret_cast<const __nv_bfloat16*>(x), ldx};
  a.n_in = 2;
  if (relu && y != nullptr) { a.in[2] = {reinterpret_cast<const __nv_bfloat16*>(y), ldy}; a.n_in = 3; a.mask_from_y = 1; }
  a.rows = rows; a.C = C; a.hp = hp; a.wp = wp; a.relu = relu;
  a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.beta = beta; a.sums = sums; a.inv_count = (float)(1.0 / count);
  a.dx = reinterpret_cast<__nv_bfloat16*>(dx); a.lddx = lddx;
  a.dres = reinterpret_cast<__nv_bfloat16*>(dres); a.lddres = lddres; a.dres_accumulate = dres_accumulate;
  CRIS_CHECK_ARG(op_ok(dx, lddx, C), "bn_bwd_apply: dx is not a 16-byte aligned [rows, %d] matrix", C);
  CRIS_CHECK_ARG(dres == nullptr || op_ok(dres, lddres, C), "bn_bwd_apply: dres is not a 16-byte aligned [rows, %d] matrix", C);
  return launch_stream<2>(a, s);
}

}  // namespace cris
