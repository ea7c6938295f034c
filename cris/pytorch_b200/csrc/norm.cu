// norm.cu — BatchNorm (batch-statistics, two-phase) and LayerNorm kernels, forward and backward.
// Reference call sites: nn.BatchNorm2d/1d + SyncBatchNorm (model/clip.py:18-26,171-183;
// model/layers.py:8-16,262; train.py:97-98), nn.LayerNorm (model/clip.py:226-231,
// model/layers.py:199-216).  All HBM-bound: 16-byte vector accesses, one pass per tensor.
#include <algorithm>

#include "vec.cuh"

namespace cris {

// ---------------------------------------------------------------------------------------------
// column reductions over a [rows, C] matrix -> partials[block % 64][2][C] (atomic accumulation)
//   MODE 0: (x, x^2)                        batch statistics of a tensor not produced by the GEMM
//   MODE 1: (dz, dz*xhat), dz = dy*(y>0)    BatchNorm backward
//   MODE 2: (dy, -)                         bias gradient
//   MODE 3: (dy, dy*xhat) with per-ROW mean/rstd   LayerNorm gamma/beta gradients
// ---------------------------------------------------------------------------------------------
struct ColReduceArgs {
  const void* a;  long long lda; int a_fp32;     // x (mode 0) or dy
  const void* a2; long long lda2;                // optional second dy (bf16) added to a (mode 3)
  const void* y;  long long ldy;                 // post-activation output (mode 1, relu mask), bf16
  const void* x;  long long ldx; int x_fp32;     // pre-normalisation input (modes 1, 3)
  const float* mean; const float* rstd;          // per-column (mode 1) or per-row (mode 3)
  const float* scale; const float* shift;        // mode 1 with y == NULL: ReLU mask recomputed as x*scale+shift > 0
  long long rows; int C; int relu; int hp, wp;
  float* partials;
};

template <int MODE>
__global__ void __launch_bounds__(256, 2) col_reduce_kernel(const ColReduceArgs p) {
  __shared__ float sm[256 * 16];
  const int G = p.C / 8;
  const int Gp = G < 256 ? G : 256;
  const int R = 256 / Gp;
  const int tg = threadIdx.x % Gp, tr = threadIdx.x / Gp;
  const long long rpb = (p.rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rpb;
  const long long r1 = min(p.rows, r0 + rpb);
  for (int g = tg; g < G; g += Gp) {
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
    const int c = g * 8;
    float mu[8], sc[8], sh[8];
    if (MODE == 1) {
      ld8f(p.mean + c, mu);
      if (p.relu && p.y == nullptr) {
        ld8f(p.scale + c, sc);
        ld8f(p.shift + c, sh);
      }
    }
    if (tr < R) {
      // U rows per trip: all loads of a trip are issued before any is consumed (memory-level parallelism)
      constexpr int U = (MODE == 1) ? 2 : 4;
      RowWalker rw;
      rw.init(r0 + tr, MODE == 3 ? 0 : p.hp, MODE == 3 ? 0 : p.wp);
      for (long long rb = r0 + tr; rb < r1; rb += (long long)R * U) {
        float a[U][8], xv[U][8], yv[U][8];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long r = rb + (long long)u * R;
          ok[u] = r < r1 && rw.interior();
          rw.advance(R);
          if (ok[u]) {
            ld8x(p.a, r * p.lda + c, p.a_fp32, a[u]);
            if (MODE == 1 || MODE == 3) ld8x(p.x, r * p.ldx + c, p.x_fp32, xv[u]);
            if (MODE == 1 && p.relu && p.y != nullptr) ld8(reinterpret_cast<const __nv_bfloat16*>(p.y) + r * p.ldy + c, yv[u]);
            if (MODE == 3 && p.a2 != nullptr) ld8(reinterpret_cast<const __nv_bfloat16*>(p.a2) + r * p.lda2 + c, yv[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!ok[u]) continue;
          const long long r = rb + (long long)u * R;
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += a[u][i]; s1[i] += a[u][i] * a[u][i]; }
          } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s0[i] += a[u][i];
          } else if (MODE == 1) {
            if (p.relu) {
              if (p.y != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (!(yv[u][i] > 0.f)) a[u][i] = 0.f;
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (!(fmaf(xv[u][i], sc[i], sh[i]) > 0.f)) a[u][i] = 0.f;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += a[u][i]; s1[i] = fmaf(a[u][i], xv[u][i] - mu[i], s1[i]); }
          } else {
            if (p.a2 != nullptr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) a[u][i] += yv[u][i];
            }
            const float m = p.mean[r], sd = p.rstd[r];
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += a[u][i]; s1[i] += a[u][i] * (xv[u][i] - m) * sd; }
          }
        }
      }
    }
    if (MODE == 1) {  // the per-channel 1/std factor of xhat is applied once, not per element
      float rs[8];
      ld8f(p.rstd + c, rs);
#pragma unroll
      for (int i = 0; i < 8; ++i) s1[i] *= rs[i];
    }
    // reduce over the R row-lanes that share this channel group
    __syncthreads();
    if (tr < R) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { sm[threadIdx.x * 16 + i] = s0[i]; sm[threadIdx.x * 16 + 8 + i] = s1[i]; }
    }
    __syncthreads();
    if (tr == 0) {
      for (int rr = 1; rr < R; ++rr) {
        const int o = (rr * Gp + tg) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += sm[o + i]; s1[i] += sm[o + 8 + i]; }
      }
      // 64 partial rows shared by all blocks (vector fp32 atomics into a caller-zeroed buffer): the finalize
      // kernels read 64 rows instead of one row per block
      float* dst = p.partials + (size_t)(blockIdx.x & 63) * 2 * p.C;
      atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(s0[0], s0[1], s0[2], s0[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + c + 4), make_float4(s0[4], s0[5], s0[6], s0[7]));
      if (MODE != 2) {
        atomicAdd(reinterpret_cast<float4*>(dst + p.C + c), make_float4(s1[0], s1[1], s1[2], s1[3]));
        atomicAdd(reinterpret_cast<float4*>(dst + p.C + c + 4), make_float4(s1[4], s1[5], s1[6], s1[7]));
      }
    }
  }
}

// sums[i] = sum_t partials[t][i]: 32 columns x 8 tile-lanes per block, coalesced 128 B rows, fp32 tree
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n_tiles, int C2,
                                                              float* __restrict__ sums) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // 4 independent chains for ILP
  if (i < C2) {
    int t = grp;
    for (; t + 24 < n_tiles; t += 32) {
      s0 += partials[(size_t)t * C2 + i];
      s1 += partials[(size_t)(t + 8) * C2 + i];
      s2 += partials[(size_t)(t + 16) * C2 + i];
      s3 += partials[(size_t)(t + 24) * C2 + i];
    }
    for (; t < n_tiles; t += 8) s0 += partials[(size_t)t * C2 + i];
  }
  sm[grp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp == 0 && i < C2) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += sm[g][lane];
    sums[i] = s;
  }
}

__global__ void bn_coeffs_kernel(const float* __restrict__ sums, double count, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                 float* running_var, float* scale, float* shift, float* mean_out, float* invstd_out,
                                 int C, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = (double)sums[c] / count;
    double v = (double)sums[C + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    if (running_mean != nullptr) {
      const double unb = count > 1 ? v * count / (count - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float inv = rsqrtf(var + eps);
  const float sc = gamma[c] * inv;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (invstd_out) invstd_out[c] = inv;
}

// eval mode, every BatchNorm of the model in ONE launch: table entries {gamma, beta, running_mean, running_var, out, C,
// c0}; out = [scale | shift | mean | invstd] (4C floats); block -> entry by binary search over the channel prefix c0
struct BnEvalEntry {
  const float* gamma; const float* beta; const float* rm; const float* rv;
  float* out;
  int C;
  int c0;  // exclusive prefix sum of ceil(C / 128) blocks
};
__global__ void bn_coeffs_multi_kernel(const BnEvalEntry* __restrict__ tab, int n, float eps) {
  int lo = 0, hi = n - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].c0 <= b) lo = mid; else hi = mid - 1;
  }
  const BnEvalEntry e = tab[lo];
  const int c = (b - e.c0) * 128 + threadIdx.x;
  if (c >= e.C) return;
  const float mean = e.rm[c], inv = rsqrtf(e.rv[c] + eps);
  const float sc = e.gamma[c] * inv;
  e.out[c] = sc;
  e.out[e.C + c] = e.beta[c] - mean * sc;
  e.out[2 * e.C + c] = mean;
  e.out[3 * e.C + c] = inv;
}

__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ scale,
                                const float* __restrict__ shift, const __nv_bfloat16* __restrict__ resid,
                                long long ldr, __nv_bfloat16* __restrict__ y, long long ldy, long long rows, int C,
                                int relu, int hp, int wp) {
  const int G = C / 8;
  const long long total = rows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / G;
    const int c = (int)(i - r * G) * 8;
    float v[8];
    if (interior_row(r, hp, wp)) {
      float sc[8], sh[8];
      ld8(x + r * ldx + c, v);
      ld8f(scale + c, sc);
      ld8f(shift + c, sh);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
      if (resid != nullptr) {
        float rv[8];
        ld8(resid + r * ldr + c, rv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += rv[k];
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    st8(y + r * ldy + c, v);
  }
}

// dx = gamma*invstd*(dz - s0/cnt - xhat*s1/cnt), dz = dy*(y>0); optional dres (+)= dz
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                    const __nv_bfloat16* __restrict__ y, long long ldy,
                                    const __nv_bfloat16* __restrict__ x, long long ldx,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ sums, float inv_count,
                                    __nv_bfloat16* __restrict__ dx, long long lddx, __nv_bfloat16* __restrict__ dres,
                                    long long lddres, int dres_accumulate, long long rows, int C, int relu, int hp,
                                    int wp) {
  const int G = C / 8;
  const long long total = rows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / G;
    const int c = (int)(i - r * G) * 8;
    float o[8], dz[8];
    if (interior_row(r, hp, wp)) {
      float xv[8], mu[8], is[8], ga[8], s0[8], s1[8];
      ld8(dy + r * lddy + c, dz);
      ld8(x + r * ldx + c, xv);
      ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga);
      if (relu) {
        if (y != nullptr) {
          float yv[8];
          ld8(y + r * ldy + c, yv);
#pragma unroll
          for (int k = 0; k < 8; ++k) if (!(yv[k] > 0.f)) dz[k] = 0.f;
        } else {  // no residual: the forward value was fma(x, scale, shift) with scale = gamma*invstd
          float be[8];
          ld8f(beta + c, be);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float scl = ga[k] * is[k];
            if (!(fmaf(xv[k], scl, be[k] - mu[k] * scl) > 0.f)) dz[k] = 0.f;
          }
        }
      }
      ld8f(sums + c, s0); ld8f(sums + C + c, s1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mu[k]) * is[k];
        o[k] = ga[k] * is[k] * (dz[k] - s0[k] * inv_count - xh * s1[k] * inv_count);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = dz[k] = 0.f;
    }
    st8(dx + r * lddx + c, o);
    if (dres != nullptr) {
      if (dres_accumulate) {
        float old[8];
        ld8(dres + r * lddres + c, old);
#pragma unroll
        for (int k = 0; k < 8; ++k) dz[k] += old[k];
      }
      st8(dres + r * lddres + c, dz);
    }
  }
}

// ---- 32-bit / magic-division variants of the two kernels above (CRIS_B200_FASTDIV=1, work < 2^31) ------------
struct BnIndex {
  FastDiv dG, dHW, dW;
  unsigned G, hw, wp, hp;
  // flat work index -> (row, 8-channel group); returns whether the row is an interior pixel
  __device__ __forceinline__ bool decode(unsigned i, unsigned& r, unsigned& c) const {
    r = dG.div(i);
    c = (i - r * G) * 8u;
    if (wp == 0u) return true;
    const unsigned rr = r - dHW.div(r) * hw;
    const unsigned h = dW.div(rr), w = rr - h * wp;
    return (h >= 1u) && (h <= hp - 2u) && (w >= 1u) && (w <= wp - 2u);
  }
};

__global__ void bn_apply_fast_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const __nv_bfloat16* __restrict__ resid, long long ldr,
                                     __nv_bfloat16* __restrict__ y, long long ldy, unsigned total, int relu,
                                     BnIndex ix) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r, c;
    const bool in = ix.decode(i, r, c);
    float v[8];
    if (in) {
      float sc[8], sh[8];
      ld8(x + (long long)r * ldx + c, v);
      ld8f(scale + c, sc);
      ld8f(shift + c, sh);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
      if (resid != nullptr) {
        float rv[8];
        ld8(resid + (long long)r * ldr + c, rv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += rv[k];
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    st8(y + (long long)r * ldy + c, v);
  }
}

// coef = [A | B | K | shf] (4 x C floats, bn_bwd_coeffs_kernel): dx = A*dz + B*x + K, mask = fma(x, A, shf) > 0
__global__ void bn_bwd_coeffs_kernel(const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ sums, float inv_count, float* __restrict__ coef,
                                     int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float A = gamma[c] * invstd[c];
  const float B = -A * invstd[c] * sums[C + c] * inv_count;
  coef[c] = A;
  coef[C + c] = B;
  coef[2 * C + c] = -A * sums[c] * inv_count - B * mean[c];
  coef[3 * C + c] = beta[c] - mean[c] * A;
}

__global__ void bn_bwd_apply_fast_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                         const __nv_bfloat16* __restrict__ y, long long ldy,
                                         const __nv_bfloat16* __restrict__ x, long long ldx,
                                         const float* __restrict__ coef, int C, __nv_bfloat16* __restrict__ dx,
                                         long long lddx, __nv_bfloat16* __restrict__ dres, long long lddres,
                                         int dres_accumulate, unsigned total, int relu, BnIndex ix) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r, c;
    const bool in = ix.decode(i, r, c);
    float o[8], dz[8];
    if (in) {
      float xv[8], A[8], Bc[8], Kc[8];
      ld8(dy + (long long)r * lddy + c, dz);
      ld8(x + (long long)r * ldx + c, xv);
      ld8f(coef + c, A);
      ld8f(coef + C + c, Bc);
      ld8f(coef + 2 * C + c, Kc);
      if (relu) {
        if (y != nullptr) {
          float yv[8];
          ld8(y + (long long)r * ldy + c, yv);
#pragma unroll
          for (int k = 0; k < 8; ++k) if (!(yv[k] > 0.f)) dz[k] = 0.f;
        } else {
          float shf[8];
          ld8f(coef + 3 * C + c, shf);
#pragma unroll
          for (int k = 0; k < 8; ++k) if (!(fmaf(xv[k], A[k], shf[k]) > 0.f)) dz[k] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], dz[k], fmaf(Bc[k], xv[k], Kc[k]));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = dz[k] = 0.f;
    }
    st8(dx + (long long)r * lddx + c, o);
    if (dres != nullptr) {
      if (dres_accumulate) {
        float old[8];
        ld8(dres + (long long)r * lddres + c, old);
#pragma unroll
        for (int k = 0; k < 8; ++k) dz[k] += old[k];
      }
      st8(dres + (long long)r * lddres + c, dz);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C % 128 == 0, C <= 2048.  y = xhat*gamma+beta (bf16 or fp32),
// optional y2 = y + add[row % add_period] (bf16) — the "+ positional encoding" copy fed to q/k.
// ---------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
    layernorm_fwd_kernel(const void* __restrict__ x, int x_fp32, long long ldx, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ add, long long ldadd,
                         int add_period, void* __restrict__ y, int y_fp32, long long ldy,
                         __nv_bfloat16* __restrict__ y2, long long ldy2, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = C / 128;
  float v[MAXV][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      ld4x(x, row * ldx + i * 128 + lane * 4, x_fp32, v[i]);
      s += v[i][0] + v[i][1] + v[i][2] + v[i][3];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float d = v[i][k] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const int c = i * 128 + lane * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 b = *reinterpret_cast<const float4*>(beta + c);
      float o[4];
      o[0] = (v[i][0] - mean) * rstd * g.x + b.x;
      o[1] = (v[i][1] - mean) * rstd * g.y + b.y;
      o[2] = (v[i][2] - mean) * rstd * g.z + b.z;
      o[3] = (v[i][3] - mean) * rstd * g.w + b.w;
      if (y != nullptr) st4x(y, row * ldy + c, y_fp32, o);
      if (y2 != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(add + (row % add_period) * ldadd + c);
        o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
        st4x(y2, row * ldy2 + c, 0, o);
      }
    }
  }
}

// dx (+)= rstd*(g - mean(g) - xhat*mean(g*xhat)), g = (dy + dy2)*gamma; the parameter gradients
// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy ride along in registers (each lane owns fixed columns) and are
// flushed once per block: shared-memory reduction over the 8 warps, then one fp32 atomic per column.
template <int MAXV>
__global__ void __launch_bounds__(256, 2)
    layernorm_bwd_kernel(const void* __restrict__ dy, int dy_fp32, long long lddy,
                         const __nv_bfloat16* __restrict__ dy2, long long lddy2, const void* __restrict__ x,
                         int x_fp32, long long ldx, const float* __restrict__ gamma, const float* __restrict__ mean,
                         const float* __restrict__ rstd, void* __restrict__ dx, int dx_fp32, long long lddx,
                         int dx_accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                         int C) {
  __shared__ float red[2 * MAXV * 128];
  const int lane = threadIdx.x & 31;
  const int nv = C / 128;
  float dg[MAXV][4], db[MAXV][4];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) dg[i][k] = db[i][k] = 0.f;
  }
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    float g[MAXV][4], xh[MAXV][4];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (i < nv) {
        const int c = i * 128 + lane * 4;
        float d[4], xv[4];
        ld4x(dy, row * lddy + c, dy_fp32, d);
        if (dy2 != nullptr) {
          float d2[4];
          ld4x(dy2, row * lddy2 + c, 0, d2);
          d[0] += d2[0]; d[1] += d2[1]; d[2] += d2[2]; d[3] += d2[3];
        }
        ld4x(x, row * ldx + c, x_fp32, xv);
        const float4 gv = *reinterpret_cast<const float4*>(gamma + c);  // L1-resident
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xh[i][k] = (xv[k] - mu) * rs;
          g[i][k] = d[k] * ga[k];
          sg += g[i][k];
          sgx += g[i][k] * xh[i][k];
          db[i][k] += d[k];
          dg[i][k] = fmaf(d[k], xh[i][k], dg[i][k]);
        }
      }
    }
    if (dx == nullptr) continue;
    const float mg = warp_sum(sg) / C, mgx = warp_sum(sgx) / C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (i < nv) {
        const int c = i * 128 + lane * 4;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = rs * (g[i][k] - mg - xh[i][k] * mgx);
        if (dx_accumulate) {
          float old[4];
          ld4x(dx, row * lddx + c, dx_fp32, old);
          o[0] += old[0]; o[1] += old[1]; o[2] += old[2]; o[3] += old[3];
        }
        st4x(dx, row * lddx + c, dx_fp32, o);
      }
    }
  }
  if (dgamma == nullptr) return;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&red[i * 128 + lane * 4 + k], dg[i][k]);
        atomicAdd(&red[C + i * 128 + lane * 4 + k], db[i][k]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

// ---- wide rows (C > 512, C % 256 == 0): one block of C/8 threads per row, 8 contiguous columns per thread ----
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* sm) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // the previous round's readers are done with sm
  if (lane == 0) sm[w] = make_float2(a, b);
  __syncthreads();
  float2 t = lane < nw ? sm[lane] : make_float2(0.f, 0.f);
  t.x = warp_sum(t.x);
  t.y = warp_sum(t.y);
  return t;
}

__global__ void __launch_bounds__(256)
    layernorm_fwd_wide_kernel(const void* __restrict__ x, int x_fp32, long long ldx, const float* __restrict__ gamma,
                              const float* __restrict__ beta, const float* __restrict__ add, long long ldadd,
                              int add_period, void* __restrict__ y, int y_fp32, long long ldy,
                              __nv_bfloat16* __restrict__ y2, long long ldy2, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out, long long rows, int C, float eps) {
  __shared__ float2 sm[32];
  const int c = threadIdx.x * 8;
  float ga[8], be[8];
  ld8f(gamma + c, ga);
  ld8f(beta + c, be);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    float v[8];
    ld8x(x, row * ldx + c, x_fp32, v);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    const float mean = block_sum2(s, 0.f, sm).x / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum2(q, 0.f, sm).x / C + eps);
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (v[k] - mean) * rstd * ga[k] + be[k];
    if (y != nullptr) st8x(y, row * ldy + c, y_fp32, o);
    if (y2 != nullptr) {
      float a[8];
      ld8f(add + (row % add_period) * ldadd + c, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += a[k];
      st8(y2 + row * ldy2 + c, o);
    }
  }
}

__global__ void __launch_bounds__(256)
    layernorm_bwd_wide_kernel(const void* __restrict__ dy, int dy_fp32, long long lddy,
                              const __nv_bfloat16* __restrict__ dy2, long long lddy2, const void* __restrict__ x,
                              int x_fp32, long long ldx, const float* __restrict__ gamma,
                              const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ dx,
                              int dx_fp32, long long lddx, int dx_accumulate, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, long long rows, int C) {
  __shared__ float2 sm[32];
  const int c = threadIdx.x * 8;
  float ga[8], dg[8], db[8];
  ld8f(gamma + c, ga);
#pragma unroll
  for (int k = 0; k < 8; ++k) dg[k] = db[k] = 0.f;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float d[8], xh[8], g[8];
    ld8x(dy, row * lddy + c, dy_fp32, d);
    if (dy2 != nullptr) {
      float d2[8];
      ld8(dy2 + row * lddy2 + c, d2);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] += d2[k];
    }
    ld8x(x, row * ldx + c, x_fp32, xh);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xh[k] = (xh[k] - mu) * rs;
      g[k] = d[k] * ga[k];
      sg += g[k];
      sgx += g[k] * xh[k];
      db[k] += d[k];
      dg[k] = fmaf(d[k], xh[k], dg[k]);
    }
    if (dx == nullptr) continue;
    const float2 t = block_sum2(sg, sgx, sm);
    const float mg = t.x / C, mgx = t.y / C;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rs * (g[k] - mg - xh[k] * mgx);
    if (dx_accumulate) {
      float old[8];
      ld8x(dx, row * lddx + c, dx_fp32, old);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += old[k];
    }
    st8x(dx, row * lddx + c, dx_fp32, o);
  }
  if (dgamma == nullptr) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(dgamma + c + k, dg[k]);
    atomicAdd(dbeta + c + k, db[k]);
  }
}

// ---- fused small kernels (one launch instead of three) --------------------------------------------------
// forward: partials[n_tiles][2][C] -> sums -> scale/shift/mean/invstd (+ running statistics)
__global__ void __launch_bounds__(256) bn_finalize_fwd_kernel(const float* __restrict__ partials, int n_tiles, int C,
                                                              float* __restrict__ sums, double count,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float momentum, float* running_mean,
                                                              float* running_var, float* scale, float* shift,
                                                              float* mean_out, float* invstd_out) {
  __shared__ float sm[8][2][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (int t = grp; t < n_tiles; t += 8) {
      s0 += partials[(size_t)t * 2 * C + c];
      s1 += partials[(size_t)t * 2 * C + C + c];
    }
  }
  sm[grp][0][lane] = s0;
  sm[grp][1][lane] = s1;
  __syncthreads();
  if (grp == 0 && c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) { a += sm[g][0][lane]; b += sm[g][1][lane]; }
    sums[c] = a;
    sums[C + c] = b;
    const double m = (double)a / count;
    double v = (double)b / count - m * m;
    if (v < 0) v = 0;
    if (running_mean != nullptr) {
      const double unb = count > 1 ? v * count / (count - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
    const float inv = rsqrtf((float)v + eps);
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - (float)m * sc;
    mean_out[c] = (float)m;
    invstd_out[c] = inv;
  }
}
// backward: partials -> sums (for dx) and the parameter gradients dbeta = sum dz, dgamma = sum dz*xhat
__global__ void __launch_bounds__(256) stats_finalize_bwd_kernel(const float* __restrict__ partials, int n_tiles, int C,
                                                                 float* __restrict__ sums, float* __restrict__ g0,
                                                                 float* __restrict__ g1) {
  __shared__ float sm[8][2][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (int t = grp; t < n_tiles; t += 8) {
      s0 += partials[(size_t)t * 2 * C + c];
      s1 += partials[(size_t)t * 2 * C + C + c];
    }
  }
  sm[grp][0][lane] = s0;
  sm[grp][1][lane] = s1;
  __syncthreads();
  if (grp == 0 && c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) { a += sm[g][0][lane]; b += sm[g][1][lane]; }
    if (sums != nullptr) { sums[c] = a; sums[C + c] = b; }
    if (g0 != nullptr) g0[c] = a;
    if (g1 != nullptr) g1[c] = b;
  }
}

}  // namespace cris

namespace cris {  // bn_stream.cu: shared-memory-staged streaming versions of the three big BatchNorm passes
bool bn_stream_ok(long long rows, int C, int hp, int wp);
int bn_apply_stream(const void* x, long long ldx, const float* scale, const float* shift, const void* resid, long long ldr,
                    void* y, long long ldy, long long rows, int C, int relu, int hp, int wp, cudaStream_t s);
int bn_reduce_stream(int mode, const void* a0, long long lda, const void* y, long long ldy, const void* x, long long ldx,
                     const float* mean, const float* invstd, const float* scale, const float* shift, long long rows, int C,
                     int relu, int hp, int wp, float* partials, int n_part, cudaStream_t s, void* dzm_out = nullptr,
                     long long lddzm = 0);
int bn_bwd_apply_stream(const void* dy, long long lddy, const void* y, long long ldy, const void* x, long long ldx,
                        const float* mean, const float* invstd, const float* gamma, const float* beta, const float* sums,
                        double count, void* dx, long long lddx, void* dres, long long lddres, int dres_accumulate,
                        long long rows, int C, int relu, int hp, int wp, cudaStream_t s);
static bool al16(const void* p, long long ld) {
  return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0);
}
}  // namespace cris

using namespace cris;

extern "C" {

int cris_col_reduce(int mode, const void* a, int64_t lda, int a_fp32, const void* a2, int64_t lda2, const void* y,
                    int64_t ldy, const void* x, int64_t ldx, int x_fp32, const float* mean, const float* rstd,
                    const float* scale, const float* shift, int64_t rows, int C, int relu, int hp, int wp,
                    float* partials, int n_blocks, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0 && C <= 2048, "col_reduce: C=%d must be a multiple of 8 and <= 2048", C);
  CRIS_CHECK_ARG(n_blocks >= 1, "col_reduce: n_blocks");
  ColReduceArgs p{a, lda, a_fp32, a2, lda2, y, ldy, x, ldx, x_fp32, mean, rstd, scale, shift, rows, C, relu, hp, wp, partials};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if ((mode == 0 || mode == 1) && !a_fp32 && !x_fp32 && a2 == nullptr && bn_stream_ok(rows, C, hp, wp) && al16(a, lda) &&
      (mode == 0 || (x != nullptr && al16(x, ldx) && al16(y, ldy))))
    return bn_reduce_stream(mode, a, lda, y, ldy, x, ldx, mean, rstd, scale, shift, rows, C, relu, hp, wp, partials,
                            n_blocks < 64 ? n_blocks : 64, s);
  switch (mode) {
    case 0: col_reduce_kernel<0><<<n_blocks, 256, 0, s>>>(p); break;
    case 1: col_reduce_kernel<1><<<n_blocks, 256, 0, s>>>(p); break;
    case 2: col_reduce_kernel<2><<<n_blocks, 256, 0, s>>>(p); break;
    case 3: col_reduce_kernel<3><<<n_blocks, 256, 0, s>>>(p); break;
    default: set_error("col_reduce: bad mode %d", mode); return -1;
  }
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_bn_bwd_reduce_masked(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x, int64_t ldx,
                              const float* mean, const float* invstd, int64_t rows, int C, int hp, int wp, void* dzm,
                              int64_t lddzm, float* partials, int n_blocks, void* stream) {
  CRIS_CHECK_ARG(dy && y && x && mean && invstd && dzm && partials && n_blocks >= 1, "bn_bwd_reduce_masked: null argument");
  CRIS_CHECK_ARG(bn_stream_ok(rows, C, hp, wp) && al16(dy, lddy) && al16(y, ldy) && al16(x, ldx) && al16(dzm, lddzm),
                 "bn_bwd_reduce_masked: rows=%lld C=%d outside the streaming kernel's domain (power-of-two C in [8, 2048], "
                 "rows >= 2048, 16-byte aligned operands): use cris_col_reduce + cris_bn_bwd_apply", (long long)rows, C);
  return bn_reduce_stream(1, dy, lddy, y, ldy, x, ldx, mean, invstd, nullptr, nullptr, rows, C, 1, hp, wp, partials,
                          n_blocks < 64 ? n_blocks : 64, reinterpret_cast<cudaStream_t>(stream), dzm, lddzm);
}

int cris_bn_reduce_partials(const float* partials, int n_tiles, int C, float* sums, void* stream) {
  reduce_partials_kernel<<<(2 * C + 31) / 32, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partials, n_tiles,
                                                                                               2 * C, sums);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_bn_finalize_fwd(const float* partials, int n_tiles, int C, float* sums, double count, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                         float* scale, float* shift, float* mean, float* invstd, void* stream) {
  bn_finalize_fwd_kernel<<<(C + 31) / 32, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partials, n_tiles, C, sums, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_stats_finalize_bwd(const float* partials, int n_tiles, int C, float* sums, float* g0, float* g1, void* stream) {
  stats_finalize_bwd_kernel<<<(C + 31) / 32, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partials, n_tiles, C, sums,
                                                                                              g0, g1);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_bn_eval_entry_bytes(void) { return (int)sizeof(BnEvalEntry); }
int cris_bn_coeffs_multi(const void* table_dev, int n_entries, int n_blocks, float eps, void* stream) {
  CRIS_CHECK_ARG(table_dev != nullptr && n_entries >= 1 && n_blocks >= 1, "cris_bn_coeffs_multi: bad table");
  bn_coeffs_multi_kernel<<<n_blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const BnEvalEntry*>(table_dev), n_entries, eps);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_bn_coeffs(const float* sums, double count, const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                   int C, int training, void* stream) {
  bn_coeffs_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sums, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd, C, training);
  CRIS_LAUNCH_OK();
  return 0;
}

static bool fastdiv_enabled() {
  const char* e = getenv("CRIS_B200_FASTDIV");  // read per call: tests flip it; default off until measured
  return e != nullptr && e[0] == '1';
}

static BnIndex make_bn_index(int C, int hp, int wp) {
  BnIndex ix;
  ix.G = (unsigned)(C / 8);
  ix.hp = (unsigned)(hp > 0 ? hp : 0);
  ix.wp = (unsigned)(wp > 0 ? wp : 0);
  ix.hw = ix.hp * ix.wp;
  ix.dG = FastDiv(ix.G);
  ix.dHW = FastDiv(ix.hw ? ix.hw : 1);
  ix.dW = FastDiv(ix.wp ? ix.wp : 1);
  return ix;
}

int cris_bn_apply(const void* x, int64_t ldx, const float* scale, const float* shift, const void* resid, int64_t ldr,
                  void* y, int64_t ldy, int64_t rows, int C, int relu, int hp, int wp, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "bn_apply: C=%d must be a multiple of 8", C);
  if (bn_stream_ok(rows, C, hp, wp) && al16(x, ldx) && al16(resid, ldr) && al16(y, ldy))
    return bn_apply_stream(x, ldx, scale, shift, resid, ldr, y, ldy, rows, C, relu, hp, wp,
                           reinterpret_cast<cudaStream_t>(stream));
  const long long work = rows * (C / 8);
  if (fastdiv_enabled() && work < (1ll << 31) && rows < (1ll << 31)) {
    bn_apply_fast_kernel<<<grid_for(work, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, scale, shift, reinterpret_cast<const __nv_bfloat16*>(resid), ldr,
        reinterpret_cast<__nv_bfloat16*>(y), ldy, (unsigned)work, relu, make_bn_index(C, hp, wp));
    CRIS_LAUNCH_OK();
    return 0;
  }
  bn_apply_kernel<<<grid_for(work, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, scale, shift, reinterpret_cast<const __nv_bfloat16*>(resid), ldr,
      reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, C, relu, hp, wp);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_bn_bwd_apply(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* x, int64_t ldx,
                      const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const float* sums, double count, void* dx, int64_t lddx, void* dres, int64_t lddres,
                      int dres_accumulate, int64_t rows, int C, int relu, int hp, int wp, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "bn_bwd_apply: C=%d must be a multiple of 8", C);
  const long long work = rows * (C / 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (bn_stream_ok(rows, C, hp, wp) && al16(dy, lddy) && al16(y, ldy) && al16(x, ldx) && al16(dx, lddx) &&
      al16(dres, lddres))
    return bn_bwd_apply_stream(dy, lddy, y, ldy, x, ldx, mean, invstd, gamma, beta, sums, count, dx, lddx, dres, lddres,
                               dres_accumulate, rows, C, relu, hp, wp, s);
  if (fastdiv_enabled() && work < (1ll << 31) && rows < (1ll << 31)) {
    // per-channel coefficients once (a [4, C] scratch vector owned by the library, stream-ordered reuse)
    static float* coef = nullptr;
    static int coef_cap = 0;
    static int coef_dev = -1;
    int dev = 0;
    CRIS_CUDA_OK(cudaGetDevice(&dev));
    if (coef == nullptr || coef_cap < C || coef_dev != dev) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &st);
      CRIS_CHECK_ARG(st == cudaStreamCaptureStatusNone, "bn_bwd_apply: coefficient scratch must be sized before capture");
      if (coef != nullptr && coef_dev == dev) cudaFree(coef);
      coef_cap = C > 4096 ? C : 4096;
      CRIS_CUDA_OK(cudaMalloc(&coef, (size_t)4 * coef_cap * sizeof(float)));
      coef_dev = dev;
    }
    bn_bwd_coeffs_kernel<<<(C + 127) / 128, 128, 0, s>>>(mean, invstd, gamma, beta, sums, (float)(1.0 / count), coef, C);
    CRIS_LAUNCH_OK();
    bn_bwd_apply_fast_kernel<<<grid_for(work, 256), 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), ldy,
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, coef, C, reinterpret_cast<__nv_bfloat16*>(dx), lddx,
        reinterpret_cast<__nv_bfloat16*>(dres), lddres, dres_accumulate, (unsigned)work, relu, make_bn_index(C, hp, wp));
    CRIS_LAUNCH_OK();
    return 0;
  }
  bn_bwd_apply_kernel<<<grid_for(work, 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), ldy,
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, mean, invstd, gamma, beta, sums, (float)(1.0 / count),
      reinterpret_cast<__nv_bfloat16*>(dx), lddx, reinterpret_cast<__nv_bfloat16*>(dres), lddres, dres_accumulate, rows, C,
      relu, hp, wp);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_layernorm_fwd(const void* x, int x_fp32, int64_t ldx, const float* gamma, const float* beta, const float* add,
                       int64_t ldadd, int add_period, void* y, int y_fp32, int64_t ldy, void* y2, int64_t ldy2,
                       float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
  CRIS_CHECK_ARG(C % 128 == 0 && C <= 2048, "layernorm: C=%d must be a multiple of 128 and <= 2048", C);
  CRIS_CHECK_ARG(C <= 512 || C % 256 == 0, "layernorm: C=%d > 512 must be a multiple of 256", C);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int period = add_period > 0 ? add_period : 1;
  if (C <= 512)
    layernorm_fwd_kernel<4><<<(int)((rows + 7) / 8), 256, 0, s>>>(x, x_fp32, ldx, gamma, beta, add, ldadd, period, y,
                                                                  y_fp32, ldy, reinterpret_cast<__nv_bfloat16*>(y2),
                                                                  ldy2, mean, rstd, rows, C, eps);
  else
    layernorm_fwd_wide_kernel<<<(int)std::min<int64_t>(rows, 148 * 8), C / 8, 0, s>>>(
        x, x_fp32, ldx, gamma, beta, add, ldadd, period, y, y_fp32, ldy, reinterpret_cast<__nv_bfloat16*>(y2), ldy2,
        mean, rstd, rows, C, eps);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_layernorm_bwd(const void* dy, int dy_fp32, int64_t lddy, const void* dy2, int64_t lddy2, const void* x,
                       int x_fp32, int64_t ldx, const float* gamma, const float* mean, const float* rstd, void* dx,
                       int dx_fp32, int64_t lddx, int dx_accumulate, float* dgamma, float* dbeta, int64_t rows, int C,
                       void* stream) {
  CRIS_CHECK_ARG(C % 128 == 0 && C <= 2048, "layernorm: C=%d must be a multiple of 128 and <= 2048", C);
  CRIS_CHECK_ARG(C <= 512 || C % 256 == 0, "layernorm: C=%d > 512 must be a multiple of 256", C);
  CRIS_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  CRIS_CHECK_ARG(dx != nullptr || dgamma != nullptr, "layernorm_bwd: nothing to compute");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (C <= 512)
    layernorm_bwd_kernel<4><<<(int)std::min<int64_t>((rows + 7) / 8, 148 * 2), 256, 0, s>>>(
        dy, dy_fp32, lddy, reinterpret_cast<const __nv_bfloat16*>(dy2), lddy2, x, x_fp32, ldx, gamma, mean, rstd, dx,
        dx_fp32, lddx, dx_accumulate, dgamma, dbeta, rows, C);
  else
    layernorm_bwd_wide_kernel<<<(int)std::min<int64_t>(rows, 148 * 8), C / 8, 0, s>>>(
        dy, dy_fp32, lddy, reinterpret_cast<const __nv_bfloat16*>(dy2), lddy2, x, x_fp32, ldx, gamma, mean, rstd, dx,
        dx_fp32, lddx, dx_accumulate, dgamma, dbeta, rows, C);
  CRIS_LAUNCH_OK();
  return 0;
}
}
