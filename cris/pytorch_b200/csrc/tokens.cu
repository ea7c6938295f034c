// tokens.cu — row softmax (attention probabilities) fwd/bwd, token embedding, EOT gather, residual
// add + dropout, QuickGELU, casts / packing, small reductions.
// Reference call sites: softmax/dropout inside F.multi_head_attention_forward (model/clip.py:119-139,
// 255-260; model/layers.py:235,240-243), nn.Embedding + positional add (clip.py:440-443), EOT gather
// (clip.py:451-452), nn.Dropout + residual adds (layers.py:237,245,249), QuickGELU (clip.py:234-236).
#include "vec.cuh"

namespace cris {

// ---- softmax over rows of S[nb][Lq][ld] (scores already scaled by the GEMM alpha) ----------------
// One warp per row, 16-byte vector accesses: lane handles 8 consecutive keys per vector, NV vectors per lane
// (Lk <= 256*NV; ld % 8 == 0).  causal: key j > query i masked.  key padding: token id word[b][j] == 0 masked
// (pad_mask of model/segmenter.py:37), b = batch_index / heads.  P = softmax; Pd (optional) = dropout(P)/(1-p).
template <int NV>
__global__ void __launch_bounds__(256)
    softmax_fwd_kernel(const __nv_bfloat16* __restrict__ S, __nv_bfloat16* __restrict__ P,
                       __nv_bfloat16* __restrict__ Pd, long long ld, long long batch_stride, int nb, int Lq, int Lk,
                       int heads, const long long* __restrict__ kpm, int causal, float p_drop, uint64_t seed,
                       const uint64_t* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const int lane = threadIdx.x & 31;
  const long long rid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (rid >= (long long)nb * Lq) return;
  const int bi = (int)(rid / Lq), qi = (int)(rid % Lq);
  const long long off = (long long)bi * batch_stride + (long long)qi * ld;
  const long long* km = kpm ? kpm + (long long)(bi / heads) * Lk : nullptr;  // token ids; id 0 = padding
  float v[NV][8];
  float mx = -INFINITY;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int j0 = (e * 32 + lane) * 8;
    if (j0 < ld) ld8(S + off + j0, v[e]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + i;
      if (j >= Lk || (causal && j > qi) || (km && km[j] == 0)) v[e][i] = -INFINITY;
      mx = fmaxf(mx, v[e][i]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < NV; ++e)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[e][i] = (v[e][i] == -INFINITY) ? 0.f : __expf(v[e][i] - mx);
      sum += v[e][i];
    }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  const uint32_t th = drop_thresh16(p_drop);
  const uint32_t rh = drop_row_hash(seed, (uint64_t)rid);  // same (row, column) masks as csrc/attention.cu
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int j0 = (e * 32 + lane) * 8;
    if (j0 < ld) {
      float pr[8], pd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pr[i] = v[e][i] * inv;
        pd[i] = (Pd != nullptr && drop_keep_rc(rh, (uint32_t)(j0 + i), th)) ? pr[i] * keep_scale : 0.f;
      }
      st8(P + off + j0, pr);
      if (Pd != nullptr) st8(Pd + off + j0, pd);
    }
  }
}

// dS = P * (dPm - sum_j P*dPm), dPm = dP * dropmask/(1-p);  written in place of dP (bf16)
template <int NV>
__global__ void __launch_bounds__(256)
    softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ dP, long long ld,
                       long long batch_stride, int nb, int Lq, int Lk, float p_drop, uint64_t seed,
                       const uint64_t* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const int lane = threadIdx.x & 31;
  const long long rid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (rid >= (long long)nb * Lq) return;
  const int bi = (int)(rid / Lq), qi = (int)(rid % Lq);
  const long long off = (long long)bi * batch_stride + (long long)qi * ld;
  const uint32_t th = drop_thresh16(p_drop);
  const uint32_t rh = drop_row_hash(seed, (uint64_t)rid);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float pv[NV][8], dv[NV][8];
  float dot = 0.f;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int j0 = (e * 32 + lane) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[e][i] = dv[e][i] = 0.f;
    if (j0 < ld) {
      ld8(P + off + j0, pv[e]);
      ld8(dP + off + j0, dv[e]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (j0 + i >= Lk) { pv[e][i] = 0.f; dv[e][i] = 0.f; }
        else if (p_drop > 0.f) dv[e][i] = drop_keep_rc(rh, (uint32_t)(j0 + i), th) ? dv[e][i] * keep_scale : 0.f;
        dot += pv[e][i] * dv[e][i];
      }
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int j0 = (e * 32 + lane) * 8;
    if (j0 < ld) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pv[e][i] * (dv[e][i] - dot);
      st8(dP + off + j0, o);
    }
  }
}

// ---- token embedding + positional embedding -> fp32 residual stream ----------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ word, const float* __restrict__ table,
                                 const float* __restrict__ pos, float* __restrict__ x, int B, int L, int C) {
  const int G = C / 4;
  const long long total = (long long)B * L * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long t = i / G;
    const int l = (int)(t % L);
    const long long tokid = word[t];
    const float4 a = *reinterpret_cast<const float4*>(table + tokid * C + g * 4);
    const float4 b = *reinterpret_cast<const float4*>(pos + (long long)l * C + g * 4);
    *reinterpret_cast<float4*>(x + t * C + g * 4) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
// dtable[word] += dx (atomic), dpos[l] += sum_b dx
__global__ void embed_bwd_kernel(const long long* __restrict__ word, const float* __restrict__ dx,
                                 float* __restrict__ dtable, float* __restrict__ dpos, int B, int L, int C) {
  const long long total = (long long)B * L * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int l = (int)(t % L);
    const float g = dx[i];
    atomicAdd(dtable + word[t] * C + c, g);
    atomicAdd(dpos + (long long)l * C + c, g);
  }
}

// ---- EOT gather: out[b, :] = x[b*L + argmax_l word[b, l], :] (first maximum, like torch.argmax) ----
__global__ void eot_gather_kernel(const long long* __restrict__ word, const void* __restrict__ x, int x_fp32,
                                  long long ldx, __nv_bfloat16* __restrict__ out, long long ldo, int L, int C) {
  const int b = blockIdx.x;
  int best = 0;
  long long bv = word[(long long)b * L];
  for (int l = 1; l < L; ++l) {
    const long long v = word[(long long)b * L + l];
    if (v > bv) { bv = v; best = l; }
  }
  const long long row = (long long)b * L + best;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = x_fp32 ? reinterpret_cast<const float*>(x)[row * ldx + c]
                           : bf2f(reinterpret_cast<const __nv_bfloat16*>(x)[row * ldx + c]);
    out[(long long)b * ldo + c] = f2bf(v);
  }
}
// dx[b*L + eot, :] += dout[b, :]
__global__ void eot_scatter_kernel(const long long* __restrict__ word, const void* __restrict__ dout, int d_fp32,
                                   long long ldd, void* __restrict__ dx, int dx_fp32, long long lddx, int L, int C) {
  const int b = blockIdx.x;
  int best = 0;
  long long bv = word[(long long)b * L];
  for (int l = 1; l < L; ++l) {
    const long long v = word[(long long)b * L + l];
    if (v > bv) { bv = v; best = l; }
  }
  const long long row = (long long)b * L + best;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = d_fp32 ? reinterpret_cast<const float*>(dout)[(long long)b * ldd + c]
                           : bf2f(reinterpret_cast<const __nv_bfloat16*>(dout)[(long long)b * ldd + c]);
    if (dx_fp32) reinterpret_cast<float*>(dx)[row * lddx + c] += g;
    else {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(dx) + row * lddx + c;
      *p = f2bf(bf2f(*p) + g);
    }
  }
}

// ---- generic elementwise over a [rows, C] matrix (C % 4 == 0) ---------------------------------------
//   OP 0: out = a (+ b)                                     (cast / add; any dtype mix)
//   OP 1: out = a + dropout(b)                              residual add with dropout (mask from seed)
//   OP 2: out = dropout(a)                                  (backward of OP 1 w.r.t. b; also plain dropout)
//   OP 3: out = quickgelu(a)
//   OP 4: out = b * quickgelu'(a)                           (a = pre-activation, b = upstream grad)
//   OP 5: out = (a > 0) ? b : 0                             relu backward (a = activation output)
struct EwArgs {
  const void* a; int a_fp32; long long lda;
  const void* b; int b_fp32; long long ldb;
  void* out; int out_fp32; long long ldo;
  long long rows; int C;
  float p_drop; uint64_t seed; const uint64_t* seed_dev;
};
template <int OP>
__global__ void ew_kernel(const EwArgs p) {
  const int G = p.C / 4;
  const long long total = p.rows * G;
  const uint32_t th = drop_thresh(p.p_drop);
  const float ks = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  const uint64_t seed = p.seed + (p.seed_dev != nullptr ? *p.seed_dev : 0ull);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / G;
    const int c = (int)(i - r * G) * 4;
    float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, o[4];
    if (p.a != nullptr) ld4x(p.a, r * p.lda + c, p.a_fp32, a);
    if (p.b != nullptr) ld4x(p.b, r * p.ldb + c, p.b_fp32, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t idx = (uint64_t)(r * p.C + c + k);
      if (OP == 0) o[k] = a[k] + b[k];
      else if (OP == 1) o[k] = a[k] + ((p.p_drop > 0.f) ? (drop_keep(seed, idx, th) ? b[k] * ks : 0.f) : b[k]);
      else if (OP == 2) o[k] = (p.p_drop > 0.f) ? (drop_keep(seed, idx, th) ? a[k] * ks : 0.f) : a[k];
      else if (OP == 3) o[k] = a[k] / (1.f + __expf(-1.702f * a[k]));
      else if (OP == 4) {
        const float s = 1.f / (1.f + __expf(-1.702f * a[k]));
        o[k] = b[k] * (s + 1.702f * a[k] * s * (1.f - s));
      } else o[k] = a[k] > 0.f ? b[k] : 0.f;
    }
    st4x(p.out, r * p.ldo + c, p.out_fp32, o);
  }
}

// ---- weight packing: fp32 OIHW [Cout][Cin][taps] -> bf16 [Cout][taps][cin_pad] (zero padded) ----------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                        int Cin, int taps, int cin_pad, const float* __restrict__ row_scale = nullptr) {
  const long long total = (long long)Cout * taps * cin_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int t = (int)((i / cin_pad) % taps);
    const int co = (int)(i / ((long long)cin_pad * taps));
    const float sc = row_scale != nullptr ? row_scale[co] : 1.f;
    out[i] = f2bf(ci < Cin ? w[((long long)co * Cin + ci) * taps + t] * sc : 0.f);
  }
}
// weight gradient of a k x k conv from the wgrad GEMM's tap-blocked layout back to the reference's OIHW:
// gw[co][ci][t] = acc[co][t][ci] (fp32; acc = [Cout][taps][cin_pad], the layout the TMA reductions of the wgrad
// GEMM write with unit stride; model/clip.py:17-25 autograd of nn.Conv2d).  Reads are coalesced along ci.
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ acc, float* __restrict__ gw, int Cout, int Cin,
                                         int taps, int cin_pad) {
  const long long total = (long long)Cout * taps * cin_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int t = (int)((i / cin_pad) % taps);
    const int co = (int)(i / ((long long)cin_pad * taps));
    if (ci < Cin) gw[((long long)co * Cin + ci) * taps + t] = acc[i];
  }
}
// data-gradient weights of a 3x3 conv as a FORWARD conv operand: out[ci][t'][co] = w[co][ci][8 - t'] (taps mirrored,
// channels transposed), bf16 [Cin][9][cout_pad] zero padded: dx = conv3x3(dz, out) (experimental halo path)
__global__ void pack_conv_weight_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                              int Cin, int cout_pad) {
  const long long total = (long long)Cin * 9 * cout_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    const int t = (int)((i / cout_pad) % 9);
    const int ci = (int)(i / ((long long)cout_pad * 9));
    out[i] = f2bf(co < Cout ? w[((long long)co * Cin + ci) * 9 + (8 - t)] : 0.f);
  }
}
// fp32 [rows][cols] -> bf16 [rows][ld] (zero padded), optional per-row scale
__global__ void pack_matrix_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long rows,
                                   int cols, int ld, const float* __restrict__ row_scale = nullptr) {
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const long long r = i / ld;
    out[i] = f2bf(c < cols ? w[r * cols + c] * (row_scale != nullptr ? row_scale[r] : 1.f) : 0.f);
  }
}
// Every bf16 kernel-layout weight copy of the model refreshed by ONE launch (was one pack launch per parameter, 143 per
// step): a device table of entries, block -> entry by binary search over chunk prefix sums (the multi-tensor Adam
// kernel's scheme).  taps == 1: matrix [rows][cols] -> [rows][ld]; taps > 1: conv OIHW -> [Cout][taps][ld = cin_pad].
struct PackEntry {
  const float* src;
  __nv_bfloat16* dst;
  long long rows;     // output rows (Cout)
  int cols;           // matrix: source columns; conv: Cin
  int ld;             // destination pitch (matrix) / cin_pad (conv)
  int taps;
  int pad_;
  long long chunk0;   // exclusive prefix sum of ceil(dst elements / kPackChunk)
  const float* row_scale;  // optional per-output-row factor (eval-mode BatchNorm folded into the convolution)
};
constexpr int kPackChunk = 8192;

__global__ void __launch_bounds__(256) pack_multi_kernel(const PackEntry* __restrict__ tab, int n) {
  int lo = 0, hi = n - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk0 <= b) lo = mid; else hi = mid - 1;
  }
  const PackEntry e = tab[lo];
  const long long total = e.rows * (long long)e.taps * e.ld;
  const long long i0 = (b - e.chunk0) * kPackChunk, i1 = min(total, i0 + kPackChunk);
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int c = (int)(i % e.ld);
    float v = 0.f;
    long long r;
    if (e.taps == 1) {
      r = i / e.ld;
      if (c < e.cols) v = e.src[r * e.cols + c];
    } else {
      const int t = (int)((i / e.ld) % e.taps);
      r = i / ((long long)e.ld * e.taps);
      if (c < e.cols) v = e.src[(r * e.cols + c) * e.taps + t];
    }
    if (e.row_scale != nullptr) v *= e.row_scale[r];
    e.dst[i] = f2bf(v);
  }
}
// out[t, c] (+)= sum_b in[b*T + t, c]   (batch reduction of token gradients -> shared positional term)
__global__ void batch_reduce_kernel(const void* __restrict__ in, int in_fp32, long long ldin, float* __restrict__ out,
                                    long long ldo, int B, int T, int C, int accumulate) {
  const long long total = (long long)T * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)(i / C);
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long idx = ((long long)b * T + t) * ldin + c;
      s += in_fp32 ? reinterpret_cast<const float*>(in)[idx] : bf2f(reinterpret_cast<const __nv_bfloat16*>(in)[idx]);
    }
    if (accumulate) out[(long long)t * ldo + c] += s;
    else out[(long long)t * ldo + c] = s;
  }
}
// small dense fp32 matmul (batch-independent glue only: bicubic positional-embedding resize,
// 169x49 by 49xC): out[m, c] = sum_k R[m, k] * X[k, c]   or with transpose_r: out[k, c] = sum_m R[m,k] * X[m,c]
__global__ void small_matmul_kernel(const float* __restrict__ R, const float* __restrict__ X, float* __restrict__ out,
                                    int M, int K, int C, int transpose_r, int accumulate) {
  const int rows_out = transpose_r ? K : M;
  const int red = transpose_r ? M : K;
  const long long total = (long long)rows_out * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int ro = (int)(i / C);
    float s = 0.f;
    for (int j = 0; j < red; ++j) {
      const float rv = transpose_r ? R[(long long)j * K + ro] : R[(long long)ro * K + j];
      s = fmaf(rv, X[(long long)j * C + c], s);
    }
    if (accumulate) out[i] += s;
    else out[i] = s;
  }
}

}  // namespace cris

using namespace cris;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int cris_softmax_fwd(const void* S, void* P, void* Pd, int64_t ld, int64_t batch_stride, int nb, int Lq, int Lk,
                     int heads, const int64_t* kpm_word, int causal, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* stream) {
  const long long* kpm = reinterpret_cast<const long long*>(kpm_word);
  CRIS_CHECK_ARG(Lk >= 1 && Lk <= 768 && ld % 8 == 0 && ld >= Lk && batch_stride % 8 == 0,
                 "softmax: Lk=%d ld=%lld unsupported (Lk <= 768, ld %% 8 == 0)", Lk, (long long)ld);
  const int grid = (int)(((long long)nb * Lq + 7) / 8);
  if (Lk <= 256)
    softmax_fwd_kernel<1><<<grid, 256, 0, STREAM>>>(CBF(S), BF(P), BF(Pd), ld, batch_stride, nb, Lq, Lk, heads, kpm,
                                                     causal, p_drop, seed, seed_dev);
  else
    softmax_fwd_kernel<3><<<grid, 256, 0, STREAM>>>(CBF(S), BF(P), BF(Pd), ld, batch_stride, nb, Lq, Lk, heads, kpm,
                                                     causal, p_drop, seed, seed_dev);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_softmax_bwd(const void* P, void* dP, int64_t ld, int64_t batch_stride, int nb, int Lq, int Lk, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* stream) {
  CRIS_CHECK_ARG(Lk >= 1 && Lk <= 768 && ld % 8 == 0 && ld >= Lk && batch_stride % 8 == 0,
                 "softmax: Lk=%d ld=%lld unsupported (Lk <= 768, ld %% 8 == 0)", Lk, (long long)ld);
  const int grid = (int)(((long long)nb * Lq + 7) / 8);
  if (Lk <= 256)
    softmax_bwd_kernel<1><<<grid, 256, 0, STREAM>>>(CBF(P), BF(dP), ld, batch_stride, nb, Lq, Lk, p_drop, seed, seed_dev);
  else
    softmax_bwd_kernel<3><<<grid, 256, 0, STREAM>>>(CBF(P), BF(dP), ld, batch_stride, nb, Lq, Lk, p_drop, seed, seed_dev);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_embed_fwd(const int64_t* word, const float* table, const float* pos, float* x, int B, int L, int C,
                   void* stream) {
  CRIS_CHECK_ARG(C % 4 == 0, "embed: C=%d", C);
  embed_fwd_kernel<<<grid_for((long long)B * L * (C / 4), 256), 256, 0, STREAM>>>(
      reinterpret_cast<const long long*>(word), table, pos, x, B, L, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_embed_bwd(const int64_t* word, const float* dx, float* dtable, float* dpos, int B, int L, int C,
                   void* stream) {
  embed_bwd_kernel<<<grid_for((long long)B * L * C, 256), 256, 0, STREAM>>>(reinterpret_cast<const long long*>(word),
                                                                           dx, dtable, dpos, B, L, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_eot_gather(const int64_t* word, const void* x, int x_fp32, int64_t ldx, void* out, int64_t ldo, int B, int L,
                    int C, void* stream) {
  eot_gather_kernel<<<B, 128, 0, STREAM>>>(reinterpret_cast<const long long*>(word), x, x_fp32, ldx, BF(out), ldo, L,
                                           C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_eot_scatter(const int64_t* word, const void* dout, int d_fp32, int64_t ldd, void* dx, int dx_fp32,
                     int64_t lddx, int B, int L, int C, void* stream) {
  eot_scatter_kernel<<<B, 128, 0, STREAM>>>(reinterpret_cast<const long long*>(word), dout, d_fp32, ldd, dx, dx_fp32,
                                            lddx, L, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_elementwise(int op, const void* a, int a_fp32, int64_t lda, const void* b, int b_fp32, int64_t ldb, void* out,
                     int out_fp32, int64_t ldo, int64_t rows, int C, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* stream) {
  CRIS_CHECK_ARG(C % 4 == 0, "elementwise: C=%d must be a multiple of 4", C);
  EwArgs p{a, a_fp32, lda, b, b_fp32, ldb, out, out_fp32, ldo, rows, C, p_drop, seed, seed_dev};
  const int grid = grid_for(rows * (C / 4), 256);
  switch (op) {
    case 0: ew_kernel<0><<<grid, 256, 0, STREAM>>>(p); break;
    case 1: ew_kernel<1><<<grid, 256, 0, STREAM>>>(p); break;
    case 2: ew_kernel<2><<<grid, 256, 0, STREAM>>>(p); break;
    case 3: ew_kernel<3><<<grid, 256, 0, STREAM>>>(p); break;
    case 4: ew_kernel<4><<<grid, 256, 0, STREAM>>>(p); break;
    case 5: ew_kernel<5><<<grid, 256, 0, STREAM>>>(p); break;
    default: set_error("elementwise: bad op %d", op); return -1;
  }
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_conv_weight(const float* w, void* out, int Cout, int Cin, int taps, int cin_pad, void* stream) {
  pack_conv_weight_kernel<<<grid_for((long long)Cout * taps * cin_pad, 256), 256, 0, STREAM>>>(w, BF(out), Cout, Cin,
                                                                                             taps, cin_pad);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_entry_bytes(void) { return (int)sizeof(PackEntry); }
int cris_pack_chunk_elems(void) { return kPackChunk; }
int cris_pack_multi(const void* table_dev, int n_entries, long long n_chunks, void* stream) {
  CRIS_CHECK_ARG(table_dev != nullptr && n_entries >= 1 && n_chunks >= 1 && n_chunks < (1ll << 31),
                 "cris_pack_multi: bad table (%d entries, %lld chunks)", n_entries, n_chunks);
  pack_multi_kernel<<<(unsigned)n_chunks, 256, 0, STREAM>>>(static_cast<const PackEntry*>(table_dev), n_entries);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_conv_weight_scaled(const float* w, const float* row_scale, void* out, int Cout, int Cin, int taps,
                                 int cin_pad, void* stream) {
  CRIS_CHECK_ARG(w && row_scale && out, "pack_conv_weight_scaled: null argument");
  pack_conv_weight_kernel<<<grid_for((long long)Cout * taps * cin_pad, 256), 256, 0, STREAM>>>(w, BF(out), Cout, Cin,
                                                                                             taps, cin_pad, row_scale);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_matrix_scaled(const float* w, const float* row_scale, void* out, int64_t rows, int cols, int ld,
                            void* stream) {
  CRIS_CHECK_ARG(w && row_scale && out, "pack_matrix_scaled: null argument");
  pack_matrix_kernel<<<grid_for(rows * ld, 256), 256, 0, STREAM>>>(w, BF(out), rows, cols, ld, row_scale);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_unpack_conv_wgrad(const float* acc, float* gw, int Cout, int Cin, int taps, int cin_pad, void* stream) {
  CRIS_CHECK_ARG(acc && gw && cin_pad >= Cin && taps >= 1, "unpack_conv_wgrad: bad argument");
  unpack_conv_wgrad_kernel<<<grid_for((long long)Cout * taps * cin_pad, 256), 256, 0, STREAM>>>(acc, gw, Cout, Cin, taps,
                                                                                              cin_pad);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_conv_weight_dgrad(const float* w, void* out, int Cout, int Cin, int cout_pad, void* stream) {
  CRIS_CHECK_ARG(w && out && cout_pad >= Cout, "pack_conv_weight_dgrad: bad argument");
  pack_conv_weight_dgrad_kernel<<<grid_for((long long)Cin * 9 * cout_pad, 256), 256, 0, STREAM>>>(w, BF(out), Cout, Cin,
                                                                                               cout_pad);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_pack_matrix(const float* w, void* out, int64_t rows, int cols, int ld, void* stream) {
  pack_matrix_kernel<<<grid_for(rows * ld, 256), 256, 0, STREAM>>>(w, BF(out), rows, cols, ld);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_batch_reduce(const void* in, int in_fp32, int64_t ldin, float* out, int64_t ldo, int B, int T, int C,
                      int accumulate, void* stream) {
  batch_reduce_kernel<<<grid_for((long long)T * C, 256), 256, 0, STREAM>>>(in, in_fp32, ldin, out, ldo, B, T, C,
                                                                          accumulate);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_small_matmul(const float* R, const float* X, float* out, int M, int K, int C, int transpose_r, int accumulate,
                      void* stream) {
  const int rows_out = transpose_r ? K : M;
  small_matmul_kernel<<<grid_for((long long)rows_out * C, 256), 256, 0, STREAM>>>(R, X, out, M, K, C, transpose_r,
                                                                                 accumulate);
  CRIS_LAUNCH_OK();
  return 0;
}
}
