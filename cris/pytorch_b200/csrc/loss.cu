// loss.cu — the text-to-pixel head: per-sample dynamic 3x3 correlation of the projected visual
// feature with the text-generated kernel, fused with the nearest-neighbour mask resize and
// BCE-with-logits (mean) — forward and backward.
// Reference: Projector.forward grouped F.conv2d (model/layers.py:71-84) + F.interpolate(mask,'nearest')
// + F.binary_cross_entropy_with_logits (model/segmenter.py:56-59).  HBM-bound: the [B,H+2,W+2,C] feature
// is read once (neighbouring taps hit L1/L2); one warp reduces one output pixel.
#include "vec.cuh"

namespace cris {

// t[b, c*9 + tap] (fp32, pitch ldt) -> smem k[(tap*8 + j)*G + g] for channel c = 8g + j (G = C/8): lane g of a
// warp reads consecutive words (no bank conflicts); bias = t[b, C*9]
__device__ __forceinline__ void stage_kernel(const float* __restrict__ t, long long ldt, int b, int C, float* sk) {
  const int G = C / 8;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int c = i / 9, tap = i - c * 9;
    sk[(tap * 8 + (c & 7)) * G + (c >> 3)] = t[(long long)b * ldt + i];
  }
}

constexpr int kDynPix = 128;  // consecutive output pixels per block (their 3x3 windows overlap in L1)

// One warp per output pixel, lane = channel group of 8 (G <= 32: the lane's 72 kernel weights live in registers).
template <bool REGS>
__global__ void __launch_bounds__(256)
    dynconv_bce_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ t,
                           long long ldt, const float* __restrict__ mask, int Hm, int Wm, float* __restrict__ pred,
                           float* __restrict__ mask_out, float* __restrict__ loss_sum, int B, int H, int W, int C,
                           float inv_n, unsigned* __restrict__ metric_counts, float metric_thr) {
  extern __shared__ float sk[];  // [9][8][G]
  __shared__ float s_loss[8];
  __shared__ unsigned s_cnt[8][2];
  unsigned n_inter = 0, n_union = 0;  // trainMetricGPU (utils/misc.py:114-129) fused: sigmoid(pred) >= thr vs target != 0
  const int b = blockIdx.y;
  stage_kernel(t, ldt, b, C, sk);
  __syncthreads();
  const float bias = t[(long long)b * ldt + 9 * C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wp = W + 2, G = C / 8;
  const int sh = mask ? Hm / H : 1, sw = mask ? Wm / W : 1;
  float kr[REGS ? 72 : 1];
  if (REGS) {
#pragma unroll
    for (int i = 0; i < 72; ++i) kr[i] = lane < G ? sk[i * G + lane] : 0.f;
  }
  float lsum = 0.f;
  const int npix = H * W;
  const int p1 = min(npix, (blockIdx.x + 1) * kDynPix);
  for (int p = blockIdx.x * kDynPix + warp; p < p1; p += 8) {
    const int h = p / W, w = p - h * W;
    const __nv_bfloat16* xb = x + (((long long)b * (H + 2) + h) * wp + w) * ldx;
    float acc = 0.f;
    if (REGS) {
      if (lane < G) {
        float v[9][8];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) ld8(xb + ((long long)ky * wp + kx) * ldx + lane * 8, v[ky * 3 + kx]);
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = fmaf(v[tp][k], kr[tp * 8 + k], acc);
      }
    } else {
      for (int g = lane; g < G; g += 32) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
          float v[8];
          ld8(xb + ((long long)(tp / 3) * wp + (tp % 3)) * ldx + g * 8, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = fmaf(v[k], sk[(tp * 8 + k) * G + g], acc);
        }
      }
    }
    acc = warp_sum(acc) + bias;
    if (lane == 0) {
      pred[(long long)b * npix + p] = acc;
      if (mask != nullptr) {
        const float tg = mask[((long long)b * Hm + (long long)h * sh) * Wm + (long long)w * sw];
        mask_out[(long long)b * npix + p] = tg;
        lsum += fmaxf(acc, 0.f) - acc * tg + log1pf(__expf(-fabsf(acc)));
        const bool o = (1.f / (1.f + expf(-acc))) >= metric_thr, g1 = tg != 0.f;
        n_inter += (o && g1) ? 1u : 0u;
        n_union += (o || g1) ? 1u : 0u;
      }
    }
  }
  if (mask != nullptr) {
    if (lane == 0) { s_loss[warp] = lsum; s_cnt[warp][0] = n_inter; s_cnt[warp][1] = n_union; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      unsigned ci = 0, cu = 0;
      for (int i = 0; i < 8; ++i) { s += s_loss[i]; ci += s_cnt[i][0]; cu += s_cnt[i][1]; }
      atomicAdd(loss_sum, s * inv_n);
      if (metric_counts != nullptr) { atomicAdd(metric_counts + 2 * b, ci); atomicAdd(metric_counts + 2 * b + 1, cu); }
    }
  }
}

// dl[b,h,w] = g * (sigmoid(pred) - target) / n ; dt[b, 9C] += sum dl  (bias gradient)
__global__ void __launch_bounds__(256)
    bce_dlogit_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ g,
                      float* __restrict__ dl, float* __restrict__ dt, long long lddt, int npix, int C, float inv_n) {
  __shared__ float s_part[8];
  const int b = blockIdx.y;
  const float gs = g[0] * inv_n;
  float s = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const long long i = (long long)b * npix + p;
    const float l = pred[i];
    const float d = gs * (1.f / (1.f + __expf(-l)) - target[i]);
    dl[i] = d;
    s += d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += s_part[i];
    atomicAdd(dt + (long long)b * lddt + 9 * C, tot);
  }
}

// dX[b, h+1, w+1, c] = sum_tap dl[b, h-(ky-1), w-(kx-1)] * k[tap][c]; zero border
template <bool REGS>
__global__ void __launch_bounds__(256)
    dynconv_bwd_x_kernel(const float* __restrict__ dl, const float* __restrict__ t, long long ldt,
                         __nv_bfloat16* __restrict__ dx, long long lddx, int B, int H, int W, int C) {
  extern __shared__ float sk[];
  const int b = blockIdx.y;
  stage_kernel(t, ldt, b, C, sk);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = H + 2, wp = W + 2, G = C / 8;
  float kr[REGS ? 72 : 1];
  if (REGS) {
#pragma unroll
    for (int i = 0; i < 72; ++i) kr[i] = lane < G ? sk[i * G + lane] : 0.f;
  }
  const int r1 = min(hp * wp, (blockIdx.x + 1) * kDynPix);
  for (int r = blockIdx.x * kDynPix + warp; r < r1; r += 8) {
    const int hq = r / wp, wq = r - hq * wp;
    const bool interior = hq >= 1 && hq <= H && wq >= 1 && wq <= W;
    float d[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int hl = hq - 1 - (ky - 1), wl = wq - 1 - (kx - 1);
        d[ky * 3 + kx] = (interior && hl >= 0 && hl < H && wl >= 0 && wl < W)
                             ? dl[((long long)b * H + hl) * W + wl] : 0.f;
      }
    __nv_bfloat16* out = dx + ((long long)b * hp * wp + r) * lddx;
    if (REGS) {
      if (lane < G) {
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(d[tp], kr[tp * 8 + k], o[k]);
        st8(out + lane * 8, o);
      }
    } else {
      for (int g = lane; g < G; g += 32) {
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(d[tp], sk[(tp * 8 + k) * G + g], o[k]);
        st8(out + g * 8, o);
      }
    }
  }
}

// dt[b, c*9 + tap] += sum_{h,w} dl[b, h-(ky-1), w-(kx-1)] * x[b, h+1, w+1, c]
// block = (row band, image); thread = channel; 9 accumulators per thread.
__global__ void __launch_bounds__(256)
    dynconv_bwd_k_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ x, long long ldx,
                         float* __restrict__ dt, long long lddt, int B, int H, int W, int C, int rows_per_block) {
  extern __shared__ float sdl[];  // [(rows_per_block + 2)][W + 2] window of dl with zero halo
  const int b = blockIdx.y;
  const int h0 = blockIdx.x * rows_per_block;
  const int h1 = min(H, h0 + rows_per_block);
  const int ws = W + 2;
  for (int i = threadIdx.x; i < (rows_per_block + 2) * ws; i += blockDim.x) {
    const int rr = i / ws, cc = i - rr * ws;
    const int hl = h0 - 1 + rr, wl = cc - 1;
    sdl[i] = (hl >= 0 && hl < H && wl >= 0 && wl < W && rr < h1 - h0 + 2) ? dl[((long long)b * H + hl) * W + wl] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) acc[tp] = 0.f;
    for (int h = h0; h < h1; ++h) {
      for (int w = 0; w < W; ++w) {
        const float xv = bf2f(x[(((long long)b * (H + 2) + h + 1) * (W + 2) + w + 1) * ldx + c]);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            // logit pixel (h-(ky-1), w-(kx-1)) lives at window row (h-h0+1)-(ky-1), col (w+1)-(kx-1)
            acc[ky * 3 + kx] = fmaf(xv, sdl[(h - h0 + 2 - ky) * ws + (w + 2 - kx)], acc[ky * 3 + kx]);
      }
    }
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) atomicAdd(dt + (long long)b * lddt + c * 9 + tp, acc[tp]);
  }
}

}  // namespace cris

using namespace cris;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" {

/* x: padded NHWC bf16 [B,H+2,W+2,C]; t: fp32 [B, 9C+1] (pitch ldt) = proj.txt(state) in the reference's
 * (c, ky, kx) order; mask: fp32 [B,1,Hm,Wm] or NULL (eval); pred/mask_out: fp32 [B,H,W]; loss_sum: fp32
 * scalar, must be zeroed by the caller. */
int cris_dynconv_bce_fwd(const void* x, int64_t ldx, const float* t, int64_t ldt, const float* mask, int Hm, int Wm,
                         float* pred, float* mask_out, float* loss_sum, unsigned* metric_counts, float metric_thr, int B,
                         int H, int W, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0 && 9 * C * 4 <= 96 * 1024, "dynconv: C=%d unsupported", C);
  CRIS_CHECK_ARG(mask == nullptr || (Hm % H == 0 && Wm % W == 0), "dynconv: mask %dx%d not an integer multiple", Hm, Wm);
  CRIS_SET_SMEM_ONCE(dynconv_bce_fwd_kernel<false>, 96 * 1024);
  CRIS_SET_SMEM_ONCE(dynconv_bce_fwd_kernel<true>, 96 * 1024);
  dim3 grid((H * W + kDynPix - 1) / kDynPix, B);
  const float inv_n = 1.f / ((float)B * H * W);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (C <= 256)
    dynconv_bce_fwd_kernel<true><<<grid, 256, 9 * C * 4, STREAM>>>(xb, ldx, t, ldt, mask, Hm, Wm, pred, mask_out,
                                                                  loss_sum, B, H, W, C, inv_n, metric_counts, metric_thr);
  else
    dynconv_bce_fwd_kernel<false><<<grid, 256, 9 * C * 4, STREAM>>>(xb, ldx, t, ldt, mask, Hm, Wm, pred, mask_out,
                                                                   loss_sum, B, H, W, C, inv_n, metric_counts, metric_thr);
  CRIS_LAUNCH_OK();
  return 0;
}

/* backward: g = device scalar dLoss; writes dl (workspace fp32 [B,H,W]), dx (padded bf16), and
 * accumulates dt (fp32 [B, 9C+1], pitch lddt, zeroed by the caller). */
int cris_dynconv_bce_bwd(const void* x, int64_t ldx, const float* t, int64_t ldt, const float* pred,
                         const float* target, const float* g, float* dl, void* dx, int64_t lddx, float* dt,
                         int64_t lddt, int B, int H, int W, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0 && 9 * C * 4 <= 96 * 1024, "dynconv: C=%d unsupported", C);
  CRIS_SET_SMEM_ONCE(dynconv_bwd_x_kernel<false>, 96 * 1024);
  CRIS_SET_SMEM_ONCE(dynconv_bwd_x_kernel<true>, 96 * 1024);
  const int npix = H * W;
  const float inv_n = 1.f / ((float)B * npix);
  bce_dlogit_kernel<<<dim3((npix + 255) / 256, B), 256, 0, STREAM>>>(pred, target, g, dl, dt, lddt, npix, C, inv_n);
  CRIS_LAUNCH_OK();
  dim3 gx(((H + 2) * (W + 2) + kDynPix - 1) / kDynPix, B);
  if (C <= 256)
    dynconv_bwd_x_kernel<true><<<gx, 256, 9 * C * 4, STREAM>>>(dl, t, ldt, reinterpret_cast<__nv_bfloat16*>(dx), lddx, B,
                                                              H, W, C);
  else
    dynconv_bwd_x_kernel<false><<<gx, 256, 9 * C * 4, STREAM>>>(dl, t, ldt, reinterpret_cast<__nv_bfloat16*>(dx), lddx, B,
                                                               H, W, C);
  CRIS_LAUNCH_OK();
  const int rpb = 4;
  dim3 gk((H + rpb - 1) / rpb, B);
  dynconv_bwd_k_kernel<<<gk, 256, (rpb + 2) * (W + 2) * 4, STREAM>>>(
      dl, reinterpret_cast<const __nv_bfloat16*>(x), ldx, dt, lddt, B, H, W, C, rpb);
  CRIS_LAUNCH_OK();
  return 0;
}
}
