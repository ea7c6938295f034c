// spatial.cu — HBM-bound layout / resampling kernels on the padded-NHWC activation layout
// ([N, H+2, W+2, C] bf16 with a zero border): avg-pool 2x2, bilinear x2 upsample, per-image channel
// gating, padded<->token conversion, CoordConv channels, fp32 NCHW image -> stem conv.
// Reference call sites: nn.AvgPool2d (model/clip.py:23,35,184), F.avg_pool2d (model/layers.py:297),
// nn.Upsample / F.interpolate bilinear (model/layers.py:54,56,293,304), f5*state (layers.py:290),
// reshape/permute glue (clip.py:113-118,140; layers.py:166,179), CoordConv.add_coord (layers.py:30-39),
// stem conv1 (clip.py:165-170).
#include "vec.cuh"

namespace cris {

// ---- avg-pool 2x2 stride 2 -------------------------------------------------------------------
__global__ void avgpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                                    long long ldy, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, G = C / 8;
  const int hpo = Ho + 2, wpo = Wo + 2, wpi = W + 2;
  const long long total = (long long)N * hpo * wpo * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % wpo);
    const int h = (int)((r / wpo) % hpo);
    const int n = (int)(r / ((long long)wpo * hpo));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= Ho && w >= 1 && w <= Wo) {
      const long long base = ((long long)n * (H + 2) + (2 * (h - 1) + 1)) * wpi + (2 * (w - 1) + 1);
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float v[8];
          ld8(x + (base + dy * wpi + dx) * ldx + g * 8, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += v[k];
        }
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] *= 0.25f;
    }
    st8(y + r * ldy + g * 8, o);
  }
}

// dx (+)= 0.25 * dy[h/2, w/2]
__global__ void avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                    __nv_bfloat16* __restrict__ dx, long long lddx, int accumulate, int N, int H,
                                    int W, int C) {
  const int G = C / 8;
  const int hpi = H + 2, wpi = W + 2, wpo = W / 2 + 2, hpo = H / 2 + 2;
  const long long total = (long long)N * hpi * wpi * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % wpi);
    const int h = (int)((r / wpi) % hpi);
    const int n = (int)(r / ((long long)wpi * hpi));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= H && w >= 1 && w <= W) {
      const long long ro = ((long long)n * hpo + ((h - 1) / 2 + 1)) * wpo + ((w - 1) / 2 + 1);
      ld8(dy + ro * lddy + g * 8, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] *= 0.25f;
      if (accumulate) {
        float old[8];
        ld8(dx + r * lddx + g * 8, old);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += old[k];
      }
    }
    st8(dx + r * lddx + g * 8, o);
  }
}

// ---- bilinear x2, align_corners=False --------------------------------------------------------
// source taps of destination index d (PyTorch upsample_bilinear2d semantics)
__device__ __forceinline__ void bilin_src(int d, int n_in, int& i0, int& i1, float& lam) {
  float s = (d + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = min(i0 + 1, n_in - 1);
  lam = s - (float)i0;
}

__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                      __nv_bfloat16* __restrict__ y, long long ldy, int N, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, G = C / 8;
  const int hpo = Ho + 2, wpo = Wo + 2, wpi = W + 2, hpi = H + 2;
  const long long total = (long long)N * hpo * wpo * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % wpo);
    const int h = (int)((r / wpo) % hpo);
    const int n = (int)(r / ((long long)wpo * hpo));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= Ho && w >= 1 && w <= Wo) {
      int h0, h1, w0, w1;
      float lh, lw;
      bilin_src(h - 1, H, h0, h1, lh);
      bilin_src(w - 1, W, w0, w1, lw);
      const long long b = (long long)n * hpi;
      float v00[8], v01[8], v10[8], v11[8];
      ld8(x + ((b + h0 + 1) * wpi + w0 + 1) * ldx + g * 8, v00);
      ld8(x + ((b + h0 + 1) * wpi + w1 + 1) * ldx + g * 8, v01);
      ld8(x + ((b + h1 + 1) * wpi + w0 + 1) * ldx + g * 8, v10);
      ld8(x + ((b + h1 + 1) * wpi + w1 + 1) * ldx + g * 8, v11);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o[k] = (1.f - lh) * ((1.f - lw) * v00[k] + lw * v01[k]) + lh * ((1.f - lw) * v10[k] + lw * v11[k]);
    }
    st8(y + r * ldy + g * 8, o);
  }
}

// gather form of the transpose: each input pixel collects from the <= 4x4 outputs that read it
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                      __nv_bfloat16* __restrict__ dx, long long lddx, int accumulate, int N, int H,
                                      int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, G = C / 8;
  const int hpo = Ho + 2, wpo = Wo + 2, wpi = W + 2, hpi = H + 2;
  const long long total = (long long)N * hpi * wpi * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % wpi);
    const int h = (int)((r / wpi) % hpi);
    const int n = (int)(r / ((long long)wpi * hpi));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= H && w >= 1 && w <= W) {
      const int hi = h - 1, wi = w - 1;
      for (int oh = max(0, 2 * hi - 1); oh <= min(Ho - 1, 2 * hi + 2); ++oh) {
        int a0, a1;
        float la;
        bilin_src(oh, H, a0, a1, la);
        const float wh = (a0 == hi ? 1.f - la : 0.f) + (a1 == hi ? la : 0.f);
        if (wh == 0.f) continue;
        for (int ow = max(0, 2 * wi - 1); ow <= min(Wo - 1, 2 * wi + 2); ++ow) {
          int b0, b1;
          float lb;
          bilin_src(ow, W, b0, b1, lb);
          const float ww = (b0 == wi ? 1.f - lb : 0.f) + (b1 == wi ? lb : 0.f);
          if (ww == 0.f) continue;
          float v[8];
          ld8(dy + (((long long)n * hpo + oh + 1) * wpo + ow + 1) * lddy + g * 8, v);
          const float f = wh * ww;
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += f * v[k];
        }
      }
      if (accumulate) {
        float old[8];
        ld8(dx + r * lddx + g * 8, old);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += old[k];
      }
    }
    st8(dx + r * lddx + g * 8, o);
  }
}

// ---- per-image channel gating: y[r, c] = x[r, c] * s[image(r), c] ------------------------------
__global__ void mul_bcast_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                 const __nv_bfloat16* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ y,
                                 long long ldy, long long rows, int rows_per_image, int C) {
  const int G = C / 8;
  const long long total = rows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / G;
    const int c = (int)(i - r * G) * 8;
    float a[8], b[8];
    ld8(x + r * ldx + c, a);
    ld8(s + (r / rows_per_image) * lds + c, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= b[k];
    st8(y + r * ldy + c, a);
  }
}
// ds[n, c] = sum over the rows of image n of dy*x   (one block per (image, 8-channel group... ) )
__global__ void mul_bcast_bwd_s_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                       const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ ds,
                                       int rows_per_image, int C) {
  const int n = blockIdx.y;
  const int c = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * 8;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = lane; r < rows_per_image; r += 32) {
    const long long row = (long long)n * rows_per_image + r;
    float a[8], b[8];
    ld8(dy + row * lddy + c, a);
    ld8(x + row * ldx + c, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += a[k] * b[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) st8f(ds + (long long)n * C + c, acc);
}

// ---- padded NHWC <-> dense tokens ------------------------------------------------------------
// tok[n*H*W + h*W + w, :] = x[n, h+1, w+1, :] (+ add[h*W+w, :])
__global__ void padded_to_tokens_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                        const float* __restrict__ add, long long ldadd, void* __restrict__ tok,
                                        int tok_fp32, long long ldt, int N, int H, int W, int C) {
  const int G = C / 8;
  const long long total = (long long)N * H * W * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long t = i / G;
    const int w = (int)(t % W);
    const int h = (int)((t / W) % H);
    const int n = (int)(t / ((long long)W * H));
    float v[8];
    ld8(x + (((long long)n * (H + 2) + h + 1) * (W + 2) + w + 1) * ldx + g * 8, v);
    if (add != nullptr) {
      float a[8];
      ld8f(add + (long long)(h * W + w) * ldadd + g * 8, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += a[k];
    }
    st8x(tok, t * ldt + g * 8, tok_fp32, v);
  }
}
// y[n, h+1, w+1, :] = tok[...] on interior rows, zero on the border
__global__ void tokens_to_padded_kernel(const void* __restrict__ tok, int tok_fp32, long long ldt,
                                        __nv_bfloat16* __restrict__ y, long long ldy, int N, int H, int W, int C) {
  const int G = C / 8;
  const int hp = H + 2, wp = W + 2;
  const long long total = (long long)N * hp * wp * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long r = i / G;
    const int w = (int)(r % wp);
    const int h = (int)((r / wp) % hp);
    const int n = (int)(r / ((long long)wp * hp));
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= H && w >= 1 && w <= W)
      ld8x(tok, (((long long)n * H + h - 1) * W + w - 1) * ldt + g * 8, tok_fp32, v);
    st8(y + r * ldy + g * 8, v);
  }
}

// ---- CoordConv coordinate channels: buf[r, c0] = x in [-1,1], buf[r, c0+1] = y, rest of the 8-group 0
__global__ void coord_fill_kernel(__nv_bfloat16* __restrict__ buf, long long ld, int c0, int N, int H, int W) {
  const int hp = H + 2, wp = W + 2;
  const long long total = (long long)N * hp * wp;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total;
       r += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(r % wp);
    const int h = (int)((r / wp) % hp);
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h >= 1 && h <= H && w >= 1 && w <= W) {
      v[0] = W > 1 ? -1.f + 2.f * (w - 1) / (W - 1) : -1.f;  // torch.linspace(-1, 1, W)
      v[1] = H > 1 ? -1.f + 2.f * (h - 1) / (H - 1) : -1.f;
    }
    st8(buf + r * ld + c0, v);
  }
}

// ---- stem conv1: fp32 NCHW image, 3x3 stride 2 pad 1, 3 -> Cout (<= 64), output padded NHWC bf16 (pre-BN)
// one thread per (output pixel, 8 output channels); weights [Cout][27] fp32 in smem.
__global__ void __launch_bounds__(256)
    stem_conv1_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, __nv_bfloat16* __restrict__ z,
                          long long ldz, int N, int Hin, int Win, int Cout) {
  extern __shared__ float sw[];  // [27][Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;
    sw[i] = w[co * 27 + k];
  }
  __syncthreads();
  const int Ho = Hin / 2, Wo = Win / 2, G = Cout / 8;
  const int hp = Ho + 2, wp = Wo + 2;
  const long long total = (long long)N * hp * wp * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long r = i / G;
    const int wq = (int)(r % wp);
    const int hq = (int)((r / wp) % hp);
    const int n = (int)(r / ((long long)wp * hp));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (hq >= 1 && hq <= Ho && wq >= 1 && wq <= Wo) {
      const int ho = hq - 1, wo = wq - 1;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int hi = 2 * ho + ky - 1;
          if (hi < 0 || hi >= Hin) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int wi = 2 * wo + kx - 1;
            if (wi < 0 || wi >= Win) continue;
            const float v = __ldg(img + (((long long)n * 3 + ci) * Hin + hi) * Win + wi);
            const float* wk = sw + (ci * 9 + ky * 3 + kx) * Cout + g * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaf(v, wk[k], o[k]);
          }
        }
    }
    st8(z + r * ldz + g * 8, o);
  }
}
// wgrad of stem conv1: dW[co][ci][ky][kx] = sum_pixels dz[pix, co] * img[...]; block-partial + atomics
__global__ void __launch_bounds__(256)
    stem_conv1_wgrad_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dz, long long lddz,
                            float* __restrict__ dw, int N, int Hin, int Win, int Cout) {
  // thread = (k in 0..26, co) pair handled as: blockDim = (Cout, 8); k loops
  const int Ho = Hin / 2, Wo = Win / 2;
  const int hp = Ho + 2, wp = Wo + 2;
  const long long npix = (long long)N * Ho * Wo;
  const long long ppb = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * ppb, p1 = min(npix, p0 + ppb);
  const int co = threadIdx.x % Cout;
  const int kslot = threadIdx.x / Cout;        // 0 .. blockDim.x/Cout - 1
  const int kslots = blockDim.x / Cout;
  float acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.f;
  for (long long pidx = p0 + kslot; pidx < p1; pidx += kslots) {
    const int wo = (int)(pidx % Wo);
    const int ho = (int)((pidx / Wo) % Ho);
    const int n = (int)(pidx / ((long long)Wo * Ho));
    const float d = bf2f(dz[(((long long)n * hp + ho + 1) * wp + wo + 1) * lddz + co]);
    if (d == 0.f) continue;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hi = 2 * ho + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int wi = 2 * wo + kx - 1;
          if (hi >= 0 && hi < Hin && wi >= 0 && wi < Win)
            acc[ci * 9 + ky * 3 + kx] += d * __ldg(img + (((long long)n * 3 + ci) * Hin + hi) * Win + wi);
        }
      }
  }
#pragma unroll
  for (int k = 0; k < 27; ++k) atomicAdd(dw + co * 27 + k, acc[k]);
}

// ---- stem im2col: fp32 NCHW image -> bf16 patches [N, H/2+2, W/2+2, 32] (27 taps (ci,ky,kx) of the 3x3 stride-2
// pad-1 stem convolution + 5 zero channels), so that conv1 and its wgrad run on the tcgen05 GEMM core.
__global__ void __launch_bounds__(256)
    stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int N, int Hin, int Win) {
  const int Ho = Hin / 2, Wo = Win / 2;
  const int hp = Ho + 2, wp = Wo + 2;
  const long long total = (long long)N * hp * wp * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 3);
    const long long r = i >> 2;
    const int wq = (int)(r % wp);
    const int hq = (int)((r / wp) % hp);
    const int n = (int)(r / ((long long)wp * hp));
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (hq >= 1 && hq <= Ho && wq >= 1 && wq <= Wo) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = g * 8 + k;
        if (t < 27) {
          const int ci = t / 9, ky = (t % 9) / 3, kx = t % 3;
          const int hi = 2 * (hq - 1) + ky - 1, wi = 2 * (wq - 1) + kx - 1;
          if (hi >= 0 && hi < Hin && wi >= 0 && wi < Win)
            o[k] = __ldg(img + (((long long)n * 3 + ci) * Hin + hi) * Win + wi);
        }
      }
    }
    st8(out + r * 32 + g * 8, o);
  }
}

}  // namespace cris

using namespace cris;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int cris_avgpool2_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int N, int H, int W, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: bad shape H=%d W=%d C=%d", H, W, C);
  const long long work = (long long)N * (H / 2 + 2) * (W / 2 + 2) * (C / 8);
  avgpool2_fwd_kernel<<<grid_for(work, 256), 256, 0, STREAM>>>(CBF(x), ldx, BF(y), ldy, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_avgpool2_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int accumulate, int N, int H, int W, int C,
                      void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: bad shape H=%d W=%d C=%d", H, W, C);
  const long long work = (long long)N * (H + 2) * (W + 2) * (C / 8);
  avgpool2_bwd_kernel<<<grid_for(work, 256), 256, 0, STREAM>>>(CBF(dy), lddy, BF(dx), lddx, accumulate, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int N, int H, int W, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "upsample2x: C=%d", C);
  const long long work = (long long)N * (2 * H + 2) * (2 * W + 2) * (C / 8);
  upsample2x_fwd_kernel<<<grid_for(work, 256), 256, 0, STREAM>>>(CBF(x), ldx, BF(y), ldy, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int accumulate, int N, int H, int W,
                        int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "upsample2x: C=%d", C);
  const long long work = (long long)N * (H + 2) * (W + 2) * (C / 8);
  upsample2x_bwd_kernel<<<grid_for(work, 256), 256, 0, STREAM>>>(CBF(dy), lddy, BF(dx), lddx, accumulate, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_mul_bcast(const void* x, int64_t ldx, const void* s, int64_t lds, void* y, int64_t ldy, int64_t rows,
                   int rows_per_image, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "mul_bcast: C=%d", C);
  mul_bcast_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, STREAM>>>(CBF(x), ldx, CBF(s), lds, BF(y), ldy, rows,
                                                                      rows_per_image, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_mul_bcast_bwd_s(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* ds, int n_images,
                         int rows_per_image, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "mul_bcast_bwd_s: C=%d", C);
  dim3 grid((C / 8 + 7) / 8, n_images);
  mul_bcast_bwd_s_kernel<<<grid, 256, 0, STREAM>>>(CBF(dy), lddy, CBF(x), ldx, ds, rows_per_image, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_padded_to_tokens(const void* x, int64_t ldx, const float* add, int64_t ldadd, void* tok, int tok_fp32,
                          int64_t ldt, int N, int H, int W, int C, void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "padded_to_tokens: C=%d", C);
  padded_to_tokens_kernel<<<grid_for((long long)N * H * W * (C / 8), 256), 256, 0, STREAM>>>(
      CBF(x), ldx, add, ldadd, tok, tok_fp32, ldt, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_tokens_to_padded(const void* tok, int tok_fp32, int64_t ldt, void* y, int64_t ldy, int N, int H, int W, int C,
                          void* stream) {
  CRIS_CHECK_ARG(C % 8 == 0, "tokens_to_padded: C=%d", C);
  tokens_to_padded_kernel<<<grid_for((long long)N * (H + 2) * (W + 2) * (C / 8), 256), 256, 0, STREAM>>>(
      tok, tok_fp32, ldt, BF(y), ldy, N, H, W, C);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_coord_fill(void* buf, int64_t ld, int c0, int N, int H, int W, void* stream) {
  CRIS_CHECK_ARG(c0 % 8 == 0, "coord_fill: c0=%d must be a multiple of 8", c0);
  coord_fill_kernel<<<grid_for((long long)N * (H + 2) * (W + 2), 256), 256, 0, STREAM>>>(BF(buf), ld, c0, N, H, W);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_stem_im2col(const float* img, void* out, int N, int Hin, int Win, void* stream) {
  CRIS_CHECK_ARG(Hin % 2 == 0 && Win % 2 == 0, "stem_im2col: odd image size");
  const long long work = (long long)N * (Hin / 2 + 2) * (Win / 2 + 2) * 4;
  stem_im2col_kernel<<<grid_for(work, 256, 148 * 32), 256, 0, STREAM>>>(img, BF(out), N, Hin, Win);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_stem_conv1_fwd(const float* img, const float* w, void* z, int64_t ldz, int N, int Hin, int Win, int Cout,
                        void* stream) {
  CRIS_CHECK_ARG(Cout % 8 == 0 && Cout <= 64 && Hin % 2 == 0 && Win % 2 == 0, "stem_conv1: bad shape");
  const long long work = (long long)N * (Hin / 2 + 2) * (Win / 2 + 2) * (Cout / 8);
  stem_conv1_fwd_kernel<<<grid_for(work, 256, 148 * 32), 256, 27 * Cout * 4, STREAM>>>(img, w, BF(z), ldz, N, Hin, Win,
                                                                                     Cout);
  CRIS_LAUNCH_OK();
  return 0;
}
int cris_stem_conv1_wgrad(const float* img, const void* dz, int64_t lddz, float* dw, int N, int Hin, int Win, int Cout,
                          void* stream) {
  CRIS_CHECK_ARG(Cout <= 256 && 256 % Cout == 0, "stem_conv1_wgrad: Cout=%d must divide 256", Cout);
  stem_conv1_wgrad_kernel<<<148 * 4, 256, 0, STREAM>>>(img, CBF(dz), lddz, dw, N, Hin, Win, Cout);
  CRIS_LAUNCH_OK();
  return 0;
}
}
