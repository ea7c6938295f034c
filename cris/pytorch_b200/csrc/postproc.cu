// postproc.cu — evaluation post-processing on the GPU: logits -> IoU against the ground-truth mask of the ORIGINAL
// photo, the step right after the model in the reference's validate()/inference() (engine/engine.py:101-124,172-190):
//
//   pred = sigmoid(logits) ; pred = F.interpolate(pred, (416,416), mode='bicubic', align_corners=True)
//   pred = cv2.warpAffine(pred, mat_inv, (w, h), flags=cv2.INTER_CUBIC, borderValue=0.) ; pred = pred > 0.35
//   iou  = sum(pred & gt) / (sum(pred | gt) + 1e-6)
//
// The reference copies every prediction to the host and calls OpenCV per sample.  Here two kernels do it for a whole
// batch: (1) sigmoid + ATen-semantics bicubic upsampling, (2) OpenCV-semantics affine warp (fixed-point coordinates
// at 1/32 pixel, float cubic kernels A = -0.75 from a 32-entry table, zero border), threshold, and the per-sample
// intersection / union counts — only 16 bytes per sample travel back.  Semantics restated and pinned against torch +
// cv2 in oracle/postproc_oracle.py.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace cris {

__constant__ float c_cubic_tab[32 * 4];  // OpenCV initInterTab1D(INTER_CUBIC): weights at x = i/32

// ATen get_cubic_upsample_coefficients (A = -0.75), float
__device__ __forceinline__ void aten_cubic(float t, float* c) {
  const float A = -0.75f;
  const float x1 = t + 1.f, x2 = 1.f - t, x3 = x2 + 1.f;
  c[0] = ((A * x1 - 5.f * A) * x1 + 8.f * A) * x1 - 4.f * A;
  c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void postproc_upsample_kernel(const float* __restrict__ logits, float* __restrict__ out, int B, int H, int W,
                                         int OH, int OW) {
  const long long total = (long long)B * OH * OW;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    const int oy = (int)((i / OW) % OH);
    const int b = (int)(i / ((long long)OW * OH));
    const float ry = sy * (float)oy, rx = sx * (float)ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float wy[4], wx[4];
    aten_cubic(ry - (float)iy, wy);
    aten_cubic(rx - (float)ix, wx);
    const float* src = logits + (long long)b * H * W;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int yy = min(max(iy - 1 + r, 0), H - 1);
      float row = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = min(max(ix - 1 + c, 0), W - 1);
        const float p = 1.f / (1.f + __expf(-src[yy * W + xx]));  // sigmoid of the logit (engine.py:103)
        row = __fadd_rn(row, __fmul_rn(p, wx[c]));
      }
      acc = __fadd_rn(acc, __fmul_rn(row, wy[r]));
    }
    out[i] = acc;
  }
}

struct WarpSample {
  double m[6];        // the matrix handed to cv2.warpAffine (maps SOURCE -> destination; inverted here like OpenCV does)
  int h, w;           // destination (original photo) size
  long long off;      // offset of this sample's ground-truth mask / prediction in the packed byte buffers
};

__global__ void postproc_warp_iou_kernel(const float* __restrict__ up, int SH, int SW, const WarpSample* __restrict__ samples,
                                         const uint8_t* __restrict__ gt, uint8_t* __restrict__ pred_out, float thr,
                                         unsigned long long* __restrict__ counts) {
  const int b = blockIdx.y;
  const WarpSample s = samples[b];
  // imgwarp.cpp: invert the 2x3 matrix in double precision
  double M0 = s.m[0], M1 = s.m[1], M2 = s.m[2], M3 = s.m[3], M4 = s.m[4], M5 = s.m[5];
  double D = M0 * M4 - M1 * M3;
  D = D != 0.0 ? 1.0 / D : 0.0;
  const double A11 = M4 * D, A22 = M0 * D;
  M0 = A11; M1 *= -D; M3 *= -D; M4 = A22;
  const double b1 = -M0 * M2 - M1 * M5, b2 = -M3 * M2 - M4 * M5;
  M2 = b1; M5 = b2;
  const float* src = up + (long long)b * SH * SW;
  const long long npix = (long long)s.h * s.w;
  unsigned inter = 0, uni = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % s.w), y = (int)(i / s.w);
    // fixed point, AB_BITS = 10, rounded to 1/32 pixel (INTER_BITS = 5); saturate_cast<int>(double) = round half even
    const int adelta = __double2int_rn(M0 * (double)x * 1024.0), bdelta = __double2int_rn(M3 * (double)x * 1024.0);
    const int X0 = __double2int_rn((M1 * (double)y + M2) * 1024.0) + 16, Y0 = __double2int_rn((M4 * (double)y + M5) * 1024.0) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = min(max(X >> 5, -32768), 32767) - 1, sy = min(max(Y >> 5, -32768), 32767) - 1;
    const float* wx = c_cubic_tab + (X & 31) * 4;
    const float* wy = c_cubic_tab + (Y & 31) * 4;
    float acc = 0.f;
    if (sx + 3 >= 0 && sx < SW && sy + 3 >= 0 && sy < SH) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int yy = sy + r;
        float row = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int xx = sx + c;
          const float v = (yy >= 0 && yy < SH && xx >= 0 && xx < SW) ? src[yy * SW + xx] : 0.f;  // BORDER_CONSTANT 0
          row = __fadd_rn(row, __fmul_rn(v, __fmul_rn(wy[r], wx[c])));
        }
        acc = __fadd_rn(acc, row);
      }
    }
    const bool p = acc > thr;
    const bool g = gt != nullptr && gt[s.off + i] != 0;
    if (pred_out != nullptr) pred_out[s.off + i] = p ? 1 : 0;
    inter += (p && g) ? 1u : 0u;
    uni += (p || g) ? 1u : 0u;
  }
  // block reduction -> two 64-bit atomics per block
  __shared__ unsigned s_i[32], s_u[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_i[warp] = inter; s_u[warp] = uni; }
  __syncthreads();
  if (warp == 0) {
    inter = lane < (int)(blockDim.x >> 5) ? s_i[lane] : 0u;
    uni = lane < (int)(blockDim.x >> 5) ? s_u[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      inter += __shfl_xor_sync(0xffffffffu, inter, o);
      uni += __shfl_xor_sync(0xffffffffu, uni, o);
    }
    if (lane == 0) {
      atomicAdd(&counts[2 * b], (unsigned long long)inter);
      atomicAdd(&counts[2 * b + 1], (unsigned long long)uni);
    }
  }
}

static int upload_cubic_tab() {
  // OpenCV interpolateCubic in float, evaluated by the HOST compiler (no fused multiply-add contraction)
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  static int dev_mask_done[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (dev_mask_done[dev & 63]) return 0;
  float tab[32 * 4];
  const float A = -0.75f;
  for (int i = 0; i < 32; ++i) {
    volatile float x = (float)i * (1.f / 32.f);
    volatile float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    volatile float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
    volatile float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    tab[i * 4 + 0] = c0; tab[i * 4 + 1] = c1; tab[i * 4 + 2] = c2;
    tab[i * 4 + 3] = 1.f - c0 - c1 - c2;
  }
  (void)once;
  err = cudaMemcpyToSymbol(c_cubic_tab, tab, sizeof(tab));
  if (err != cudaSuccess) {
    set_error("postproc: cubic table upload failed: %s", cudaGetErrorString(err));
    return -2;
  }
  dev_mask_done[dev & 63] = 1;
  return 0;
}

}  // namespace cris

using namespace cris;

extern "C" {

int cris_postproc_sample_bytes(void) { return (int)sizeof(WarpSample); }

int cris_postproc_upsample(const float* logits, float* prob_up, int B, int H, int W, int OH, int OW, void* stream) {
  CRIS_CHECK_ARG(logits && prob_up && B >= 1 && H >= 1 && W >= 1 && OH >= 1 && OW >= 1, "postproc_upsample: bad argument");
  const long long total = (long long)B * OH * OW;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  postproc_upsample_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, prob_up, B, H, W, OH, OW);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_postproc_warp_iou(const float* prob_up, int B, int SH, int SW, const void* samples_dev, const uint8_t* gt,
                           uint8_t* pred_out, float thr, unsigned long long* counts, long long max_pixels, void* stream) {
  CRIS_CHECK_ARG(prob_up && samples_dev && counts && B >= 1 && B <= 65535 && max_pixels >= 1, "postproc_warp_iou: bad argument");
  if (int rc = upload_cubic_tab()) return rc;
  long long gx = (max_pixels + 255) / 256;
  if (gx > 1024) gx = 1024;
  dim3 grid((unsigned)gx, (unsigned)B);
  postproc_warp_iou_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      prob_up, SH, SW, reinterpret_cast<const WarpSample*>(samples_dev), gt, pred_out, thr, counts);
  CRIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
