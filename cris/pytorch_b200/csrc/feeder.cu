// feeder.cu — the input transform of the reference's dataset on the GPU, the step right BEFORE the model
// (utils/dataset.py:148-163 letterbox, :210-221 convert):
//
//   img  = cv2.warpAffine(photo_rgb_u8, mat, (416,416), flags=INTER_CUBIC, borderValue=[mean*255])
//   mask = cv2.warpAffine(mask_u8,      mat, (416,416), flags=INTER_LINEAR, borderValue=0.) / 255.
//   img  = img.transpose(2,0,1).float().div_(255.).sub_(mean).div_(std)
//
// The reference does this per sample on loader workers and ships fp32 tensors (2.1 MB + 0.7 MB per sample) to the
// GPU; here the decoded 8-bit photos go up (~0.9 MB for 480x640) and ONE kernel per batch writes the model's fp32
// NCHW input and the fp32 mask.  8-bit warpAffine semantics restated exactly (oracle/feeder_oracle.py, pinned
// bit-exact against cv2): fixed-point destination coordinates (AB_BITS 10, 1/32 pixel), integer 2-D weight tables
// (float outer product * 2^15 rounded to short, one central tap corrected so the sum is 2^15), (sum + 2^14) >> 15
// saturated; constant border: whole footprint outside -> border value, else outside taps contribute it.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace cris {

__device__ short g_cubic_i[1024 * 16];   // [fy*32 + fx][row*4 + col]
__device__ short g_linear_i[1024 * 4];   // [fy*32 + fx][row*2 + col]

struct FeedSample {
  double m[6];         // the matrix handed to cv2.warpAffine (source -> destination)
  int h, w;            // source photo size
  long long img_off;   // byte offset of this sample's RGB HWC photo in `images`
  long long mask_off;  // byte offset of its [h, w] mask in `masks`, or -1
};

struct FeedConst {
  int cval[3];
  float mean[3], stdv[3];
};

template <int K>
__device__ __forceinline__ int warp_tap_sum(const uint8_t* __restrict__ src, int cn, int c, int sh, int sw, int sx, int sy,
                                            const short* __restrict__ w, int cval) {
  // remapBicubic / remapBilinear with BORDER_CONSTANT: cval * ONE + sum over inside taps of (S - cval) * w
  int sum = cval << 15;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int yy = sy + i;
    if (yy < 0 || yy >= sh) continue;
    const uint8_t* row = src + ((long long)yy * sw) * cn + c;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int xx = sx + j;
      if (xx >= 0 && xx < sw) sum += ((int)row[(long long)xx * cn] - cval) * (int)w[i * K + j];
    }
  }
  const int v = (sum + (1 << 14)) >> 15;
  return min(max(v, 0), 255);
}

__global__ void feeder_letterbox_kernel(const uint8_t* __restrict__ images, const uint8_t* __restrict__ masks,
                                        const FeedSample* __restrict__ samples, int OH, int OW, FeedConst k,
                                        float* __restrict__ img_out, float* __restrict__ mask_out) {
  const int b = blockIdx.y;
  const FeedSample s = samples[b];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OH * OW) return;
  const int x = i % OW, y = i / OW;
  // imgwarp.cpp: invert the 2x3 matrix in double precision
  double M0 = s.m[0], M1 = s.m[1], M2 = s.m[2], M3 = s.m[3], M4 = s.m[4], M5 = s.m[5];
  double D = M0 * M4 - M1 * M3;
  D = D != 0.0 ? 1.0 / D : 0.0;
  const double A11 = M4 * D, A22 = M0 * D;
  M0 = A11; M1 *= -D; M3 *= -D; M4 = A22;
  const double b1 = -M0 * M2 - M1 * M5, b2 = -M3 * M2 - M4 * M5;
  M2 = b1; M5 = b2;
  const int adelta = __double2int_rn(M0 * (double)x * 1024.0), bdelta = __double2int_rn(M3 * (double)x * 1024.0);
  const int X0 = __double2int_rn((M1 * (double)y + M2) * 1024.0) + 16, Y0 = __double2int_rn((M4 * (double)y + M5) * 1024.0) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int ix = min(max(X >> 5, -32768), 32767), iy = min(max(Y >> 5, -32768), 32767);
  const int frac = (Y & 31) * 32 + (X & 31);
  {
    const int sx = ix - 1, sy = iy - 1;
    const bool outside = sx >= s.w || sx + 4 <= 0 || sy >= s.h || sy + 4 <= 0;
    const uint8_t* src = images + s.img_off;
    const short* w = g_cubic_i + frac * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = outside ? k.cval[c] : warp_tap_sum<4>(src, 3, c, s.h, s.w, sx, sy, w, k.cval[c]);
      // convert(): float32, three separately rounded operations
      const float t = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), k.mean[c]), k.stdv[c]);
      img_out[(((long long)b * 3 + c) * OH + y) * OW + x] = t;
    }
  }
  if (mask_out != nullptr) {
    int v = 0;
    if (s.mask_off >= 0) {
      const bool outside = ix >= s.w || ix + 2 <= 0 || iy >= s.h || iy + 2 <= 0;
      if (!outside) v = warp_tap_sum<2>(masks + s.mask_off, 1, 0, s.h, s.w, ix, iy, g_linear_i + frac * 4, 0);
    }
    mask_out[((long long)b * OH + y) * OW + x] = (float)((double)v / 255.0);   // numpy uint8 / 255. (float64), then .float()
  }
}

// initInterTab2D(method, fixpt = true) evaluated by the HOST compiler (float, no fused multiply-add contraction)
static void build_fixed_tab(const float* tab1d, int k, short* out) {
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      short* it = out + (i * 32 + j) * k * k;
      int isum = 0;
      for (int k1 = 0; k1 < k; ++k1) {
        const float vy = tab1d[i * k + k1];
        for (int k2 = 0; k2 < k; ++k2) {
          volatile float v = vy * tab1d[j * k + k2];
          volatile float sc = v * 32768.f;
          long r = lrintf(sc);                       // saturate_cast<short>(float): round half to even, saturate
          r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
          it[k1 * k + k2] = (short)r;
          isum += (int)r;
        }
      }
      if (isum != 32768) {
        const int diff = isum - 32768, n = k * k, k2h = k / 2;
        int big = k2h * k + k2h, small = big;
        // OpenCV walks k1, k2 over {k/2, k/2 + 1}: entries past this record read as 0 (not yet written)
        auto at = [&](int idx) { return idx < n ? (int)it[idx] : 0; };
        for (int k1 = k2h; k1 < k2h + 2; ++k1)
          for (int k2 = k2h; k2 < k2h + 2; ++k2) {
            const int idx = k1 * k + k2;
            if (at(idx) < at(small)) small = idx;
            else if (at(idx) > at(big)) big = idx;
          }
        const int tgt = diff < 0 ? big : small;
        if (tgt < n) it[tgt] = (short)(it[tgt] - diff);
      }
    }
}

static int upload_feeder_tabs() {
  static std::mutex mu;
  static int done[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev & 63]) return 0;
  float cub[32 * 4], lin[32 * 2];
  const float A = -0.75f;
  for (int i = 0; i < 32; ++i) {
    volatile float x = (float)i * (1.f / 32.f);
    volatile float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    volatile float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
    volatile float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    cub[i * 4 + 0] = c0; cub[i * 4 + 1] = c1; cub[i * 4 + 2] = c2;
    cub[i * 4 + 3] = 1.f - c0 - c1 - c2;
    lin[i * 2 + 0] = 1.f - x; lin[i * 2 + 1] = x;
  }
  static short tc[1024 * 16], tl[1024 * 4];
  build_fixed_tab(cub, 4, tc);
  build_fixed_tab(lin, 2, tl);
  cudaError_t e = cudaMemcpyToSymbol(g_cubic_i, tc, sizeof(tc));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_linear_i, tl, sizeof(tl));
  if (e != cudaSuccess) {
    set_error("feeder: weight table upload failed: %s", cudaGetErrorString(e));
    return -2;
  }
  done[dev & 63] = 1;
  return 0;
}

}  // namespace cris

using namespace cris;

extern "C" {

int cris_feeder_sample_bytes(void) { return (int)sizeof(FeedSample); }

int cris_feeder_letterbox(const uint8_t* images, const uint8_t* masks, const void* samples_dev, int B, int OH, int OW,
                          const double* border3, const float* mean3, const float* std3, float* img_out, float* mask_out,
                          void* stream) {
  CRIS_CHECK_ARG(images && samples_dev && border3 && mean3 && std3 && img_out && B >= 1 && B <= 65535 && OH >= 1 && OW >= 1 &&
                     (long long)OH * OW < (1ll << 30),
                 "feeder_letterbox: bad argument");
  CRIS_CHECK_ARG(mask_out == nullptr || masks != nullptr, "feeder_letterbox: mask_out without masks");
  if (int rc = upload_feeder_tabs()) return rc;
  FeedConst k;
  for (int c = 0; c < 3; ++c) {
    double r = nearbyint(border3[c]);   // saturate_cast<uchar>(double): round half to even, saturate
    k.cval[c] = (int)(r < 0 ? 0 : (r > 255 ? 255 : r));
    k.mean[c] = mean3[c];
    k.stdv[c] = std3[c];
  }
  dim3 grid((unsigned)(((long long)OH * OW + 255) / 256), (unsigned)B);
  feeder_letterbox_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      images, masks, reinterpret_cast<const FeedSample*>(samples_dev), OH, OW, k, img_out, mask_out);
  CRIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
