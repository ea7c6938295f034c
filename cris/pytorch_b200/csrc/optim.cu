// optim.cu — multi-tensor Adam step in ONE kernel launch per parameter group (SURVEY §8f "next" row).
// Reference: torch.optim.Adam(param_list, lr, weight_decay) driven through GradScaler
// (train.py:105-111; engine/engine.py:52-57): unscale by 1/loss_scale, skip the step when a non-finite gradient was
// found, L2 weight decay folded into the gradient, bias-corrected moments.  HBM-bound: 28 B per parameter
// (p r/w, g r, m r/w, v r/w), float4 accesses, one pass.
#include "common.cuh"

namespace cris {

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;       // elements
  long long chunk0;  // index of this tensor's first chunk (exclusive prefix sum of chunk counts)
};

constexpr int kAdamChunk = 8192;  // elements per block

// omb1 / omb2 = (float)(1 - beta) evaluated in double on the host, as torch does: 1.f - 0.999f is off by 5e-5
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float inv_scale, float wd, float omb1,
                                            float b2, float omb2, float eps, float step_size, float bc2_sqrt) {
  g *= inv_scale;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(omb1, g - m, m);                     // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(omb2, g * g, b2 * v);                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = fmaf(-step_size, m / denom, p);           // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
    adam_kernel(const AdamTensor* __restrict__ tab, int n_tensors, float omb1, float b2, float omb2, float eps, float wd,
                float step_size, float bc2_sqrt, const float* __restrict__ grad_scale,
                const float* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: skip the whole step
  const float inv_scale = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
  // block -> tensor: binary search over the chunk prefix sums
  int lo = 0, hi = n_tensors - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk0 <= b) lo = mid; else hi = mid - 1;
  }
  const AdamTensor t = tab[lo];
  const long long e0 = (b - t.chunk0) * kAdamChunk;
  const long long e1 = min(t.n, e0 + kAdamChunk);
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
  if (vec) {
    const long long q1 = e0 + ((e1 - e0) & ~3ll);
    for (long long i = e0 + threadIdx.x * 4; i < q1; i += 256 * 4) {
      float4 p = *reinterpret_cast<float4*>(t.p + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<float4*>(t.m + i);
      float4 v = *reinterpret_cast<float4*>(t.v + i);
      adam_update(p.x, g.x, m.x, v.x, inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
      adam_update(p.y, g.y, m.y, v.y, inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
      adam_update(p.z, g.z, m.z, v.z, inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
      adam_update(p.w, g.w, m.w, v.w, inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
    }
    for (long long i = q1 + threadIdx.x; i < e1; i += 256)
      adam_update(t.p[i], t.g[i], t.m[i], t.v[i], inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
  } else {
    for (long long i = e0 + threadIdx.x; i < e1; i += 256)
      adam_update(t.p[i], t.g[i], t.m[i], t.v[i], inv_scale, wd, omb1, b2, omb2, eps, step_size, bc2_sqrt);
  }
}

}  // namespace cris

using namespace cris;

extern "C" {

int cris_adam_table_entry_bytes(void) { return (int)sizeof(AdamTensor); }
int cris_adam_chunk_elems(void) { return kAdamChunk; }

int cris_adam_step(const void* table_dev, int n_tensors, long long n_chunks, double lr, double beta1, double beta2,
                   double eps, double weight_decay, double step, const float* grad_scale, const float* found_inf,
                   void* stream) {
  CRIS_CHECK_ARG(table_dev != nullptr && n_tensors >= 1 && n_chunks >= 1 && n_chunks < (1ll << 31),
                 "cris_adam_step: bad table (%d tensors, %lld chunks)", n_tensors, n_chunks);
  CRIS_CHECK_ARG(step >= 1.0, "cris_adam_step: step %g must be >= 1 (the step count AFTER this update)", step);
  const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  adam_kernel<<<(unsigned)n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const AdamTensor*>(table_dev), n_tensors, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)eps, (float)weight_decay, step_size, bc2_sqrt,
      grad_scale, found_inf);
  CRIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
