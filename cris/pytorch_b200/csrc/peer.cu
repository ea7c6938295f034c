// peer.cu — cross-GPU exchange of BatchNorm statistics over NVLink peer memory (one kernel, no NCCL).
//
// Replaces the collective inside torch.nn.SyncBatchNorm (reference: train.py:97-98 converts every BatchNorm; the
// forward all-gathers (mean, invstd, count), the backward all-reduces (sum_dy, sum_dy_xmu)).  Here each rank owns
// one cudaMalloc'ed buffer that every peer maps through CUDA IPC:
//
//   [ counters u32[SLOTS] | flags u32[SLOTS][8] | data f32[2 parities][SLOTS][SLOT_FLOATS] ]
//
// One exchange = one slot: the rank stores its vector into its own data slot, release-stores the slot's epoch into
// every peer's flag word, spins (acquire) until every peer's epoch arrived in its own flag words, then sums all
// ranks' vectors in rank order (bit-identical result on every rank).  The epoch is a per-slot use counter kept in
// device memory, so the kernel has no host-side state and can be replayed from a CUDA graph.  Data slots are
// double-buffered on the epoch's parity: a rank can only reach use e+2 of a slot after every peer signalled use
// e+1, i.e. after every peer finished reading use e.
#include <string.h>

#include "common.cuh"

namespace cris {

constexpr int kPeerSlots = CRIS_PEER_MAX_SLOTS;
constexpr int kPeerWorld = CRIS_PEER_MAX_WORLD;
constexpr int kSlotFloats = CRIS_PEER_SLOT_FLOATS;
constexpr size_t kCntBytes = (size_t)kPeerSlots * 4;
constexpr size_t kFlagBytes = (size_t)kPeerSlots * kPeerWorld * 4;
constexpr size_t kHdrBytes = kCntBytes + kFlagBytes;
constexpr size_t kDataBytes = (size_t)2 * kPeerSlots * kSlotFloats * 4;
// second region (fused SyncBatchNorm exchange, "push" protocol): a ring of kRing sites, per site one counter, one flag
// word per peer and one data vector PER SOURCE RANK (peers store straight into the reader's memory):
//   [ cnt u32[kRing] | flags u32[kRing][8] | data f32[kRing][8][SLOT_FLOATS] ]
constexpr int kRing = 32;
constexpr size_t kRingOff = kHdrBytes + kDataBytes;
constexpr size_t kRingCnt = (size_t)kRing * 4;
constexpr size_t kRingFlag = (size_t)kRing * kPeerWorld * 4;
constexpr size_t kRingData = (size_t)kRing * kPeerWorld * kSlotFloats * 4;
constexpr size_t kTotalBytes = kRingOff + kRingCnt + kRingFlag + kRingData;

struct PeerPtrs {
  unsigned char* p[kPeerWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerPtrs pp, int world, int rank, int slot,
                                                             const float* in, float* out,
                                                             int n, long long timeout_cycles) {
  unsigned char* me = pp.p[rank];
  unsigned* cnt = reinterpret_cast<unsigned*>(me) + slot;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = *cnt + 1u;
  __syncthreads();
  const unsigned e = s_epoch;
  const size_t data_off = ((size_t)(e & 1u) * kPeerSlots + (size_t)slot) * kSlotFloats;
  float* mine = reinterpret_cast<float*>(me + kHdrBytes) + data_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = in[i];
  __threadfence_system();
  __syncthreads();
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    // announce: my slot data for epoch e is visible
    unsigned* theirs = reinterpret_cast<unsigned*>(pp.p[t] + kCntBytes) + (size_t)slot * kPeerWorld + rank;
    st_release_sys(theirs, e);
    // wait for the peer's announcement in my own memory
    const unsigned* f = reinterpret_cast<const unsigned*>(me + kCntBytes) + (size_t)slot * kPeerWorld + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        printf("cris peer exchange: rank %d timed out waiting for rank %d (slot %d, epoch %u)\n", rank, t, slot, e);
        __trap();
      }
      __nanosleep(64);
    }
  }
  if (t == 0) *cnt = e;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      const float* src = reinterpret_cast<const float*>(pp.p[r] + kHdrBytes) + data_off;
      s += (r == rank) ? mine[i] : ld_relaxed_sys(src + i);
    }
    out[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One SyncBatchNorm exchange site in ONE kernel (was: reduce_partials -> peer_allreduce -> bn_coeffs, three launches
// that each rank-synchronise through the graph):
//   1. partials[n_tiles][2][C] -> this rank's sums (shared memory)
//   2. PUSH: every thread stores its elements straight into each peer's ring slot [site][my rank] over NVLink
//      (posted remote stores: one one-way trip; the old protocol signalled first and then loaded remotely = a flag
//      trip plus a load round trip per peer), one system fence + release flag per peer, spin on the own flag words
//   3. sum the `world` vectors in rank order from LOCAL memory (bit-identical on every rank)
//   4. forward : scale/shift/mean/invstd + running statistics from the GLOBAL sums (bn_coeffs_kernel's arithmetic)
//      backward: parameter gradients from the LOCAL sums (DDP averages them), GLOBAL sums for bn_bwd_apply
// A ring of kRing slots is enough: a rank can be at most one site ahead of its slowest peer (every site is a
// barrier), and a slot is rewritten kRing sites later.
// Reference: torch.nn.SyncBatchNorm forward/backward (train.py:97-98).
struct BnSyncArgs {
  PeerPtrs pp;
  int world, rank, slot;
  const float* partials; int n_tiles; int C;
  int backward;
  double count;                      // forward: elements per channel over ALL ranks
  const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var;
  float* scale; float* shift; float* mean; float* invstd;   // forward outputs
  float* sums_out; float* g0; float* g1;                      // backward outputs
  long long timeout_cycles;
};

__global__ void __launch_bounds__(1024) peer_bn_sync_kernel(const BnSyncArgs a) {
  __shared__ float s_loc[kSlotFloats];
  __shared__ unsigned s_epoch;
  const int n = 2 * a.C;
  const int slot = a.slot % kRing;
  unsigned char* me = a.pp.p[a.rank];
  unsigned* cnt = reinterpret_cast<unsigned*>(me + kRingOff) + slot;
  if (threadIdx.x == 0) s_epoch = *cnt + 1u;
  // 1. local reduction of the partial rows
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int t = 0;
    for (; t + 3 < a.n_tiles; t += 4) {
      s0 += a.partials[(size_t)t * n + i];
      s1 += a.partials[(size_t)(t + 1) * n + i];
      s2 += a.partials[(size_t)(t + 2) * n + i];
      s3 += a.partials[(size_t)(t + 3) * n + i];
    }
    for (; t < a.n_tiles; ++t) s0 += a.partials[(size_t)t * n + i];
    s_loc[i] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  const unsigned e = s_epoch;
  if (a.backward) {  // parameter gradients are the LOCAL sums
    for (int i = threadIdx.x; i < a.C; i += blockDim.x) {
      if (a.g0) a.g0[i] = s_loc[i];
      if (a.g1) a.g1[i] = s_loc[a.C + i];
    }
  }
  // 2. push my vector into every peer's slot [site][my rank]
  for (int r = 0; r < a.world; ++r) {
    if (r == a.rank) continue;
    float* dst = reinterpret_cast<float*>(a.pp.p[r] + kRingOff + kRingCnt + kRingFlag) +
                 ((size_t)slot * kPeerWorld + a.rank) * kSlotFloats;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = s_loc[i];
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < a.world && t != a.rank) {
    __threadfence_system();  // the CTA's remote stores (ordered before this point by the barrier) become visible ...
    unsigned* theirs = reinterpret_cast<unsigned*>(a.pp.p[t] + kRingOff + kRingCnt) + (size_t)slot * kPeerWorld + a.rank;
    st_release_sys(theirs, e);  // ... before the flag
    const unsigned* f = reinterpret_cast<const unsigned*>(me + kRingOff + kRingCnt) + (size_t)slot * kPeerWorld + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > a.timeout_cycles) {
        printf("cris peer bn sync: rank %d timed out waiting for rank %d (site %d, epoch %u)\n", a.rank, t, a.slot, e);
        __trap();
      }
    }
  }
  if (t == 0) *cnt = e;
  __syncthreads();
  // 3. + 4.
  const float* mine = reinterpret_cast<const float*>(me + kRingOff + kRingCnt + kRingFlag) + (size_t)slot * kPeerWorld * kSlotFloats;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float g0s = 0.f, g1s = 0.f;
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) { g0s += s_loc[c]; g1s += s_loc[a.C + c]; }
      else {
        g0s += ld_relaxed_sys(mine + (size_t)r * kSlotFloats + c);
        g1s += ld_relaxed_sys(mine + (size_t)r * kSlotFloats + a.C + c);
      }
    }
    if (a.backward) {
      a.sums_out[c] = g0s;
      a.sums_out[a.C + c] = g1s;
    } else {
      const double m = (double)g0s / a.count;
      double v = (double)g1s / a.count - m * m;
      if (v < 0) v = 0;
      const float meanf = (float)m, varf = (float)v;
      if (a.running_mean != nullptr) {
        const double unb = a.count > 1 ? v * a.count / (a.count - 1) : v;
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * meanf;
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
      }
      const float inv = rsqrtf(varf + a.eps);
      const float sc = a.gamma[c] * inv;
      a.scale[c] = sc;
      a.shift[c] = a.beta[c] - meanf * sc;
      a.mean[c] = meanf;
      a.invstd[c] = inv;
    }
  }
}

}  // namespace cris

using namespace cris;

extern "C" {

size_t cris_peer_buffer_bytes(void) { return kTotalBytes; }

int cris_peer_buffer_create(void** dev_ptr, unsigned char* handle_out) {
  CRIS_CHECK_ARG(dev_ptr && handle_out, "cris_peer_buffer_create: null argument");
  void* p = nullptr;
  CRIS_CUDA_OK(cudaMalloc(&p, kTotalBytes));
  CRIS_CUDA_OK(cudaMemset(p, 0, kTotalBytes));
  CRIS_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return -2;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == CRIS_PEER_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

int cris_peer_buffer_open(const unsigned char* handle, void** dev_ptr) {
  CRIS_CHECK_ARG(handle && dev_ptr, "cris_peer_buffer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CRIS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return 0;
}

int cris_peer_buffer_close(void* dev_ptr, int owned) {
  if (!dev_ptr) return 0;
  if (owned) {
    CRIS_CUDA_OK(cudaFree(dev_ptr));
  } else {
    CRIS_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  }
  return 0;
}

int cris_peer_allreduce_f32(void* const* peer_ptrs, int world, int rank, int slot, const float* in, float* out, int n,
                            double timeout_s, void* stream) {
  CRIS_CHECK_ARG(peer_ptrs && in && out, "cris_peer_allreduce_f32: null argument");
  CRIS_CHECK_ARG(world >= 2 && world <= kPeerWorld, "cris_peer_allreduce_f32: world %d outside 2..%d", world, kPeerWorld);
  CRIS_CHECK_ARG(rank >= 0 && rank < world, "cris_peer_allreduce_f32: rank %d of %d", rank, world);
  CRIS_CHECK_ARG(slot >= 0 && slot < kPeerSlots, "cris_peer_allreduce_f32: slot %d outside 0..%d", slot, kPeerSlots - 1);
  CRIS_CHECK_ARG(n >= 1 && n <= kSlotFloats, "cris_peer_allreduce_f32: n %d outside 1..%d", n, kSlotFloats);
  PeerPtrs pp;
  for (int r = 0; r < kPeerWorld; ++r) pp.p[r] = r < world ? static_cast<unsigned char*>(peer_ptrs[r]) : nullptr;
  for (int r = 0; r < world; ++r) CRIS_CHECK_ARG(pp.p[r], "cris_peer_allreduce_f32: peer %d not mapped", r);
  if (timeout_s <= 0) timeout_s = 120.0;
  const long long cycles = (long long)(timeout_s * 1.9e9);
  peer_allreduce_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(pp, world, rank, slot, in, out, n, cycles);
  CRIS_LAUNCH_OK();
  return 0;
}


static int fill_peers(PeerPtrs& pp, void* const* peer_ptrs, int world, int rank, const char* who) {
  CRIS_CHECK_ARG(peer_ptrs != nullptr, "%s: null peer table", who);
  CRIS_CHECK_ARG(world >= 2 && world <= kPeerWorld, "%s: world %d outside 2..%d", who, world, kPeerWorld);
  CRIS_CHECK_ARG(rank >= 0 && rank < world, "%s: rank %d of %d", who, rank, world);
  for (int r = 0; r < kPeerWorld; ++r) pp.p[r] = r < world ? static_cast<unsigned char*>(peer_ptrs[r]) : nullptr;
  for (int r = 0; r < world; ++r) CRIS_CHECK_ARG(pp.p[r], "%s: peer %d not mapped", who, r);
  return 0;
}

int cris_peer_bn_sync_fwd(void* const* peer_ptrs, int world, int rank, int site, const float* partials, int n_tiles, int C,
                          double global_count, const float* gamma, const float* beta, float eps, float momentum,
                          float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                          double timeout_s, void* stream) {
  BnSyncArgs a{};
  if (int rc = fill_peers(a.pp, peer_ptrs, world, rank, "cris_peer_bn_sync_fwd")) return rc;
  CRIS_CHECK_ARG(partials && gamma && beta && scale && shift && mean && invstd, "cris_peer_bn_sync_fwd: null argument");
  CRIS_CHECK_ARG(C >= 1 && 2 * C <= kSlotFloats && n_tiles >= 1 && site >= 0, "cris_peer_bn_sync_fwd: C=%d n_tiles=%d", C, n_tiles);
  a.world = world; a.rank = rank; a.slot = site; a.partials = partials; a.n_tiles = n_tiles; a.C = C; a.backward = 0;
  a.count = global_count; a.gamma = gamma; a.beta = beta; a.eps = eps; a.momentum = momentum;
  a.running_mean = running_mean; a.running_var = running_var;
  a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd;
  a.timeout_cycles = (long long)((timeout_s > 0 ? timeout_s : 120.0) * 1.9e9);
  peer_bn_sync_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a);
  CRIS_LAUNCH_OK();
  return 0;
}

int cris_peer_bn_sync_bwd(void* const* peer_ptrs, int world, int rank, int site, const float* partials, int n_tiles, int C,
                          float* sums_out, float* grad_beta, float* grad_gamma, double timeout_s, void* stream) {
  BnSyncArgs a{};
  if (int rc = fill_peers(a.pp, peer_ptrs, world, rank, "cris_peer_bn_sync_bwd")) return rc;
  CRIS_CHECK_ARG(partials && sums_out, "cris_peer_bn_sync_bwd: null argument");
  CRIS_CHECK_ARG(C >= 1 && 2 * C <= kSlotFloats && n_tiles >= 1 && site >= 0, "cris_peer_bn_sync_bwd: C=%d n_tiles=%d", C, n_tiles);
  a.world = world; a.rank = rank; a.slot = site; a.partials = partials; a.n_tiles = n_tiles; a.C = C; a.backward = 1;
  a.sums_out = sums_out; a.g0 = grad_beta; a.g1 = grad_gamma;
  a.timeout_cycles = (long long)((timeout_s > 0 ? timeout_s : 120.0) * 1.9e9);
  peer_bn_sync_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a);
  CRIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
