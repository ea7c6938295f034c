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

}  // namespace cris

using namespace cris;

extern "C" {

size_t cris_peer_buffer_bytes(void) { return kHdrBytes + kDataBytes; }

int cris_peer_buffer_create(void** dev_ptr, unsigned char* handle_out) {
  CRIS_CHECK_ARG(dev_ptr && handle_out, "cris_peer_buffer_create: null argument");
  void* p = nullptr;
  CRIS_CUDA_OK(cudaMalloc(&p, kHdrBytes + kDataBytes));
  CRIS_CUDA_OK(cudaMemset(p, 0, kHdrBytes + kDataBytes));
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

}  // extern "C"
