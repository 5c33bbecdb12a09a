// peer.cu — one-shot all-reduce of a short vector over NVLink peer memory, in ONE kernel and without NCCL: the
// exchange SynchronizedBatchNorm2d needs per layer (forward [sum x | sum x^2 | n], backward [sum dy | sum dy*xhat];
// reference modeling/sync_batchnorm/batchnorm.py:90-111 does ReduceAddCoalesced + Broadcast between DataParallel threads).
//
// Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into every
// peer's address space) with two slots of `cap` bytes and a signal pad of one 32-bit word per peer.  Call number
// `epoch` (the same on every rank: the ranks run the same layer sequence):
//   1. copy the local vector into MY slot epoch & 1, make it visible system-wide (fence),
//   2. store `epoch` into word [my rank] of every peer's signal pad (st.release.sys over NVLink),
//   3. wait until all words of MY signal pad have reached `epoch` (ld.acquire.sys),
//   4. every element = the sum over the ranks' slots in RANK ORDER (P2P loads): the same fixed order on every rank, so
//      all ranks hold bit-identical results (deterministic; an NCCL ring does not promise that across rank counts).
// Two slots suffice: a rank can only start call e+2 after every peer has signalled e+1, i.e. has finished reading
// the slot of call e.  The wait is bounded (~2 s of polling): a missing peer makes the call fail, not hang.
// One block; vectors are a few KB (2C+1 floats).  Different GPUs only — ranks sharing one GPU must not use this.
#include "common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int PEER_MAX_WORLD = 16;
struct PeerParams {
  void* bufs[PEER_MAX_WORLD];          // symmetric buffers of all ranks (peer-mapped addresses)
  uint32_t* signals[PEER_MAX_WORLD];   // signal pads of all ranks
  int rank, world; uint32_t epoch; long long cap_bytes;
};

template <typename T>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(T* __restrict__ vec, int n, PeerParams p, uint32_t* __restrict__ epoch_dev, int* __restrict__ status) {
  __shared__ int failed;
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    failed = 0;
    // the call counter lives on the DEVICE when epoch_dev is given, so a CUDA graph that contains this launch can be
    // replayed: every replay is a new call (a host-side counter would be frozen into the graph)
    if (epoch_dev) { epoch_s = *epoch_dev + 1u; *epoch_dev = epoch_s; } else epoch_s = p.epoch;
  }
  __syncthreads();
  p.epoch = epoch_s;
  const int slot = (int)(p.epoch & 1u);
  T* mine = reinterpret_cast<T*>(static_cast<char*>(p.bufs[p.rank]) + (size_t)slot * p.cap_bytes);
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = vec[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < p.world) st_release_sys(p.signals[threadIdx.x] + p.rank, p.epoch);
  if (threadIdx.x < p.world) {
    const uint32_t* flag = p.signals[p.rank] + threadIdx.x;
    long long spins = 0;
    while ((int)(ld_acquire_sys(flag) - p.epoch) < 0) {
      __nanosleep(64);
      if (++spins > 20000000ll) { failed = 1; break; }      // ~2 s: a peer never arrived
    }
  }
  __syncthreads();
  if (failed) { if (threadIdx.x == 0 && status) *status = 1; return; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T s = 0;
    for (int r = 0; r < p.world; ++r)
      s += reinterpret_cast<const volatile T*>(static_cast<const char*>(p.bufs[r]) + (size_t)slot * p.cap_bytes)[i];
    vec[i] = s;
  }
}

}  // namespace

/* vec (device, fp32 if elem_bytes == 4 else fp64, n elements) <- sum over the ranks, in place.  buf_ptrs / signal_ptrs:
 * HOST arrays of `world` peer-mapped addresses (symmetric buffer with 2 slots of cap_bytes each; signal pad of >= world
 * uint32 words, zero-initialised).  epoch: 1, 2, 3, ... — the same on every rank; or epoch_dev: a device counter (starts at
 * 0) that the kernel increments itself, which makes the launch replayable inside a CUDA graph.  status_dev: device int set
 * to 1 if a peer did not arrive within ~2 s (or NULL). */
extern "C" int add_peer_allreduce(void* vec, int n, int elem_bytes, const uint64_t* buf_ptrs, const uint64_t* signal_ptrs, int rank,
                                  int world, uint32_t epoch, uint32_t* epoch_dev, int64_t cap_bytes, int* status_dev, void* stream) {
  ADD_CHECK_ARG(vec && n > 0 && buf_ptrs && signal_ptrs && world >= 1 && rank >= 0 && rank < world && (epoch > 0 || epoch_dev));
  ADD_CHECK_SUP(world <= PEER_MAX_WORLD && (elem_bytes == 4 || elem_bytes == 8) && (int64_t)n * elem_bytes <= cap_bytes);
  PeerParams p;
  for (int r = 0; r < world; ++r) { p.bufs[r] = (void*)buf_ptrs[r]; p.signals[r] = (uint32_t*)signal_ptrs[r]; }
  p.rank = rank; p.world = world; p.epoch = epoch; p.cap_bytes = cap_bytes;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 4) peer_allreduce_kernel<float><<<1, 256, 0, s>>>((float*)vec, n, p, epoch_dev, status_dev);
  else peer_allreduce_kernel<double><<<1, 256, 0, s>>>((double*)vec, n, p, epoch_dev, status_dev);
  ADD_RETURN_LAUNCH();
}
