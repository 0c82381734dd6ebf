// Cross-GPU signalling for the fused compute+collective kernels: release/acquire flags in peer-mapped (NVLink) memory,
// raised by the LAST block of a producer kernel and polled by the consumer kernel's own blocks -- no barrier kernels,
// no NCCL call, no host round trip (reference: SOK's all2all = ncclSend/Recv groups + D2H count exchange + stream sync,
// addons/sparse_operation_kit/legacy/kit_cc_impl/embedding/dispatcher/all2all_input_dispatcher.cu:248-285).
//
// Protocol.  Every rank keeps a monotonic step counter `state[0]` (bumped by k_sp_step_end once per forward / training
// step on every rank, so all ranks agree on it).  A producer on rank r that finished phase `ch` of step e stores
// e + 1 into flags[ch][r] OF EVERY PEER with st.release.sys; a consumer waits until flags[ch][src] >= e + 1 with
// ld.acquire.sys.  Flags are never reset (wrap-safe signed comparison), so a captured CUDA graph replays the protocol.
#pragma once
#include "common.cuh"

extern "C" {
struct DrPeers {
  void* ptr[16];     // ptr[r] = this symmetric buffer as mapped in the local address space for rank r
};
// flags: uint32 [kSpChannels][16] per rank (symmetric).  state: int32 [16] local: [0] step epoch, [8 + ch] block-done counters.
struct DrSpSync {
  DrPeers flags;
  int32_t* state;
  int32_t rank;
  int32_t W;
};
}

namespace drc {

constexpr int kSpChannels = 8;
enum { SP_CH_DEDUP = 0, SP_CH_ROWS = 1, SP_CH_GRAD = 2, SP_CH_DENSE = 3, SP_CH_AUX = 4 };

__device__ __forceinline__ uint32_t sp_epoch(const DrSpSync& s) { return (uint32_t)ld_volatile_i32(&s.state[0]) + 1u; }

// Called by ALL threads of EVERY block as the last thing the kernel does with the data being published.
__device__ __forceinline__ void sp_signal_last_block(const DrSpSync& s, int ch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                     // this block's writes (ordered by the barrier) before the count
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    const int prev = atomicAdd(&s.state[8 + ch], 1);
    if (prev == (int)total - 1) {
      s.state[8 + ch] = 0;
      __threadfence_system();                                   // acquire the other blocks' writes, then publish system-wide
      const uint32_t ep = sp_epoch(s);
      for (int r = 0; r < s.W; ++r)
        st_release_sys(reinterpret_cast<uint32_t*>(s.flags.ptr[r]) + ch * 16 + s.rank, ep);
    }
  }
}

// One-thread poll of one source's flag (caller brackets it with __syncthreads()).
__device__ __forceinline__ void sp_wait_one(const DrSpSync& s, int ch, int src) {
  const uint32_t ep = sp_epoch(s);
  const uint32_t* f = reinterpret_cast<const uint32_t*>(s.flags.ptr[s.rank]) + ch * 16 + src;
  while ((int32_t)(ld_acquire_sys(f) - ep) < 0) __nanosleep(40);
}

// All threads of the block call it: returns once every rank's flag of channel `ch` reached this step.
__device__ __forceinline__ void sp_wait_all(const DrSpSync& s, int ch) {
  if ((int)threadIdx.x < s.W) sp_wait_one(s, ch, threadIdx.x);
  __syncthreads();
}

}  // namespace drc
