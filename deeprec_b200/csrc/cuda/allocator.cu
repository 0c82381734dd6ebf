// GPU TensorPool allocator (reference: TF_GPU_ALLOCATOR=tensorpool, common_runtime/gpu_tensorpool_allocator.*): the planned pool
// of common/tensor_pool.h over cudaMalloc, exported with the signatures torch.cuda.memory.CUDAPluggableAllocator expects.
// One pool per device; blocks are tagged with the stream they were last used on, so reuse never crosses streams.
//
//   alloc = torch.cuda.memory.CUDAPluggableAllocator(libdeeprec_cuda.so, "dr_tp_cuda_malloc", "dr_tp_cuda_free")
//   torch.cuda.memory.change_current_allocator(alloc)        # before the first CUDA allocation
//   ... each step: lib.dr_tp_cuda_step_end(device)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "../common/tensor_pool.h"

namespace {
constexpr int kMaxDev = 16;
dr::TensorPool* g_pool[kMaxDev] = {};
std::mutex g_mu;
int64_t g_small = 32 << 10;          // reference: allocations below 32 KB bypass the pool
int g_collect = 3, g_replan = 8;

void* DevAlloc(size_t n, void* ctx) {
  int prev = 0; cudaGetDevice(&prev);
  const int dev = (int)(intptr_t)ctx;
  if (prev != dev) cudaSetDevice(dev);
  void* p = nullptr;
  if (cudaMalloc(&p, n) != cudaSuccess) {
    cudaGetLastError(); p = nullptr;
    // GPU virtual memory (docs/docs_en/GPU-Virtual-Memory.md, TF_GPU_VMEM): when device memory is exhausted fall back to managed memory,
    // which the driver pages between host and device -- slower, but the job keeps running instead of failing with OOM
    static const bool vmem = [] { const char* e = getenv("DEEPREC_GPU_VMEM"); if (!e) e = getenv("TF_GPU_VMEM"); return e && (e[0] == '1' || e[0] == 't' || e[0] == 'T'); }();
    if (vmem && cudaMallocManaged(&p, n) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
  }
  if (prev != dev) cudaSetDevice(prev);
  return p;
}
void DevFree(void* p, void*) { cudaFree(p); }

dr::TensorPool* Pool(int dev) {
  if (dev < 0 || dev >= kMaxDev) return nullptr;
  std::lock_guard<std::mutex> l(g_mu);
  if (!g_pool[dev]) {
    if (const char* e = getenv("DEEPREC_TENSORPOOL_SMALL_BYTES")) g_small = atoll(e);
    // GPU-Memory-Optimization.md: statistics run over steps [START_STATISTIC_STEP, STOP_STATISTIC_STEP); STABLE_STATISTIC_STEP (the CPU
    // side's name for the length of the collection window) is accepted too
    int start = 0;
    if (const char* e = getenv("STABLE_STATISTIC_STEP")) g_collect = atoi(e) > 0 ? atoi(e) : g_collect;
    if (const char* e = getenv("START_STATISTIC_STEP")) start = atoi(e) > 0 ? atoi(e) : 0;
    if (const char* e = getenv("STOP_STATISTIC_STEP")) { const int stop = atoi(e); if (stop > start) g_collect = stop - start; }
    g_pool[dev] = new dr::TensorPool(DevAlloc, DevFree, (void*)(intptr_t)dev, (size_t)g_small, g_collect, g_replan);
    g_pool[dev]->SetStartStep(start);
  }
  return g_pool[dev];
}
}  // namespace

extern "C" {

void* dr_tp_cuda_malloc(ssize_t size, int device, cudaStream_t stream) {
  dr::TensorPool* p = Pool(device);
  return p ? p->Alloc((size_t)size, (uint64_t)(uintptr_t)stream) : nullptr;
}

void dr_tp_cuda_free(void* ptr, ssize_t /*size*/, int device, cudaStream_t /*stream*/) {
  if (dr::TensorPool* p = Pool(device)) p->Free(ptr);
}

void dr_tp_cuda_step_end(int device) { if (dr::TensorPool* p = Pool(device)) p->StepEnd(); }

void dr_tp_cuda_stats(int device, int64_t* out9) {
  dr::TensorPool* p = Pool(device);
  if (!p) return;
  const dr::TensorPoolStats s = p->Stats();
  const int64_t v[9] = {s.phase, s.steps, s.pool_bytes, s.pool_hits, s.pool_misses, s.small_bypass, s.backend_allocs, s.live_pool_blocks, s.replans};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
}

}  // extern "C"
