// C ABI of the host TensorPool (common/tensor_pool.h): 64-byte aligned host memory, optionally page-locked by the
// caller (the python wrapper registers slabs with cudaHostRegister when a GPU is present).
#include <cstdlib>

#include "../common/tensor_pool.h"

namespace {
void* HostAlloc(size_t n, void*) { void* p = nullptr; return posix_memalign(&p, 64, (n + 63) & ~size_t(63)) == 0 ? p : nullptr; }
void HostFree(void* p, void*) { free(p); }
}  // namespace

extern "C" {

void* dr_tp_create(int64_t small_threshold, int collect_steps, int replan_misses) {
  return new dr::TensorPool(HostAlloc, HostFree, nullptr, (size_t)small_threshold, collect_steps, replan_misses);
}
void dr_tp_destroy(void* h) { delete static_cast<dr::TensorPool*>(h); }
void dr_tp_set_start_step(void* h, int s) { static_cast<dr::TensorPool*>(h)->SetStartStep(s); }
void* dr_tp_alloc(void* h, int64_t bytes, uint64_t stream) { return static_cast<dr::TensorPool*>(h)->Alloc((size_t)bytes, stream); }
void dr_tp_free(void* h, void* p) { static_cast<dr::TensorPool*>(h)->Free(p); }
void dr_tp_step_end(void* h) { static_cast<dr::TensorPool*>(h)->StepEnd(); }
void dr_tp_stats(void* h, int64_t* out9) {
  const dr::TensorPoolStats s = static_cast<dr::TensorPool*>(h)->Stats();
  const int64_t v[9] = {s.phase, s.steps, s.pool_bytes, s.pool_hits, s.pool_misses, s.small_bypass, s.backend_allocs, s.live_pool_blocks, s.replans};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
}
int64_t dr_tp_class_bytes(int64_t bytes) { return (int64_t)dr::TensorPool::ClassBytes(dr::TensorPool::ClassOf((size_t)bytes)); }

}  // extern "C"
