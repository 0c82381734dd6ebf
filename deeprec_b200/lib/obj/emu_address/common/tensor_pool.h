// TensorPool: a planned pool allocator for the per-step temporaries of a training / serving loop.
//
// Reference behaviour (common_runtime/{memory_planner,tensorpool_allocator,gpu_memory_planner,gpu_tensorpool_allocator}.*,
// docs/docs_en/{CPU,GPU}-Memory-Optimization.md; TF_GPU_ALLOCATOR=tensorpool, START/STABLE/MAX_STATISTIC_STEP):
// allocation sizes are recorded for the first steps, a plan (how many buffers of which size class are live at once) is
// derived from them, and from then on allocations are served from pre-carved buffers; requests the plan does not cover go
// to the underlying allocator.  Small allocations bypass the pool.
//
// Design here:
//   * size classes: powers of two subdivided 4x (<= 19 % internal fragmentation), from `small_threshold` upward;
//   * COLLECT phase (`collect_steps` steps): pass-through to the backend, per class track live / peak-live counts;
//   * PLAN: one backend slab per class = peak_live blocks, carved into a lock-protected LIFO free list;
//   * SERVE: O(1) pop / push; a miss falls through to the backend and is counted; when misses exceed `replan_misses`
//     the pool re-collects for one step and grows the affected classes (STABLE_STATISTIC_STEP behaviour);
//   * `stream` tags (GPU): a block freed on stream S is only handed out again on S, so reuse never needs an event.
// The backend is two function pointers, so the same code serves malloc'ed host memory (tested on CPU), pinned host memory
// and cudaMalloc'ed device memory (csrc/cuda/allocator.cu, plugged into PyTorch via CUDAPluggableAllocator).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace dr {

struct TensorPoolStats {
  int64_t phase;            // 0 collect, 1 serve
  int64_t steps;
  int64_t pool_bytes;       // bytes held in slabs
  int64_t pool_hits, pool_misses, small_bypass;
  int64_t backend_allocs;   // calls that reached the backend (collect + misses + small)
  int64_t live_pool_blocks;
  int64_t replans;
};

class TensorPool {
 public:
  using AllocFn = void* (*)(size_t bytes, void* ctx);
  using FreeFn = void (*)(void* p, void* ctx);

  TensorPool(AllocFn a, FreeFn f, void* ctx, size_t small_threshold, int collect_steps, int replan_misses)
      : alloc_(a), free_(f), ctx_(ctx), small_(std::max<size_t>(small_threshold, 256)), collect_steps_(std::max(1, collect_steps)),
        replan_misses_(std::max(1, replan_misses)) {}

  ~TensorPool() { for (void* s : slabs_) free_(s, ctx_); }

  // START_STATISTIC_STEP: the first `s` steps (graph initialisation, warm-up) are passed through and NOT counted into the plan; the
  // collect phase is steps [s, s + collect_steps).
  void SetStartStep(int s) { std::lock_guard<std::mutex> l(mu_); start_step_ = std::max(0, s); }

  static int ClassOf(size_t bytes) {          // 4 sub-classes per power of two
    if (bytes <= 256) return 3;                                            // ClassBytes(3) == 256
    const size_t b = bytes - 1;
    const int hi = 63 - __builtin_clzll((unsigned long long)b);            // floor(log2(bytes-1))
    const int sub = (int)((b >> (hi - 2)) & 3);
    return (hi - 7) * 4 + sub;                                             // 257..320 -> class 4 (320 B), 321..384 -> class 5, ...
  }
  static size_t ClassBytes(int c) {
    const int hi = c / 4 + 7, sub = c % 4;
    return ((size_t)(4 + sub + 1)) << (hi - 2);
  }

  void* Alloc(size_t bytes, uint64_t stream = 0) {
    if (bytes == 0) return nullptr;
    std::lock_guard<std::mutex> l(mu_);
    if (bytes < small_) { ++st_.small_bypass; return Backend(bytes); }
    const int c = ClassOf(bytes);
    if ((int)cls_.size() <= c) cls_.resize(c + 1);
    Class& k = cls_[c];
    if (++k.live > k.peak) k.peak = k.live;
    if (st_.phase == 1) {
      auto& fl = k.free[stream];
      if (!fl.empty()) {
        void* p = fl.back(); fl.pop_back();
        owner_[p] = {c, stream, true};
        ++st_.pool_hits; ++st_.live_pool_blocks;
        return p;
      }
      if (!k.unassigned.empty()) {       // carved but never used on any stream yet
        void* p = k.unassigned.back(); k.unassigned.pop_back();
        owner_[p] = {c, stream, true};
        ++st_.pool_hits; ++st_.live_pool_blocks;
        return p;
      }
      ++st_.pool_misses; ++misses_since_plan_;
    }
    void* p = Backend(ClassBytes(c));
    if (p) owner_[p] = {c, stream, false};
    return p;
  }

  void Free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> l(mu_);
    auto it = owner_.find(p);
    if (it == owner_.end()) { free_(p, ctx_); return; }       // small bypass
    const Owner o = it->second;
    owner_.erase(it);
    Class& k = cls_[o.cls];
    --k.live;
    if (o.pooled) { k.free[o.stream].push_back(p); --st_.live_pool_blocks; }
    else free_(p, ctx_);
  }

  // Step boundary: ends the collect phase after `collect_steps`, re-plans when the plan proved too small.
  void StepEnd() {
    std::lock_guard<std::mutex> l(mu_);
    ++st_.steps;
    if (st_.phase == 0 && start_step_ > 0 && st_.steps == start_step_) for (auto& k : cls_) k.peak = k.live;     // statistics start here
    if (st_.phase == 0 && st_.steps >= start_step_ + collect_steps_) { Plan(); st_.phase = 1; }
    else if (st_.phase == 1 && misses_since_plan_ >= replan_misses_) { Plan(); ++st_.replans; }
    if (st_.phase == 1) for (auto& k : cls_) k.peak = k.live;      // collect: max over the whole phase; serve: per step window
  }

  TensorPoolStats Stats() { std::lock_guard<std::mutex> l(mu_); return st_; }

 private:
  struct Class {
    int64_t live = 0, peak = 0, planned = 0;
    std::unordered_map<uint64_t, std::vector<void*>> free;    // per stream tag
    std::vector<void*> unassigned;
  };
  struct Owner { int cls; uint64_t stream; bool pooled; };

  void* Backend(size_t bytes) { ++st_.backend_allocs; return alloc_(bytes, ctx_); }

  void Plan() {                                // grow every class to its observed peak-live count
    for (size_t c = 0; c < cls_.size(); ++c) {
      Class& k = cls_[c];
      const int64_t want = std::max(k.peak, k.live), add = want - k.planned;
      if (add <= 0) continue;
      const size_t cb = ClassBytes((int)c);
      char* slab = static_cast<char*>(alloc_(cb * (size_t)add, ctx_));
      if (!slab) continue;
      slabs_.push_back(slab);
      st_.pool_bytes += (int64_t)(cb * (size_t)add);
      for (int64_t i = 0; i < add; ++i) k.unassigned.push_back(slab + (size_t)i * cb);
      k.planned = want;
    }
    misses_since_plan_ = 0;
  }

  AllocFn alloc_; FreeFn free_; void* ctx_;
  size_t small_; int collect_steps_, replan_misses_, start_step_ = 0;
  std::mutex mu_;
  std::vector<Class> cls_;
  std::unordered_map<void*, Owner> owner_;
  std::vector<void*> slabs_;
  TensorPoolStats st_{};
  int64_t misses_since_plan_ = 0;
};

}  // namespace dr
