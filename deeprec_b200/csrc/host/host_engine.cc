// Host-tier embedding engine (C++17, no CUDA): the CPU EmbeddingVariable implementation and
// the DRAM tier of the multi-tier store.  C ABI, loaded with ctypes (deeprec_b200/_native.py).
//
// Parity map (behaviour, not structure) to the reference:
//   EmbeddingVar::GetEmbeddings / LookupOrCreateKey   framework/embedding/embedding_var.h:142-219
//   LocklessHashMap / DenseHashMap KV                 framework/embedding/cpu_hash_map_kv.h:23-223, dense_hash_map_kv.h:27-160
//   CounterFilterPolicy / BloomFilterPolicy           framework/embedding/counter_filter_policy.h:35-189, bloom_filter_policy.h:33-450
//   GlobalStep / L2Weight shrink at save              framework/embedding/globalstep_shrink_policy.h:43-58, l2weight_shrink_policy.h:46-64
//   KvResourceSparseApply* (8 rules, WithCounts)      core/kernels/training_ali_ops.cc:73-3200
//   ckpt bucketing (key % 1000) + N->M re-shard       framework/embedding/storage.h:255-288, embedding_var_restore.cc:131
//   IndicesIncrRecorder (dirty keys)                  core/kernels/incr_save_restore_ops.h:347
//
// Design: partitioned open-addressing KV with wait-free reads (acquire loads) and CAS inserts;
// per-partition shared_mutex is taken shared by every op and exclusive only while a
// partition rehashes or is compacted by eviction.  Key metadata (freq, version, row index,
// dirty bit) is SoA in chunked arrays so pointers stay stable while the table grows; rows
// come from a chunked slab with a free list (the EVAllocator analogue: no per-row malloc).
#ifdef DR_USE_OPENMP
#include <omp.h>
#endif
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <vector>

#include "../common/ev_types.h"

namespace dr {

// ------------------------------------------------------------------------------------
// Thread pool (intra-op sharding; the reference uses TF's Shard() over worker threads)
// ------------------------------------------------------------------------------------
#ifdef DR_USE_OPENMP
// Production build: parallel loops run on the process's OpenMP runtime -- the SAME libgomp pool PyTorch's CPU kernels use.  A private
// std::thread pool next to it is starved by libgomp's spin-waiting workers between torch ops (measured: the native DLRM interaction
// took 12.8 ms inside a training step vs 3.4 ms stand-alone); sharing the runtime removes the oversubscription.
class ThreadPool {
 public:
  explicit ThreadPool(int n) : n_(n < 1 ? 1 : n) {}
  int size() const { return n_; }
  void ParallelFor(int64_t n, int64_t min_grain, const std::function<void(int64_t, int64_t)>& fn) {
    if (n <= 0) return;
    const int shards = (int)std::min<int64_t>(n_ + 1, (n + min_grain - 1) / min_grain);
    if (shards <= 1 || omp_in_parallel()) { fn(0, n); return; }
    const int64_t per = (n + shards - 1) / shards;
#pragma omp parallel for schedule(static, 1) num_threads(shards)
    for (int s = 0; s < shards; ++s) {
      const int64_t b = s * per, e = std::min(n, b + per);
      if (b < e) fn(b, e);
    }
  }
 private:
  int n_;
};
#else
// Sanitizer / stand-alone build: a private pool of std::threads (ThreadSanitizer cannot see into an uninstrumented libgomp).
class ThreadPool {
 public:
  explicit ThreadPool(int n) : stop_(false) {
    if (n < 1) n = 1;
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { Loop(); });
  }
  ~ThreadPool() {
    { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // Runs fn(begin,end) over [0,n) in roughly equal shards; caller participates.
  void ParallelFor(int64_t n, int64_t min_grain, const std::function<void(int64_t, int64_t)>& fn) {
    if (n <= 0) return;
    int shards = (int)std::min<int64_t>(size() + 1, (n + min_grain - 1) / min_grain);
    if (shards <= 1) { fn(0, n); return; }
    // the completion count is only touched under dmu: the waiter can then not observe zero (and destroy dmu / dcv, which live on its
    // stack) while the last worker is still between its decrement and its notify -- ThreadSanitizer caught exactly that window
    int remaining = shards - 1;
    std::mutex dmu; std::condition_variable dcv;
    int64_t per = (n + shards - 1) / shards;
    for (int s = 1; s < shards; ++s) {
      int64_t b = s * per, e = std::min(n, b + per);
      Submit([&, b, e] {
        if (b < e) fn(b, e);
        std::lock_guard<std::mutex> l(dmu);
        if (--remaining == 0) dcv.notify_one();
      });
    }
    fn(0, std::min(n, per));
    std::unique_lock<std::mutex> l(dmu);
    dcv.wait(l, [&] { return remaining == 0; });
  }
  void Submit(std::function<void()> f) {
    { std::lock_guard<std::mutex> l(mu_); q_.push_back(std::move(f)); }
    cv_.notify_one();
  }
 private:
  void Loop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        f = std::move(q_.front()); q_.erase(q_.begin());
      }
      f();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<std::function<void()>> q_;
  std::mutex mu_; std::condition_variable cv_; bool stop_;
};
#endif

static ThreadPool* GlobalPool() {
  static ThreadPool* p = [] {
    int n = (int)std::thread::hardware_concurrency();
#ifdef DR_USE_OPENMP
    n = omp_get_max_threads();                                 // follows OMP_NUM_THREADS / torch.set_num_threads at first use
#endif
    if (const char* e = getenv("DEEPREC_HOST_THREADS")) n = atoi(e);
    if (n > 64) n = 64;
    return new ThreadPool(std::max(1, n - 1));
  }();
  return p;
}

// ------------------------------------------------------------------------------------
// Chunked array: stable addresses under growth, lock-free reads.
// ------------------------------------------------------------------------------------
// Large blocks (>= 2 MiB) are 2 MiB-aligned and advised as transparent huge pages (the host runs THP in `madvise` mode on most
// distributions); smaller ones are plain 64-byte aligned allocations.  free() releases either.
static constexpr int64_t kHugePage = int64_t(2) << 20;
static inline void* HugeAlloc(size_t bytes) {
  void* p = nullptr;
  const bool huge = bytes >= (size_t)kHugePage;
  if (posix_memalign(&p, huge ? (size_t)kHugePage : 64, bytes) != 0) abort();
#ifdef MADV_HUGEPAGE
  static const bool thp = getenv("DEEPREC_HOST_THP") && atoi(getenv("DEEPREC_HOST_THP")) != 0;
  if (huge && thp) madvise(p, bytes, MADV_HUGEPAGE);
#endif
  return p;
}

template <typename T, int kLog2Chunk = 16>
class ChunkedArray {
 public:
  static constexpr int64_t kChunk = int64_t(1) << kLog2Chunk;
  static constexpr int64_t kMaxChunks = 1 << 16;
  explicit ChunkedArray(int64_t width = 1) : width_(width) {
    chunks_ = new std::atomic<T*>[kMaxChunks];
    for (int64_t i = 0; i < kMaxChunks; ++i) chunks_[i].store(nullptr, std::memory_order_relaxed);
  }
  ~ChunkedArray() {
    for (T* p : bases_) free(p);
    delete[] chunks_;
  }
  T* at(int64_t idx) {
    T* c = chunks_[idx >> kLog2Chunk].load(std::memory_order_acquire);
    return c + (idx & (kChunk - 1)) * width_;
  }
  // Entries are NOT initialised here: every user resets an entry when it hands it out (ResetMeta / InitRow), and a fill pass would touch
  // (page-fault) memory long before it is used.  The first kSmallChunks chunks are separate allocations (a 50-row table stays small);
  // after that chunks come in 2 MiB-aligned slabs advised as transparent huge pages -- creating rows then page-faults once per 2 MiB
  // instead of once per 4 KiB (the dominant cost of inserting fresh keys), and random row reads of a multi-GB table miss the TLB less.
  void EnsureCapacity(int64_t n, const T& /*fill*/) {
    int64_t need = (n + kChunk - 1) >> kLog2Chunk;
    if (need <= nchunks_.load(std::memory_order_acquire)) return;
    std::lock_guard<std::mutex> l(mu_);
    const int64_t chunk_elems = kChunk * width_;
    const int64_t chunk_bytes = (int64_t)sizeof(T) * chunk_elems;
    for (int64_t c = nchunks_.load(); c < need;) {
      int64_t group = 1;
      if (c >= kSmallChunks) { group = (kHugePage + chunk_bytes - 1) / chunk_bytes; if (group < 1) group = 1; if (c + group > kMaxChunks) group = kMaxChunks - c; }
      if (group < 1) abort();                              // more than kMaxChunks chunks: index space exhausted
      T* base = static_cast<T*>(HugeAlloc((size_t)(group * chunk_bytes)));
      bases_.push_back(base);
      for (int64_t g = 0; g < group; ++g) chunks_[c + g].store(base + g * chunk_elems, std::memory_order_release);
      c += group;
      nchunks_.store(c, std::memory_order_release);
    }
  }
  int64_t capacity() const { return nchunks_.load() << kLog2Chunk; }
  int64_t width() const { return width_; }
 private:
  int64_t width_;
  std::atomic<T*>* chunks_;
  std::atomic<int64_t> nchunks_{0};
  std::mutex mu_;
  std::vector<T*> bases_;              // what free() gets: one pointer per chunk (small) or per slab of chunks
  static constexpr int64_t kSmallChunks = 8;
};

// ------------------------------------------------------------------------------------
// Partitioned open-addressing KV: key -> meta index.
// ------------------------------------------------------------------------------------
static constexpr int64_t kEmptyKey = INT64_MIN;

// One 16-byte slot per entry: key and value index share a cache line, so a probe costs one DRAM access instead of two.
struct KVSlot { std::atomic<int64_t> key; std::atomic<int32_t> val; int32_t pad; };
// slots / cap are read by every probe, size and the lock word are written by every insert: three separate cache lines
struct alignas(64) KVPart {
  KVSlot* slots = nullptr;
  int64_t cap = 0;
  alignas(64) std::atomic<int64_t> size{0};
  alignas(64) std::shared_mutex mu;
  ~KVPart() { free(slots); }
  void Alloc(int64_t c) {
    cap = c;
    slots = static_cast<KVSlot*>(HugeAlloc(sizeof(KVSlot) * (size_t)c));
    for (int64_t i = 0; i < c; ++i) { slots[i].key.store(kEmptyKey, std::memory_order_relaxed); slots[i].val.store(-1, std::memory_order_relaxed); slots[i].pad = 0; }
  }
};

class HostKV {
 public:
  HostKV(int nparts, int64_t init_cap) : nparts_(std::max(1, nparts)), parts_(new KVPart[nparts_]) {
    int64_t per = 64;
    while (per * nparts_ < init_cap * 2) per <<= 1;
    for (int p = 0; p < nparts_; ++p) parts_[p].Alloc(per);
  }
  ~HostKV() { delete[] parts_; }
  int PartOf(int64_t key) const { return (int)((dr_mix64((uint64_t)key) >> 40) % (uint64_t)nparts_); }

  // wait-free find; returns meta index or -1
  int32_t Find(int64_t key) {
    if (key == kEmptyKey) return -1;                   // the reserved padding / empty-slot value is never a stored key
    KVPart& P = parts_[PartOf(key)];
    std::shared_lock<std::shared_mutex> l(P.mu);
    return FindLocked(P, key);
  }
  // Batch readers: hold every partition's shared lock ONCE for a whole key range instead of once per key (the per-key lock/unlock
  // is an atomic RMW on a line shared by all reader threads: ~2/3 of the cost of a random lookup), then probe with FindNoLock and
  // hide the DRAM latency of the probes with PrefetchSlot.
  class SharedAll {
   public:
    explicit SharedAll(HostKV& kv) : kv_(kv) { for (int p = 0; p < kv_.nparts_; ++p) kv_.parts_[p].mu.lock_shared(); }
    ~SharedAll() { for (int p = kv_.nparts_ - 1; p >= 0; --p) kv_.parts_[p].mu.unlock_shared(); }
    SharedAll(const SharedAll&) = delete;
   private:
    HostKV& kv_;
  };
  int32_t FindNoLock(int64_t key) { return key == kEmptyKey ? -1 : FindLocked(parts_[PartOf(key)], key); }
  void PrefetchSlot(int64_t key) const {
    const KVPart& P = parts_[PartOf(key)];
    const uint64_t pos = dr_mix64((uint64_t)key) & (uint64_t)(P.cap - 1);
    __builtin_prefetch(&P.slots[pos]);
  }
  // find or insert; `alloc` is called exactly once by the inserting thread to get a meta index.
  template <typename Alloc>
  int32_t FindOrInsert(int64_t key, Alloc&& alloc, bool* inserted) {
    KVPart& P = parts_[PartOf(key)];
    *inserted = false;
    for (;;) {
      bool need_grow = false;
      {
        std::shared_lock<std::shared_mutex> l(P.mu);
        uint64_t mask = P.cap - 1;
        uint64_t pos = dr_mix64((uint64_t)key) & mask;
        for (int64_t probes = 0; probes < P.cap; ++probes, pos = (pos + 1) & mask) {
          int64_t k = P.slots[pos].key.load(std::memory_order_acquire);
          if (k == key) return WaitVal(P, pos);
          if (k == kEmptyKey) {
            if ((P.size.load(std::memory_order_relaxed) + 1) * 10 > P.cap * 7) { need_grow = true; break; }
            int64_t expected = kEmptyKey;
            if (P.slots[pos].key.compare_exchange_strong(expected, key, std::memory_order_acq_rel)) {
              P.size.fetch_add(1, std::memory_order_relaxed);
              int32_t v = alloc();
              P.slots[pos].val.store(v, std::memory_order_release);
              *inserted = true;
              return v;
            }
            if (expected == key) return WaitVal(P, pos);
            // lost the slot to another key: keep probing from the same slot's successor
          }
        }
        if (!need_grow) need_grow = true;  // table full
      }
      Grow(P);
    }
  }
  int64_t Size() const { int64_t s = 0; for (int p = 0; p < nparts_; ++p) s += parts_[p].size.load(); return s; }

  int NumParts() const { return nparts_; }
  // Make room for `extra` more keys up front (restore of a checkpoint: one rebuild per partition instead of one per doubling).
  template <typename PF> void Reserve(int64_t extra, PF&& parallel_for) {
    const int64_t per = extra / nparts_ + extra / (nparts_ * 8) + 64;            // uniform hashing: +12.5 % slack per partition
    parallel_for(nparts_, [&](int64_t pb, int64_t pe) {
      for (int64_t p = pb; p < pe; ++p) {
        KVPart& P = parts_[p];
        std::unique_lock<std::shared_mutex> l(P.mu);
        int64_t cap = P.cap;
        while ((P.size.load() + per) * 10 > cap * 7) cap <<= 1;
        if (cap == P.cap) continue;
        std::vector<std::pair<int64_t, int32_t>> items; items.reserve((size_t)P.size.load());
        for (int64_t i = 0; i < P.cap; ++i) {
          int64_t k = P.slots[i].key.load(std::memory_order_relaxed);
          if (k != kEmptyKey) items.emplace_back(k, P.slots[i].val.load(std::memory_order_relaxed));
        }
        Rebuild(P, cap, items);
      }
    });
  }
  // Iterate the (key, idx) pairs of one partition / of all partitions (caller guarantees no concurrent writers, as in Save).  Whole-table
  // passes (snapshot, eviction, dirty reset) run one partition per worker.
  template <typename F> void ForEachInPart(int p, F&& f) {
    KVPart& P = parts_[p];
    std::shared_lock<std::shared_mutex> l(P.mu);
    for (int64_t i = 0; i < P.cap; ++i) {
      int64_t k = P.slots[i].key.load(std::memory_order_acquire);
      if (k == kEmptyKey) continue;
      // a CAS insert running under the same shared lock publishes the key first and the value index second: a slot whose value is still
      // -1 is not part of the table yet (its metadata does not exist) -- skip it, as the point readers' WaitVal would wait for it
      const int32_t v = P.slots[i].val.load(std::memory_order_acquire);
      if (v >= 0) f(k, v);
    }
  }
  template <typename F> void ForEach(F&& f) { for (int p = 0; p < nparts_; ++p) ForEachInPart(p, f); }
  template <typename F> void ForEachInPartNoLock(int p, F&& f) {          // caller holds ExclusiveAll
    KVPart& P = parts_[p];
    for (int64_t i = 0; i < P.cap; ++i) {
      int64_t k = P.slots[i].key.load(std::memory_order_acquire);
      if (k == kEmptyKey) continue;
      const int32_t v = P.slots[i].val.load(std::memory_order_acquire);
      if (v >= 0) f(k, v);
    }
  }
  class ExclusiveAll {                                                     // stops inserts, growth and batch readers on every partition
   public:
    explicit ExclusiveAll(HostKV& kv) : kv_(kv) { for (int p = 0; p < kv_.nparts_; ++p) kv_.parts_[p].mu.lock(); }
    ~ExclusiveAll() { for (int p = kv_.nparts_ - 1; p >= 0; --p) kv_.parts_[p].mu.unlock(); }
    ExclusiveAll(const ExclusiveAll&) = delete;
   private:
    HostKV& kv_;
  };
  // Remove every key for which pred(key, idx) is true (partition rebuilt under exclusive lock; nothing is rebuilt when nothing goes).
  template <typename Pred> int64_t RemoveIfInPart(int p, Pred&& pred) {
    int64_t removed = 0;
    KVPart& P = parts_[p];
    std::unique_lock<std::shared_mutex> l(P.mu);
    std::vector<std::pair<int64_t, int32_t>> keep; keep.reserve(P.size.load());
    for (int64_t i = 0; i < P.cap; ++i) {
      int64_t k = P.slots[i].key.load(std::memory_order_relaxed);
      if (k == kEmptyKey) continue;
      int32_t v = P.slots[i].val.load(std::memory_order_relaxed);
      if (pred(k, v)) ++removed; else keep.emplace_back(k, v);
    }
    if (removed) Rebuild(P, P.cap, keep);
    return removed;
  }
  template <typename Pred> int64_t RemoveIf(Pred&& pred) {
    int64_t removed = 0;
    for (int p = 0; p < nparts_; ++p) removed += RemoveIfInPart(p, pred);
    return removed;
  }
 private:
  static int32_t WaitVal(KVPart& P, uint64_t pos) {
    int32_t v;
    while ((v = P.slots[pos].val.load(std::memory_order_acquire)) < 0) std::this_thread::yield();
    return v;
  }
  static int32_t FindLocked(KVPart& P, int64_t key) {
    uint64_t mask = P.cap - 1;
    uint64_t pos = dr_mix64((uint64_t)key) & mask;
    for (int64_t probes = 0; probes < P.cap; ++probes, pos = (pos + 1) & mask) {
      int64_t k = P.slots[pos].key.load(std::memory_order_acquire);
      if (k == key) return WaitVal(P, pos);
      if (k == kEmptyKey) return -1;
    }
    return -1;
  }
  static void Rebuild(KVPart& P, int64_t newcap, const std::vector<std::pair<int64_t, int32_t>>& items) {
    free(P.slots);
    P.Alloc(newcap);
    uint64_t mask = newcap - 1;
    for (auto& kv : items) {
      uint64_t pos = dr_mix64((uint64_t)kv.first) & mask;
      while (P.slots[pos].key.load(std::memory_order_relaxed) != kEmptyKey) pos = (pos + 1) & mask;
      P.slots[pos].key.store(kv.first, std::memory_order_relaxed);
      P.slots[pos].val.store(kv.second, std::memory_order_relaxed);
    }
    P.size.store((int64_t)items.size());
  }
  static void Grow(KVPart& P) {
    std::unique_lock<std::shared_mutex> l(P.mu);
    if ((P.size.load() + 1) * 10 <= P.cap * 7) return;  // someone else grew it
    std::vector<std::pair<int64_t, int32_t>> items; items.reserve(P.size.load());
    for (int64_t i = 0; i < P.cap; ++i) {
      int64_t k = P.slots[i].key.load(std::memory_order_relaxed);
      if (k != kEmptyKey) items.emplace_back(k, P.slots[i].val.load(std::memory_order_relaxed));
    }
    Rebuild(P, P.cap * 2, items);
  }
  int nparts_;
  KVPart* parts_;
};

// ------------------------------------------------------------------------------------
// Counting Bloom filter (CBFFilter). k = ceil(log2(1/p)), m = ceil(n*|ln p|/ln^2 2)
// (embedding_config.h:72-95).  Counters saturate at their width.
// ------------------------------------------------------------------------------------
class CountingBloom {
 public:
  CountingBloom(int64_t n, double p, int bits) : bits_(bits) {
    if (p <= 0 || p >= 1) p = 0.01;
    if (n < 1) n = 1;
    k_ = std::max(1, (int)std::ceil(std::log2(1.0 / p)));
    m_ = std::max<int64_t>(8, (int64_t)std::ceil((double)n * std::fabs(std::log(p)) / (std::log(2.0) * std::log(2.0))));
    bytes_per_ = bits / 8;
    data_.reset(new std::atomic<uint8_t>[m_ * bytes_per_]);
    for (int64_t i = 0; i < m_ * bytes_per_; ++i) data_[i].store(0, std::memory_order_relaxed);
    maxv_ = bits >= 63 ? INT64_MAX : ((int64_t(1) << bits) - 1);
  }
  // add `count` to the k counters, return the new minimum
  int64_t AddAndMin(int64_t key, int64_t count) {
    int64_t mn = INT64_MAX;
    for (int i = 0; i < k_; ++i) {
      int64_t idx = (int64_t)(dr_hash_seed((uint64_t)key, (uint64_t)i) % (uint64_t)m_);
      mn = std::min(mn, AddAt(idx, count));
    }
    return mn;
  }
  int64_t Min(int64_t key) const {
    int64_t mn = INT64_MAX;
    for (int i = 0; i < k_; ++i) {
      int64_t idx = (int64_t)(dr_hash_seed((uint64_t)key, (uint64_t)i) % (uint64_t)m_);
      mn = std::min(mn, Load(idx));
    }
    return mn;
  }
  int k() const { return k_; }
  int64_t m() const { return m_; }
  int bits() const { return bits_; }
  int64_t nbytes() const { return m_ * bytes_per_; }
  uint8_t* raw() { return reinterpret_cast<uint8_t*>(data_.get()); }
 private:
  int64_t Load(int64_t idx) const {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(data_.get()) + idx * bytes_per_;
    switch (bits_) {
      case 8: return *p;
      case 16: { uint16_t v; memcpy(&v, p, 2); return v; }
      case 32: { uint32_t v; memcpy(&v, p, 4); return v; }
      default: { int64_t v; memcpy(&v, p, 8); return v; }
    }
  }
  template <typename U> int64_t AddT(int64_t idx, int64_t count) {
    auto* a = reinterpret_cast<std::atomic<U>*>(reinterpret_cast<uint8_t*>(data_.get()) + idx * sizeof(U));
    U cur = a->load(std::memory_order_relaxed);
    for (;;) {
      int64_t nv = std::min<int64_t>(maxv_, (int64_t)cur + count);
      if (a->compare_exchange_weak(cur, (U)nv, std::memory_order_relaxed)) return nv;
    }
  }
  int64_t AddAt(int64_t idx, int64_t count) {
    switch (bits_) {
      case 8: return AddT<uint8_t>(idx, count);
      case 16: return AddT<uint16_t>(idx, count);
      case 32: return AddT<uint32_t>(idx, count);
      default: return AddT<uint64_t>(idx, count);
    }
  }
  int bits_, k_, bytes_per_;
  int64_t m_, maxv_;
  std::unique_ptr<std::atomic<uint8_t>[]> data_;
};

// ------------------------------------------------------------------------------------
// Host EmbeddingVariable
// ------------------------------------------------------------------------------------
struct SnapItem { int32_t bucket; int32_t idx; int64_t key; };

class HostEV {
 public:
  explicit HostEV(const DrEvConfig& c)
      : cfg_(c), stride_(dr_row_stride(c.dim, c.num_slots, c.has_scalars)),
        kv_(c.num_partitions > 0 ? c.num_partitions : 16, std::max<int64_t>(1024, c.init_capacity)),
        meta_(1), rows_(stride_) {
    default_.assign((size_t)(std::max<int64_t>(1, c.default_value_dim) * c.dim), 0.0f);
    if (c.filter_type == DR_FILTER_BLOOM)
      bloom_.reset(new CountingBloom(c.bloom_max_elements, c.bloom_fpp, c.bloom_counter_bits > 0 ? c.bloom_counter_bits : 32));
  }
  const DrEvConfig& cfg() const { return cfg_; }
  int64_t stride() const { return stride_; }
  void SetDefault(const float* m) { memcpy(default_.data(), m, default_.size() * sizeof(float)); }
  int64_t Size() const { return admitted_.load(); }           // admitted keys (total_count())
  int64_t TotalKeys() const { return kv_.Size(); }            // incl. un-admitted (counter filter)

  // ---- forward: read-only gather (embedding_var.h:202-219) ---------------------------
  void Lookup(const int64_t* keys, int64_t n, float* out) {
    GlobalPool()->ParallelFor(n, 2048, [&](int64_t b, int64_t e) { LookupRange(keys, b, e, out, cfg_.dim); });
  }
  // keys[b, e) -> out + i * out_stride (serial; callers parallelise over ranges / tables)
  void LookupRange(const int64_t* keys, int64_t b, int64_t e, float* out, int64_t out_stride) {
    const int64_t dim = cfg_.dim;
    constexpr int W = 16;                                  // software pipeline: slots of W keys, then their metadata, then their rows
    HostKV::SharedAll guard(kv_);
    for (int64_t i0 = b; i0 < e; i0 += W) {
      const int n = (int)std::min<int64_t>(W, e - i0);
      if (i0 + W < e) for (int j = 0; j < (int)std::min<int64_t>(W, e - i0 - W); ++j) kv_.PrefetchSlot(keys[i0 + W + j]);
      int32_t row[W];
      for (int j = 0; j < n; ++j) {
        if (keys[i0 + j] == kEmptyKey) { row[j] = -3; continue; }       // DR_PAD_KEY: "no id here" (padding of a dense [B, L] id tensor)
        const int32_t idx = kv_.FindNoLock(keys[i0 + j]);
        row[j] = idx;
        if (idx >= 0) __builtin_prefetch(meta_.at(idx));
      }
      for (int j = 0; j < n; ++j) {
        if (row[j] == -3) continue;
        row[j] = row[j] >= 0 ? RowOf(row[j]) : -1;
        if (row[j] >= 0) __builtin_prefetch(rows_.at(row[j]));
      }
      for (int j = 0; j < n; ++j) {
        const int64_t i = i0 + j;
        float* o = out + i * out_stride;
        if (row[j] == -3) {
          std::fill(o, o + dim, 0.f);                                      // padding reads zeros, is never created, counted or updated
        } else if (row[j] >= 0) {
          memcpy(o, rows_.at(row[j]), dim * sizeof(float));
        } else if (cfg_.filter_type != DR_FILTER_NONE && cfg_.filter_freq > 0) {
          std::fill(o, o + dim, cfg_.default_value_no_permission);
        } else {
          memcpy(o, DefaultRow(keys[i]), dim * sizeof(float));
        }
      }
    }
  }
  // samples [b0, b1) of a dense [B, L] id tensor: out[b] = sum of rows (PAD skipped; unseen / un-admitted ids read what Lookup reads)
  void LookupPooledRange(const int64_t* ids, int64_t b0, int64_t b1, int64_t L, float* out, int64_t out_stride) {
    const int64_t dim = cfg_.dim;
    HostKV::SharedAll guard(kv_);
    std::vector<int32_t> row((size_t)L);
    for (int64_t b = b0; b < b1; ++b) {
      const int64_t* k = ids + b * L;
      if (b + 1 < b1) for (int64_t j = 0; j < L; ++j) kv_.PrefetchSlot(ids[(b + 1) * L + j]);       // next sample's probes in flight
      for (int64_t j = 0; j < L; ++j) {
        if (k[j] == kEmptyKey) { row[(size_t)j] = -3; continue; }
        const int32_t idx = kv_.FindNoLock(k[j]);
        row[(size_t)j] = idx >= 0 ? RowOf(idx) : -1;
        if (row[(size_t)j] >= 0) __builtin_prefetch(rows_.at(row[(size_t)j]));
      }
      float* o = out + b * out_stride;
      std::fill(o, o + dim, 0.f);
      for (int64_t j = 0; j < L; ++j) {
        const int32_t r = row[(size_t)j];
        if (r == -3) continue;
        if (r >= 0) { const float* src = rows_.at(r); for (int64_t d = 0; d < dim; ++d) o[d] += src[d]; }
        else if (cfg_.filter_type != DR_FILTER_NONE && cfg_.filter_freq > 0) { for (int64_t d = 0; d < dim; ++d) o[d] += cfg_.default_value_no_permission; }
        else { const float* src = DefaultRow(k[j]); for (int64_t d = 0; d < dim; ++d) o[d] += src[d]; }
      }
    }
  }
  // gather a slot (or the trailing scalars with slot == num_slots+1) for inspection / ckpt
  void LookupSlot(const int64_t* keys, int64_t n, int slot, float* out) {
    const int64_t dim = cfg_.dim;
    for (int64_t i = 0; i < n; ++i) {
      int32_t idx = kv_.Find(keys[i]);
      int32_t r = idx >= 0 ? RowOf(idx) : -1;
      float* o = out + i * dim;
      if (r >= 0) memcpy(o, rows_.at(r) + slot * dim, dim * sizeof(float));
      else std::fill(o, o + dim, slot == 0 ? 0.f : cfg_.slot_init[slot - 1]);
    }
  }
  // metadata reads: same probe pipeline as LookupRange (batch lock, slots of W keys in flight, then their metadata lines)
  template <typename F> void ForEachMetaIndex(const int64_t* keys, int64_t n, F&& f) {
    GlobalPool()->ParallelFor(n, 4096, [&](int64_t b, int64_t e) {
      constexpr int W = 16;
      HostKV::SharedAll guard(kv_);
      for (int64_t i0 = b; i0 < e; i0 += W) {
        const int m = (int)std::min<int64_t>(W, e - i0);
        if (i0 + W < e) for (int j = 0; j < (int)std::min<int64_t>(W, e - i0 - W); ++j) kv_.PrefetchSlot(keys[i0 + W + j]);
        int32_t idx[W];
        for (int j = 0; j < m; ++j) { idx[j] = kv_.FindNoLock(keys[i0 + j]); if (idx[j] >= 0) __builtin_prefetch(meta_.at(idx[j])); }
        for (int j = 0; j < m; ++j) f(i0 + j, idx[j]);
      }
    });
  }
  void GetFreq(const int64_t* keys, int64_t n, int64_t* out) {
    ForEachMetaIndex(keys, n, [&](int64_t i, int32_t idx) { out[i] = idx >= 0 ? FreqOf(idx) : (bloom_ ? bloom_->Min(keys[i]) : 0); });
  }
  void GetVersion(const int64_t* keys, int64_t n, int64_t* out) {
    ForEachMetaIndex(keys, n, [&](int64_t i, int32_t idx) { out[i] = idx >= 0 ? VersionOf(idx) : -1; });
  }

  // ---- LookupOrCreateKey with admission (counter_filter_policy.h:106-139) -------------
  // returns row index (>=0) when the key is admitted, -1 otherwise.
  // A block of ApplyRange reserves the metadata / row indices of all its new keys with ONE fetch_add each (the per-key global
  // counters sat on one cache line and made 8 inserting threads slower than one).
  struct Reserve { int64_t meta_next = 0, meta_end = 0, row_next = 0, row_end = 0, admitted = 0; };

  // `known` = meta index already resolved by a batched read-only probe (ApplyRange), or -1
  int32_t LookupOrCreate(int64_t key, int64_t count, int64_t step, int32_t known = -1, Reserve* rs = nullptr) {
    if (cfg_.is_inference) {
      int32_t idx = known >= 0 ? known : kv_.Find(key);
      return idx >= 0 ? RowOf(idx) : -1;
    }
    if (bloom_) {
      int32_t idx = known >= 0 ? known : kv_.Find(key);
      if (idx < 0) {
        int64_t mn = bloom_->AddAndMin(key, count);
        if (mn < cfg_.filter_freq) return -1;
      }
    }
    bool inserted = false;
    int32_t idx = known >= 0 ? known : kv_.FindOrInsert(key, [this, rs] { return AllocMeta(rs); }, &inserted);
    int64_t* f = (&meta_.at(idx)->freq);
    int64_t nf = __atomic_add_fetch(f, count, __ATOMIC_RELAXED);
    __atomic_store_n((&meta_.at(idx)->version), step, __ATOMIC_RELAXED);
    __atomic_store_n((&meta_.at(idx)->dirty), (uint8_t)1, __ATOMIC_RELAXED);
    int32_t* rp = (&meta_.at(idx)->row);
    int32_t r = __atomic_load_n(rp, __ATOMIC_ACQUIRE);
    if (r >= 0) return r;
    bool admit = bloom_ ? true : (cfg_.filter_type == DR_FILTER_NONE || nf >= cfg_.filter_freq);
    if (!admit) return -1;
    // claim allocation: -1 -> -2 (pending)
    int32_t expect = -1;
    if (__atomic_compare_exchange_n(rp, &expect, -2, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
      r = AllocRow(rs);
      InitRow(r, key);
      __atomic_store_n(rp, r, __ATOMIC_RELEASE);
      if (rs) ++rs->admitted; else admitted_.fetch_add(1);
      return r;
    }
    while ((r = __atomic_load_n(rp, __ATOMIC_ACQUIRE)) < 0) std::this_thread::yield();
    return r;
  }

  // ---- sparse apply on de-duplicated keys (training_ali_ops.cc) ------------------------
  void Apply(const int64_t* keys, const float* grads, const int64_t* counts, int64_t n, const DrOptHyper& hp) {
    GlobalPool()->ParallelFor(n, 512, [&](int64_t b, int64_t e) { ApplyRange(keys, grads, counts, b, e, hp); });
  }
  // serial body over [b, e): callers parallelise over key ranges (Apply) or over tables (group apply)
  void ApplyRange(const int64_t* keys, const float* grads, const int64_t* counts, int64_t b, int64_t e, const DrOptHyper& hp) {
    const int64_t dim = cfg_.dim;
    const float alpha = dr_adam_alpha(hp);
    // Blocks of kBlock keys: (1) a read-only pass under ONE shared-lock acquisition resolves the keys that already exist (in steady
    // state almost all of them) with prefetched probes and pulls the slots / metadata of this block into the cache; (2) the update
    // pass right behind it finds everything it touches still cached, and only the misses go through the locking find-or-insert.
    // (Remove / Shrink never run concurrently with Apply, so an index found in (1) stays valid.)
    constexpr int64_t kBlock = 256;
    int32_t known_blk[kBlock];
    std::vector<float> newacc(dim);
    for (int64_t blk = b; blk < e; blk += kBlock) {
      const int64_t be = std::min(e, blk + kBlock);
      int32_t* known = known_blk - blk;                       // known[i] for i in [blk, be)
      {
        HostKV::SharedAll guard(kv_);
        constexpr int W = 16;
        for (int64_t i0 = blk; i0 < be; i0 += W) {
          const int n = (int)std::min<int64_t>(W, be - i0);
          if (i0 + W < be) for (int j = 0; j < (int)std::min<int64_t>(W, be - i0 - W); ++j) kv_.PrefetchSlot(keys[i0 + W + j]);
          for (int j = 0; j < n; ++j) {
            if (keys[i0 + j] == kEmptyKey) { known[i0 + j] = -3; continue; }      // padding id: skipped below
            const int32_t idx = kv_.FindNoLock(keys[i0 + j]);
            known[i0 + j] = idx;
            if (idx >= 0) __builtin_prefetch(meta_.at(idx), 1);
          }
        }
        for (int64_t i = blk; i < be; ++i)
          if (known[i] >= 0) { const int32_t r = RowOf(known[i]); if (r >= 0) __builtin_prefetch(rows_.at(r), 1); }
      }
      Reserve rs;
      if (!cfg_.is_inference && !bloom_ && n_free_meta_.load(std::memory_order_relaxed) == 0 && n_free_rows_.load(std::memory_order_relaxed) == 0) {
        int64_t misses = 0;
        for (int64_t i = blk; i < be; ++i) misses += known[i] == -1;
        if (misses) {
          rs.meta_next = next_meta_.fetch_add(misses); rs.meta_end = rs.meta_next + misses;
          meta_.EnsureCapacity(rs.meta_end, Meta{0, -1, -1, 0, {0}});
          if (cfg_.filter_type == DR_FILTER_NONE) {            // every new key is admitted at once: its row can be reserved too
            rs.row_next = next_row_.fetch_add(misses); rs.row_end = rs.row_next + misses;
            rows_.EnsureCapacity(rs.row_end, 0.f);
          }
        }
      }
      for (int64_t i = blk; i < be; ++i) {
        if (known[i] == -3) continue;
        int32_t r = LookupOrCreate(keys[i], counts ? counts[i] : 1, hp.global_step, known[i], &rs);
        if (r < 0) continue;
        float* row = rows_.at(r);
        const float* g = grads + i * dim;
        float* s0 = row + dim; float* s1 = row + 2 * dim;
        float dummy0 = 0.f, dummy1 = 0.f;
        if (hp.kind == DR_OPT_FTRL) {
          float sq = 0.f;
          for (int64_t d = 0; d < dim; ++d) { float l = dr_ftrl_linear(hp, g[d], row[d], s0[d], s1[d], newacc[d]); sq += l * l; }
          float norm = std::sqrt(sq);
          for (int64_t d = 0; d < dim; ++d) {
            row[d] = dr_ftrl_weight(hp, s1[d], newacc[d], norm);
            s0[d] += g[d] * g[d];   // reference advances accum with the raw gradient
          }
          continue;
        }
        bool decay_now = false;
        if (hp.kind == DR_OPT_ADAGRAD_DECAY) {
          float* sc = row + dim * (1 + cfg_.num_slots);
          if (hp.decay_step > 0 && (float)(hp.global_step / hp.decay_step) > sc[0]) { decay_now = true; sc[0] += 1.0f; }
        }
        const int ns = cfg_.num_slots;
        for (int64_t d = 0; d < dim; ++d)
          dr_apply_elem(hp.kind, hp, alpha, decay_now, g[d], row[d], ns > 0 ? s0[d] : dummy0, ns > 1 ? s1[d] : dummy1);
      }
      if (rs.admitted) admitted_.fetch_add(rs.admitted);
    }
  }

  // ---- eviction (only inside save; single_tier_storage.h:235-261) -----------------------
  int64_t Shrink(int64_t global_step) {
    const bool gs = cfg_.steps_to_live > 0;
    const bool l2 = cfg_.l2_weight_threshold >= 0.f;
    if (!gs && !l2) return 0;
    // one hash partition per worker; each collects what it freed, the free lists are extended once at the end
    const int np = kv_.NumParts();
    std::vector<std::vector<int32_t>> freed_meta((size_t)np), freed_rows((size_t)np);
    std::atomic<int64_t> removed{0};
    GlobalPool()->ParallelFor(np, 1, [&](int64_t pb, int64_t pe) {
      for (int64_t p = pb; p < pe; ++p) {
        removed.fetch_add(kv_.RemoveIfInPart((int)p, [&](int64_t key, int32_t idx) {
          (void)key;
          int32_t r = *(&meta_.at(idx)->row);
          bool evict = false;
          if (gs) {
            int64_t* v = (&meta_.at(idx)->version);
            if (*v == -1) *v = global_step;                       // globalstep_shrink_policy.h:50
            else if (global_step - *v > cfg_.steps_to_live) evict = true;
          }
          if (!evict && l2 && r >= 0) {
            const float* row = rows_.at(r); float s = 0.f;
            for (int64_t d = 0; d < cfg_.dim; ++d) s += row[d] * row[d];
            if (0.5f * s < cfg_.l2_weight_threshold) evict = true;  // l2weight_shrink_policy.h:52
          }
          if (evict) {
            if (r >= 0) freed_rows[(size_t)p].push_back(r);
            freed_meta[(size_t)p].push_back(idx);
          }
          return evict;
        }));
      }
    });
    FreeBulk(freed_meta, freed_rows);
    return removed.load();
  }
  int64_t Remove(const int64_t* keys, int64_t n) {
    std::vector<int64_t> ks(keys, keys + n); std::sort(ks.begin(), ks.end());
    std::vector<int32_t> freed;
    int64_t removed = kv_.RemoveIf([&](int64_t key, int32_t idx) {
      if (!std::binary_search(ks.begin(), ks.end(), key)) return false;
      int32_t r = *(&meta_.at(idx)->row);
      if (r >= 0) { FreeRow(r); admitted_.fetch_sub(1); }
      freed.push_back(idx);
      return true;
    });
    for (int32_t idx : freed) FreeMeta(idx);
    return removed;
  }

  // ---- snapshot for checkpoint / elastic export -------------------------------------
  // dirty_only: incremental checkpoint (keys touched since the last ClearDirty()).
  // part filter: keep key%1000%part_num == part_id (sharded snapshot for elastic scaling).
  void SnapshotBegin(int dirty_only, int part_id, int part_num, int64_t* n_admitted, int64_t* n_filtered) {
    // Two scans of the hash partitions instead of per-partition item vectors: pass 1 only counts (per partition x checkpoint bucket,
    // admitted / filtered), pass 2 writes every item straight to its final position -- the only large allocation is the result itself
    // (a 10 M-key snapshot used to touch ~3x that while the vectors grew), then every bucket is sorted by key.
    const int np = kv_.NumParts();
    auto classify = [&](int64_t key, int32_t idx, int* bucket) -> int {       // 0 = skip, 1 = admitted, 2 = filtered
      *bucket = dr_ckpt_bucket(key);
      if (part_num > 1 && *bucket % part_num != part_id) return 0;
      if (dirty_only && !*(&meta_.at(idx)->dirty)) return 0;
      return *(&meta_.at(idx)->row) >= 0 ? 1 : 2;
    };
    // Save runs between steps, so the two scans normally see the same table.  If someone does modify it in between (a parameter server
    // checkpointing while pushes arrive), the placement pass notices (a range overflows or is left short; overflowing items are dropped,
    // never written out of range) and the snapshot is retried -- the last attempt under exclusive partition locks.
    // NOTE for callers: the exclusive last attempt freezes the hash partitions (inserts, growth, batch readers) but NOT Apply / ImportCow,
    // which flip meta->row / meta->dirty of already-known indices: whoever checkpoints a live table must quiesce applies (the parameter
    // server holds its push lock across save()).  If even the last attempt disagrees with its own counting scan, the ranges are compacted
    // to what was actually placed below -- never a zeroed or stale item in the output.
    std::vector<int64_t> cnt_a, cnt_f, end_a, end_f, start_a, start_f;
    bool consistent = false;
    for (int attempt = 0; attempt < 3; ++attempt) {
      const bool exclusive = attempt == 2;
      std::unique_ptr<HostKV::ExclusiveAll> guard;
      if (exclusive) guard.reset(new HostKV::ExclusiveAll(kv_));
      auto scan = [&](int p, auto&& f) { if (exclusive) kv_.ForEachInPartNoLock(p, f); else kv_.ForEachInPart(p, f); };
      cnt_a.assign((size_t)np * 1000, 0); cnt_f.assign((size_t)np * 1000, 0);
      GlobalPool()->ParallelFor(np, 1, [&](int64_t pb, int64_t pe) {
        for (int64_t p = pb; p < pe; ++p)
          scan((int)p, [&](int64_t key, int32_t idx) {
            int b; const int c = classify(key, idx, &b);
            if (c == 1) cnt_a[(size_t)p * 1000 + b]++; else if (c == 2) cnt_f[(size_t)p * 1000 + b]++;
          });
      });
      auto offsets = [&](std::vector<int64_t>& cnt, std::vector<int64_t>& end, std::vector<int64_t>& off, std::vector<SnapItem>& dst) {
        off.assign(1001, 0); end.assign(cnt.size(), 0);
        int64_t run = 0;
        for (int b = 0; b < 1000; ++b) {
          off[b] = run;
          for (int p = 0; p < np; ++p) { const int64_t c = cnt[(size_t)p * 1000 + b]; cnt[(size_t)p * 1000 + b] = run; run += c; end[(size_t)p * 1000 + b] = run; }
        }
        off[1000] = run;
        dst.resize((size_t)run);
      };
      offsets(cnt_a, end_a, snap_adm_off_, snap_adm_);
      offsets(cnt_f, end_f, snap_flt_off_, snap_flt_);
      start_a = cnt_a; start_f = cnt_f;
      std::atomic<int64_t> dropped{0};
      GlobalPool()->ParallelFor(np, 1, [&](int64_t pb, int64_t pe) {
        for (int64_t p = pb; p < pe; ++p)
          scan((int)p, [&](int64_t key, int32_t idx) {
            int b; const int c = classify(key, idx, &b);
            if (c == 0) return;
            std::vector<int64_t>& cur = c == 1 ? cnt_a : cnt_f; const std::vector<int64_t>& end = c == 1 ? end_a : end_f;
            const size_t slot = (size_t)p * 1000 + b;
            if (cur[slot] >= end[slot]) { dropped.fetch_add(1, std::memory_order_relaxed); return; }
            (c == 1 ? snap_adm_ : snap_flt_)[(size_t)cur[slot]++] = {b, idx, key};
          });
      });
      consistent = dropped.load() == 0;
      for (size_t i = 0; consistent && i < end_a.size(); ++i) consistent = cnt_a[i] == end_a[i] && cnt_f[i] == end_f[i];
      if (consistent) break;
    }
    if (!consistent) {
      // ranges left short (or overflowed): keep exactly the items that were placed, bucket by bucket, and recompute the offsets
      auto compact = [&](std::vector<SnapItem>& dst, std::vector<int64_t>& off, const std::vector<int64_t>& start, const std::vector<int64_t>& cur) {
        std::vector<SnapItem> out; out.reserve(dst.size());
        std::vector<int64_t> noff(1001, 0);
        for (int b = 0; b < 1000; ++b) {
          noff[b] = (int64_t)out.size();
          for (int p = 0; p < np; ++p) { const size_t sl = (size_t)p * 1000 + b; for (int64_t i = start[sl]; i < cur[sl]; ++i) out.push_back(dst[(size_t)i]); }
        }
        noff[1000] = (int64_t)out.size();
        dst.swap(out); off.swap(noff);
      };
      compact(snap_adm_, snap_adm_off_, start_a, cnt_a);
      compact(snap_flt_, snap_flt_off_, start_f, cnt_f);
      fprintf(stderr, "[deeprec_host] snapshot of a table that kept changing: kept the %zu + %zu items that were placed consistently; "
                      "quiesce applies while checkpointing\n", snap_adm_.size(), snap_flt_.size());
    }
    auto sort_buckets = [&](std::vector<SnapItem>& dst, const std::vector<int64_t>& off) {
      GlobalPool()->ParallelFor(1000, 8, [&](int64_t bb, int64_t be) {
        for (int64_t b = bb; b < be; ++b)
          std::sort(dst.begin() + off[b], dst.begin() + off[b + 1], [](const SnapItem& x, const SnapItem& y) { return x.key < y.key; });
      });
    };
    sort_buckets(snap_adm_, snap_adm_off_);
    sort_buckets(snap_flt_, snap_flt_off_);
    *n_admitted = (int64_t)snap_adm_.size(); *n_filtered = (int64_t)snap_flt_.size();
  }
  void SnapshotRead(int64_t* keys, float* rows, int64_t* freqs, int64_t* versions, int64_t* part_offset,
                    int64_t* fkeys, int64_t* ffreqs, int64_t* fversions, int64_t* fpart_offset) {
    auto fill = [&](std::vector<SnapItem>& v, const std::vector<int64_t>& offs, int64_t* k, float* rw, int64_t* f, int64_t* ver, int64_t* off) {
      if (off) { if (offs.size() == 1001) memcpy(off, offs.data(), 1001 * sizeof(int64_t)); else std::fill(off, off + 1001, 0); }
      const int64_t n = (int64_t)v.size();
      GlobalPool()->ParallelFor(n, 4096, [&](int64_t b, int64_t e) {
        constexpr int64_t W = 8;                                  // metadata lines 2W ahead, row lines W ahead of the copy
        for (int64_t i = b; i < e; ++i) {
          if (i + 2 * W < e) __builtin_prefetch(meta_.at(v[(size_t)(i + 2 * W)].idx));
          if (rw && i + W < e) { const int32_t r = *(&meta_.at(v[(size_t)(i + W)].idx)->row); if (r >= 0) __builtin_prefetch(rows_.at(r)); }
          const Meta* m = meta_.at(v[(size_t)i].idx);
          if (k) k[i] = v[(size_t)i].key;
          if (f) f[i] = m->freq;
          if (ver) ver[i] = m->version;
          if (rw) memcpy(rw + i * stride_, rows_.at(m->row), stride_ * sizeof(float));
        }
      });
    };
    fill(snap_adm_, snap_adm_off_, keys, rows, freqs, versions, part_offset);
    fill(snap_flt_, snap_flt_off_, fkeys, nullptr, ffreqs, fversions, fpart_offset);
  }
  void SnapshotEnd() { snap_adm_.clear(); snap_adm_.shrink_to_fit(); snap_flt_.clear(); snap_flt_.shrink_to_fit(); }
  void ClearDirty() {
    GlobalPool()->ParallelFor(kv_.NumParts(), 1, [&](int64_t pb, int64_t pe) {
      for (int64_t p = pb; p < pe; ++p) kv_.ForEachInPart((int)p, [&](int64_t, int32_t idx) { *(&meta_.at(idx)->dirty) = 0; });
    });
  }

  // ---- import (restore / elastic import / incremental replay) ---------------------------
  // rows: [n, ncols] (ncols <= stride; missing slot columns take slot_init); rows == nullptr
  // imports filtered (un-admitted) keys.  Only keys with key%1000%part_num == part_id are kept.
  int64_t Import(const int64_t* keys, const float* rows, int64_t ncols, const int64_t* freqs,
                 const int64_t* versions, int64_t n, int part_id, int part_num, int reset_version) {
    // restore of a large table is insert-bound: key ranges go to the workers (inserts are CAS-based, growth is per partition); a key that
    // appears twice in one call keeps one row (the claim below) and the later copy wins or loses arbitrarily, as two restores would
    std::atomic<int64_t> kept{0};
    if (n >= 4096) kv_.Reserve(part_num > 1 ? n / part_num + n / (part_num * 4) : n, [](int64_t cnt, const std::function<void(int64_t, int64_t)>& fn) { GlobalPool()->ParallelFor(cnt, 1, fn); });
    // restore into an empty table: every key of a range is new, so its metadata / row indices are reserved with one fetch_add each
    const bool bulk = kv_.Size() == 0 && part_num <= 1 && n_free_meta_.load(std::memory_order_relaxed) == 0 && n_free_rows_.load(std::memory_order_relaxed) == 0;
    GlobalPool()->ParallelFor(n, 4096, [&](int64_t b, int64_t e) {
      int64_t mine = 0, new_rows = 0;
      Reserve rs;
      if (bulk) {
        rs.meta_next = next_meta_.fetch_add(e - b); rs.meta_end = rs.meta_next + (e - b);
        meta_.EnsureCapacity(rs.meta_end, Meta{0, -1, -1, 0, {0}});
        if (rows) { rs.row_next = next_row_.fetch_add(e - b); rs.row_end = rs.row_next + (e - b); rows_.EnsureCapacity(rs.row_end, 0.f); }
      }
      for (int64_t i = b; i < e; ++i) {
        int64_t key = keys[i];
        if (key == kEmptyKey) continue;
        if (part_num > 1 && dr_ckpt_bucket(key) % part_num != part_id) continue;
        bool inserted = false;
        int32_t idx = kv_.FindOrInsert(key, [this, &rs] { return AllocMeta(&rs); }, &inserted);
        __atomic_store_n(&meta_.at(idx)->freq, freqs ? freqs[i] : 0, __ATOMIC_RELAXED);
        __atomic_store_n(&meta_.at(idx)->version, (reset_version & 1) ? -1 : (versions ? versions[i] : -1), __ATOMIC_RELAXED);
        // reset_version bit 1: the rows are LIVE training state arriving from another tier (device -> host demotion), not a restore: they
        // belong in the next incremental checkpoint, so they carry the dirty bit the evicted device row had
        if (reset_version & 2) __atomic_store_n(&meta_.at(idx)->dirty, (uint8_t)1, __ATOMIC_RELAXED);
        if (rows) {
          int32_t* rp = (&meta_.at(idx)->row);
          int32_t r = __atomic_load_n(rp, __ATOMIC_ACQUIRE);
          if (r == -1) {
            int32_t expect = -1;
            if (__atomic_compare_exchange_n(rp, &expect, -2, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
              r = AllocRow(&rs);
              if (ncols < stride_) InitRow(r, key);                       // missing slot columns take their initial values
              memcpy(rows_.at(r), rows + i * ncols, std::min(ncols, stride_) * sizeof(float));
              __atomic_store_n(rp, r, __ATOMIC_RELEASE); ++new_rows; ++mine;
              continue;
            }
          }
          while ((r = __atomic_load_n(rp, __ATOMIC_ACQUIRE)) < 0) std::this_thread::yield();
          memcpy(rows_.at(r), rows + i * ncols, std::min(ncols, stride_) * sizeof(float));
        }
        ++mine;
      }
      if (new_rows) admitted_.fetch_add(new_rows);
      kept.fetch_add(mine);
    });
    return kept.load();
  }
  // Import into a table that is being READ concurrently (serving: delta update of a live model).  A new key's row is written first and
  // its index published with release semantics; an existing key gets a fresh row (copy-on-write) and the index is swapped -- readers see
  // either the complete old row or the complete new one, never a half-written row.  Replaced rows are not recycled (a reader may still be
  // copying them): they stay allocated until the table is destroyed, i.e. until the next full model update replaces it.
  int64_t ImportCow(const int64_t* keys, const float* rows, int64_t ncols, int64_t n) {
    const int64_t w = std::min(ncols, stride_);
    for (int64_t i = 0; i < n; ++i) {
      if (keys[i] == kEmptyKey) continue;
      bool inserted = false;
      const int32_t idx = kv_.FindOrInsert(keys[i], [this] { return AllocMeta(); }, &inserted);
      int32_t* rp = (&meta_.at(idx)->row);
      const int32_t old = __atomic_load_n(rp, __ATOMIC_ACQUIRE);
      const int32_t r = AllocRow();
      InitRow(r, keys[i]);
      memcpy(rows_.at(r), rows + i * ncols, (size_t)w * sizeof(float));
      __atomic_store_n(rp, r, __ATOMIC_RELEASE);
      if (old < 0) admitted_.fetch_add(1);
    }
    return n;
  }
  // rows (full stride) + metadata of specific keys; found[i] = 1 if the key owns a row (multi-tier promotion path)
  void ExportKeys(const int64_t* keys, int64_t n, float* rows, int64_t* freqs, int64_t* versions, uint8_t* found) {
    GlobalPool()->ParallelFor(n, 1024, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        int32_t idx = kv_.Find(keys[i]);
        int32_t r = idx >= 0 ? RowOf(idx) : -1;
        found[i] = r >= 0;
        freqs[i] = idx >= 0 ? FreqOf(idx) : 0;
        versions[i] = idx >= 0 ? VersionOf(idx) : -1;
        if (r >= 0) memcpy(rows + i * stride_, rows_.at(r), stride_ * sizeof(float));
      }
    });
  }
  CountingBloom* bloom() { return bloom_.get(); }

 private:
  // metadata cells are read by forward lookups while the training thread updates them: all shared accesses are atomic
  // (acquire on the row index pairs with the release store that publishes an initialised row)
  int32_t RowOf(int32_t idx) { return __atomic_load_n((&meta_.at(idx)->row), __ATOMIC_ACQUIRE); }
  int64_t FreqOf(int32_t idx) { return __atomic_load_n((&meta_.at(idx)->freq), __ATOMIC_RELAXED); }
  int64_t VersionOf(int32_t idx) { return __atomic_load_n((&meta_.at(idx)->version), __ATOMIC_RELAXED); }
  const float* DefaultRow(int64_t key) const {
    return default_.data() + dr_default_row(key, std::max<int64_t>(1, cfg_.default_value_dim)) * cfg_.dim;
  }
  void InitRow(int32_t r, int64_t key) {
    float* row = rows_.at(r);
    memcpy(row, DefaultRow(key), cfg_.dim * sizeof(float));
    for (int s = 0; s < cfg_.num_slots; ++s) std::fill(row + (1 + s) * cfg_.dim, row + (2 + s) * cfg_.dim, cfg_.slot_init[s]);
    for (int64_t d = cfg_.dim * (1 + cfg_.num_slots); d < stride_; ++d) row[d] = 0.f;
  }
  int32_t AllocMeta(Reserve* rs = nullptr) {
    if (rs && rs->meta_next < rs->meta_end) { const int64_t i = rs->meta_next++; ResetMeta((int32_t)i); return (int32_t)i; }
    if (n_free_meta_.load(std::memory_order_relaxed) > 0) {      // the free lists are empty unless something was evicted: no lock on the hot path
      std::lock_guard<std::mutex> l(free_mu_);
      if (!free_meta_.empty()) { int32_t i = free_meta_.back(); free_meta_.pop_back(); n_free_meta_.fetch_sub(1, std::memory_order_relaxed); ResetMeta(i); return i; }
    }
    int64_t i = next_meta_.fetch_add(1);
    meta_.EnsureCapacity(i + 1, Meta{0, -1, -1, 0, {0}});
    ResetMeta((int32_t)i);
    return (int32_t)i;
  }
  void ResetMeta(int32_t i) { *(&meta_.at(i)->freq) = 0; *(&meta_.at(i)->version) = -1; *(&meta_.at(i)->row) = -1; *(&meta_.at(i)->dirty) = 0; }
  void FreeMeta(int32_t i) { std::lock_guard<std::mutex> l(free_mu_); free_meta_.push_back(i); n_free_meta_.fetch_add(1, std::memory_order_relaxed); }
  int32_t AllocRow(Reserve* rs = nullptr) {
    if (rs && rs->row_next < rs->row_end) return (int32_t)rs->row_next++;
    if (n_free_rows_.load(std::memory_order_relaxed) > 0) {
      std::lock_guard<std::mutex> l(free_mu_);
      if (!free_rows_.empty()) { int32_t r = free_rows_.back(); free_rows_.pop_back(); n_free_rows_.fetch_sub(1, std::memory_order_relaxed); return r; }
    }
    int64_t r = next_row_.fetch_add(1);
    rows_.EnsureCapacity(r + 1, 0.f);
    return (int32_t)r;
  }
  void FreeRow(int32_t r) { std::lock_guard<std::mutex> l(free_mu_); free_rows_.push_back(r); n_free_rows_.fetch_add(1, std::memory_order_relaxed); }
  void FreeBulk(const std::vector<std::vector<int32_t>>& metas, const std::vector<std::vector<int32_t>>& rows) {
    std::lock_guard<std::mutex> l(free_mu_);
    int64_t nm = 0, nr = 0;
    for (auto& v : metas) { free_meta_.insert(free_meta_.end(), v.begin(), v.end()); nm += (int64_t)v.size(); }
    for (auto& v : rows) { free_rows_.insert(free_rows_.end(), v.begin(), v.end()); nr += (int64_t)v.size(); }
    n_free_meta_.fetch_add(nm, std::memory_order_relaxed); n_free_rows_.fetch_add(nr, std::memory_order_relaxed);
    admitted_.fetch_sub(nr);
  }

  DrEvConfig cfg_;
  int64_t stride_;
  HostKV kv_;
  // One 32-byte record per key: a probe that found the key touches ONE cache line for frequency, version, row index and dirty flag
  // (four parallel arrays cost four DRAM misses per key on tables larger than the caches -- the same AoS lesson as the device table's
  // 32-byte slot).
  struct alignas(32) Meta { int64_t freq; int64_t version; int32_t row; uint8_t dirty; uint8_t pad[11]; };
  ChunkedArray<Meta> meta_;
  ChunkedArray<float, 12> rows_;       // 4096 rows per chunk
  std::vector<float> default_;
  std::unique_ptr<CountingBloom> bloom_;
  alignas(64) std::atomic<int64_t> next_meta_{0};
  alignas(64) std::atomic<int64_t> next_row_{0};
  alignas(64) std::atomic<int64_t> admitted_{0};
  std::mutex free_mu_;
  std::vector<int32_t> free_meta_, free_rows_;
  std::atomic<int64_t> n_free_meta_{0}, n_free_rows_{0};
  std::vector<SnapItem> snap_adm_, snap_flt_;
  std::vector<int64_t> snap_adm_off_, snap_flt_off_;      // 1001 checkpoint-bucket offsets of the current snapshot
};

}  // namespace dr

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

void* dr_host_ev_create(const DrEvConfig* cfg) { return new dr::HostEV(*cfg); }
void dr_host_ev_destroy(void* h) { delete static_cast<dr::HostEV*>(h); }
int64_t dr_host_ev_stride(void* h) { return static_cast<dr::HostEV*>(h)->stride(); }
void dr_host_ev_set_default(void* h, const float* m) { static_cast<dr::HostEV*>(h)->SetDefault(m); }
int64_t dr_host_ev_size(void* h) { return static_cast<dr::HostEV*>(h)->Size(); }
int64_t dr_host_ev_total_keys(void* h) { return static_cast<dr::HostEV*>(h)->TotalKeys(); }
void dr_host_ev_lookup(void* h, const int64_t* keys, int64_t n, float* out) { static_cast<dr::HostEV*>(h)->Lookup(keys, n, out); }
void dr_host_ev_lookup_slot(void* h, const int64_t* keys, int64_t n, int slot, float* out) { static_cast<dr::HostEV*>(h)->LookupSlot(keys, n, slot, out); }
void dr_host_ev_get_freq(void* h, const int64_t* keys, int64_t n, int64_t* out) { static_cast<dr::HostEV*>(h)->GetFreq(keys, n, out); }
void dr_host_ev_get_version(void* h, const int64_t* keys, int64_t n, int64_t* out) { static_cast<dr::HostEV*>(h)->GetVersion(keys, n, out); }
void dr_host_ev_apply(void* h, const int64_t* keys, const float* grads, const int64_t* counts, int64_t n, const DrOptHyper* hp) {
  static_cast<dr::HostEV*>(h)->Apply(keys, grads, counts, n, *hp);
}
int64_t dr_host_ev_shrink(void* h, int64_t step) { return static_cast<dr::HostEV*>(h)->Shrink(step); }
int64_t dr_host_ev_remove(void* h, const int64_t* keys, int64_t n) { return static_cast<dr::HostEV*>(h)->Remove(keys, n); }
void dr_host_ev_snapshot_begin(void* h, int dirty_only, int part_id, int part_num, int64_t* na, int64_t* nf) {
  static_cast<dr::HostEV*>(h)->SnapshotBegin(dirty_only, part_id, part_num, na, nf);
}
void dr_host_ev_snapshot_read(void* h, int64_t* keys, float* rows, int64_t* freqs, int64_t* versions, int64_t* poff,
                              int64_t* fkeys, int64_t* ffreqs, int64_t* fversions, int64_t* fpoff) {
  static_cast<dr::HostEV*>(h)->SnapshotRead(keys, rows, freqs, versions, poff, fkeys, ffreqs, fversions, fpoff);
}
void dr_host_ev_snapshot_end(void* h) { static_cast<dr::HostEV*>(h)->SnapshotEnd(); }
void dr_host_ev_clear_dirty(void* h) { static_cast<dr::HostEV*>(h)->ClearDirty(); }
int64_t dr_host_ev_import(void* h, const int64_t* keys, const float* rows, int64_t ncols, const int64_t* freqs,
                          const int64_t* versions, int64_t n, int part_id, int part_num, int reset_version) {
  return static_cast<dr::HostEV*>(h)->Import(keys, rows, ncols, freqs, versions, n, part_id, part_num, reset_version);
}
int64_t dr_host_ev_import_cow(void* h, const int64_t* keys, const float* rows, int64_t ncols, int64_t n) {
  return static_cast<dr::HostEV*>(h)->ImportCow(keys, rows, ncols, n);
}
void dr_host_ev_export_keys(void* h, const int64_t* keys, int64_t n, float* rows, int64_t* freqs, int64_t* versions, uint8_t* found) {
  static_cast<dr::HostEV*>(h)->ExportKeys(keys, n, rows, freqs, versions, found);
}
int64_t dr_host_bloom_info(void* h, int64_t* k, int64_t* m, int64_t* bits) {
  auto* b = static_cast<dr::HostEV*>(h)->bloom();
  if (!b) return 0;
  *k = b->k(); *m = b->m(); *bits = b->bits();
  return b->nbytes();
}
void dr_host_bloom_read(void* h, uint8_t* out) {
  auto* b = static_cast<dr::HostEV*>(h)->bloom();
  if (b) memcpy(out, b->raw(), b->nbytes());
}
void dr_host_bloom_write(void* h, const uint8_t* in) {
  auto* b = static_cast<dr::HostEV*>(h)->bloom();
  if (b) memcpy(b->raw(), in, b->nbytes());
}

// ---- host unique-with-counts (optimizer.py:91 dedup of sparse grads; unique_ali_op.cc) -----
// out_unique[nu], out_inverse[n], out_counts[nu]; returns nu.  Order = first occurrence.
int64_t dr_host_unique(const int64_t* keys, int64_t n, int64_t* out_unique, int64_t* out_inverse, int64_t* out_counts) {
  int64_t cap = 16; while (cap < n * 2) cap <<= 1;
  std::vector<int64_t> tk(cap, dr::kEmptyKey); std::vector<int64_t> tv(cap, -1);
  uint64_t mask = cap - 1; int64_t nu = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t key = keys[i]; uint64_t pos = dr_mix64((uint64_t)key) & mask;
    for (;;) {
      if (tk[pos] == key && tv[pos] >= 0) { out_inverse[i] = tv[pos]; out_counts[tv[pos]]++; break; }
      if (tv[pos] < 0) { tk[pos] = key; tv[pos] = nu; out_unique[nu] = key; out_counts[nu] = 1; out_inverse[i] = nu; ++nu; break; }
      pos = (pos + 1) & mask;
    }
  }
  return nu;
}
// segment-sum of per-occurrence grads into per-unique grads
void dr_host_segment_sum(const float* grads, const int64_t* inverse, int64_t n, int64_t dim, float* out, int64_t nu) {
  memset(out, 0, sizeof(float) * nu * dim);
  for (int64_t i = 0; i < n; ++i) {
    float* o = out + inverse[i] * dim; const float* g = grads + i * dim;
    for (int64_t d = 0; d < dim; ++d) o[d] += g[d];
  }
}

// ---- fused entry points for the framework's CPU path (one native call per step instead of 3-4 per table) ----------------------
}  // extern "C"  (templates below need C++ linkage)
namespace {
// dedup (first-occurrence order) + per-unique gradient sums of one table; grads row i at grads + i * row_stride
struct DedupScratch { std::vector<int64_t> uniq, inv, cnt; std::vector<float> gsum; int64_t nu = 0; };
// GradAt(i) -> pointer to the gradient row of occurrence i (strided tensor, or one row shared by a bag of ids)
template <class GradAt>
void DedupAndSumSerialT(const int64_t* ids, int64_t n, GradAt grad_at, int64_t dim, DedupScratch* s) {
  s->uniq.resize(n); s->inv.resize(n); s->cnt.resize(n);
  s->nu = n ? dr_host_unique(ids, n, s->uniq.data(), s->inv.data(), s->cnt.data()) : 0;
  s->gsum.assign((size_t)(s->nu * dim), 0.f);
  for (int64_t i = 0; i < n; ++i) {
    float* o = s->gsum.data() + s->inv[i] * dim; const float* g = grad_at(i);
    for (int64_t d = 0; d < dim; ++d) o[d] += g[d];
  }
}

// Large batches (sequence models push B x L ids into one table): bucket the occurrences by key hash, then every bucket is de-duplicated
// and summed by its own thread (a key lives in exactly one bucket, so buckets are independent); results are concatenated.
// The unique order differs from first-occurrence order, which no consumer depends on.
template <class GradAt>
void DedupAndSumT(const int64_t* ids, int64_t n, GradAt grad_at, int64_t dim, DedupScratch* s, bool allow_parallel = true) {
  const int nb = std::min<int>(64, (dr::GlobalPool()->size() + 1) * 4);
  if (!allow_parallel || n < 16384 || nb < 4) { DedupAndSumSerialT(ids, n, grad_at, dim, s); return; }
  std::vector<int32_t> bucket_of((size_t)n), order((size_t)n);
  std::vector<int64_t> start((size_t)nb + 1, 0);
  dr::GlobalPool()->ParallelFor(n, 8192, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) bucket_of[(size_t)i] = (int32_t)((dr_mix64((uint64_t)ids[i]) >> 33) % (uint64_t)nb);
  });
  for (int64_t i = 0; i < n; ++i) start[(size_t)bucket_of[(size_t)i] + 1]++;
  for (int k = 0; k < nb; ++k) start[(size_t)k + 1] += start[(size_t)k];
  { std::vector<int64_t> cur(start.begin(), start.end() - 1); for (int64_t i = 0; i < n; ++i) order[(size_t)cur[(size_t)bucket_of[(size_t)i]]++] = (int32_t)i; }
  std::vector<DedupScratch> parts((size_t)nb);
  dr::GlobalPool()->ParallelFor(nb, 1, [&](int64_t b, int64_t e) {
    std::vector<int64_t> keys;
    for (int64_t k = b; k < e; ++k) {
      const int64_t lo = start[(size_t)k], m = start[(size_t)k + 1] - lo;
      DedupScratch& P = parts[(size_t)k];
      keys.resize((size_t)m);
      for (int64_t j = 0; j < m; ++j) keys[(size_t)j] = ids[order[(size_t)(lo + j)]];
      P.uniq.resize((size_t)m); P.inv.resize((size_t)m); P.cnt.resize((size_t)m);
      P.nu = m ? dr_host_unique(keys.data(), m, P.uniq.data(), P.inv.data(), P.cnt.data()) : 0;
      P.gsum.assign((size_t)(P.nu * dim), 0.f);
      for (int64_t j = 0; j < m; ++j) {
        float* o = P.gsum.data() + P.inv[(size_t)j] * dim; const float* g = grad_at((int64_t)order[(size_t)(lo + j)]);
        for (int64_t d = 0; d < dim; ++d) o[d] += g[d];
      }
    }
  });
  int64_t nu = 0;
  for (auto& P : parts) nu += P.nu;
  s->nu = nu; s->uniq.resize((size_t)nu); s->cnt.resize((size_t)nu); s->gsum.resize((size_t)(nu * dim));
  int64_t off = 0;
  for (auto& P : parts) {
    if (P.nu) {
      memcpy(s->uniq.data() + off, P.uniq.data(), sizeof(int64_t) * (size_t)P.nu);
      memcpy(s->cnt.data() + off, P.cnt.data(), sizeof(int64_t) * (size_t)P.nu);
      memcpy(s->gsum.data() + off * dim, P.gsum.data(), sizeof(float) * (size_t)(P.nu * dim));
    }
    off += P.nu;
  }
}

void DedupAndSum(const int64_t* ids, int64_t n, const float* grads, int64_t row_stride, int64_t dim, DedupScratch* s, bool allow_parallel = true) {
  DedupAndSumT(ids, n, [=](int64_t i) { return grads + i * row_stride; }, dim, s, allow_parallel);
}
}  // namespace
extern "C" {

// unique + segment-sum + apply of ONE table in one call; grads row i at grads + i * row_stride (strided views need no copy)
void dr_host_ev_apply_raw(void* h, const int64_t* ids, int64_t n, const float* grads, int64_t row_stride, const DrOptHyper* hp) {
  auto* ev = static_cast<dr::HostEV*>(h);
  DedupScratch s;
  DedupAndSum(ids, n, grads, row_stride, ev->cfg().dim, &s);
  ev->Apply(s.uniq.data(), s.gsum.data(), s.cnt.data(), s.nu, *hp);
}

// ONE de-duplicated apply for everything a table received in a step: `nseg` segments, segment s = n[s] ids whose occurrence i uses the
// gradient row grads[s] + (i / group[s]) * row_stride[s]  (group 1: one row per occurrence -- a plain lookup; group L: the bag of L ids
// of sample i / L shares one row -- a sum-pooled lookup, whose [B*L, dim] per-occurrence gradient is therefore never materialised).
void dr_host_ev_apply_multi(void* h, int nseg, const int64_t* const* ids, const int64_t* n, const float* const* grads, const int64_t* row_stride,
                            const int64_t* group, const DrOptHyper* hp) {
  auto* ev = static_cast<dr::HostEV*>(h);
  int64_t total = 0;
  for (int s = 0; s < nseg; ++s) total += n[s];
  if (total == 0) return;
  std::vector<int64_t> all((size_t)total);
  std::vector<const float*> gp((size_t)total);
  int64_t o = 0;
  for (int s = 0; s < nseg; ++s) {
    const int64_t g = group[s] > 0 ? group[s] : 1;
    memcpy(all.data() + o, ids[s], sizeof(int64_t) * (size_t)n[s]);
    for (int64_t i = 0; i < n[s]; ++i) gp[(size_t)(o + i)] = grads[s] + (i / g) * row_stride[s];
    o += n[s];
  }
  DedupScratch sc;
  DedupAndSumT(all.data(), total, [&](int64_t i) { return gp[(size_t)i]; }, ev->cfg().dim, &sc);
  ev->Apply(sc.uniq.data(), sc.gsum.data(), sc.cnt.data(), sc.nu, *hp);
}

// Sum-pooled lookup of a dense [B, L] id tensor (PAD_KEY = unused position): out[b] = sum of the rows of sample b's ids, written at
// out + b * out_stride.  The [B, L, dim] intermediate of lookup + mask + sum never exists.
void dr_host_ev_lookup_pooled(void* h, const int64_t* ids, int64_t B, int64_t L, float* out, int64_t out_stride) {
  auto* ev = static_cast<dr::HostEV*>(h);
  dr::GlobalPool()->ParallelFor(B, 64, [&](int64_t b0, int64_t b1) { ev->LookupPooledRange(ids, b0, b1, L, out, out_stride); });
}

// T tables of equal dim, one id per (table, sample): keys feature-major [T][B] -> out sample-major [B][T][dim]
// (the layout the interaction layers consume: no per-table tensors, no stack).  Parallel over (table, key-chunk) tiles.
void dr_host_group_lookup(void** hs, int T, const int64_t* keys, int64_t B, float* out) {
  if (T <= 0 || B <= 0) return;
  const int64_t dim = static_cast<dr::HostEV*>(hs[0])->cfg().dim;
  const int64_t chunk = 1024, chunks = (B + chunk - 1) / chunk;
  if ((int64_t)T * B < 2048) {                     // online-serving sized requests: a parallel region would cost more than the probes
    for (int t = 0; t < T; ++t) static_cast<dr::HostEV*>(hs[t])->LookupRange(keys + (int64_t)t * B, 0, B, out + (int64_t)t * dim, (int64_t)T * dim);
    return;
  }
  dr::GlobalPool()->ParallelFor((int64_t)T * chunks, 1, [&](int64_t b, int64_t e) {
    for (int64_t w = b; w < e; ++w) {
      const int64_t t = w / chunks, c = w % chunks, lo = c * chunk, hi = std::min(B, lo + chunk);
      static_cast<dr::HostEV*>(hs[t])->LookupRange(keys + t * B, lo, hi, out + t * dim, (int64_t)T * dim);
    }
  });
}

// The matching update: ids [T][B], grads [B][T][dim] (as produced by autograd for the lookup above).  Tables are independent, so
// with many tables each one is processed serially on its own worker (dedup + sums + apply); with few tables they run one after
// another, each using the whole pool.
void dr_host_group_apply_raw(void** hs, int T, const int64_t* ids, int64_t B, const float* grads, const DrOptHyper* hp) {
  if (T <= 0 || B <= 0) return;
  const int64_t dim = static_cast<dr::HostEV*>(hs[0])->cfg().dim, row_stride = (int64_t)T * dim;
  if (T >= dr::GlobalPool()->size() + 1) {
    dr::GlobalPool()->ParallelFor(T, 1, [&](int64_t b, int64_t e) {
      DedupScratch s;
      for (int64_t t = b; t < e; ++t) {
        auto* ev = static_cast<dr::HostEV*>(hs[t]);
        DedupAndSum(ids + t * B, B, grads + t * dim, row_stride, dim, &s, /*allow_parallel=*/false);   // already inside a parallel region
        ev->ApplyRange(s.uniq.data(), s.gsum.data(), s.cnt.data(), 0, s.nu, *hp);
      }
    });
  } else {
    for (int t = 0; t < T; ++t) dr_host_ev_apply_raw(hs[t], ids + (int64_t)t * B, B, grads + (int64_t)t * dim, row_stride, hp);
  }
}

// ---- DLRM dot interaction on the CPU path (modelzoo/dlrm/train.py:121-133) ------------------------------------------------------
// feats = [dense | embs] ([F = T + 1, D] per sample); out = [dense | strict lower triangle of feats featsT, row-major (i, j < i)].
// The composite PyTorch form (cat + bmm + advanced-index gather, and an index_put in the backward) moves ~10x the bytes it needs;
// here each sample is handled in cache by one thread.
}  // extern "C"  (templates below need C++ linkage)

namespace {
constexpr int kMaxF = 64;      // features per sample handled by the register-blocked path (T + 1 <= 64)

// fwd: gram rows as small GEMV against the transposed features (vectorised over j, no horizontal sums)
template <int D>
void DotFwdRange(const float* __restrict dense, const float* __restrict embs, int64_t b0, int64_t b1, int T, float* __restrict out) {
  const int F = T + 1, P = F * (F - 1) / 2, Fp = (F + 7) & ~7;
  alignas(64) float ft[D * kMaxF];
  alignas(64) float acc[4 * kMaxF];
  for (int64_t b = b0; b < b1; ++b) {
    const float* d = dense + b * D; const float* e = embs + b * (int64_t)T * D;
    float* o = out + b * (int64_t)(D + P);
    for (int k = 0; k < D; ++k) {                      // ft[k][j] = feat[j][k]
      float* r = ft + k * Fp;
      r[0] = d[k];
      for (int j = 1; j < F; ++j) r[j] = e[(int64_t)(j - 1) * D + k];
      for (int j = F; j < Fp; ++j) r[j] = 0.f;
    }
    for (int k = 0; k < D; ++k) o[k] = d[k];
    float* z = o + D;
    // four gram rows per pass: one load of ft[k] feeds four independent accumulator rows (one row alone is a chain of D dependent FMAs)
    for (int i0 = 1; i0 < F; i0 += 4) {
      const int ib = std::min(4, F - i0);
      const int n = std::min(Fp, (i0 + ib - 1 + 7) & ~7);            // row i needs columns j < i; the block needs j < i0 + ib - 1
      const float* f0 = e + (int64_t)(i0 - 1) * D;
      const float* f1 = e + (int64_t)(std::min(i0 + 1, F - 1) - 1) * D; const float* f2 = e + (int64_t)(std::min(i0 + 2, F - 1) - 1) * D; const float* f3 = e + (int64_t)(std::min(i0 + 3, F - 1) - 1) * D;
      for (int r = 0; r < 4; ++r) for (int j = 0; j < n; ++j) acc[r * kMaxF + j] = 0.f;
      for (int k = 0; k < D; ++k) {
        const float a0 = f0[k], a1 = f1[k], a2 = f2[k], a3 = f3[k]; const float* r = ft + k * Fp;
        for (int j = 0; j < n; ++j) { const float v = r[j]; acc[j] += a0 * v; acc[kMaxF + j] += a1 * v; acc[2 * kMaxF + j] += a2 * v; acc[3 * kMaxF + j] += a3 * v; }
      }
      for (int r = 0; r < ib; ++r) for (int j = 0; j < i0 + r; ++j) *z++ = acc[r * kMaxF + j];
    }
  }
}

// bwd: dfeat[i] = sum_j S[i][j] feat[j] with S the symmetric zero-diagonal matrix built from dz (vectorised over D)
template <int D>
void DotBwdRange(const float* __restrict dense, const float* __restrict embs, const float* __restrict dz, int64_t b0, int64_t b1, int T,
                 float* __restrict ddense, float* __restrict dembs) {
  const int F = T + 1, P = F * (F - 1) / 2;
  alignas(64) float S[kMaxF * kMaxF];
  alignas(64) float feat[kMaxF * D];
  for (int64_t b = b0; b < b1; ++b) {
    const float* d = dense + b * D; const float* e = embs + b * (int64_t)T * D;
    const float* g = dz + b * (int64_t)(D + P); const float* gz = g + D;
    for (int k = 0; k < D; ++k) feat[k] = d[k];
    for (int j = 1; j < F; ++j) for (int k = 0; k < D; ++k) feat[j * D + k] = e[(int64_t)(j - 1) * D + k];
    for (int i = 0; i < F; ++i) S[i * F + i] = 0.f;
    for (int i = 1; i < F; ++i) for (int j = 0; j < i; ++j) { const float w = *gz++; S[i * F + j] = w; S[j * F + i] = w; }
    // four output rows at a time: one load of feat[j] feeds four independent accumulator rows (a single row is a chain of F dependent FMAs)
    constexpr int IB = 4;
    for (int i0 = 0; i0 < F; i0 += IB) {
      const int ib = std::min(IB, F - i0);
      float accb[IB][D];
      for (int r = 0; r < IB; ++r) for (int k = 0; k < D; ++k) accb[r][k] = (i0 + r == 0) ? g[k] : 0.f;   // feature 0 also receives the pass-through gradient
      const float* s0 = S + (size_t)i0 * F;
      const float* s1 = S + (size_t)std::min(i0 + 1, F - 1) * F; const float* s2 = S + (size_t)std::min(i0 + 2, F - 1) * F; const float* s3 = S + (size_t)std::min(i0 + 3, F - 1) * F;
      for (int j = 0; j < F; ++j) {
        const float* fj = feat + j * D;
        const float w0 = s0[j], w1 = s1[j], w2 = s2[j], w3 = s3[j];
        for (int k = 0; k < D; ++k) { const float f = fj[k]; accb[0][k] += w0 * f; accb[1][k] += w1 * f; accb[2][k] += w2 * f; accb[3][k] += w3 * f; }
      }
      for (int r = 0; r < ib; ++r) {
        const int i = i0 + r;
        float* dst = i == 0 ? ddense + b * D : dembs + b * (int64_t)T * D + (int64_t)(i - 1) * D;
        for (int k = 0; k < D; ++k) dst[k] = accb[r][k];
      }
    }
  }
}

// generic (any D, any T) scalar fallbacks
void DotFwdGeneric(const float* dense, const float* embs, int64_t b0, int64_t b1, int T, int D, float* out) {
  const int F = T + 1, P = F * (F - 1) / 2;
  for (int64_t b = b0; b < b1; ++b) {
    const float* d = dense + b * D; const float* e = embs + b * (int64_t)T * D;
    float* o = out + b * (int64_t)(D + P);
    memcpy(o, d, sizeof(float) * D);
    float* z = o + D;
    for (int i = 1; i < F; ++i) {
      const float* fi = e + (int64_t)(i - 1) * D;
      for (int j = 0; j < i; ++j) {
        const float* fj = j == 0 ? d : e + (int64_t)(j - 1) * D;
        float a = 0.f;
        for (int k = 0; k < D; ++k) a += fi[k] * fj[k];
        *z++ = a;
      }
    }
  }
}
void DotBwdGeneric(const float* dense, const float* embs, const float* dz, int64_t b0, int64_t b1, int T, int D, float* ddense, float* dembs) {
  const int F = T + 1, P = F * (F - 1) / 2;
  for (int64_t b = b0; b < b1; ++b) {
    const float* d = dense + b * D; const float* e = embs + b * (int64_t)T * D;
    const float* g = dz + b * (int64_t)(D + P);
    float* dd = ddense + b * D; float* de = dembs + b * (int64_t)T * D;
    memcpy(dd, g, sizeof(float) * D);
    memset(de, 0, sizeof(float) * (size_t)T * D);
    const float* gz = g + D;
    for (int i = 1; i < F; ++i) {
      const float* fi = e + (int64_t)(i - 1) * D; float* gi = de + (int64_t)(i - 1) * D;
      for (int j = 0; j < i; ++j) {
        const float w = *gz++;
        const float* fj = j == 0 ? d : e + (int64_t)(j - 1) * D;
        float* gj = j == 0 ? dd : de + (int64_t)(j - 1) * D;
        for (int k = 0; k < D; ++k) { gi[k] += w * fj[k]; gj[k] += w * fi[k]; }
      }
    }
  }
}
}  // namespace

extern "C" {

void dr_host_dot_interaction_fwd(const float* dense, const float* embs, int64_t B, int T, int D, float* out) {
  const bool blocked = T + 1 <= kMaxF;
  dr::GlobalPool()->ParallelFor(B, 64, [&](int64_t b0, int64_t b1) {
    if (blocked && D == 16) DotFwdRange<16>(dense, embs, b0, b1, T, out);
    else if (blocked && D == 8) DotFwdRange<8>(dense, embs, b0, b1, T, out);
    else if (blocked && D == 32) DotFwdRange<32>(dense, embs, b0, b1, T, out);
    else if (blocked && D == 64) DotFwdRange<64>(dense, embs, b0, b1, T, out);
    else DotFwdGeneric(dense, embs, b0, b1, T, D, out);
  });
}
// dz [B, D + P] -> ddense [B, D], dembs [B, T, D]
void dr_host_dot_interaction_bwd(const float* dense, const float* embs, const float* dz, int64_t B, int T, int D, float* ddense, float* dembs) {
  const bool blocked = T + 1 <= kMaxF;
  dr::GlobalPool()->ParallelFor(B, 64, [&](int64_t b0, int64_t b1) {
    if (blocked && D == 16) DotBwdRange<16>(dense, embs, dz, b0, b1, T, ddense, dembs);
    else if (blocked && D == 8) DotBwdRange<8>(dense, embs, dz, b0, b1, T, ddense, dembs);
    else if (blocked && D == 32) DotBwdRange<32>(dense, embs, dz, b0, b1, T, ddense, dembs);
    else if (blocked && D == 64) DotBwdRange<64>(dense, embs, dz, b0, b1, T, ddense, dembs);
    else DotBwdGeneric(dense, embs, dz, b0, b1, T, D, ddense, dembs);
  });
}

int dr_host_num_threads() { return dr::GlobalPool()->size() + 1; }

}  // extern "C"
