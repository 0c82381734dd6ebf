// Host IO runtime (C++17): checkpoint tensor bundle, bounded staging queue (TensorBuffer /
// PrefetchRunner analogue), elastic WorkQueue, synthetic Criteo/Taobao batch generators.
//
// Parity map to the reference:
//   BundleWriter/Reader streaming of EV tensors     framework/embedding/embedding_var_ckpt_data.cc:156-246 (8 MiB buffer)
//   TensorBufferPut/Take/Cancel/Close/Size          core/kernels/tensor_buffer_ops.{h,cc}:94,124
//   PrefetchRunner                                  cc/training/prefetch_runner.{h,cc}
//   WorkQueue ops                                   core/kernels/work_queue_ops.cc, python/ops/work_queue.py:113-598
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../common/bundle.h"

namespace dr {

// ---------------------------------------------------------------------------------------
// Bounded staging queue of tickets (the payload tensors stay owned by the framework layer).
// put blocks while full, take blocks while empty; close wakes everyone; cancel drops pending.
// ---------------------------------------------------------------------------------------
class StagingQueue {
 public:
  explicit StagingQueue(int64_t capacity) : cap_(std::max<int64_t>(1, capacity)) {}
  // 0 ok, 1 timeout, 2 closed/cancelled
  int Put(int64_t ticket, int64_t timeout_ms) {
    std::unique_lock<std::mutex> l(mu_);
    auto pred = [&] { return closed_ || cancelled_ || (int64_t)q_.size() < cap_; };
    if (timeout_ms < 0) not_full_.wait(l, pred);
    else if (!not_full_.wait_for(l, std::chrono::milliseconds(timeout_ms), pred)) return 1;
    if (closed_ || cancelled_) return 2;
    q_.push_back(ticket);
    not_empty_.notify_one();
    return 0;
  }
  int Take(int64_t* ticket, int64_t timeout_ms) {
    std::unique_lock<std::mutex> l(mu_);
    auto pred = [&] { return !q_.empty() || closed_ || cancelled_; };
    if (timeout_ms < 0) not_empty_.wait(l, pred);
    else if (!not_empty_.wait_for(l, std::chrono::milliseconds(timeout_ms), pred)) return 1;
    if (q_.empty()) return 2;
    *ticket = q_.front(); q_.pop_front();
    not_full_.notify_one();
    return 0;
  }
  void Close() { std::lock_guard<std::mutex> l(mu_); closed_ = true; not_empty_.notify_all(); not_full_.notify_all(); }
  // cancel: reject producers and drop staged items until Resume (TensorBufferCancel semantics)
  int64_t Cancel(int64_t* dropped, int64_t max_out) {
    std::lock_guard<std::mutex> l(mu_);
    cancelled_ = true; int64_t n = 0;
    while (!q_.empty()) { if (n < max_out) dropped[n] = q_.front(); ++n; q_.pop_front(); }
    not_full_.notify_all(); not_empty_.notify_all();
    return n;
  }
  void Resume() { std::lock_guard<std::mutex> l(mu_); cancelled_ = false; }
  int64_t Size() { std::lock_guard<std::mutex> l(mu_); return (int64_t)q_.size(); }
  bool closed() { std::lock_guard<std::mutex> l(mu_); return closed_; }
 private:
  int64_t cap_; std::deque<int64_t> q_; std::mutex mu_; std::condition_variable not_full_, not_empty_;
  bool closed_ = false, cancelled_ = false;
};

// ---------------------------------------------------------------------------------------
// WorkQueue: global list of work items, epochs, optional shuffle, resumable position.
// ---------------------------------------------------------------------------------------
class WorkQueue {
 public:
  WorkQueue(std::vector<std::string> items, int64_t num_epochs, bool shuffle, uint64_t seed)
      : items_(std::move(items)), num_epochs_(num_epochs), shuffle_(shuffle), seed_(seed) { StartEpoch(); }
  // returns item or empty + done=true when all epochs are consumed
  bool Take(std::string* out) {
    std::lock_guard<std::mutex> l(mu_);
    for (;;) {
      if (pos_ < (int64_t)order_.size()) { *out = items_[order_[pos_++]]; ++taken_; return true; }
      if (num_epochs_ > 0 && epoch_ + 1 >= num_epochs_) return false;
      ++epoch_; StartEpoch();
      if (items_.empty()) return false;
    }
  }
  void Add(const std::string& item) { std::lock_guard<std::mutex> l(mu_); items_.push_back(item); order_.push_back((int64_t)items_.size() - 1); }
  void State(int64_t* epoch, int64_t* pos, int64_t* taken) { std::lock_guard<std::mutex> l(mu_); *epoch = epoch_; *pos = pos_; *taken = taken_; }
  void Restore(int64_t epoch, int64_t pos) { std::lock_guard<std::mutex> l(mu_); epoch_ = epoch; StartEpoch(); pos_ = std::min<int64_t>(pos, order_.size()); }
  int64_t Remaining() { std::lock_guard<std::mutex> l(mu_); return (int64_t)order_.size() - pos_; }
 private:
  void StartEpoch() {
    order_.resize(items_.size());
    for (size_t i = 0; i < order_.size(); ++i) order_[i] = (int64_t)i;
    if (shuffle_) { std::mt19937_64 g(seed_ + (uint64_t)epoch_ * 0x9e3779b97f4a7c15ULL); std::shuffle(order_.begin(), order_.end(), g); }
    pos_ = 0;
  }
  std::vector<std::string> items_; std::vector<int64_t> order_;
  int64_t num_epochs_, epoch_ = 0, pos_ = 0, taken_ = 0; bool shuffle_; uint64_t seed_; std::mutex mu_;
};

// ---------------------------------------------------------------------------------------
// Synthetic generators (no datasets in the sandbox): Criteo-shaped click logs and
// Taobao-shaped behaviour sequences.  Ids follow a truncated power law so that the dedup /
// admission / cache paths see realistic skew.
// ---------------------------------------------------------------------------------------
struct XorShift { uint64_t s; explicit XorShift(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ULL + 0x1234567ULL) {}
  uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); } };

static inline int64_t PowerLawId(XorShift& r, int64_t card, double alpha) {
  // inverse-CDF sample of p(x) ~ x^-alpha on [1, card]; alpha == 0 -> uniform
  double u = r.uni();
  if (alpha <= 0.0) return (int64_t)(u * card);
  double x;
  if (std::fabs(alpha - 1.0) < 1e-9) x = std::pow((double)card, u);
  else { double a = 1.0 - alpha; x = std::pow(u * (std::pow((double)card, a) - 1.0) + 1.0, 1.0 / a); }
  int64_t id = (int64_t)x - 1; if (id < 0) id = 0; if (id >= card) id = card - 1;
  // scatter the rank so hot ids are not numerically adjacent
  return (int64_t)((uint64_t)id * 0x9E3779B97F4A7C15ULL % (uint64_t)card);
}

}  // namespace dr

extern "C" {

// ---- bundle ---------------------------------------------------------------------------
void* dr_bundle_writer_open(const char* prefix) { auto* w = new dr::BundleWriter(prefix); if (!w->ok()) { delete w; return nullptr; } return w; }
int dr_bundle_writer_add(void* w, const char* name, const char* dtype, const int64_t* shape, int ndim, const void* data, int64_t nbytes) {
  return static_cast<dr::BundleWriter*>(w)->Add(name, dtype, shape, ndim, data, nbytes);
}
int dr_bundle_writer_close(void* w) { auto* p = static_cast<dr::BundleWriter*>(w); int rc = p->Close(); delete p; return rc; }
void* dr_bundle_reader_open(const char* prefix) { auto* r = new dr::BundleReader(prefix); if (!r->ok()) { delete r; return nullptr; } return r; }
void dr_bundle_reader_close(void* r) { delete static_cast<dr::BundleReader*>(r); }
int64_t dr_bundle_reader_count(void* r) { return (int64_t)static_cast<dr::BundleReader*>(r)->entries().size(); }
// writes name/dtype into caller buffers; shape into shape[8]; returns ndim or -1
int dr_bundle_reader_entry(void* r, int64_t i, char* name, int name_cap, char* dtype, int dtype_cap, int64_t* shape, int64_t* nbytes) {
  auto& es = static_cast<dr::BundleReader*>(r)->entries();
  if (i < 0 || i >= (int64_t)es.size()) return -1;
  auto& e = es[i];
  snprintf(name, name_cap, "%s", e.name.c_str()); snprintf(dtype, dtype_cap, "%s", e.dtype.c_str());
  for (size_t d = 0; d < e.shape.size() && d < 8; ++d) shape[d] = e.shape[d];
  *nbytes = e.nbytes;
  return (int)e.shape.size();
}
int dr_bundle_reader_read(void* r, const char* name, void* dst, int64_t dst_bytes, int verify) {
  auto* br = static_cast<dr::BundleReader*>(r);
  auto* e = br->Find(name);
  if (!e) return -10;
  if (e->nbytes > dst_bytes) return -11;
  return br->Read(*e, dst, verify);
}

// ---- staging queue ----------------------------------------------------------------------
void* dr_stage_create(int64_t capacity) { return new dr::StagingQueue(capacity); }
void dr_stage_destroy(void* q) { delete static_cast<dr::StagingQueue*>(q); }
int dr_stage_put(void* q, int64_t ticket, int64_t timeout_ms) { return static_cast<dr::StagingQueue*>(q)->Put(ticket, timeout_ms); }
int dr_stage_take(void* q, int64_t* ticket, int64_t timeout_ms) { return static_cast<dr::StagingQueue*>(q)->Take(ticket, timeout_ms); }
void dr_stage_close(void* q) { static_cast<dr::StagingQueue*>(q)->Close(); }
int64_t dr_stage_cancel(void* q, int64_t* dropped, int64_t max_out) { return static_cast<dr::StagingQueue*>(q)->Cancel(dropped, max_out); }
void dr_stage_resume(void* q) { static_cast<dr::StagingQueue*>(q)->Resume(); }
int64_t dr_stage_size(void* q) { return static_cast<dr::StagingQueue*>(q)->Size(); }

// ---- work queue -------------------------------------------------------------------------
void* dr_wq_create(const char* items_nl, int64_t num_epochs, int shuffle, uint64_t seed) {
  std::vector<std::string> items; std::string cur;
  for (const char* p = items_nl; *p; ++p) { if (*p == '\n') { if (!cur.empty()) items.push_back(cur); cur.clear(); } else cur.push_back(*p); }
  if (!cur.empty()) items.push_back(cur);
  return new dr::WorkQueue(std::move(items), num_epochs, shuffle != 0, seed);
}
void dr_wq_destroy(void* q) { delete static_cast<dr::WorkQueue*>(q); }
int dr_wq_take(void* q, char* out, int cap) {
  std::string s; if (!static_cast<dr::WorkQueue*>(q)->Take(&s)) return -1;
  snprintf(out, cap, "%s", s.c_str()); return (int)s.size();
}
void dr_wq_add(void* q, const char* item) { static_cast<dr::WorkQueue*>(q)->Add(item); }
void dr_wq_state(void* q, int64_t* epoch, int64_t* pos, int64_t* taken) { static_cast<dr::WorkQueue*>(q)->State(epoch, pos, taken); }
void dr_wq_restore(void* q, int64_t epoch, int64_t pos) { static_cast<dr::WorkQueue*>(q)->Restore(epoch, pos); }
int64_t dr_wq_remaining(void* q) { return static_cast<dr::WorkQueue*>(q)->Remaining(); }

// ---- synthetic data ----------------------------------------------------------------------
// Criteo-shaped: dense [batch, num_dense] fp32 (log-normal-ish, >=0), ids feature-major
// [num_tables, batch] int64, labels [batch] fp32.  Deterministic in (seed, batch index).
void dr_gen_criteo(uint64_t seed, int64_t batch, int num_dense, int num_tables, const int64_t* cards, double alpha,
                   float* dense, int64_t* ids, float* labels, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  auto work = [&](int t) {
    int64_t b0 = batch * t / nthreads, b1 = batch * (t + 1) / nthreads;
    dr::XorShift r(seed * 1315423911ULL + (uint64_t)t * 2654435761ULL + 17);
    for (int64_t b = b0; b < b1; ++b) {
      for (int d = 0; d < num_dense; ++d) { double u = r.uni(); dense[b * num_dense + d] = (float)std::log1p(u * u * 100.0); }
      labels[b] = r.uni() < 0.25 ? 1.0f : 0.0f;
    }
    for (int f = 0; f < num_tables; ++f)
      for (int64_t b = b0; b < b1; ++b) ids[(int64_t)f * batch + b] = dr::PowerLawId(r, cards[f], alpha);
  };
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

// Taobao-shaped (DIN/DIEN/BST): user id, target item, target category, history of
// (item, category) pairs with variable length <= max_len (padded with -1), labels.
void dr_gen_taobao(uint64_t seed, int64_t batch, int max_len, int64_t n_users, int64_t n_items, int64_t n_cats, double alpha,
                   int64_t* user, int64_t* item, int64_t* cat, int64_t* hist_item, int64_t* hist_cat, int32_t* hist_len, float* labels) {
  dr::XorShift r(seed * 7919ULL + 3);
  for (int64_t b = 0; b < batch; ++b) {
    user[b] = dr::PowerLawId(r, n_users, alpha);
    item[b] = dr::PowerLawId(r, n_items, alpha);
    cat[b] = item[b] % n_cats;
    int len = 1 + (int)(r.uni() * max_len); if (len > max_len) len = max_len;
    hist_len[b] = len;
    for (int t = 0; t < max_len; ++t) {
      if (t < len) { int64_t it = dr::PowerLawId(r, n_items, alpha); hist_item[b * max_len + t] = it; hist_cat[b * max_len + t] = it % n_cats; }
      else { hist_item[b * max_len + t] = -1; hist_cat[b * max_len + t] = -1; }
    }
    labels[b] = r.uni() < 0.5 ? 1.0f : 0.0f;
  }
}

}  // extern "C"
