// SsdHashStore: log-structured key -> row store on a file system (the SSDHASH tier).
//
// Reference behaviour (framework/embedding/ssd_hash_kv.h:139-810, emb_file.h, ssd_record_descriptor.h): an in-memory
// key -> {file, offset} index, append-only ".emb" files, a write buffer, and a compaction pass (synchronous or on a
// background thread, TF_SSDHASH_ASYNC_COMPACTION) that rewrites the live records of mostly-dead files.
//
// This implementation: fixed-size records [key | freq | version | row(stride floats)], one sharded index
// (mutex per shard), append through a per-store buffer with pwrite, reads with pread (page cache does the caching a
// DRAM tier above us does not), per-file live/total counters so compaction picks files by garbage ratio, and an optional
// compaction thread woken whenever a file crosses the ratio.  Checkpointing walks the index (ExportAll), so no file
// hard-copy protocol is needed.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace dr {

struct SsdPos { int32_t file; int64_t off; };

class SsdHashStore {
 public:
  SsdHashStore(const std::string& dir, int64_t stride, int64_t file_bytes, int async_compaction)
      : dir_(dir), stride_(stride), rec_bytes_(24 + stride * 4), file_bytes_(std::max<int64_t>(file_bytes, rec_bytes_ * 64)),
        async_(async_compaction != 0) {
    ::mkdir(dir_.c_str(), 0755);
    OpenNewFile();
    if (async_) worker_ = std::thread([this] { CompactionLoop(); });
  }
  ~SsdHashStore() {
    {
      std::lock_guard<std::mutex> l(cmu_);
      stop_ = true;
    }
    ccv_.notify_all();
    if (worker_.joinable()) worker_.join();
    std::lock_guard<std::mutex> l(wmu_);
    for (auto& f : files_)
      if (f.fd >= 0) { ::close(f.fd); ::unlink(f.path.c_str()); }
    ::rmdir(dir_.c_str());
  }

  int64_t Size() const { return size_.load(); }
  int64_t NumFiles() {
    std::lock_guard<std::mutex> l(wmu_);
    int64_t n = 0;
    for (auto& f : files_) n += f.fd >= 0;
    return n;
  }
  int64_t BytesOnDisk() {
    std::lock_guard<std::mutex> l(wmu_);
    int64_t n = 0;
    for (auto& f : files_) if (f.fd >= 0) n += f.total * rec_bytes_;
    return n;
  }
  int64_t Compactions() const { return compactions_.load(); }

  // append (or overwrite) n records
  void Put(const int64_t* keys, const float* rows, const int64_t* freqs, const int64_t* versions, int64_t n) {
    std::vector<char> rec(rec_bytes_);
    for (int64_t i = 0; i < n; ++i) {
      memcpy(rec.data(), &keys[i], 8);
      int64_t f = freqs ? freqs[i] : 0, v = versions ? versions[i] : -1;
      memcpy(rec.data() + 8, &f, 8); memcpy(rec.data() + 16, &v, 8);
      memcpy(rec.data() + 24, rows + i * stride_, stride_ * 4);
      SsdPos p = Append(rec.data());
      Shard& S = shard(keys[i]);
      std::lock_guard<std::mutex> l(S.mu);
      auto it = S.map.find(keys[i]);
      if (it != S.map.end()) { MarkDead(it->second.file); it->second = p; }
      else { S.map.emplace(keys[i], p); size_.fetch_add(1); }
    }
    MaybeCompact();
  }

  // read n records; found[i] = 0 leaves row i untouched
  void Get(const int64_t* keys, int64_t n, float* rows, int64_t* freqs, int64_t* versions, uint8_t* found) {
    Flush();
    std::vector<char> rec(rec_bytes_);
    for (int64_t i = 0; i < n; ++i) {
      SsdPos p;
      {
        Shard& S = shard(keys[i]);
        std::lock_guard<std::mutex> l(S.mu);
        auto it = S.map.find(keys[i]);
        if (it == S.map.end()) { found[i] = 0; continue; }
        p = it->second;
      }
      // the record may move under compaction (or still sit in the write buffer) between index read and pread:
      // verify the key, flush, re-read the index and retry
      bool ok = false;
      for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
        {
          std::shared_lock<std::shared_mutex> rd(fd_mu_);      // compaction may not close this descriptor while it is being read
          int fd = FdOf(p.file);
          ok = fd >= 0 && ::pread(fd, rec.data(), rec_bytes_, p.off) == (ssize_t)rec_bytes_ && memcmp(rec.data(), &keys[i], 8) == 0;
        }
        if (ok) break;
        Flush();
        Shard& S = shard(keys[i]);
        std::lock_guard<std::mutex> l(S.mu);
        auto it = S.map.find(keys[i]);
        if (it == S.map.end()) break;
        p = it->second;
      }
      if (!ok) { found[i] = 0; continue; }
      found[i] = 1;
      if (freqs) memcpy(&freqs[i], rec.data() + 8, 8);
      if (versions) memcpy(&versions[i], rec.data() + 16, 8);
      if (rows) memcpy(rows + i * stride_, rec.data() + 24, stride_ * 4);
    }
  }

  void Contains(const int64_t* keys, int64_t n, uint8_t* found) {
    for (int64_t i = 0; i < n; ++i) {
      Shard& S = shard(keys[i]);
      std::lock_guard<std::mutex> l(S.mu);
      found[i] = S.map.count(keys[i]) ? 1 : 0;
    }
  }

  int64_t Remove(const int64_t* keys, int64_t n) {
    int64_t removed = 0;
    for (int64_t i = 0; i < n; ++i) {
      Shard& S = shard(keys[i]);
      std::lock_guard<std::mutex> l(S.mu);
      auto it = S.map.find(keys[i]);
      if (it == S.map.end()) continue;
      MarkDead(it->second.file);
      S.map.erase(it);
      size_.fetch_sub(1); ++removed;
    }
    MaybeCompact();
    return removed;
  }

  // all live keys (for checkpoint); returns count, fills up to cap
  int64_t ExportKeys(int64_t* out, int64_t cap) {
    int64_t n = 0;
    for (auto& S : shards_) {
      std::lock_guard<std::mutex> l(S.mu);
      for (auto& kv : S.map) { if (n < cap) out[n] = kv.first; ++n; }
    }
    return n;
  }

  // rewrite the live records of every file whose dead fraction >= ratio (never the active file); returns files reclaimed
  int64_t Compact(double ratio) {
    std::lock_guard<std::mutex> g(compact_mu_);
    Flush();
    std::vector<int32_t> victims;
    {
      std::lock_guard<std::mutex> l(wmu_);
      for (int32_t f = 0; f < (int32_t)files_.size(); ++f) {
        auto& F = files_[f];
        if (F.fd < 0 || f == active_ || F.total == 0) continue;
        if ((double)(F.total - F.live.load()) / (double)F.total >= ratio) victims.push_back(f);
      }
    }
    std::vector<char> rec(rec_bytes_);
    for (int32_t f : victims) {
      int fd = FdOf(f);
      int64_t total;
      { std::lock_guard<std::mutex> l(wmu_); total = files_[f].total; }
      for (int64_t r = 0; r < total; ++r) {
        const int64_t off = r * rec_bytes_;
        if (::pread(fd, rec.data(), rec_bytes_, off) != (ssize_t)rec_bytes_) break;
        int64_t key; memcpy(&key, rec.data(), 8);
        Shard& S = shard(key);
        std::lock_guard<std::mutex> l(S.mu);
        auto it = S.map.find(key);
        if (it == S.map.end() || it->second.file != f || it->second.off != off) continue;    // dead record
        it->second = Append(rec.data());
      }
      Flush();
      std::unique_lock<std::shared_mutex> wr(fd_mu_);          // wait for in-flight readers of this file (found by ThreadSanitizer: close vs pread)
      std::lock_guard<std::mutex> l(wmu_);
      ::close(files_[f].fd); ::unlink(files_[f].path.c_str());
      files_[f].fd = -1; files_[f].total = 0; files_[f].live = 0;
    }
    compactions_.fetch_add((int64_t)victims.size());
    return (int64_t)victims.size();
  }

  void Flush() {
    std::lock_guard<std::mutex> l(wmu_);
    FlushLocked();
  }

 private:
  struct File {
    int fd = -1; std::string path; int64_t total = 0; std::atomic<int64_t> live{0};
    File() = default;
    File(File&& o) noexcept : fd(o.fd), path(std::move(o.path)), total(o.total), live(o.live.load()) {}
  };
  struct Shard { std::mutex mu; std::unordered_map<int64_t, SsdPos> map; };
  static constexpr int kShards = 64;

  Shard& shard(int64_t key) { return shards_[((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 58]; }

  void OpenNewFile() {      // wmu_ held (or constructor)
    File F;
    F.path = dir_ + "/" + std::to_string(files_.size()) + ".emb";
    F.fd = ::open(F.path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    files_.push_back(std::move(F));
    active_ = (int32_t)files_.size() - 1;
    buf_.clear(); buf_off_ = 0;
  }
  void FlushLocked() {
    if (buf_.empty()) return;
    ssize_t w = ::pwrite(files_[active_].fd, buf_.data(), buf_.size(), buf_off_);
    (void)w;
    buf_off_ += (int64_t)buf_.size();
    buf_.clear();
  }
  SsdPos Append(const char* rec) {
    std::lock_guard<std::mutex> l(wmu_);
    if ((files_[active_].total + 1) * rec_bytes_ > file_bytes_) { FlushLocked(); OpenNewFile(); }
    File& F = files_[active_];
    SsdPos p{active_, F.total * rec_bytes_};
    buf_.insert(buf_.end(), rec, rec + rec_bytes_);
    F.total += 1; F.live.fetch_add(1);
    if ((int64_t)buf_.size() >= (4 << 20)) FlushLocked();
    return p;
  }
  void MarkDead(int32_t f) {
    std::lock_guard<std::mutex> l(wmu_);
    files_[f].live.fetch_sub(1);
  }
  int FdOf(int32_t f) {
    std::lock_guard<std::mutex> l(wmu_);
    return f >= 0 && f < (int32_t)files_.size() ? files_[f].fd : -1;
  }
  void MaybeCompact() {
    if (!async_) return;
    { std::lock_guard<std::mutex> l(cmu_); wake_ = true; }
    ccv_.notify_one();
  }
  void CompactionLoop() {
    std::unique_lock<std::mutex> l(cmu_);
    while (!stop_) {
      ccv_.wait(l, [this] { return stop_ || wake_; });
      if (stop_) break;
      wake_ = false;
      l.unlock();
      Compact(0.5);
      l.lock();
    }
  }

  std::string dir_;
  int64_t stride_, rec_bytes_, file_bytes_;
  bool async_;
  Shard shards_[kShards];
  std::mutex wmu_;                       // files_, active_, write buffer
  std::vector<File> files_;
  int32_t active_ = 0;
  std::vector<char> buf_;
  int64_t buf_off_ = 0;
  std::atomic<int64_t> size_{0}, compactions_{0};
  std::mutex compact_mu_;
  std::shared_mutex fd_mu_;            // readers (Get) shared, descriptor close (Compact) exclusive
  std::mutex cmu_;
  std::condition_variable ccv_;
  bool stop_ = false, wake_ = false;
  std::thread worker_;
};

}  // namespace dr

extern "C" {
void* dr_ssd_create(const char* dir, int64_t stride, int64_t file_bytes, int async_compaction) {
  return new dr::SsdHashStore(dir, stride, file_bytes, async_compaction);
}
void dr_ssd_destroy(void* h) { delete static_cast<dr::SsdHashStore*>(h); }
int64_t dr_ssd_size(void* h) { return static_cast<dr::SsdHashStore*>(h)->Size(); }
int64_t dr_ssd_num_files(void* h) { return static_cast<dr::SsdHashStore*>(h)->NumFiles(); }
int64_t dr_ssd_bytes(void* h) { return static_cast<dr::SsdHashStore*>(h)->BytesOnDisk(); }
int64_t dr_ssd_compactions(void* h) { return static_cast<dr::SsdHashStore*>(h)->Compactions(); }
void dr_ssd_put(void* h, const int64_t* keys, const float* rows, const int64_t* freqs, const int64_t* versions, int64_t n) {
  static_cast<dr::SsdHashStore*>(h)->Put(keys, rows, freqs, versions, n);
}
void dr_ssd_get(void* h, const int64_t* keys, int64_t n, float* rows, int64_t* freqs, int64_t* versions, uint8_t* found) {
  static_cast<dr::SsdHashStore*>(h)->Get(keys, n, rows, freqs, versions, found);
}
void dr_ssd_contains(void* h, const int64_t* keys, int64_t n, uint8_t* found) { static_cast<dr::SsdHashStore*>(h)->Contains(keys, n, found); }
int64_t dr_ssd_remove(void* h, const int64_t* keys, int64_t n) { return static_cast<dr::SsdHashStore*>(h)->Remove(keys, n); }
int64_t dr_ssd_export_keys(void* h, int64_t* out, int64_t cap) { return static_cast<dr::SsdHashStore*>(h)->ExportKeys(out, cap); }
int64_t dr_ssd_compact(void* h, double ratio) { return static_cast<dr::SsdHashStore*>(h)->Compact(ratio); }
void dr_ssd_flush(void* h) { static_cast<dr::SsdHashStore*>(h)->Flush(); }
}
