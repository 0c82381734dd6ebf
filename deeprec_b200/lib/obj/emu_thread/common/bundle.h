// Checkpoint tensor bundle: <prefix>.data (raw, 64-byte aligned records) + <prefix>.index (text).
// Header-only so the host engine (training checkpoints) and the serving runtime (model load / delta update)
// share one implementation.  Streaming 8 MiB writer, CRC32 per tensor, atomic publish (data first, index last).
#pragma once
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

namespace dr {

// CRC-32 (IEEE 802.3, reflected 0xEDB88320), slicing-by-8: eight table lookups per 8 input bytes instead of one per byte -- a multi-GB
// EmbeddingVariable dump is checksummed at memory-copy-like speed instead of ~0.35 GB/s.  Tables are built once (thread-safe static).
struct Crc32Tables {
  uint32_t t[8][256];
  Crc32Tables() {
    for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; t[0][i] = c; }
    for (uint32_t i = 0; i < 256; ++i) for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
inline uint32_t Crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
  static const Crc32Tables tb;
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { crc = tb.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8); --n; }
  while (n >= 8) {
    uint64_t v; memcpy(&v, p, 8);                       // little-endian host (x86-64 / aarch64)
    const uint32_t lo = crc ^ (uint32_t)v, hi = (uint32_t)(v >> 32);
    crc = tb.t[7][lo & 0xFF] ^ tb.t[6][(lo >> 8) & 0xFF] ^ tb.t[5][(lo >> 16) & 0xFF] ^ tb.t[4][lo >> 24] ^
          tb.t[3][hi & 0xFF] ^ tb.t[2][(hi >> 8) & 0xFF] ^ tb.t[1][(hi >> 16) & 0xFF] ^ tb.t[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = tb.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

// crc(A || B) from crc(A), crc(B) and |B|: multiply crc(A) by x^(8|B|) in GF(2)[x] / P(x) by repeated squaring of the "shift one zero
// bit" operator (the classic combine construction) -- lets large tensors be checksummed in parallel chunks.
inline uint32_t Crc32Combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  auto times = [](const uint32_t* mat, uint32_t vec) { uint32_t sum = 0; for (; vec; vec >>= 1, ++mat) if (vec & 1) sum ^= *mat; return sum; };
  auto square = [&](uint32_t* sq, const uint32_t* mat) { for (int n = 0; n < 32; ++n) sq[n] = times(mat, mat[n]); };
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;                                   // operator for one zero bit
  for (int n = 1; n < 32; ++n) odd[n] = 1u << (n - 1);
  square(even, odd);                                      // two zero bits
  square(odd, even);                                      // four zero bits
  do {                                                    // first square gives one zero byte, then 2, 4, ... bytes
    square(even, odd);
    if (len2 & 1) crc1 = times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    square(odd, even);
    if (len2 & 1) crc1 = times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}

// Whole-tensor checksum: chunks in parallel on the OpenMP runtime when the translation unit is built with it (the host library).
inline uint32_t Crc32Large(const uint8_t* p, size_t n) {
#ifdef _OPENMP
  constexpr size_t kMinChunk = size_t(4) << 20;
  int parts = (int)std::min<size_t>((size_t)omp_get_max_threads(), n / kMinChunk);
  if (parts > 1 && !omp_in_parallel()) {
    std::vector<uint32_t> crc((size_t)parts);
    const size_t per = ((n + parts - 1) / parts + 7) & ~size_t(7);
#pragma omp parallel for schedule(static, 1) num_threads(parts)
    for (int i = 0; i < parts; ++i) {
      const size_t b = std::min(n, (size_t)i * per), e = std::min(n, b + per);
      crc[(size_t)i] = Crc32(p + b, e - b);
    }
    uint32_t c = crc[0];
    for (int i = 1; i < parts; ++i) { const size_t b = std::min(n, (size_t)i * per), e = std::min(n, b + per); c = Crc32Combine(c, crc[(size_t)i], e - b); }
    return c;
  }
#endif
  return Crc32(p, n);
}

// ---------------------------------------------------------------------------------------
// Tensor bundle: <prefix>.data (raw, 64-byte aligned records) + <prefix>.index (text).
// ---------------------------------------------------------------------------------------
struct BundleEntry { std::string name, dtype; std::vector<int64_t> shape; int64_t offset = 0, nbytes = 0; uint32_t crc = 0; };

class BundleWriter {
 public:
  explicit BundleWriter(const std::string& prefix) : prefix_(prefix) {
    f_ = fopen((prefix + ".data.tmp").c_str(), "wb");
    buf_.resize(8 << 20);
    if (f_) setvbuf(f_, buf_.data(), _IOFBF, buf_.size());
  }
  bool ok() const { return f_ != nullptr; }
  int Add(const char* name, const char* dtype, const int64_t* shape, int ndim, const void* data, int64_t nbytes) {
    if (!f_) return -1;
    int64_t pad = (64 - (off_ & 63)) & 63;
    static const char zeros[64] = {0};
    if (pad) { if (fwrite(zeros, 1, (size_t)pad, f_) != (size_t)pad) { failed_ = true; return -2; } off_ += pad; }
    BundleEntry e; e.name = name; e.dtype = dtype; e.shape.assign(shape, shape + ndim); e.offset = off_; e.nbytes = nbytes;
    e.crc = Crc32Large(static_cast<const uint8_t*>(data), (size_t)nbytes);
    if (nbytes && fwrite(data, 1, (size_t)nbytes, f_) != (size_t)nbytes) { failed_ = true; return -2; }
    off_ += nbytes;
    entries_.push_back(std::move(e));
    return 0;
  }
  int Close() {
    if (!f_) return -1;
    // a short write (ENOSPC, quota) surfaces in fwrite, fflush or fclose: NEVER publish the index of a truncated data file --
    // the caller prunes older checkpoints only after a successful close
    const bool flush_bad = fflush(f_) != 0 || ferror(f_) != 0;
    const bool close_bad = fclose(f_) != 0;
    f_ = nullptr;
    if (failed_ || flush_bad || close_bad) { remove((prefix_ + ".data.tmp").c_str()); return -5; }
    FILE* fi = fopen((prefix_ + ".index.tmp").c_str(), "w");
    if (!fi) return -2;
    fprintf(fi, "DEEPREC_B200_BUNDLE 1 %zu\n", entries_.size());
    for (auto& e : entries_) {
      fprintf(fi, "%s\t%s\t%zu", e.name.c_str(), e.dtype.c_str(), e.shape.size());
      for (auto d : e.shape) fprintf(fi, "\t%lld", (long long)d);
      fprintf(fi, "\t%lld\t%lld\t%u\n", (long long)e.offset, (long long)e.nbytes, e.crc);
    }
    const bool index_bad = ferror(fi) != 0;
    if (fclose(fi) != 0 || index_bad) { remove((prefix_ + ".index.tmp").c_str()); remove((prefix_ + ".data.tmp").c_str()); return -6; }
    // atomic publish: data first, index last (a reader that sees the index sees complete data)
    if (rename((prefix_ + ".data.tmp").c_str(), (prefix_ + ".data").c_str()) != 0) return -3;
    if (rename((prefix_ + ".index.tmp").c_str(), (prefix_ + ".index").c_str()) != 0) return -4;
    return 0;
  }
  ~BundleWriter() { if (f_) fclose(f_); }
 private:
  std::string prefix_; FILE* f_ = nullptr; std::vector<char> buf_; int64_t off_ = 0; std::vector<BundleEntry> entries_; bool failed_ = false;
};

class BundleReader {
 public:
  explicit BundleReader(const std::string& prefix) : prefix_(prefix) {
    FILE* fi = fopen((prefix + ".index").c_str(), "r");
    if (!fi) return;
    char magic[64]; int ver; size_t n;
    if (fscanf(fi, "%63s %d %zu\n", magic, &ver, &n) != 3 || std::string(magic) != "DEEPREC_B200_BUNDLE") { fclose(fi); return; }
    std::vector<char> line(1 << 16);
    for (size_t i = 0; i < n; ++i) {
      if (!fgets(line.data(), (int)line.size(), fi)) break;
      std::vector<std::string> tok; char* save = nullptr;
      for (char* t = strtok_r(line.data(), "\t\n", &save); t; t = strtok_r(nullptr, "\t\n", &save)) tok.emplace_back(t);
      if (tok.size() < 6) continue;
      BundleEntry e; e.name = tok[0]; e.dtype = tok[1]; size_t nd = std::stoul(tok[2]);
      if (tok.size() != 3 + nd + 3) continue;
      for (size_t d = 0; d < nd; ++d) e.shape.push_back(std::stoll(tok[3 + d]));
      e.offset = std::stoll(tok[3 + nd]); e.nbytes = std::stoll(tok[4 + nd]); e.crc = (uint32_t)std::stoul(tok[5 + nd]);
      index_[e.name] = entries_.size(); entries_.push_back(std::move(e));
    }
    fclose(fi);
    f_ = fopen((prefix + ".data").c_str(), "rb");
  }
  ~BundleReader() { if (f_) fclose(f_); }
  bool ok() const { return f_ != nullptr; }
  const std::vector<BundleEntry>& entries() const { return entries_; }
  const BundleEntry* Find(const std::string& n) const { auto it = index_.find(n); return it == index_.end() ? nullptr : &entries_[it->second]; }
  int Read(const BundleEntry& e, void* dst, int verify) {
    std::lock_guard<std::mutex> l(mu_);
    if (fseeko(f_, e.offset, SEEK_SET) != 0) return -1;
    if (e.nbytes && fread(dst, 1, (size_t)e.nbytes, f_) != (size_t)e.nbytes) return -2;
    if (verify && Crc32Large(static_cast<const uint8_t*>(dst), (size_t)e.nbytes) != e.crc) return -3;
    return 0;
  }
 private:
  std::string prefix_; FILE* f_ = nullptr; std::vector<BundleEntry> entries_; std::map<std::string, size_t> index_; std::mutex mu_;
};



// fp32 view of a tensor that a low-precision conversion may have stored as bf16 / f16 / int8 (tools/low_precision_optimize.py: int8 tensors carry a
// sibling `<name>/scale`, one fp32 per leading-dimension row or one for the whole tensor).  The serving runtimes read every float tensor through
// this, so a converted saved model (2-4x smaller to ship, deltas included) loads like the original.
inline bool ReadAsFloat(BundleReader& r, const std::string& name, std::vector<float>* out, std::vector<int64_t>* shape = nullptr) {
  const BundleEntry* e = r.Find(name);
  if (!e) return false;
  if (shape) *shape = e->shape;
  if (e->dtype != "bf16" && e->dtype != "f16" && e->dtype != "i8") {           // fp32 under any spelling ("f32", writers of other tools: "float32"): raw bytes
    out->resize((size_t)e->nbytes / 4); return r.Read(*e, out->data(), 1) == 0;
  }
  std::vector<uint8_t> raw((size_t)e->nbytes);
  if (r.Read(*e, raw.data(), 1) != 0) return false;
  if (e->dtype == "bf16") {
    const size_t n = raw.size() / 2; out->resize(n);
    for (size_t i = 0; i < n; ++i) { uint16_t h; memcpy(&h, raw.data() + 2 * i, 2); const uint32_t u = (uint32_t)h << 16; memcpy(&(*out)[i], &u, 4); }
    return true;
  }
  if (e->dtype == "f16") {
    const size_t n = raw.size() / 2; out->resize(n);
    for (size_t i = 0; i < n; ++i) {
      uint16_t h; memcpy(&h, raw.data() + 2 * i, 2);
      const uint32_t sign = (uint32_t)(h & 0x8000) << 16; uint32_t ex = (h >> 10) & 0x1F, man = h & 0x3FF, u;
      if (ex == 0) {
        if (man == 0) u = sign;
        else { int sh = 0; while (!(man & 0x400)) { man <<= 1; ++sh; } man &= 0x3FF; u = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (man << 13); }
      } else if (ex == 31) u = sign | 0x7F800000u | (man << 13);
      else u = sign | ((ex - 15 + 127) << 23) | (man << 13);
      memcpy(&(*out)[i], &u, 4);
    }
    return true;
  }
  if (e->dtype == "i8") {
    std::vector<float> sc;
    const BundleEntry* se = r.Find(name + "/scale");
    if (!se || se->dtype != "f32") return false;
    sc.resize((size_t)se->nbytes / 4);
    if (r.Read(*se, sc.data(), 1) != 0 || sc.empty()) return false;
    const size_t n = raw.size(); out->resize(n);
    const size_t rows = sc.size(), per = rows ? n / rows : n;
    if (rows > 1 && per * rows != n) return false;
    for (size_t i = 0; i < n; ++i) (*out)[i] = (float)(int8_t)raw[i] * sc[rows > 1 ? i / per : 0];
    return true;
  }
  return false;
}

}  // namespace dr
