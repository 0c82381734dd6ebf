// Native concurrency stress test of the host runtime through its C ABI, meant to be built with -fsanitize=thread (and =address).
// Mirrors the reference's gtest tier on the storage engine (core/kernels/embedding_variable_ops_test.cc: TestLookupRemoveConcurrency,
// TestFeatureFilterParallel, lockless insert/remove/export, SSDHashKV sync/async compaction) -- SURVEY §4 / §5.2.
//
// What may legitimately run concurrently in the framework (and therefore here):
//   * Lookup (read-only forward) from many threads while ONE training thread applies updates (Apply parallelises internally).
//     Structural safety (index growth, chunk allocation, metadata publication) is what is checked here; reading the VALUES of a
//     row that is being updated at the same moment is Hogwild by design (async embedding stage / async PS, exactly as in the
//     reference), so the readers of this test look up keys the concurrent updates do not touch;
//   * Remove / Shrink / Snapshot / Import are step-boundary operations: concurrent with Lookup of OTHER keys, never with Apply;
//   * SsdHashStore Put/Get/Remove/Compact from any thread; StagingQueue / WorkQueue are MPMC.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../deeprec_b200/csrc/common/ev_types.h"

extern "C" {
void* dr_host_ev_create(const DrEvConfig* cfg);
void dr_host_ev_destroy(void* h);
int64_t dr_host_ev_stride(void* h);
void dr_host_ev_set_default(void* h, const float* m);
int64_t dr_host_ev_size(void* h);
void dr_host_ev_lookup(void* h, const int64_t* keys, int64_t n, float* out);
void dr_host_ev_get_freq(void* h, const int64_t* keys, int64_t n, int64_t* out);
void dr_host_ev_apply(void* h, const int64_t* keys, const float* grads, const int64_t* counts, int64_t n, const DrOptHyper* hp);
int64_t dr_host_ev_shrink(void* h, int64_t step);
int64_t dr_host_ev_remove(void* h, const int64_t* keys, int64_t n);
void dr_host_ev_snapshot_begin(void* h, int dirty_only, int part_id, int part_num, int64_t* na, int64_t* nf);
void dr_host_ev_snapshot_read(void* h, int64_t* keys, float* rows, int64_t* freqs, int64_t* versions, int64_t* poff, int64_t* fkeys,
                              int64_t* ffreqs, int64_t* fversions, int64_t* fpoff);
void dr_host_ev_snapshot_end(void* h);
int64_t dr_host_ev_import(void* h, const int64_t* keys, const float* rows, int64_t ncols, const int64_t* freqs, const int64_t* versions,
                          int64_t n, int part_id, int part_num, int reset_version);
void dr_host_ev_clear_dirty(void* h);
int64_t dr_host_ev_total_keys(void* h);
void* dr_ssd_create(const char* dir, int64_t stride, int64_t file_bytes, int async_compaction);
void dr_ssd_destroy(void* h);
int64_t dr_ssd_size(void* h);
void dr_ssd_put(void* h, const int64_t* keys, const float* rows, const int64_t* freqs, const int64_t* versions, int64_t n);
void dr_ssd_get(void* h, const int64_t* keys, int64_t n, float* rows, int64_t* freqs, int64_t* versions, uint8_t* found);
int64_t dr_ssd_remove(void* h, const int64_t* keys, int64_t n);
int64_t dr_ssd_compact(void* h, double ratio);
void* dr_stage_create(int64_t capacity);
void dr_stage_destroy(void* q);
int dr_stage_put(void* q, int64_t ticket, int64_t timeout_ms);
int dr_stage_take(void* q, int64_t* ticket, int64_t timeout_ms);
void dr_stage_close(void* q);
}

static int g_fail = 0;
#define CHECK(cond)                                                             \
  do {                                                                          \
    if (!(cond)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

static DrEvConfig MakeCfg(int dim, int filter_freq) {
  DrEvConfig c;
  memset(&c, 0, sizeof(c));
  c.dim = dim; c.num_slots = 1; c.init_capacity = 1024;
  c.filter_type = filter_freq > 0 ? 1 : 0; c.filter_freq = filter_freq; c.bloom_counter_bits = 32;
  c.l2_weight_threshold = -1.f; c.default_value_dim = 16; c.record_freq = c.record_version = 1; c.num_partitions = 8;
  c.slot_init[0] = 0.1f;
  return c;
}
static DrOptHyper Adagrad(int64_t step) {
  DrOptHyper hp;
  memset(&hp, 0, sizeof(hp));
  hp.kind = 1; hp.lr = 0.1f; hp.init_accum = 0.1f; hp.global_step = step; hp.epsilon = 1e-8f; hp.beta1 = 0.9f; hp.beta2 = 0.999f; hp.lr_power = -0.5f;
  return hp;
}

// 1. readers hammer Lookup over a growing key space while the trainer inserts + updates (table growth under concurrent reads)
static void TestLookupWhileTraining() {
  const int dim = 8;
  DrEvConfig cfg = MakeCfg(dim, 0);
  void* ev = dr_host_ev_create(&cfg);
  std::vector<float> def(16 * dim, 0.5f);
  dr_host_ev_set_default(ev, def.data());
  std::vector<int64_t> keys(4096), counts(4096, 1);
  std::vector<float> grads(4096 * dim, 0.01f);
  constexpr int64_t kStable = 8192, kNewBase = 1000000;
  for (int c = 0; c < 2; ++c) {      // pre-train the stable range the readers use
    for (int i = 0; i < 4096; ++i) keys[i] = c * 4096 + i;
    DrOptHyper hp = Adagrad(0);
    dr_host_ev_apply(ev, keys.data(), grads.data(), counts.data(), 4096, &hp);
  }
  std::vector<float> ref(kStable * dim);
  {
    std::vector<int64_t> all(kStable);
    for (int64_t i = 0; i < kStable; ++i) all[i] = i;
    dr_host_ev_lookup(ev, all.data(), kStable, ref.data());
  }
  std::atomic<bool> stop{false};
  std::vector<std::thread> readers;
  for (int t = 0; t < 4; ++t)
    readers.emplace_back([&, t] {
      std::vector<int64_t> k(512), fr(512);
      std::vector<float> out(512 * dim);
      uint64_t x = 88172645463325252ull + t;
      while (!stop.load()) {
        for (auto& kk : k) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; kk = (int64_t)(x % kStable); }
        dr_host_ev_lookup(ev, k.data(), (int64_t)k.size(), out.data());
        for (int i = 0; i < 512; ++i) if (out[i * dim] != ref[k[i] * dim]) { CHECK(false && "stable row changed or torn"); break; }
        // metadata of keys that are being inserted right now: any value is fine, the read must be race-free
        for (auto& kk : k) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; kk = kNewBase + (int64_t)(x % (40 * 4096)); }
        dr_host_ev_get_freq(ev, k.data(), (int64_t)k.size(), fr.data());
      }
    });
  for (int step = 0; step < 40; ++step) {
    for (int i = 0; i < 4096; ++i) keys[i] = kNewBase + (int64_t)step * 4096 + i;     // unique within the call (Apply's contract), all new
    DrOptHyper hp = Adagrad(step + 1);
    dr_host_ev_apply(ev, keys.data(), grads.data(), counts.data(), 4096, &hp);
  }
  stop = true;
  for (auto& r : readers) r.join();
  CHECK(dr_host_ev_size(ev) == kStable + 40 * 4096);
  dr_host_ev_destroy(ev);
}

// 2. counter-filter admission from many threads' worth of duplicates + remove / snapshot at step boundaries with readers running
static void TestFilterRemoveSnapshot() {
  const int dim = 4;
  DrEvConfig cfg = MakeCfg(dim, 3);
  void* ev = dr_host_ev_create(&cfg);
  std::vector<float> def(16 * dim, 0.25f);
  dr_host_ev_set_default(ev, def.data());
  std::vector<int64_t> keys(2000), counts(2000, 1);
  std::vector<float> grads(2000 * dim, 0.1f);
  for (int i = 0; i < 2000; ++i) keys[i] = i;
  for (int step = 0; step < 2; ++step) { DrOptHyper hp = Adagrad(step); dr_host_ev_apply(ev, keys.data(), grads.data(), counts.data(), 2000, &hp); }
  CHECK(dr_host_ev_size(ev) == 0);                       // freq 2 < 3: nothing admitted yet
  { DrOptHyper hp = Adagrad(2); dr_host_ev_apply(ev, keys.data(), grads.data(), counts.data(), 2000, &hp); }
  CHECK(dr_host_ev_size(ev) == 2000);
  std::atomic<bool> stop{false};
  std::thread reader([&] {
    std::vector<int64_t> k(256);
    std::vector<float> out(256 * dim);
    while (!stop.load()) {
      for (int i = 0; i < 256; ++i) k[i] = 1000 + i;     // keys that are NOT being removed
      dr_host_ev_lookup(ev, k.data(), 256, out.data());
    }
  });
  std::vector<int64_t> rm(500);
  for (int i = 0; i < 500; ++i) rm[i] = i;
  CHECK(dr_host_ev_remove(ev, rm.data(), 500) == 500);
  int64_t na = 0, nf = 0;
  dr_host_ev_snapshot_begin(ev, 0, 0, 1, &na, &nf);
  CHECK(na == 1500);
  std::vector<int64_t> sk(na), sf(na), sv(na), poff(1001), fk(nf + 1), ff(nf + 1), fv(nf + 1), fpoff(1001);
  std::vector<float> rows((size_t)na * dr_host_ev_stride(ev));
  dr_host_ev_snapshot_read(ev, sk.data(), rows.data(), sf.data(), sv.data(), poff.data(), fk.data(), ff.data(), fv.data(), fpoff.data());
  dr_host_ev_snapshot_end(ev);
  for (int64_t i = 0; i < na; ++i) CHECK(sk[i] >= 500 && sf[i] == 3);
  stop = true;
  reader.join();
  dr_host_ev_destroy(ev);
}

// 2b. checkpoint-scale passes on the worker pool: snapshot of a 40k-key table (every partition scanned by a different worker, parallel
// bucket sort), parallel import into a second table -- with every key present twice in the batch and readers probing the table while it
// grows -- then a parallel eviction pass and a second import that recycles the freed rows.
static void TestParallelSnapshotImportShrink() {
  const int dim = 4;
  const int64_t n = 40000;
  DrEvConfig cfg = MakeCfg(dim, 0);
  cfg.steps_to_live = 5;
  void* src = dr_host_ev_create(&cfg);
  std::vector<float> def(16 * dim, 0.5f);
  dr_host_ev_set_default(src, def.data());
  std::vector<int64_t> keys(n), counts(n, 1);
  std::vector<float> grads((size_t)n * dim, 0.01f);
  for (int64_t i = 0; i < n; ++i) keys[i] = i * 7919 + 13;
  { DrOptHyper hp = Adagrad(1); dr_host_ev_apply(src, keys.data(), grads.data(), counts.data(), n, &hp); }
  int64_t na = 0, nf = 0;
  dr_host_ev_snapshot_begin(src, 0, 0, 1, &na, &nf);
  CHECK(na == n && nf == 0);
  const int64_t stride = dr_host_ev_stride(src);
  std::vector<int64_t> sk(na), sf(na), sv(na), poff(1001), fk(1), ff(1), fv(1), fpoff(1001);
  std::vector<float> rows((size_t)na * stride);
  dr_host_ev_snapshot_read(src, sk.data(), rows.data(), sf.data(), sv.data(), poff.data(), fk.data(), ff.data(), fv.data(), fpoff.data());
  dr_host_ev_snapshot_end(src);
  CHECK(poff[0] == 0 && poff[1000] == n);
  for (int b = 0; b < 1000; ++b)
    for (int64_t i = poff[b]; i < poff[b + 1]; ++i) {
      CHECK(sk[i] % 1000 == b && sf[i] == 1 && sv[i] == 1);
      if (i > poff[b]) CHECK(sk[i - 1] < sk[i]);                         // bucket-major, key-sorted inside a bucket
    }
  // import while readers probe the growing table (new rows are published with release semantics; patching EXISTING rows under readers
  // is ImportCow's job, not Import's), then the same batch with every key twice into a third table: one row per key
  void* dst = dr_host_ev_create(&cfg);
  dr_host_ev_set_default(dst, def.data());
  std::atomic<bool> stop{false};
  std::thread reader([&] {
    std::vector<int64_t> k(512), fr(512);
    std::vector<float> out(512 * dim);
    uint64_t x = 1;
    while (!stop.load()) {
      for (int i = 0; i < 512; ++i) { x = x * 6364136223846793005ull + 1442695040888963407ull; k[i] = sk[(x >> 33) % (uint64_t)n]; }
      dr_host_ev_lookup(dst, k.data(), 512, out.data());
      dr_host_ev_get_freq(dst, k.data(), 512, fr.data());
      for (int i = 0; i < 512; ++i) CHECK(fr[i] == 0 || fr[i] == 1);
    }
  });
  CHECK(dr_host_ev_import(dst, sk.data(), rows.data(), stride, sf.data(), sv.data(), n, 0, 1, 0) == n);
  stop = true; reader.join();
  CHECK(dr_host_ev_size(dst) == n && dr_host_ev_total_keys(dst) == n);
  std::vector<float> got((size_t)n * dim);
  dr_host_ev_lookup(dst, sk.data(), n, got.data());
  for (int64_t i = 0; i < n; ++i) for (int d = 0; d < dim; ++d) CHECK(got[i * dim + d] == rows[i * stride + d]);
  {
    void* dup = dr_host_ev_create(&cfg);
    dr_host_ev_set_default(dup, def.data());
    std::vector<int64_t> k2(sk); k2.insert(k2.end(), sk.begin(), sk.end());
    std::vector<int64_t> f2(sf); f2.insert(f2.end(), sf.begin(), sf.end());
    std::vector<int64_t> v2(sv); v2.insert(v2.end(), sv.begin(), sv.end());
    std::vector<float> r2(rows); r2.insert(r2.end(), rows.begin(), rows.end());
    CHECK(dr_host_ev_import(dup, k2.data(), r2.data(), stride, f2.data(), v2.data(), 2 * n, 0, 1, 0) == 2 * n);
    CHECK(dr_host_ev_size(dup) == n && dr_host_ev_total_keys(dup) == n);     // a duplicated key owns exactly one row
    dr_host_ev_destroy(dup);
  }
  // eviction (every key has version 1, steps_to_live 5): parallel per-partition rebuild, free lists extended once
  CHECK(dr_host_ev_shrink(dst, 3) == 0 && dr_host_ev_size(dst) == n);
  CHECK(dr_host_ev_shrink(dst, 100) == n && dr_host_ev_size(dst) == 0 && dr_host_ev_total_keys(dst) == 0);
  CHECK(dr_host_ev_import(dst, sk.data(), rows.data(), stride, sf.data(), sv.data(), n, 1, 3, 1) > 0);   // partition 1 of 3, versions reset
  dr_host_ev_snapshot_begin(dst, 0, 0, 1, &na, &nf);
  std::vector<int64_t> tk(na), tf(na), tv(na);
  std::vector<float> trow((size_t)na * stride);
  dr_host_ev_snapshot_read(dst, tk.data(), trow.data(), tf.data(), tv.data(), poff.data(), fk.data(), ff.data(), fv.data(), fpoff.data());
  dr_host_ev_snapshot_end(dst);
  CHECK(na == dr_host_ev_size(dst));
  for (int64_t i = 0; i < na; ++i) CHECK((tk[i] % 1000) % 3 == 1 && tv[i] == -1);
  dr_host_ev_clear_dirty(dst);
  dr_host_ev_snapshot_begin(dst, 1, 0, 1, &na, &nf);
  CHECK(na == 0);
  dr_host_ev_snapshot_end(dst);
  dr_host_ev_destroy(src); dr_host_ev_destroy(dst);
}

// 3. SSD store: writers, readers, removers and compaction all at once
static void TestSsdStore(const char* dir) {
  const int stride = 8;
  void* s = dr_ssd_create(dir, stride, (24 + stride * 4) * 256, /*async_compaction=*/1);
  std::atomic<bool> stop{false};
  std::vector<std::thread> th;
  for (int t = 0; t < 3; ++t)
    th.emplace_back([&, t] {
      std::vector<int64_t> k(128), f(128), v(128);
      std::vector<float> r(128 * stride);
      for (int rep = 0; rep < 60; ++rep) {
        for (int i = 0; i < 128; ++i) { k[i] = t * 100000 + (rep % 6) * 128 + i; f[i] = rep; v[i] = rep; for (int d = 0; d < stride; ++d) r[i * stride + d] = (float)k[i]; }
        dr_ssd_put(s, k.data(), r.data(), f.data(), v.data(), 128);
      }
    });
  th.emplace_back([&] {
    std::vector<int64_t> k(128), f(128), v(128);
    std::vector<float> r(128 * stride);
    std::vector<uint8_t> found(128);
    uint64_t x = 1234567;
    while (!stop.load()) {
      for (int i = 0; i < 128; ++i) { x = x * 6364136223846793005ull + 1442695040888963407ull; k[i] = (int64_t)((x >> 33) % 3) * 100000 + (int64_t)((x >> 12) % 768); }
      dr_ssd_get(s, k.data(), 128, r.data(), f.data(), v.data(), found.data());
      for (int i = 0; i < 128; ++i) if (found[i]) CHECK(r[i * stride] == (float)k[i]);      // a found record is never torn / foreign
    }
  });
  th.emplace_back([&] { while (!stop.load()) dr_ssd_compact(s, 0.3); });
  for (int t = 0; t < 3; ++t) th[t].join();
  stop = true;
  th[3].join(); th[4].join();
  CHECK(dr_ssd_size(s) == 3 * 768);
  std::vector<int64_t> k(768);
  for (int i = 0; i < 768; ++i) k[i] = i;
  CHECK(dr_ssd_remove(s, k.data(), 768) == 768);
  CHECK(dr_ssd_size(s) == 2 * 768);
  dr_ssd_destroy(s);
}

// 4. staging queue: producers / consumers, every ticket exactly once
static void TestStagingQueue() {
  void* q = dr_stage_create(4);
  std::atomic<int64_t> sum{0}, taken{0};
  std::vector<std::thread> th;
  for (int p = 0; p < 3; ++p) th.emplace_back([&, p] { for (int i = 0; i < 500; ++i) CHECK(dr_stage_put(q, p * 1000 + i, 10000) == 0); });
  for (int c = 0; c < 2; ++c)
    th.emplace_back([&] {
      int64_t t;
      while (dr_stage_take(q, &t, 10000) == 0) { sum += t; ++taken; }
    });
  for (int p = 0; p < 3; ++p) th[p].join();
  dr_stage_close(q);
  th[3].join(); th[4].join();
  int64_t expect = 0;
  for (int p = 0; p < 3; ++p) for (int i = 0; i < 500; ++i) expect += p * 1000 + i;
  CHECK(taken.load() == 1500 && sum.load() == expect);
  dr_stage_destroy(q);
}

int main(int argc, char** argv) {
  TestLookupWhileTraining();
  TestFilterRemoveSnapshot();
  TestParallelSnapshotImportShrink();
  TestSsdStore(argc > 1 ? argv[1] : "/tmp/deeprec_ssd_stress");
  TestStagingQueue();
  if (g_fail) { fprintf(stderr, "%d checks failed\n", g_fail); return 1; }
  printf("HOST_STRESS_OK\n");
  return 0;
}
