// Stress driver for the native parameter-server data plane (csrc/host/ps_server.cc) under ThreadSanitizer / AddressSanitizer:
// W worker threads hammer ONE server with fused PULL / PUSH messages over real TCP connections while the main thread toggles the scaling
// fence (frozen flag + definition version).  Checked: every accepted push was applied exactly once (frequency of a key == accepted pushes that
// carried it ... counted through the row values of an SGD table), pulls never observe a torn message, STALE answers leave the stream in sync,
// the in-flight count drains to zero under the fence, stop() joins every thread with connections still open.
#include <chrono>
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
void dr_host_ev_set_default(void* h, const float* m);
void dr_host_ev_lookup(void* h, const int64_t* keys, int64_t n, float* out);
void* dr_ps_server_start(const char* bind_addr, int port, int* bound_port);
int dr_ps_server_add_table(void* sv, void* host_ev, int dim, const DrOptHyper* hp);
void dr_ps_server_set_def(void* sv, int def_version, int frozen);
int dr_ps_server_inflight(void* sv);
void dr_ps_server_stats(void* sv, uint64_t* out);
void dr_ps_server_stop(void* sv);
void* dr_ps_client_connect(const char* host, int port);
void dr_ps_client_close(void* c);
int dr_ps_client_send(void* c, int op, int def_version, int nt, const int* table_ids, const int64_t* n, const int64_t* const* keys, const float* const* grads, const int* dims);
int dr_ps_client_recv(void* c, int nt, const int64_t* n, const int* dims, float* const* rows, uint64_t* aux);
}

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  constexpr int D = 8, W = 4, STEPS = 300, NK = 64, KEYS = 500;
  DrEvConfig c{};
  c.dim = D; c.num_slots = 0; c.has_scalars = 0; c.init_capacity = 1024; c.default_value_dim = 1; c.num_partitions = 4; c.record_freq = 1; c.record_version = 1;
  c.l2_weight_threshold = -1.f;
  void* ev[2] = {dr_host_ev_create(&c), dr_host_ev_create(&c)};
  std::vector<float> zero(D, 0.f);
  for (void* e : ev) dr_host_ev_set_default(e, zero.data());                 // rows start at 0: after training row[k][0] = -lr * (sum of pushed grads)
  int port = 0;
  void* sv = dr_ps_server_start("127.0.0.1", 0, &port);
  CHECK(sv && port > 0);
  DrOptHyper hp{}; hp.kind = DR_OPT_SGD; hp.lr = 1.0f;
  int tid[2] = {dr_ps_server_add_table(sv, ev[0], D, &hp), dr_ps_server_add_table(sv, ev[1], D, &hp)};
  CHECK(tid[0] == 0 && tid[1] == 1);

  std::atomic<int> def{0};
  std::atomic<bool> failed{false};
  std::vector<std::vector<double>> applied(W, std::vector<double>(2 * KEYS, 0.0));    // per worker: sum of gradients the server ACCEPTED, per (table, key)
  std::atomic<uint64_t> stale{0}, accepted{0};
  auto worker = [&](int w) {
    void* pull = dr_ps_client_connect("127.0.0.1", port);
    void* push = dr_ps_client_connect("127.0.0.1", port);
    if (!pull || !push) { failed = true; return; }
    uint64_t rng = 0x9E3779B97F4A7C15ull * (w + 1);
    auto next = [&] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    std::vector<int64_t> k0(NK), k1(NK); std::vector<float> g0(NK * D), g1(NK * D), r0(NK * D), r1(NK * D);
    for (int s = 0; s < STEPS && !failed; ++s) {
      for (int i = 0; i < NK; ++i) { k0[i] = (int64_t)(next() % KEYS); k1[i] = (int64_t)(next() % KEYS); }
      for (int i = 0; i < NK; ++i) for (int d = 0; d < D; ++d) { g0[i * D + d] = (float)((i + d) % 3 + 1); g1[i * D + d] = (float)((i * 2 + d) % 5 + 1); }
      const int ids[2] = {0, 1}; const int64_t n[2] = {NK, NK}; const int dims[2] = {D, D};
      const int64_t* keys[2] = {k0.data(), k1.data()}; const float* grads[2] = {g0.data(), g1.data()}; float* rows[2] = {r0.data(), r1.data()};
      const int dv = def.load();
      if (dr_ps_client_send(pull, 1, dv, 2, ids, n, keys, nullptr, dims) != 0) { failed = true; break; }
      int rc = dr_ps_client_recv(pull, 2, n, dims, rows, nullptr);
      if (rc < 0 || rc == 2) { failed = true; break; }
      if (rc == 0) for (int i = 0; i < NK * D; ++i) if (!(r0[i] <= 0.f && r1[i] <= 0.f)) { failed = true; break; }     // SGD with positive grads, lr 1: rows only decrease
      if (dr_ps_client_send(push, 2, dv, 2, ids, n, keys, grads, dims) != 0) { failed = true; break; }
      uint64_t aux = 0;
      rc = dr_ps_client_recv(push, 0, nullptr, nullptr, nullptr, &aux);
      if (rc < 0 || rc == 2) { failed = true; break; }
      if (rc == 1) { stale++; continue; }
      accepted++;
      for (int i = 0; i < NK; ++i) { applied[w][(size_t)k0[i]] += g0[i * D]; applied[w][(size_t)(KEYS + k1[i])] += g1[i * D]; }
    }
    dr_ps_client_close(pull); dr_ps_client_close(push);
  };
  std::vector<std::thread> ts;
  for (int w = 0; w < W; ++w) ts.emplace_back(worker, w);
  // the scaling fence, toggled while the workers run: freeze -> wait for a drained server -> bump the definition -> thaw
  for (int round = 0; round < 20 && !failed; ++round) {
    std::this_thread::sleep_for(std::chrono::milliseconds(3));
    dr_ps_server_set_def(sv, def.load(), 1);
    // time-bounded (not spin-count-bounded): under TSAN on a loaded box a handler thread can stay descheduled for seconds.  ONE observation of
    // zero after the freeze means drained (nothing can be admitted any more); later reads may see the transient +1 of a request that Admit() is
    // in the middle of rejecting, so the verdict is the observation itself, not a second read
    bool drained = false;
    for (auto t0 = std::chrono::steady_clock::now(); std::chrono::steady_clock::now() - t0 < std::chrono::seconds(120);) {
      if (dr_ps_server_inflight(sv) == 0) { drained = true; break; }
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    CHECK(drained);
    def.fetch_add(1);
    dr_ps_server_set_def(sv, def.load(), 0);
  }
  for (auto& t : ts) t.join();
  CHECK(!failed.load());
  uint64_t st[6]; dr_ps_server_stats(sv, st);
  CHECK(st[1] == accepted.load() * 2);                      // table-level pushes applied == accepted messages x 2 tables
  CHECK(st[4] >= stale.load());
  // every accepted gradient was applied exactly once: row[k][0] == -sum of accepted g[., 0]
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < KEYS; ++k) {
      double want = 0; for (int w = 0; w < W; ++w) want += applied[w][(size_t)(t * KEYS + k)];
      int64_t key = k; float row[D];
      dr_host_ev_lookup(ev[t], &key, 1, row);
      if (!(row[0] <= 0.f && (double)-row[0] > want - 0.5 && (double)-row[0] < want + 0.5)) { fprintf(stderr, "table %d key %d: row %f want %f\n", t, k, row[0], -want); return 1; }
    }
  // stop() with a connection still open must join its thread
  void* idle = dr_ps_client_connect("127.0.0.1", port);
  CHECK(idle != nullptr);
  dr_ps_server_stop(sv);
  dr_ps_client_close(idle);
  for (void* e : ev) dr_host_ev_destroy(e);
  printf("PS_STRESS_OK accepted=%llu stale=%llu\n", (unsigned long long)accepted.load(), (unsigned long long)stale.load());
  return 0;
}
