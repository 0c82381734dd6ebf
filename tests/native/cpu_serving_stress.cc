// Sanitizer tier for the CPU serving runtime (csrc/host/cpu_serving.cc), built with -fsanitize=thread and =address:
// a tiny DLRM-shaped model is written with the bundle writer, served through the C ABI from several client threads (compact and
// protobuf requests, RR and MOD session selection) WHILE the updater thread applies delta updates and swaps in new full versions.
// Checks: every reply is well-formed and finite, versions only move forward, garbage is rejected, nothing races / leaks / overflows.
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../deeprec_b200/csrc/common/bundle.h"
#include "../../deeprec_b200/csrc/common/predict_pb.h"

extern "C" {
void* dr_cpu_initialize(const char* model_entry, const char* model_config, int* state);
int dr_cpu_process(void* model_buf, const void* input_data, int input_size, void** output_data, int* output_size);
int dr_cpu_get_serving_model_info(void* model_buf, void** output_data, int* output_size);
void dr_cpu_serving_release(void* model_buf);
void dr_cpu_serving_free(void* p);
}

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static constexpr int ND = 5, T = 3, D = 8, H0 = 16, TOP0 = 16;      // bottom MLP 5 -> 16 -> 8, top MLP inter -> 16, logits
static int pad8(int n) { return (n + 7) / 8 * 8; }

static void AddF(dr::BundleWriter& w, const std::string& name, const std::vector<float>& v, std::vector<int64_t> shape) {
  CHECK(w.Add(name.c_str(), "float32", shape.data(), (int)shape.size(), v.data(), (int64_t)v.size() * 4) == 0);
}
static void AddI(dr::BundleWriter& w, const std::string& name, const std::vector<int64_t>& v) {
  int64_t n = (int64_t)v.size();
  CHECK(w.Add(name.c_str(), "int64", &n, 1, v.data(), n * 8) == 0);
}

static void WriteDense(dr::BundleWriter& w, std::mt19937& rng) {
  std::normal_distribution<float> nd(0.f, 0.3f);
  auto rnd = [&](size_t n) { std::vector<float> v(n); for (auto& x : v) x = nd(rng); return v; };
  const int bot[2] = {H0, D}; int k = ND;
  for (int l = 0; l < 2; ++l) {
    const std::string nm = "mlp_bot_" + std::to_string(l); const int N = bot[l], Kp = pad8(k);
    AddF(w, "dense/" + nm + "/kernel", rnd((size_t)N * Kp), {N, Kp}); AddF(w, "dense/" + nm + "/bias", rnd(N), {N});
    AddF(w, "dense/" + nm + "/bn_gamma", std::vector<float>(N, 1.f), {N}); AddF(w, "dense/" + nm + "/bn_beta", rnd(N), {N});
    AddF(w, "bn/" + nm + "/moving_mean", rnd(N), {N}); AddF(w, "bn/" + nm + "/moving_variance", std::vector<float>(N, 1.f), {N});
    k = N;
  }
  const int F = T + 1, inter = D + F * (F - 1) / 2;
  AddF(w, "dense/mlp_top_0/kernel", rnd((size_t)TOP0 * pad8(inter)), {TOP0, pad8(inter)}); AddF(w, "dense/mlp_top_0/bias", rnd(TOP0), {TOP0});
  AddF(w, "dense/logits/kernel", rnd(TOP0), {TOP0}); AddF(w, "dense/logits/bias", rnd(1), {1});
}

static void WriteFull(const std::string& root, const std::string& name, int64_t version, std::mt19937& rng) {
  const std::string dir = root + "/" + name;
  mkdir(dir.c_str(), 0755); mkdir((dir + "/variables").c_str(), 0755);
  {
    dr::BundleWriter w(dir + "/variables/variables");
    WriteDense(w, rng);
    std::normal_distribution<float> nd(0.f, 0.5f);
    for (int t = 0; t < T; ++t) {
      std::vector<int64_t> keys; std::vector<float> vals, def(64 * D);
      for (int i = 0; i < 200; ++i) keys.push_back(i * 3 + t);
      vals.resize(keys.size() * D); for (auto& x : vals) x = nd(rng); for (auto& x : def) x = nd(rng);
      const std::string b = "table/" + std::to_string(t);
      AddI(w, b + "-keys", keys); AddF(w, b + "-values", vals, {(int64_t)keys.size(), D}); AddF(w, b + "-default", def, {64, D});
    }
    CHECK(w.Close() == 0);
  }
  FILE* f = fopen((dir + "/saved_model.json").c_str(), "w");
  fprintf(f, "{\"model\": \"dlrm\", \"version\": %lld, \"num_dense\": %d, \"num_tables\": %d, \"embedding_dim\": %d, \"mlp_bot\": [%d, %d], \"mlp_top\": [%d], "
             "\"bn_eps\": 0.001, \"variables\": \"variables/variables\"}", (long long)version, ND, T, D, H0, D, TOP0);
  fclose(f);
}

static void WriteDelta(const std::string& root, int64_t version, std::mt19937& rng) {
  mkdir((root + "/.incr").c_str(), 0755);
  dr::BundleWriter w(root + "/.incr/delta-" + std::to_string(version));
  WriteDense(w, rng);
  std::normal_distribution<float> nd(0.f, 0.5f);
  for (int t = 0; t < T; ++t) {
    std::vector<int64_t> keys; for (int i = 0; i < 40; ++i) keys.push_back((int64_t)(rng() % 900));
    std::vector<float> vals(keys.size() * D); for (auto& x : vals) x = nd(rng);
    const std::string b = "table/" + std::to_string(t);
    AddI(w, b + "-sparse_incr_keys", keys); AddF(w, b + "-sparse_incr_values", vals, {(int64_t)keys.size(), D});
  }
  CHECK(w.Close() == 0);
}

static void WriteVersions(const std::string& root, const std::string& full_dir, int64_t full_v, const std::vector<int64_t>& deltas) {
  std::string s = "{\"full\": {\"version\": " + std::to_string(full_v) + ", \"dir\": \"" + full_dir + "\"}, \"deltas\": [";
  for (size_t i = 0; i < deltas.size(); ++i)
    s += std::string(i ? ", " : "") + "{\"version\": " + std::to_string(deltas[i]) + ", \"base\": " + std::to_string(full_v) + ", \"prefix\": \"" + root + "/.incr/delta-" + std::to_string(deltas[i]) + "\"}";
  s += "]}";
  FILE* f = fopen((root + "/serving_versions.json.tmp").c_str(), "w"); fputs(s.c_str(), f); fclose(f);
  rename((root + "/serving_versions.json.tmp").c_str(), (root + "/serving_versions.json").c_str());
}

static std::string WireRequest(std::mt19937& rng, int B) {
  drpb::WireReq h{drpb::kWireReqMagic, 1, (uint32_t)B, ND, T, 0};
  std::string s(reinterpret_cast<const char*>(&h), sizeof(h));
  std::vector<float> d((size_t)B * ND); for (auto& x : d) x = (float)(rng() % 100) / 30.f;
  std::vector<int64_t> ids((size_t)T * B); for (auto& x : ids) x = (int64_t)(rng() % 1200);       // some stored, some unseen
  s.append(reinterpret_cast<const char*>(d.data()), d.size() * 4); s.append(reinterpret_cast<const char*>(ids.data()), ids.size() * 8);
  return s;
}

static std::string ProtoRequest(std::mt19937& rng, int B) {
  drpb::Request r;
  drpb::Array d; d.dtype = drpb::DT_FLOAT; d.shape = {B, ND}; for (int i = 0; i < B * ND; ++i) d.f32.push_back((float)(rng() % 100) / 30.f);
  drpb::Array i; i.dtype = drpb::DT_INT64; i.shape = {T, B}; for (int k = 0; k < B * T; ++k) i.i64.push_back((int64_t)(rng() % 1200));
  r.inputs.emplace_back("dense", d); r.inputs.emplace_back("ids", i);
  std::string s; drpb::EncodeRequest(r, &s); return s;
}

int main(int argc, char** argv) {
  CHECK(argc > 1);
  const std::string root = argv[1];
  mkdir(root.c_str(), 0755);
  std::mt19937 rng(5);
  WriteFull(root, "v1", 10, rng);
  WriteVersions(root, root + "/v1", 10, {});
  const std::string cfg = "{\"session_num\": 3, \"select_session_policy\": \"MOD\", \"max_batch\": 16, \"checkpoint_dir\": \"" + root +
                          "\", \"model_update_interval_ms\": 20, \"intra_op_parallelism_threads\": 1, \"enable_batching\": true, "
                          "\"max_batch_size\": 24, \"batch_timeout_micros\": 300}";     // requests of <= 12 rows are merged, larger ones bypass the batcher
  int state = -1;
  void* h = dr_cpu_initialize((root + "/v1").c_str(), cfg.c_str(), &state);
  CHECK(h && state == 0);

  std::atomic<bool> stop{false};
  std::atomic<int64_t> served{0}, max_version{0};
  auto client = [&](int tid) {
    std::mt19937 r(100 + tid);
    int64_t last = 0;
    while (!stop) {
      const int B = 1 + (int)(r() % 40);                      // > max_batch exercises the chunking
      const bool proto = (r() & 3) == 0;
      const std::string req = proto ? ProtoRequest(r, B) : WireRequest(r, B);
      void* out = nullptr; int n = 0;
      const int rc = dr_cpu_process(h, req.data(), (int)req.size(), &out, &n);
      CHECK(rc == 200 && out && n > 0);
      int64_t version = -1;
      if (proto) {
        drpb::Response resp; CHECK(drpb::ParseResponse(out, (size_t)n, &resp));
        for (auto& kv : resp.outputs) {
          if (kv.first == "probabilities") { CHECK((int)kv.second.f32.size() == B); for (float p : kv.second.f32) CHECK(p >= 0.f && p <= 1.f); }
          if (kv.first == "model_version") version = kv.second.i64[0];
        }
      } else {
        drpb::WireResp rh; memcpy(&rh, out, sizeof(rh));
        CHECK(rh.magic == drpb::kWireRespMagic && (int)rh.batch == B && n == (int)(sizeof(rh) + (size_t)B * 4));
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(out) + sizeof(rh));
        for (int i = 0; i < B; ++i) CHECK(p[i] >= 0.f && p[i] <= 1.f);
        version = rh.model_version;
      }
      CHECK(version >= last);                                  // a client never sees the model go backwards
      last = version;
      int64_t mv = max_version.load(); while (version > mv && !max_version.compare_exchange_weak(mv, version)) {}
      dr_cpu_serving_free(out);
      if ((r() & 63) == 0) { void* o2 = nullptr; int n2 = 0; CHECK(dr_cpu_process(h, "garbage!", 8, &o2, &n2) == 500); }
      served++;
    }
  };
  std::vector<std::thread> ts;
  for (int i = 0; i < 4; ++i) ts.emplace_back(client, i);

  // publisher: two deltas on v10, a full v20, a delta on it, a broken full version (must be skipped), a full v30
  auto wait_ms = [](int ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); };
  wait_ms(150);
  WriteDelta(root, 11, rng); WriteVersions(root, root + "/v1", 10, {11}); wait_ms(150);
  WriteDelta(root, 12, rng); WriteVersions(root, root + "/v1", 10, {11, 12}); wait_ms(150);
  WriteFull(root, "v2", 20, rng); WriteVersions(root, root + "/v2", 20, {}); wait_ms(250);
  WriteDelta(root, 21, rng); WriteVersions(root, root + "/v2", 20, {21}); wait_ms(150);
  mkdir((root + "/broken").c_str(), 0755); { FILE* f = fopen((root + "/broken/saved_model.json").c_str(), "w"); fputs("{\"version\": 25", f); fclose(f); }
  WriteVersions(root, root + "/broken", 25, {}); wait_ms(200);
  WriteFull(root, "v3", 30, rng); WriteVersions(root, root + "/v3", 30, {}); wait_ms(300);
  stop = true;
  for (auto& t : ts) t.join();

  void* info = nullptr; int n = 0;
  CHECK(dr_cpu_get_serving_model_info(h, &info, &n) == 200);
  const std::string s(static_cast<const char*>(info), (size_t)n);
  dr_cpu_serving_free(info);
  printf("%s\nserved %lld requests, newest version seen by a client: %lld\n", s.c_str(), (long long)served.load(), (long long)max_version.load());
  CHECK(s.find("\"model_version\": 30") != std::string::npos && s.find("\"full_updates\": 2") != std::string::npos);
  CHECK(max_version.load() == 30 && served.load() > 50);
  CHECK(s.find("\"merged_requests\": 0,") == std::string::npos && s.find("\"max_batch_size\": 24") != std::string::npos);
  dr_cpu_serving_release(h);
  printf("CPU_SERVING_STRESS_OK\n");
  return 0;
}
