// Sanitizer tier for the network-facing / allocator pieces of the host runtime (built with -fsanitize=address,undefined and =thread):
//   * protobuf codec (common/predict_pb.h): random byte strings, truncations and bit flips of valid requests must never crash,
//     read out of bounds or loop; valid requests must round-trip;
//   * TensorPool (common/tensor_pool.h): concurrent alloc / free / step_end from several threads, blocks never overlap.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../deeprec_b200/csrc/common/mini_json.h"
#include "../../deeprec_b200/csrc/common/predict_pb.h"
#include "../../deeprec_b200/csrc/common/tensor_pool.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static std::string ValidRequest(std::mt19937_64& rng, int B, int nd, int ns, bool per_feature) {
  drpb::Request r; r.signature_name = "serving_default"; r.output_filter = {"probabilities"};
  std::uniform_real_distribution<float> fd(-3, 3);
  if (!per_feature) {
    drpb::Array d; d.dtype = drpb::DT_FLOAT; d.shape = {B, nd}; for (int i = 0; i < B * nd; ++i) d.f32.push_back(fd(rng));
    drpb::Array i; i.dtype = drpb::DT_INT64; i.shape = {ns, B}; for (int k = 0; k < B * ns; ++k) i.i64.push_back((int64_t)rng() >> (rng() % 60));
    r.inputs.emplace_back("dense", d); r.inputs.emplace_back("ids", i);
  } else {
    for (int c = 0; c < nd; ++c) { drpb::Array d; d.dtype = drpb::DT_FLOAT; d.shape = {B}; for (int b = 0; b < B; ++b) d.f32.push_back(fd(rng)); r.inputs.emplace_back("I" + std::to_string(c + 1), d); }
    for (int t = 0; t < ns; ++t) { drpb::Array i; i.dtype = drpb::DT_INT64; i.shape = {B}; for (int b = 0; b < B; ++b) i.i64.push_back((int64_t)(rng() % 100000) - 50); r.inputs.emplace_back("C" + std::to_string(t + 1), i); }
  }
  std::string s; drpb::EncodeRequest(r, &s); return s;
}

static void FuzzCodec() {
  std::mt19937_64 rng(1234);
  int64_t accepted = 0, rejected = 0;
  for (int iter = 0; iter < 3000; ++iter) {
    const int B = 1 + (int)(rng() % 9), nd = 1 + (int)(rng() % 13), ns = 1 + (int)(rng() % 26);
    std::string valid = ValidRequest(rng, B, nd, ns, iter & 1);
    {   // round trip of the valid message
      drpb::Request r; std::string wire, err;
      CHECK(drpb::ParseRequest(valid.data(), valid.size(), &r));
      CHECK(drpb::RequestToWire(r, nd, ns, &wire, &err));
      CHECK(wire.size() == sizeof(drpb::WireReq) + (size_t)B * nd * 4 + (size_t)B * ns * 8);
    }
    std::string m = valid;
    switch (iter % 4) {
      case 0: m.resize(rng() % (m.size() + 1)); break;                                                    // truncation
      case 1: for (int k = 0; k < 1 + (int)(rng() % 8); ++k) m[rng() % m.size()] ^= (char)(1u << (rng() % 8)); break;   // bit flips
      case 2: { size_t n = rng() % 200; m.resize(n); for (auto& c : m) c = (char)rng(); break; }          // noise
      case 3: m.insert(rng() % m.size(), std::string(1 + rng() % 12, (char)0xff)); break;                 // runaway varints
    }
    // parse from an exactly-sized heap buffer so that ASAN sees any over-read
    std::vector<uint8_t> buf(m.begin(), m.end());
    drpb::Request r; std::string wire, err;
    if (drpb::ParseRequest(buf.data(), buf.size(), &r) && drpb::RequestToWire(r, nd, ns, &wire, &err)) ++accepted; else ++rejected;
    drpb::Response resp; (void)drpb::ParseResponse(buf.data(), buf.size(), &resp);
    std::string out; (void)drpb::WireToResponse(buf.data(), buf.size(), r.output_filter, &out);
  }
  CHECK(rejected > 500);
  printf("codec fuzz: %lld mutated requests accepted, %lld rejected\n", (long long)accepted, (long long)rejected);
}

static void FuzzJson() {
  std::mt19937_64 rng(77);
  const std::string valid = "{\"session_num\": 4, \"select_session_policy\": \"MOD\", \"mlp_bot\": [512, 256, 64, 16], \"full\": {\"version\": 12, \"dir\": \"/m/v12\"}, "
                            "\"deltas\": [{\"version\": 13, \"base\": 12, \"prefix\": \"/m/.incr/delta-13\"}], \"ok\": true, \"none\": null, \"x\": -1.5e3}";
  drjson::JVal j;
  CHECK(drjson::ParseJson(valid, &j) && j.n("session_num", 0) == 4 && j.s("select_session_policy", "") == "MOD" && j.get("deltas")->arr.size() == 1);
  int64_t ok = 0, bad = 0;
  for (int iter = 0; iter < 4000; ++iter) {
    std::string m = valid;
    switch (iter % 4) {
      case 0: m.resize(rng() % (m.size() + 1)); break;
      case 1: for (int k = 0; k < 1 + (int)(rng() % 6); ++k) m[rng() % m.size()] ^= (char)(1u << (rng() % 8)); break;
      case 2: { size_t n = rng() % 120; m.resize(n); for (auto& c : m) c = (char)rng(); break; }
      case 3: m = std::string(1 + rng() % 200, "[{\""[rng() % 3]); break;                     // deep nesting must be cut off, not overflow the stack
    }
    std::string exact(m.data(), m.size());                                                     // exactly-sized heap copy: over-reads are visible to ASAN
    drjson::JVal v;
    if (drjson::ParseJson(exact, &v)) ++ok; else ++bad;
  }
  CHECK(bad > 1000);
  printf("json fuzz: %lld parsed, %lld rejected\n", (long long)ok, (long long)bad);
}

static void* HostAlloc(size_t n, void*) { return malloc(n); }
static void HostFree(void* p, void*) { free(p); }

static void StressPool() {
  dr::TensorPool pool(HostAlloc, HostFree, nullptr, 1024, 2, 4);
  std::atomic<bool> stop{false};
  std::atomic<int64_t> bad{0};
  auto worker = [&](int tid) {
    std::mt19937_64 rng(tid);
    std::vector<std::pair<uint8_t*, size_t>> live;
    for (int it = 0; it < 4000; ++it) {
      if (live.size() < 6 && (rng() & 1)) {
        size_t n = 512 + rng() % 70000;
        auto* p = static_cast<uint8_t*>(pool.Alloc(n, (uint64_t)(tid % 2)));
        CHECK(p != nullptr);
        memset(p, tid + 1, n);
        live.emplace_back(p, n);
      } else if (!live.empty()) {
        size_t k = rng() % live.size();
        auto [p, n] = live[k];
        for (size_t i = 0; i < n; i += 97) if (p[i] != (uint8_t)(tid + 1)) { bad++; break; }      // nobody else wrote into my block
        pool.Free(p);
        live[k] = live.back(); live.pop_back();
      }
    }
    for (auto& pn : live) pool.Free(pn.first);
  };
  std::thread stepper([&] { while (!stop) { pool.StepEnd(); std::this_thread::sleep_for(std::chrono::microseconds(200)); } });
  std::vector<std::thread> ts;
  for (int t = 0; t < 4; ++t) ts.emplace_back(worker, t);
  for (auto& t : ts) t.join();
  stop = true; stepper.join();
  const dr::TensorPoolStats s = pool.Stats();
  CHECK(bad == 0);
  CHECK(s.live_pool_blocks == 0);
  CHECK(s.phase == 1 && s.pool_hits > 0);
  printf("pool stress: %lld hits, %lld misses, %lld replans, %lld pool bytes\n", (long long)s.pool_hits, (long long)s.pool_misses,
         (long long)s.replans, (long long)s.pool_bytes);
}

int main() {
  FuzzCodec();
  FuzzJson();
  StressPool();
  printf("CODEC_POOL_OK\n");
  return 0;
}
