// Native load generator for the CPU Processor C ABI (python client threads top out at ~10 k requests/s on the GIL):
//   cpu_serving_load <libdeeprec_host.so> <saved_model_dir> <config json> <threads> <requests per thread> <rows per request> <num_dense> <num_sparse>
// Prints one JSON line: qps, samples/s, p50 / p99 latency.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

struct WireReq { uint32_t magic, version, batch, num_dense, num_sparse, reserved; };   // csrc/common/predict_pb.h
using InitFn = void* (*)(const char*, const char*, int*);
using ProcFn = int (*)(void*, const void*, int, void**, int*);
using InfoFn = int (*)(void*, void**, int*);
using FreeFn = void (*)(void*);
using RelFn = void (*)(void*);

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: see the header comment\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 1; }
  auto init = (InitFn)dlsym(lib, "dr_cpu_initialize"); auto proc = (ProcFn)dlsym(lib, "dr_cpu_process");
  auto info = (InfoFn)dlsym(lib, "dr_cpu_get_serving_model_info"); auto fre = (FreeFn)dlsym(lib, "dr_cpu_serving_free"); auto rel = (RelFn)dlsym(lib, "dr_cpu_serving_release");
  const int threads = atoi(argv[4]), per = atoi(argv[5]), rows = atoi(argv[6]), nd = atoi(argv[7]), ns = atoi(argv[8]);
  int state = -1;
  void* h = init(argv[2], argv[3], &state);
  if (!h || state != 0) { fprintf(stderr, "initialize failed: %d\n", state); return 1; }
  std::vector<std::vector<double>> lat((size_t)threads);
  auto client = [&](int t) {
    std::mt19937_64 rng(1234 + t);
    std::string req(sizeof(WireReq) + (size_t)rows * nd * 4 + (size_t)ns * rows * 8, '\0');
    WireReq hd{0x51525244u, 1u, (uint32_t)rows, (uint32_t)nd, (uint32_t)ns, 0u};
    memcpy(&req[0], &hd, sizeof(hd));
    float* d = reinterpret_cast<float*>(&req[sizeof(hd)]); int64_t* ids = reinterpret_cast<int64_t*>(&req[sizeof(hd) + (size_t)rows * nd * 4]);
    lat[(size_t)t].reserve((size_t)per);
    for (int i = 0; i < per; ++i) {
      for (int k = 0; k < rows * nd; ++k) d[k] = (float)(rng() % 1000) / 500.f - 1.f;
      for (int k = 0; k < ns * rows; ++k) ids[k] = (int64_t)(rng() % 1000);
      void* out = nullptr; int n = 0;
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = proc(h, req.data(), (int)req.size(), &out, &n);
      lat[(size_t)t].push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      if (rc != 200) { fprintf(stderr, "process -> %d\n", rc); exit(1); }
      fre(out);
    }
  };
  for (int warm = 0; warm < 1; ++warm) { std::vector<double> keep; client(0); lat[0].clear(); }
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; ++t) ts.emplace_back(client, t);
  for (auto& t : ts) t.join();
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::vector<double> all;
  for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end());
  void* js = nullptr; int jn = 0; info(h, &js, &jn);
  printf("{\"client_threads\": %d, \"rows_per_request\": %d, \"requests\": %zu, \"qps\": %.1f, \"samples_per_s\": %.1f, \"p50_ms\": %.4f, \"p99_ms\": %.4f, \"model_info\": %.*s}\n",
         threads, rows, all.size(), all.size() / wall, all.size() * (double)rows / wall, all[all.size() / 2], all[(size_t)(all.size() * 0.99)], jn, (const char*)js);
  fre(js); rel(h);
  return 0;
}
