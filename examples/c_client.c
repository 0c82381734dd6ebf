/* A C integrator of the serving ABI (what an RPC front-end does): dlopen the runtime, initialize a model directory, send one protobuf
 * PredictRequest and one compact request, print the probabilities.
 *
 *   gcc examples/c_client.c -Ideeprec_b200/csrc/include -ldl -o /tmp/c_client
 *   /tmp/c_client deeprec_b200/lib/libdeeprec_host.so <saved_model_dir>
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deeprec_processor.h"

#define ND 13
#define NS 26
#define B 4

typedef void* (*init_fn)(const char*, const char*, int*);
typedef int (*process_fn)(void*, const void*, int, void**, int*);
typedef int (*info_fn)(void*, void**, int*);
typedef void (*free_fn)(void*);
typedef int (*enc_fn)(const float*, const int64_t*, int64_t, int, int, int, const char*, const char*, void**, int64_t*);
typedef int64_t (*dec_fn)(const void*, int64_t, float*, int64_t, int64_t*);

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s libdeeprec_host.so <saved_model_dir>\n", argv[0]); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  init_fn init = (init_fn)dlsym(lib, "dr_cpu_initialize"); process_fn proc = (process_fn)dlsym(lib, "dr_cpu_process");
  info_fn info = (info_fn)dlsym(lib, "dr_cpu_get_serving_model_info"); free_fn sfree = (free_fn)dlsym(lib, "dr_cpu_serving_free");
  free_fn release = (free_fn)dlsym(lib, "dr_cpu_serving_release"); enc_fn enc = (enc_fn)dlsym(lib, "dr_pb_encode_request");
  dec_fn dec = (dec_fn)dlsym(lib, "dr_pb_decode_response"); free_fn pbfree = (free_fn)dlsym(lib, "dr_pb_free");
  if (!init || !proc || !info || !sfree || !release || !enc || !dec || !pbfree) { fprintf(stderr, "missing symbol\n"); return 1; }

  int state = -1;
  void* model = init(argv[2], "{\"session_num\": 2, \"model_update_interval_ms\": 0}", &state);
  if (!model || state != 0) { fprintf(stderr, "initialize failed (%d)\n", state); return 1; }

  float dense[B * ND]; int64_t ids[NS * B];
  for (int i = 0; i < B * ND; ++i) dense[i] = (float)(i % 7) * 0.25f;
  for (int t = 0; t < NS; ++t) for (int b = 0; b < B; ++b) ids[t * B + b] = (int64_t)((t * 31 + b * 7) % 50);

  /* 1. protobuf PredictRequest (per-feature inputs I1..I13, C1..C26) */
  void* req = NULL; int64_t req_n = 0;
  if (enc(dense, ids, B, ND, NS, 1, "serving_default", "", &req, &req_n) != 0) return 1;
  void* out = NULL; int out_n = 0;
  int rc = proc(model, req, (int)req_n, &out, &out_n);
  float p_pb[B]; int64_t version = -1;
  if (rc != 200 || dec(out, out_n, p_pb, B, &version) != B) { fprintf(stderr, "protobuf request failed (%d)\n", rc); return 1; }
  pbfree(req); sfree(out);

  /* 2. the compact format */
  struct dr_wire_request h = {0x51525244u, 1u, B, ND, NS, 0u};
  char wire[sizeof(h) + sizeof(dense) + sizeof(ids)];
  memcpy(wire, &h, sizeof(h)); memcpy(wire + sizeof(h), dense, sizeof(dense)); memcpy(wire + sizeof(h) + sizeof(dense), ids, sizeof(ids));
  rc = proc(model, wire, (int)sizeof(wire), &out, &out_n);
  struct dr_wire_response rh; memcpy(&rh, out, sizeof(rh));
  const float* p_wire = (const float*)((const char*)out + sizeof(rh));
  if (rc != 200 || rh.batch != B) { fprintf(stderr, "compact request failed (%d)\n", rc); return 1; }
  for (int b = 0; b < B; ++b) {
    printf("sample %d: p(protobuf) = %.6f  p(compact) = %.6f\n", b, p_pb[b], p_wire[b]);
    if (p_pb[b] != p_wire[b]) { fprintf(stderr, "encodings disagree\n"); return 1; }
  }
  sfree(out);
  void* js = NULL; int js_n = 0;
  info(model, &js, &js_n); printf("model version %lld; %.*s\n", (long long)version, js_n, (const char*)js); sfree(js);
  release(model);
  printf("C_CLIENT_OK\n");
  return 0;
}
