// C ABI over common/predict_pb.h: protobuf PredictRequest / PredictResponse (wire-compatible with the reference's
// serving/processor/serving/predict.proto) <-> the serving runtime's compact format.  Used by the python serving layer,
// the HTTP front-end (`:predict_proto`) and C clients; the GPU runtime includes the same header and accepts protobuf
// requests in `process()` directly.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../common/predict_pb.h"

namespace {
thread_local std::string g_err;

int Emit(const std::string& s, void** out, int64_t* out_n) {
  *out_n = (int64_t)s.size();
  *out = malloc(s.size() ? s.size() : 1);
  if (!*out) return -2;
  memcpy(*out, s.data(), s.size());
  return 0;
}
}  // namespace

extern "C" {

const char* dr_pb_last_error() { return g_err.c_str(); }
void dr_pb_free(void* p) { free(p); }

// PredictRequest bytes -> compact "DRRQ" request for a model with (num_dense, num_sparse) inputs.
int dr_pb_request_to_wire(const void* pb, int64_t n, int num_dense, int num_sparse, void** out, int64_t* out_n) {
  drpb::Request r;
  if (!drpb::ParseRequest(pb, (size_t)n, &r)) { g_err = "malformed PredictRequest"; return -1; }
  std::string w;
  if (!drpb::RequestToWire(r, num_dense, num_sparse, &w, &g_err)) return -1;
  return Emit(w, out, out_n);
}

// compact "DRRS" response (+ the request it answers, for output_filter; may be null) -> PredictResponse bytes.
int dr_pb_response_from_wire(const void* wire, int64_t n, const void* request_pb, int64_t request_n, void** out, int64_t* out_n) {
  drpb::Request r;
  if (request_pb && request_n > 0 && !drpb::ParseRequest(request_pb, (size_t)request_n, &r)) { g_err = "malformed PredictRequest"; return -1; }
  std::string o;
  if (!drpb::WireToResponse(wire, (size_t)n, r.output_filter, &o)) { g_err = "malformed wire response"; return -1; }
  return Emit(o, out, out_n);
}

// Client side: build a PredictRequest.  per_feature = 0: inputs {"dense": float[B,nd], "ids": int64[ns,B]};
// per_feature = 1: inputs I1..I<nd> (float [B]) and C1..C<ns> (int64 [B]), the modelzoo naming.
int dr_pb_encode_request(const float* dense, const int64_t* ids, int64_t B, int nd, int ns, int per_feature, const char* signature,
                         const char* output_filter, void** out, int64_t* out_n) {
  drpb::Request r;
  if (signature) r.signature_name = signature;
  if (output_filter && *output_filter) r.output_filter.emplace_back(output_filter);
  if (!per_feature) {
    drpb::Array d; d.dtype = drpb::DT_FLOAT; d.shape = {B, nd}; d.f32.assign(dense, dense + B * nd);
    drpb::Array i; i.dtype = drpb::DT_INT64; i.shape = {ns, B}; i.i64.assign(ids, ids + (int64_t)ns * B);
    r.inputs.emplace_back("dense", std::move(d)); r.inputs.emplace_back("ids", std::move(i));
  } else {
    for (int c = 0; c < nd; ++c) {
      drpb::Array d; d.dtype = drpb::DT_FLOAT; d.shape = {B}; d.f32.resize(B);
      for (int64_t b = 0; b < B; ++b) d.f32[b] = dense[b * nd + c];
      r.inputs.emplace_back("I" + std::to_string(c + 1), std::move(d));
    }
    for (int t = 0; t < ns; ++t) {
      drpb::Array i; i.dtype = drpb::DT_INT64; i.shape = {B}; i.i64.assign(ids + (int64_t)t * B, ids + (int64_t)(t + 1) * B);
      r.inputs.emplace_back("C" + std::to_string(t + 1), std::move(i));
    }
  }
  std::string o; drpb::EncodeRequest(r, &o);
  return Emit(o, out, out_n);
}

// Client side: PredictResponse -> probabilities (returns the element count, or -1) and the model version.
int64_t dr_pb_decode_response(const void* pb, int64_t n, float* probs, int64_t cap, int64_t* model_version) {
  drpb::Response r;
  if (!drpb::ParseResponse(pb, (size_t)n, &r)) { g_err = "malformed PredictResponse"; return -1; }
  int64_t count = 0;
  if (model_version) *model_version = -1;
  for (const auto& kv : r.outputs) {
    if (kv.first == "probabilities") { count = (int64_t)kv.second.f32.size(); if (probs) memcpy(probs, kv.second.f32.data(), (size_t)std::min(count, cap) * 4); }
    else if (kv.first == "model_version" && model_version && !kv.second.i64.empty()) *model_version = kv.second.i64[0];
  }
  return count;
}

// probabilities per row of a PredictResponse (the last dimension of a rank-2 "probabilities" output: multi-task models), 1 otherwise
int64_t dr_pb_response_cols(const void* pb, int64_t n) {
  drpb::Response r;
  if (!drpb::ParseResponse(pb, (size_t)n, &r)) { g_err = "malformed PredictResponse"; return -1; }
  for (const auto& kv : r.outputs)
    if (kv.first == "probabilities" && kv.second.shape.size() == 2 && kv.second.shape[1] > 0) return kv.second.shape[1];
  return 1;
}

}  // extern "C"
