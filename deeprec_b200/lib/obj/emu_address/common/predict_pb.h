// Protobuf wire codec for the serving request/response messages (header-only, no protoc / libprotobuf).
//
// Wire-compatible with the reference's `tensorflow.eas` messages (serving/processor/serving/predict.proto):
//   ArrayShape      { repeated int64 dim = 1 [packed] }
//   ArrayProto      { ArrayDataType dtype = 1; ArrayShape array_shape = 2; float_val = 3; double_val = 4; int_val = 5;
//                     string_val = 6; int64_val = 7; bool_val = 8 }   (numeric repeated fields packed; unpacked also accepted)
//   PredictRequest  { string signature_name = 1; map<string, ArrayProto> inputs = 2; repeated string output_filter = 3 }
//   PredictResponse { map<string, ArrayProto> outputs = 1 }
// so an existing client SDK (Go / Java / Python demos under serving/sdk/) can talk to this runtime unchanged.
//
// The runtime's own compact format ("DRRQ"/"DRRS", serving_runtime.cu) stays the fast path; `RequestToWire` /
// `WireToResponse` translate between the two, so `process()` accepts either encoding.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace drpb {

enum DType : int { DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_UINT8 = 4, DT_INT16 = 5, DT_INT8 = 6, DT_STRING = 7, DT_INT64 = 9, DT_BOOL = 10 };

struct Array {
  int dtype = DT_INVALID;
  std::vector<int64_t> shape;
  std::vector<float> f32;
  std::vector<double> f64;
  std::vector<int32_t> i32;
  std::vector<int64_t> i64;
  std::vector<uint8_t> b8;
  std::vector<std::string> str;
  int64_t NumElements() const { int64_t n = 1; for (int64_t d : shape) n *= d; return shape.empty() ? (int64_t)std::max({f32.size(), f64.size(), i32.size(), i64.size(), b8.size(), str.size()}) : n; }
};

struct Request {
  std::string signature_name;
  std::vector<std::pair<std::string, Array>> inputs;
  std::vector<std::string> output_filter;
};

struct Response { std::vector<std::pair<std::string, Array>> outputs; };

// ---------------------------------------------------------------- reader
struct Reader {
  const uint8_t* p; const uint8_t* end; bool ok = true;
  Reader(const void* d, size_t n) : p(static_cast<const uint8_t*>(d)), end(p + n) {}
  bool Done() const { return p >= end || !ok; }
  uint64_t Varint() {
    uint64_t v = 0; int shift = 0;
    while (p < end && shift < 64) {
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
    ok = false; return 0;
  }
  uint32_t Fixed32() { if (end - p < 4) { ok = false; return 0; } uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint64_t Fixed64() { if (end - p < 8) { ok = false; return 0; } uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
  Reader Sub() {
    uint64_t n = Varint();
    if (!ok || n > (uint64_t)(end - p)) { ok = false; return Reader(p, 0); }
    Reader r(p, (size_t)n); p += n; return r;
  }
  void Skip(int wt) {
    switch (wt) {
      case 0: Varint(); break;
      case 1: Fixed64(); break;
      case 2: Sub(); break;
      case 5: Fixed32(); break;
      default: ok = false;          // groups (3/4) are not used by these messages
    }
  }
};

inline bool ParseShape(Reader r, std::vector<int64_t>* dims) {
  while (!r.Done()) {
    uint64_t tag = r.Varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
    if (f == 1 && wt == 2) { Reader s = r.Sub(); while (!s.Done()) dims->push_back((int64_t)s.Varint()); if (!s.ok) return false; }
    else if (f == 1 && wt == 0) dims->push_back((int64_t)r.Varint());
    else r.Skip(wt);
  }
  return r.ok;
}

inline bool ParseArray(Reader r, Array* a) {
  while (!r.Done()) {
    uint64_t tag = r.Varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
    switch (f) {
      case 1: if (wt == 0) a->dtype = (int)r.Varint(); else r.Skip(wt); break;
      case 2: if (wt == 2) { if (!ParseShape(r.Sub(), &a->shape)) return false; } else r.Skip(wt); break;
      case 3:
        if (wt == 2) { Reader s = r.Sub(); size_t n = (size_t)(s.end - s.p) / 4, o = a->f32.size(); a->f32.resize(o + n); if (n) memcpy(a->f32.data() + o, s.p, n * 4); }
        else if (wt == 5) { uint32_t v = r.Fixed32(); float x; memcpy(&x, &v, 4); a->f32.push_back(x); }
        else r.Skip(wt);
        break;
      case 4:
        if (wt == 2) { Reader s = r.Sub(); size_t n = (size_t)(s.end - s.p) / 8, o = a->f64.size(); a->f64.resize(o + n); if (n) memcpy(a->f64.data() + o, s.p, n * 8); }
        else if (wt == 1) { uint64_t v = r.Fixed64(); double x; memcpy(&x, &v, 8); a->f64.push_back(x); }
        else r.Skip(wt);
        break;
      case 5:
        if (wt == 2) { Reader s = r.Sub(); while (!s.Done()) a->i32.push_back((int32_t)(int64_t)s.Varint()); if (!s.ok) return false; }
        else if (wt == 0) a->i32.push_back((int32_t)(int64_t)r.Varint());
        else r.Skip(wt);
        break;
      case 6:
        if (wt == 2) { Reader s = r.Sub(); a->str.emplace_back(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p)); }
        else r.Skip(wt);
        break;
      case 7:
        if (wt == 2) { Reader s = r.Sub(); while (!s.Done()) a->i64.push_back((int64_t)s.Varint()); if (!s.ok) return false; }
        else if (wt == 0) a->i64.push_back((int64_t)r.Varint());
        else r.Skip(wt);
        break;
      case 8:
        if (wt == 2) { Reader s = r.Sub(); while (!s.Done()) a->b8.push_back(s.Varint() ? 1 : 0); if (!s.ok) return false; }
        else if (wt == 0) a->b8.push_back(r.Varint() ? 1 : 0);
        else r.Skip(wt);
        break;
      default: r.Skip(wt);
    }
  }
  return r.ok;
}

inline bool ParseMapEntry(Reader r, std::string* key, Array* val) {
  while (!r.Done()) {
    uint64_t tag = r.Varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
    if (f == 1 && wt == 2) { Reader s = r.Sub(); key->assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p)); }
    else if (f == 2 && wt == 2) { if (!ParseArray(r.Sub(), val)) return false; }
    else r.Skip(wt);
  }
  return r.ok;
}

inline bool ParseRequest(const void* data, size_t n, Request* req) {
  Reader r(data, n);
  while (!r.Done()) {
    uint64_t tag = r.Varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
    if (f == 1 && wt == 2) { Reader s = r.Sub(); req->signature_name.assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p)); }
    else if (f == 2 && wt == 2) {
      std::string k; Array a;
      if (!ParseMapEntry(r.Sub(), &k, &a)) return false;
      auto it = std::find_if(req->inputs.begin(), req->inputs.end(), [&](const auto& kv) { return kv.first == k; });
      if (it != req->inputs.end()) it->second = std::move(a); else req->inputs.emplace_back(std::move(k), std::move(a));   // map: last entry wins
    }
    else if (f == 3 && wt == 2) { Reader s = r.Sub(); req->output_filter.emplace_back(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p)); }
    else r.Skip(wt);
  }
  return r.ok;
}

inline bool ParseResponse(const void* data, size_t n, Response* resp) {
  Reader r(data, n);
  while (!r.Done()) {
    uint64_t tag = r.Varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
    if (f == 1 && wt == 2) { std::string k; Array a; if (!ParseMapEntry(r.Sub(), &k, &a)) return false; resp->outputs.emplace_back(std::move(k), std::move(a)); }
    else r.Skip(wt);
  }
  return r.ok;
}

// ---------------------------------------------------------------- writer
inline void PutVarint(std::string* o, uint64_t v) { while (v >= 0x80) { o->push_back((char)(v | 0x80)); v >>= 7; } o->push_back((char)v); }
inline void PutTag(std::string* o, int field, int wt) { PutVarint(o, ((uint64_t)field << 3) | (uint64_t)wt); }
inline void PutBytes(std::string* o, int field, const void* d, size_t n) { PutTag(o, field, 2); PutVarint(o, n); o->append(static_cast<const char*>(d), n); }

inline void EncodeArray(const Array& a, std::string* o) {
  if (a.dtype) { PutTag(o, 1, 0); PutVarint(o, (uint64_t)a.dtype); }
  {
    std::string dims, shape;
    for (int64_t d : a.shape) PutVarint(&dims, (uint64_t)d);
    if (!dims.empty()) PutBytes(&shape, 1, dims.data(), dims.size());
    PutBytes(o, 2, shape.data(), shape.size());
  }
  if (!a.f32.empty()) PutBytes(o, 3, a.f32.data(), a.f32.size() * 4);
  if (!a.f64.empty()) PutBytes(o, 4, a.f64.data(), a.f64.size() * 8);
  if (!a.i32.empty()) { std::string s; for (int32_t v : a.i32) PutVarint(&s, (uint64_t)(int64_t)v); PutBytes(o, 5, s.data(), s.size()); }
  for (const auto& s : a.str) PutBytes(o, 6, s.data(), s.size());
  if (!a.i64.empty()) { std::string s; s.reserve(a.i64.size() * 4); for (int64_t v : a.i64) PutVarint(&s, (uint64_t)v); PutBytes(o, 7, s.data(), s.size()); }
  if (!a.b8.empty()) { std::string s; for (uint8_t v : a.b8) s.push_back((char)(v ? 1 : 0)); PutBytes(o, 8, s.data(), s.size()); }
}

inline void EncodeMapEntry(int field, const std::string& key, const Array& a, std::string* o) {
  std::string val, entry;
  EncodeArray(a, &val);
  PutBytes(&entry, 1, key.data(), key.size());
  PutBytes(&entry, 2, val.data(), val.size());
  PutBytes(o, field, entry.data(), entry.size());
}

inline void EncodeRequest(const Request& r, std::string* o) {
  if (!r.signature_name.empty()) PutBytes(o, 1, r.signature_name.data(), r.signature_name.size());
  for (const auto& kv : r.inputs) EncodeMapEntry(2, kv.first, kv.second, o);
  for (const auto& f : r.output_filter) PutBytes(o, 3, f.data(), f.size());
}

inline void EncodeResponse(const Response& r, std::string* o) { for (const auto& kv : r.outputs) EncodeMapEntry(1, kv.first, kv.second, o); }

// ---------------------------------------------------------------- PredictRequest <-> the runtime's compact wire format
#pragma pack(push, 1)
struct WireReq { uint32_t magic, version, batch, num_dense, num_sparse, reserved; };
struct WireResp { uint32_t magic, batch, status, reserved; int64_t model_version; };
#pragma pack(pop)
constexpr uint32_t kWireReqMagic = 0x51525244, kWireRespMagic = 0x53525244;

inline bool IsWireRequest(const void* d, size_t n) { uint32_t m = 0; if (n >= 4) memcpy(&m, d, 4); return m == kWireReqMagic; }

// numeric suffix ordering: "C2" < "C10"; names without digits sort lexicographically before
inline bool NaturalLess(const std::string& a, const std::string& b) {
  auto split = [](const std::string& s, std::string* head, long long* num) {
    size_t i = s.size(); while (i > 0 && isdigit((unsigned char)s[i - 1])) --i;
    *head = s.substr(0, i); *num = i < s.size() ? atoll(s.c_str() + i) : -1;
  };
  std::string ha, hb; long long na, nb; split(a, &ha, &na); split(b, &hb, &nb);
  return ha != hb ? ha < hb : na < nb;
}

template <class T> inline void AppendAs(const Array& a, std::vector<T>* out) {
  switch (a.dtype) {
    case DT_FLOAT: for (float v : a.f32) out->push_back((T)v); break;
    case DT_DOUBLE: for (double v : a.f64) out->push_back((T)v); break;
    case DT_INT64: for (int64_t v : a.i64) out->push_back((T)v); break;
    case DT_BOOL: for (uint8_t v : a.b8) out->push_back((T)v); break;
    default: for (int32_t v : a.i32) out->push_back((T)v);       // INT32 / INT16 / INT8 / UINT8 travel in int_val
  }
}

inline bool IsFloatType(int dt) { return dt == DT_FLOAT || dt == DT_DOUBLE; }

// Input conventions (both produced by the client SDK in serving/):
//   (a) "dense": float [B, num_dense], "ids": int64 [num_sparse, B] (feature-major) or [B, num_sparse] when the shape says so;
//   (b) one input per feature, modelzoo naming: float inputs (I1..I13, each [B] or [B,1]) form the dense columns in natural
//       order, integer inputs (C1..C26, each [B] or [B,1]) the sparse features in natural order.
inline bool RequestToWire(const Request& r, int num_dense, int num_sparse, std::string* out, std::string* err) {
  const Array *dense = nullptr, *ids = nullptr;
  for (const auto& kv : r.inputs) { if (kv.first == "dense") dense = &kv.second; else if (kv.first == "ids") ids = &kv.second; }
  std::vector<float> D; std::vector<int64_t> I; int64_t B = 0;
  if (dense && ids) {
    AppendAs(*dense, &D); AppendAs(*ids, &I);
    if (num_dense <= 0 || D.size() % (size_t)num_dense) { *err = "dense element count is not a multiple of num_dense"; return false; }
    B = (int64_t)D.size() / num_dense;
    if ((int64_t)I.size() != B * num_sparse) { *err = "ids element count != batch * num_sparse"; return false; }
    const bool sample_major = ids->shape.size() == 2 && ids->shape[0] == B && ids->shape[1] == num_sparse && B != num_sparse;
    if (sample_major) { std::vector<int64_t> T(I.size()); for (int64_t b = 0; b < B; ++b) for (int t = 0; t < num_sparse; ++t) T[(size_t)t * B + b] = I[(size_t)b * num_sparse + t]; I.swap(T); }
  } else {
    std::vector<const std::pair<std::string, Array>*> fcols, icols;
    for (const auto& kv : r.inputs) (IsFloatType(kv.second.dtype) ? fcols : icols).push_back(&kv);
    auto by_name = [](const auto* a, const auto* b) { return NaturalLess(a->first, b->first); };
    std::sort(fcols.begin(), fcols.end(), by_name); std::sort(icols.begin(), icols.end(), by_name);
    if ((int)fcols.size() != num_dense || (int)icols.size() != num_sparse) { *err = "expected " + std::to_string(num_dense) + " float and " + std::to_string(num_sparse) + " integer inputs, got " + std::to_string(fcols.size()) + " and " + std::to_string(icols.size()); return false; }
    std::vector<std::vector<float>> cols(fcols.size());
    for (size_t c = 0; c < fcols.size(); ++c) AppendAs(fcols[c]->second, &cols[c]);
    B = cols.empty() ? -1 : (int64_t)cols[0].size();
    for (const auto* kv : icols) { size_t o = I.size(); AppendAs(kv->second, &I); int64_t n = (int64_t)(I.size() - o); if (B < 0) B = n; if (n != B) { *err = "input " + kv->first + " has a different batch size"; return false; } }
    for (const auto& c : cols) if ((int64_t)c.size() != B) { *err = "dense inputs have different batch sizes"; return false; }
    D.resize((size_t)B * num_dense);
    for (int64_t b = 0; b < B; ++b) for (int c = 0; c < num_dense; ++c) D[(size_t)b * num_dense + c] = cols[c][b];
  }
  if (B <= 0) { *err = "empty batch"; return false; }
  WireReq h{kWireReqMagic, 1, (uint32_t)B, (uint32_t)num_dense, (uint32_t)num_sparse, 0};
  out->assign(reinterpret_cast<const char*>(&h), sizeof(h));
  out->append(reinterpret_cast<const char*>(D.data()), D.size() * 4);
  out->append(reinterpret_cast<const char*>(I.data()), I.size() * 8);
  return true;
}

// DRRS -> PredictResponse{"probabilities": float[B], "model_version": int64[1]} honouring the request's output_filter.
inline bool WireToResponse(const void* wire, size_t n, const std::vector<std::string>& output_filter, std::string* out) {
  WireResp h;
  if (n < sizeof(h)) return false;
  memcpy(&h, wire, sizeof(h));
  const size_t no = h.reserved > 1 ? h.reserved : 1;                 // multi-task models: `reserved` = probabilities per row (sample-major)
  if (h.magic != kWireRespMagic || n < sizeof(h) + (size_t)h.batch * no * 4) return false;
  auto wanted = [&](const char* name) { return output_filter.empty() || std::find(output_filter.begin(), output_filter.end(), name) != output_filter.end(); };
  Response r;
  if (wanted("probabilities")) {
    Array a; a.dtype = DT_FLOAT; a.f32.resize((size_t)h.batch * no);
    if (no > 1) a.shape = {(int64_t)h.batch, (int64_t)no}; else a.shape = {(int64_t)h.batch};
    if (h.batch) memcpy(a.f32.data(), static_cast<const uint8_t*>(wire) + sizeof(h), (size_t)h.batch * no * 4);
    r.outputs.emplace_back("probabilities", std::move(a));
  }
  if (wanted("model_version")) { Array a; a.dtype = DT_INT64; a.shape = {1}; a.i64 = {h.model_version}; r.outputs.emplace_back("model_version", std::move(a)); }
  EncodeResponse(r, out);
  return true;
}

}  // namespace drpb
