// CSV record decoding for the input pipeline: id lists, key:value lists and plain number rows -> sparse (row, id[, value]) triples or dense rows.
// Records arrive as ONE packed byte buffer + offsets (no per-record Python objects); counting and filling are two parallel passes over the
// records, so a 8192-record batch of multi-hot features is decoded in tens of microseconds per core instead of a Python loop per token.
//
// Reference: kernels/trans_csv_ali_ops.cc (TransCsvID2Sparse / ID2Dense / KV2Sparse / KV2Dense / ToDense, sharded over the intra-op pool with
// work_sharder), kernels/string_split_and_pad_ali_op.cc.  Semantics kept: the second dimension of the result is `max_id` (ids / keys must be
// < max_id; max_id = -1048575 in the reference = "detect from the data"), empty tokens are skipped, malformed numbers are an error.
#include <cstdint>
#include <cstdlib>
#include <cstring>

extern "C" {

// tokens per record (empty tokens skipped); returns the total
int64_t dr_csv_count(const char* buf, const int64_t* offs, int64_t n, char delim, int64_t* counts) {
  int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int64_t r = 0; r < n; ++r) {
    const char* p = buf + offs[r]; const char* e = buf + offs[r + 1];
    int64_t c = 0; bool in_tok = false;
    for (; p < e; ++p) {
      if (*p == delim) { in_tok = false; continue; }
      if (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') continue;
      if (!in_tok) { in_tok = true; ++c; }
    }
    counts[r] = c; total += c;
  }
  return total;
}

static inline const char* skip_ws(const char* p, const char* e) { while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p; return p; }

// one token [p, q): parse an int64; returns false on garbage
static inline bool parse_i64(const char* p, const char* q, int64_t* out) {
  p = skip_ws(p, q);
  bool neg = false;
  if (p < q && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
  if (p >= q) return false;
  int64_t v = 0; bool any = false;
  for (; p < q && *p >= '0' && *p <= '9'; ++p) { v = v * 10 + (*p - '0'); any = true; }
  p = skip_ws(p, q);
  if (!any || p != q) return false;
  *out = neg ? -v : v;
  return true;
}
static inline bool parse_f32(const char* p, const char* q, float* out) {
  p = skip_ws(p, q);
  char tmp[64];
  const size_t len = (size_t)(q - p);
  if (len == 0 || len >= sizeof(tmp)) return false;
  memcpy(tmp, p, len); tmp[len] = 0;
  char* end = nullptr;
  const float v = strtof(tmp, &end);
  while (*end == ' ' || *end == '\t' || *end == '\r' || *end == '\n') ++end;
  if (end == tmp || *end != 0) return false;
  *out = v;
  return true;
}

// ids of every record -> (row, id) pairs at row_start[r]...; returns the number of malformed tokens (0 = ok).  Ids outside [0, max_id) (max_id >= 0)
// are written as -1 (the caller drops them).
int64_t dr_csv_ids(const char* buf, const int64_t* offs, int64_t n, char delim, int64_t max_id, const int64_t* row_start, int64_t* out_rows, int64_t* out_ids) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t r = 0; r < n; ++r) {
    const char* p = buf + offs[r]; const char* e = buf + offs[r + 1];
    int64_t w = row_start[r];
    while (p <= e) {
      const char* q = (const char*)memchr(p, delim, (size_t)(e - p));
      if (!q) q = e;
      if (skip_ws(p, q) != q) {
        int64_t v;
        if (!parse_i64(p, q, &v)) ++bad;
        else { out_rows[w] = r; out_ids[w] = (max_id >= 0 && (v < 0 || v >= max_id)) ? -1 : v; ++w; }
      }
      p = q + 1;
    }
  }
  return bad;
}

// "key:value" tokens -> (row, key, value)
int64_t dr_csv_kvs(const char* buf, const int64_t* offs, int64_t n, char delim, char kv_delim, int64_t max_id, const int64_t* row_start, int64_t* out_rows,
                   int64_t* out_keys, float* out_vals) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t r = 0; r < n; ++r) {
    const char* p = buf + offs[r]; const char* e = buf + offs[r + 1];
    int64_t w = row_start[r];
    while (p <= e) {
      const char* q = (const char*)memchr(p, delim, (size_t)(e - p));
      if (!q) q = e;
      if (skip_ws(p, q) != q) {
        const char* c = (const char*)memchr(p, kv_delim, (size_t)(q - p));
        int64_t k; float v;
        if (!c || !parse_i64(p, c, &k) || !parse_f32(c + 1, q, &v)) ++bad;
        else { out_rows[w] = r; out_keys[w] = (max_id >= 0 && (k < 0 || k >= max_id)) ? -1 : k; out_vals[w] = v; ++w; }
      }
      p = q + 1;
    }
  }
  return bad;
}

// plain numbers -> dense rows [n, ncols], left-aligned, zero-padded, extra columns dropped
int64_t dr_csv_to_dense(const char* buf, const int64_t* offs, int64_t n, char delim, int64_t ncols, float* out) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t r = 0; r < n; ++r) {
    const char* p = buf + offs[r]; const char* e = buf + offs[r + 1];
    float* row = out + r * ncols;
    for (int64_t c = 0; c < ncols; ++c) row[c] = 0.f;
    int64_t c = 0;
    while (p <= e) {
      const char* q = (const char*)memchr(p, delim, (size_t)(e - p));
      if (!q) q = e;
      if (skip_ws(p, q) != q) {
        float v;
        if (!parse_f32(p, q, &v)) ++bad;
        else { if (c < ncols) row[c] = v; ++c; }
      }
      p = q + 1;
    }
  }
  return bad;
}

// StringSplitAndPad on ids: every record split into at most max_len int64 tokens, padded with pad_value -> [n, max_len]
int64_t dr_csv_split_pad_ids(const char* buf, const int64_t* offs, int64_t n, char delim, int64_t max_len, int64_t pad_value, int64_t* out) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t r = 0; r < n; ++r) {
    const char* p = buf + offs[r]; const char* e = buf + offs[r + 1];
    int64_t* row = out + r * max_len;
    int64_t c = 0;
    while (p <= e && c < max_len) {
      const char* q = (const char*)memchr(p, delim, (size_t)(e - p));
      if (!q) q = e;
      if (skip_ws(p, q) != q) {
        int64_t v;
        if (!parse_i64(p, q, &v)) ++bad; else row[c++] = v;
      }
      p = q + 1;
    }
    for (; c < max_len; ++c) row[c] = pad_value;
  }
  return bad;
}

}  // extern "C"
