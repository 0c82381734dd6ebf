// Remote feature store client: RESP2 (Redis serialization protocol) over TCP, written against the protocol (no hiredis).
//
// Reference behaviour (serving/processor/storage/redis_feature_store.{h,cc}, feature_store_mgr.h): with
// `feature_store_type: "redis"` the serving graph's EV gathers become KvLookup ops and model updates become KvInsert /
// KvImport ops against a Redis instance shared by all serving replicas.  Here one row is one Redis string:
//   key   = "<prefix>:<feature key, decimal>"        (prefix = "<model>/<version>/<table>")
//   value = dim x fp32, little endian
// Batched operations are pipelined (one write of many MGET / MSET commands, then one pass over the replies), so a
// lookup of N keys costs ceil(N / kChunk) commands in a single round trip.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr int64_t kChunk = 512;          // keys per MGET / MSET command

struct Reply {
  char type = 0;                         // '+', '-', ':', '$', '*', or 'n' (nil)
  std::string str;
  int64_t num = 0;
  std::vector<Reply> items;
};

struct RedisConn {
  int fd = -1;
  int timeout_ms = 5000;
  std::mutex mu;
  std::string err;
  std::string rbuf;                      // unread bytes
  size_t rpos = 0;

  ~RedisConn() { if (fd >= 0) close(fd); }

  bool Connect(const char* host, int port) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC; hints.ai_socktype = SOCK_STREAM;
    char portstr[16]; snprintf(portstr, sizeof(portstr), "%d", port);
    if (getaddrinfo(host, portstr, &hints, &res) != 0 || !res) { err = std::string("cannot resolve ") + host; return false; }
    for (addrinfo* a = res; a; a = a->ai_next) {
      fd = socket(a->ai_family, a->ai_socktype, a->ai_protocol);
      if (fd < 0) continue;
      if (connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
      close(fd); fd = -1;
    }
    freeaddrinfo(res);
    if (fd < 0) { err = std::string("cannot connect to ") + host + ":" + portstr + ": " + strerror(errno); return false; }
    int one = 1; setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    return true;
  }

  bool WriteAll(const std::string& s) {
    size_t off = 0;
    while (off < s.size()) {
      ssize_t w = send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
      if (w < 0) { if (errno == EINTR) continue; err = std::string("send: ") + strerror(errno); return false; }
      off += (size_t)w;
    }
    return true;
  }

  bool Fill() {
    pollfd p{fd, POLLIN, 0};
    int r = poll(&p, 1, timeout_ms);
    if (r == 0) { err = "timeout waiting for the server"; return false; }
    if (r < 0) { if (errno == EINTR) return Fill(); err = std::string("poll: ") + strerror(errno); return false; }
    char tmp[1 << 16];
    ssize_t n = recv(fd, tmp, sizeof(tmp), 0);
    if (n == 0) { err = "connection closed by the server"; return false; }
    if (n < 0) { if (errno == EINTR) return Fill(); err = std::string("recv: ") + strerror(errno); return false; }
    if (rpos > (1u << 20) && rpos * 2 > rbuf.size()) { rbuf.erase(0, rpos); rpos = 0; }
    rbuf.append(tmp, (size_t)n);
    return true;
  }

  bool ReadLine(std::string* line) {
    for (;;) {
      size_t e = rbuf.find("\r\n", rpos);
      if (e != std::string::npos) { line->assign(rbuf, rpos, e - rpos); rpos = e + 2; return true; }
      if (!Fill()) return false;
    }
  }

  bool ReadExact(size_t n, std::string* out) {
    while (rbuf.size() - rpos < n + 2) if (!Fill()) return false;
    out->assign(rbuf, rpos, n); rpos += n + 2;
    return true;
  }

  bool ReadReply(Reply* r, int depth = 0) {
    std::string line;
    if (!ReadLine(&line) || line.empty()) { if (err.empty()) err = "protocol error: empty reply line"; return false; }
    r->type = line[0];
    switch (line[0]) {
      case '+': case '-': r->str = line.substr(1); return true;
      case ':': r->num = atoll(line.c_str() + 1); return true;
      case '$': {
        long long n = atoll(line.c_str() + 1);
        if (n < 0) { r->type = 'n'; return true; }
        return ReadExact((size_t)n, &r->str);
      }
      case '*': {
        long long n = atoll(line.c_str() + 1);
        if (n < 0) { r->type = 'n'; return true; }
        if (depth > 4) { err = "protocol error: reply nested too deep"; return false; }
        r->items.resize((size_t)n);
        for (auto& it : r->items) if (!ReadReply(&it, depth + 1)) return false;
        return true;
      }
      default: err = "protocol error: unknown reply type '" + line.substr(0, 1) + "'"; return false;
    }
  }
};

void AppendBulk(std::string* o, const char* d, size_t n) {
  char h[32]; int k = snprintf(h, sizeof(h), "$%zu\r\n", n);
  o->append(h, (size_t)k); o->append(d, n); o->append("\r\n");
}
void AppendBulk(std::string* o, const std::string& s) { AppendBulk(o, s.data(), s.size()); }
void AppendHeader(std::string* o, size_t argc) { char h[32]; int k = snprintf(h, sizeof(h), "*%zu\r\n", argc); o->append(h, (size_t)k); }

std::string RowKey(const char* prefix, int64_t key) {
  char b[32]; snprintf(b, sizeof(b), ":%lld", (long long)key);
  return std::string(prefix) + b;
}

bool Simple(RedisConn* c, const std::vector<std::string>& args, Reply* r) {
  std::string cmd; AppendHeader(&cmd, args.size());
  for (const auto& a : args) AppendBulk(&cmd, a);
  if (!c->WriteAll(cmd) || !c->ReadReply(r)) return false;
  if (r->type == '-') { c->err = r->str; return false; }
  return true;
}

}  // namespace

extern "C" {

void* dr_redis_connect(const char* host, int port, int timeout_ms, const char* password, int db) {
  auto* c = new RedisConn();
  c->timeout_ms = timeout_ms > 0 ? timeout_ms : 5000;
  if (!c->Connect(host, port)) return c;          // caller checks dr_redis_ok()
  Reply r;
  if (password && *password && !Simple(c, {"AUTH", password}, &r)) { close(c->fd); c->fd = -1; return c; }
  if (db > 0 && !Simple(c, {"SELECT", std::to_string(db)}, &r)) { close(c->fd); c->fd = -1; return c; }
  return c;
}

int dr_redis_ok(void* h) { return h && static_cast<RedisConn*>(h)->fd >= 0; }
const char* dr_redis_last_error(void* h) { return h ? static_cast<RedisConn*>(h)->err.c_str() : "null connection"; }
void dr_redis_close(void* h) { delete static_cast<RedisConn*>(h); }

int dr_redis_ping(void* h) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  Reply r; return Simple(c, {"PING"}, &r) && r.str == "PONG" ? 0 : -1;
}

int64_t dr_redis_dbsize(void* h) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  Reply r; return Simple(c, {"DBSIZE"}, &r) ? r.num : -1;
}

int dr_redis_flushdb(void* h) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  Reply r; return Simple(c, {"FLUSHDB"}, &r) ? 0 : -1;
}

int dr_redis_set(void* h, const char* key, const void* val, int64_t n) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  Reply r; return Simple(c, {"SET", key, std::string(static_cast<const char*>(val), (size_t)n)}, &r) ? 0 : -1;
}

// returns the value length (copied up to cap), -1 when missing, -2 on error
int64_t dr_redis_get(void* h, const char* key, void* out, int64_t cap) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  Reply r;
  if (!Simple(c, {"GET", key}, &r)) return -2;
  if (r.type == 'n') return -1;
  memcpy(out, r.str.data(), (size_t)std::min<int64_t>(cap, (int64_t)r.str.size()));
  return (int64_t)r.str.size();
}

// rows [n, dim] fp32 -> pipelined MSET commands.  Returns n, or -1 on error.
int64_t dr_redis_mset_rows(void* h, const char* prefix, const int64_t* keys, int64_t n, const float* rows, int dim) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  std::string cmd; int64_t ncmd = 0;
  for (int64_t off = 0; off < n; off += kChunk, ++ncmd) {
    const int64_t m = std::min(kChunk, n - off);
    AppendHeader(&cmd, (size_t)(1 + 2 * m)); AppendBulk(&cmd, "MSET", 4);
    for (int64_t i = 0; i < m; ++i) {
      AppendBulk(&cmd, RowKey(prefix, keys[off + i]));
      AppendBulk(&cmd, reinterpret_cast<const char*>(rows + (off + i) * dim), (size_t)dim * 4);
    }
  }
  if (!c->WriteAll(cmd)) return -1;
  bool ok = true;
  for (int64_t i = 0; i < ncmd; ++i) { Reply r; if (!c->ReadReply(&r)) return -1; if (r.type == '-') { c->err = r.str; ok = false; } }
  return ok ? n : -1;
}

// pipelined MGET: found rows are written to rows[i], found[i] = 1; missing rows are left untouched.  Returns #found or -1.
int64_t dr_redis_mget_rows(void* h, const char* prefix, const int64_t* keys, int64_t n, float* rows, int dim, uint8_t* found) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  std::string cmd; int64_t ncmd = 0;
  for (int64_t off = 0; off < n; off += kChunk, ++ncmd) {
    const int64_t m = std::min(kChunk, n - off);
    AppendHeader(&cmd, (size_t)(1 + m)); AppendBulk(&cmd, "MGET", 4);
    for (int64_t i = 0; i < m; ++i) AppendBulk(&cmd, RowKey(prefix, keys[off + i]));
  }
  if (!c->WriteAll(cmd)) return -1;
  int64_t hits = 0; bool ok = true;
  for (int64_t ci = 0; ci < ncmd; ++ci) {
    Reply r;
    if (!c->ReadReply(&r)) return -1;
    const int64_t off = ci * kChunk, m = std::min(kChunk, n - off);
    if (r.type != '*' || (int64_t)r.items.size() != m) { c->err = r.type == '-' ? r.str : "protocol error: MGET reply shape"; ok = false; continue; }
    for (int64_t i = 0; i < m; ++i) {
      const Reply& it = r.items[(size_t)i];
      const bool hit = it.type == '$' && it.str.size() == (size_t)dim * 4;
      if (found) found[off + i] = hit ? 1 : 0;
      if (hit) { memcpy(rows + (off + i) * dim, it.str.data(), (size_t)dim * 4); ++hits; }
    }
  }
  return ok ? hits : -1;
}

int64_t dr_redis_del_rows(void* h, const char* prefix, const int64_t* keys, int64_t n) {
  auto* c = static_cast<RedisConn*>(h); std::lock_guard<std::mutex> l(c->mu);
  std::string cmd; int64_t ncmd = 0;
  for (int64_t off = 0; off < n; off += kChunk, ++ncmd) {
    const int64_t m = std::min(kChunk, n - off);
    AppendHeader(&cmd, (size_t)(1 + m)); AppendBulk(&cmd, "DEL", 3);
    for (int64_t i = 0; i < m; ++i) AppendBulk(&cmd, RowKey(prefix, keys[off + i]));
  }
  if (!c->WriteAll(cmd)) return -1;
  int64_t removed = 0;
  for (int64_t i = 0; i < ncmd; ++i) { Reply r; if (!c->ReadReply(&r)) return -1; if (r.type == ':') removed += r.num; }
  return removed;
}

}  // extern "C"
