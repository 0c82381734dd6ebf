// Native data plane of the asynchronous parameter-server mode: sparse pulls / pushes over plain TCP, served by C++ threads straight on the
// HostEV engine -- no Python, no GIL, no pickling on the hot path.
//
// Reference components replaced: StarServer (PS runtime with pull / push semantics and lock-free execution on the PS,
// contrib/star_server/*, kernels/star_run_graph_op.cc:157) and the GRPC++ / seastar tensor transport (contrib/star/seastar/*: "control plane
// stays gRPC, tensors travel over the fast path").  Same split here: torch.distributed.rpc stays the control plane (create variables,
// elastic scaling, checkpoints -- parallel/ps.py), this file is the tensor path:
//
//   * one acceptor thread + one thread per worker connection (a job has tens of workers, not thousands: share-nothing per connection, blocking
//     reads, TCP_NODELAY; every connection owns its receive / send buffers)
//   * PULL  = FuseRecv: ONE message carries the keys of every table of the step; rows come back in request order
//   * PUSH  = sparse gradients of every table in ONE message; the server dedups + applies through dr_host_ev_apply_raw; the optimizer
//     hyper-state (global step, beta powers) advances under a per-table mutex; rows sit behind a per-table reader / writer lock (pulls
//     shared, a push exclusive for its apply; tables independent) -- DEEPREC_PS_HOGWILD=1 drops it for the reference's lock-free execution
//   * elastic scaling fence: every request carries the server-definition version it was partitioned under; a frozen server or a version
//     mismatch answers STALE without touching a row; in-flight requests are counted so IsReadyScaling can wait for a drained server
//
// Wire format (little endian).  Request: u32 magic 'DRPS', u32 op (1 pull | 2 push | 3 ping), i32 def_version (-1: do not check),
// u32 n_tables, then per table: u32 table_id, u32 n, i64 keys[n] (+ f32 grads[n * dim] for a push).
// Response: u32 magic, u32 status (0 ok | 1 stale | 2 error), u64 aux (pushes applied so far); a pull appends f32 rows[n * dim] per table.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "../common/ev_types.h"

extern "C" {
void dr_host_ev_lookup(void* h, const int64_t* keys, int64_t n, float* out);
void dr_host_ev_apply_raw(void* h, const int64_t* ids, int64_t n, const float* grads, int64_t row_stride, const DrOptHyper* hp);
}

namespace {

constexpr uint32_t kMagic = 0x53505244;          // "DRPS"
enum : uint32_t { OP_PULL = 1, OP_PUSH = 2, OP_PING = 3 };
enum : uint32_t { ST_OK = 0, ST_STALE = 1, ST_ERROR = 2 };
constexpr uint32_t kMaxTablesPerMsg = 4096;
constexpr uint64_t kMaxKeysPerTable = 1ull << 28;

bool ReadAll(int fd, void* buf, size_t n) {
  uint8_t* p = static_cast<uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
    if (r <= 0) return false;
    p += r; n -= (size_t)r;
  }
  return true;
}
bool WriteAll(int fd, const void* buf, size_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r <= 0) return false;
    p += r; n -= (size_t)r;
  }
  return true;
}

struct PsTable {
  void* ev = nullptr; int dim = 0;
  DrOptHyper hp{}; std::mutex mu;                 // optimizer hyper-state: step counter / beta powers advance once per push
  // rows: pulls of a table run concurrently (shared), a push excludes them for the duration of its apply (exclusive) -- a row is never
  // read half-updated and two pushes never interleave on a row; tables are independent.  DEEPREC_PS_HOGWILD=1 drops this lock: the
  // reference's lock-free PS execution (updates may be lost / torn when two workers hit the same key at once).
  std::shared_mutex rows;
};
const bool kHogwild = [] { const char* e = getenv("DEEPREC_PS_HOGWILD"); return e && e[0] == '1'; }();

struct PsServer {
  int listen_fd = -1, port = 0;
  std::thread acceptor;
  std::mutex mu;                                    // tables (append-only), connection list
  std::vector<std::unique_ptr<PsTable>> tables;
  std::vector<int> conns; std::vector<std::thread> workers;
  std::atomic<bool> stop{false};
  std::atomic<int> frozen{0}, def_version{0}, inflight{0};
  std::atomic<uint64_t> pulls{0}, pushes{0}, bytes_in{0}, bytes_out{0}, stale{0};

  PsTable* Table(uint32_t id) { std::lock_guard<std::mutex> l(mu); return id < tables.size() ? tables[id].get() : nullptr; }

  // admission of one request under the scaling fence: counted in `inflight` BEFORE the staleness check can pass, so a drained server stays drained
  bool Admit(int32_t dv) {
    // Dekker-style handshake with dr_ps_server_set_def / dr_ps_server_inflight (store frozen; load inflight): all four accesses are seq_cst
    inflight.fetch_add(1);
    if (frozen.load() || (dv >= 0 && dv != def_version.load())) {
      inflight.fetch_sub(1);
      stale.fetch_add(1, std::memory_order_relaxed);
      return false;
    }
    return true;
  }

  void Serve(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    std::vector<uint8_t> in, out;
    struct Item { PsTable* t; uint32_t n; size_t keys_off, grads_off; };
    std::vector<Item> items;
    for (;;) {
      uint32_t head[4];
      if (!ReadAll(fd, head, sizeof(head))) break;
      if (head[0] != kMagic || head[3] > kMaxTablesPerMsg) break;
      const uint32_t op = head[1], nt = head[3];
      int32_t dv; memcpy(&dv, &head[2], 4);
      // the payload is read completely before anything is decided: a rejected request leaves the stream in sync
      in.clear(); items.clear();
      bool ok = true;
      uint64_t total_rows_bytes = 0;
      for (uint32_t i = 0; i < nt && ok; ++i) {
        uint32_t th[2];
        if (!ReadAll(fd, th, sizeof(th))) { ok = false; break; }
        PsTable* t = Table(th[0]);
        const uint64_t n = th[1];
        if (!t || n > kMaxKeysPerTable) { ok = false; break; }
        const size_t kbytes = (size_t)n * 8, gbytes = op == OP_PUSH ? (size_t)n * t->dim * 4 : 0;
        const size_t off = in.size();
        in.resize(off + kbytes + gbytes);
        if (!ReadAll(fd, in.data() + off, kbytes + gbytes)) { ok = false; break; }
        items.push_back({t, (uint32_t)n, off, off + kbytes});
        total_rows_bytes += (uint64_t)n * t->dim * 4;
      }
      if (!ok) break;                                  // malformed stream or closed socket: drop the connection
      bytes_in.fetch_add(sizeof(head) + in.size() + (uint64_t)nt * 8, std::memory_order_relaxed);
      uint32_t status = ST_OK;
      size_t body = 0;
      if (op == OP_PING) {
      } else if (!Admit(dv)) {
        status = ST_STALE;
      } else {
        if (op == OP_PULL) {
          out.resize(16 + (size_t)total_rows_bytes);
          size_t o = 16;
          for (const Item& it : items) {
            // keys sit at an arbitrary offset inside the receive buffer: 8-byte alignment is not guaranteed -> aligned copy
            std::vector<int64_t> keys(it.n);
            if (it.n) memcpy(keys.data(), in.data() + it.keys_off, (size_t)it.n * 8);
            if (it.n) {
              std::shared_lock<std::shared_mutex> rl(it.t->rows, std::defer_lock);
              if (!kHogwild) rl.lock();
              dr_host_ev_lookup(it.t->ev, keys.data(), it.n, reinterpret_cast<float*>(out.data() + o));
            }
            o += (size_t)it.n * it.t->dim * 4;
          }
          body = o - 16;
          pulls.fetch_add(1, std::memory_order_relaxed);
        } else if (op == OP_PUSH) {
          for (const Item& it : items) {
            if (!it.n) continue;
            DrOptHyper hp;
            {
              std::lock_guard<std::mutex> l(it.t->mu);
              hp = it.t->hp;
              it.t->hp.global_step += 1;
              if (hp.kind == DR_OPT_ADAM || hp.kind == DR_OPT_ADAMW || hp.kind == DR_OPT_ADAM_ASYNC || hp.kind == DR_OPT_ADAM_ASYNC_RMSPROP) { it.t->hp.beta1_power *= hp.beta1; it.t->hp.beta2_power *= hp.beta2; }
            }
            std::vector<int64_t> keys(it.n); std::vector<float> grads((size_t)it.n * it.t->dim);
            memcpy(keys.data(), in.data() + it.keys_off, (size_t)it.n * 8);
            memcpy(grads.data(), in.data() + it.grads_off, grads.size() * 4);
            {
              std::unique_lock<std::shared_mutex> wl(it.t->rows, std::defer_lock);
              if (!kHogwild) wl.lock();
              dr_host_ev_apply_raw(it.t->ev, keys.data(), it.n, grads.data(), it.t->dim, &hp);     // dedup + segment-sum + row-wise optimizer
            }
            pushes.fetch_add(1, std::memory_order_relaxed);
          }
        } else {
          status = ST_ERROR;
        }
        inflight.fetch_sub(1);
      }
      if (out.size() < 16) out.resize(16);
      const uint64_t aux = pushes.load(std::memory_order_relaxed);
      memcpy(out.data(), &kMagic, 4); memcpy(out.data() + 4, &status, 4); memcpy(out.data() + 8, &aux, 8);
      if (!WriteAll(fd, out.data(), 16 + (status == ST_OK ? body : 0))) break;
      bytes_out.fetch_add(16 + (status == ST_OK ? body : 0), std::memory_order_relaxed);
    }
    ::close(fd);
  }

  void Accept() {
    while (!stop.load()) {
      sockaddr_in peer{}; socklen_t len = sizeof(peer);
      const int fd = ::accept(listen_fd, reinterpret_cast<sockaddr*>(&peer), &len);
      if (fd < 0) { if (stop.load()) return; continue; }
      std::lock_guard<std::mutex> l(mu);
      if (stop.load()) { ::close(fd); return; }
      conns.push_back(fd);
      workers.emplace_back([this, fd] { Serve(fd); });
    }
  }
};

struct PsClient {
  int fd = -1;
  std::vector<uint8_t> buf;
};

}  // namespace

extern "C" {

// port 0 = ephemeral; *bound_port receives the port the server listens on.  Returns nullptr on failure.
void* dr_ps_server_start(const char* bind_addr, int port, int* bound_port) {
  auto s = std::make_unique<PsServer>();
  s->listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
  if (s->listen_fd < 0) return nullptr;
  int one = 1;
  setsockopt(s->listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in a{}; a.sin_family = AF_INET; a.sin_port = htons((uint16_t)port);
  if (!bind_addr || !*bind_addr || inet_pton(AF_INET, bind_addr, &a.sin_addr) != 1) a.sin_addr.s_addr = htonl(INADDR_ANY);
  if (::bind(s->listen_fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0 || ::listen(s->listen_fd, 128) != 0) { ::close(s->listen_fd); return nullptr; }
  socklen_t len = sizeof(a);
  getsockname(s->listen_fd, reinterpret_cast<sockaddr*>(&a), &len);
  s->port = ntohs(a.sin_port);
  if (bound_port) *bound_port = s->port;
  PsServer* raw = s.release();
  raw->acceptor = std::thread([raw] { raw->Accept(); });
  return raw;
}

// Registers a HostEV (handle of dr_host_ev_create) with its optimizer hyper-state; returns the table id workers address it by.
int dr_ps_server_add_table(void* sv, void* host_ev, int dim, const DrOptHyper* hp) {
  auto* s = static_cast<PsServer*>(sv);
  auto t = std::make_unique<PsTable>();
  t->ev = host_ev; t->dim = dim; t->hp = *hp;
  std::lock_guard<std::mutex> l(s->mu);
  s->tables.push_back(std::move(t));
  return (int)s->tables.size() - 1;
}

void dr_ps_server_set_def(void* sv, int def_version, int frozen) {
  auto* s = static_cast<PsServer*>(sv);
  s->def_version.store(def_version);
  s->frozen.store(frozen);
}
int dr_ps_server_inflight(void* sv) { return static_cast<PsServer*>(sv)->inflight.load(); }

// out[6] = pulls, pushes (table-level), bytes in, bytes out, stale rejections, connections
void dr_ps_server_stats(void* sv, uint64_t* out) {
  auto* s = static_cast<PsServer*>(sv);
  out[0] = s->pulls.load(); out[1] = s->pushes.load(); out[2] = s->bytes_in.load(); out[3] = s->bytes_out.load(); out[4] = s->stale.load();
  std::lock_guard<std::mutex> l(s->mu);
  out[5] = s->conns.size();
}

void dr_ps_server_stop(void* sv) {
  auto* s = static_cast<PsServer*>(sv);
  s->stop.store(true);
  ::shutdown(s->listen_fd, SHUT_RDWR); ::close(s->listen_fd);
  if (s->acceptor.joinable()) s->acceptor.join();
  std::vector<std::thread> ws;
  {
    std::lock_guard<std::mutex> l(s->mu);
    for (int fd : s->conns) ::shutdown(fd, SHUT_RDWR);       // wakes the blocking reads; Serve() closes the descriptor
    ws.swap(s->workers);
  }
  for (auto& t : ws) if (t.joinable()) t.join();
  delete s;
}

// Stable partition of a key batch by owning server (key % 1000 % num_ps with a non-negative remainder -- the 1000-bucket rule checkpoints use):
// order[0 .. n) lists the positions grouped by owner (ascending position inside a group), counts[p] = keys owned by server p.
void dr_ps_partition(const int64_t* keys, int64_t n, int num_ps, int64_t* order, int64_t* counts) {
  std::vector<int32_t> own((size_t)n);
  for (int p = 0; p < num_ps; ++p) counts[p] = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t b = keys[i] % 1000; if (b < 0) b += 1000;
    own[(size_t)i] = (int32_t)(b % num_ps);
    counts[own[(size_t)i]]++;
  }
  std::vector<int64_t> next((size_t)num_ps, 0);
  for (int p = 1; p < num_ps; ++p) next[(size_t)p] = next[(size_t)p - 1] + counts[p - 1];
  for (int64_t i = 0; i < n; ++i) order[next[(size_t)own[(size_t)i]]++] = i;
}

// ---- worker side (ctypes releases the GIL around these calls) ----------------------------------------------------------------------
void* dr_ps_client_connect(const char* host, int port) {
  const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return nullptr;
  sockaddr_in a{}; a.sin_family = AF_INET; a.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host, &a.sin_addr) != 1 || ::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) { ::close(fd); return nullptr; }
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  auto* c = new PsClient();
  c->fd = fd;
  return c;
}
void dr_ps_client_close(void* cv) { auto* c = static_cast<PsClient*>(cv); if (c) { ::close(c->fd); delete c; } }

// Sends one PULL (grads == nullptr) or PUSH message for `nt` tables: table_ids[i], n[i] keys at keys[i] (+ grads[i], n[i] * dims[i] floats).
// Returns 0, or -1 when the connection broke.
int dr_ps_client_send(void* cv, int op, int def_version, int nt, const int* table_ids, const int64_t* n, const int64_t* const* keys, const float* const* grads,
                      const int* dims) {
  auto* c = static_cast<PsClient*>(cv);
  size_t total = 16;
  for (int i = 0; i < nt; ++i) total += 8 + (size_t)n[i] * 8 + (grads ? (size_t)n[i] * dims[i] * 4 : 0);
  c->buf.resize(total);
  uint8_t* p = c->buf.data();
  const uint32_t head[4] = {kMagic, (uint32_t)op, 0, (uint32_t)nt};
  memcpy(p, head, 16); memcpy(p + 8, &def_version, 4); p += 16;
  for (int i = 0; i < nt; ++i) {
    const uint32_t th[2] = {(uint32_t)table_ids[i], (uint32_t)n[i]};
    memcpy(p, th, 8); p += 8;
    if (n[i]) { memcpy(p, keys[i], (size_t)n[i] * 8); p += (size_t)n[i] * 8; }
    if (grads && n[i]) { memcpy(p, grads[i], (size_t)n[i] * dims[i] * 4); p += (size_t)n[i] * dims[i] * 4; }
  }
  return WriteAll(c->fd, c->buf.data(), total) ? 0 : -1;
}

// Reads one response.  rows[i] (pull only) receives n[i] * dims[i] floats when the status is OK.  Returns the status (0 ok, 1 stale, 2 error) or -1
// when the connection broke; *aux = pushes the server has applied so far.
int dr_ps_client_recv(void* cv, int nt, const int64_t* n, const int* dims, float* const* rows, uint64_t* aux) {
  auto* c = static_cast<PsClient*>(cv);
  uint8_t head[16];
  if (!ReadAll(c->fd, head, 16)) return -1;
  uint32_t magic, status; memcpy(&magic, head, 4); memcpy(&status, head + 4, 4);
  if (aux) memcpy(aux, head + 8, 8);
  if (magic != kMagic) return -1;
  if (status != ST_OK || !rows) return (int)status;
  for (int i = 0; i < nt; ++i)
    if (n[i] && !ReadAll(c->fd, rows[i], (size_t)n[i] * dims[i] * 4)) return -1;
  return 0;
}

}  // extern "C"
