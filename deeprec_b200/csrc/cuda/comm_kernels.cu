// NVLink 5 / NVSwitch peer-memory layer: symmetric buffers (CUDA IPC), device-side rank barrier, and the
// fused compute+collective kernels of the model-parallel embedding / data-parallel dense step.
//
//   k_mp_lookup         id dispatch + hash probe (+admission/claim) + row gather in ONE kernel: the owner of a
//                       table loads the requesters' id columns over NVLink (coalesced 8 B/sample P2P loads),
//                       probes its table, and stores bf16 rows straight into each requester's feature-major
//                       activation buffer (coalesced 32 B/sample P2P stores).  Replaces SOK's selectKernel ->
//                       NCCL all2all(counts) -> D2H + cudaStreamSynchronize -> NCCL all2all(keys) -> get_insert ->
//                       gather -> NCCL all2all(vectors) -> reorderKernel   (SURVEY §3.4, C2/C3/S1/S2/S3).
//   k_mp_sparse_grad    sparse-gradient return + dedup: the owner pulls each requester's gradient columns over
//                       NVLink and reduces them into the per-unique-key buffer with vectorised L2 atomics; the
//                       row-wise optimizer (k_apply) follows on the same stream.  Replaces gatherExKernel ->
//                       NCCL all2all(grads) -> unique -> unsorted_segment_sum -> sparse apply (C4/K10/K7).
//   k_allreduce_apply   dense gradient all-reduce fused with the optimizer: every rank loads all peers' gradient
//                       shards over NVLink in a fixed order (bitwise identical sums on all ranks), applies the
//                       update rule and writes fp32 master weights in one pass.  Replaces Horovod ncclAllReduce +
//                       separate Apply* op (C1/K9).
//   k_rank_barrier      flag barrier over peer memory (st.release.sys / ld.acquire.sys), epoch kept on device so a
//                       captured CUDA graph replays it.
#include "table.cuh"

using namespace drc;

extern "C" {
struct DrPeers {
  void* ptr[16];     // ptr[r] = this buffer as mapped in the local address space for rank r
};
}

namespace {

constexpr int kMaxRanks = 16;

// owner of a key of a ROW-sharded table (decorrelated from the probe hash)
__device__ __forceinline__ int row_owner(int64_t key, int W) { return (int)((dr_mix64((uint64_t)key ^ 0x5bd1e9955bd1e995ULL) >> 33) % (uint64_t)W); }
constexpr int kMaxChannels = 16;

// signals layout (per rank, symmetric): uint32 flags[kMaxChannels][kMaxRanks]; epochs[kMaxChannels] lives in LOCAL memory
__global__ void k_rank_barrier(DrPeers sig, uint32_t* __restrict__ epochs, int channel, int rank, int world) {
  pdl_sync();
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) { epoch = epochs[channel] + 1; epochs[channel] = epoch; }
  __syncthreads();
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();
    uint32_t* remote = reinterpret_cast<uint32_t*>(sig.ptr[r]) + channel * kMaxRanks + rank;
    st_release_sys(remote, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(sig.ptr[rank]) + channel * kMaxRanks + r;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(20); }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Requester-side dispatch for ROW-SHARDED tables: bucket this rank's ids of row table r by owning rank
// (row_owner(key) = hash % W) into compact per-owner lists that the owners read over NVLink -- the owner then probes only
// the ~B/W keys it owns instead of scanning all W*B ids of the table (SOK's selectKernel, all2all_input_dispatcher.cu:37-129,
// without the count exchange + host sync: counts live in peer-readable memory and are read after the step's first barrier).
//   bkt_key [nr][W][B] int64, bkt_b [nr][W][B] int32 (sample index), cnt [nr][W] int32 (zeroed before the launch).
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mp_partition(const int64_t* __restrict__ ids /* [T][B] local */, const int32_t* __restrict__ row_tg, int nr,
                                                      int W, int64_t B, int64_t* __restrict__ bkt_key, int32_t* __restrict__ bkt_b,
                                                      int32_t* __restrict__ cnt) {
  pdl_sync();
  const int64_t n = (int64_t)nr * B;
  const int64_t n32 = (n + 31) & ~int64_t(31);
  const unsigned lane = threadIdx.x & 31;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = i < n;
    const int64_t ii = live ? i : n - 1;
    const int r = (int)(ii / B);
    const int64_t b = ii % B;
    const int64_t key = ids[(int64_t)row_tg[r] * B + b];
    const int o = row_owner(key, W);
    const int mk = live ? (r * W + o) : -(int)(lane + 1);
    const unsigned same = __match_any_sync(0xffffffffu, mk);      // lanes of the warp going to the same (table, owner) bucket
    if (!live) continue;
    const int leader = __ffs(same) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(&cnt[mk], __popc(same));
    base = __shfl_sync(same, base, leader);
    const int64_t slot = (int64_t)mk * B + base + __popc(same & ((1u << lane) - 1u));
    bkt_key[slot] = key;
    bkt_b[slot] = (int32_t)b;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Fused dispatch + probe + gather.  Two item spaces, processed by the same blocks in 256-item chunks:
//   table-wise tables: item ((j * W + s) * B + b), j = local table, s = requesting rank, b = sample (ids read from the
//                      requester's id columns);
//   row-sharded tables: item (r, s, e), e < cnt_s[r][me]: the e-th entry of requester s's bucket for me.
// pos_out / brow use the same [table][s][B] layout for both (row tables after the table-wise ones).
// -----------------------------------------------------------------------------------------------------------------
struct MpRow {                 // row-sharded part of the exchange (nr == 0 disables it)
  DrPeers bkt_key, bkt_b, cnt; // requester-side buckets (peer-mapped)
  int32_t* cnt_local;          // [nr][W] counts copied by the owner in the forward (the backward iterates the same lists)
  int32_t* brow;               // [nr][W][B] sample index of every owned row item (backward reads demb[tg][b] of requester s)
  int nr;
};

template <int LPR>   // lanes per row = dim / 4 (float4 per lane), power of two <= 32
__global__ void __launch_bounds__(256) k_mp_lookup(const DrDeviceTable* __restrict__ tables, const int32_t* __restrict__ table_map,
                                                   const int32_t* __restrict__ table_global,   // [nl + nr] global table id
                                                   int rank, int nl, int W, int64_t B, int T, DrPeers ids_peers /* int64 [T][B] */,
                                                   DrPeers emb_peers /* bf16 [T][B][D] */, MpRow row, int train,
                                                   const int64_t* __restrict__ step_ptr, int32_t* __restrict__ pos_out,
                                                   int64_t* __restrict__ ulist, int32_t* __restrict__ nunique, int64_t ulist_cap) {
  pdl_sync();
  __shared__ int32_t s_pos[256];
  __shared__ int64_t s_key[256];
  __shared__ int32_t s_b[256];
  __shared__ int32_t s_cnt[16 * 16];
  __shared__ TouchSmem s_touch;
  (void)step_ptr; (void)T;
  const int nrw = row.nr * W;
  for (int i = threadIdx.x; i < nrw; i += blockDim.x) {
    const int r = i / W, s = i % W;
    const int32_t c = reinterpret_cast<const int32_t*>(row.cnt.ptr[s])[r * W + rank];
    s_cnt[i] = c;
    if (blockIdx.x == 0) row.cnt_local[i] = c;
  }
  __syncthreads();
  const int64_t CB = (B + 255) / 256;
  const int64_t chunks_tw = (int64_t)nl * W * CB, chunks = chunks_tw + (int64_t)nrw * CB;
  for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    int jt, s; int64_t e0, cnt;                     // jt = index into table_map / table_global
    if (c < chunks_tw) { jt = (int)(c / (W * CB)); s = (int)((c / CB) % W); e0 = (c % CB) * 256; cnt = B; }
    else { const int64_t rc = c - chunks_tw; const int r = (int)(rc / (W * CB)); s = (int)((rc / CB) % W); e0 = (rc % CB) * 256; cnt = s_cnt[r * W + s]; jt = nl + r; }
    if (e0 >= cnt) continue;                        // block-uniform
    const bool is_row = c >= chunks_tw;
    const int tg = table_global[jt];
    const DrDeviceTable& TB = tables[table_map[jt]];
    const int64_t slot_base = ((int64_t)jt * W + s) * B;           // into pos_out (and, minus the table-wise part, brow)
    // ---- phase 1: one thread per key: peer load + probe / insert / admission; bookkeeping aggregated over the block
    {
      const int64_t e = e0 + threadIdx.x;
      const bool live = e < cnt;
      int64_t key = 0, pos = -2;
      int32_t b = 0;
      bool touch = false;
      if (live) {
        if (is_row) {
          const int64_t off = ((int64_t)(jt - nl) * W + rank) * B + e;
          key = reinterpret_cast<const int64_t*>(row.bkt_key.ptr[s])[off];
          b = reinterpret_cast<const int32_t*>(row.bkt_b.ptr[s])[off];
          row.brow[((int64_t)(jt - nl) * W + s) * B + e] = b;
        } else {
          b = (int32_t)e;
          key = reinterpret_cast<const int64_t*>(ids_peers.ptr[s])[(int64_t)tg * B + e];
        }
        if (!train || TB.is_inference) {
          pos = table_find(TB, key);
        } else {
          bool inserted = false, skip = false;
          if (TB.filter_type == DR_FILTER_BLOOM) {
            pos = table_find(TB, key);
            if (pos < 0) {
              if (bloom_add_min(TB, key, 1u) < (uint32_t)TB.filter_freq) skip = true;
              else pos = table_find_or_insert(TB, key, &inserted);
            }
          } else {
            pos = table_find_or_insert(TB, key, &inserted);
          }
          if (!skip && pos < 0) TB.counters[CTR_OVERFLOW] = 1;
          if (inserted) atomicAdd(&TB.counters[CTR_NKEYS], 1);
          touch = !skip && pos >= 0;
        }
        pos_out[slot_base + e] = (int32_t)pos;
      }
      s_pos[threadIdx.x] = live ? (int32_t)pos : -2;
      s_key[threadIdx.x] = key;
      s_b[threadIdx.x] = b;
      if (train) table_touch_block(tables, touch, pos, table_map[jt], ulist, nunique, ulist_cap, s_touch);
    }
    __syncthreads();
    // ---- phase 2: LPR lanes per row copy fp32 row -> bf16 into the requester's buffer over NVLink
    constexpr int ROWS_PER_IT = 256 / LPR;
    const int lane = threadIdx.x % LPR;
    for (int it = 0; it < LPR; ++it) {
      const int li = it * ROWS_PER_IT + threadIdx.x / LPR;
      if (s_pos[li] != -2) {
        const float* src = table_read_ptr(TB, s_key[li], s_pos[li]);
        float4 v = src ? *reinterpret_cast<const float4*>(src + 4 * lane)
                       : make_float4(TB.no_permission, TB.no_permission, TB.no_permission, TB.no_permission);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(emb_peers.ptr[s]) + ((int64_t)tg * B + s_b[li]) * (4 * LPR) + 4 * lane;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    }
    __syncthreads();
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Sparse-gradient pull + dedup: gsum[tag[pos_i]] += peer_demb[s][t][b][:]   (same two item spaces as k_mp_lookup)
// -----------------------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) k_mp_sparse_grad(const DrDeviceTable* __restrict__ tables, const int32_t* __restrict__ table_map,
                                                        const int32_t* __restrict__ table_global, int nl, int W, int64_t B,
                                                        DrPeers demb_peers /* bf16 [T][B][D] */, const int32_t* __restrict__ pos,
                                                        const int32_t* __restrict__ cnt_local, const int32_t* __restrict__ brow, int nr,
                                                        float* __restrict__ gsum, int C) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int dim = 4 * LPR;
  constexpr int IPC = 256 / LPR;                 // items per chunk-iteration of one block
  float* s_acc = reinterpret_cast<float*>(smem_raw);
  int32_t* s_tag = reinterpret_cast<int32_t*>(smem_raw + (size_t)C * dim * 4);
  for (int e = threadIdx.x; e < C * dim; e += blockDim.x) s_acc[e] = 0.f;
  for (int e = threadIdx.x; e < C; e += blockDim.x) s_tag[e] = -1;
  __syncthreads();
  const int lane = threadIdx.x % LPR;
  const int gleader = (threadIdx.x & 31) / LPR * LPR;
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << gleader);
  const int nrw = nr * W;
  const int64_t CB = (B + IPC - 1) / IPC;
  const int64_t chunks_tw = (int64_t)nl * W * CB, chunks = chunks_tw + (int64_t)nrw * CB;
  for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    int jt, s; int64_t e0, cnt;
    if (c < chunks_tw) { jt = (int)(c / (W * CB)); s = (int)((c / CB) % W); e0 = (c % CB) * IPC; cnt = B; }
    else { const int64_t rc = c - chunks_tw; const int r = (int)(rc / (W * CB)); s = (int)((rc / CB) % W); e0 = (rc % CB) * IPC; cnt = cnt_local[r * W + s]; jt = nl + r; }
    const int64_t e = e0 + threadIdx.x / LPR;
    if (e >= cnt) continue;
    const int64_t slot = ((int64_t)jt * W + s) * B + e;
    const int32_t p = pos[slot];
    if (p < 0) continue;
    const int64_t b = c < chunks_tw ? e : (int64_t)brow[((int64_t)(jt - nl) * W + s) * B + e];
    const DrDeviceTable& TBa = tables[table_map[jt]];
    const int32_t u = TBa.slots[p].tag;
    if (u < 0) continue;
    const int Ce = TBa.capacity <= (1 << 17) ? C : 0;
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(demb_peers.ptr[s]) + ((int64_t)table_global[jt] * B + b) * dim + 4 * lane;
    uint2 raw = *reinterpret_cast<const uint2*>(src);
    float2 a = unpack_bf16x2(raw.x), cc = unpack_bf16x2(raw.y);
    const float4 g = make_float4(a.x, a.y, cc.x, cc.y);
    combine_add<LPR>(s_tag, s_acc, Ce, dim, u, lane, gmask, gleader, &g, 1, gsum);
  }
  __syncthreads();
  flush_combining_cache(s_tag, s_acc, C, dim, gsum);
}

// -----------------------------------------------------------------------------------------------------------------
// Dense all-reduce (one-shot over peer memory, fixed summation order) fused with the optimizer update.
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_allreduce_apply(DrPeers grad_peers, int W, float* __restrict__ w, float* __restrict__ s0,
                                                         float* __restrict__ s1, int64_t n4 /* n / 4 */, const DrOptHyper* __restrict__ hp_dev,
                                                         float* __restrict__ reduced_out) {
  pdl_sync();
  DrOptHyper hp = {};
  if (hp_dev) hp = *hp_dev;
  const float alpha = hp_dev ? dr_adam_alpha(hp) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < W; ++r) {
      int4 raw = ld_nc_v4(reinterpret_cast<const float4*>(grad_peers.ptr[r]) + i);
      g.x += __int_as_float(raw.x); g.y += __int_as_float(raw.y); g.z += __int_as_float(raw.z); g.w += __int_as_float(raw.w);
    }
    if (reduced_out) reinterpret_cast<float4*>(reduced_out)[i] = g;
    if (w) {
      float4 wv = reinterpret_cast<float4*>(w)[i];
      float4 a = s0 ? reinterpret_cast<float4*>(s0)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 b = s1 ? reinterpret_cast<float4*>(s1)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      dr_apply_elem(hp.kind, hp, alpha, false, g.x, wv.x, a.x, b.x);
      dr_apply_elem(hp.kind, hp, alpha, false, g.y, wv.y, a.y, b.y);
      dr_apply_elem(hp.kind, hp, alpha, false, g.z, wv.z, a.z, b.z);
      dr_apply_elem(hp.kind, hp, alpha, false, g.w, wv.w, a.w, b.w);
      reinterpret_cast<float4*>(w)[i] = wv;
      if (s0) reinterpret_cast<float4*>(s0)[i] = a;
      if (s1) reinterpret_cast<float4*>(s1)[i] = b;
    }
  }
}

inline int grid_for(int64_t n, int block, int max_blocks = kNumSMs * 8) {
  int64_t b = (n + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

extern "C" {

// ---- device / IPC plumbing (this library links its own static cudart: make its current device explicit) ----------
int dr_cuda_set_device(int dev) { DR_CUDA_CHECK(cudaSetDevice(dev)); return 0; }
int dr_cuda_set_sparse_blocks_per_sm(int n) { sparse_blocks_per_sm() = n < 1 ? 1 : n; return 0; }
int dr_cuda_get_device() { int d = -1; cudaGetDevice(&d); return d; }

int dr_comm_alloc(int64_t bytes, void** out) {
  DR_CUDA_CHECK(cudaMalloc(out, (size_t)bytes));
  DR_CUDA_CHECK(cudaMemset(*out, 0, (size_t)bytes));
  return 0;
}
int dr_comm_free(void* p) { DR_CUDA_CHECK(cudaFree(p)); return 0; }
int dr_comm_get_handle(void* p, void* handle64) {
  cudaIpcMemHandle_t h;
  DR_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int dr_comm_open_handle(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DR_CUDA_CHECK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int dr_comm_close_handle(void* p) { DR_CUDA_CHECK(cudaIpcCloseMemHandle(p)); return 0; }
int dr_comm_can_access_peer(int dev, int peer) { int ok = 0; cudaDeviceCanAccessPeer(&ok, dev, peer); return ok; }

int dr_comm_barrier(const DrPeers* sig, uint32_t* epochs, int channel, int rank, int world, cudaStream_t s) {
  DR_PDL_LAUNCH((k_rank_barrier), 1, 32, 0, s, *sig, epochs, channel, rank, world);
  DR_LAUNCH_CHECK();
  return 0;
}

int dr_comm_mp_partition(const int64_t* ids, const int32_t* row_tg, int nr, int W, int64_t B, int64_t* bkt_key, int32_t* bkt_b, int32_t* cnt,
                         cudaStream_t s) {
  if (nr <= 0) return 0;
  if (nr * W > 256 || W > 16) return -2;
  DR_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (size_t)nr * W * 4, s));
  DR_PDL_LAUNCH((k_mp_partition), grid_for((int64_t)nr * B, 256, kNumSMs * sparse_blocks_per_sm()), 256, 0, s, ids, row_tg, nr, W, B, bkt_key, bkt_b, cnt);
  DR_LAUNCH_CHECK();
  return 0;
}

// table_map / table_global hold the nl table-wise local tables first, then the nr row-sharded tables.
int dr_comm_mp_lookup(const DrDeviceTable* tables_dev, const int32_t* table_map, const int32_t* table_global, int rank, int nl, int nr, int W, int64_t B,
                      int T, int dim, const DrPeers* ids_peers, const DrPeers* emb_peers, const DrPeers* bkt_key_peers, const DrPeers* bkt_b_peers,
                      const DrPeers* cnt_peers, int32_t* cnt_local, int32_t* brow, int train, const int64_t* step_ptr, int32_t* pos_out,
                      int64_t* ulist, int32_t* nunique, int64_t ulist_cap, cudaStream_t s) {
  if (nl + nr == 0) return 0;
  if (nr * W > 256 || W > 16) return -2;
  MpRow row{};
  row.nr = nr;
  if (nr > 0) { row.bkt_key = *bkt_key_peers; row.bkt_b = *bkt_b_peers; row.cnt = *cnt_peers; row.cnt_local = cnt_local; row.brow = brow; }
  const int64_t chunks = (int64_t)(nl + nr) * W * ((B + 255) / 256);
  int grid = grid_for(chunks * 256, 256, kNumSMs * sparse_blocks_per_sm());
  switch (dim / 4) {
    case 2: DR_PDL_LAUNCH((k_mp_lookup<2>), grid, 256, 0, s, tables_dev, table_map, table_global, rank, nl, W, B, T, *ids_peers, *emb_peers, row, train, step_ptr, pos_out, ulist, nunique, ulist_cap); break;
    case 4: DR_PDL_LAUNCH((k_mp_lookup<4>), grid, 256, 0, s, tables_dev, table_map, table_global, rank, nl, W, B, T, *ids_peers, *emb_peers, row, train, step_ptr, pos_out, ulist, nunique, ulist_cap); break;
    case 8: DR_PDL_LAUNCH((k_mp_lookup<8>), grid, 256, 0, s, tables_dev, table_map, table_global, rank, nl, W, B, T, *ids_peers, *emb_peers, row, train, step_ptr, pos_out, ulist, nunique, ulist_cap); break;
    case 16: DR_PDL_LAUNCH((k_mp_lookup<16>), grid, 256, 0, s, tables_dev, table_map, table_global, rank, nl, W, B, T, *ids_peers, *emb_peers, row, train, step_ptr, pos_out, ulist, nunique, ulist_cap); break;
    case 32: DR_PDL_LAUNCH((k_mp_lookup<32>), grid, 256, 0, s, tables_dev, table_map, table_global, rank, nl, W, B, T, *ids_peers, *emb_peers, row, train, step_ptr, pos_out, ulist, nunique, ulist_cap); break;
    default: return -3;
  }
  DR_LAUNCH_CHECK();
  return 0;
}

int dr_comm_mp_sparse_grad(const DrDeviceTable* tables_dev, const int32_t* table_map, const int32_t* table_global, int nl, int nr, int W, int64_t B,
                           int dim, const DrPeers* demb_peers, const int32_t* pos, const int32_t* cnt_local, const int32_t* brow, float* gsum,
                           cudaStream_t s) {
  if (nl + nr == 0) return 0;
  int lpr = dim / 4;
  const int64_t chunks = (int64_t)(nl + nr) * W * ((B + 256 / lpr - 1) / (256 / lpr));
  int grid = grid_for(chunks * 256, 256, kNumSMs * sparse_blocks_per_sm());
  const int C = combining_cache_slots(dim);
  const size_t smem = (size_t)C * dim * 4 + (size_t)C * 4;
  switch (lpr) {
    case 2: DR_PDL_LAUNCH((k_mp_sparse_grad<2>), grid, 256, smem, s, tables_dev, table_map, table_global, nl, W, B, *demb_peers, pos, cnt_local, brow, nr, gsum, C); break;
    case 4: DR_PDL_LAUNCH((k_mp_sparse_grad<4>), grid, 256, smem, s, tables_dev, table_map, table_global, nl, W, B, *demb_peers, pos, cnt_local, brow, nr, gsum, C); break;
    case 8: DR_PDL_LAUNCH((k_mp_sparse_grad<8>), grid, 256, smem, s, tables_dev, table_map, table_global, nl, W, B, *demb_peers, pos, cnt_local, brow, nr, gsum, C); break;
    case 16: DR_PDL_LAUNCH((k_mp_sparse_grad<16>), grid, 256, smem, s, tables_dev, table_map, table_global, nl, W, B, *demb_peers, pos, cnt_local, brow, nr, gsum, C); break;
    case 32: DR_PDL_LAUNCH((k_mp_sparse_grad<32>), grid, 256, smem, s, tables_dev, table_map, table_global, nl, W, B, *demb_peers, pos, cnt_local, brow, nr, gsum, C); break;
    default: return -3;
  }
  DR_LAUNCH_CHECK();
  return 0;
}

// n must be a multiple of 4.  w == null => pure all-reduce into reduced_out.
int dr_comm_allreduce_apply(const DrPeers* grad_peers, int W, float* w, float* s0, float* s1, int64_t n, const DrOptHyper* hp_dev,
                            float* reduced_out, cudaStream_t s) {
  if (n % 4) return -2;
  DR_PDL_LAUNCH((k_allreduce_apply), grid_for(n / 4, 256, kNumSMs * 4), 256, 0, s, *grad_peers, W, w, s0, s1, n / 4, hp_dev, reduced_out);
  DR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
