// tcgen05 / TMEM / TMA GEMM kernels for the dense (MLP) path -- hand-written for sm_100a.
//
//   k_gemm_tn      D[M,N] = A[M,K] * B[N,K]^T (+bias)(ReLU)(* relu-mask)      bf16 in, fp32 TMEM accumulate, bf16 out
//                  persistent, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (one elected
//                  thread issues tcgen05.mma), warps2-5 = epilogue (tcgen05.ld -> regs -> fused epilogue ->
//                  16B global stores); 4-stage smem ring (128B-swizzled K-major tiles), double-buffered TMEM
//                  accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//                  Used for: Linear forward (A = activations, B = W[N,K]) and dX (A = dY, B = W^T[K,N]).
//   k_gemm_nt_splitk  dW[N_out,K_in] += dY[b,N_out]^T * X[b,K_in]            both operands MN-major straight from
//                  the row-major activations (no transposes materialised), split over the batch, fp32
//                  red.global.add.v4 epilogue.
//
// The reference has no tensor-core kernel of its own: its MLP GEMMs are cuBLAS/cuBLASLt calls
// (stream_executor/cuda/cuda_blas.cc:431-455, kernels/matmul_op_fused.cc) -- SURVEY §2.14 row L1.
#include <cuda.h>

#include "common.cuh"

using namespace drc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;    // 6 warps
constexpr int kGemmThreadsV2 = 320;  // 10 warps: TMA, MMA, 8 epilogue

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 2-D bf16 tensor map: dims {inner, outer}, row pitch in bytes, box {box_inner, box_outer}, 128B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -100;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101 - (int)r;
}

struct Epilogue {
  const float* bias;              // [N] or null
  const __nv_bfloat16* mask_src;  // aux tile [M, ld_mask] or null
  __nv_bfloat16* out;             // [M, ldc]
  float* out_f32;                 // optional fp32 copy of the output (or null; v1 path only)
  int64_t ldc, ld_mask;
  int relu;
  int aux_mode;                   // 1: out *= (aux > 0) (ReLU backward)   2: aux only feeds S2 (S2 += out * aux)
  float* S1;                      // optional column sums over M:  S1[n] += sum_m out[m, n]
  float* S2;                      // optional:  S2[n] += sum_m out[m,n] * (aux_mode == 2 ? aux[m,n] : out[m,n])
};

template <int BLOCK_N>
struct SmemLayoutTN {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (BLOCK_N < 8 ? 8 : BLOCK_N) * BLOCK_K * 2;
  static constexpr int kBBytesAligned = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytesAligned;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epilogue ep) {
  using L = SmemLayoutTN<BLOCK_N>;
  constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_sync();      // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kABytes + L::kBBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (single elected thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N < 16 ? 16 : BLOCK_N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 B (= 16 bf16) inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);            // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= Epilogue warps (TMEM -> registers -> global) =================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
      constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
        uint32_t r[32];
        if constexpr (CH == 32) {
          tmem_ld_32x32(t_row + c0, r);
        } else {
          uint32_t r16[16];
          tmem_ld_32x16(t_row + c0, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = r16[j];
        }
        tmem_ld_wait();
        const int col0 = n_blk * BLOCK_N + c0;
        if (row < M && col0 < N) {
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              if (col0 + j < N) {
                float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (ep.mask_src) {
            const __nv_bfloat16* ms = ep.mask_src + (int64_t)row * ep.ld_mask + col0;
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              if (col0 + j < N) {
                int4 raw = *reinterpret_cast<const int4*>(ms + j);
                const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = unpack_bf16x2(w[e]);
                  if (!(f.x > 0.f)) v[j + 2 * e] = 0.f;
                  if (!(f.y > 0.f)) v[j + 2 * e + 1] = 0.f;
                }
              }
            }
          }
          __nv_bfloat16* dst = ep.out + (int64_t)row * ep.ldc + col0;
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            if (col0 + j < N) {   // N is a multiple of 8 (checked on the host)
              int4 pk;
              pk.x = (int)pack_bf16x2(v[j], v[j + 1]); pk.y = (int)pack_bf16x2(v[j + 2], v[j + 3]);
              pk.z = (int)pack_bf16x2(v[j + 4], v[j + 5]); pk.w = (int)pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<int4*>(dst + j) = pk;
            }
          }
          if (ep.out_f32) {
            float* d32 = ep.out_f32 + (int64_t)row * ep.ldc + col0;
#pragma unroll
            for (int j = 0; j < CH; j += 4)
              if (col0 + j < N) *reinterpret_cast<float4*>(d32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// -------------------------------------------------------------------------------------------------
// k_gemm_tn_2cta: the same GEMM on a CTA PAIR (thread-block cluster of 2 = the two SMs of one TPC, `cta_group::2`).
//   * one 256 x BLOCK_N output tile per pair: CTA r owns rows [r*128, r*128+128) of it (its TMEM holds 128 lanes x BLOCK_N fp32
//     columns, double-buffered), loads ITS 128 rows of A and ITS HALF (BLOCK_N / 2 rows) of B per k block -- every B byte is
//     fetched once per 256 output rows instead of once per 128: the per-SM L2 -> smem traffic that bounds k_gemm_tn_v2 at these
//     shapes (34-40 % of copy bandwidth, profiles/ncu_summary.md) drops by BLOCK_N / (128 + BLOCK_N) of the B share;
//   * both CTAs' TMA loads complete on the LEADER's full barrier (`cp.async.bulk.tensor...cta_group::2`, barrier address with
//     the peer bit cleared); the leader's elected thread issues ONE `tcgen05.mma.cta_group::2` (M = 256) per 16-wide k step that
//     reads both CTAs' shared memory and writes both CTAs' TMEM; `tcgen05.commit.cta_group::2 ... multicast::cluster` frees the
//     smem stage in both CTAs / hands the accumulator to both epilogues;
//   * the epilogue warps of both CTAs release the accumulator with a remote `mbarrier.arrive` on the leader's barrier (`mapa`).
// Direct-store epilogue (bias / ReLU / ReLU-backward mask / optional fp32 copy), i.e. the k_gemm_tn feature set.
// Opt-in (DEEPREC_GEMM_2CTA=1 / dr_cuda_set_gemm_2cta): written after the round's GPU budget was spent -- first hardware run is
// tests/test_gpu_zzzzzz_gemm_2cta.py.  Reference: none (cuBLAS call, stream_executor/cuda/cuda_blas.cc:431-455).
// -------------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the pair's even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes land on the pair LEADER's mbarrier (executed by both CTAs; destination = the executing CTA's smem)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once every MMA issued so far has retired
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in cluster CTA `rank`
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}

template <int BLOCK_N>
struct SmemLayoutTN2CTA {
  static constexpr int kStages2 = BLOCK_N >= 256 ? 5 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;             // my 128 rows of A
  static constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;       // my half of B (multiple of 1024 for BLOCK_N >= 16)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = kStages2 * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_gemm_tn_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epilogue ep) {
  using L = SmemLayoutTN2CTA<BLOCK_N>;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "M = 256 pair MMA: N multiple of 16 in [16, 256]; halves of 32+");
  constexpr int S = L::kStages2;
  constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * L::kStageBytes);
  uint64_t* full_bar = bars;                  // [S]  used in the leader only (both CTAs' TMA bytes land there)
  uint64_t* empty_bar = bars + S;             // [S]  per CTA: multicast commit from the leader's MMA thread
  uint64_t* tfull_bar = bars + 2 * S;         // [2]  per CTA: multicast commit
  uint64_t* tempty_bar = bars + 2 * S + 2;    // [2]  leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_pairs * n_tiles;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int first_tile = (int)cluster_id_x(), tile_step = (int)cluster_nid_x();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) {      // the same warp index in both CTAs performs the pair allocation
    tmem_alloc_2cta(tmem_ptr, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits of BOTH CTAs are visible before any remote arrive / multicast commit / peer TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer (one per CTA: my A rows, my half of B) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
        const int row0 = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M;
        const int col0 = n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (L::kABytes + L::kBBytes));   // both CTAs' bytes
          tma_load_2d_2cta(sa, &tmA, &full_bar[stage], kb * BLOCK_K, row0);
          tma_load_2d_2cta(sb, &tmB, &full_bar[stage], kb * BLOCK_K, col0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: ONE thread of the LEADER CTA drives both SMs' tensor cores =================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_2cta(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit_2cta(&tfull_bar[acc]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= Epilogue warps of both CTAs: my 128 rows of the pair tile =================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
        const int col0 = n_blk * BLOCK_N + c0;
        if (row < M && col0 < N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < N) {
                float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (ep.mask_src) {
            const __nv_bfloat16* ms = ep.mask_src + (int64_t)row * ep.ld_mask + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < N) {
                int4 raw = *reinterpret_cast<const int4*>(ms + j);
                const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = unpack_bf16x2(w[e]);
                  if (!(f.x > 0.f)) v[j + 2 * e] = 0.f;
                  if (!(f.y > 0.f)) v[j + 2 * e + 1] = 0.f;
                }
              }
            }
          }
          __nv_bfloat16* dst = ep.out + (int64_t)row * ep.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < N) {   // N is a multiple of 8 (checked on the host)
              int4 pk;
              pk.x = (int)pack_bf16x2(v[j], v[j + 1]); pk.y = (int)pack_bf16x2(v[j + 2], v[j + 3]);
              pk.z = (int)pack_bf16x2(v[j + 4], v[j + 5]); pk.w = (int)pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<int4*>(dst + j) = pk;
            }
          }
          if (ep.out_f32) {
            float* d32 = ep.out_f32 + (int64_t)row * ep.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < N) *reinterpret_cast<float4*>(d32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);     // the leader's MMA thread owns the accumulator hand-back
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  __syncwarp();         // re-converge the single-lane role branches before the aligned cluster barrier
  tc_fence_before();
  cluster_sync_all();   // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  if (warp == 1) tmem_dealloc_2cta(tmem_base, kTmemCols);
}

// -------------------------------------------------------------------------------------------------
// k_gemm_tn_v2: same mainloop, production epilogue.
//   * accumulator -> registers (tcgen05.ld) -> fused math -> bf16 -> 128B-swizzled smem staging -> TMA store
//     (cp.async.bulk.tensor, full 128 B lines; M/N tails clipped by the tensor map), double-buffered per warp;
//   * the aux tile (ReLU-backward mask, or the BatchNorm-backward partner activation) is fetched with
//     coalesced 16 B loads (8 lanes per 128 B row segment) through the same swizzled staging layout;
//   * optional per-column batch statistics S1 += sum(out), S2 += sum(out*out | out*aux) are reduced with a
//     31-shuffle warp transpose-reduce and one global atomic per column per 32-row slab -- this is what
//     removes the separate BatchNorm-statistics / bias-gradient passes over the activations.
// -------------------------------------------------------------------------------------------------
template <int BLOCK_N> struct StagesFor { static constexpr int value = BLOCK_N >= 256 ? 3 : 4; };

template <int BLOCK_N>
struct SmemLayoutTN2 {
  static constexpr int kSt = StagesFor<BLOCK_N>::value;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreBytes = 8 * 4096;          // 8 epilogue warps x [32 rows][128 B] staging
  static constexpr int kAuxBytes = 8 * 4096;            // 8 epilogue warps x [32 rows][128 B] aux tile
  static constexpr int kBiasCols = 1024;                // bias vector staged in smem once per CTA (N <= kBiasCols)
  static constexpr int kBiasBytes = kBiasCols * 4;
  static constexpr int kBResBytes = 0;                  // (B-resident variant only)
  static constexpr int kBSliceBytes = kBBytes;
  static constexpr int kMaxKb = 1 << 30;
  static constexpr int kTotal = kSt * kStageBytes + kStoreBytes + kAuxBytes + kBiasBytes + 1024 + 256;
};

// B-resident ("weight-stationary") variant of the same kernel.  In the MLP GEMMs the B operand is the layer's weight matrix:
// every M tile of a CTA multiplies the SAME [BLOCK_N, K] slab, yet the streaming kernel re-fetches it from L2 for every tile
// (B is 2/3 of the shared-memory ingest at BLOCK_N = 256).  Here a CTA owns one N block for its whole life, loads the slab once
// (all K blocks, <= kMaxKb x BLOCK_N x 128 B) and then streams only A through the TMA ring; the aux tile and the store staging
// share one buffer to make room.  Tile order: CTA c -> n_blk = c % n_tiles, m_blk = c / n_tiles, += gridDim / n_tiles.
template <int BLOCK_N>
struct SmemLayoutTN3 {
  static constexpr int kSt = BLOCK_N >= 128 ? 3 : 4;    // A-only stages
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes;
  static constexpr int kStoreBytes = 8 * 4096;
  static constexpr int kAuxBytes = 0;                   // aliased onto the store staging (a lane reads its aux chunk before it overwrites it)
  static constexpr int kBiasCols = 1024;
  static constexpr int kBiasBytes = kBiasCols * 4;
  static constexpr int kMaxKb = 8;                      // K <= 512
  static constexpr int kBSliceBytes = kBBytes;
  static constexpr int kBResBytes = kMaxKb * kBSliceBytes;
  static constexpr int kTotal = kBResBytes + kSt * kStageBytes + kStoreBytes + kBiasBytes + 1024 + 256;
};

template <int BLOCK_N, bool BRES> struct LayoutSel { using type = SmemLayoutTN2<BLOCK_N>; };
template <int BLOCK_N> struct LayoutSel<BLOCK_N, true> { using type = SmemLayoutTN3<BLOCK_N>; };

// lane c ends with the sum over the warp's 32 rows of column c (v is destroyed)
__device__ __forceinline__ float warp_col_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float keep = upper ? v[k + off] : v[k];
      const float send = upper ? v[k] : v[k + off];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int BLOCK_N, bool BRES = false>
__global__ void __launch_bounds__(kGemmThreadsV2, 1)
k_gemm_tn_v2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
             int M, int N, int K, Epilogue ep) {
  using L = typename LayoutSel<BLOCK_N, BRES>::type;
  constexpr int kSt = L::kSt;
  constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres_base = smem_al;                         // [kMaxKb][BLOCK_N x 128 B] resident weight slab (BRES only; 0 bytes otherwise)
  uint8_t* smem = smem_al + L::kBResBytes;              // TMA ring
  uint8_t* store_base = smem + kSt * L::kStageBytes;
  uint8_t* aux_base = store_base + (BRES ? 0 : L::kStoreBytes);
  float* s_bias = reinterpret_cast<float*>(store_base + L::kStoreBytes + L::kAuxBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(store_base + L::kStoreBytes + L::kAuxBytes + L::kBiasBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kSt;
  uint64_t* tfull_bar = bars + 2 * kSt;
  uint64_t* tempty_bar = bars + 2 * kSt + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kSt + 4);
  uint64_t* bfull_bar = bars + 2 * kSt + 5;             // BRES: the weight slab has landed
  // the epilogue's bias reads were its largest stall (global loads, long scoreboard, per 32-column half): stage the vector once
  const bool smem_bias = ep.bias != nullptr && N <= L::kBiasCols;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    for (int i = 0; i < kSt; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    if constexpr (BRES) mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_sync();      // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail
  if (smem_bias) {   // (the bias may be produced by the previous kernel: BatchNorm folding)
    for (int i = threadIdx.x; i < L::kBiasCols; i += blockDim.x) s_bias[i] = i < N ? ep.bias[i] : 0.f;
    __syncthreads();
  }
  // tile schedule: streaming kernel = round-robin over all (m, n) tiles; B-resident kernel = fixed n block, strided m blocks.
  // (Spelled out inside every role branch so that the single-thread producer / MMA loops keep their indices in uniform registers.)
#define DR_TILE_LOOP(it)                                                                                       \
  for (int it = BRES ? (int)blockIdx.x / n_tiles : (int)blockIdx.x; it < (BRES ? m_tiles : num_tiles);       \
       it += BRES ? (int)gridDim.x / n_tiles : (int)gridDim.x)
#define DR_TILE_M(it) (BRES ? (it) : (it) / n_tiles)
#define DR_TILE_N(it) (BRES ? (int)blockIdx.x % n_tiles : (it) % n_tiles)
#define DR_TILE_ANY() ((BRES ? (int)blockIdx.x / n_tiles : (int)blockIdx.x) < (BRES ? m_tiles : num_tiles))

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if constexpr (BRES) {
        if (DR_TILE_ANY()) {          // the whole [BLOCK_N, K] weight slab, once
          mbar_expect_tx(bfull_bar, (uint32_t)(num_kb * L::kBSliceBytes));
          for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(bres_base + kb * L::kBSliceBytes, &tmB, bfull_bar, kb * BLOCK_K, DR_TILE_N(0) * BLOCK_N);
        }
      }
      DR_TILE_LOOP(it) {
        const int m_blk = DR_TILE_M(it), n_blk = DR_TILE_N(it);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          if constexpr (!BRES) {
            uint8_t* sb = sa + L::kABytes;
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if constexpr (BRES) {
        if (DR_TILE_ANY()) { mbar_wait(bfull_bar, 0); tc_fence_after(); }
      }
      DR_TILE_LOOP(it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = BRES ? smem_u32(bres_base + kb * L::kBSliceBytes) : sa + L::kABytes;
          const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps =================
    // 8 epilogue warps: two per TMEM lane quarter (q = warp % 4 is a hardware rule), interleaved over the 64-column
    // chunks, so every SM sub-partition hosts two independent epilogue instruction streams (latency hiding).
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* my_store = store_base + ew * 4096;
    uint8_t* my_aux = aux_base + ew * 4096;
    const uint32_t swz = (uint32_t)(lane & 7);
    int acc = 0; uint32_t acc_phase = 0;
    DR_TILE_LOOP(it) {
      const int m_blk = DR_TILE_M(it), n_blk = DR_TILE_N(it);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
        const int col0 = n_blk * BLOCK_N + c0;
        if (col0 >= N) break;                           // warp-uniform
        if (((c0 >> 6) & 1) != half) continue;          // the sibling warp of this lane quarter owns this chunk
        if constexpr (BRES) {   // aux and store staging share one buffer: the previous chunk's TMA store must have read it first
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        // ---- aux tile [32 rows x 64 cols] -> swizzled smem (coalesced: 8 lanes cover one 128 B row segment)
        if (ep.mask_src) {
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3), ch = lane & 7;
            int4 val = make_int4(0, 0, 0, 0);
            if (row0 + r < M && col0 + ch * 8 < N) val = ld_nc_v4(ep.mask_src + (int64_t)(row0 + r) * ep.ld_mask + col0 + ch * 8);
            *reinterpret_cast<int4*>(my_aux + r * 128 + ((ch ^ (r & 7)) << 4)) = val;
          }
          __syncwarp();
        }
        // the staging buffer was handed to TMA one chunk ago: wait until it has been read
        if constexpr (!BRES) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        uint8_t* stg = my_store;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c0 + h * 32, r);
          tmem_ld_wait();
          const int colh = col0 + h * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (smem_bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {      // broadcast ld.shared.v4 (columns >= N hold 0)
              const float4 b = *reinterpret_cast<const float4*>(s_bias + colh + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          } else if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (colh + j < N) {
                float4 b = *reinterpret_cast<const float4*>(ep.bias + colh + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          float w[32];     // aux values (only materialised when needed)
          if (ep.mask_src) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              int4 raw = *reinterpret_cast<const int4*>(my_aux + lane * 128 + (((uint32_t)(h * 4 + j4) ^ swz) << 4));
              const uint32_t ww[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { float2 f = unpack_bf16x2(ww[e]); w[j4 * 8 + 2 * e] = f.x; w[j4 * 8 + 2 * e + 1] = f.y; }
            }
            if (ep.aux_mode == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (!(w[j] > 0.f)) v[j] = 0.f;
            }
          }
          // ---- bf16 pack -> swizzled staging
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            int4 pk;
            pk.x = (int)pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]); pk.y = (int)pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]);
            pk.z = (int)pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]); pk.w = (int)pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]);
            *reinterpret_cast<int4*>(stg + lane * 128 + (((uint32_t)(h * 4 + j4) ^ swz) << 4)) = pk;
          }
          // ---- fused column statistics (rows beyond M contribute nothing)
          if (ep.S1) {
            float p[32];
            if (ep.S2) {
              float s2[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) s2[j] = row_ok ? v[j] * ((ep.aux_mode == 2) ? w[j] : v[j]) : 0.f;
              const float t2 = warp_col_reduce(s2, lane);
              if (colh + lane < N) atomicAdd(ep.S2 + colh + lane, t2);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = row_ok ? v[j] : 0.f;
            const float t1 = warp_col_reduce(p, lane);
            if (colh + lane < N) atomicAdd(ep.S1 + colh + lane, t1);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < M) {
          tma_store_2d(&tmC, stg, col0, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

#undef DR_TILE_LOOP
#undef DR_TILE_M
#undef DR_TILE_N
#undef DR_TILE_ANY

// -------------------------------------------------------------------------------------------------
// dW[N_out, K_in] += sum_b dY[b, N_out] * X[b, K_in]   (split over b; fp32 atomics epilogue)
//   A operand = dY^T  : M = N_out, MN-major (tensor map over dY: inner = N_out, outer = batch)
//   B operand = X^T   : N = K_in , MN-major (tensor map over X : inner = K_in , outer = batch)
// smem per stage: A = 2 chunks [BLOCK_KB rows][64 elems]; B = BLOCK_N/64 such chunks.
// -------------------------------------------------------------------------------------------------
constexpr int BLOCK_KB = 64;            // batch rows (reduction) per pipeline stage
constexpr int kChunkBytes = BLOCK_KB * 128;

template <int BLOCK_N>
struct SmemLayoutNT {
  static constexpr int kAChunks = BLOCK_M / 64;
  static constexpr int kBChunks = BLOCK_N / 64;
  static constexpr int kStageBytes = (kAChunks + kBChunks) * kChunkBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 + 256;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_nt_splitk(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Mo /*N_out*/, int No /*K_in*/,
                 int batch, int rows_per_split, float* __restrict__ dW, int64_t ldw) {
  using L = SmemLayoutNT<BLOCK_N>;
  constexpr int kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y, split = blockIdx.z;
  const int b0 = split * rows_per_split;
  const int b1 = min(batch, b0 + rows_per_split);
  const int num_kb = (b1 - b0 + BLOCK_KB - 1) / BLOCK_KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_sync();      // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail
  if (num_kb <= 0) { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, kTmemCols); return; }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kAChunks * kChunkBytes;
        mbar_expect_tx(&full_bar[stage], L::kStageBytes);
        const int brow = b0 + kb * BLOCK_KB;   // rows beyond `batch` are zero-filled by TMA; rows in [b1, batch) of a
                                               // short last stage belong to the next split: masked by box clamp below
#pragma unroll
        for (int c = 0; c < L::kAChunks; ++c) tma_load_2d(sa + c * kChunkBytes, &tmA, &full_bar[stage], m_blk * BLOCK_M + c * 64, brow);
#pragma unroll
        for (int c = 0; c < L::kBChunks; ++c) tma_load_2d(sb + c * kChunkBytes, &tmB, &full_bar[stage], n_blk * BLOCK_N + c * 64, brow);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kAChunks * kChunkBytes;
        // MN-major SW128: LBO = distance between 64-element MN chunks, SBO = 8 k-rows (1024 B)
        const uint64_t adesc = umma_desc_sw128(sa, kChunkBytes, 1024);
        const uint64_t bdesc = umma_desc_sw128(sb, kChunkBytes, 1024);
#pragma unroll
        for (int k = 0; k < BLOCK_KB / UMMA_K; ++k) {
          // 16 k-rows = 2048 B  -> +128 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tfull_bar);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int row = m_blk * BLOCK_M + q * 32 + lane;      // N_out index
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(t_row + c0, r);
      tmem_ld_wait();
      const int col0 = n_blk * BLOCK_N + c0;               // K_in index
      if (row < Mo) {
        float* dst = dW + (int64_t)row * ldw + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col0 + j + 3 < No) {
            red_add_v4_f32(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
            for (int e = 0; e < 4; ++e) if (col0 + j + e < No) atomicAdd(dst + j + e, __uint_as_float(r[j + e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN>
int launch_tn(const CUtensorMap& ta, const void* B, int M, int N, int K, int64_t ldb, const Epilogue& ep, int max_ctas, cudaStream_t s) {
  CUtensorMap tb;
  int rc = make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BLOCK_K, BN < 8 ? 8 : BN);
  if (rc) return rc;
  using L = SmemLayoutTN<BN>;
  static DrPerDeviceOnce attr_once; bool& attr_set = attr_once();
  if (!attr_set) {
    DR_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  if (grid > max_ctas) grid = max_ctas;
  DR_PDL_LAUNCH((k_gemm_tn<BN>), grid, kGemmThreads, L::kTotal, s, ta, tb, M, N, K, ep);
  DR_LAUNCH_CHECK();
  return 0;
}

int make_tmap_c(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes) {
  return make_tmap(m, ptr, inner, outer, pitch_bytes, 64, 32);     // 64 cols (128 B) x 32 rows, 128B swizzle
}

template <int BN, bool BRES = false>
int launch_tn_v2(const CUtensorMap& ta, const void* B, int M, int N, int K, int64_t ldb, const Epilogue& ep, int max_ctas, cudaStream_t s) {
  CUtensorMap tb, tc;
  int rc = make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BLOCK_K, BN);
  if (rc) return rc;
  rc = make_tmap_c(&tc, ep.out, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldc * 2);
  if (rc) return rc;
  using L = typename LayoutSel<BN, BRES>::type;
  static_assert(L::kTotal <= 227 * 1024, "shared-memory budget");
  static DrPerDeviceOnce attr_once; bool& attr_set = attr_once();
  if (!attr_set) {
    DR_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn_v2<BN, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  if (grid > max_ctas) grid = max_ctas;
  if (BRES) grid = grid / n_tiles * n_tiles;        // every CTA owns one n block: the grid is a whole number of n-block groups
  DR_PDL_LAUNCH((k_gemm_tn_v2<BN, BRES>), grid, kGemmThreadsV2, L::kTotal, s, ta, tb, tc, M, N, K, ep);
  DR_LAUNCH_CHECK();
  return 0;
}

// DEEPREC_GEMM_BRES=1 selects the B-resident (weight-stationary) kernel where it applies (K <= 512, enough CTAs for one per n block).
inline int& gemm_bres_enabled() { static int v = [] { const char* e = getenv("DEEPREC_GEMM_BRES"); return (e && e[0] == '1') ? 1 : 0; }(); return v; }

// CTA-pair launcher: grid = 2 x min(pair tiles, max_ctas / 2); the cluster shape is a kernel attribute (__cluster_dims__).
template <int BN>
int launch_tn_2cta(const CUtensorMap& ta, const void* B, int M, int N, int K, int64_t ldb, const Epilogue& ep, int max_ctas, cudaStream_t s) {
  CUtensorMap tb;
  int rc = make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BLOCK_K, BN / 2);
  if (rc) return rc;
  using L = SmemLayoutTN2CTA<BN>;
  static_assert(L::kTotal <= 227 * 1024, "shared-memory budget");
  static DrPerDeviceOnce attr_once; bool& attr_set = attr_once();
  if (!attr_set) {
    DR_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn_2cta<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + BN - 1) / BN;
  int pairs = m_pairs * n_tiles;
  if (pairs > max_ctas / 2) pairs = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
  k_gemm_tn_2cta<BN><<<2 * pairs, kGemmThreads, L::kTotal, s>>>(ta, tb, M, N, K, ep);
  DR_LAUNCH_CHECK();
  return 0;
}

// DEEPREC_GEMM_2CTA=1 routes the direct-store-epilogue shapes (no fused statistics) with N >= 64 to the CTA-pair kernel.
inline int& gemm_2cta_enabled() { static int v = [] { const char* e = getenv("DEEPREC_GEMM_2CTA"); return (e && e[0] == '1') ? 1 : 0; }(); return v; }

template <int BN>
int launch_nt(const CUtensorMap& ta, const CUtensorMap& tb, int Mo, int No, int batch, int splits, float* dW, int64_t ldw, cudaStream_t s) {
  using L = SmemLayoutNT<BN>;
  static DrPerDeviceOnce attr_once; bool& attr_set = attr_once();
  if (!attr_set) {
    DR_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_nt_splitk<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int rows_per_split = ((batch + splits - 1) / splits + BLOCK_KB - 1) / BLOCK_KB * BLOCK_KB;
  splits = (batch + rows_per_split - 1) / rows_per_split;
  dim3 grid((Mo + BLOCK_M - 1) / BLOCK_M, (No + BN - 1) / BN, splits);
  DR_PDL_LAUNCH((k_gemm_nt_splitk<BN>), grid, kGemmThreads, L::kTotal, s, ta, tb, Mo, No, batch, rows_per_split, dW, ldw);
  DR_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int dr_cuda_gemm_tn_ex(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const float* bias, int relu,
                       const void* mask_src, int64_t ld_mask, int aux_mode, void* out, int64_t ldc, float* out_f32, float* S1, float* S2,
                       int max_ctas, int force_v1, cudaStream_t s);

// out[M,N] = A[M,K](lda) * B[N,K](ldb)^T (+bias)(relu)(*mask).  Requirements: bf16, K-major, lda/ldb/ldc multiples of 8,
// N multiple of 8, pointers 16B-aligned.  out_f32 optional.
int dr_cuda_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const float* bias, int relu,
                    const void* mask_src, int64_t ld_mask, void* out, int64_t ldc, float* out_f32, int max_ctas, cudaStream_t s) {
  return dr_cuda_gemm_tn_ex(A, lda, B, ldb, M, N, K, bias, relu, mask_src, ld_mask, mask_src ? 1 : 0, out, ldc, out_f32, nullptr, nullptr,
                            max_ctas, 0, s);
}

// Extended entry: aux_mode 1 = ReLU-backward mask, 2 = statistics partner; S1/S2 = fused column statistics (v2 epilogue,
// N >= 64 and no fp32 copy).  force_v1 selects the direct-store epilogue (kept for A/B measurements).
int dr_cuda_gemm_tn_ex(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const float* bias, int relu,
                       const void* mask_src, int64_t ld_mask, int aux_mode, void* out, int64_t ldc, float* out_f32, float* S1, float* S2,
                       int max_ctas, int force_v1, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 8)) return -2;
  if (max_ctas <= 0) max_ctas = kNumSMs;
  CUtensorMap ta;
  int rc = make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  Epilogue ep{bias, (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)out, out_f32, ldc, ld_mask, relu, aux_mode, S1, S2};
  if (gemm_2cta_enabled() && !force_v1 && N >= 64 && !S1 && !S2 && aux_mode != 2) {
    if (N <= 128) return launch_tn_2cta<128>(ta, B, M, N, K, ldb, ep, max_ctas, s);
    return launch_tn_2cta<256>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  }
  const bool v2 = !force_v1 && N > 32 && out_f32 == nullptr;
  if (v2 && gemm_bres_enabled() && (K + BLOCK_K - 1) / BLOCK_K <= SmemLayoutTN3<128>::kMaxKb) {
    if (N <= 64) return launch_tn_v2<64, true>(ta, B, M, N, K, ldb, ep, max_ctas, s);
    if ((N + 127) / 128 <= max_ctas) return launch_tn_v2<128, true>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  }
  if (v2) {
    if (N <= 64) return launch_tn_v2<64>(ta, B, M, N, K, ldb, ep, max_ctas, s);
    if (N <= 128) return launch_tn_v2<128>(ta, B, M, N, K, ldb, ep, max_ctas, s);
    return launch_tn_v2<256>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  }
  if (S1 || S2 || aux_mode == 2) return -4;    // statistics need the v2 epilogue
  if (N <= 16) return launch_tn<16>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  if (N <= 32) return launch_tn<32>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  if (N <= 64) return launch_tn<64>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  if (N <= 128) return launch_tn<128>(ta, B, M, N, K, ldb, ep, max_ctas, s);
  return launch_tn<256>(ta, B, M, N, K, ldb, ep, max_ctas, s);
}

// dW[N_out,K_in](ldw, fp32, accumulated into) += dY[batch,N_out](ldy)^T * X[batch,K_in](ldx)
// A/B switch for the B-resident GEMM variant (same effect as DEEPREC_GEMM_BRES, settable at run time; returns the previous value).
int dr_cuda_set_gemm_2cta(int on) { int prev = gemm_2cta_enabled(); gemm_2cta_enabled() = on ? 1 : 0; return prev; }
int dr_cuda_set_gemm_bres(int on) { int prev = gemm_bres_enabled(); gemm_bres_enabled() = on ? 1 : 0; return prev; }

int dr_cuda_gemm_dw(const void* dY, int64_t ldy, const void* X, int64_t ldx, int batch, int N_out, int K_in, float* dW, int64_t ldw,
                    int splits, cudaStream_t s) {
  if (batch <= 0 || N_out <= 0 || K_in <= 0) return 0;
  if ((ldy % 8) || (ldx % 8)) return -2;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, dY, (uint64_t)N_out, (uint64_t)batch, (uint64_t)ldy * 2, 64, BLOCK_KB);
  if (rc) return rc;
  rc = make_tmap(&tb, X, (uint64_t)K_in, (uint64_t)batch, (uint64_t)ldx * 2, 64, BLOCK_KB);
  if (rc) return rc;
  int m_tiles = (N_out + BLOCK_M - 1) / BLOCK_M;
  if (K_in <= 64) {
    if (splits <= 0) splits = max(1, kNumSMs / m_tiles);
    return launch_nt<64>(ta, tb, N_out, K_in, batch, splits, dW, ldw, s);
  }
  if (K_in <= 128) {
    if (splits <= 0) splits = max(1, kNumSMs / m_tiles);
    return launch_nt<128>(ta, tb, N_out, K_in, batch, splits, dW, ldw, s);
  }
  int n_tiles = (K_in + 255) / 256;
  if (splits <= 0) splits = max(1, kNumSMs / (m_tiles * n_tiles));
  return launch_nt<256>(ta, tb, N_out, K_in, batch, splits, dW, ldw, s);
}

}  // extern "C"
