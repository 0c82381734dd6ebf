#line 1 "/root/repo/deeprec_b200/csrc/cuda/interaction_kernels.cu"
// Feature-interaction kernels: DLRM pairwise dot (strict lower triangle of F F^T, concatenated with the
// dense vector), DeepFM second-order FM term, DIN attention pooling.  One warp per sample; lane i owns
// feature row i in registers, rows are broadcast through shared memory.
//
// Reference: modelzoo/dlrm/train.py:121-133 (_dot_op: matmul + boolean_mask), modelzoo/deepfm/train.py:178-187,
// modelzoo/din/train.py:143-188.  The interaction is a batch of 27x16 Gram matrices -- 23 kFLOP per sample
// against 1.6 kB of traffic -- so it is HBM-bound by two orders of magnitude; it reads the embedding rows
// exactly once, straight from the (peer-written) feature-major receive buffer, and emits the padded bf16
// activation the tcgen05 top-MLP GEMM consumes through TMA.
#include "sp_sync.cuh"

using namespace drc;

namespace {

template <int D>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* p, float (&f)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 8) {
    int4 raw = ld_nc_v4(p + c);
    const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 v = unpack_bf16x2(w[e]); f[c + 2 * e] = v.x; f[c + 2 * e + 1] = v.y; }
  }
}
template <int D>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* p, const float (&f)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 8) {
    int4 pk;
    pk.x = (int)pack_bf16x2(f[c], f[c + 1]); pk.y = (int)pack_bf16x2(f[c + 2], f[c + 3]);
    pk.z = (int)pack_bf16x2(f[c + 4], f[c + 5]); pk.w = (int)pack_bf16x2(f[c + 6], f[c + 7]);
    *reinterpret_cast<int4*>(p + c) = pk;
  }
}

// Z[b] = [ x[b] (D) | { <F_i, F_j> : 0 <= j < i < F } | 0-pad ],  F_0 = x[b], F_t = emb[t-1][b]
template <int D>
__global__ void __launch_bounds__(256) k_dot_fwd(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ emb,
                                                 int64_t emb_stride_t, int64_t emb_stride_b, int T, int64_t B,
                                                 __nv_bfloat16* __restrict__ Z, int64_t ldz) {
  constexpr int WPB = 8;
  using emu_sh_4613001 = float[WPB][32][D]; emu_sh_4613001& sF = *reinterpret_cast<emu_sh_4613001*>(emu::shared_var(4613001, sizeof(emu_sh_4613001)));
  using emu_sh_4613002 = __nv_bfloat16[WPB][512]; emu_sh_4613002& sZ = *reinterpret_cast<emu_sh_4613002*>(emu::shared_var(4613002, sizeof(emu_sh_4613002)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = T + 1;
  for (int64_t b = (int64_t)blockIdx.x * WPB + warp; b < B; b += (int64_t)gridDim.x * WPB) {
    float f[D];
    if (lane < F) {
      const __nv_bfloat16* src = lane == 0 ? x + b * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + b * emb_stride_b;
      load_row_bf16<D>(src, f);
#pragma unroll
      for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(&sF[warp][lane][c]) = make_float4(f[c], f[c + 1], f[c + 2], f[c + 3]);
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) sZ[warp][c] = __float2bfloat16(f[c]);
      }
    }
    __syncwarp();
    const int base = D + lane * (lane - 1) / 2;
    for (int j = 0; j < F - 1; ++j) {
      float g = 0.f;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        float4 v = *reinterpret_cast<const float4*>(&sF[warp][j][c]);
        g += f[c] * v.x + f[c + 1] * v.y + f[c + 2] * v.z + f[c + 3] * v.w;
      }
      if (lane < F && j < lane) sZ[warp][base + j] = __float2bfloat16(g);
    }
    const int used = D + F * (F - 1) / 2;
    for (int c = used + lane; c < ldz; c += 32) sZ[warp][c] = __float2bfloat16(0.f);
    __syncwarp();
    for (int c = lane * 8; c < ldz; c += 256)
      *reinterpret_cast<int4*>(Z + b * ldz + c) = *reinterpret_cast<const int4*>(&sZ[warp][c]);
    __syncwarp();
  }
}

// dF_i = sum_{j != i} S_ij F_j (+ dZ[0:D] for i = 0),  S symmetric from the lower-triangle grads.
template <int D>
__global__ void __launch_bounds__(256) k_dot_bwd(const __nv_bfloat16* __restrict__ dZ, int64_t ldz, const __nv_bfloat16* __restrict__ x,
                                                 int64_t ldx, const __nv_bfloat16* __restrict__ emb, int64_t emb_stride_t,
                                                 int64_t emb_stride_b, int T, int64_t B, __nv_bfloat16* __restrict__ dx, int64_t lddx,
                                                 __nv_bfloat16* __restrict__ demb, int64_t demb_stride_t, int64_t demb_stride_b) {
  constexpr int WPB = 8;
  using emu_sh_4613003 = float[WPB][32][D]; emu_sh_4613003& sF = *reinterpret_cast<emu_sh_4613003*>(emu::shared_var(4613003, sizeof(emu_sh_4613003)));
  using emu_sh_4613004 = float[WPB][512]; emu_sh_4613004& sG = *reinterpret_cast<emu_sh_4613004*>(emu::shared_var(4613004, sizeof(emu_sh_4613004)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = T + 1;
  const int used = D + F * (F - 1) / 2;
  for (int64_t b = (int64_t)blockIdx.x * WPB + warp; b < B; b += (int64_t)gridDim.x * WPB) {
    float f[D];
    if (lane < F) {
      const __nv_bfloat16* src = lane == 0 ? x + b * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + b * emb_stride_b;
      load_row_bf16<D>(src, f);
#pragma unroll
      for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(&sF[warp][lane][c]) = make_float4(f[c], f[c + 1], f[c + 2], f[c + 3]);
    }
    for (int c = lane * 8; c < ldz; c += 256) {
      int4 raw = ld_nc_v4(dZ + b * ldz + c);
      const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 v = unpack_bf16x2(w[e]); sG[warp][c + 2 * e] = v.x; sG[warp][c + 2 * e + 1] = v.y; }
    }
    __syncwarp();
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = (lane == 0) ? sG[warp][c] : 0.f;
    const int mybase = D + lane * (lane - 1) / 2;
    for (int j = 0; j < F; ++j) {
      float s = 0.f;
      if (lane < F) {
        if (j < lane) s = sG[warp][mybase + j];
        else if (j > lane) s = sG[warp][D + j * (j - 1) / 2 + lane];
      }
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        float4 v = *reinterpret_cast<const float4*>(&sF[warp][j][c]);
        acc[c] += s * v.x; acc[c + 1] += s * v.y; acc[c + 2] += s * v.z; acc[c + 3] += s * v.w;
      }
    }
    (void)used;
    if (lane == 0) store_row_bf16<D>(dx + b * lddx, acc);
    else if (lane < F) store_row_bf16<D>(demb + (int64_t)(lane - 1) * demb_stride_t + b * demb_stride_b, acc);
    __syncwarp();
  }
}

// -------------------------------------------------------------------------------------------------
// DeepFM second-order term: fm[b, :] = 0.5 * ((sum_t e_t)^2 - sum_t e_t^2)  and its backward
//   d e_t = dfm * (sum_t e - e_t).   emb: [T][B][D] bf16 (strided), out fp32/bf16 [B, D].
// -------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_fm_fwd(const __nv_bfloat16* __restrict__ emb, int64_t st, int64_t sb, int T, int64_t B,
                                                __nv_bfloat16* __restrict__ out, int64_t ldo, float* __restrict__ sum_out) {
  constexpr int LPR = D / 8;                 // lanes per sample (8 elements per lane)
  const int lane = threadIdx.x % LPR;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x / LPR;
  for (int64_t b = gid; b < B; b += gstride) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      float f[8];
      load_row_bf16<8>(emb + t * st + b * sb + lane * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.5f * (s[j] * s[j] - q[j]);
    store_row_bf16<8>(out + b * ldo + lane * 8, o);
    if (sum_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sum_out[b * D + lane * 8 + j] = s[j];
    }
  }
}
template <int D>
__global__ void __launch_bounds__(256) k_fm_bwd(const __nv_bfloat16* __restrict__ dfm, int64_t ldd, const __nv_bfloat16* __restrict__ emb,
                                                int64_t st, int64_t sb, const float* __restrict__ sum_in, int T, int64_t B,
                                                __nv_bfloat16* __restrict__ demb, int64_t dst, int64_t dsb, int accumulate) {
  constexpr int LPR = D / 8;
  const int lane = threadIdx.x % LPR;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x / LPR;
  for (int64_t b = gid; b < B; b += gstride) {
    float g[8], s[8];
    load_row_bf16<8>(dfm + b * ldd + lane * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = sum_in[b * D + lane * 8 + j];
    for (int t = 0; t < T; ++t) {
      float f[8], o[8];
      load_row_bf16<8>(emb + t * st + b * sb + lane * 8, f);
      if (accumulate) load_row_bf16<8>(demb + t * dst + b * dsb + lane * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (accumulate ? o[j] : 0.f) + g[j] * (s[j] - f[j]);
      store_row_bf16<8>(demb + t * dst + b * dsb + lane * 8, o);
    }
  }
}


#ifndef DR_CUDA_EMU   // tcgen05 / TMEM: hardware only (the emulation build runs the SIMT kernels above for every shape)
// =================================================================================================
// tcgen05 versions (D = 16, T + 1 <= 32): FOUR samples share one 128-row UMMA tile.
//   forward :  G = A A^T with A[128 x 16] = 4 samples x 32 (padded) feature rows.  One tcgen05.mma (M=128, N=128, K=16)
//              per 4 samples; warp w reads the w-th diagonal 32x32 block of the accumulator straight out of TMEM
//              (tcgen05.ld 32x32b: lane i <- row i of its sample's Gram matrix) and emits the strict lower triangle.
//   backward:  dF = S F with S the symmetric 32x32 gradient matrix of a sample; A = blockdiag(S_0..S_3) [128 x 128],
//              B = F^T [16 x 128]; eight tcgen05.mma k-steps (M=128, N=16, K=128); lane i of warp w <- dF_i.
// Operands are written by the warps into NO-SWIZZLE K-major core-matrix layout (8 rows x 16 B core matrices,
// LBO = next 8-element K chunk, SBO = next 8-row group), accumulators live in TMEM.  The SIMT kernels above remain
// as the general-shape fallback.
// =================================================================================================
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  return d;                          // layout_type = 0 (SWIZZLE_NONE)
}

// INDIRECT: `emb` is the requester-side unique-row buffer urow[gs][16] (written by the owners over NVLink) and feature t of sample b
// is row inv[b][t] of it (-1 = padding -> zeros); the kernel first waits for every owner's ROWS flag (sp_sync.cuh).
template <bool INDIRECT>
__global__ void __launch_bounds__(128) k_dot_fwd_tc(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ emb,
                                                    int64_t emb_stride_t, int64_t emb_stride_b, int T, int64_t B,
                                                    __nv_bfloat16* __restrict__ Z, int64_t ldz, const int32_t* __restrict__ inv, int ldinv,
                                                    DrSpSync sync) {
  pdl_sync();
  if (INDIRECT) sp_wait_all(sync, SP_CH_ROWS);
  constexpr int D = 16;
  using emu_sh_4613005 = uint8_t[128 * 32]; emu_sh_4613005& sA = *reinterpret_cast<emu_sh_4613005*>(emu::shared_var(4613005, sizeof(emu_sh_4613005)));          // [16 row-groups][2 K-chunks][8 rows][16 B]
  using emu_sh_4613006 = __nv_bfloat16[4][512]; emu_sh_4613006& sZ = *reinterpret_cast<emu_sh_4613006*>(emu::shared_var(4613006, sizeof(emu_sh_4613006)));
  using emu_sh_4613007 = uint64_t; emu_sh_4613007& bar = *reinterpret_cast<emu_sh_4613007*>(emu::shared_var(4613007, sizeof(emu_sh_4613007)));
  using emu_sh_4613008 = uint32_t; emu_sh_4613008& tmem_slot = *reinterpret_cast<emu_sh_4613008*>(emu::shared_var(4613008, sizeof(emu_sh_4613008)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = T + 1;
  for (int i = threadIdx.x; i < 128 * 32 / 16; i += 128) reinterpret_cast<int4*>(sA)[i] = make_int4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_addr = smem_u32(sA);
  const uint64_t desc = umma_desc_noswz(a_addr, 128, 256);
  constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
  const int R = warp * 32 + lane;                                  // my row in the tile
  uint8_t* my_row = sA + (R >> 3) * 256 + (R & 7) * 16;
  const int64_t ngroups = (B + 3) / 4;
  uint32_t phase = 0;
  // software pipeline: the rows of group g+1 are fetched into registers while group g is in the tensor core / epilogue
  int4 c0 = make_int4(0, 0, 0, 0), c1 = c0;
  int32_t gsn = -1;                                   // INDIRECT: row index of my feature for the NEXT group to fetch
  auto fetch_idx = [&](int64_t gg) {
    const int64_t bb = gg * 4 + warp;
    gsn = (gg < ngroups && lane >= 1 && lane < F && bb < B) ? inv[bb * ldinv + lane - 1] : -1;
  };
  auto fetch = [&](int64_t gg, int4& r0, int4& r1) {
    const int64_t bb = gg * 4 + warp;
    if (gg < ngroups && lane < F && bb < B) {
      if (INDIRECT) {
        r0 = r1 = make_int4(0, 0, 0, 0);
        if (lane == 0) { r0 = ld_nc_v4(x + bb * ldx); r1 = ld_nc_v4(x + bb * ldx + 8); }
        else if (gsn >= 0) { const __nv_bfloat16* src = emb + (int64_t)gsn * D; r0 = ld_v4_volatile(src); r1 = ld_v4_volatile(src + 8); }
      } else {
        const __nv_bfloat16* src = lane == 0 ? x + bb * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + bb * emb_stride_b;
        r0 = ld_nc_v4(src); r1 = ld_nc_v4(src + 8);
      }
    }
    if (INDIRECT) {
      fetch_idx(gg + gridDim.x);                      // index of the group after this one: in flight for a whole iteration
    } else {
      // two more groups ahead: L2 prefetch only (one group of register loads per warp cannot cover the DRAM latency)
      const int64_t pb = (gg + 2 * (int64_t)gridDim.x) * 4 + warp;
      if (lane < F && pb < B) prefetch_l2(lane == 0 ? x + pb * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + pb * emb_stride_b);
    }
  };
  if (INDIRECT) fetch_idx(blockIdx.x);
  fetch(blockIdx.x, c0, c1);
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t b = g * 4 + warp;
    const bool live = b < B;
    if (lane < F && live) {
      *reinterpret_cast<int4*>(my_row) = c0;
      *reinterpret_cast<int4*>(my_row + 128) = c1;
      if (lane == 0) { *reinterpret_cast<int4*>(&sZ[warp][0]) = c0; *reinterpret_cast<int4*>(&sZ[warp][8]) = c1; }
    }
    fetch(g + gridDim.x, c0, c1);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      umma_bf16(tmem, desc, desc, idesc, 0u);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + warp * 32, r);
    tmem_ld_wait();
    if (live) {
      const int base = D + lane * (lane - 1) / 2;
#pragma unroll
      for (int j = 0; j < 31; ++j)
        if (j < lane && lane < F) sZ[warp][base + j] = __float2bfloat16(__uint_as_float(r[j]));
      const int used = D + F * (F - 1) / 2;
      for (int c = used + lane; c < ldz; c += 32) sZ[warp][c] = __float2bfloat16(0.f);
      __syncwarp();
      for (int c = lane * 8; c < ldz; c += 256)
        *reinterpret_cast<int4*>(Z + b * ldz + c) = *reinterpret_cast<const int4*>(&sZ[warp][c]);
    }
    tc_fence_before();
    __syncthreads();          // all TMEM reads + smem tile reads done before the next group overwrites them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// INDIRECT: features come from urow[inv[b][t]] (unique-first pipeline); the per-sample gradient rows still go to the feature-major
// demb buffer, which k_sp_segsum (sparse_pipeline.cu) pre-reduces per distinct key on the side stream.  (A first version reduced
// them here with shared-memory fp32 atomics: those compile to ATOMS.CAST.SPIN loops and made this kernel 2.6x slower -- see
// profiles/r2_notes.md.)
template <bool INDIRECT>
__global__ void __launch_bounds__(128) k_dot_bwd_tc(const __nv_bfloat16* __restrict__ dZ, int64_t ldz, const __nv_bfloat16* __restrict__ x,
                                                    int64_t ldx, const __nv_bfloat16* __restrict__ emb, int64_t emb_stride_t,
                                                    int64_t emb_stride_b, int T, int64_t B, __nv_bfloat16* __restrict__ dx, int64_t lddx,
                                                    __nv_bfloat16* __restrict__ demb, int64_t demb_stride_t, int64_t demb_stride_b,
                                                    const int32_t* __restrict__ inv, int ldinv) {
  pdl_sync();
  constexpr int D = 16;
  uint8_t* dyn = (uint8_t*)emu::dyn_smem();
  uint8_t* sA = dyn;                                   // blockdiag(S): [16 row-groups][16 K-chunks][8 rows][16 B] = 32 KB
  uint8_t* sB = dyn + 32768;                           // F^T: [2 d-groups][16 K-chunks][8 d][16 B] = 4 KB
  __nv_bfloat16* sG = reinterpret_cast<__nv_bfloat16*>(dyn + 32768 + 4096);   // [4][512] dZ rows

  using emu_sh_4613009 = uint64_t; emu_sh_4613009& bar = *reinterpret_cast<emu_sh_4613009*>(emu::shared_var(4613009, sizeof(emu_sh_4613009)));
  using emu_sh_4613010 = uint32_t; emu_sh_4613010& tmem_slot = *reinterpret_cast<emu_sh_4613010*>(emu::shared_var(4613010, sizeof(emu_sh_4613010)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = T + 1;
  for (int i = threadIdx.x; i < (32768 + 4096) / 16; i += 128) reinterpret_cast<int4*>(dyn)[i] = make_int4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint64_t adesc = umma_desc_noswz(smem_u32(sA), 128, 2048);
  const uint64_t bdesc = umma_desc_noswz(smem_u32(sB), 128, 2048);
  constexpr uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
  const int R = warp * 32 + lane;
  __nv_bfloat16* myG = sG + warp * 512;
  const int64_t ngroups = (B + 3) / 4;
  uint32_t phase = 0;
  // software pipeline: dZ chunks + the feature row of group g+1 are fetched while group g is being processed
  int4 pz0 = make_int4(0, 0, 0, 0), pz1 = pz0, pf0 = pz0, pf1 = pz0;
  int32_t gsn = -1;                                   // INDIRECT: row index of my feature for the NEXT group to fetch
  auto fetch_idx = [&](int64_t gg) {
    const int64_t bb = gg * 4 + warp;
    gsn = (gg < ngroups && lane >= 1 && lane < F && bb < B) ? inv[bb * ldinv + lane - 1] : -1;
  };
  auto fetch = [&](int64_t gg) {
    const int64_t bb = gg * 4 + warp;
    pz0 = pz1 = pf0 = pf1 = make_int4(0, 0, 0, 0);
    if (gg < ngroups && bb < B) {
      if (lane * 8 < ldz) pz0 = ld_nc_v4(dZ + bb * ldz + lane * 8);
      if (lane * 8 + 256 < ldz) pz1 = ld_nc_v4(dZ + bb * ldz + lane * 8 + 256);
      if (lane < F) {
        if (INDIRECT) {
          if (lane == 0) { pf0 = ld_nc_v4(x + bb * ldx); pf1 = ld_nc_v4(x + bb * ldx + 8); }
          else if (gsn >= 0) {
            const __nv_bfloat16* src = emb + (int64_t)gsn * D;
            pf0 = ld_nc_v4(src); pf1 = ld_nc_v4(src + 8);
          }
        } else {
          const __nv_bfloat16* src = lane == 0 ? x + bb * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + bb * emb_stride_b;
          pf0 = ld_nc_v4(src); pf1 = ld_nc_v4(src + 8);
        }
      }
    }
    if (INDIRECT) fetch_idx(gg + gridDim.x);
    // two more groups ahead: L2 prefetch only (one group of register loads per warp cannot cover the DRAM latency)
    const int64_t pb = (gg + 2 * (int64_t)gridDim.x) * 4 + warp;
    if (pb < B) {
      if (lane * 8 < ldz) prefetch_l2(dZ + pb * ldz + lane * 8);
      if (lane * 8 + 256 < ldz) prefetch_l2(dZ + pb * ldz + lane * 8 + 256);
      if (!INDIRECT && lane < F) prefetch_l2(lane == 0 ? x + pb * ldx : emb + (int64_t)(lane - 1) * emb_stride_t + pb * emb_stride_b);
    }
  };
  if (INDIRECT) fetch_idx(blockIdx.x);
  fetch(blockIdx.x);
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t b = g * 4 + warp;
    const bool live = b < B;
    // ---- stage dZ row and F^T from the prefetched registers
    if (lane * 8 < ldz) *reinterpret_cast<int4*>(myG + lane * 8) = pz0;
    if (lane * 8 + 256 < ldz) *reinterpret_cast<int4*>(myG + lane * 8 + 256) = pz1;
    if (lane < F) {
      const int4 c0 = pf0, c1 = pf1;
      const uint32_t w[8] = {(uint32_t)c0.x, (uint32_t)c0.y, (uint32_t)c0.z, (uint32_t)c0.w, (uint32_t)c1.x, (uint32_t)c1.y, (uint32_t)c1.z, (uint32_t)c1.w};
      // B[n = d][k = R] at (d/8)*2048 + (R/8)*128 + (d%8)*16 + (R%8)*2
      uint8_t* colbase = sB + (R >> 3) * 128 + (R & 7) * 2;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const uint16_t v = (d & 1) ? (uint16_t)(w[d >> 1] >> 16) : (uint16_t)(w[d >> 1] & 0xFFFF);
        *reinterpret_cast<uint16_t*>(colbase + (d >> 3) * 2048 + (d & 7) * 16) = v;
      }
    }
    fetch(g + gridDim.x);
    __syncwarp();
    // ---- my row of S (32 columns of the diagonal block): S[i][j] = dG[max][min], zero on the diagonal / padding
    {
      uint32_t packed[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        uint32_t lo = 0, hi = 0;
        const int j0 = 2 * jj, j1 = 2 * jj + 1;
        if (lane < F) {
          if (j0 < F && j0 != lane) { const int hiI = max(lane, j0), loI = min(lane, j0); lo = *reinterpret_cast<const uint16_t*>(myG + D + hiI * (hiI - 1) / 2 + loI); }
          if (j1 < F && j1 != lane) { const int hiI = max(lane, j1), loI = min(lane, j1); hi = *reinterpret_cast<const uint16_t*>(myG + D + hiI * (hiI - 1) / 2 + loI); }
        }
        packed[jj] = lo | (hi << 16);
      }
      // A[R][C = 32*warp + j] at (R/8)*2048 + (C/8)*128 + (R%8)*16 + (C%8)*2
      uint8_t* rowbase = sA + (R >> 3) * 2048 + (warp * 4) * 128 + (R & 7) * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<int4*>(rowbase + c * 128) = make_int4((int)packed[4 * c], (int)packed[4 * c + 1], (int)packed[4 * c + 2], (int)packed[4 * c + 3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)     // 16 K elements = 2 chunks = 256 B  -> +16 in the (addr >> 4) field
        umma_bf16(tmem, adesc + (uint64_t)(k * 16), bdesc + (uint64_t)(k * 16), idesc, k != 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t r[16];
    tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    if (live && lane < F) {
      float acc[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[d] = __uint_as_float(r[d]);
      if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] += __bfloat162float(myG[d]);
        store_row_bf16<16>(dx + b * lddx, acc);
      } else {
        store_row_bf16<16>(demb + (int64_t)(lane - 1) * demb_stride_t + b * demb_stride_b, acc);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

#endif  // DR_CUDA_EMU

inline int grid_for(int64_t n, int block, int max_blocks = kNumSMs * 8) {
  int64_t b = (n + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

#ifdef DR_CUDA_EMU
static int& dot_force_simt() { static int v = 1; return v; }
#else
static int& dot_force_simt() { static int v = 0; return v; }
#endif

extern "C" {

int dr_cuda_dot_set_simt(int v) { dot_force_simt() = v; return 0; }

// unique-first pipeline variants: features gathered from urow through inv (fwd), gradients pre-reduced into ugrad (bwd)
int dr_cuda_dot_interaction_fwd_u(const void* x, int64_t ldx, const void* urow, const int32_t* inv, int ldinv, int T, int D, int64_t B, void* Z,
                                  int64_t ldz, const DrSpSync* sync, cudaStream_t s) {
  if (T + 1 > 32 || ldz > 512 || ldz % 8 || D + (T + 1) * T / 2 > ldz || D != 16) return -2;
#ifdef DR_CUDA_EMU
  return -100;
#else
  int g = grid_for((B + 3) / 4, 1, kNumSMs * 4);
  DR_PDL_LAUNCH((k_dot_fwd_tc<true>), g, 128, 0, s, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)urow, 0, 0, T, B, (__nv_bfloat16*)Z, ldz, inv, ldinv, *sync);
  DR_LAUNCH_CHECK();
  return 0;
#endif
}

int dr_cuda_dot_interaction_bwd_u(const void* dZ, int64_t ldz, const void* x, int64_t ldx, const void* urow, const int32_t* inv, int ldinv,
                                  int T, int D, int64_t B, void* dx, int64_t lddx, void* demb, int64_t demb_stride_t, int64_t demb_stride_b,
                                  cudaStream_t s) {
  if (T + 1 > 32 || ldz > 512 || ldz % 8 || D != 16) return -2;
#ifdef DR_CUDA_EMU
  return -100;
#else
  constexpr int kSmem = 32768 + 4096 + 4 * 1024;
  static DrPerDeviceOnce attr_once; bool& attr = attr_once();
  if (!attr) { DR_CUDA_CHECK(cudaFuncSetAttribute(k_dot_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); attr = true; }
  int g = grid_for((B + 3) / 4, 1, kNumSMs * 4);
  DR_PDL_LAUNCH((k_dot_bwd_tc<true>), g, 128, kSmem, s, (const __nv_bfloat16*)dZ, ldz, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)urow, 0, 0, T, B,
                (__nv_bfloat16*)dx, lddx, (__nv_bfloat16*)demb, demb_stride_t, demb_stride_b, inv, ldinv);
  DR_LAUNCH_CHECK();
  return 0;
#endif
}

int dr_cuda_dot_interaction_fwd(const void* x, int64_t ldx, const void* emb, int64_t emb_stride_t, int64_t emb_stride_b, int T, int D,
                                int64_t B, void* Z, int64_t ldz, cudaStream_t s) {
  if (T + 1 > 32 || ldz > 512 || ldz % 8 || D + (T + 1) * T / 2 > ldz) return -2;
#ifndef DR_CUDA_EMU
  if (D == 16 && !dot_force_simt()) {
    int g = grid_for((B + 3) / 4, 1, kNumSMs * 4);
    DR_PDL_LAUNCH((k_dot_fwd_tc<false>), g, 128, 0, s, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B, (__nv_bfloat16*)Z, ldz,
                  (const int32_t*)nullptr, 0, DrSpSync{});
    DR_LAUNCH_CHECK();
    return 0;
  }
#endif
  int grid = grid_for((B + 7) / 8, 1, kNumSMs * 6);
  switch (D) {
    case 8: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_dot_fwd<8>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B, (__nv_bfloat16*)Z, ldz); }); break;
    case 16: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_dot_fwd<16>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B, (__nv_bfloat16*)Z, ldz); }); break;
    case 32: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_dot_fwd<32>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B, (__nv_bfloat16*)Z, ldz); }); break;
    default: return -3;
  }
  DR_LAUNCH_CHECK();
  return 0;
}

int dr_cuda_dot_interaction_bwd(const void* dZ, int64_t ldz, const void* x, int64_t ldx, const void* emb, int64_t emb_stride_t,
                                int64_t emb_stride_b, int T, int D, int64_t B, void* dx, int64_t lddx, void* demb, int64_t demb_stride_t,
                                int64_t demb_stride_b, cudaStream_t s) {
  if (T + 1 > 32 || ldz > 512 || ldz % 8) return -2;
#ifndef DR_CUDA_EMU
  if (D == 16 && !dot_force_simt()) {
    constexpr int kSmem = 32768 + 4096 + 4 * 1024;
    static DrPerDeviceOnce attr_once; bool& attr = attr_once();
    if (!attr) { DR_CUDA_CHECK(cudaFuncSetAttribute(k_dot_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); attr = true; }
    int g = grid_for((B + 3) / 4, 1, kNumSMs * 4);
    DR_PDL_LAUNCH((k_dot_bwd_tc<false>), g, 128, kSmem, s, (const __nv_bfloat16*)dZ, ldz, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B,
                                       (__nv_bfloat16*)dx, lddx, (__nv_bfloat16*)demb, demb_stride_t, demb_stride_b, (const int32_t*)nullptr, 0);
    DR_LAUNCH_CHECK();
    return 0;
  }
#endif
  int grid = grid_for((B + 7) / 8, 1, kNumSMs * 6);
#define BWD(DD) emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_dot_bwd<DD>((const __nv_bfloat16*)dZ, ldz, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)emb, emb_stride_t, emb_stride_b, T, B, (__nv_bfloat16*)dx, lddx, (__nv_bfloat16*)demb, demb_stride_t, demb_stride_b); })
  switch (D) { case 8: BWD(8); break; case 16: BWD(16); break; case 32: BWD(32); break; default: return -3; }
#undef BWD
  DR_LAUNCH_CHECK();
  return 0;
}

int dr_cuda_fm_fwd(const void* emb, int64_t st, int64_t sb, int T, int D, int64_t B, void* out, int64_t ldo, float* sum_out, cudaStream_t s) {
  int grid = grid_for(B * (D / 8), 256);
  switch (D) {
    case 8: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_fm_fwd<8>((const __nv_bfloat16*)emb, st, sb, T, B, (__nv_bfloat16*)out, ldo, sum_out); }); break;
    case 16: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_fm_fwd<16>((const __nv_bfloat16*)emb, st, sb, T, B, (__nv_bfloat16*)out, ldo, sum_out); }); break;
    case 32: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_fm_fwd<32>((const __nv_bfloat16*)emb, st, sb, T, B, (__nv_bfloat16*)out, ldo, sum_out); }); break;
    case 64: emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_fm_fwd<64>((const __nv_bfloat16*)emb, st, sb, T, B, (__nv_bfloat16*)out, ldo, sum_out); }); break;
    default: return -3;
  }
  DR_LAUNCH_CHECK();
  return 0;
}

int dr_cuda_fm_bwd(const void* dfm, int64_t ldd, const void* emb, int64_t st, int64_t sb, const float* sum_in, int T, int D, int64_t B,
                   void* demb, int64_t dst, int64_t dsb, int accumulate, cudaStream_t s) {
  int grid = grid_for(B * (D / 8), 256);
#define FMB(DD) emu::launch(dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(s), [&] { k_fm_bwd<DD>((const __nv_bfloat16*)dfm, ldd, (const __nv_bfloat16*)emb, st, sb, sum_in, T, B, (__nv_bfloat16*)demb, dst, dsb, accumulate); })
  switch (D) { case 8: FMB(8); break; case 16: FMB(16); break; case 32: FMB(32); break; case 64: FMB(64); break; default: return -3; }
#undef FMB
  DR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
