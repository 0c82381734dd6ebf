// Common device helpers for the sm_100a kernel library: PTX wrappers (mbarrier, TMA, tcgen05,
// system-scope acquire/release for NVLink peer signalling), small vector types, error macros.
#pragma once
#include <cstdlib>
#ifdef DR_CUDA_EMU
#include "emu/cuda_emu.h"      // g++ build of the SIMT kernels: every CUDA thread is a host thread (CPU CI, ASAN / TSAN over kernel code)
#else
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>

#include "../common/ev_types.h"

#define DR_CUDA_CHECK(expr)                                                                 \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      fprintf(stderr, "[deeprec_cuda] %s failed at %s:%d: %s\n", #expr, __FILE__, __LINE__, \
              cudaGetErrorString(_e));                                                      \
      return (int)_e;                                                                       \
    }                                                                                       \
  } while (0)

#define DR_LAUNCH_CHECK()                       \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) {                    \
      fprintf(stderr, "[deeprec_cuda] launch failed at %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                           \
    }                                           \
  } while (0)

namespace drc {

inline int& pdl_enabled() { static int v = [] { const char* e = getenv("DEEPREC_PDL"); return (e && e[0] == '0') ? 0 : 1; }(); return v; }

#ifdef DR_CUDA_EMU
#define DR_PDL_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (emu::launch(dim3(grid), dim3(block), (size_t)(smem), stream, [&] { kernel(__VA_ARGS__); }), cudaSuccess)
#else
template <typename... KArgs, typename... Args>
inline cudaError_t dr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DR_PDL_LAUNCH(kernel, grid, block, smem, stream, ...) drc::dr_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#endif


#ifdef DR_CUDA_EMU
constexpr int kNumSMs = 2;      // emulated blocks run one after the other: the grid-stride kernels get small grids (any grid size must be correct)
#else
constexpr int kNumSMs = 148;
#endif

// cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs (serving ProcessorGroup: one replica per GPU in one process)
// must raise the dynamic shared-memory limit on each of them.  One flag per (call site, device).
struct DrPerDeviceOnce {
  bool done[64] = {};
  bool& operator()() { int d = 0; cudaGetDevice(&d); return done[d & 63]; }
};

// Resident-block budget (per SM) of the sparse-path kernels.  They run on a side stream next to the tcgen05 GEMMs
// (1 CTA/SM, ~200 KB smem, 192 threads): capping them leaves thread slots so both streams really overlap.
inline int& sparse_blocks_per_sm() { static int v = 16; return v; }
constexpr int64_t kEmptyKey = INT64_MIN;
constexpr int64_t kTombKey = INT64_MIN + 1;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

#ifdef DR_CUDA_EMU
// the same helpers on host atomics (seq_cst where the device code uses release / acquire at system scope)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_sync() {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ int4 ld_nc_v4(const void* p) { int4 r; memcpy(&r, p, 16); return r; }
__device__ __forceinline__ int4 ld_v4_volatile(const void* p) { int4 r; memcpy(&r, p, 16); return r; }
__device__ __forceinline__ void st_na_v4(void* p, const int4& v) { memcpy(p, &v, 16); }
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  atomicAdd(p, a); atomicAdd(p + 1, b); atomicAdd(p + 2, c); atomicAdd(p + 3, d);
}
#else
// ---- system-scope signalling over NVLink (peer flags live in IPC-mapped memory) ---------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// streaming 16-byte accesses that do not pollute L1 (peer data is L2-bypassed anyway)
// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// Kernels on the training-step path are launched with cudaLaunchAttributeProgrammaticStreamSerialization (dr_launch_pdl below):
// the next kernel's CTAs may become resident and run their prologue (barrier init, TMEM alloc, descriptor prefetch, index
// math) while the previous kernel drains.  griddepcontrol.wait blocks until the predecessor grid has completed and its
// memory is visible, so NOTHING that touches global memory may precede pdl_sync() in such a kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// An early trigger lets the dependent grid's CTAs become resident while the primary still runs; measured on the DLRM step this
// steals SM slots from the primary (-4 %), so the trigger is left implicit (at grid completion) unless DEEPREC_PDL_EARLY is built in.
#ifdef DEEPREC_PDL_EARLY
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }
#else
__device__ __forceinline__ void pdl_sync() { pdl_wait(); }
#endif

// pull a line into L2 without occupying a register / scoreboard slot (software prefetch for latency-bound streaming kernels)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ int4 ld_nc_v4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_v4_volatile(const void* p) {
  int4 r;
  asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// vectorised fp32 reduction (sm_90+): one L2 atomic transaction for 4 floats
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor) -----------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tensor-map box: no shared memory, no barrier -- lets a producer run further ahead of the smem ring
__device__ __forceinline__ void tma_prefetch_2d(const void* desc, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(desc), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(desc), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 (5th-gen tensor cores, TMEM) ---------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread t gets lane base+t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#endif  // DR_CUDA_EMU

// volatile scalar loads of words that other threads update with atomics (probe-then-CAS): plain volatile on the device; relaxed atomic
// loads in the emulation build, which is what they mean -- ThreadSanitizer then reports only the races that are NOT this idiom
#ifdef DR_CUDA_EMU
__device__ __forceinline__ int64_t ld_volatile_i64(const int64_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ int64_t ld_volatile_i64(const int64_t* p) { return *reinterpret_cast<const volatile int64_t*>(p); }
__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
#endif

// idempotent metadata stores that race BY DESIGN with the volatile reads above (every writer stores the same "dirty" mark; the claim winner publishes
// its list index to readers that accept either value): plain stores on the device, relaxed atomic stores in the emulation build
#ifdef DR_CUDA_EMU
#define DR_ST_RACY(lhs, v) __atomic_store_n(&(lhs), (v), __ATOMIC_RELAXED)
#else
#define DR_ST_RACY(lhs, v) (lhs) = (v)
#endif

// 16-byte / 8-byte volatile loads of a slot's metadata words (same PTX text as before the emulation build existed: the SASS of the hot
// probe kernels is unchanged); on the host: relaxed atomic word loads, so ThreadSanitizer sees them as the benign races they are
#ifdef DR_CUDA_EMU
#define DR_LD_V4_VOLATILE(hi, p)                                                                   \
  do {                                                                                             \
    const int* _q = reinterpret_cast<const int*>(p);                                               \
    (hi).x = __atomic_load_n(_q, __ATOMIC_RELAXED); (hi).y = __atomic_load_n(_q + 1, __ATOMIC_RELAXED); \
    (hi).z = __atomic_load_n(_q + 2, __ATOMIC_RELAXED); (hi).w = __atomic_load_n(_q + 3, __ATOMIC_RELAXED); \
  } while (0)
#define DR_LD_V2_VOLATILE_U32(v, p)                                                                \
  do {                                                                                             \
    const unsigned* _q = reinterpret_cast<const unsigned*>(p);                                     \
    (v).x = __atomic_load_n(_q, __ATOMIC_RELAXED); (v).y = __atomic_load_n(_q + 1, __ATOMIC_RELAXED); \
  } while (0)
#else
#define DR_LD_V4_VOLATILE(hi, p) \
  asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"((hi).x), "=r"((hi).y), "=r"((hi).z), "=r"((hi).w) : "l"(p))
#define DR_LD_V2_VOLATILE_U32(v, p) asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"((v).x), "=r"((v).y) : "l"(p) : "memory")
#endif

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (InstrDescriptor): c=f32, a=b=bf16, K-major or MN-major operands.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace drc
