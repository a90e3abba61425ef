// Shared device helpers for the sm_100a kernels: PTX wrappers for mbarrier, TMA bulk copies,
// tcgen05 (UMMA / TMEM), proxy fences, and the TF32 rounding the reference applies
// (wmma::__float_to_tf32 == cvt.rna.tf32.f32, /root/reference TCGNN_kernel.cu:436-444).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tcgnn_b200.h"

namespace tcgnn {

// ---------------------------------------------------------------------------------------------
// Plan layout shared by plan.cu / spmm_tc.cu / sddmm_tc.cu
// ---------------------------------------------------------------------------------------------
struct __align__(16) TileMeta {  // one 16x8 TC block of the reference's SGT, 64 bytes
  int32_t cols[8];     // X row gathered for condensed column c (-1: padding, contributes zeros)
  uint32_t mask[4];    // bit (r*8+c): word r/4, bit (r%4)*8+c  <=> edge (window row r, condensed col c)
  int32_t win;         // owning row window
  int32_t edge_ofs;    // offset of the tile's first edge in tile order (exclusive scan of popcounts)
  uint32_t flags;      // bit0: first tile of its window, bit1: last tile of its window
  int32_t reserved;
};
static_assert(sizeof(TileMeta) == 64, "TileMeta must be 64 bytes");

constexpr uint32_t kTileFirst = 1u;
constexpr uint32_t kTileLast = 2u;

struct PlanView {  // device pointers, passed by value to kernels
  const TileMeta* tiles;        // [num_tiles]
  const int32_t* win_tile_ptr;  // [num_windows + 1] exclusive scan of blockPartition
  const int32_t* eperm;         // [num_pairs] tile order -> CSR edge id (lazy; weighted SpMM / SDDMM)
  const int32_t* slice_ptr;     // [grid + 1] tile range of every persistent CTA (balanced by tiles + windows)
  int32_t num_nodes;            // rows of the plan
  int32_t num_cols;             // rows of X
  int32_t row_base;             // global id of row 0 (row panels)
  int32_t num_windows;
  int32_t num_tiles;
  int32_t num_pairs;
};

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits (crt/mma.h:96-103).
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32_rna4(float4 v) {
  return make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// explicit shared-space accesses (addresses are 32-bit shared-window byte addresses)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier (by shared-window address; the uint64_t* overloads convert)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive that releases a buffer the caller has only READ: `dep` must be computed from every value loaded out of that
// buffer and can never equal 0x80000000 (OR of non-negative ids, -1 paddings and words shifted right by one).
// An mbarrier arrive is an independent shared-memory operation: it can complete while earlier ld.shared instructions
// of the warp still wait in a backed-up load/store queue (the gathers keep it full), and the buffer's next writer --
// a TMA bulk copy released by this very arrive -- then overwrites what they have yet to read.  Seen as ~1 wrong tile
// per 10^8 in SDDMM at D = 256 (profiles/r02c_*, r02d_*).  The arrive count is computed from `dep` (always 1), so the
// instruction cannot issue before the loads' results are in their registers -- a real register dependence, which
// neither the compiler nor the assembler can fold away.
__device__ __forceinline__ void mbar_arrive_after_loads(uint32_t bar, uint32_t dep) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 c;\n"
      "setp.eq.u32 p, %1, 0x80000000;\n"
      "selp.b32 c, 2, 1, p;\n"
      "mbarrier.arrive.shared::cta.b64 _, [%0], c;\n"
      "}\n" ::"r"(bar),
      "r"(dep)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread until the phase completes or a time limit passes, whichever is first; without a
// hint the limit is short and a waiting warp comes back to poll (a shared-memory access each time) every ~100
// cycles.  TCGNN_WAIT_HINT_NS > 0 passes that many nanoseconds as the suspend-time hint.
#ifndef TCGNN_WAIT_HINT_NS
#define TCGNN_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if TCGNN_WAIT_HINT_NS > 0
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(static_cast<uint32_t>(TCGNN_WAIT_HINT_NS))
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// non-blocking test (try_wait may suspend the thread for a while when the phase is not complete)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a protocol bug must surface as a launch failure, never as a hung GPU.  The limit is
// wall-clock (globaltimer), far above anything a correct pipeline waits for even when the context is time-sliced
// or runs under a debugger / sanitizer; -DTCGNN_WATCHDOG_NS=0 compiles it out.  Before trapping it says why, so
// the host sees "pipeline watchdog" next to the sticky launch failure.
#ifndef TCGNN_WATCHDOG_NS
#define TCGNN_WATCHDOG_NS 30000000000ULL   // 30 s
#endif
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void watchdog_trap(uint32_t bar, uint32_t parity) {
  printf("tcgnn: pipeline watchdog -- block %d thread %d waited > %llu ms on mbarrier 0x%x (parity %u)\n", blockIdx.x,
         threadIdx.x, static_cast<unsigned long long>(TCGNN_WATCHDOG_NS / 1000000ULL), bar, parity);
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (TCGNN_WATCHDOG_NS != 0 && (++spins & 0xFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > TCGNN_WATCHDOG_NS) watchdog_trap(bar, parity);
    }
  }
}
// Same, for waiters that expect to wait long (epilogue warps): sleep between polls so they do not
// compete for issue slots and shared-memory bandwidth with the warps on the critical path.
template <unsigned kSleepNs = 64>
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(kSleepNs);
    if (TCGNN_WATCHDOG_NS != 0 && (++spins & 0xFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > TCGNN_WATCHDOG_NS) watchdog_trap(bar, parity);
    }
  }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { mbar_init(smem_u32(bar), count); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { mbar_arrive(smem_u32(bar)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  mbar_arrive_expect_tx(smem_u32(bar), bytes);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  return mbar_try_wait(smem_u32(bar), parity);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait(smem_u32(bar), parity); }

// ---------------------------------------------------------------------------------------------
// proxy fences / TMA 1-D bulk copy (SASS: UBLKCP)
// ---------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  tma_bulk_g2s(smem_dst, gmem_src, bytes, smem_u32(bar));
}
// same, destination given as a shared-window address, with an L2 eviction policy (createpolicy)
__device__ __forceinline__ void tma_bulk_g2s_hint(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_dst),
      "l"(gmem_src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// shared -> global bulk copies (TMA engine, SASS UBLKCP / UBLKRED): they do not queue behind the SM's
// load/store pipeline.  Addresses 16-byte aligned, size a multiple of 16.  Tracked by bulk groups.
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes)
               : "memory");
}
// global[i] += shared[i] (fp32), element-wise atomic at L2
__device__ __forceinline__ void tma_bulk_s2g_add_f32(void* gmem_dst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_bulk_wait_group_read() {   // sources of all but the N newest groups have been read
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_bulk_wait_group() {        // all but the N newest groups are complete
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// L2 eviction policies: feature rows are re-read by many windows (keep), the tile stream and the
// output are touched once (let them go first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// 32-bit read-only global load, raw bits (register gathers of the TS SpMM).  volatile: the loads stay where they are
// written, i.e. all of a stage's loads are issued before the warp blocks on anything.  -DTCGNN_TS_L1_NOALLOC: do not
// keep the line in L1.
__device__ __forceinline__ uint32_t ldg_nc_u32(const char* p) {
  uint32_t v;
#ifdef TCGNN_TS_L1_NOALLOC
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
#else
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
#endif
  return v;
}

// ---------------------------------------------------------------------------------------------
// cp.async (SASS: LDGSTS): 16-byte global -> shared copies that bypass registers; completion is
// reported to an mbarrier, so the issuing warp never waits for the data itself
// ---------------------------------------------------------------------------------------------
// src_bytes < 16: the remainder of the 16 bytes is zero-filled (src_bytes == 0 reads nothing)
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_16_hint(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes,
                                                 uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(smem_dst), "l"(gmem_src),
               "r"(src_bytes), "l"(policy)
               : "memory");
}
// Feature-row gathers of the SpMM / SDDMM kernels: with or without an L2 evict_last hint (-DTCGNN_X_GATHER_HINT=1).
// The hint's policy operand costs issue slots on every LDGSTS; whether keeping X in L2 pays for it is a measured,
// per-kernel choice (profiles/r02m_ldgsts_flavour_ab.txt).
#ifndef TCGNN_X_GATHER_HINT
#define TCGNN_X_GATHER_HINT 0
#endif
__device__ __forceinline__ uint64_t x_gather_policy() {   // once per warp, outside the loops
#if TCGNN_X_GATHER_HINT
  return l2_policy_evict_last();
#else
  return 0;
#endif
}
__device__ __forceinline__ void cp_async_16_x(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes,
                                              uint64_t policy) {
#if TCGNN_X_GATHER_HINT
  cp_async_16_hint(smem_dst, gmem_src, src_bytes, policy);
#else
  (void)policy;
  cp_async_16(smem_dst, gmem_src, src_bytes);
#endif
}
// one arrival on `bar` (counted in its init count: .noinc) once all prior cp.async of this thread landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) { cp_async_mbar_arrive_noinc(smem_u32(bar)); }
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {   // at most N most recent groups still pending
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result) {  // whole warp, .sync.aligned
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: pow2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) { tmem_alloc<kCols>(smem_u32(smem_result)); }
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in tensor memory (lane = row m of A, one 32-bit column per k: 128 lanes x 8 columns for
// M128 K8 tf32; K-major only).  The operand fetch of an MMA then reads shared memory for B alone.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: thread t of the warp writes 8 consecutive columns of TMEM lane (warp%4)*32+t
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { umma_commit(smem_u32(bar)); }
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp reads TMEM lane (warp%4)*32+t.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor")
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor: start address, leading/stride byte offsets (all >>4),
// descriptor version 1 (Blackwell) at bits [46,48), swizzle mode at bits [61,64).
constexpr uint64_t kSwizzleNone = 0, kSwizzle128BBase32B = 1, kSwizzle128B = 2;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint64_t swizzle) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;   // [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;   // [32,46)
  d |= 1ull << 46;                                                // version = 1
  d |= swizzle << 61;                                             // layout type
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate.
//   [4,6) D format = 1 (f32)   [7,10) A format = 2 (tf32)   [10,13) B format = 2 (tf32)
//   bit 15 A major (1 = MN-major)   bit 16 B major (1 = MN-major)
//   [17,23) N >> 3              [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t m, uint32_t n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Byte offset of the 16-byte chunk `chunk16` (0..7) of 128-byte row `row` inside a 1024-byte
// 128B-swizzle atom whose base is 1024-byte aligned (chunk index XOR row, Swizzle<3,4,3>).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

// Same rows, but the 32-byte swizzle granule of SWIZZLE_128B_BASE32B (Swizzle<2,5,2>: address bits
// [5,7) ^= bits [7,9)), the only shared-memory layout tcgen05 accepts for MN-major tf32 operands.
// `row` is the 128-byte row inside a 512-byte aligned atom of 4 rows (8 rows = two stacked atoms).
__device__ __forceinline__ uint32_t sw128_base32_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ ((row & 3u) << 1)) << 4);
}

}  // namespace tcgnn
