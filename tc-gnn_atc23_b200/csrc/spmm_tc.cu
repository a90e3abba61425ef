// SpMM (neighbour aggregation) on 5th-gen tensor cores, sm_100a.
//
// Replaces spmm_forward_cuda_kernel / spmmAGNN_forward_cuda_kernel of the reference
// (/root/reference TCGNN_conv/TCGNN_kernel.cu:336-454, :459-578).  Same mathematics, i.e. per
// 16-row window w and per condensed 16x8 TC block t
//     Y[16w:16w+16, :] += A_t (16x8, 0/1 or edge weight, tf32) . tf32(X[cols_t, :]) (8xD)
// but mapped the other way round onto tcgen05.mma so that the gathered feature rows need no
// transpose:   Yt (D x 16)  +=  Xg^T (D x 8, "MN-major": every gathered row is contiguous along M)
//                               . A_t^T (8 x 16, K-major)
//   M = 128 features per MMA (DBLK blocks of 128 per pass), N = 16 window rows, K = 8 neighbours.
// The accumulator lives in TMEM (lane = feature, column = window row); nothing is rescanned:
// the plan (plan.cu) delivers, per tile, the 8 rows to gather and a 128-bit occupancy mask.
//
// Two kernels per call:
//   1. tf32_round_pack_kernel (round_pack.cu): Xr = cvt.rna.tf32(X) once per element (the reference
//      rounds inside its inner loop, TCGNN_kernel.cu:441-443; tcgen05 would otherwise truncate),
//      packed with a 16-byte aligned leading dimension so every gather below is a 128-bit access;
//   2. spmm_tc_kernel, CTA = 12 warps, persistent over a contiguous, equally sized slice of the
//      global tile stream (balances hub windows: a window cut by a slice boundary is combined with
//      fp32 atomics):
//   warps 0-3   epilogue     TMEM -> registers -> coalesced global stores (lane = feature)
//   warp  4     MMA issuer   warp-uniform loop, one elected lane issues tcgen05.mma / tcgen05.commit; per
//                            stage it reads ONE word (which tiles open / close a window); a full stage
//                            inside one window is G back-to-back MMAs with precomputed descriptors
//   warp  5     meta loader  TMA bulk copies (UBLKCP, L2 evict_first) of the tile records into a 16-deep
//                            ring that runs far ahead of the data stages
//   warps 6-11  producers    warp p OWNS the stages k = p (mod 6): it gathers all G tiles of the stage with
//                            asynchronous 128-bit copies (cp.async / LDGSTS, zero-fill for padding, no
//                            cache hint) straight into the swizzled (128B rows, 32B granule) MN-major A
//                            tiles, expands the occupancy masks (or edge weights) into the K-major B tiles,
//                            writes the stage's open/close word, and publishes the stage with ONE mbarrier
//                            arrive once cp.async.wait_group says its copies have landed.
// History of this pipeline (profiles/r01b_ablations.txt, r01c_ablations.txt, r01d_*.txt): per-thread
// cp.async.mbarrier arrivals (129 per stage) serialise in the LSU: 160 ns per tile.  Four producer warps that
// all take part in every stage, publishing with a lag: 66 ns per tile, bound by the ~600-cycle serial chain
// (barrier waits, fences, address arithmetic) each warp walks per stage.  Stage ownership puts six such chains
// in parallel and needs 1 arrival per stage.
// Pipeline: S data stages of G tiles (A + B), three mbarrier rings (meta_full/meta_empty, full/empty);
// NACC TMEM accumulators with acc_full / acc_empty so the epilogue overlaps the next windows.
// All shared-memory metadata reads are explicit ld.shared (the generic-address loads the compiler
// emits for a runtime-aligned dynamic smem base cost the MMA thread ~300 cycles per tile).
#include <stdio.h>
#include <stdlib.h>

#include "plan.h"

namespace tcgnn {

namespace {

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kMetaWarp = 5;
constexpr int kProducerWarp0 = 6;     // producer warp p owns the stages p, p + P, ...
constexpr int kAcc = 4;               // TMEM accumulator ring
constexpr int kBTileBytes = TCGNN_BLK_H * TCGNN_BLK_W * 4;  // 512
// Profiling switches are compiled in only with -DTCGNN_DEBUG_SWITCHES (build.py --debug-switches): the production
// library ignores TCGNN_ABLATE / TCGNN_TUNE / TCGNN_TRACE / TCGNN_PRESET, and the ablation branches fold away.
#ifdef TCGNN_DEBUG_SWITCHES
constexpr bool kDebugSwitches = true;
#else
constexpr bool kDebugSwitches = false;
#endif
// ablation switches (env TCGNN_ABLATE, profiling builds only -- results are wrong when set): skip the row
// gathers / the MMAs after a window's first / the B-tile construction / all but one MMA of a full stage
constexpr uint32_t kAblateGather = 1u, kAblateMma = 2u, kAblateBuild = 4u, kAblateMmaFast = 8u;
constexpr uint32_t kAblateOutput = 256u;   // no bulk output stores
// L2 policy switches (env TCGNN_TUNE overrides the default kTuneDefault): tile stream evict_first / output rows
// written with streaming stores.  The feature-row gathers carry NO cache hint: an evict_last policy operand on every
// LDGSTS cost 3.5 % (reddit-like R-MAT 2.81 -> 2.72 ms, uniform 4.76 -> 4.59 ms; profiles/r02m_ldgsts_flavour_ab.txt)
// and the zero-fill size operand nothing.
constexpr uint32_t kTuneMetaFirst = 32u, kTuneYStream = 64u;
constexpr uint32_t kTuneDefault = kTuneMetaFirst | kTuneYStream;
// op mode (not a debug switch): Y += A X -- every window is combined with the bulk reduce-add / fp32 atomics and
// nothing is cleared first (TCGNN_ACCUMULATE: per-source-panel partial products of the sharded path)
constexpr uint32_t kFlagAccumulate = 1u << 24;

// DBLK: 128-feature blocks per pass; G: tiles per pipeline stage; S: data stages; P: producer TEAMS (a team owns
// whole stages); L: own stages a team keeps in flight before it publishes the oldest
// STAGED: the epilogue stages a window's output in shared memory and writes it with one bulk copy
// TEAM: warps per team.  Measured with tools/l2_gather_bench (profiles/r02a_*): one warp sustains only ~6-9 bytes per
// clock of LDGSTS row gathers however many tiles it keeps in flight (8-10 outstanding 512-byte instructions), so six
// gathering warps cap an SM at ~38-54 B/clk = 5600-8000 B/clk for the chip, while twelve reach ~10,800 and sixteen
// ~12,400.  Two warps per stage double the gather issue rate without touching the per-stage barrier protocol.
// TS: the gathered rows never touch shared memory.  The TEAM = 4 warps of a team sit on the four TMEM lane quadrants;
// warp q loads features [32q, 32q + 32) of all G x 8 gathered rows of a stage into REGISTERS (coalesced 128-byte
// loads, 64 in flight per thread), writes them to tensor memory with tcgen05.st (lane = feature, one column per
// neighbour) and the MMA takes its A operand from there, so no gathered byte crosses shared memory.  Measured
// (profiles/r02m_spmm_ts_timings.txt, r02m_trace_spmm_ts_*): the layout works (tests/test_gpu_umma_layouts.py), an
// MMA with A in TMEM costs 55 cycles instead of 39, and the kernel is 1.9x SLOWER than the shared-memory pipeline:
// a warp-wide 4-byte load moves 128 bytes where an LDGSTS moves 512, every one of the four quadrant warps redoes the
// address arithmetic of all 64 rows of a stage, and a warp needs ~36 cycles per memory instruction either way
// (2300 cycles to issue a stage).  Features must lie along the TMEM lanes, so wider loads would need a transpose
// across warps.  Compiled in profiling builds only (TCGNN_SPMM_TS=1).
template <int DBLK, int G, int S_, int P, int L, bool STAGED, int TEAM = 2, bool TS = false>
struct Cfg {
  static constexpr bool kTs = TS;
  static constexpr int kTeam = TEAM;
  static constexpr int kTilesPerMember = G / TEAM;
  static constexpr bool kStaged = STAGED;
  static constexpr int kG = G;
  static constexpr int kATileBytes = DBLK * 4096;              // DBLK*4 swizzle atoms of 8 rows x 128 B
  static constexpr int kAStageBytes = TS ? 0 : kG * kATileBytes;
  static constexpr uint32_t kAccCols = kAcc * DBLK * 16;       // TMEM: accumulator ring (64 / 128 columns) ...
  static constexpr uint32_t kAStageCols = G * DBLK * 8;        // ... then (TS) the A ring: 8 columns per tile and block
  static constexpr int kBStageBytes = kG * kBTileBytes;
  static constexpr int kMetaStageBytes = kG * static_cast<int>(sizeof(TileMeta));
  static constexpr int kStages = S_;                           // data ring (A + B tiles)
  static constexpr int kProducers = P;
  static constexpr int kOwnLag = L;
  static constexpr int kThreads = (kProducerWarp0 + P * TEAM) * 32;
  static constexpr int kMetaStages = 16;                       // tile-record ring, prefetched far ahead of the data
  static constexpr uint32_t kTmemCols = TS ? 512u : kAccCols;  // a power of two
  static constexpr int kBarBytes = (2 * kMetaStages + 2 * kStages + 2 * kAcc) * 8;
  static constexpr int kYStageBytes = STAGED ? TCGNN_BLK_H * DBLK * 128 * 4 : 0;   // one window of output (8 / 16 KB)
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + 2 * kYStageBytes +
                                    kMetaStages * kMetaStageBytes + kBarBytes + kStages * 4 /*info words*/ + 16 +
                                    1024 /*alignment slack*/;
  static_assert(kG >= 1 && kG <= 8, "the open/close word holds 8 tile bits");
  static_assert(G % TEAM == 0, "the members of a team gather equal shares of a stage");
  static_assert(!TS || (TEAM == 4 && L == 0 && kAccCols + S_ * kAStageCols <= 512),
                "TS: one warp per TMEM lane quadrant, eager publication, accumulators + A ring within 512 columns");
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  static_assert(kMetaStages >= (L + 1) * P || kMetaStages >= 16, "records of every in-flight own stage stay resident");
  // L == 0: a team publishes the stage it has just filled as soon as its copies have landed (it sits in
  // cp.async.wait_group for the tail of the flight and builds the next stage's B values afterwards, while the MMAs
  // consume this one); L >= 1: the stage is published at the top of the L-th following own iteration.
  static_assert(L >= 0 && (L - 1) * P < S_, "a warp may not wait for the slot of an own stage it has not published yet");
};

struct SliceInfo {
  int32_t t0, t1;        // tile range of this CTA
  int32_t w_first;       // window of tile t0
  int32_t n_windows;     // windows touched
  bool partial_first;    // first window starts before t0  -> atomics
  bool partial_last;     // last window ends after t1      -> atomics
};

__device__ __forceinline__ SliceInfo slice_info(const PlanView& pv) {
  SliceInfo s;
  s.t0 = pv.slice_ptr[blockIdx.x];   // the kernels are launched with plan->grid CTAs
  s.t1 = pv.slice_ptr[blockIdx.x + 1];
  s.w_first = 0;
  s.n_windows = 0;
  s.partial_first = s.partial_last = false;
  if (s.t1 > s.t0) {
    const TileMeta* a = pv.tiles + s.t0;
    const TileMeta* b = pv.tiles + (s.t1 - 1);
    s.w_first = a->win;
    s.n_windows = b->win - a->win + 1;
    s.partial_first = (a->flags & kTileFirst) == 0;
    s.partial_last = (b->flags & kTileLast) == 0;
  }
  return s;
}

// Rows of windows cut by a slice boundary are accumulated with atomics: clear them first.
__global__ void spmm_zero_partial_rows(PlanView pv, float* __restrict__ y, int64_t ldy, int32_t dim) {
  const SliceInfo s = slice_info(pv);
  if (s.t1 <= s.t0) return;
  for (int side = 0; side < 2; ++side) {
    const bool partial = side == 0 ? s.partial_first : s.partial_last;
    if (!partial) continue;
    if (side == 1 && s.n_windows == 1 && s.partial_first) continue;  // same window, already cleared
    const int32_t w = side == 0 ? s.w_first : s.w_first + s.n_windows - 1;
    const int32_t row0 = w * TCGNN_BLK_H;
    const int32_t rows = min(TCGNN_BLK_H, pv.num_nodes - row0);
    for (int i = threadIdx.x; i < rows * dim; i += blockDim.x) y[(row0 + i / dim) * ldy + i % dim] = 0.0f;
  }
}

// Edge weights (CSR order) -> tile order, rounded to tf32 like the reference (TCGNN_kernel.cu:560-563).
__global__ void permute_weights_kernel(const int32_t* __restrict__ eperm, const float* __restrict__ w,
                                       float* __restrict__ out, int32_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = tf32_rna(w[eperm[i]]);
}

// Timeline probe (env TCGNN_TRACE=1, tools/trace.py): block 0 records clock64 at fixed points of the first
// kTraceStages stages -- role 0 = MMA warp, 1 = producer warp 0, 2 = meta loader; 8 points per stage.
constexpr int kTraceStages = 512;
constexpr int kTraceCtas = 160;             // + per-CTA {cycles, tiles, windows, 0}
constexpr int kTraceWords = 3 * kTraceStages * 8 + kTraceCtas * 4 + kTraceStages * 8;   // + epilogue warp 0, per window
__device__ __forceinline__ void trace_put(long long* trace, int role, int32_t k, int point) {
  if (k < kTraceStages && (threadIdx.x & 31) == 0) trace[(role * kTraceStages + k) * 8 + point] = clock64();
}

// TS gathers of one tile: element [m][r] = feature block m of gathered row r at this thread's feature, straight into
// registers.  A tile without padding (all but a window's last) takes the branch-free path: one IMAD.WIDE and one load
// per element.  Returns the OR of the record's column ids (see mbar_arrive_after_loads).
template <int DBLK>
__device__ __forceinline__ uint32_t ts_gather_tile(uint32_t record, const char* const (&base)[DBLK],
                                                   const uint32_t (&row_bytes)[DBLK], uint32_t (&v)[DBLK][8]) {
  const int4 c0 = lds_v4(record), c1 = lds_v4(record + 16);   // rows to gather (-1: padding)
  const int32_t cols[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const uint32_t all = static_cast<uint32_t>(c0.x | c0.y | c0.z | c0.w | c1.x | c1.y | c1.z | c1.w);
  if (static_cast<int32_t>(all) >= 0) {   // warp-uniform
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int m = 0; m < DBLK; ++m)
        v[m][r] = ldg_nc_u32(base[m] + static_cast<uint64_t>(static_cast<uint32_t>(cols[r])) * row_bytes[m]);
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int m = 0; m < DBLK; ++m) {
        v[m][r] = 0u;
        if (cols[r] >= 0)
          v[m][r] = ldg_nc_u32(base[m] + static_cast<uint64_t>(static_cast<uint32_t>(cols[r])) * row_bytes[m]);
      }
  }
  return all;
}

template <class C, int DBLK>
__global__ void __launch_bounds__(C::kThreads, 1)
spmm_tc_kernel(PlanView pv, const float* __restrict__ x /* tf32-rounded, 16B aligned */, int64_t ldx /* % 4 == 0 */,
               const float* __restrict__ wperm, float* __restrict__ y, int64_t ldy,
               int32_t dim /* <= DBLK*128, features of this pass */, uint32_t flags /* kAblate* | kTune* bits */,
               long long* __restrict__ trace /* nullable */) {
  constexpr int kG = C::kG;
  constexpr int S = C::kStages;
  constexpr int MS = C::kMetaStages;
  constexpr int kProducers = C::kProducers;
  constexpr int OWN_LAG = C::kOwnLag;
  extern __shared__ uint8_t smem_raw[];
  // shared-space byte addresses (ld.shared / st.shared / descriptors / bulk copies all take these)
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem;                                       // [S][G][DBLK*4 atoms][8][128B]
  const uint32_t b_smem = a_smem + S * C::kAStageBytes;               // [S][G][512B]
  const uint32_t y_smem = b_smem + S * C::kBStageBytes;               // [2][16][dim] output staging (bulk stores)
  const uint32_t m_smem = y_smem + 2 * C::kYStageBytes;               // [MS][G] TileMeta
  const uint32_t bars = m_smem + MS * C::kMetaStageBytes;
  const uint32_t meta_full = bars, meta_empty = bars + 8 * MS;
  const uint32_t full = bars + 16 * MS, empty = full + 8 * S;
  const uint32_t acc_full = empty + 8 * S, acc_empty = acc_full + 8 * kAcc;
  const uint32_t info_smem = acc_empty + 8 * kAcc;                    // [S] open/close word per stage
  const uint32_t tmem_slot = info_smem + 4 * S;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const SliceInfo sl = slice_info(pv);
  const int32_t n_tiles = sl.t1 - sl.t0;
  const int32_t n_stages = (n_tiles + kG - 1) / kG;
  const bool tr = kDebugSwitches && trace != nullptr && blockIdx.x == ((flags >> 16) & 0xFFu);   // env TCGNN_TRACE_CTA
  const long long t_start = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < MS; ++s) {
      mbar_init(meta_full + 8 * s, 1);                // expect_tx arrive of the meta loader + the copy's bytes
      mbar_init(meta_empty + 8 * s, C::kTeam);        // the stage's producer warps are done with the records
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full + 8 * s, C::kTeam);              // the stage's producer warps: copies landed, B tiles written
      mbar_init(empty + 8 * s, 1);                    // tcgen05.commit
    }
    for (int b = 0; b < kAcc; ++b) {
      mbar_init(acc_full + 8 * b, 1);                 // tcgen05.commit
      mbar_init(acc_empty + 8 * b, kEpiWarps);        // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds_u32(tmem_slot);

  if (warp < kEpiWarps) {
    // ===================================== epilogue =====================================
    // TMEM -> registers -> row-major staging tile in shared memory -> one bulk copy per output row (TMA engine).
    // Plain st.global from here queue behind the producers' gathers in the SM's load/store pipeline: ~120
    // cycles per store instruction, 3800 cycles per window at D = 256 (profiles/r01e_*trace*).  Windows cut by
    // a slice boundary use the bulk reduce-add instead of fp32 atomics.
    const int q = warp;  // TMEM lane quadrant == warp id % 4
    const bool stream_y = (flags & kTuneYStream) != 0;
    // one bulk copy per window needs a contiguous, 16-byte aligned output panel
    const bool bulk = C::kStaged && ldy == dim && (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    const bool issuer = threadIdx.x == 0;
    const uint32_t row_bytes = static_cast<uint32_t>(dim) * 4u;
    for (int32_t wl = 0; wl < sl.n_windows; ++wl) {
      const int b = wl % kAcc;
      const bool tre = tr && q == 0 && wl < kTraceStages;
      long long* trow = trace + 3 * kTraceStages * 8 + kTraceCtas * 4 + wl * 8;
      if (tre && lane == 0) trow[0] = clock64();
      mbar_wait_backoff(acc_full + 8 * b, (wl / kAcc) & 1);
      tc_fence_after();
      if (tre && lane == 0) trow[1] = clock64();
      const int32_t w = sl.w_first + wl;
      const bool use_atomic = (flags & kFlagAccumulate) != 0 || (wl == 0 && sl.partial_first) ||
                              (wl == sl.n_windows - 1 && sl.partial_last);
      const int32_t row0 = w * TCGNN_BLK_H;
      const uint32_t ybuf = y_smem + (wl & 1) * C::kYStageBytes;
#pragma unroll
      for (int m = 0; m < DBLK; ++m) {
        const int f = m * 128 + q * 32 + lane;       // feature owned by this thread
        if (m * 128 + q * 32 < dim) {                // warp-uniform
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (b * DBLK + m) * 16, v);
          tmem_ld_wait();
          if (tre && lane == 0) trow[2 + m] = clock64();
          if (f < dim) {
            if (bulk) {
#pragma unroll
              for (int i = 0; i < 16; ++i) sts_f32(ybuf + i * row_bytes + f * 4, __uint_as_float(v[i]));
            } else {
              float* yp = y + static_cast<int64_t>(row0) * ldy + f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (row0 + i < pv.num_nodes) {
                  if (use_atomic) atomicAdd(yp + i * ldy, __uint_as_float(v[i]));
                  else if (stream_y) __stcs(yp + i * ldy, __uint_as_float(v[i]));
                  else yp[i * ldy] = __uint_as_float(v[i]);
                }
              }
            }
          }
        }
      }
      if (tre && lane == 0) trow[4] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * b);   // the accumulator is in registers / shared memory now
      if (bulk) {
        fence_proxy_async_smem();                      // staging writes -> visible to the bulk copies
        // the copies of window wl-1 have read the other buffer: after the barrier everybody may overwrite it
        if (issuer) tma_bulk_wait_group_read<0>();
        named_barrier_sync(1, kEpiWarps * 32);
        if (issuer && !(kDebugSwitches && (flags & kAblateOutput))) {
          const int rows = min(TCGNN_BLK_H, pv.num_nodes - row0);
          float* yp = y + static_cast<int64_t>(row0) * ldy;
          // contiguous output: the window's rows are one block (a bulk copy costs ~200 cycles whatever its size)
          if (use_atomic) tma_bulk_s2g_add_f32(yp, ybuf, rows * row_bytes);
          else tma_bulk_s2g(yp, ybuf, rows * row_bytes);
          tma_bulk_commit_group();
        }
      }
      if (tre && lane == 0) trow[5] = clock64();
    }
    if (bulk && issuer) tma_bulk_wait_group<0>();      // shared memory must outlive the copies
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer ===================================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses and descriptors in uniform
    // registers); one elected lane issues tcgen05.mma / tcgen05.commit.  Per tile: one descriptor add, one MMA.
    constexpr uint32_t idesc = make_idesc_tf32(128, 16, /*A MN-major (TMEM: K-major)*/ !C::kTs, /*B K-major*/ false);
    // A: MN-major tf32 -> SWIZZLE_128B_BASE32B atoms of 32 features x 4 k-rows (4 x 128 B): feature blocks
    // 1024 B apart (LBO), the two k-halves of a K=8 MMA 512 B apart (SBO)
    const uint64_t adesc0 = make_smem_desc(0, 1024, 512, kSwizzle128BBase32B);
    // B: K-major, no swizzle: 8x16B core matrices; K chunks 128 B apart (LBO), 8-row groups 256 B apart (SBO)
    const uint64_t bdesc0 = make_smem_desc(0, 128, 256, kSwizzleNone);
    const bool skip_mma = kDebugSwitches && (flags & kAblateMma) != 0;
    int32_t wl = 0;        // windows opened so far
    uint32_t acc = tmem_base;
    int b = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int32_t k = 0; k < n_stages; ++k) {
      mbar_wait(full + 8 * s, ph);   // the producer fenced its generic-proxy writes before arriving
      tc_fence_after();
      if (tr) trace_put(trace, 0, k, 0);
      // bits [0,G): tile opens a window, [8,8+G): closes, [16,..): tiles.  The shuffle tells the compiler the
      // word is warp-uniform, so the branches below stay uniform.
      const uint32_t info = __shfl_sync(0xffffffffu, lds_u32(info_smem + 4 * s), 0);
      const int nt = static_cast<int>(info >> 16);
      const uint64_t adesc_s = adesc0 | static_cast<uint64_t>(((a_smem + s * C::kAStageBytes) & 0x3FFFFu) >> 4);
      const uint64_t bdesc_s = bdesc0 | static_cast<uint64_t>(((b_smem + s * C::kBStageBytes) & 0x3FFFFu) >> 4);
      const uint32_t a_tmem_s = tmem_base + C::kAccCols + s * C::kAStageCols;   // TS: lane 0, first column of the stage
      if (tr) trace_put(trace, 0, k, 1);
      if (info == (static_cast<uint32_t>(kG) << 16) && !skip_mma) {
        // common case in dense windows: a full stage strictly inside one window -> G back-to-back MMAs
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < kG; ++j) {
            if (kDebugSwitches && j > 0 && (flags & kAblateMmaFast)) break;
#pragma unroll
            for (int m = 0; m < DBLK; ++m) {
              if constexpr (C::kTs)
                umma_tf32_ts(acc + m * 16, a_tmem_s + (j * DBLK + m) * 8,
                             bdesc_s + static_cast<uint64_t>((j * kBTileBytes) >> 4), idesc, 1u);
              else
                umma_tf32(acc + m * 16, adesc_s + static_cast<uint64_t>((j * C::kATileBytes + m * 4096) >> 4),
                          bdesc_s + static_cast<uint64_t>((j * kBTileBytes) >> 4), idesc, 1u);
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          if (j < nt) {
            const bool first = (info >> j) & 1u;
            const bool last = (info >> (8 + j)) & 1u;
            if (first) {
              b = wl % kAcc;
              acc = tmem_base + b * DBLK * 16;
              mbar_wait(acc_empty + 8 * b, ((wl / kAcc) & 1) ^ 1);
              tc_fence_after();
            }
            if (elect_one()) {
              if (!skip_mma || first) {
#pragma unroll
                for (int m = 0; m < DBLK; ++m) {
                  if constexpr (C::kTs)
                    umma_tf32_ts(acc + m * 16, a_tmem_s + (j * DBLK + m) * 8,
                                 bdesc_s + static_cast<uint64_t>((j * kBTileBytes) >> 4), idesc, first ? 0u : 1u);
                  else
                    umma_tf32(acc + m * 16, adesc_s + static_cast<uint64_t>((j * C::kATileBytes + m * 4096) >> 4),
                              bdesc_s + static_cast<uint64_t>((j * kBTileBytes) >> 4), idesc, first ? 0u : 1u);
                }
              }
              if (last) umma_commit(acc_full + 8 * b);
            }
            if (last) ++wl;
          }
        }
      }
      if (tr) trace_put(trace, 0, k, 2);
      if (elect_one()) umma_commit(empty + 8 * s);
      if (tr) trace_put(trace, 0, k, 3);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    // the last commit must land in this CTA's shared memory before the CTA may retire
    if (n_stages > 0) mbar_wait(empty + 8 * ((n_stages - 1) % S), ((n_stages - 1) / S) & 1);
  } else if (warp == kMetaWarp) {
    // ===================================== meta loader (TMA) ============================
    if (lane == 0) {
      const uint64_t policy = (flags & kTuneMetaFirst) ? l2_policy_evict_first() : l2_policy_evict_normal();
      int ms = 0;
      uint32_t mph = 0;
      for (int32_t k = 0; k < n_stages; ++k) {
        if (tr) trace_put(trace, 2, k, 0);
        mbar_wait(meta_empty + 8 * ms, mph ^ 1u);
        if (tr) trace_put(trace, 2, k, 1);
        const int32_t g0 = sl.t0 + k * kG;
        const uint32_t bytes = static_cast<uint32_t>(min(kG, sl.t1 - g0)) * sizeof(TileMeta);
        mbar_arrive_expect_tx(meta_full + 8 * ms, bytes);
        tma_bulk_g2s_hint(m_smem + ms * C::kMetaStageBytes, pv.tiles + g0, bytes, meta_full + 8 * ms, policy);
        if (tr) trace_put(trace, 2, k, 2);
        if (++ms == MS) { ms = 0; mph ^= 1u; }
      }
    }
  } else {
    // ===================================== producers ====================================
    // team `p` owns the stages p, p + P, ...; its member `member` gathers / builds the tiles [j_lo, j_lo + kTpm)
    constexpr int kTpm = C::kTilesPerMember;
    const int p = (warp - kProducerWarp0) / C::kTeam;
    const int member = (warp - kProducerWarp0) % C::kTeam;
    const int j_lo = member * kTpm;
    const int nvec = (dim + 3) >> 2;          // 16-byte vectors per feature row in this pass
    constexpr int kVecPerLane = DBLK * 8;     // 8 rows x DBLK*32 vectors / 32 lanes
    const bool skip_gather = kDebugSwitches && (flags & kAblateGather) != 0;
    const bool skip_build = kDebugSwitches && (flags & kAblateBuild) != 0;
    const uint64_t policy = x_gather_policy();
    // my 16-byte vector of a full-width row; the empty asm keeps the sum in ONE register pair, so a row address is a
    // single IMAD.WIDE (the compiler otherwise re-adds the kernel parameter after every multiply)
    uint64_t x_lane_bits = reinterpret_cast<uint64_t>(x) + static_cast<uint64_t>(lane) * 16u;
    asm volatile("" : "+l"(x_lane_bits));
    const char* x_lane = reinterpret_cast<const char*>(x_lane_bits);
    const uint32_t row_bytes = static_cast<uint32_t>(ldx) * 4u;
    // B tile: lane -> 16-byte chunk `lane` of the tile: [n/8][k/4][n%8] x 4 floats (k%4)
    const int bn = (lane >> 4) * 8 + (lane & 7);
    const int bword = bn >> 2;
    const int bshift = (bn & 3) * 8 + ((lane >> 3) & 1) * 4;
    int32_t published = p;                    // oldest own stage not yet published
    // TS: my feature of block m is m * 128 + 32 * (warp % 4) + lane == my TMEM lane; a lane past `dim` reads x[0]
    // over and over (stride 0) -- its accumulator rows are never stored
    const char* ts_base[DBLK];
    uint32_t ts_row_bytes[DBLK];
#pragma unroll
    for (int m = 0; m < DBLK; ++m) {
      const int f = m * 128 + (warp & 3) * 32 + lane;
      ts_base[m] = reinterpret_cast<const char*>(x) + (f < dim ? f * 4 : 0);
      ts_row_bytes[m] = f < dim ? static_cast<uint32_t>(ldx) * 4u : 0u;
    }
    for (int32_t k = p; k < n_stages; k += kProducers) {
      const int s = k % S;
      const int ms = k % MS;
      const int32_t g0 = sl.t0 + k * kG;
      const int nt = min(kG, sl.t1 - g0);
      const bool trp = tr && p == 0 && member == 0;
      if (trp) trace_put(trace, 1, k / kProducers, 0);
      mbar_wait(meta_full + 8 * ms, (k / MS) & 1);
      if (trp) trace_put(trace, 1, k / kProducers, 1);
      const uint32_t meta = m_smem + ms * C::kMetaStageBytes;
      // B values of the stage
      float4 bv[kTpm];
      uint32_t ring_dep = 0u;   // every word read from the record ring, see mbar_arrive_after_loads
      if (wperm == nullptr) {
        // pattern: my four cells are four bits of one mask word (expanding them through a 16-entry float4 table in
        // shared memory instead of the selects was 2 % slower: profiles/r02o_timings_final.txt)
#pragma unroll
        for (int jj = 0; jj < kTpm; ++jj) {
          const int j = j_lo + jj;
          const uint32_t mw = (j < nt && !skip_build) ? lds_u32(meta + j * 64 + 32 + bword * 4) : 0u;
          ring_dep |= mw >> 1;
          const uint32_t nib = (mw >> bshift) & 0xFu;
          bv[jj] = make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f,
                               (nib & 8u) ? 1.0f : 0.0f);
        }
      } else {
        // weighted: a tile's weights are one contiguous run of the tile-ordered array (mask-bit order).  One
        // coalesced load per tile (all issued before any is used), then every lane picks its <= 4 values by
        // rank with shuffles -- global load instructions are expensive next to the gathers.
        uint32_t nibs[kTpm];
        int ranks[kTpm];
        int counts[kTpm];
        float wv[kTpm];
        const float* runs[kTpm];
#pragma unroll
        for (int jj = 0; jj < kTpm; ++jj) {
          const int j = j_lo + jj;
          nibs[jj] = 0u;
          ranks[jj] = 0;
          counts[jj] = 0;
          wv[jj] = 0.0f;
          runs[jj] = wperm;
          if (j < nt && !skip_build) {
            const int4 m4 = lds_v4(meta + j * 64 + 32);
            const uint32_t w0 = static_cast<uint32_t>(m4.x), w1 = static_cast<uint32_t>(m4.y),
                           w2 = static_cast<uint32_t>(m4.z), w3 = static_cast<uint32_t>(m4.w);
            const uint32_t mine = bword == 0 ? w0 : (bword == 1 ? w1 : (bword == 2 ? w2 : w3));
            ring_dep |= (w0 | w1 | w2 | w3) >> 1;
            nibs[jj] = (mine >> bshift) & 0xFu;
            // rank of the first of my four bits among the tile's set bits (bit order r*8+c)
            ranks[jj] = __popc(mine & ((1u << bshift) - 1u)) + (bword > 0 ? __popc(w0) : 0) +
                       (bword > 1 ? __popc(w1) : 0) + (bword > 2 ? __popc(w2) : 0);
            counts[jj] = __popc(w0) + __popc(w1) + __popc(w2) + __popc(w3);
            runs[jj] = wperm + static_cast<int32_t>(lds_u32(meta + j * 64 + 52));
            if (counts[jj] <= 32 && lane < counts[jj]) wv[jj] = __ldg(runs[jj] + lane);
          }
        }
#pragma unroll
        for (int j = 0; j < kTpm; ++j) {
          const uint32_t nib = nibs[j];
          const int r0 = ranks[j], r1 = r0 + (nib & 1u), r2 = r1 + ((nib >> 1) & 1u), r3 = r2 + ((nib >> 2) & 1u);
          if (counts[j] <= 32) {   // warp-uniform
            const float x0 = __shfl_sync(0xffffffffu, wv[j], r0 & 31), x1 = __shfl_sync(0xffffffffu, wv[j], r1 & 31),
                        x2 = __shfl_sync(0xffffffffu, wv[j], r2 & 31), x3 = __shfl_sync(0xffffffffu, wv[j], r3 & 31);
            bv[j] = make_float4((nib & 1u) ? x0 : 0.0f, (nib & 2u) ? x1 : 0.0f, (nib & 4u) ? x2 : 0.0f,
                                (nib & 8u) ? x3 : 0.0f);
          } else {                 // dense tile (> 32 of 128 cells): direct loads
            bv[j] = make_float4((nib & 1u) ? __ldg(runs[j] + r0) : 0.0f, (nib & 2u) ? __ldg(runs[j] + r1) : 0.0f,
                                (nib & 4u) ? __ldg(runs[j] + r2) : 0.0f, (nib & 8u) ? __ldg(runs[j] + r3) : 0.0f);
          }
        }
      }
      // open/close word for the MMA warp: lane j looks at tile j
      bool first = false, last = false;
      if (lane < nt) {
        const uint32_t tf = lds_u32(meta + lane * 64 + 56);
        first = (tf & kTileFirst) != 0 || (g0 + lane == sl.t0);
        last = (tf & kTileLast) != 0 || (g0 + lane == sl.t1 - 1);
      }
      const uint32_t fm = __ballot_sync(0xffffffffu, first), lm = __ballot_sync(0xffffffffu, last);
      ring_dep |= fm | lm;
      if constexpr (C::kTs) {
        // ---- gather into registers: all loads of the stage are in flight before the slot is even free ----
        if (trp) trace_put(trace, 1, k / kProducers, 2);
        uint32_t v[kG][DBLK][8];
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          if (j < nt && !skip_gather) {
            ring_dep |= ts_gather_tile<DBLK>(meta + j * 64, ts_base, ts_row_bytes, v[j]);
          } else {
#pragma unroll
            for (int m = 0; m < DBLK; ++m)
#pragma unroll
              for (int r = 0; r < 8; ++r) v[j][m][r] = 0u;
          }
        }
        if (trp) trace_put(trace, 1, k / kProducers, 3);
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(meta_empty + 8 * ms, ring_dep);   // the records are in registers
        if (trp) trace_put(trace, 1, k / kProducers, 4);
        mbar_wait(empty + 8 * s, ((k / S) & 1) ^ 1u);   // the MMAs of stage k - S have read this A slot and B slot
        tc_fence_after();
        if (trp) trace_put(trace, 1, k / kProducers, 5);
        const uint32_t a_slot = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + C::kAccCols + s * C::kAStageCols;
#pragma unroll
        for (int j = 0; j < kG; ++j)
          if (j < nt) {
#pragma unroll
            for (int m = 0; m < DBLK; ++m) tmem_st_32x32b_x8(a_slot + (j * DBLK + m) * 8, v[j][m]);
          }
        if (trp) trace_put(trace, 1, k / kProducers, 6);
        if (lane == 0 && member == 0) sts_u32(info_smem + 4 * s, fm | (lm << 8) | (static_cast<uint32_t>(nt) << 16));
#pragma unroll
        for (int jj = 0; jj < kTpm; ++jj)
          if (j_lo + jj < nt) sts_v4(b_smem + s * C::kBStageBytes + (j_lo + jj) * kBTileBytes + lane * 16, bv[jj]);
        tmem_st_wait();
        fence_proxy_async_smem();      // B tiles -> the MMA's operand fetch
        tc_fence_before();             // A columns -> the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * s);
        published += kProducers;
        if (trp) trace_put(trace, 1, k / kProducers, 7);
        continue;
      }
      // publish the oldest own stage once its copies have landed -- BEFORE blocking on a free slot, so a
      // landed stage never waits for the MMAs of an older one
      if (trp) trace_put(trace, 1, k / kProducers, 2);
      if (OWN_LAG > 0 && k - published >= OWN_LAG * kProducers) {
        cp_async_wait_group<(OWN_LAG > 0 ? OWN_LAG - 1 : 0)>();
        if (trp) trace_put(trace, 1, k / kProducers, 3);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * (published % S));
        published += kProducers;
      }
      if (trp) trace_put(trace, 1, k / kProducers, 4);
      mbar_wait(empty + 8 * s, ((k / S) & 1) ^ 1u);   // slot consumed by the MMAs of stage k - S
      if (trp) trace_put(trace, 1, k / kProducers, 5);
      if (!skip_gather) {
#pragma unroll 2
        for (int jj = 0; jj < kTpm; ++jj) {
          const int j = j_lo + jj;
          if (j < nt) {
            const uint32_t a_tile = a_smem + s * C::kAStageBytes + j * C::kATileBytes;
            const int4 c0 = lds_v4(meta + j * 64), c1 = lds_v4(meta + j * 64 + 16);   // rows to gather (-1: padding)
            const int32_t cols[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            const uint32_t any = static_cast<uint32_t>(c0.x | c0.y | c0.z | c0.w | c1.x | c1.y | c1.z | c1.w);
            ring_dep |= any;
            if (nvec == DBLK * 32 && static_cast<int32_t>(any) >= 0) {
              // full-width rows, no padding row (every tile of a window but its last): one multiply-add and one
              // copy per row and 128-feature block, nothing else
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const char* src = x_lane + static_cast<uint64_t>(static_cast<uint32_t>(cols[r])) * row_bytes;
#pragma unroll
                for (int h = 0; h < DBLK; ++h) {
                  const int v = h * 32 + lane;
                  cp_async_16_x(a_tile + (v >> 3) * 1024 + sw128_base32_offset(r, v & 7), src + h * 512, 16u, policy);
                }
              }
            } else if (nvec == DBLK * 32) {
              // full-width rows: lane -> vector `lane` (+32) of each of the 8 gathered rows
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int32_t col = cols[r];
                // one 32 x 32 -> 64 bit multiply-add per row: my 16 bytes of row `col` (row stride < 4 GB:
                // x_is_prerounded); the clamp keeps a padding row's unused address inside X
                const char* src = x_lane + static_cast<uint64_t>(static_cast<uint32_t>(col < 0 ? 0 : col)) * row_bytes;
                const uint32_t bytes = col < 0 ? 0u : 16u;   // padding column: zero-fill
#pragma unroll
                for (int h = 0; h < DBLK; ++h) {
                  const int v = h * 32 + lane;
                  cp_async_16_x(a_tile + (v >> 3) * 1024 + sw128_base32_offset(r, v & 7), src + h * 512, bytes, policy);
                }
              }
            } else {
              // narrow rows (dim < DBLK*128): (row, vector) pairs r-major over the warp
              const int items = 8 * nvec;
#pragma unroll
              for (int t = 0; t < kVecPerLane; ++t) {
                const int item = t * 32 + lane;
                if (item < items) {
                  const int r = item / nvec;
                  const int v = item - r * nvec;
                  int32_t col = cols[0];
#pragma unroll
                  for (int i = 1; i < 8; ++i) col = r == i ? cols[i] : col;
                  const float* src = x + static_cast<int64_t>(col < 0 ? 0 : col) * ldx + v * 4;
                  cp_async_16_x(a_tile + (v >> 3) * 1024 + sw128_base32_offset(r, v & 7), src, col < 0 ? 0u : 16u, policy);
                }
              }
            }
          }
        }
      }
      cp_async_commit_group();
      if (trp) trace_put(trace, 1, k / kProducers, 6);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_after_loads(meta_empty + 8 * ms, ring_dep);   // every lane of this warp HAS the records in registers
        if (member == 0) sts_u32(info_smem + 4 * s, fm | (lm << 8) | (static_cast<uint32_t>(nt) << 16));
      }
#pragma unroll
      for (int jj = 0; jj < kTpm; ++jj)
        if (j_lo + jj < nt) sts_v4(b_smem + s * C::kBStageBytes + (j_lo + jj) * kBTileBytes + lane * 16, bv[jj]);
      if (OWN_LAG == 0) {
        cp_async_wait_all();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * s);
        published += kProducers;
      }
      if (trp) trace_put(trace, 1, k / kProducers, 7);
    }
    // drain: publish the own stages still in flight
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (; published < n_stages; published += kProducers) mbar_arrive(full + 8 * (published % S));
  }

  // ===================================== teardown =======================================
  tc_fence_before();
  __syncthreads();
  if (kDebugSwitches && trace != nullptr && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
    long long* row = trace + 3 * kTraceStages * 8 + blockIdx.x * 4;
    row[0] = clock64() - t_start;
    row[1] = n_tiles;
    row[2] = sl.n_windows;
    row[3] = 0;
  }
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

uint32_t kernel_flags() {
  if (!kDebugSwitches) return kTuneDefault;
  static const uint32_t flags = [] {
    const char* a = getenv("TCGNN_ABLATE");
    const char* t = getenv("TCGNN_TUNE");
    const uint32_t ablate = a ? static_cast<uint32_t>(strtoul(a, nullptr, 0)) & 0x30Fu : 0u;
    const uint32_t tune = t ? static_cast<uint32_t>(strtoul(t, nullptr, 0)) & (kTuneMetaFirst | kTuneYStream)
                            : kTuneDefault;
    const char* c = getenv("TCGNN_TRACE_CTA");
    const uint32_t cta = c ? (static_cast<uint32_t>(atoi(c)) & 0xFFu) << 16 : 0u;   // bits [16,24)
    return ablate | tune | cta;
  }();
  return flags;
}

// TCGNN_TRACE=<path> (profiling builds): after every launch, block 0's timeline is written to <path>
// (3 roles x 512 stages x 8 int64)
long long* trace_buffer() {
  if (!kDebugSwitches) return nullptr;
  static long long* buf = [] {
    long long* p = nullptr;
    if (getenv("TCGNN_TRACE") != nullptr) {
      if (cudaMalloc(&p, sizeof(long long) * kTraceWords) != cudaSuccess) p = nullptr;
    }
    return p;
  }();
  if (buf != nullptr) cudaMemset(buf, 0, sizeof(long long) * kTraceWords);
  return buf;
}
void trace_dump(const long long* trace, cudaStream_t stream) {
  static long long host[kTraceWords];
  if (cudaStreamSynchronize(stream) != cudaSuccess) return;
  if (cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  if (FILE* f = fopen(getenv("TCGNN_TRACE"), "wb")) {
    fwrite(host, 1, sizeof(host), f);
    fclose(f);
  }
}

template <class C, int DBLK>
cudaError_t launch_kernel(const tcgnn_plan* plan, const PlanView& pv, int grid, const float* xr, int64_t ldr,
                          const float* wperm, float* y, int64_t ldy, int32_t dim, uint32_t mode_flags,
                          cudaStream_t stream) {
  // the opt-in to > 48 KB of dynamic shared memory is per device and per kernel instantiation
  static std::mutex attr_mu;
  static bool attr_set[64] = {};
  const int dev = plan->device;
  {
    std::lock_guard<std::mutex> lock(attr_mu);
    if (dev >= 64 || !attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(spmm_tc_kernel<C, DBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           C::kSmemBytes);
      if (e != cudaSuccess) return e;
      if (dev < 64) attr_set[dev] = true;
    }
  }
  long long* trace = trace_buffer();
  spmm_tc_kernel<C, DBLK><<<grid, C::kThreads, C::kSmemBytes, stream>>>(pv, xr, ldr, wperm, y, ldy, dim,
                                                                         kernel_flags() | mode_flags, trace);
  count_launch();
  if (trace != nullptr) trace_dump(trace, stream);
  return cudaGetLastError();
}

// pipeline shape: env TCGNN_PRESET selects among the compiled variants (profiling builds; 0 = production choice)
int preset_setting() {
  if (!kDebugSwitches) return 0;
  static const int v = [] {
    const char* e = getenv("TCGNN_PRESET");
    return e ? atoi(e) : 0;
  }();
  return v;
}

template <int DBLK>
cudaError_t launch_pass(const tcgnn_plan* plan, const PlanView& pv, int grid, const float* xr, int64_t ldr,
                        const float* wperm, float* y, int64_t ldy, int32_t dim, uint32_t mode_flags,
                        cudaStream_t stream) {
  if (!(mode_flags & kFlagAccumulate)) {
    spmm_zero_partial_rows<<<grid, 128, 0, stream>>>(pv, y, ldy, dim);
    count_launch();
  }
  constexpr int T = 8 / DBLK;   // tiles in 32 KB of A
  // Gathering warps per stage and publication lag (tuning knobs TCGNN_SPMM_TEAM / TCGNN_SPMM_LAG; results are
  // identical).  Measured on B200 (profiles/r02d_spmm_team_lag_ab.txt): the weighted kernel gains 1.25-1.4x from a
  // second warp per stage (its B tiles cost a load and four shuffles per lane and tile), the D = 256 kernels 3 %,
  // the unweighted D <= 128 kernel loses 2-3 % to the extra warps; publishing right after landing (lag 0) is worth
  // 1-5 % except for the weighted D = 256 kernel.
  const bool weighted = wperm != nullptr;
  static const int team_env = [] {
    const char* e = getenv("TCGNN_SPMM_TEAM");
    return e != nullptr ? atoi(e) : 0;
  }();
  static const int lag_env = [] {
    const char* e = getenv("TCGNN_SPMM_LAG");
    return e != nullptr ? atoi(e) : -1;
  }();
  const int team = team_env == 1 || team_env == 2 ? team_env : ((weighted || DBLK == 2) ? 2 : 1);
  const int lag = lag_env == 0 || lag_env == 1 ? lag_env : ((weighted && DBLK == 2) ? 1 : 0);
#define TCGNN_LAUNCH_TL(G, S, P, STAGED, TEAM, LAG) \
  launch_kernel<Cfg<DBLK, G, S, P, LAG, STAGED, TEAM>, DBLK>(plan, pv, grid, xr, ldr, wperm, y, ldy, dim, mode_flags, stream)
#define TCGNN_LAUNCH(G, S, P, STAGED)                                                          \
  return team == 1 ? (lag == 1 ? TCGNN_LAUNCH_TL(G, S, P, STAGED, 1, 1) : TCGNN_LAUNCH_TL(G, S, P, STAGED, 1, 0)) \
                   : (lag == 1 ? TCGNN_LAUNCH_TL(G, S, P, STAGED, 2, 1) : TCGNN_LAUNCH_TL(G, S, P, STAGED, 2, 0))
  // Dense windows (hundreds of tiles each: reddit): the deepest ring, output written straight from registers
  // (rare).  Sparse windows: one output tile every few tiles -- stage it and hand it to the TMA engine; the
  // staging buffers cost one pipeline slot.
  int preset = preset_setting();
  if (preset == 0) preset = static_cast<int64_t>(plan->num_tiles) >= 256LL * plan->num_windows ? 1 : 2;
  // (Y += ... keeps this choice: with >= 256 tiles per window the 16 scalar atomics per thread and window of the
  // register epilogue cost ~3 % while the staged shape gives up a pipeline stage -- measured on 2 GPUs, r02g)
  // TCGNN_SPMM_SHAPE=1 (tuning knob): half-size stages, twice as many of them, one gathering warp each -- the same
  // bytes in flight, recycled at twice the granularity
  static const int shape = [] {
    const char* e = getenv("TCGNN_SPMM_SHAPE");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (shape == 1) {
    if (preset == 1) return lag == 1 ? TCGNN_LAUNCH_TL(T / 2, 12, 12, false, 1, 1) : TCGNN_LAUNCH_TL(T / 2, 12, 12, false, 1, 0);
    return lag == 1 ? TCGNN_LAUNCH_TL(T / 2, 10, 10, true, 1, 1) : TCGNN_LAUNCH_TL(T / 2, 10, 10, true, 1, 0);
  }
#ifdef TCGNN_DEBUG_SWITCHES
  // Register gathers + A operand in tensor memory (Cfg::kTs), profiling builds only: TCGNN_SPMM_TS=1 selects it,
  // TCGNN_SPMM_TS_GROUPS the number of gathering teams (3: 18 warps, 4: 22 warps).  Correct (tests/test_gpu_spmm.py
  // passes with it) and 1.9x SLOWER than the shared-memory pipeline -- kept as the measured alternative.
  static const int ts_env = [] {
    const char* e = getenv("TCGNN_SPMM_TS");
    return e != nullptr ? atoi(e) : 0;
  }();
  static const int ts_groups = [] {
    const char* e = getenv("TCGNN_SPMM_TS_GROUPS");
    return e != nullptr ? atoi(e) : 3;
  }();
  if (ts_env != 0) {
#define TCGNN_LAUNCH_TS(STAGED, NG) \
  launch_kernel<Cfg<DBLK, T, 6, NG, 0, STAGED, 4, true>, DBLK>(plan, pv, grid, xr, ldr, wperm, y, ldy, dim, mode_flags, stream)
    if (preset == 1) return ts_groups == 4 ? TCGNN_LAUNCH_TS(false, 4) : TCGNN_LAUNCH_TS(false, 3);
    return ts_groups == 4 ? TCGNN_LAUNCH_TS(true, 4) : TCGNN_LAUNCH_TS(true, 3);
#undef TCGNN_LAUNCH_TS
  }
#endif
  switch (preset) {
    case 1: TCGNN_LAUNCH(T, 6, 6, false);
#ifdef TCGNN_DEBUG_SWITCHES
    case 5: TCGNN_LAUNCH(T, 5, 4, true);
    case 6: TCGNN_LAUNCH(T, 5, 5, false);
#endif
    default: TCGNN_LAUNCH(T, 5, 5, true);
  }
#undef TCGNN_LAUNCH
#undef TCGNN_LAUNCH_TL
}

}  // namespace

// edge_weight: CSR edge order (permuted + rounded here) unless TCGNN_W_TILE_ORDER says it already is the plan's
// tile-ordered, tf32-rounded array (what the fused AGNN entry's SDDMM leaves behind).
int spmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y, int64_t ldy,
                int32_t dim, uint32_t op_flags, cudaStream_t stream, int row_chunk) {
  PlanView pv = plan->view();
  if (row_chunk >= 0) {
    if (row_chunk + 1 >= static_cast<int>(plan->row_chunk_win.size()) || plan->chunk_slice_ptr == nullptr) {
      set_last_error("spmm: row chunk %d does not exist", row_chunk);
      return TCGNN_ERR_INVALID_ARG;
    }
    pv.slice_ptr = plan->chunk_slice_ptr + static_cast<size_t>(row_chunk) * (plan->grid + 1);
  }
  const float* wperm = nullptr;
  if (edge_weight != nullptr && plan->num_pairs > 0) {
    if (op_flags & TCGNN_W_TILE_ORDER) {
      wperm = edge_weight;
    } else {
      int st = plan_ensure_eperm(plan, stream);
      if (st != TCGNN_OK) return st;
      st = plan_ensure_scratch(plan, &plan->weight_perm, static_cast<size_t>(plan->num_pairs));
      if (st != TCGNN_OK) return st;
      int g = (plan->num_pairs + 255) / 256;
      if (g > 148 * 16) g = 148 * 16;
      permute_weights_kernel<<<g, 256, 0, stream>>>(plan->eperm, edge_weight, plan->weight_perm, plan->num_pairs);
      count_launch();
      wperm = plan->weight_perm;
    }
  }
  // Xr = tf32_rna(X), packed [num_cols, ldr] with ldr % 4 == 0 (16-byte aligned rows)
  int64_t ldr = (static_cast<int64_t>(dim) + 3) / 4 * 4;
  const float* xr = nullptr;
  if (x_is_prerounded(x, ldx, dim, op_flags)) {
    xr = x;      // the caller rounded X already (tcgnn_round_tf32) and its rows are 16-byte aligned
    ldr = ldx;
  } else {
    int st = round_pack_launch(plan, x, ldx, dim, ldr, stream, &xr);
    if (st != TCGNN_OK) return st;
  }
  // one persistent CTA per SM, each owning a slice of the tile stream (plan->slice_ptr)
  const int grid = plan->grid;
  const uint32_t mode = (op_flags & TCGNN_ACCUMULATE) ? kFlagAccumulate : 0u;
  for (int32_t f0 = 0; f0 < dim; f0 += 256) {
    const int32_t d = dim - f0 < 256 ? dim - f0 : 256;
    cudaError_t e = d > 128 ? launch_pass<2>(plan, pv, grid, xr + f0, ldr, wperm, y + f0, ldy, d, mode, stream)
                            : launch_pass<1>(plan, pv, grid, xr + f0, ldr, wperm, y + f0, ldy, d, mode, stream);
    if (e != cudaSuccess) {
      set_last_error("spmm kernel launch failed: %s", cudaGetErrorString(e));
      return TCGNN_ERR_CUDA;
    }
  }
  return TCGNN_OK;
}

}  // namespace tcgnn
