// SDDMM (edge scores) on 5th-gen tensor cores, sm_100a.
//
// Replaces sddmm_forward_cuda_kernel of the reference (/root/reference
// TCGNN_conv/TCGNN_kernel.cu:584-728):  out[e] = sum_k tf32(X[row(e),k]) * tf32(X[col(e),k]).
// Per row window w and per group of up to 16 TC blocks (= 128 condensed columns) one dense
// contraction on tcgen05.mma:
//     S^T (128 gathered cols x 16 window rows) = Xg (128 x D, K-major) . Xw^T (D x 16, K-major)
//   M = 128 (TMEM lane = condensed column), N = 16 (TMEM column = window row), K = D in steps of 8.
// Both operands are rows of X, i.e. naturally K-major: the same 128B-swizzled row image serves as
// A (gathered rows) and as B (the window's own 16 rows).  Only positions that are edges are
// written (tile occupancy mask), in tile order; a second pass restores CSR edge order.
//
// A pipeline stage is one 64-feature chunk of one group: A 32 KB + B 4 KB, as two 32-feature sub-blocks (the
// 128B-swizzled K-major image holds 32 floats per row); 8 MMAs per stage amortise the MMA warp's per-stage latency.  The group's 16 tile records
// travel once per group through their own ring (TMA bulk copy), far ahead of the data.  X is first rounded
// to tf32 (cvt.rna) and packed once per call (round_pack.cu), so the row gathers are plain asynchronous
// copies.  CTA = 10 warps:
//   warps 0-3  epilogue     TMEM -> registers -> masked scatter of edge values
//   warp  4    MMA issuer   (warp-uniform loop, one elected lane issues)
//   warp  5    meta loader  (TMA)
//   warps 6-11 producers   warp p OWNS the stages k = p (mod 6): cp.async (LDGSTS, zero-fill) of the 128
//                           gathered rows + the window's 16 rows, 32 features each, into the 128B-swizzled
//                           K-major images; it publishes the stage with ONE mbarrier arrive once
//                           cp.async.wait_group says the copies have landed (see spmm_tc.cu for the history
//                           of this pipeline: per-thread cp.async.mbarrier arrivals and all-warps-per-stage
//                           schemes were 2-2.5x slower).
#include <stdlib.h>

#include "plan.h"

namespace tcgnn {

namespace {

constexpr int kStages = 5;
constexpr int kMetaStages = 16;                           // ring of group records
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kMetaWarp = 5;
constexpr int kProducerWarp0 = 6;
constexpr int kProducers = 5;                             // producer team p owns the stages p, p + 5, ...
// kTeam (template parameter of the kernel): warps per team, each gathers an equal share of a stage's rows
constexpr int kAcc = 4;
constexpr int kGroupTiles = 16;
constexpr int kChunk = 64;                                 // features per pipeline stage: two 32-feature sub-blocks
constexpr int kASubBytes = 128 * 128;                      // 128 rows x 32 floats, 128B-swizzled K-major
constexpr int kBSubBytes = TCGNN_BLK_H * 128;              // 16 rows x 32 floats
constexpr int kAStageBytes = 2 * kASubBytes;               // 32 KB
constexpr int kBStageBytes = 2 * kBSubBytes;               // 4 KB
constexpr int kMetaTileBytes = kGroupTiles * static_cast<int>(sizeof(TileMeta));
constexpr int kMetaStageBytes = kMetaTileBytes + 16;       // + header {tile_start, ntiles, win, 0}
constexpr uint32_t kTmemCols = kAcc * 16;
constexpr int kOutStageBytes = kGroupTiles * TCGNN_BLK_H * TCGNN_BLK_W * 4;   // a group's edge values, compacted (8 KB)
constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + 2 * kOutStageBytes +
                           kMetaStages * kMetaStageBytes + (2 * kMetaStages + 2 * kStages + 2 * kAcc) * 8 + 64 + 1024;
static_assert(kProducers <= kStages, "a warp may not wait for the slot of an own stage it has not published yet");
static_assert(kMetaStages >= 2 * kProducers, "one feature chunk per group: every in-flight stage is another group");
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
constexpr uint32_t kDbgTrailingCopy = 1u;                  // diagnostics: one more (zero-fill) cp.async closes every stage

template <int kTeam>
__global__ void __launch_bounds__((kProducerWarp0 + kProducers * kTeam) * 32, 1)
sddmm_tc_kernel(PlanView pv, const int4* __restrict__ groups, int32_t num_groups,
                const float* __restrict__ x /* tf32-rounded, 16B aligned */, int64_t ldx /* % 4 == 0 */,
                float* __restrict__ out_raw /* nullable: scores, tile order */,
                float* __restrict__ out_att /* nullable: tf32_rna(score * *scale), tile order */,
                const float* __restrict__ scale /* nullable (1.0); device scalar */, int32_t dim, uint32_t flags) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem;
  const uint32_t b_smem = a_smem + kStages * kAStageBytes;
  const uint32_t o_smem = b_smem + kStages * kBStageBytes;      // [2] compacted edge values of a group
  const uint32_t m_smem = o_smem + 2 * kOutStageBytes;
  const uint32_t bars = m_smem + kMetaStages * kMetaStageBytes;
  const uint32_t meta_full = bars, meta_empty = bars + 8 * kMetaStages;
  const uint32_t full = bars + 16 * kMetaStages, empty = full + 8 * kStages;
  const uint32_t acc_full = empty + 8 * kStages, acc_empty = acc_full + 8 * kAcc;
  const uint32_t tmem_slot = acc_empty + 8 * kAcc;
  uint8_t* const smem_gen = smem_raw + (smem - smem_u32(smem_raw));   // generic view (header store)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // every group costs the same here (nkc stages of 128 + 16 rows, however many of its 16 tiles exist), so the
  // persistent CTAs take equal shares of the GROUPS -- not of the tiles: R-MAT graphs have long runs of one-tile
  // windows, i.e. one-tile groups
  const int32_t g_lo = static_cast<int32_t>(static_cast<int64_t>(num_groups) * blockIdx.x / gridDim.x);
  const int32_t g_hi = static_cast<int32_t>(static_cast<int64_t>(num_groups) * (blockIdx.x + 1) / gridDim.x);
  const int32_t n_groups = g_hi - g_lo;
  const int32_t nkc = (dim + kChunk - 1) / kChunk;   // 64-feature chunks
  const int32_t n_stages = n_groups * nkc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMetaStages; ++s) {
      mbar_init(meta_full + 8 * s, 1);
      mbar_init(meta_empty + 8 * s, nkc * kTeam);  // the owners of the group's nkc stages have taken the row ids
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + 8 * s, kTeam);              // the stage's producer warps, once their copies have landed
      mbar_init(empty + 8 * s, 1);                 // tcgen05.commit
    }
    for (int b = 0; b < kAcc; ++b) {
      mbar_init(acc_full + 8 * b, 1);
      mbar_init(acc_empty + 8 * b, kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds_u32(tmem_slot);

  if (warp < kEpiWarps) {
    // ===================================== epilogue =====================================
    // The edge values of a group are one contiguous run of the tile-ordered output (tiles are consecutive, a
    // tile's edges are in mask-bit order).  They are compacted in shared memory and written with coalesced
    // stores: ~n/128 store instructions per thread instead of up to 16 scattered ones (a global store
    // instruction costs ~100 cycles of the SM's load/store pipeline next to the gathers, spmm_tc.cu).
    const int q = warp;
    const int m = q * 32 + lane;       // condensed column inside the group
    const int tt = m >> 3, c = m & 7;  // tile inside the group, column inside the tile
    const int tid = threadIdx.x;       // 0..127
    const float att_scale = scale != nullptr ? __ldg(scale) : 1.0f;
    // The group's tile records (occupancy masks, output offsets) come from global memory through two dependent loads
    // (work unit -> records): they are fetched one group ahead (the work unit two ahead), so their L2 latency is
    // hidden behind the previous group's epilogue instead of sitting between "accumulator ready" and its drain --
    // the four epilogue warps see one accumulator every two pipeline stages.
    struct Records {
      uint4 mask, last_mask;
      int32_t edge_ofs, first_ofs, last_ofs, ntiles;
    };
    auto fetch = [&](const int4 g) {
      Records r;
      r.mask = make_uint4(0, 0, 0, 0);
      r.edge_ofs = 0;
      r.ntiles = g.y;
      if (tt < g.y) {
        const TileMeta* t = pv.tiles + g.x + tt;
        r.mask = *reinterpret_cast<const uint4*>(t->mask);
        r.edge_ofs = t->edge_ofs;
      }
      // first / one-past-last output index of the group (same addresses for all threads: broadcast loads)
      r.first_ofs = pv.tiles[g.x].edge_ofs;
      const TileMeta* tl = pv.tiles + g.x + g.y - 1;
      r.last_mask = *reinterpret_cast<const uint4*>(tl->mask);
      r.last_ofs = tl->edge_ofs;
      return r;
    };
    int4 g_ahead = n_groups > 1 ? groups[g_lo + 1] : make_int4(0, 1, 0, 0);
    Records ahead = n_groups > 0 ? fetch(groups[g_lo]) : Records{};
    for (int32_t gl = 0; gl < n_groups; ++gl) {
      const Records rec = ahead;
      if (gl + 1 < n_groups) {
        ahead = fetch(g_ahead);
        if (gl + 2 < n_groups) g_ahead = groups[g_lo + gl + 2];
      }
      const uint4 mask = rec.mask;
      const int32_t edge_ofs = rec.edge_ofs;
      const int32_t e0 = rec.first_ofs;
      const uint4 ml = rec.last_mask;
      const int32_t n_out = rec.last_ofs + __popc(ml.x) + __popc(ml.y) + __popc(ml.z) + __popc(ml.w) - e0;
      const int b = gl % kAcc;
      mbar_wait_backoff(acc_full + 8 * b, (gl / kAcc) & 1);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * b);
      // bit (r*8+c): word r/4, bit (r%4)*8+c ; rank in bit order = position in tile-ordered output
      const uint32_t obuf = o_smem + (gl & 1) * kOutStageBytes;
      const uint32_t mw[4] = {mask.x, mask.y, mask.z, mask.w};
      int base_rank = edge_ofs - e0;
#pragma unroll
      for (int wd = 0; wd < 4; ++wd) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int bit = rr * 8 + c;
          if (mw[wd] & (1u << bit)) {
            const int rank = base_rank + __popc(mw[wd] & ((1u << bit) - 1u));
            sts_u32(obuf + rank * 4, v[wd * 4 + rr]);
          }
        }
        base_rank += __popc(mw[wd]);
      }
      // double-buffered: the barrier of group gl also orders everybody's reads of group gl-1's buffer before
      // the writes of group gl+1 into it
      named_barrier_sync(1, kEpiWarps * 32);
      // Coalesced tile-order stores only: writing CSR order from here (out[eperm[i]], 32 scattered sectors per store
      // instruction) next to the producers' gathers made the kernel 1.2-1.9x slower (profiles/r02a_*: reddit R-MAT
      // 4.9 -> 9.0 ms) -- the LSU is the contended unit.  The fused AGNN path needs no CSR order at all: it leaves
      // the scaled, tf32-rounded attention right here for the weighted SpMM.
      for (int i = tid; i < n_out; i += kEpiWarps * 32) {
        const float v = __uint_as_float(lds_u32(obuf + i * 4));
        if (out_raw != nullptr) out_raw[e0 + i] = v;
        if (out_att != nullptr) out_att[e0 + i] = tf32_rna(v * att_scale);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer ===================================
    // warp-uniform loop, one elected lane issues
    constexpr uint32_t idesc = make_idesc_tf32(128, 16, false, false);
    // K-major, 128B swizzle: 8-row groups 1024 B apart (SBO); K steps advance the start address by 32 B
    const uint64_t desc0 = make_smem_desc(0, 16, 1024, kSwizzle128B);
    int s = 0;
    uint32_t ph = 0;
    int32_t kc = 0, gl = 0;
    for (int32_t k = 0; k < n_stages; ++k) {
      const int b = gl % kAcc;
      if (kc == 0) {
        mbar_wait(acc_empty + 8 * b, ((gl / kAcc) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(full + 8 * s, ph);   // producers fenced their generic-proxy writes before arriving
      tc_fence_after();
      const uint64_t adesc = desc0 | static_cast<uint64_t>(((a_smem + s * kAStageBytes) & 0x3FFFFu) >> 4);
      const uint64_t bdesc = desc0 | static_cast<uint64_t>(((b_smem + s * kBStageBytes) & 0x3FFFFu) >> 4);
      const int ksteps = min(8, (dim - kc * kChunk + 7) >> 3);   // K = 8 per MMA; 4 per 32-feature sub-block
      if (elect_one()) {
        if (ksteps == 8) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32(tmem_base + b * 16, adesc + static_cast<uint64_t>(((ks >> 2) * kASubBytes >> 4) + (ks & 3) * 2),
                      bdesc + static_cast<uint64_t>(((ks >> 2) * kBSubBytes >> 4) + (ks & 3) * 2), idesc,
                      (kc | ks) != 0 ? 1u : 0u);
        } else {
          for (int ks = 0; ks < ksteps; ++ks)
            umma_tf32(tmem_base + b * 16, adesc + static_cast<uint64_t>(((ks >> 2) * kASubBytes >> 4) + (ks & 3) * 2),
                      bdesc + static_cast<uint64_t>(((ks >> 2) * kBSubBytes >> 4) + (ks & 3) * 2), idesc,
                      (kc | ks) != 0 ? 1u : 0u);
        }
        umma_commit(empty + 8 * s);
        if (kc == nkc - 1) umma_commit(acc_full + 8 * b);
      }
      if (++kc == nkc) { kc = 0; ++gl; }
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    // the last commit must land in this CTA's shared memory before the CTA may retire
    if (n_stages > 0) mbar_wait(empty + 8 * ((n_stages - 1) % kStages), ((n_stages - 1) / kStages) & 1);
  } else if (warp == kMetaWarp) {
    // ===================================== meta loader (TMA) ============================
    if (lane == 0) {
      int4 nxt = n_groups > 0 ? groups[g_lo] : make_int4(0, 0, 0, 0);
      int ms = 0;
      uint32_t mph = 0;
      for (int32_t gl = 0; gl < n_groups; ++gl) {
        const int4 grp = nxt;
        if (gl + 1 < n_groups) nxt = groups[g_lo + gl + 1];   // prefetch the next work unit
        mbar_wait(meta_empty + 8 * ms, mph ^ 1u);
        *reinterpret_cast<int4*>(smem_gen + (m_smem - smem) + ms * kMetaStageBytes + kMetaTileBytes) =
            make_int4(grp.x, grp.y, grp.z, 0);                // ordered before the readers by the arrive below
        const uint32_t bytes = static_cast<uint32_t>(grp.y) * sizeof(TileMeta);
        mbar_arrive_expect_tx(meta_full + 8 * ms, bytes);
        tma_bulk_g2s(smem_gen + (m_smem - smem) + ms * kMetaStageBytes, pv.tiles + grp.x, bytes, meta_full + 8 * ms);
        if (++ms == kMetaStages) { ms = 0; mph ^= 1u; }
      }
    }
  } else {
    // ===================================== producers ====================================
    const int pw = (warp - kProducerWarp0) / kTeam;   // team
    const int member = (warp - kProducerWarp0) % kTeam;
    constexpr int kRows = 128 + TCGNN_BLK_H;          // gathered rows + the window's own rows
    constexpr int kPerLane = kRows * 8 / 32 / kTeam;  // 18 16-byte vectors per lane, stage and feature sub-block
    static_assert((kRows * 8 / 32) % kTeam == 0, "the members of a team gather equal shares of a stage");
    const int u_lo = member * kPerLane;
    const int nvec = (dim + 3) >> 2;                  // valid 16-byte vectors per row
    const int v = lane & 7;                           // my vector inside the 32-feature chunk
    const int rsub = lane >> 3;                       // my row inside each group of 4 rows
    const uint64_t policy = x_gather_policy();
    // row address = one 32 x 32 -> 64 bit multiply-add (row stride < 4 GB: x_is_prerounded); the empty asm keeps
    // the base in one register pair
    uint64_t x_bits = reinterpret_cast<uint64_t>(x);
    asm volatile("" : "+l"(x_bits));
    const char* x_bytes = reinterpret_cast<const char*>(x_bits);
    const uint32_t row_bytes = static_cast<uint32_t>(ldx) * 4u;
    int32_t published = pw;                           // oldest own stage not yet published
    int32_t gl = pw / nkc, kc = pw % nkc;
    for (int32_t k = pw; k < n_stages; k += kProducers) {
      const int s = k % kStages;
      const int ms = gl % kMetaStages;
      mbar_wait(meta_full + 8 * ms, (gl / kMetaStages) & 1);
      const uint32_t meta = m_smem + ms * kMetaStageBytes;
      const int4 hdr = lds_v4(meta + kMetaTileBytes);
      const int32_t ntiles = hdr.y, win = hdr.z;
      int32_t node[kPerLane];                         // row u*4 + rsub of the stage image
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) {
        const int row = (u_lo + u) * 4 + rsub;
        node[u] = -1;
        if (row < 128) {
          if ((row >> 3) < ntiles) node[u] = static_cast<int32_t>(lds_u32(meta + (row >> 3) * 64 + (row & 7) * 4));
        } else {
          const int32_t r = win * TCGNN_BLK_H + (row - 128);
          node[u] = r < pv.num_nodes ? r + pv.row_base : -1;   // the window's own rows (global ids)
        }
      }
      // the ring slot may be refilled as soon as this arrive lands: every row id must be IN its register first
      uint32_t node_or = 0u;
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) node_or |= static_cast<uint32_t>(node[u]);
      const uint32_t dep = node_or | static_cast<uint32_t>(ntiles ^ win);
      __syncwarp();
      if (lane == 0) mbar_arrive_after_loads(meta_empty + 8 * ms, dep);
      // publish the previous own stage once its copies have landed -- BEFORE blocking on a free slot
      if (k - published >= kProducers) {
        cp_async_wait_group<0>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * (published % kStages));
        published += kProducers;
      }
      mbar_wait(empty + 8 * s, ((k / kStages) & 1) ^ 1u);    // slot consumed by the MMAs of stage k - kStages
      const uint32_t a_stage = a_smem + s * kAStageBytes;
      const uint32_t b_stage = b_smem + s * kBStageBytes;
      const bool second = dim - kc * kChunk > 32;            // the MMAs read the second sub-block too
      // every row of the stage exists and the chunk holds all 64 features (all stages of a full group at D % 64 == 0):
      // one multiply-add and two copies per row, no clamps, no zero-fill predicates
      const bool plain = __all_sync(0xffffffffu, static_cast<int32_t>(node_or) >= 0) && dim - kc * kChunk >= kChunk;
      if (plain) {
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
          const int row = (u_lo + u) * 4 + rsub;
          const bool is_a = row < 128;
          const uint32_t dst = (is_a ? a_stage + (row >> 3) * 1024 : b_stage + ((row - 128) >> 3) * 1024) +
                               sw128_offset(row & 7, v);
          const char* rowp = x_bytes + static_cast<uint64_t>(static_cast<uint32_t>(node[u])) * row_bytes +
                             (kc * (kChunk / 4) + v) * 16;
          cp_async_16_x(dst, rowp, 16u, policy);
          cp_async_16_x(dst + (is_a ? kASubBytes : kBSubBytes), rowp + 128, 16u, policy);
        }
      } else
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) {
        const int row = (u_lo + u) * 4 + rsub;
        const bool is_a = row < 128;
        const uint32_t dst = (is_a ? a_stage + (row >> 3) * 1024 : b_stage + ((row - 128) >> 3) * 1024) +
                             sw128_offset(row & 7, v);
        const uint32_t sub_step = is_a ? kASubBytes : kBSubBytes;
        const char* rowp = x_bytes + static_cast<uint64_t>(static_cast<uint32_t>(node[u] >= 0 ? node[u] : 0)) * row_bytes;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub == 1 && !second) break;
          const int vg = kc * (kChunk / 4) + sub * 8 + v;    // vector index inside the feature row
          const bool valid = node[u] >= 0 && vg < nvec;
          cp_async_16_x(dst + sub * sub_step, rowp + (valid ? vg * 16 : 0), valid ? 16u : 0u, policy);  // zero-fill
        }
      }
      if (flags & kDbgTrailingCopy) cp_async_16(tmem_slot + 16 + (lane & 1) * 16, x, 0u);
      cp_async_commit_group();
      kc += kProducers;
      while (kc >= nkc) { kc -= nkc; ++gl; }
    }
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (; published < n_stages; published += kProducers) mbar_arrive(full + 8 * (published % kStages));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// tile order -> CSR edge order
__global__ void unpermute_kernel(const int32_t* __restrict__ eperm, const float* __restrict__ in,
                                 float* __restrict__ out, int32_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[eperm[i]] = in[i];
}

}  // namespace

// edge_out_csr (nullable): scores in CSR edge order.  tile_out (nullable): tf32_rna(score * *scale) in the plan's
// tile order -- the weight operand of spmm_launch(TCGNN_W_TILE_ORDER).
int sddmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out_csr, float* tile_out,
                 const float* scale, int32_t dim, uint32_t op_flags, cudaStream_t stream) {
  if (plan->num_edges == 0) return TCGNN_OK;
  int st = TCGNN_OK;
  float* raw_tile = nullptr;
  if (edge_out_csr != nullptr) {   // CSR order = tile order + one permutation pass
    st = plan_ensure_eperm(plan, stream);
    if (st == TCGNN_OK) st = plan_ensure_scratch(plan, &plan->sddmm_perm, static_cast<size_t>(plan->num_pairs));
    if (st != TCGNN_OK) return st;
    raw_tile = plan->sddmm_perm;
  }
  st = plan_ensure_groups(plan, stream);
  if (st != TCGNN_OK) return st;
  // warps per producer team: a tuning knob (results are identical), TCGNN_SDDMM_TEAM=1|2|3
  static const int team = [] {
    const char* e = getenv("TCGNN_SDDMM_TEAM");
    const int v = e ? atoi(e) : 2;
    return v >= 1 && v <= 3 ? v : 2;
  }();
  static const uint32_t dbg = [] {
    const char* e = getenv("TCGNN_SDDMM_DBG");
    return e ? static_cast<uint32_t>(atoi(e)) & kDbgTrailingCopy : 0u;
  }();
  static std::mutex attr_mu;
  static bool attr_set[64] = {};
  cudaError_t e;
  {
    std::lock_guard<std::mutex> lock(attr_mu);
    if (plan->device >= 64 || !attr_set[plan->device]) {
      e = cudaFuncSetAttribute(sddmm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(sddmm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(sddmm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) {
        set_last_error("cudaFuncSetAttribute(sddmm) failed: %s", cudaGetErrorString(e));
        return TCGNN_ERR_CUDA;
      }
      if (plan->device < 64) attr_set[plan->device] = true;
    }
  }
  const int grid = plan->grid;
  int64_t ldr = (static_cast<int64_t>(dim) + 3) / 4 * 4;
  const float* xr = nullptr;
  if (x_is_prerounded(x, ldx, dim, op_flags)) {
    xr = x;
    ldr = ldx;
  } else {
    st = round_pack_launch(plan, x, ldx, dim, ldr, stream, &xr);
    if (st != TCGNN_OK) return st;
  }
  if (edge_out_csr != nullptr && static_cast<int64_t>(plan->num_pairs) < plan->num_edges) {
    // duplicated (row, col) pairs: only one edge of each pair receives the value (as in the reference)
    e = cudaMemsetAsync(edge_out_csr, 0, sizeof(float) * static_cast<size_t>(plan->num_edges), stream);
    if (e != cudaSuccess) {
      set_last_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
      return TCGNN_ERR_CUDA;
    }
  }
  const PlanView pv = plan->view();
  const uint32_t kflags = dbg;
#define TCGNN_SDDMM(T)                                                                                      \
  sddmm_tc_kernel<T><<<grid, (kProducerWarp0 + kProducers * T) * 32, kSmemBytes, stream>>>(                \
      pv, plan->groups, plan->num_groups, xr, ldr, raw_tile, tile_out, scale, dim, kflags)
  if (team == 1) TCGNN_SDDMM(1);
  else if (team == 3) TCGNN_SDDMM(3);
  else TCGNN_SDDMM(2);
#undef TCGNN_SDDMM
  count_launch();
  if (edge_out_csr != nullptr) {
    int g = (plan->num_pairs + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    unpermute_kernel<<<g, 256, 0, stream>>>(plan->eperm, raw_tile, edge_out_csr, plan->num_pairs);
    count_launch();
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("sddmm kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  return TCGNN_OK;
}

// Fused AGNN edge pipeline (reference gnn_conv.py:125-132: SDDMM -> x attention_w -> weighted SpMM): the SDDMM
// epilogue leaves tf32_rna(score * attention_w) in tile order, which is exactly the B-operand stream of the weighted
// SpMM -- no unpermute / permute passes, no [E] round trip through the caller, X rounded once for both kernels.
int agnn_launch(tcgnn_plan* plan, const float* x, int64_t ldx, const float* attention_w, float* y, int64_t ldy,
                float* att_tile_out, float* edge_out_csr, int32_t dim, uint32_t op_flags, cudaStream_t stream) {
  float* att = att_tile_out;
  if (att == nullptr) {
    int st = plan_ensure_scratch(plan, &plan->weight_perm, static_cast<size_t>(plan->num_pairs));
    if (st != TCGNN_OK) return st;
    att = plan->weight_perm;
  }
  // X is rounded once for both kernels when its width allows 16-byte rows; a ragged width makes each kernel pack
  // its own zero-padded copy (as the separate entry points do)
  const float* xr = x;
  int64_t ldr = ldx;
  uint32_t flags = op_flags;
  if (!x_is_prerounded(x, ldx, dim, op_flags)) {
    flags &= ~TCGNN_X_IS_TF32;
    if ((dim & 3) == 0) {
      int st = round_pack_launch(plan, x, ldx, dim, dim, stream, &xr);
      if (st != TCGNN_OK) return st;
      ldr = dim;
      flags |= TCGNN_X_IS_TF32;
    }
  }
  if (plan->num_edges > 0) {
    int st = sddmm_launch(plan, xr, ldr, edge_out_csr, att, attention_w, dim, flags, stream);
    if (st != TCGNN_OK) return st;
  }
  return spmm_launch(plan, xr, ldr, plan->num_pairs > 0 ? att : nullptr, y, ldy, dim,
                     (flags & ~TCGNN_ACCUMULATE) | TCGNN_W_TILE_ORDER, stream);
}

}  // namespace tcgnn
