/*
 * tcgnn_b200 -- C ABI of the B200-native TC-GNN aggregation path (SGT + SpMM + SDDMM).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * The Python extension module `TCGNN` (tc-gnn_atc23_b200/csrc/binding.cpp) is a thin
 * torch-aware caller of exactly these entry points; INTEGRATION.md shows how the reference's
 * own TCGNN.cpp would bind them.  All functions return 0 on success or a negative
 * tcgnn_status; they never call exit(), never synchronise the device unless stated, and
 * launch on the stream they are given (`stream` is a cudaStream_t passed as void*).
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   tcgnn_sgt_cpu      <- preprocess            TCGNN_conv/TCGNN.cpp:172-226
 *   tcgnn_sgt_cuda     <- preprocess_gpu        TCGNN_conv/TCGNN.cpp:229-256 (+ fill_edgeToRow /
 *                         fill_window stubs     TCGNN_conv/TCGNN_kernel.cu:21-119)
 *   tcgnn_spmm_f32     <- spmm_forward_cuda     TCGNN_conv/TCGNN_kernel.cu:175-220 (kernel :336-454)
 *                         spmmAGNN_forward_cuda TCGNN_conv/TCGNN_kernel.cu:227-279 (kernel :459-578)
 *   tcgnn_sddmm_f32    <- sddmm_forward_cuda    TCGNN_conv/TCGNN_kernel.cu:286-327 (kernel :584-728)
 *   tcgnn_agnn_f32     <- the AGNN edge pipeline of gnn_conv.py:125-132 (forward_ef -> torch.mm with attention_w ->
 *                         transpose/contiguous -> forward_AGNN) as one call
 *   tcgnn_csr_transpose<- (new) A^T for a backward pass that is correct on directed graphs; the reference re-uses
 *                         the forward CSR, i.e. assumes A == A^T (gnn_conv.py:76-85)
 *   tcgnn_plan_*       <- (new) the kernel-side layout derived once per graph from the SGT arrays;
 *                         the reference re-derives it inside every kernel launch by rescanning all
 *                         window edges per tile (TCGNN_kernel.cu:399-408).
 */
#ifndef TCGNN_B200_H
#define TCGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCGNN_BLK_H 16 /* TCGNN_conv/config.h:4 */
#define TCGNN_BLK_W 8  /* TCGNN_conv/config.h:5 */

typedef enum tcgnn_status {
  TCGNN_OK = 0,
  TCGNN_ERR_INVALID_ARG = -1, /* null pointer, negative size, unsupported blk_h / blk_w, bad alignment */
  TCGNN_ERR_CUDA = -2,        /* a CUDA runtime call or a kernel launch failed (see tcgnn_last_error) */
  TCGNN_ERR_NO_DEVICE = -3,   /* no CUDA device / not an sm_100 device */
  TCGNN_ERR_OOM = -4,         /* host or device allocation failed */
  TCGNN_ERR_OVERFLOW = -5     /* tile or edge count does not fit the 32-bit plan indices */
} tcgnn_status;

typedef struct tcgnn_plan tcgnn_plan; /* opaque, owns device memory */

/* Library / diagnostics ------------------------------------------------------------------- */
int tcgnn_version(void);                    /* 10000*major + 100*minor + patch */
const char* tcgnn_status_string(int status);
const char* tcgnn_last_error(void);         /* thread-local detail of the last failure */

/* SGT (sparse-graph translation) ------------------------------------------------------------
 * Bit-exact with the reference `preprocess` (TCGNN.cpp:172-226):
 *   edge_to_row[e]      = row owning edge e
 *   S_w                 = sorted unique col_idx over the edges of window w (blk_h rows)
 *   edge_to_col[e]      = rank of col_idx[e] in S_w
 *   block_partition[w]  = ceil(max(|S_w|, 1) / blk_w)        (empty window -> 1, as the reference)
 * *tc_blocks_out (nullable) receives sum(block_partition) (+1 when num_nodes % blk_h == 0, the
 * total the reference prints at TCGNN.cpp:225; its out-of-bounds write is not reproduced).
 * Host version: all pointers are host memory; num_threads <= 0 means all hardware threads.
 * Device version: all pointers are device memory; *tc_blocks_out is host memory and, when
 * non-null, the call synchronises `stream` to fill it. */
int tcgnn_sgt_cpu(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges,
                  int32_t blk_h, int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col,
                  int32_t* edge_to_row, int64_t* tc_blocks_out, int32_t num_threads);
int tcgnn_sgt_cuda(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges,
                   int32_t blk_h, int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col,
                   int32_t* edge_to_row, int64_t* tc_blocks_out, void* stream);

/* Row-panel variant (1-D destination-row sharding): row_ptr describes `num_rows` consecutive rows
 * (rebased to start at 0) whose col_idx are global node ids in [0, num_cols).  The SGT of a panel
 * that starts on a window boundary equals the panel's slice of the whole graph's SGT. */
int tcgnn_sgt_cuda_panel(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                         int64_t num_edges, int32_t blk_h, int32_t blk_w, int32_t* block_partition,
                         int32_t* edge_to_col, int32_t* edge_to_row, int64_t* tc_blocks_out, void* stream);

/* Plan ---------------------------------------------------------------------------------------
 * Derives the kernel-side tile stream from the caller's SGT arrays (all device pointers, the
 * same five arrays the reference kernels take, TCGNN_kernel.cu:336-346).  One 64-byte record
 * per 16x8 TC block: the 8 gathered feature rows, a 128-bit occupancy mask, the owning window
 * and the offset of its edges in tile order.  Synchronises `stream` (it needs the tile total). */
int tcgnn_plan_create(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                      const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_nodes,
                      int64_t num_edges, int32_t num_windows, void* stream, tcgnn_plan** plan_out);
/* Row-panel variant for 1-D destination-row sharding (new; the reference is single-GPU): the plan
 * covers `num_rows` consecutive rows of a graph with `num_cols` nodes.  row_ptr is the panel's own
 * CSR pointer rebased to start at 0, the SGT arrays are the panel's slices of the global ones
 * (edge_to_row rebased to the panel), col_idx keeps GLOBAL node ids in [0, num_cols).  X passed to
 * the kernels has num_cols rows; SpMM writes the panel's num_rows output rows; SDDMM reads the
 * panel's own feature rows at X[row_base + r].  tcgnn_plan_create == panel with num_cols ==
 * num_rows and row_base == 0.  row_base == -1: the plan's rows are not rows of X at all (a partial
 * product over one source panel's packed rows: col_idx indexes that block, num_cols may be smaller than
 * num_rows); such a plan serves SpMM only. */
int tcgnn_plan_create_panel(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                            const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_rows,
                            int32_t num_cols, int32_t row_base, int64_t num_edges, int32_t num_windows,
                            void* stream, tcgnn_plan** plan_out);
int tcgnn_plan_destroy(tcgnn_plan* plan);
/* info[0]=num_nodes info[1]=num_edges info[2]=num_windows info[3]=num_tiles info[4]=plan bytes
 * info[5]=distinct (row,col) pairs info[6]=device ordinal info[7]=SM count */
int tcgnn_plan_info(const tcgnn_plan* plan, int64_t info[8]);

/* SpMM: Y[i,:] = sum_{e in row i} w_e * tf32(X[col_idx[e],:])      (fp32 accumulate)
 * edge_weight == NULL  -> w_e = 1 (pattern; GCN/GIN/SAG aggregation, TCGNN_kernel.cu:336-454)
 * edge_weight != NULL  -> w_e = tf32(edge_weight[e]), CSR edge order (AGNN, :459-578)
 * X: [num_nodes, dim] row-major with leading dimension ldx (floats); Y likewise with ldy; Y is
 * fully overwritten (rows of empty windows become 0).  Any dim >= 1 (the reference drops
 * dim % 16 tails and columns >= 128). */
int tcgnn_spmm_f32(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight,
                   float* y, int64_t ldy, int32_t dim, void* stream);

/* SDDMM: edge_out[e] = sum_k tf32(X[row(e),k]) * tf32(X[col_idx[e],k])   (TCGNN_kernel.cu:584-728)
 * edge_out: [num_edges] fp32, CSR edge order, fully overwritten. */
int tcgnn_sddmm_f32(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out, int32_t dim,
                    void* stream);

/* Pre-rounded operands.  The kernels consume cvt.rna.tf32(X) (what the reference's wmma::__float_to_tf32 does per
 * use, TCGNN_kernel.cu:436-444); tcgnn_spmm_f32 / tcgnn_sddmm_f32 make that copy on every call.  A caller that
 * feeds the same X to several ops (AGNN: SDDMM + weighted SpMM) or that ships X between GPUs (row-panel
 * sharding: round the local panel once, all-gather the rounded rows) rounds once with tcgnn_round_tf32 and passes
 * TCGNN_X_IS_TF32 to the *_ex entry points.  The flag is honoured when x is 16-byte aligned, ldx % 4 == 0,
 * dim % 4 == 0 and ldx < 2^30 (otherwise the op packs a copy as usual).  out: [rows, ldo] with ldo % 4 == 0, 16-byte aligned; columns
 * [dim, ldo) are zero-filled. */
#define TCGNN_X_IS_TF32 1u
/* SpMM only: Y += A X instead of Y = A X (nothing is cleared; every window is combined with fp32 reduce-adds).  Used
 * by the sharded path, which adds one partial product per source panel as that panel's rows arrive. */
#define TCGNN_ACCUMULATE 2u
/* SpMM only: `edge_weight` is not in CSR edge order but already the plan's tile-ordered, tf32-rounded weight stream
 * ([pairs] floats, pairs = tcgnn_plan_info()[5]) -- what tcgnn_agnn_f32 leaves in `att_tile_out`. */
#define TCGNN_W_TILE_ORDER 4u
int tcgnn_round_tf32(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t dim, void* stream);
/* Same, fused with the exchange of row-panel sharding: `out` may be PEER memory (a P2P-mapped pointer into another
 * GPU's gathered matrix) -- the rounded rows are written straight over NVLink -- and tcgnn_round_tf32_multicast
 * takes an NVSwitch multicast address (multimem.st: one store reaches the buffer of every GPU of the group). */
int tcgnn_round_tf32_multicast(const float* x, int64_t ldx, float* out_mc, int64_t ldo, int64_t rows, int32_t dim,
                               void* stream);
/* Copies the row segments [seg_begin_rows[i], seg_end_rows[i]) (rows of `ld` floats, ld % 4 == 0) of the matrix at
 * `src` to the same rows of the matrices at peers[0..n_peers) -- P2P-mapped pointers into other GPUs' copies -- in
 * one launch (second phase of the balanced exchange of sharding.py).  peers / segment arrays are HOST arrays
 * (<= 16 entries each). */
int tcgnn_push_rows(const float* src, float* const* peers, int32_t n_peers, const int64_t* seg_begin_rows,
                    const int64_t* seg_end_rows, int32_t n_segs, int64_t ld, void* stream);
int tcgnn_spmm_f32_ex(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y,
                      int64_t ldy, int32_t dim, uint32_t flags, void* stream);
int tcgnn_sddmm_f32_ex(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out, int32_t dim,
                       uint32_t flags, void* stream);

/* Fused AGNN edge pipeline (reference gnn_conv.py:125-132):
 *     score[e] = <tf32(X[row(e)]), tf32(X[col(e)])>           (SDDMM, TCGNN_kernel.cu:584-728)
 *     att[e]   = score[e] * attention_w[0]                     (torch.mm with the [1, n_heads = 1] parameter)
 *     Y        = (A o tf32(att)) tf32(X)                        (weighted SpMM, TCGNN_kernel.cu:459-578)
 * in one call: X is rounded once, the SDDMM epilogue writes tf32(att) in the plan's tile order, and the weighted SpMM
 * consumes that stream directly -- no CSR-order [E] round trip, no permutation passes.  attention_w: DEVICE pointer
 * to one float (the layer's parameter; NULL = 1.0).  att_tile_out (nullable): [pairs] floats that receive tf32(att)
 * in tile order, for a backward pass via tcgnn_spmm_f32_ex(..., TCGNN_W_TILE_ORDER); NULL uses plan scratch.
 * edge_out (nullable): [num_edges] raw scores in CSR edge order (== tcgnn_sddmm_f32).  Results are bit-identical to
 * the three separate calls. */
int tcgnn_agnn_f32(tcgnn_plan* plan, const float* x, int64_t ldx, const float* attention_w, float* y, int64_t ldy,
                   float* att_tile_out, float* edge_out, int32_t dim, uint32_t flags, void* stream);

/* A^T of a device CSR (unsorted rows allowed): row_ptr_t [num_cols + 1], col_idx_t [num_edges] and, when non-null,
 * edge_map_t [num_edges] = CSR edge id in A of every edge of A^T (to carry edge weights over).  The order of the
 * entries inside a row of A^T is unspecified (SGT, the plan and every kernel here accept unsorted rows; results
 * depend only on the set of (row, col) pairs).  num_rows x num_cols is A's shape.  Synchronises `stream`. */
int tcgnn_csr_transpose(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                        int64_t num_edges, int32_t* row_ptr_t, int32_t* col_idx_t, int32_t* edge_map_t, void* stream);

/* dst[i, :] = src[rows[i], :] for i < n_rows (rows of `ld` floats, ld % 4 == 0, 16-byte aligned): packs the feature
 * rows another GPU's panel references into a contiguous block for the exchange (sharding.py). */
int tcgnn_gather_rows(const float* src, int64_t ld, const int32_t* rows, int64_t n_rows, float* dst, void* stream);

/* Stream-ordered flag wait for the overlapped exchange: blocks `stream` (one spinning thread, ld.acquire.sys) until
 * (int32)(*flag - value) >= 0.  `flag` is device memory a peer GPU's copy engine writes after its rows have landed.
 * timeout_ms > 0: gives up after that long and sets *error_out (device int32, nullable) to 1 instead of hanging. */
int tcgnn_stream_wait_flag(const int32_t* flag, int32_t value, int32_t timeout_ms, int32_t* error_out, void* stream);
/* Same, the expected value read from device memory when the wait executes (a step counter the caller bumps on the
 * device): the launch is identical every step, so the whole exchange step can be replayed as a CUDA graph. */
int tcgnn_stream_wait_flag_dev(const int32_t* flag, const int32_t* value_dev, int32_t timeout_ms, int32_t* error_out,
                               void* stream);

/* SpMM with HOST feature / result buffers (the end-to-end path of a caller whose features live in host memory;
 * page-locked buffers for full PCIe speed).  x_host: [num_cols, ldx], y_host: [num_nodes, ldy] in host memory;
 * edge_weight stays a DEVICE pointer (CSR edge order) or NULL.  The H2D copy of X and the D2H copy of Y run on
 * plan-owned copy streams around the kernels.  Ordered
 * on `stream` like every other call: it waits for work queued before it, and `stream` waits for the last copy,
 * so synchronising `stream` (or an event recorded on it) makes y_host valid.  Device staging is plan-owned. */
int tcgnn_spmm_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, const float* edge_weight, float* y_host,
                        int64_t ldy, int32_t dim, void* stream);

/* SDDMM / fused AGNN with HOST buffers, same conventions as tcgnn_spmm_f32_host: x_host [num_cols, ldx] in, results
 * out to host memory (edge_out_host [num_edges]; y_host [num_nodes, ldy]; either AGNN output may be NULL except
 * y_host).  attention_w stays a DEVICE pointer. */
int tcgnn_sddmm_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, float* edge_out_host, int32_t dim,
                         void* stream);
int tcgnn_agnn_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, const float* attention_w, float* y_host,
                        int64_t ldy, float* edge_out_host, int32_t dim, void* stream);

/* Bring-up / layout diagnostics (used by tests/test_gpu_umma_layouts.py): copies the two byte images
 * into 1024-byte aligned shared memory, issues `ksteps` tcgen05.mma.kind::tf32 (M=128) whose
 * descriptors are adesc/bdesc (start-address field 0) plus the image base plus k*a_step_bytes /
 * k*b_step_bytes, and returns the 128 x ncols fp32 accumulator (row = TMEM lane) in d_out (HOST
 * memory).  ncols must be 16, 32 or 64 and equal the N of idesc.  Synchronises the stream.
 * adesc == 0: the A operand is taken from TENSOR memory instead -- a_image is then the row-major fp32 matrix
 * [128][8 * ksteps] (a_bytes = 4096 * ksteps, ksteps <= 24), written with tcgen05.st (lane = row, k-step s in columns
 * 64 + 8 s ...); a_step_bytes is ignored and idesc must say K-major A. */
int tcgnn_debug_umma(const void* a_image, int32_t a_bytes, const void* b_image, int32_t b_bytes,
                     uint64_t adesc, uint64_t bdesc, uint32_t idesc, int32_t ksteps, int32_t a_step_bytes,
                     int32_t b_step_bytes, float* d_out, int32_t ncols, void* stream);

/* Issue-rate diagnostic (tools/umma_bench.py): `n_mma` back-to-back tcgen05.mma.kind::tf32 from one CTA per
 * grid slot, rotating over n_acc accumulators (acc_stride_cols TMEM columns apart) and n_a / n_b operand
 * tiles in zero-filled shared memory; cycles_out[0] = SM cycles until the last MMA was issued,
 * cycles_out[1] = until the commit after it arrived (block 0).  Synchronises the stream.
 * adesc == 0: A from tensor memory -- groups of 8 MMAs over n_a (<= 56) operand slots of 8 columns, B tiles 512 bytes
 * apart, one accumulator (n_acc, acc_stride_cols, a_step_bytes, n_b, b_step_bytes ignored). */
int tcgnn_debug_umma_bench(uint64_t adesc, uint64_t bdesc, uint32_t idesc, int32_t n_mma, int32_t n_acc,
                           int32_t acc_stride_cols, int32_t n_a, int32_t a_step_bytes, int32_t n_b,
                           int32_t b_step_bytes, int32_t grid, int64_t cycles_out[2], void* stream);

/* Number of kernels launched by the calling thread through this library since the last reset
 * (bench.py's gpu_launches). */
int64_t tcgnn_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* TCGNN_B200_H */
