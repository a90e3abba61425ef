// Host-side plan object behind the opaque `tcgnn_plan` of include/tcgnn_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "common.cuh"

struct tcgnn_plan {
  // borrowed device pointers of the caller's graph (must outlive the plan)
  const int32_t* row_ptr = nullptr;
  const int32_t* col_idx = nullptr;
  const int32_t* edge_to_col = nullptr;
  const int32_t* edge_to_row = nullptr;
  int32_t num_nodes = 0;   // rows covered by the plan (panel rows)
  int32_t num_cols = 0;    // rows of X (== num_nodes unless the plan is a row panel of a larger graph)
  int32_t row_base = 0;    // global id of the plan's row 0
  int64_t num_edges = 0;
  int32_t num_windows = 0;
  int32_t num_tiles = 0;
  int32_t num_pairs = 0;  // distinct (row, col) pairs == set bits over all tile masks
  int device = 0;
  int num_sms = 0;
  // owned device memory
  tcgnn::TileMeta* tiles = nullptr;   // [num_tiles + 1]
  int32_t* win_tile_ptr = nullptr;    // [num_windows + 1]
  int32_t* slice_ptr = nullptr;       // [grid + 1] tile range per persistent CTA
  int grid = 1;                       // persistent CTAs the kernels are launched with
  int32_t* eperm = nullptr;           // [num_pairs]   lazy (weighted SpMM / SDDMM)
  float* weight_perm = nullptr;       // [num_pairs]   lazy: edge weights in tile order
  float* sddmm_perm = nullptr;        // [num_pairs]   lazy: SDDMM scores in tile order (before the CSR permutation)
  float* x_round = nullptr;           // lazy, grows: tf32-rounded, 16B-row-aligned copy of the current X
  size_t x_round_cap = 0;             // floats
  int4* groups = nullptr;             // [num_groups]  lazy: SDDMM work units {tile_start, ntiles, win, 0}
  int32_t num_groups = 0;
  int32_t* flag = nullptr;            // device error counter
  // launches over a window sub-range (host-buffer pipeline): chunk r covers the windows
  // [row_chunk_win[r], row_chunk_win[r+1]) and has its own balanced CTA slices
  std::vector<int32_t> row_chunk_win;   // host copy, R + 1 entries (empty: no row chunks)
  int32_t* chunk_slice_ptr = nullptr;   // device [R][grid + 1]
  // CSR / SGT arrays a derived plan owns itself (column-chunk sub-plans); freed with the plan
  void* owned[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // host-buffer entry points (tcgnn_*_f32_host): device staging + copy streams + the chunk pipeline, all lazy
  float* host_x_dev = nullptr;        // [num_cols * dim]
  float* host_y_dev = nullptr;        // [num_nodes * dim]
  float* host_e_dev = nullptr;        // [num_edges]
  size_t host_x_cap = 0, host_y_cap = 0, host_e_cap = 0;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> host_ev;
  std::vector<tcgnn_plan*> col_chunks;  // sub-plans over the column ranges [col_chunk_bounds[j], col_chunk_bounds[j+1])
  std::vector<int32_t> col_chunk_bounds;
  bool host_pipeline_tried = false;
  std::mutex mu;                      // guards the lazy members
  std::mutex host_mu;                 // serialises the host-buffer entry points of one plan

  tcgnn::PlanView view() const {
    tcgnn::PlanView v;
    v.tiles = tiles;
    v.win_tile_ptr = win_tile_ptr;
    v.eperm = eperm;
    v.slice_ptr = slice_ptr;
    v.num_nodes = num_nodes;
    v.num_cols = num_cols;
    v.row_base = row_base;
    v.num_windows = num_windows;
    v.num_tiles = num_tiles;
    v.num_pairs = num_pairs;
    return v;
  }
};

namespace tcgnn {

void set_last_error(const char* fmt, ...);
void count_launch(int n = 1);

int plan_create(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_nodes, int32_t num_cols,
                int32_t row_base, int64_t num_edges, int32_t num_windows, cudaStream_t stream,
                tcgnn_plan** plan_out);
int plan_destroy(tcgnn_plan* plan);
int plan_ensure_eperm(tcgnn_plan* plan, cudaStream_t stream);
int plan_ensure_scratch(tcgnn_plan* plan, float** slot, size_t count);
int plan_ensure_groups(tcgnn_plan* plan, cudaStream_t stream);   // synchronises the stream on first use
// Xr = cvt.rna.tf32(X[:, :dim]) packed as [num_cols, ldr] (ldr % 4 == 0) into the plan's scratch (round_pack.cu).
// Ops on one plan must be stream-ordered (they share this scratch).
int round_pack_launch(tcgnn_plan* plan, const float* x, int64_t ldx, int32_t dim, int64_t ldr, cudaStream_t stream,
                      const float** xr_out);

// kernels (spmm_tc.cu / sddmm_tc.cu / sgt_gpu.cu / umma_probe.cu)
// row_chunk >= 0: only the windows of that row chunk (plan_set_row_chunks); -1: the whole plan
int spmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y, int64_t ldy,
                int32_t dim, uint32_t op_flags, cudaStream_t stream, int row_chunk = -1);
// balanced CTA slices for launches over the window ranges [win_bounds[r], win_bounds[r+1]), r < n_chunks
int plan_set_row_chunks(tcgnn_plan* plan, const int32_t* win_bounds, int n_chunks, cudaStream_t stream);
// sub-plan of `parent` restricted to the columns [c0, c1) (ids rebased to c0): its X operand is rows [c0, c1) of the
// parent's.  Owns its filtered CSR and SGT arrays.  Synchronises the stream.
int plan_create_column_chunk(const tcgnn_plan* parent, int32_t c0, int32_t c1, cudaStream_t stream,
                             tcgnn_plan** plan_out);
int csr_transpose_launch(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                         int64_t num_edges, int32_t* row_ptr_t, int32_t* col_idx_t, int32_t* edge_map_t,
                         cudaStream_t stream);
int gather_rows_launch(const float* src, int64_t ld, const int32_t* rows, int64_t n_rows, float* dst,
                       cudaStream_t stream);
int wait_flag_launch(const int32_t* flag, int32_t value, const int32_t* value_dev, int32_t timeout_ms,
                     int32_t* error_out, cudaStream_t stream);
enum HostOp { kHostSpmm = 0, kHostSddmm = 1, kHostAgnn = 2 };
int host_op_launch(tcgnn_plan* plan, int op, const float* x_host, int64_t ldx, const float* dev_arg, float* y_host,
                   int64_t ldy, float* e_host, int32_t dim, cudaStream_t stream);
int sddmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out_csr, float* tile_out,
                 const float* scale, int32_t dim, uint32_t op_flags, cudaStream_t stream);
int agnn_launch(tcgnn_plan* plan, const float* x, int64_t ldx, const float* attention_w, float* y, int64_t ldy,
                float* att_tile_out, float* edge_out_csr, int32_t dim, uint32_t op_flags, cudaStream_t stream);
// TCGNN_X_IS_TF32 is honoured only for 16-byte aligned rows of a width that is a multiple of 4: otherwise the last
// 16-byte vector of a row would pull in whatever follows the first `dim` columns in the caller's matrix (columns
// that SDDMM would contract over), so the op packs its own zero-padded copy instead.  The same for a row stride of
// 2^30 elements or more: the kernels address gathered rows with a 32-bit byte stride.
inline bool x_is_prerounded(const float* x, int64_t ldx, int32_t dim, uint32_t op_flags) {
  return (op_flags & TCGNN_X_IS_TF32) != 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 3) == 0 &&
         (dim & 3) == 0 && ldx < (int64_t{1} << 30);
}
int push_rows_launch(const float* src, float* const* peers, int32_t n_peers, const int64_t* seg_begin_rows,
                     const int64_t* seg_end_rows, int32_t n_segs, int64_t ld, cudaStream_t stream);
int round_tf32_launch(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t dim, int multimem,
                      cudaStream_t stream);
int sgt_cuda(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int32_t num_cols, int64_t num_edges,
             int32_t blk_h,
             int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col, int32_t* edge_to_row,
             int64_t* tc_blocks_out, cudaStream_t stream);
int debug_umma(const void* a_image, int32_t a_bytes, const void* b_image, int32_t b_bytes, uint64_t adesc,
               uint64_t bdesc, uint32_t idesc, int32_t ksteps, int32_t a_step_bytes, int32_t b_step_bytes,
               float* d_out, int32_t ncols, cudaStream_t stream);
int debug_umma_bench(uint64_t adesc, uint64_t bdesc, uint32_t idesc, int32_t n_mma, int32_t n_acc, int32_t acc_stride,
                     int32_t n_a, int32_t a_step, int32_t n_b, int32_t b_step, int32_t grid, int64_t* cycles_out,
                     cudaStream_t stream);

}  // namespace tcgnn
