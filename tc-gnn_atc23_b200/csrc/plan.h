// Host-side plan object behind the opaque `tcgnn_plan` of include/tcgnn_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <mutex>

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
  float* sddmm_perm = nullptr;        // [num_pairs]   lazy: SDDMM results in tile order
  float* x_round = nullptr;           // lazy, grows: tf32-rounded, 16B-row-aligned copy of the current X
  size_t x_round_cap = 0;             // floats
  int4* groups = nullptr;             // [num_groups]  lazy: SDDMM work units {tile_start, ntiles, win, 0}
  int32_t num_groups = 0;
  int32_t* flag = nullptr;            // device error counter
  // host-buffer entry point (tcgnn_spmm_f32_host): device staging + copy streams, lazy
  float* host_x_dev = nullptr;        // [num_cols * dim] column chunks, packed
  float* host_y_dev = nullptr;        // [num_nodes * dim]
  size_t host_x_cap = 0, host_y_cap = 0;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t host_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // start, h2d[2], kernel[2], done
  std::mutex mu;                      // guards the lazy members

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
int spmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y, int64_t ldy,
                int32_t dim, uint32_t op_flags, cudaStream_t stream);
int sddmm_launch(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out, int32_t dim, uint32_t op_flags,
                 cudaStream_t stream);
int push_rows_launch(const float* src, float* const* peers, int32_t n_peers, const int64_t* seg_begin_rows,
                     const int64_t* seg_end_rows, int32_t n_segs, int64_t ld, cudaStream_t stream);
int spmm_host_launch(tcgnn_plan* plan, const float* x_host, int64_t ldx, const float* edge_weight, float* y_host,
                     int64_t ldy, int32_t dim, cudaStream_t stream);
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
