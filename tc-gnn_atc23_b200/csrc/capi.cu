// extern "C" surface of include/tcgnn_b200.h: argument validation, error strings, launch counter.
#include <stdarg.h>
#include <stdio.h>

#include "plan.h"

namespace tcgnn {

static thread_local char g_last_error[512] = "";
static thread_local int64_t g_launches = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int sgt_cpu(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges, int32_t blk_h,
            int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col, int32_t* edge_to_row,
            int64_t* tc_blocks_out, int32_t num_threads);

}  // namespace tcgnn

using namespace tcgnn;

extern "C" {

int tcgnn_version(void) { return 100; /* 0.1.0 */ }

const char* tcgnn_status_string(int status) {
  switch (status) {
    case TCGNN_OK: return "ok";
    case TCGNN_ERR_INVALID_ARG: return "invalid argument";
    case TCGNN_ERR_CUDA: return "CUDA error";
    case TCGNN_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case TCGNN_ERR_OOM: return "out of memory";
    case TCGNN_ERR_OVERFLOW: return "32-bit index overflow";
    default: return "unknown status";
  }
}

const char* tcgnn_last_error(void) { return g_last_error; }

int64_t tcgnn_launch_count(int reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int tcgnn_sgt_cpu(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges,
                  int32_t blk_h, int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col,
                  int32_t* edge_to_row, int64_t* tc_blocks_out, int32_t num_threads) {
  return sgt_cpu(row_ptr, col_idx, num_nodes, num_edges, blk_h, blk_w, block_partition, edge_to_col, edge_to_row,
                 tc_blocks_out, num_threads);
}

int tcgnn_sgt_cuda_panel(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                         int64_t num_edges, int32_t blk_h, int32_t blk_w, int32_t* block_partition,
                         int32_t* edge_to_col, int32_t* edge_to_row, int64_t* tc_blocks_out, void* stream) {
  if (row_ptr == nullptr || block_partition == nullptr || num_rows < 0 || num_cols < 0 || num_edges < 0 ||
      blk_h <= 0 || blk_w <= 0 ||
      (num_edges > 0 && (col_idx == nullptr || edge_to_col == nullptr || edge_to_row == nullptr))) {
    set_last_error("tcgnn_sgt_cuda: bad argument");
    return TCGNN_ERR_INVALID_ARG;
  }
  return sgt_cuda(row_ptr, col_idx, num_rows, num_cols, num_edges, blk_h, blk_w, block_partition, edge_to_col,
                  edge_to_row, tc_blocks_out, static_cast<cudaStream_t>(stream));
}

int tcgnn_sgt_cuda(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges,
                   int32_t blk_h, int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col,
                   int32_t* edge_to_row, int64_t* tc_blocks_out, void* stream) {
  return tcgnn_sgt_cuda_panel(row_ptr, col_idx, num_nodes, num_nodes, num_edges, blk_h, blk_w, block_partition,
                              edge_to_col, edge_to_row, tc_blocks_out, stream);
}

int tcgnn_plan_create_panel(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                            const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_rows, int32_t num_cols,
                            int32_t row_base, int64_t num_edges, int32_t num_windows, void* stream,
                            tcgnn_plan** plan_out) {
  if (plan_out == nullptr) {
    set_last_error("tcgnn_plan_create: plan_out is null");
    return TCGNN_ERR_INVALID_ARG;
  }
  *plan_out = nullptr;
  const int64_t expect_windows = (static_cast<int64_t>(num_rows) + TCGNN_BLK_H - 1) / TCGNN_BLK_H;
  if (row_ptr == nullptr || block_partition == nullptr || num_rows <= 0 || num_edges < 0 ||
      num_edges > 0x7FFFFFFFLL || num_windows != expect_windows || row_base < -1 || num_cols < 1 ||
      (row_base >= 0 && static_cast<int64_t>(row_base) + num_rows > num_cols) ||
      (num_edges > 0 && (col_idx == nullptr || edge_to_col == nullptr || edge_to_row == nullptr))) {
    set_last_error("tcgnn_plan_create: bad argument (num_rows=%d num_cols=%d row_base=%d num_edges=%lld "
                   "num_windows=%d, expected %lld windows of %d rows)",
                   num_rows, num_cols, row_base, static_cast<long long>(num_edges), num_windows,
                   static_cast<long long>(expect_windows), TCGNN_BLK_H);
    return TCGNN_ERR_INVALID_ARG;
  }
  return plan_create(row_ptr, col_idx, block_partition, edge_to_col, edge_to_row, num_rows, num_cols, row_base,
                     num_edges, num_windows, static_cast<cudaStream_t>(stream), plan_out);
}

int tcgnn_plan_create(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                      const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_nodes, int64_t num_edges,
                      int32_t num_windows, void* stream, tcgnn_plan** plan_out) {
  return tcgnn_plan_create_panel(row_ptr, col_idx, block_partition, edge_to_col, edge_to_row, num_nodes, num_nodes, 0,
                                 num_edges, num_windows, stream, plan_out);
}

int tcgnn_plan_destroy(tcgnn_plan* plan) { return plan_destroy(plan); }

int tcgnn_plan_info(const tcgnn_plan* plan, int64_t info[8]) {
  if (plan == nullptr || info == nullptr) return TCGNN_ERR_INVALID_ARG;
  info[0] = plan->num_nodes;
  info[1] = plan->num_edges;
  info[2] = plan->num_windows;
  info[3] = plan->num_tiles;
  info[4] = static_cast<int64_t>(sizeof(TileMeta)) * (plan->num_tiles + 1) +
            4 * (static_cast<int64_t>(plan->num_windows) + 1) + (plan->eperm ? 4LL * plan->num_pairs : 0) +
            (plan->weight_perm ? 4LL * plan->num_pairs : 0) + (plan->sddmm_perm ? 4LL * plan->num_pairs : 0) +
            (plan->groups ? 16LL * plan->num_groups : 0);
  info[5] = plan->num_pairs;
  info[6] = plan->device;
  info[7] = plan->num_sms;
  return TCGNN_OK;
}

static int check_op(const char* what, const tcgnn_plan* plan, const float* x, int64_t ldx, const void* out,
                    int32_t dim) {
  if (plan == nullptr || x == nullptr || out == nullptr || dim < 1 || ldx < dim) {
    set_last_error("%s: bad argument (plan=%p x=%p out=%p dim=%d ldx=%lld)", what, (const void*)plan,
                   (const void*)x, out, dim, static_cast<long long>(ldx));
    return TCGNN_ERR_INVALID_ARG;
  }
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != plan->device) {
    set_last_error("%s: current device %d differs from the plan's device %d", what, dev, plan->device);
    return TCGNN_ERR_INVALID_ARG;
  }
  return TCGNN_OK;
}

int tcgnn_spmm_f32_ex(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y, int64_t ldy,
                      int32_t dim, uint32_t flags, void* stream) {
  int st = check_op("tcgnn_spmm_f32", plan, x, ldx, y, dim);
  if (st != TCGNN_OK) return st;
  if (ldy < dim) {
    set_last_error("tcgnn_spmm_f32: ldy < dim");
    return TCGNN_ERR_INVALID_ARG;
  }
  if ((flags & TCGNN_W_TILE_ORDER) && edge_weight == nullptr) {
    set_last_error("tcgnn_spmm_f32: TCGNN_W_TILE_ORDER without a weight array");
    return TCGNN_ERR_INVALID_ARG;
  }
  return spmm_launch(plan, x, ldx, edge_weight, y, ldy, dim, flags, static_cast<cudaStream_t>(stream));
}

int tcgnn_spmm_f32(tcgnn_plan* plan, const float* x, int64_t ldx, const float* edge_weight, float* y, int64_t ldy,
                   int32_t dim, void* stream) {
  return tcgnn_spmm_f32_ex(plan, x, ldx, edge_weight, y, ldy, dim, 0u, stream);
}

int tcgnn_spmm_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, const float* edge_weight, float* y_host,
                        int64_t ldy, int32_t dim, void* stream) {
  int st = check_op("tcgnn_spmm_f32_host", plan, x_host, ldx, y_host, dim);
  if (st != TCGNN_OK) return st;
  if (ldy < dim) {
    set_last_error("tcgnn_spmm_f32_host: ldy < dim");
    return TCGNN_ERR_INVALID_ARG;
  }
  return host_op_launch(plan, kHostSpmm, x_host, ldx, edge_weight, y_host, ldy, nullptr, dim,
                        static_cast<cudaStream_t>(stream));
}

int tcgnn_sddmm_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, float* edge_out_host, int32_t dim,
                         void* stream) {
  if (plan != nullptr && plan->num_edges == 0) return TCGNN_OK;
  int st = check_op("tcgnn_sddmm_f32_host", plan, x_host, ldx, edge_out_host, dim);
  if (st != TCGNN_OK) return st;
  return host_op_launch(plan, kHostSddmm, x_host, ldx, nullptr, nullptr, 0, edge_out_host, dim,
                        static_cast<cudaStream_t>(stream));
}

int tcgnn_agnn_f32_host(tcgnn_plan* plan, const float* x_host, int64_t ldx, const float* attention_w, float* y_host,
                        int64_t ldy, float* edge_out_host, int32_t dim, void* stream) {
  int st = check_op("tcgnn_agnn_f32_host", plan, x_host, ldx, y_host, dim);
  if (st != TCGNN_OK) return st;
  if (ldy < dim) {
    set_last_error("tcgnn_agnn_f32_host: ldy < dim");
    return TCGNN_ERR_INVALID_ARG;
  }
  return host_op_launch(plan, kHostAgnn, x_host, ldx, attention_w, y_host, ldy, edge_out_host, dim,
                        static_cast<cudaStream_t>(stream));
}

int tcgnn_agnn_f32(tcgnn_plan* plan, const float* x, int64_t ldx, const float* attention_w, float* y, int64_t ldy,
                   float* att_tile_out, float* edge_out, int32_t dim, uint32_t flags, void* stream) {
  int st = check_op("tcgnn_agnn_f32", plan, x, ldx, y, dim);
  if (st != TCGNN_OK) return st;
  if (plan->row_base < 0) {
    set_last_error("tcgnn_agnn_f32: the plan was created with row_base = -1 (its rows are not rows of X): SpMM only");
    return TCGNN_ERR_INVALID_ARG;
  }
  if (ldy < dim) {
    set_last_error("tcgnn_agnn_f32: ldy < dim");
    return TCGNN_ERR_INVALID_ARG;
  }
  return agnn_launch(plan, x, ldx, attention_w, y, ldy, att_tile_out, edge_out, dim, flags & TCGNN_X_IS_TF32,
                     static_cast<cudaStream_t>(stream));
}

int tcgnn_csr_transpose(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                        int64_t num_edges, int32_t* row_ptr_t, int32_t* col_idx_t, int32_t* edge_map_t, void* stream) {
  if (row_ptr == nullptr || row_ptr_t == nullptr || num_rows < 0 || num_cols < 0 || num_edges < 0 ||
      num_edges > 0x7FFFFFFFLL || (num_edges > 0 && (col_idx == nullptr || col_idx_t == nullptr))) {
    set_last_error("tcgnn_csr_transpose: bad argument");
    return TCGNN_ERR_INVALID_ARG;
  }
  return csr_transpose_launch(row_ptr, col_idx, num_rows, num_cols, num_edges, row_ptr_t, col_idx_t, edge_map_t,
                              static_cast<cudaStream_t>(stream));
}

int tcgnn_gather_rows(const float* src, int64_t ld, const int32_t* rows, int64_t n_rows, float* dst, void* stream) {
  if (n_rows < 0 || ld < 4 || (ld & 3) != 0 ||
      (n_rows > 0 && (src == nullptr || rows == nullptr || dst == nullptr ||
                      (reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0))) {
    set_last_error("tcgnn_gather_rows: bad argument (ld %% 4 == 0, 16-byte aligned src / dst)");
    return TCGNN_ERR_INVALID_ARG;
  }
  return gather_rows_launch(src, ld, rows, n_rows, dst, static_cast<cudaStream_t>(stream));
}

int tcgnn_stream_wait_flag(const int32_t* flag, int32_t value, int32_t timeout_ms, int32_t* error_out, void* stream) {
  if (flag == nullptr || (reinterpret_cast<uintptr_t>(flag) & 3) != 0) {
    set_last_error("tcgnn_stream_wait_flag: flag is null or misaligned");
    return TCGNN_ERR_INVALID_ARG;
  }
  return wait_flag_launch(flag, value, nullptr, timeout_ms, error_out, static_cast<cudaStream_t>(stream));
}

int tcgnn_stream_wait_flag_dev(const int32_t* flag, const int32_t* value_dev, int32_t timeout_ms, int32_t* error_out,
                               void* stream) {
  if (flag == nullptr || value_dev == nullptr || (reinterpret_cast<uintptr_t>(flag) & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(value_dev) & 3) != 0) {
    set_last_error("tcgnn_stream_wait_flag_dev: flag / value_dev is null or misaligned");
    return TCGNN_ERR_INVALID_ARG;
  }
  return wait_flag_launch(flag, 0, value_dev, timeout_ms, error_out, static_cast<cudaStream_t>(stream));
}

int tcgnn_sddmm_f32_ex(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out, int32_t dim, uint32_t flags,
                       void* stream) {
  if (plan != nullptr && plan->num_edges == 0) return TCGNN_OK;
  int st = check_op("tcgnn_sddmm_f32", plan, x, ldx, edge_out, dim);
  if (st != TCGNN_OK) return st;
  if (plan->row_base < 0) {
    set_last_error("tcgnn_sddmm_f32: the plan was created with row_base = -1 (its rows are not rows of X): SpMM only");
    return TCGNN_ERR_INVALID_ARG;
  }
  return sddmm_launch(plan, x, ldx, edge_out, nullptr, nullptr, dim, flags & TCGNN_X_IS_TF32,
                      static_cast<cudaStream_t>(stream));
}

int tcgnn_sddmm_f32(tcgnn_plan* plan, const float* x, int64_t ldx, float* edge_out, int32_t dim, void* stream) {
  return tcgnn_sddmm_f32_ex(plan, x, ldx, edge_out, dim, 0u, stream);
}

int tcgnn_push_rows(const float* src, float* const* peers, int32_t n_peers, const int64_t* seg_begin_rows,
                    const int64_t* seg_end_rows, int32_t n_segs, int64_t ld, void* stream) {
  if (src == nullptr || peers == nullptr || n_peers < 0 || n_peers > 16 || n_segs < 0 || n_segs > 16 ||
      (n_segs > 0 && (seg_begin_rows == nullptr || seg_end_rows == nullptr)) || ld < 4 || (ld & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(src) & 15) != 0) {
    set_last_error("tcgnn_push_rows: bad argument (<= 16 peers, <= 16 segments, ld %% 4 == 0, 16-byte aligned bases)");
    return TCGNN_ERR_INVALID_ARG;
  }
  for (int i = 0; i < n_peers; ++i)
    if (peers[i] == nullptr || (reinterpret_cast<uintptr_t>(peers[i]) & 15) != 0) {
      set_last_error("tcgnn_push_rows: peer %d is null or not 16-byte aligned", i);
      return TCGNN_ERR_INVALID_ARG;
    }
  for (int i = 0; i < n_segs; ++i)
    if (seg_begin_rows[i] < 0 || seg_end_rows[i] < seg_begin_rows[i]) {
      set_last_error("tcgnn_push_rows: bad segment %d", i);
      return TCGNN_ERR_INVALID_ARG;
    }
  return push_rows_launch(src, peers, n_peers, seg_begin_rows, seg_end_rows, n_segs, ld, static_cast<cudaStream_t>(stream));
}

static int round_tf32_checked(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t dim,
                              int multimem, void* stream) {
  if (x == nullptr || out == nullptr || rows < 0 || dim < 1 || ldx < dim || ldo < dim || (ldo & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(out) & 15) != 0) {
    set_last_error("tcgnn_round_tf32: bad argument (out must be 16-byte aligned with ldo %% 4 == 0)");
    return TCGNN_ERR_INVALID_ARG;
  }
  return round_tf32_launch(x, ldx, out, ldo, rows, dim, multimem, static_cast<cudaStream_t>(stream));
}

int tcgnn_round_tf32(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t dim, void* stream) {
  return round_tf32_checked(x, ldx, out, ldo, rows, dim, 0, stream);
}

int tcgnn_round_tf32_multicast(const float* x, int64_t ldx, float* out_mc, int64_t ldo, int64_t rows, int32_t dim,
                               void* stream) {
  return round_tf32_checked(x, ldx, out_mc, ldo, rows, dim, 1, stream);
}

int tcgnn_debug_umma(const void* a_image, int32_t a_bytes, const void* b_image, int32_t b_bytes, uint64_t adesc,
                     uint64_t bdesc, uint32_t idesc, int32_t ksteps, int32_t a_step_bytes, int32_t b_step_bytes,
                     float* d_out, int32_t ncols, void* stream) {
  return debug_umma(a_image, a_bytes, b_image, b_bytes, adesc, bdesc, idesc, ksteps, a_step_bytes, b_step_bytes,
                    d_out, ncols, static_cast<cudaStream_t>(stream));
}

int tcgnn_debug_umma_bench(uint64_t adesc, uint64_t bdesc, uint32_t idesc, int32_t n_mma, int32_t n_acc,
                           int32_t acc_stride_cols, int32_t n_a, int32_t a_step_bytes, int32_t n_b,
                           int32_t b_step_bytes, int32_t grid, int64_t cycles_out[2], void* stream) {
  return debug_umma_bench(adesc, bdesc, idesc, n_mma, n_acc, acc_stride_cols, n_a, a_step_bytes, n_b, b_step_bytes,
                          grid, cycles_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
