// Graph utilities around the aggregation kernels (all integer / byte work, HBM- or latency-bound; nothing here is
// reshaped into a GEMM):
//   * csr_transpose      A^T of a device CSR, so that the backward pass can aggregate over the transposed graph.  The
//                        reference re-uses the forward CSR for dX (gnn_conv.py:76-85), which is only correct for a
//                        symmetric adjacency.
//   * column-chunk plans the sub-graph of a plan restricted to a column range [c0, c1): Y = sum_j A[:, chunk j] X[chunk j].
//                        The host-buffer entry points start aggregating chunk j as soon as its rows of X have arrived
//                        over PCIe (host_entry.cu); the sharded path does the same per source GPU (sharding.py).
//   * gather_rows        packs the feature rows another GPU's panel references into one contiguous block.
//   * wait_flag          stream-ordered wait on a flag a peer GPU's copy engine writes after its rows have landed.
#include "scan.cuh"

namespace tcgnn {

namespace {

int grid_for(int64_t n, int threads, int cap = 148 * 32) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}

// ---------------------------------------------------------------------------------------------
// transpose
// ---------------------------------------------------------------------------------------------
__global__ void count_cols_kernel(const int32_t* __restrict__ col_idx, int64_t num_edges, int32_t num_cols,
                                  int32_t* __restrict__ counts, int32_t* __restrict__ bad) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < num_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t c = col_idx[e];
    if (c < 0 || c >= num_cols) { atomicAdd(bad, 1); continue; }
    atomicAdd(&counts[c], 1);
  }
}

// One thread per edge of A (its row by binary search, balanced for power-law rows): claim the next free slot of
// row col(e) of A^T.  The order inside a row of A^T is whatever the atomics give: SGT and the plan builder accept
// unsorted rows, and the tile stream (hence every result) depends only on the set of (row, col) pairs.
__global__ void scatter_transpose_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                         int32_t num_rows, int32_t num_cols, int64_t num_edges,
                                         const int32_t* __restrict__ row_ptr_t, int32_t* __restrict__ cursor,
                                         int32_t* __restrict__ col_idx_t, int32_t* __restrict__ edge_map_t) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < num_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t c = col_idx[e];
    if (c < 0 || c >= num_cols) continue;
    int32_t lo = 0, hi = num_rows;   // last row r with row_ptr[r] <= e
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (row_ptr[mid] <= e) lo = mid; else hi = mid;
    }
    const int32_t pos = row_ptr_t[c] + atomicAdd(&cursor[c], 1);
    col_idx_t[pos] = lo;
    if (edge_map_t != nullptr) edge_map_t[pos] = static_cast<int32_t>(e);
  }
}

// ---------------------------------------------------------------------------------------------
// column-chunk sub-graphs
// ---------------------------------------------------------------------------------------------
// one warp per row: edges with c0 <= col < c1
__global__ void count_in_range_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                      int32_t num_rows, int32_t c0, int32_t c1, int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < num_rows; r += nwarps) {
    int32_t n = 0;
    for (int32_t e = row_ptr[r] + lane; e < row_ptr[r + 1]; e += 32) {
      const int32_t c = col_idx[e];
      n += (c >= c0 && c < c1) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) counts[r] = n;
  }
}
// one warp per row: order-preserving compaction (ballot + prefix popcount), column ids rebased to c0
__global__ void compact_in_range_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                        int32_t num_rows, int32_t c0, int32_t c1, const int32_t* __restrict__ sub_ptr,
                                        int32_t* __restrict__ sub_col) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < num_rows; r += nwarps) {
    int32_t out = sub_ptr[r];
    const int32_t e1 = row_ptr[r + 1];
    for (int32_t e0 = row_ptr[r]; e0 < e1; e0 += 32) {
      const int32_t e = e0 + lane;
      const int32_t c = e < e1 ? col_idx[e] : -1;
      const bool keep = c >= c0 && c < c1;
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (keep) sub_col[out + __popc(m & ((1u << lane) - 1u))] = c - c0;
      out += __popc(m);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exchange helpers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float4* __restrict__ src, int64_t ldv, const int32_t* __restrict__ rows, int64_t n_rows,
                   float4* __restrict__ dst) {
  const int64_t total = n_rows * ldv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / ldv;
    const int64_t v = i - r * ldv;
    dst[i] = __ldg(src + static_cast<int64_t>(rows[r]) * ldv + v);
  }
}

// value_dev != nullptr: the expected value is read from device memory when the kernel runs (a step counter the
// caller bumps on the device), so the launch is the same every step and can live in a CUDA graph
__global__ void wait_flag_kernel(const int32_t* flag, int32_t value, const int32_t* value_dev,
                                 unsigned long long timeout_ns, int32_t* error_out) {
  if (value_dev != nullptr) value = *value_dev;
  const uint64_t t0 = global_timer_ns();
  for (;;) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v - value >= 0) return;
    __nanosleep(256);
    if (timeout_ns != 0 && global_timer_ns() - t0 > timeout_ns) {
      if (error_out != nullptr) *error_out = 1;
      return;   // never hang the stream: the caller checks error_out
    }
  }
}

}  // namespace

#define GOPS_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_last_error("%s failed: %s", #expr, cudaGetErrorString(_e));                    \
      status = (_e == cudaErrorMemoryAllocation) ? TCGNN_ERR_OOM : TCGNN_ERR_CUDA;       \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)

int csr_transpose_launch(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_rows, int32_t num_cols,
                         int64_t num_edges, int32_t* row_ptr_t, int32_t* col_idx_t, int32_t* edge_map_t,
                         cudaStream_t stream) {
  int status = TCGNN_OK;
  int32_t* counts = nullptr;    // [num_cols] in-degrees, then the per-row cursors; + 1 error counter
  int32_t* scratch = nullptr;
  int32_t bad = 0;
  const int64_t nblk = (static_cast<int64_t>(num_cols) + kScanTile - 1) / kScanTile;
  GOPS_CUDA(cudaMalloc(&counts, sizeof(int32_t) * (static_cast<size_t>(num_cols) + 1)));
  GOPS_CUDA(cudaMalloc(&scratch, sizeof(int32_t) * (static_cast<size_t>(nblk) + 2)));
  GOPS_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (static_cast<size_t>(num_cols) + 1), stream));
  if (num_edges > 0) {
    count_cols_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(col_idx, num_edges, num_cols, counts,
                                                                   counts + num_cols);
    count_launch();
  }
  GOPS_CUDA(exclusive_scan(LoadI32{counts}, num_cols, row_ptr_t, scratch, stream));
  GOPS_CUDA(cudaMemcpyAsync(&bad, counts + num_cols, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GOPS_CUDA(cudaStreamSynchronize(stream));
  if (bad != 0) {
    set_last_error("tcgnn_csr_transpose: %d column ids outside [0, %d)", bad, num_cols);
    status = TCGNN_ERR_INVALID_ARG;
    goto done;
  }
  GOPS_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * static_cast<size_t>(num_cols), stream));
  if (num_edges > 0) {
    scatter_transpose_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(row_ptr, col_idx, num_rows, num_cols,
                                                                          num_edges, row_ptr_t, counts, col_idx_t,
                                                                          edge_map_t);
    count_launch();
    GOPS_CUDA(cudaGetLastError());
    GOPS_CUDA(cudaStreamSynchronize(stream));   // `counts` is freed below
  }
done:
  if (counts) cudaFree(counts);
  if (scratch) cudaFree(scratch);
  return status;
}

int plan_create_column_chunk(const tcgnn_plan* parent, int32_t c0, int32_t c1, cudaStream_t stream,
                             tcgnn_plan** plan_out) {
  int status = TCGNN_OK;
  *plan_out = nullptr;
  const int32_t n = parent->num_nodes;
  const int32_t nwin = parent->num_windows;
  int32_t* counts = nullptr;
  int32_t* scratch = nullptr;
  int32_t* arr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // sub_ptr, sub_col, bp, e2c, e2r
  int32_t total = 0;
  const int64_t nblk = (static_cast<int64_t>(n) + kScanTile - 1) / kScanTile;
  tcgnn_plan* sub = nullptr;
  GOPS_CUDA(cudaMalloc(&counts, sizeof(int32_t) * static_cast<size_t>(n)));
  GOPS_CUDA(cudaMalloc(&scratch, sizeof(int32_t) * (static_cast<size_t>(nblk) + 2)));
  GOPS_CUDA(cudaMalloc(&arr[0], sizeof(int32_t) * (static_cast<size_t>(n) + 1)));
  count_in_range_kernel<<<grid_for(static_cast<int64_t>(n) * 32, 256), 256, 0, stream>>>(parent->row_ptr, parent->col_idx,
                                                                                         n, c0, c1, counts);
  count_launch();
  GOPS_CUDA(exclusive_scan(LoadI32{counts}, n, arr[0], scratch, stream));
  GOPS_CUDA(cudaMemcpyAsync(&total, arr[0] + n, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GOPS_CUDA(cudaStreamSynchronize(stream));
  {
    const size_t ne = static_cast<size_t>(total > 0 ? total : 1);
    GOPS_CUDA(cudaMalloc(&arr[1], sizeof(int32_t) * ne));
    GOPS_CUDA(cudaMalloc(&arr[2], sizeof(int32_t) * static_cast<size_t>(nwin)));
    GOPS_CUDA(cudaMalloc(&arr[3], sizeof(int32_t) * ne));
    GOPS_CUDA(cudaMalloc(&arr[4], sizeof(int32_t) * ne));
  }
  if (total > 0) {
    compact_in_range_kernel<<<grid_for(static_cast<int64_t>(n) * 32, 256), 256, 0, stream>>>(
        parent->row_ptr, parent->col_idx, n, c0, c1, arr[0], arr[1]);
    count_launch();
  }
  GOPS_CUDA(cudaGetLastError());
  status = sgt_cuda(arr[0], arr[1], n, c1 - c0, total, TCGNN_BLK_H, TCGNN_BLK_W, arr[2], arr[3], arr[4], nullptr, stream);
  if (status != TCGNN_OK) goto done;
  status = plan_create(arr[0], arr[1], arr[2], arr[3], arr[4], n, c1 - c0, /*row_base: SpMM only*/ -1, total, nwin, stream,
                       &sub);
  if (status != TCGNN_OK) goto done;
  for (int i = 0; i < 5; ++i) {
    sub->owned[i] = arr[i];
    arr[i] = nullptr;
  }
  *plan_out = sub;
done:
  if (counts) cudaFree(counts);
  if (scratch) cudaFree(scratch);
  for (int32_t* a : arr)
    if (a) cudaFree(a);
  return status;
}

int gather_rows_launch(const float* src, int64_t ld, const int32_t* rows, int64_t n_rows, float* dst,
                       cudaStream_t stream) {
  if (n_rows <= 0) return TCGNN_OK;
  const int64_t ldv = ld >> 2;
  gather_rows_kernel<<<grid_for(n_rows * ldv, 256, 148 * 8), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(src), ldv, rows, n_rows, reinterpret_cast<float4*>(dst));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("gather_rows_kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  return TCGNN_OK;
}

int wait_flag_launch(const int32_t* flag, int32_t value, const int32_t* value_dev, int32_t timeout_ms,
                     int32_t* error_out, cudaStream_t stream) {
  wait_flag_kernel<<<1, 1, 0, stream>>>(flag, value, value_dev,
                                        timeout_ms > 0 ? static_cast<unsigned long long>(timeout_ms) * 1000000ULL : 0ULL,
                                        error_out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("wait_flag_kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  return TCGNN_OK;
}

}  // namespace tcgnn
