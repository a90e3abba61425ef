// GPU SGT (sparse-graph translation).  The reference only has a stub here: `preprocess_gpu`
// (/root/reference TCGNN_conv/TCGNN.cpp:229-256) fills edgeToRow (TCGNN_kernel.cu:21-40) and its
// `fill_window` body is commented out (:42-80).  This implements the full `preprocess` semantics
// (TCGNN.cpp:172-226) on the device, bit-exact with the host version:
//   * edge_to_row: one thread per edge, binary search in row_ptr (balanced for power-law rows);
//   * windows with <= kSmallMax edges: one CTA per window, bitonic sort + dedup + rank in smem;
//   * larger windows (hub rows): queued and handled by persistent CTAs with a private bitmap over
//     the column ids: set bits, prefix-popcount, rank(c) = prefix[c/32] + popc(bits below c).
// Integer-only and bandwidth/latency bound; no tensor cores involved.
#include "plan.h"

namespace tcgnn {

namespace {

constexpr int kSmallThreads = 256;
constexpr int kSmallMax = 4096;        // keys sorted in shared memory per window
constexpr int kLargeThreads = 1024;

__global__ void edge_to_row_kernel(const int32_t* __restrict__ row_ptr, int32_t num_nodes, int64_t num_edges,
                                   int32_t* __restrict__ edge_to_row) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < num_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // last row r with row_ptr[r] <= e  (rows with no edges are skipped naturally)
    int32_t lo = 0, hi = num_nodes;
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (row_ptr[mid] <= e) lo = mid; else hi = mid;
    }
    edge_to_row[e] = lo;
  }
}

__global__ void __launch_bounds__(kSmallThreads)
sgt_small_windows_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                         int32_t num_nodes, int32_t num_windows, int32_t blk_h, int32_t blk_w,
                         int32_t* __restrict__ block_partition, int32_t* __restrict__ edge_to_col,
                         int32_t* __restrict__ large_list, int32_t* __restrict__ large_count,
                         unsigned long long* __restrict__ total_blocks) {
  __shared__ uint32_t keys[kSmallMax];
  __shared__ uint32_t uniq[kSmallMax];
  __shared__ int32_t warp_sums[kSmallThreads / 32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x;
  for (int32_t w = blockIdx.x; w < num_windows; w += gridDim.x) {
    const int64_t r0 = static_cast<int64_t>(w) * blk_h;
    const int64_t r1 = min(r0 + blk_h, static_cast<int64_t>(num_nodes));
    const int32_t s = row_ptr[r0], t = row_ptr[r1];
    const int32_t len = t - s;
    if (len > kSmallMax) {
      if (tid == 0) large_list[atomicAdd(large_count, 1)] = w;
      continue;
    }
    if (len == 0) {
      if (tid == 0) {
        block_partition[w] = 1;   // the reference's empty-window artefact (TCGNN.cpp:160,216)
        atomicAdd(total_blocks, 1ull);
      }
      continue;
    }
    int32_t p2 = 1;
    while (p2 < len) p2 <<= 1;
    for (int i = tid; i < p2; i += kSmallThreads) keys[i] = i < len ? static_cast<uint32_t>(col_idx[s + i]) : 0xFFFFFFFFu;
    __syncthreads();
    // bitonic sort, ascending, ids compared as unsigned (TCGNN.cpp:205-209)
    for (int32_t k = 2; k <= p2; k <<= 1) {
      for (int32_t j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < p2; i += kSmallThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const uint32_t a = keys[i], b = keys[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    // dedup: exclusive scan of head flags over the first `len` sorted keys
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < len; base += kSmallThreads) {
      const int i = base + tid;
      const int32_t flag = (i < len && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
      int32_t incl = flag;
#pragma unroll
      for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, ofs);
        if ((tid & 31) >= ofs) incl += v;
      }
      if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
      __syncthreads();
      int32_t prefix = carry_s;
      for (int q = 0; q < (tid >> 5); ++q) prefix += warp_sums[q];
      if (flag) uniq[prefix + incl - 1] = keys[i];
      __syncthreads();
      if (tid == kSmallThreads - 1) carry_s = prefix + incl;
      __syncthreads();
    }
    const int32_t nu = carry_s;
    if (tid == 0) {
      const int32_t bp = (max(nu, 1) + blk_w - 1) / blk_w;   // TCGNN.cpp:216
      block_partition[w] = bp;
      atomicAdd(total_blocks, static_cast<unsigned long long>(bp));
    }
    for (int i = tid; i < len; i += kSmallThreads) {          // TCGNN.cpp:220-223
      const uint32_t key = static_cast<uint32_t>(col_idx[s + i]);
      int32_t lo = 0, hi = nu;
      while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (uniq[mid] < key) lo = mid + 1; else hi = mid;
      }
      edge_to_col[s + i] = lo;
    }
    __syncthreads();
  }
}

// Persistent CTAs; CTA b owns bitmap/prefix slices [b * words, (b+1) * words).
__global__ void __launch_bounds__(kLargeThreads)
sgt_large_windows_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                         int32_t num_nodes, int32_t num_cols, int32_t blk_h, int32_t blk_w, const int32_t* __restrict__ large_list,
                         const int32_t* __restrict__ large_count, uint32_t* __restrict__ bitmaps,
                         uint32_t* __restrict__ prefixes, int32_t words, int32_t* __restrict__ block_partition,
                         int32_t* __restrict__ edge_to_col, unsigned long long* __restrict__ total_blocks,
                         int32_t* __restrict__ bad) {
  __shared__ uint32_t warp_sums[kLargeThreads / 32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x;
  uint32_t* bits = bitmaps + static_cast<size_t>(blockIdx.x) * words;
  uint32_t* pre = prefixes + static_cast<size_t>(blockIdx.x) * words;
  const int32_t n_large = *large_count;
  for (int32_t li = blockIdx.x; li < n_large; li += gridDim.x) {
    const int32_t w = large_list[li];
    const int64_t r0 = static_cast<int64_t>(w) * blk_h;
    const int64_t r1 = min(r0 + blk_h, static_cast<int64_t>(num_nodes));
    const int32_t s = row_ptr[r0], t = row_ptr[r1];
    for (int i = tid; i < words; i += kLargeThreads) bits[i] = 0u;
    __syncthreads();
    for (int32_t e = s + tid; e < t; e += kLargeThreads) {
      const uint32_t c = static_cast<uint32_t>(col_idx[e]);
      if (c >= static_cast<uint32_t>(num_cols)) { atomicAdd(bad, 1); continue; }
      atomicOr(&bits[c >> 5], 1u << (c & 31));
    }
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < words; base += kLargeThreads) {
      const int i = base + tid;
      const uint32_t cnt = i < words ? __popc(bits[i]) : 0u;
      uint32_t incl = cnt;
#pragma unroll
      for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, ofs);
        if ((tid & 31) >= ofs) incl += v;
      }
      if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
      __syncthreads();
      uint32_t prefix = carry_s;
      for (int q = 0; q < (tid >> 5); ++q) prefix += warp_sums[q];
      if (i < words) pre[i] = prefix + incl - cnt;
      __syncthreads();
      if (tid == kLargeThreads - 1) carry_s = prefix + incl;
      __syncthreads();
    }
    const int32_t nu = static_cast<int32_t>(carry_s);
    if (tid == 0) {
      const int32_t bp = (max(nu, 1) + blk_w - 1) / blk_w;
      block_partition[w] = bp;
      atomicAdd(total_blocks, static_cast<unsigned long long>(bp));
    }
    for (int32_t e = s + tid; e < t; e += kLargeThreads) {
      const uint32_t c = static_cast<uint32_t>(col_idx[e]);
      if (c >= static_cast<uint32_t>(num_cols)) continue;
      edge_to_col[e] = static_cast<int32_t>(pre[c >> 5] + __popc(bits[c >> 5] & ((1u << (c & 31)) - 1u)));
    }
    __syncthreads();
  }
}

}  // namespace

int sgt_cuda(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int32_t num_cols, int64_t num_edges,
             int32_t blk_h,
             int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col, int32_t* edge_to_row,
             int64_t* tc_blocks_out, cudaStream_t stream) {
  const int64_t num_windows64 = (static_cast<int64_t>(num_nodes) + blk_h - 1) / blk_h;
  const int32_t num_windows = static_cast<int32_t>(num_windows64);
  if (num_windows == 0) {
    if (tc_blocks_out) *tc_blocks_out = 1;   // N == 0: the reference still runs one empty loop trip
    return TCGNN_OK;
  }
  int status = TCGNN_ERR_CUDA;
  cudaError_t e = cudaSuccess;
  int32_t* large_list = nullptr;     // [num_windows] + count + bad
  unsigned long long* total = nullptr;
  uint32_t* bitmaps = nullptr;
  int32_t host_counts[2] = {0, 0};
  unsigned long long host_total = 0;
  const int32_t words = (num_cols + 31) / 32;   // hub-window bitmap spans the column id range
  int sms = 148;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) goto done;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if ((e = cudaMalloc(&large_list, sizeof(int32_t) * (static_cast<size_t>(num_windows) + 2))) != cudaSuccess) goto done;
  if ((e = cudaMalloc(&total, sizeof(unsigned long long))) != cudaSuccess) goto done;
  if ((e = cudaMemsetAsync(large_list + num_windows, 0, 2 * sizeof(int32_t), stream)) != cudaSuccess) goto done;
  if ((e = cudaMemsetAsync(total, 0, sizeof(unsigned long long), stream)) != cudaSuccess) goto done;
  if (num_edges > 0) {
    int64_t g = (num_edges + 255) / 256;
    if (g > sms * 32) g = sms * 32;
    edge_to_row_kernel<<<static_cast<int>(g), 256, 0, stream>>>(row_ptr, num_nodes, num_edges, edge_to_row);
    count_launch();
  }
  {
    const int grid = num_windows < sms * 64 ? num_windows : sms * 64;
    sgt_small_windows_kernel<<<grid, kSmallThreads, 0, stream>>>(row_ptr, col_idx, num_nodes, num_windows, blk_h,
                                                                 blk_w, block_partition, edge_to_col, large_list,
                                                                 large_list + num_windows, total);
    count_launch();
  }
  if ((e = cudaGetLastError()) != cudaSuccess) goto done;
  if ((e = cudaMemcpyAsync(host_counts, large_list + num_windows, sizeof(int32_t), cudaMemcpyDeviceToHost, stream)) !=
      cudaSuccess)
    goto done;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) goto done;
  if (host_counts[0] > 0) {
    int grid = host_counts[0] < sms * 2 ? host_counts[0] : sms * 2;
    if ((e = cudaMalloc(&bitmaps, sizeof(uint32_t) * 2 * static_cast<size_t>(grid) * words)) != cudaSuccess) goto done;
    sgt_large_windows_kernel<<<grid, kLargeThreads, 0, stream>>>(
        row_ptr, col_idx, num_nodes, num_cols, blk_h, blk_w, large_list, large_list + num_windows, bitmaps,
        bitmaps + static_cast<size_t>(grid) * words, words, block_partition, edge_to_col, total,
        large_list + num_windows + 1);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) goto done;
  }
  if ((e = cudaMemcpyAsync(&host_total, total, sizeof(host_total), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
    goto done;
  if ((e = cudaMemcpyAsync(&host_counts[1], large_list + num_windows + 1, sizeof(int32_t), cudaMemcpyDeviceToHost,
                           stream)) != cudaSuccess)
    goto done;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) goto done;
  if (host_counts[1] != 0) {
    set_last_error("tcgnn_sgt_cuda: %d column ids are outside [0, num_cols)", host_counts[1]);
    status = TCGNN_ERR_INVALID_ARG;
    e = cudaSuccess;
    goto cleanup;
  }
  if (tc_blocks_out) *tc_blocks_out = static_cast<int64_t>(host_total) + (num_nodes % blk_h == 0 ? 1 : 0);
  status = TCGNN_OK;
done:
  if (status != TCGNN_OK) {
    set_last_error("tcgnn_sgt_cuda: %s", cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation) status = TCGNN_ERR_OOM;
  }
cleanup:
  if (large_list) cudaFree(large_list);
  if (total) cudaFree(total);
  if (bitmaps) cudaFree(bitmaps);
  return status;
}

}  // namespace tcgnn
