// Exclusive scan of int32 values produced by a load functor (three passes: block sums -> scan of sums -> add back).
// Shared by the plan builder (plan.cu) and the graph utilities (graph_ops.cu).  Integer-only, HBM-bound.
#pragma once
#include "plan.h"

namespace tcgnn {

// ------------------------------------------------------------------------------------------
// exclusive scan of int32 (three-pass: block sums -> scan of sums -> add back)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename Load>
__device__ __forceinline__ void block_scan_tile(Load load, int64_t n, int64_t base, int32_t* out, int32_t carry_in,
                                                int32_t* tile_total) {
  __shared__ int32_t warp_sums[kScanThreads / 32];
  const int tid = threadIdx.x;
  int32_t vals[kScanItems];
  int32_t thread_sum = 0;
  const int64_t first = base + static_cast<int64_t>(tid) * kScanItems;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t idx = first + i;
    vals[i] = idx < n ? load(idx) : 0;
    thread_sum += vals[i];
  }
  // warp inclusive scan of thread sums
  int32_t incl = thread_sum;
#pragma unroll
  for (int ofs = 1; ofs < 32; ofs <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, ofs);
    if ((tid & 31) >= ofs) incl += t;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  int32_t warp_prefix = 0;
  int32_t total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const int32_t s = warp_sums[w];
    if (w < (tid >> 5)) warp_prefix += s;
    total += s;
  }
  if (out != nullptr) {
    int32_t run = carry_in + warp_prefix + incl - thread_sum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = first + i;
      if (idx < n) out[idx] = run;
      run += vals[i];
    }
  }
  if (tile_total != nullptr && tid == 0) *tile_total = total;
  __syncthreads();
}

template <typename Load>
__global__ void __launch_bounds__(kScanThreads) scan_block_sums(Load load, int64_t n, int32_t* block_sums) {
  block_scan_tile(load, n, static_cast<int64_t>(blockIdx.x) * kScanTile, nullptr, 0, &block_sums[blockIdx.x]);
}
// single block: exclusive scan of block sums in place; writes the grand total to sums[nblocks]
static __global__ void __launch_bounds__(kScanThreads) scan_sums_inplace(int32_t* sums, int32_t nblocks) {
  __shared__ int32_t carry;
  __shared__ int32_t tile_total;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  struct L {
    const int32_t* p;
    __device__ int32_t operator()(int64_t i) const { return p[i]; }
  } load{sums};
  for (int64_t base = 0; base < nblocks; base += kScanTile) {
    const int32_t c = carry;
    block_scan_tile(load, nblocks, base, sums, c, &tile_total);
    if (threadIdx.x == 0) carry = c + tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nblocks] = carry;
}
template <typename Load>
__global__ void __launch_bounds__(kScanThreads) scan_apply(Load load, int64_t n, const int32_t* block_offsets,
                                                          int32_t* out) {
  block_scan_tile(load, n, static_cast<int64_t>(blockIdx.x) * kScanTile, out, block_offsets[blockIdx.x], nullptr);
}

// out[0..n) = exclusive scan, out[n] = total (also left in scratch[nblocks]).
template <typename Load>
inline cudaError_t exclusive_scan(Load load, int64_t n, int32_t* out, int32_t* scratch, cudaStream_t stream) {
  const int nblocks = static_cast<int>((n + kScanTile - 1) / kScanTile);
  if (nblocks > 0) {
    scan_block_sums<<<nblocks, kScanThreads, 0, stream>>>(load, n, scratch);
    count_launch();
  }
  scan_sums_inplace<<<1, kScanThreads, 0, stream>>>(scratch, nblocks);
  count_launch();
  if (nblocks > 0) {
    scan_apply<<<nblocks, kScanThreads, 0, stream>>>(load, n, scratch, out);
    count_launch();
  }
  return cudaMemcpyAsync(out + n, scratch + nblocks, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
}

struct LoadI32 {
  const int32_t* p;
  __device__ int32_t operator()(int64_t i) const { return p[i]; }
};

}  // namespace tcgnn
