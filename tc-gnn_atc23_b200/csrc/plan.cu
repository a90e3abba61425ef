// Plan builder: turns the reference's SGT arrays (blockPartition / edgeToColumn / edgeToRow,
// /root/reference TCGNN.cpp:172-226) into a tile stream the tcgen05 kernels can consume without
// the per-tile rescan of all window edges the reference kernels do (TCGNN_kernel.cu:399-408).
//
// All work is on the device; the only host round trip is the tile total (needed to size the
// allocation).  Integer-only, HBM-bound: plain coalesced kernels, grid-stride.
#include <stdio.h>
#include <stdlib.h>

#include <new>

#include "scan.cuh"

namespace tcgnn {

struct LoadClampedBp {  // max(blockPartition[w], 1): a window always owns at least one tile
  const int32_t* bp;
  __device__ int32_t operator()(int64_t i) const { return max(bp[i], 1); }
};
struct LoadWindowGroups {  // SDDMM work units per window: ceil(tiles / 16)
  const int32_t* win_tile_ptr;
  __device__ int32_t operator()(int64_t i) const { return (win_tile_ptr[i + 1] - win_tile_ptr[i] + 15) >> 4; }
};
struct LoadTilePopc {
  const TileMeta* tiles;
  __device__ int32_t operator()(int64_t i) const {
    const uint4 m = *reinterpret_cast<const uint4*>(tiles[i].mask);
    return __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
  }
};

// ------------------------------------------------------------------------------------------
// tile records
// ------------------------------------------------------------------------------------------
__global__ void init_tiles_kernel(TileMeta* tiles, const int32_t* __restrict__ win_tile_ptr, int32_t num_windows,
                                  int32_t num_tiles) {
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < num_tiles;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // window = last w with win_tile_ptr[w] <= g
    int32_t lo = 0, hi = num_windows;
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (win_tile_ptr[mid] <= g) lo = mid; else hi = mid;
    }
    uint32_t flags = 0;
    if (win_tile_ptr[lo] == g) flags |= kTileFirst;
    if (win_tile_ptr[lo + 1] == g + 1) flags |= kTileLast;
    uint4* rec = reinterpret_cast<uint4*>(tiles + g);
    rec[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);  // cols[0..3] = -1
    rec[1] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);  // cols[4..7] = -1
    rec[2] = make_uint4(0, 0, 0, 0);                                          // mask
    rec[3] = make_uint4(static_cast<uint32_t>(lo), 0u, flags, 0u);            // win, edge_ofs, flags, reserved
  }
}

// One thread per CSR edge: record the gathered row and set the occupancy bit of its tile.
// Edges whose SGT entries are inconsistent (column rank outside the window's tiles, row outside
// the graph) are counted in *bad and skipped instead of corrupting memory.
__global__ void scatter_edges_kernel(TileMeta* tiles, const int32_t* __restrict__ win_tile_ptr,
                                     const int32_t* __restrict__ col_idx, const int32_t* __restrict__ edge_to_col,
                                     const int32_t* __restrict__ edge_to_row, int64_t num_edges, int32_t num_nodes,
                                     int32_t num_cols, int32_t* bad) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < num_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t r = edge_to_row[e];
    const int32_t c = edge_to_col[e];
    const int32_t x = col_idx[e];
    if (r < 0 || r >= num_nodes || c < 0 || x < 0 || x >= num_cols) { atomicAdd(bad, 1); continue; }
    const int32_t w = r / TCGNN_BLK_H;
    const int32_t t0 = win_tile_ptr[w];
    const int32_t g = t0 + c / TCGNN_BLK_W;
    if (g >= win_tile_ptr[w + 1]) { atomicAdd(bad, 1); continue; }
    const int32_t rl = r % TCGNN_BLK_H, cl = c % TCGNN_BLK_W;
    tiles[g].cols[cl] = x;  // every edge of this (window, rank) carries the same column id
    atomicOr(&tiles[g].mask[rl >> 2], 1u << ((rl & 3) * 8 + cl));
  }
}

// Tile range of every persistent CTA.  A CTA's time is ~ a * tiles + b * windows (every window costs an
// accumulator hand-over and a 16 x D output tile; measured b / a = 4..14, profiles/r01e_*), so slices are
// balanced on tiles + win_cost * windows, not on tiles alone: R-MAT graphs have long runs of 1-tile windows.
// blockIdx.y = chunk: the tiles of the windows [win_bounds[c], win_bounds[c+1]) (one chunk = the whole plan when
// win_bounds is null) are cut into `grid` slices; slice_ptr is [chunks][grid + 1].
__global__ void slice_bounds_kernel(const TileMeta* __restrict__ tiles, const int32_t* __restrict__ win_tile_ptr,
                                    const int32_t* __restrict__ win_bounds, int32_t num_windows, int32_t win_cost,
                                    int32_t grid, int32_t* __restrict__ slice_ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > grid) return;
  const int32_t w_lo = win_bounds ? win_bounds[blockIdx.y] : 0;
  const int32_t w_hi = win_bounds ? win_bounds[blockIdx.y + 1] : num_windows;
  const int32_t t_lo = win_tile_ptr[w_lo], t_hi = win_tile_ptr[w_hi];
  const int64_t total = static_cast<int64_t>(t_hi - t_lo) + static_cast<int64_t>(win_cost) * (w_hi - w_lo);
  const int64_t target = total * i / grid;
  // smallest t with cost(t) = (t - t_lo) + win_cost * (windows begun before tile t) >= target
  int32_t lo = t_lo, hi = t_hi;
  while (lo < hi) {
    const int32_t mid = lo + (hi - lo) / 2;
    const int64_t cost = (mid - t_lo) + static_cast<int64_t>(win_cost) * (tiles[mid].win - w_lo);
    if (cost >= target) hi = mid; else lo = mid + 1;
  }
  slice_ptr[static_cast<int64_t>(blockIdx.y) * (grid + 1) + i] = i == grid ? t_hi : (i == 0 ? t_lo : lo);
}

// Cost of a window in TC blocks for the CTA slices.  Measured (profiles/r02k_window_cost_sweep.txt): plans whose
// windows are dense enough for the register epilogue (>= 256 TC blocks per window: 16 store instructions per thread
// and window) want 12 -- reddit-like R-MAT 3.00 -> 2.81 ms -- while the staged epilogue (one bulk copy per window)
// wants 3: products-like 11.2 ms at 3, 11.9 at 12, 13.4 at 32.  TCGNN_WIN_COST overrides.
static int window_cost(bool dense_windows) {
  static const int env = [] {
    const char* e = getenv("TCGNN_WIN_COST");
    const int v = e ? atoi(e) : -1;
    return v > 1024 ? 1024 : v;
  }();
  return env >= 0 ? env : (dense_windows ? 12 : 3);
}

__global__ void store_edge_ofs_kernel(TileMeta* tiles, const int32_t* __restrict__ tile_ofs, int32_t num_tiles) {
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < num_tiles;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x)
    tiles[g].edge_ofs = tile_ofs[g];
}

__global__ void eperm_kernel(const TileMeta* __restrict__ tiles, const int32_t* __restrict__ win_tile_ptr,
                             const int32_t* __restrict__ edge_to_col, const int32_t* __restrict__ edge_to_row,
                             int64_t num_edges, int32_t* eperm) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < num_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t r = edge_to_row[e];
    const int32_t c = edge_to_col[e];
    const int32_t w = r / TCGNN_BLK_H;
    const int32_t g = win_tile_ptr[w] + c / TCGNN_BLK_W;
    const int32_t rl = r % TCGNN_BLK_H, cl = c % TCGNN_BLK_W;
    const TileMeta& t = tiles[g];
    const int word = rl >> 2, bit = (rl & 3) * 8 + cl;
    int rank = __popc(t.mask[word] & ((1u << bit) - 1u));
    for (int i = 0; i < word; ++i) rank += __popc(t.mask[i]);
    eperm[t.edge_ofs + rank] = static_cast<int32_t>(e);  // duplicated (row, col): one edge wins, as in the reference
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return static_cast<int>(g);
}

#define PLAN_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_last_error("%s failed: %s", #expr, cudaGetErrorString(_e));                     \
      status = (_e == cudaErrorMemoryAllocation) ? TCGNN_ERR_OOM : TCGNN_ERR_CUDA;        \
      goto fail;                                                                          \
    }                                                                                     \
  } while (0)

int plan_create(const int32_t* row_ptr, const int32_t* col_idx, const int32_t* block_partition,
                const int32_t* edge_to_col, const int32_t* edge_to_row, int32_t num_nodes, int32_t num_cols,
                int32_t row_base, int64_t num_edges, int32_t num_windows, cudaStream_t stream,
                tcgnn_plan** plan_out) {
  int status = TCGNN_OK;
  tcgnn_plan* p = new (std::nothrow) tcgnn_plan();
  if (p == nullptr) return TCGNN_ERR_OOM;
  int32_t* scratch = nullptr;
  int32_t* tile_ofs = nullptr;
  int32_t host_vals[2] = {0, 0};
  int dev = 0;
  p->row_ptr = row_ptr;
  p->col_idx = col_idx;
  p->edge_to_col = edge_to_col;
  p->edge_to_row = edge_to_row;
  p->num_nodes = num_nodes;
  p->num_cols = num_cols;
  p->row_base = row_base;
  p->num_edges = num_edges;
  p->num_windows = num_windows;
  PLAN_CUDA(cudaGetDevice(&dev));
  p->device = dev;
  PLAN_CUDA(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev));
  {
    int major = 0;
    PLAN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
      set_last_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
      status = TCGNN_ERR_NO_DEVICE;
      goto fail;
    }
  }
  {
    const int64_t nblk_w = (static_cast<int64_t>(num_windows) + kScanTile - 1) / kScanTile;
    PLAN_CUDA(cudaMalloc(&p->win_tile_ptr, sizeof(int32_t) * (static_cast<size_t>(num_windows) + 1)));
    PLAN_CUDA(cudaMalloc(&scratch, sizeof(int32_t) * (static_cast<size_t>(nblk_w) + 2)));
    PLAN_CUDA(exclusive_scan(LoadClampedBp{block_partition}, num_windows, p->win_tile_ptr, scratch, stream));
    PLAN_CUDA(cudaMemcpyAsync(&host_vals[0], p->win_tile_ptr + num_windows, sizeof(int32_t), cudaMemcpyDeviceToHost,
                              stream));
    PLAN_CUDA(cudaStreamSynchronize(stream));
    PLAN_CUDA(cudaFree(scratch));
    scratch = nullptr;
  }
  if (host_vals[0] < 0) {
    set_last_error("tile count overflows int32");
    status = TCGNN_ERR_OVERFLOW;
    goto fail;
  }
  p->num_tiles = host_vals[0];
  {
    const size_t nt = static_cast<size_t>(p->num_tiles);
    const int64_t nblk_t = (static_cast<int64_t>(nt) + kScanTile - 1) / kScanTile;
    PLAN_CUDA(cudaMalloc(&p->tiles, sizeof(TileMeta) * (nt + 1)));
    PLAN_CUDA(cudaMalloc(&scratch, sizeof(int32_t) * (static_cast<size_t>(nblk_t) + 2)));
    PLAN_CUDA(cudaMalloc(&tile_ofs, sizeof(int32_t) * (nt + 1)));
    PLAN_CUDA(cudaMalloc(&p->flag, sizeof(int32_t)));
    PLAN_CUDA(cudaMemsetAsync(p->flag, 0, sizeof(int32_t), stream));
    init_tiles_kernel<<<grid_for(p->num_tiles, 256), 256, 0, stream>>>(p->tiles, p->win_tile_ptr, num_windows,
                                                                       p->num_tiles);
    count_launch();
    if (num_edges > 0) {
      scatter_edges_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(p->tiles, p->win_tile_ptr, col_idx,
                                                                        edge_to_col, edge_to_row, num_edges,
                                                                        num_nodes, num_cols, p->flag);
      count_launch();
    }
    PLAN_CUDA(exclusive_scan(LoadTilePopc{p->tiles}, p->num_tiles, tile_ofs, scratch, stream));
    store_edge_ofs_kernel<<<grid_for(p->num_tiles, 256), 256, 0, stream>>>(p->tiles, tile_ofs, p->num_tiles);
    count_launch();
    PLAN_CUDA(cudaMemcpyAsync(&host_vals[0], tile_ofs + nt, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PLAN_CUDA(cudaMemcpyAsync(&host_vals[1], p->flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PLAN_CUDA(cudaStreamSynchronize(stream));
    PLAN_CUDA(cudaGetLastError());
    p->num_pairs = host_vals[0];
    if (host_vals[1] != 0) {
      set_last_error("%d edges have SGT entries inconsistent with blockPartition / num_nodes "
                     "(edgeToColumn, edgeToRow and blockPartition must come from the same preprocess call)",
                     host_vals[1]);
      status = TCGNN_ERR_INVALID_ARG;
      goto fail;
    }
    PLAN_CUDA(cudaFree(scratch));
    scratch = nullptr;
    PLAN_CUDA(cudaFree(tile_ofs));
    tile_ofs = nullptr;
    // one persistent CTA per SM (fewer for tiny graphs: >= 8 tiles per CTA)
    p->grid = p->num_sms;
    if (p->num_tiles < p->grid * 8) p->grid = p->num_tiles / 8;
    if (p->grid < 1) p->grid = 1;
    PLAN_CUDA(cudaMalloc(&p->slice_ptr, sizeof(int32_t) * (static_cast<size_t>(p->grid) + 1)));
    slice_bounds_kernel<<<(p->grid + 256) / 256, 256, 0, stream>>>(
        p->tiles, p->win_tile_ptr, nullptr, num_windows,
        window_cost(static_cast<int64_t>(p->num_tiles) >= 256LL * num_windows), p->grid, p->slice_ptr);
    count_launch();
    PLAN_CUDA(cudaGetLastError());
    PLAN_CUDA(cudaStreamSynchronize(stream));
  }
  *plan_out = p;
  return TCGNN_OK;
fail:
  if (scratch) cudaFree(scratch);
  if (tile_ofs) cudaFree(tile_ofs);
  plan_destroy(p);
  return status;
}

int plan_ensure_eperm(tcgnn_plan* p, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(p->mu);
  if (p->eperm != nullptr || p->num_edges == 0) return TCGNN_OK;
  int32_t* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, sizeof(int32_t) * static_cast<size_t>(p->num_pairs > 0 ? p->num_pairs : 1));
  if (e != cudaSuccess) {
    set_last_error("cudaMalloc(eperm) failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_OOM;
  }
  eperm_kernel<<<grid_for(p->num_edges, 256), 256, 0, stream>>>(p->tiles, p->win_tile_ptr, p->edge_to_col,
                                                               p->edge_to_row, p->num_edges, buf);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(buf);
    set_last_error("eperm kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  p->eperm = buf;
  return TCGNN_OK;
}

__global__ void fill_groups_kernel(const int32_t* __restrict__ win_tile_ptr, const int32_t* __restrict__ win_group_ptr,
                                   int32_t num_windows, int4* groups) {
  for (int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; w < num_windows;
       w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t t0 = win_tile_ptr[w], t1 = win_tile_ptr[w + 1];
    int32_t gi = win_group_ptr[w];
    for (int32_t t = t0; t < t1; t += 16, ++gi) groups[gi] = make_int4(t, min(16, t1 - t), static_cast<int32_t>(w), 0);
  }
}

int plan_ensure_groups(tcgnn_plan* p, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(p->mu);
  if (p->groups != nullptr) return TCGNN_OK;
  int32_t* win_group_ptr = nullptr;
  int32_t* scratch = nullptr;
  int4* groups = nullptr;
  int32_t total = 0;
  const int64_t nblk = (static_cast<int64_t>(p->num_windows) + kScanTile - 1) / kScanTile;
  cudaError_t e = cudaMalloc(&win_group_ptr, sizeof(int32_t) * (static_cast<size_t>(p->num_windows) + 1));
  if (e == cudaSuccess) e = cudaMalloc(&scratch, sizeof(int32_t) * (static_cast<size_t>(nblk) + 2));
  if (e == cudaSuccess)
    e = exclusive_scan(LoadWindowGroups{p->win_tile_ptr}, p->num_windows, win_group_ptr, scratch, stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(&total, win_group_ptr + p->num_windows, sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaMalloc(&groups, sizeof(int4) * (static_cast<size_t>(total) + 1));
  if (e == cudaSuccess) {
    fill_groups_kernel<<<grid_for(p->num_windows, 256), 256, 0, stream>>>(p->win_tile_ptr, win_group_ptr,
                                                                         p->num_windows, groups);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (win_group_ptr) cudaFree(win_group_ptr);
  if (scratch) cudaFree(scratch);
  if (e != cudaSuccess) {
    if (groups) cudaFree(groups);
    set_last_error("building SDDMM groups failed: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? TCGNN_ERR_OOM : TCGNN_ERR_CUDA;
  }
  p->groups = groups;
  p->num_groups = total;
  return TCGNN_OK;
}

int plan_ensure_scratch(tcgnn_plan* p, float** slot, size_t count) {
  std::lock_guard<std::mutex> lock(p->mu);
  if (*slot != nullptr) return TCGNN_OK;
  cudaError_t e = cudaMalloc(slot, sizeof(float) * (count > 0 ? count : 1));
  if (e != cudaSuccess) {
    *slot = nullptr;
    set_last_error("cudaMalloc(scratch) failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_OOM;
  }
  return TCGNN_OK;
}

int plan_set_row_chunks(tcgnn_plan* p, const int32_t* win_bounds, int n_chunks, cudaStream_t stream) {
  if (n_chunks < 1 || win_bounds == nullptr || win_bounds[0] != 0 || win_bounds[n_chunks] != p->num_windows) {
    set_last_error("plan_set_row_chunks: bounds must run from 0 to num_windows");
    return TCGNN_ERR_INVALID_ARG;
  }
  for (int r = 0; r < n_chunks; ++r)
    if (win_bounds[r + 1] < win_bounds[r]) {
      set_last_error("plan_set_row_chunks: bounds must be non-decreasing");
      return TCGNN_ERR_INVALID_ARG;
    }
  std::lock_guard<std::mutex> lock(p->mu);
  int32_t* d_bounds = nullptr;
  int32_t* table = nullptr;
  cudaError_t e = cudaMalloc(&d_bounds, sizeof(int32_t) * (n_chunks + 1));
  if (e == cudaSuccess) e = cudaMalloc(&table, sizeof(int32_t) * static_cast<size_t>(n_chunks) * (p->grid + 1));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_bounds, win_bounds, sizeof(int32_t) * (n_chunks + 1), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    slice_bounds_kernel<<<dim3((p->grid + 256) / 256, n_chunks), 256, 0, stream>>>(
        p->tiles, p->win_tile_ptr, d_bounds, p->num_windows,
        window_cost(static_cast<int64_t>(p->num_tiles) >= 256LL * p->num_windows), p->grid, table);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // win_bounds is the caller's (pageable) memory
  if (d_bounds) cudaFree(d_bounds);
  if (e != cudaSuccess) {
    if (table) cudaFree(table);
    set_last_error("plan_set_row_chunks failed: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? TCGNN_ERR_OOM : TCGNN_ERR_CUDA;
  }
  if (p->chunk_slice_ptr) cudaFree(p->chunk_slice_ptr);
  p->chunk_slice_ptr = table;
  p->row_chunk_win.assign(win_bounds, win_bounds + n_chunks + 1);
  return TCGNN_OK;
}

int plan_destroy(tcgnn_plan* p) {
  if (p == nullptr) return TCGNN_OK;
  for (tcgnn_plan* c : p->col_chunks) plan_destroy(c);
  for (void* a : p->owned)
    if (a) cudaFree(a);
  if (p->chunk_slice_ptr) cudaFree(p->chunk_slice_ptr);
  if (p->host_e_dev) cudaFree(p->host_e_dev);
  if (p->tiles) cudaFree(p->tiles);
  if (p->win_tile_ptr) cudaFree(p->win_tile_ptr);
  if (p->slice_ptr) cudaFree(p->slice_ptr);
  if (p->eperm) cudaFree(p->eperm);
  if (p->weight_perm) cudaFree(p->weight_perm);
  if (p->sddmm_perm) cudaFree(p->sddmm_perm);
  if (p->x_round) cudaFree(p->x_round);
  if (p->groups) cudaFree(p->groups);
  if (p->flag) cudaFree(p->flag);
  if (p->host_x_dev) cudaFree(p->host_x_dev);
  if (p->host_y_dev) cudaFree(p->host_y_dev);
  if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
  for (cudaEvent_t ev : p->host_ev)
    if (ev) cudaEventDestroy(ev);
  delete p;
  return TCGNN_OK;
}

}  // namespace tcgnn
