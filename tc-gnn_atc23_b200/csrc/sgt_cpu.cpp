// Multi-threaded host SGT (sparse-graph translation), bit-exact with the reference `preprocess`
// (/root/reference TCGNN_conv/TCGNN.cpp:172-226, dedup helper :157-170), whose OpenMP pragmas are
// ignored by its own build so it runs single-threaded with a std::map per window.
// Here: windows are claimed in chunks by a pool of std::threads; each window is sorted and
// deduplicated in a thread-local buffer and every edge finds its rank by binary search.
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/tcgnn_b200.h"

namespace tcgnn {

void set_last_error(const char* fmt, ...);

int sgt_cpu(const int32_t* row_ptr, const int32_t* col_idx, int32_t num_nodes, int64_t num_edges, int32_t blk_h,
            int32_t blk_w, int32_t* block_partition, int32_t* edge_to_col, int32_t* edge_to_row,
            int64_t* tc_blocks_out, int32_t num_threads) {
  if (row_ptr == nullptr || block_partition == nullptr || num_nodes < 0 || num_edges < 0 || blk_h <= 0 ||
      blk_w <= 0 || (num_edges > 0 && (col_idx == nullptr || edge_to_col == nullptr || edge_to_row == nullptr))) {
    set_last_error("tcgnn_sgt_cpu: bad argument");
    return TCGNN_ERR_INVALID_ARG;
  }
  const int64_t num_windows = (static_cast<int64_t>(num_nodes) + blk_h - 1) / blk_h;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  int nthreads = num_threads > 0 ? num_threads : static_cast<int>(hw);
  if (nthreads > num_windows) nthreads = static_cast<int>(num_windows > 0 ? num_windows : 1);
  if (num_edges < (1 << 14)) nthreads = 1;

  std::atomic<int64_t> next_window{0};
  std::vector<int64_t> partial(static_cast<size_t>(nthreads), 0);
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(256, num_windows / (nthreads * 8 + 1)));

  auto worker = [&](int tid) {
    std::vector<uint32_t> buf;
    int64_t blocks = 0;
    for (;;) {
      const int64_t w0 = next_window.fetch_add(chunk, std::memory_order_relaxed);
      if (w0 >= num_windows) break;
      const int64_t w1 = std::min(num_windows, w0 + chunk);
      for (int64_t w = w0; w < w1; ++w) {
        const int64_t r0 = w * blk_h;
        const int64_t r1 = std::min<int64_t>(r0 + blk_h, num_nodes);
        for (int64_t r = r0; r < r1; ++r)                                  // TCGNN.cpp:194-197
          for (int32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) edge_to_row[e] = static_cast<int32_t>(r);
        const int32_t s = row_ptr[r0], t = row_ptr[r1];
        const int32_t len = t - s;
        buf.resize(static_cast<size_t>(len));
        bool sorted = true;                                                // ids compare as unsigned, :205-209
        for (int32_t i = 0; i < len; ++i) {
          buf[i] = static_cast<uint32_t>(col_idx[s + i]);
          sorted = sorted && (i == 0 || buf[i - 1] <= buf[i]);
        }
        if (!sorted) std::sort(buf.begin(), buf.end());
        const int32_t nu = static_cast<int32_t>(std::unique(buf.begin(), buf.end()) - buf.begin());   // :157-170
        const int32_t cnt = nu > 0 ? nu : 1;            // empty window: the reference's map has one entry
        block_partition[w] = (cnt + blk_w - 1) / blk_w;                    // :216
        blocks += block_partition[w];
        const uint32_t* ub = buf.data();
        for (int32_t e = s; e < t; ++e)                                    // :220-223
          edge_to_col[e] = static_cast<int32_t>(std::lower_bound(ub, ub + nu, static_cast<uint32_t>(col_idx[e])) - ub);
      }
    }
    partial[static_cast<size_t>(tid)] = blocks;
  };

  if (nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(nthreads));
    for (int i = 0; i < nthreads; ++i) pool.emplace_back(worker, i);
    for (auto& th : pool) th.join();
  }
  if (tc_blocks_out != nullptr) {
    int64_t total = 0;
    for (int64_t v : partial) total += v;
    if (num_nodes % blk_h == 0) total += 1;   // the reference's extra loop trip (TCGNN.cpp:200), printed total only
    *tc_blocks_out = total;
  }
  return TCGNN_OK;
}

}  // namespace tcgnn
