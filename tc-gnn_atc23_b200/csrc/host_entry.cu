// Host-buffer entry points: tcgnn_spmm_f32_host / tcgnn_sddmm_f32_host / tcgnn_agnn_f32_host.
//
// The caller's features live in HOST memory (page-locked for full PCIe speed); the graph, its SGT arrays and the plan
// stay resident on the device like the reference's main_tcgnn.py:56-60.  One call = host -> device copy of X, the
// kernels, device -> host copy of the result, all ordered on the caller's stream.
//
// SpMM is pipelined against both copies.  X travels in C row chunks; the plan is split once into C column-chunk
// sub-plans (graph_ops.cu), and Y = sum_j A[:, chunk j] . X[chunk j]:
//   * chunk 0 is aggregated for all windows as soon as it has landed -- while chunks 1.. are still on the bus;
//   * the remaining chunks are aggregated window range by window range (R row chunks, TCGNN_ACCUMULATE), and the
//     finished output rows of range r go back to the host while range r + 1 is being computed (output rows of a
//     window range are one contiguous block, so the DMA runs at full speed).
// Each chunk is rounded to TF32 in place once after it landed (no second N x D pass in the kernels).
// Measured on the reddit-sized R-MAT graph (D = 128): see DESIGN.md section 5.
#include <stdlib.h>

#include <algorithm>

#include "plan.h"

namespace tcgnn {

namespace {

#define HOST_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_last_error("host entry: %s failed: %s", #expr, cudaGetErrorString(_e));         \
      return _e == cudaErrorMemoryAllocation ? TCGNN_ERR_OOM : TCGNN_ERR_CUDA;            \
    }                                                                                     \
  } while (0)

int grow(float** buf, size_t* cap, size_t need, cudaStream_t stream) {
  if (*cap >= need) return TCGNN_OK;
  HOST_CUDA(cudaStreamSynchronize(stream));   // the old staging buffer may still be in use
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  HOST_CUDA(cudaMalloc(buf, need * sizeof(float)));
  *cap = need;
  return TCGNN_OK;
}

int ensure_streams(tcgnn_plan* p, size_t n_events) {
  if (p->h2d_stream == nullptr) {
    HOST_CUDA(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
    HOST_CUDA(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
  }
  while (p->host_ev.size() < n_events) {
    cudaEvent_t ev = nullptr;
    HOST_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    p->host_ev.push_back(ev);
  }
  return TCGNN_OK;
}

// "CxR" from TCGNN_HOST_CHUNKS (column chunks of X x row chunks of Y); default 4x4, 1x1 disables the pipeline
void chunk_setting(int* c, int* r) {
  int a = 4, b = 4;
  if (const char* e = getenv("TCGNN_HOST_CHUNKS")) {
    if (sscanf(e, "%dx%d", &a, &b) != 2) { a = 4; b = 4; }
  }
  *c = std::min(std::max(a, 1), 16);
  *r = std::min(std::max(b, 1), 16);
}

// Build the column-chunk sub-plans and the row chunks once per plan (synchronises the stream a few times).
int ensure_pipeline(tcgnn_plan* p, cudaStream_t stream) {
  if (p->host_pipeline_tried) return TCGNN_OK;
  p->host_pipeline_tried = true;
  int C = 4, R = 4;
  chunk_setting(&C, &R);
  // small graphs: the copies take microseconds, chunking only adds launches
  if (static_cast<int64_t>(p->num_cols) < 65536 || p->num_windows < 64 * R || C < 2) return TCGNN_OK;
  std::vector<int32_t> cb(C + 1), wb(R + 1);
  for (int j = 0; j <= C; ++j) cb[j] = static_cast<int32_t>(static_cast<int64_t>(p->num_cols) * j / C / 16 * 16);
  cb[C] = p->num_cols;
  for (int r = 0; r <= R; ++r) wb[r] = static_cast<int32_t>(static_cast<int64_t>(p->num_windows) * r / R);
  std::vector<tcgnn_plan*> subs;
  for (int j = 0; j < C; ++j) {
    tcgnn_plan* sub = nullptr;
    int st = plan_create_column_chunk(p, cb[j], cb[j + 1], stream, &sub);
    if (st == TCGNN_OK && j > 0) st = plan_set_row_chunks(sub, wb.data(), R, stream);
    if (st != TCGNN_OK) {
      if (sub) plan_destroy(sub);
      for (tcgnn_plan* s : subs) plan_destroy(s);
      return st;
    }
    subs.push_back(sub);
  }
  p->col_chunks = subs;
  p->col_chunk_bounds = cb;
  p->row_chunk_win = wb;
  return TCGNN_OK;
}

int copy_rows(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int32_t dim,
              cudaMemcpyKind kind, cudaStream_t s) {
  if (rows <= 0) return TCGNN_OK;
  if (ld_dst == dim && ld_src == dim) {
    HOST_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * static_cast<size_t>(rows) * dim, kind, s));
  } else {
    HOST_CUDA(cudaMemcpy2DAsync(dst, sizeof(float) * ld_dst, src, sizeof(float) * ld_src, sizeof(float) * dim,
                                static_cast<size_t>(rows), kind, s));
  }
  return TCGNN_OK;
}

}  // namespace

// op: kHostSpmm (dev_arg = edge weights in CSR order or null), kHostSddmm, kHostAgnn (dev_arg = attention_w scalar)
int host_op_launch(tcgnn_plan* p, int op, const float* x_host, int64_t ldx, const float* dev_arg, float* y_host,
                   int64_t ldy, float* e_host, int32_t dim, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(p->host_mu);
  const size_t need_x = static_cast<size_t>(p->num_cols) * dim;
  const size_t need_y = op == kHostSddmm ? 0 : static_cast<size_t>(p->num_nodes) * dim;
  const size_t need_e = (op == kHostSddmm || (op == kHostAgnn && e_host != nullptr)) ? static_cast<size_t>(p->num_edges) : 0;
  int st = grow(&p->host_x_dev, &p->host_x_cap, need_x, stream);
  if (st == TCGNN_OK && need_y) st = grow(&p->host_y_dev, &p->host_y_cap, need_y, stream);
  if (st == TCGNN_OK && need_e) st = grow(&p->host_e_dev, &p->host_e_cap, need_e, stream);
  if (st != TCGNN_OK) return st;
  const bool pipelined_op = op == kHostSpmm && dev_arg == nullptr && (dim & 3) == 0;
  if (pipelined_op) {
    st = ensure_pipeline(p, stream);
    if (st != TCGNN_OK) return st;
  }
  const int C = pipelined_op ? static_cast<int>(p->col_chunks.size()) : 0;
  const int R = C > 0 ? static_cast<int>(p->row_chunk_win.size()) - 1 : 0;
  st = ensure_streams(p, static_cast<size_t>(std::max(C, 1) + std::max(R, 1) + 2));
  if (st != TCGNN_OK) return st;
  cudaEvent_t* ev = p->host_ev.data();
  cudaEvent_t ev_start = ev[0], ev_done = ev[1];
  cudaEvent_t* ev_h2d = ev + 2;
  cudaEvent_t* ev_k = ev + 2 + std::max(C, 1);
  HOST_CUDA(cudaEventRecord(ev_start, stream));
  HOST_CUDA(cudaStreamWaitEvent(p->h2d_stream, ev_start, 0));
  HOST_CUDA(cudaStreamWaitEvent(p->d2h_stream, ev_start, 0));
  float* xd = p->host_x_dev;
  float* yd = p->host_y_dev;

  if (C == 0) {
    // one piece: copy in, run, copy out
    st = copy_rows(xd, dim, x_host, ldx, p->num_cols, dim, cudaMemcpyHostToDevice, p->h2d_stream);
    if (st != TCGNN_OK) return st;
    HOST_CUDA(cudaEventRecord(ev_h2d[0], p->h2d_stream));
    HOST_CUDA(cudaStreamWaitEvent(stream, ev_h2d[0], 0));
    if (op == kHostSpmm) st = spmm_launch(p, xd, dim, dev_arg, yd, dim, dim, 0u, stream);
    else if (op == kHostSddmm) st = sddmm_launch(p, xd, dim, p->host_e_dev, nullptr, nullptr, dim, 0u, stream);
    else st = agnn_launch(p, xd, dim, dev_arg, yd, dim, nullptr, need_e ? p->host_e_dev : nullptr, dim, 0u, stream);
    if (st != TCGNN_OK) return st;
    HOST_CUDA(cudaEventRecord(ev_k[0], stream));
    HOST_CUDA(cudaStreamWaitEvent(p->d2h_stream, ev_k[0], 0));
    if (need_y) {
      st = copy_rows(y_host, ldy, yd, dim, p->num_nodes, dim, cudaMemcpyDeviceToHost, p->d2h_stream);
      if (st != TCGNN_OK) return st;
    }
    if (need_e) HOST_CUDA(cudaMemcpyAsync(e_host, p->host_e_dev, sizeof(float) * need_e, cudaMemcpyDeviceToHost, p->d2h_stream));
  } else {
    const std::vector<int32_t>& cb = p->col_chunk_bounds;
    const std::vector<int32_t>& wb = p->row_chunk_win;
    for (int j = 0; j < C; ++j) {
      st = copy_rows(xd + static_cast<size_t>(cb[j]) * dim, dim, x_host + static_cast<size_t>(cb[j]) * ldx, ldx,
                     cb[j + 1] - cb[j], dim, cudaMemcpyHostToDevice, p->h2d_stream);
      if (st != TCGNN_OK) return st;
      HOST_CUDA(cudaEventRecord(ev_h2d[j], p->h2d_stream));
    }
    // each chunk is rounded in place once it has landed; the kernels then take it as is (TCGNN_X_IS_TF32)
    auto chunk_ready = [&](int j) -> int {
      HOST_CUDA(cudaStreamWaitEvent(stream, ev_h2d[j], 0));
      float* xc = xd + static_cast<size_t>(cb[j]) * dim;
      return round_tf32_launch(xc, dim, xc, dim, cb[j + 1] - cb[j], dim, 0, stream);
    };
    st = chunk_ready(0);
    if (st != TCGNN_OK) return st;
    st = spmm_launch(p->col_chunks[0], xd, dim, nullptr, yd, dim, dim, TCGNN_X_IS_TF32, stream);
    if (st != TCGNN_OK) return st;
    for (int r = 0; r < R; ++r) {
      for (int j = 1; j < C; ++j) {
        if (r == 0) {
          st = chunk_ready(j);
          if (st != TCGNN_OK) return st;
        }
        st = spmm_launch(p->col_chunks[j], xd + static_cast<size_t>(cb[j]) * dim, dim, nullptr, yd, dim, dim,
                         TCGNN_X_IS_TF32 | TCGNN_ACCUMULATE, stream, r);
        if (st != TCGNN_OK) return st;
      }
      HOST_CUDA(cudaEventRecord(ev_k[r], stream));
      HOST_CUDA(cudaStreamWaitEvent(p->d2h_stream, ev_k[r], 0));
      const int64_t row0 = static_cast<int64_t>(wb[r]) * TCGNN_BLK_H;
      const int64_t row1 = std::min<int64_t>(static_cast<int64_t>(wb[r + 1]) * TCGNN_BLK_H, p->num_nodes);
      st = copy_rows(y_host + row0 * ldy, ldy, yd + row0 * dim, dim, row1 - row0, dim, cudaMemcpyDeviceToHost,
                     p->d2h_stream);
      if (st != TCGNN_OK) return st;
    }
  }
  HOST_CUDA(cudaEventRecord(ev_done, p->d2h_stream));
  HOST_CUDA(cudaStreamWaitEvent(stream, ev_done, 0));
  return TCGNN_OK;
}

}  // namespace tcgnn
