// Python extension module `TCGNN` -- the operator surface of the reference
// (/root/reference TCGNN_conv/TCGNN.cpp:260-272: preprocess, preprocess_gpu, forward, forward_ef,
// forward_AGNN, backward, backward_ef) re-implemented as a thin torch-aware caller of the C ABI in
// include/tcgnn_b200.h.  Same positional signatures, same list-of-tensor returns, so gnn_conv.py /
// main_tcgnn.py of the reference run unchanged.  Differences, all deliberate:
//   * errors raise (TORCH_CHECK) instead of printf + exit(-1) (TCGNN_kernel.cu:211-217);
//   * kernels run on PyTorch's current stream, not the legacy default stream (:197);
//   * dtype / shape / device are validated (the reference only checks is_cuda + is_contiguous);
//   * any feature width is computed (the reference drops D % 16 tails and columns >= 128);
//   * `preprocess_gpu` is a real implementation (the reference's is a stub that prints 0).
// The kernel-side plan is derived once per graph and cached, keyed on the storages of the five
// SGT tensors (weak references + version counters, so a freed or mutated graph never hits).
#include <c10/cuda/CUDAException.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <list>
#include <mutex>
#include <vector>

#include "../../include/tcgnn_b200.h"

namespace {

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
  CHECK_CUDA(x);       \
  CHECK_CONTIGUOUS(x)
#define CHECK_I32(x) TORCH_CHECK((x).scalar_type() == torch::kInt32, #x " must be int32")
#define CHECK_F32(x) TORCH_CHECK((x).scalar_type() == torch::kFloat32, #x " must be float32")

void check_status(int status, const char* what) {
  TORCH_CHECK(status == TCGNN_OK, what, " failed: ", tcgnn_status_string(status), " -- ", tcgnn_last_error());
}

// ---------------------------------------------------------------------------------------------
// plan cache
// ---------------------------------------------------------------------------------------------
// Keyed on the storages of the five graph tensors (weak references + version counters).  Plans are handed out with
// shared ownership: an op holds its plan until its launches are queued, so eviction or a dead-graph purge on another
// thread (autograd runs backward on per-device worker threads) can never free a plan that is in use --
// tcgnn_plan_destroy runs when the last holder lets go, and its cudaFree calls wait for queued kernels.
struct TensorKey {
  const void* ptr = nullptr;
  int64_t used = 0;   // entries the plan reads (callers may pass longer arrays, like the reference's main_tcgnn.py)
  uint32_t version = 0;
  c10::weak_intrusive_ptr<c10::StorageImpl> storage;

  TensorKey(const torch::Tensor& t, int64_t used_entries)
      : ptr(t.data_ptr()),
        used(used_entries),
        version(t.is_inference() ? 0u : static_cast<uint32_t>(t._version())),
        storage(c10::weak_intrusive_ptr<c10::StorageImpl>(t.storage().getWeakStorageImpl())) {}

  bool matches(const torch::Tensor& t, int64_t used_entries) const {
    if (storage.expired()) return false;
    if (t.data_ptr() != ptr || used_entries != used || t.numel() < used) return false;
    if (t.storage().unsafeGetStorageImpl() != storage._unsafe_get_target()) return false;
    const uint32_t v = t.is_inference() ? 0u : static_cast<uint32_t>(t._version());
    return v == version;
  }
  bool alive() const { return !storage.expired(); }
};

struct PlanHolder {
  tcgnn_plan* plan = nullptr;
  std::vector<torch::Tensor> keep;   // arrays the plan borrows that this module allocated (transposed graph)
  torch::Tensor edge_map;            // transposed plans: edge of A^T -> CSR edge id of A
  ~PlanHolder() {
    if (plan != nullptr) tcgnn_plan_destroy(plan);
  }
};
using PlanRef = std::shared_ptr<PlanHolder>;

struct PlanEntry {
  std::vector<TensorKey> keys;   // nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow
  int device = 0;
  int64_t num_cols = 0;
  int64_t row_base = 0;
  bool transposed = false;
  PlanRef holder;
};

std::mutex g_cache_mu;
std::list<PlanEntry> g_cache;   // most recently used first
size_t cache_capacity() {
  static const size_t cap = [] {
    const char* e = getenv("TCGNN_PLAN_CACHE");
    const long v = e ? atol(e) : 32;   // >= one panel plan per GPU of a single-process 8-GPU run, with room to spare
    return static_cast<size_t>(v < 1 ? 1 : v);
  }();
  return cap;
}

void check_status(int status, const char* what);

PlanRef build_transposed(const torch::Tensor& nodePointer, const torch::Tensor& edgeList, int64_t num_nodes,
                         int64_t num_edges, cudaStream_t stream) {
  auto opts = nodePointer.options();
  auto holder = std::make_shared<PlanHolder>();
  auto rp_t = torch::empty({num_nodes + 1}, opts);
  auto ci_t = torch::empty({std::max<int64_t>(num_edges, 1)}, opts);
  holder->edge_map = torch::empty({std::max<int64_t>(num_edges, 1)}, opts);
  check_status(tcgnn_csr_transpose(nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(),
                                   static_cast<int32_t>(num_nodes), static_cast<int32_t>(num_nodes), num_edges,
                                   rp_t.data_ptr<int32_t>(), ci_t.data_ptr<int32_t>(),
                                   holder->edge_map.data_ptr<int32_t>(), stream),
               "tcgnn_csr_transpose");
  const int64_t nwin = (num_nodes + TCGNN_BLK_H - 1) / TCGNN_BLK_H;
  auto bp_t = torch::zeros({nwin}, opts);
  auto e2c_t = torch::zeros({std::max<int64_t>(num_edges, 1)}, opts);
  auto e2r_t = torch::zeros({std::max<int64_t>(num_edges, 1)}, opts);
  check_status(tcgnn_sgt_cuda(rp_t.data_ptr<int32_t>(), ci_t.data_ptr<int32_t>(), static_cast<int32_t>(num_nodes),
                              num_edges, TCGNN_BLK_H, TCGNN_BLK_W, bp_t.data_ptr<int32_t>(), e2c_t.data_ptr<int32_t>(),
                              e2r_t.data_ptr<int32_t>(), nullptr, stream),
               "tcgnn_sgt_cuda (transposed graph)");
  check_status(tcgnn_plan_create(rp_t.data_ptr<int32_t>(), ci_t.data_ptr<int32_t>(), bp_t.data_ptr<int32_t>(),
                                 e2c_t.data_ptr<int32_t>(), e2r_t.data_ptr<int32_t>(), static_cast<int32_t>(num_nodes),
                                 num_edges, static_cast<int32_t>(nwin), stream, &holder->plan),
               "tcgnn_plan_create (transposed graph)");
  holder->keep = {rp_t, ci_t, bp_t, e2c_t, e2r_t};
  return holder;
}

PlanRef get_plan(const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                 const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                 const torch::Tensor& edgeToRow, int64_t num_cols = -1, int64_t row_base = 0,
                 bool transposed = false) {
  const int64_t num_nodes = nodePointer.size(0) - 1;
  const int64_t num_edges = edgeList.size(0);
  const int64_t num_windows = (num_nodes + TCGNN_BLK_H - 1) / TCGNN_BLK_H;
  if (num_cols < 0) num_cols = num_nodes;
  const torch::Tensor* ts[5] = {&nodePointer, &edgeList, &blockPartition, &edgeToColumn, &edgeToRow};
  const int64_t used[5] = {num_nodes + 1, num_edges, num_windows, num_edges, num_edges};
  const int device = nodePointer.get_device();
  std::vector<PlanRef> dead;   // destroyed after the lock is released (cudaFree synchronises the device)
  std::unique_lock<std::mutex> lock(g_cache_mu);
  for (auto it = g_cache.begin(); it != g_cache.end();) {
    bool alive = true;
    for (const auto& k : it->keys) alive = alive && k.alive();
    if (!alive) {   // the graph tensors were freed: drop the plan
      dead.push_back(std::move(it->holder));
      it = g_cache.erase(it);
      continue;
    }
    bool hit = it->device == device && it->num_cols == num_cols && it->row_base == row_base &&
               it->transposed == transposed;
    for (int i = 0; hit && i < 5; ++i) hit = it->keys[i].matches(*ts[i], used[i]);
    if (hit) {
      g_cache.splice(g_cache.begin(), g_cache, it);
      return g_cache.front().holder;
    }
    ++it;
  }
  TORCH_CHECK(num_nodes >= 1, "nodePointer must have at least two entries");
  TORCH_CHECK(num_nodes <= INT32_MAX && num_edges <= INT32_MAX, "graph too large for int32 CSR");
  // '>=': the reference's driver sizes edgeToColumn / edgeToRow from the raw pair count (main_tcgnn.py:44-46,
  // dataset.py:79), which exceeds len(column_index) once scipy has merged duplicate edges
  TORCH_CHECK(edgeToColumn.size(0) >= num_edges && edgeToRow.size(0) >= num_edges,
              "edgeToColumn / edgeToRow must have at least one entry per edge");
  TORCH_CHECK(blockPartition.size(0) >= num_windows, "blockPartition must have ceil(num_nodes / 16) entries");
  PlanEntry entry;
  for (int i = 0; i < 5; ++i) entry.keys.emplace_back(*ts[i], used[i]);
  entry.device = device;
  entry.num_cols = num_cols;
  entry.row_base = row_base;
  entry.transposed = transposed;
  TORCH_CHECK(num_cols >= 1 && num_cols <= INT32_MAX && row_base >= -1 &&
                  (row_base < 0 || row_base + num_nodes <= num_cols),
              "row panel [", row_base, ", ", row_base + num_nodes, ") does not fit a graph of ", num_cols, " nodes");
  auto stream = c10::cuda::getCurrentCUDAStream(device).stream();
  if (transposed) {
    TORCH_CHECK(row_base == 0 && num_cols == num_nodes, "the transposed plan exists for whole graphs only");
    entry.holder = build_transposed(nodePointer, edgeList, num_nodes, num_edges, stream);
  } else {
    entry.holder = std::make_shared<PlanHolder>();
    const int status = tcgnn_plan_create_panel(
        nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(), blockPartition.data_ptr<int32_t>(),
        edgeToColumn.data_ptr<int32_t>(), edgeToRow.data_ptr<int32_t>(), static_cast<int32_t>(num_nodes),
        static_cast<int32_t>(num_cols), static_cast<int32_t>(row_base), num_edges, static_cast<int32_t>(num_windows),
        stream, &entry.holder->plan);
    check_status(status, "tcgnn_plan_create");
  }
  g_cache.push_front(std::move(entry));
  while (g_cache.size() > cache_capacity()) {
    dead.push_back(std::move(g_cache.back().holder));
    g_cache.pop_back();
  }
  PlanRef out = g_cache.front().holder;
  lock.unlock();
  dead.clear();
  return out;
}

void clear_plan_cache() {
  std::list<PlanEntry> old;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    old.swap(g_cache);
  }
}

constexpr int64_t kWholeGraph = -1;    // row_base conventions of the helpers below
constexpr int64_t kDetachedRows = -2;

void check_graph(const torch::Tensor& input, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                 const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                 const torch::Tensor& edgeToRow, int64_t row_base = -1 /* >= 0: row panel of a larger graph */) {
  CHECK_INPUT(input);
  CHECK_INPUT(nodePointer);
  CHECK_INPUT(edgeList);
  CHECK_INPUT(blockPartition);
  CHECK_INPUT(edgeToColumn);
  CHECK_INPUT(edgeToRow);
  CHECK_F32(input);
  CHECK_I32(nodePointer);
  CHECK_I32(edgeList);
  CHECK_I32(blockPartition);
  CHECK_I32(edgeToColumn);
  CHECK_I32(edgeToRow);
  TORCH_CHECK(input.dim() == 2, "input must be [num_nodes, dim]");
  TORCH_CHECK(nodePointer.dim() == 1 && edgeList.dim() == 1, "nodePointer / edgeList must be 1-D");
  if (row_base == kDetachedRows) {
    // partial product over one source panel: the column ids index `input`, the rows are not rows of `input`
  } else if (row_base < 0) {
    TORCH_CHECK(input.size(0) == nodePointer.size(0) - 1, "input has ", input.size(0), " rows but the graph has ",
                nodePointer.size(0) - 1, " nodes");
  } else {
    TORCH_CHECK(row_base + nodePointer.size(0) - 1 <= input.size(0), "row panel [", row_base, ", ",
                row_base + nodePointer.size(0) - 1, ") exceeds the ", input.size(0), " rows of input");
  }
  TORCH_CHECK(input.size(1) >= 1, "input must have at least one feature column");
  const auto dev = input.device();
  TORCH_CHECK(nodePointer.device() == dev && edgeList.device() == dev && blockPartition.device() == dev &&
                  edgeToColumn.device() == dev && edgeToRow.device() == dev,
              "all tensors must be on the same CUDA device");
}

// ---------------------------------------------------------------------------------------------
// operators (reference: TCGNN.cpp:63-150)
// ---------------------------------------------------------------------------------------------
const float* check_attention(const torch::Tensor& edgeAttention, const torch::Tensor& input, int64_t num_edges) {
  CHECK_INPUT(edgeAttention);
  CHECK_F32(edgeAttention);
  TORCH_CHECK(edgeAttention.device() == input.device(), "edgeAttention must be on the same device as input");
  // [n_heads, E]; like the reference kernel (TCGNN_kernel.cu:529) only head 0 is read.
  TORCH_CHECK(edgeAttention.dim() >= 1 && edgeAttention.size(-1) == num_edges,
              "edgeAttention must be [n_heads, num_edges]");
  return edgeAttention.data_ptr<float>();
}

// out (optional): accumulate into this [num_rows, dim] tensor instead of allocating (Y += A X, TCGNN_ACCUMULATE)
torch::Tensor run_spmm(const torch::Tensor& input, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                       const torch::Tensor* edgeAttention, const torch::Tensor& blockPartition,
                       const torch::Tensor& edgeToColumn, const torch::Tensor& edgeToRow, int64_t row_base,
                       bool x_is_tf32 = false, const torch::Tensor* accumulate_into = nullptr,
                       bool transposed = false) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, row_base);
  const float* weights = edgeAttention != nullptr ? check_attention(*edgeAttention, input, edgeList.size(0)) : nullptr;
  c10::cuda::CUDAGuard guard(input.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow,
                          row_base == kWholeGraph ? -1 : input.size(0),
                          row_base == kWholeGraph ? 0 : (row_base == kDetachedRows ? -1 : row_base), transposed);
  torch::Tensor wt;   // transposed graph: the weights of A's edges carried over to A^T's edge order
  if (transposed && weights != nullptr && edgeList.size(0) > 0) {
    wt = edgeAttention->reshape({-1}).slice(0, 0, edgeList.size(0)).index_select(0, plan->edge_map);
    weights = wt.data_ptr<float>();
  }
  torch::Tensor output;
  uint32_t flags = x_is_tf32 ? TCGNN_X_IS_TF32 : 0u;
  if (accumulate_into != nullptr) {
    output = *accumulate_into;
    CHECK_INPUT(output);
    CHECK_F32(output);
    TORCH_CHECK(output.dim() == 2 && output.size(0) == nodePointer.size(0) - 1 && output.size(1) == input.size(1) &&
                    output.device() == input.device(),
                "accumulate_into must be a contiguous [num_rows, dim] float32 tensor on the input's device");
    flags |= TCGNN_ACCUMULATE;
  } else {
    output = torch::empty({nodePointer.size(0) - 1, input.size(1)}, input.options());
  }
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_spmm_f32_ex(plan->plan, input.data_ptr<float>(), input.size(1), weights, output.data_ptr<float>(),
                                 output.size(1), static_cast<int32_t>(input.size(1)), flags, stream),
               "tcgnn_spmm_f32");
  return output;
}

torch::Tensor run_sddmm(const torch::Tensor& input, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                        const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                        const torch::Tensor& edgeToRow, int64_t row_base, bool x_is_tf32 = false) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, row_base);
  c10::cuda::CUDAGuard guard(input.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow,
                          row_base < 0 ? -1 : input.size(0), row_base < 0 ? 0 : row_base);
  auto output = torch::empty({edgeList.size(0)}, input.options());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_sddmm_f32_ex(plan->plan, input.data_ptr<float>(), input.size(1), output.data_ptr<float>(),
                                  static_cast<int32_t>(input.size(1)), x_is_tf32 ? TCGNN_X_IS_TF32 : 0u, stream),
               "tcgnn_sddmm_f32");
  return output;
}

std::vector<torch::Tensor> spmm_forward(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                        torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                        torch::Tensor edgeToRow) {
  return {run_spmm(input, nodePointer, edgeList, nullptr, blockPartition, edgeToColumn, edgeToRow, -1)};
}

std::vector<torch::Tensor> spmm_forward_AGNN(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                             torch::Tensor edgeAttention, torch::Tensor blockPartition,
                                             torch::Tensor edgeToColumn, torch::Tensor edgeToRow) {
  return {run_spmm(input, nodePointer, edgeList, &edgeAttention, blockPartition, edgeToColumn, edgeToRow, -1)};
}

std::vector<torch::Tensor> sddmm_forward(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                         torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                         torch::Tensor edgeToRow) {
  return {run_sddmm(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, -1)};
}

// Aggregation over the TRANSPOSED graph: dX = A^T dY, what the backward pass of Y = A X needs on a directed graph.
// The reference's `backward` re-uses the forward CSR (TCGNN.cpp:268, gnn_conv.py:76-85), i.e. assumes A == A^T; the
// transposed CSR, its SGT and its plan are derived on the device on first use and cached with the graph.
std::vector<torch::Tensor> spmm_backward_T(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                           torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                           torch::Tensor edgeToRow) {
  return {run_spmm(input, nodePointer, edgeList, nullptr, blockPartition, edgeToColumn, edgeToRow, -1, false, nullptr,
                   true)};
}
std::vector<torch::Tensor> spmm_backward_T_AGNN(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                                torch::Tensor edgeAttention, torch::Tensor blockPartition,
                                                torch::Tensor edgeToColumn, torch::Tensor edgeToRow) {
  return {run_spmm(input, nodePointer, edgeList, &edgeAttention, blockPartition, edgeToColumn, edgeToRow, -1, false,
                   nullptr, true)};
}

// Fused AGNN edge pipeline (tcgnn_agnn_f32; reference gnn_conv.py:125-132 as one call).  attention_w: the layer's
// [1, n_heads = 1] parameter (CUDA tensor, read on the device -- no host synchronisation).  Returns
// [Y, attention in tile order (tf32-rounded; feed it to forward_AGNN_tile for the backward pass), edge_feature in CSR
// order or an empty tensor when want_edge_feature is false].
std::vector<torch::Tensor> agnn_fused(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                      torch::Tensor attention_w, torch::Tensor blockPartition,
                                      torch::Tensor edgeToColumn, torch::Tensor edgeToRow, bool want_edge_feature) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, -1);
  CHECK_INPUT(attention_w);
  CHECK_F32(attention_w);
  TORCH_CHECK(attention_w.numel() >= 1 && attention_w.device() == input.device(),
              "attention_w must be a float32 CUDA tensor with at least one element on the input's device");
  c10::cuda::CUDAGuard guard(input.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  int64_t info[8];
  check_status(tcgnn_plan_info(plan->plan, info), "tcgnn_plan_info");
  auto output = torch::empty({nodePointer.size(0) - 1, input.size(1)}, input.options());
  auto att_tile = torch::empty({std::max<int64_t>(info[5], 1)}, input.options());
  auto edge_feature = torch::empty({want_edge_feature ? edgeList.size(0) : 0}, input.options());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_agnn_f32(plan->plan, input.data_ptr<float>(), input.size(1), attention_w.data_ptr<float>(),
                              output.data_ptr<float>(), output.size(1), att_tile.data_ptr<float>(),
                              want_edge_feature && edgeList.size(0) > 0 ? edge_feature.data_ptr<float>() : nullptr,
                              static_cast<int32_t>(input.size(1)), 0u, stream),
               "tcgnn_agnn_f32");
  return {output, att_tile, edge_feature};
}

// Weighted SpMM whose weights are already in the plan's tile order (the second output of forward_AGNN_fused).
std::vector<torch::Tensor> spmm_forward_AGNN_tile(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                                  torch::Tensor att_tile, torch::Tensor blockPartition,
                                                  torch::Tensor edgeToColumn, torch::Tensor edgeToRow) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, -1);
  CHECK_INPUT(att_tile);
  CHECK_F32(att_tile);
  c10::cuda::CUDAGuard guard(input.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  int64_t info[8];
  check_status(tcgnn_plan_info(plan->plan, info), "tcgnn_plan_info");
  TORCH_CHECK(att_tile.numel() >= info[5] && att_tile.device() == input.device(),
              "att_tile must hold one weight per distinct (row, col) pair of the plan (", info[5], ")");
  auto output = torch::empty({nodePointer.size(0) - 1, input.size(1)}, input.options());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_spmm_f32_ex(plan->plan, input.data_ptr<float>(), input.size(1),
                                 info[5] > 0 ? att_tile.data_ptr<float>() : nullptr, output.data_ptr<float>(),
                                 output.size(1), static_cast<int32_t>(input.size(1)),
                                 info[5] > 0 ? TCGNN_W_TILE_ORDER : 0u, stream),
               "tcgnn_spmm_f32 (tile-ordered weights)");
  return {output};
}

// Row-panel variants for 1-D destination-row sharding (new; the reference is single-GPU): `input` is the
// feature matrix the panel's column ids index (the all-gathered matrix, or one source panel's packed rows), the five
// graph tensors describe the caller's row panel (tcgnn_plan_create_panel), `row_base` is the global id of the panel's
// first row.  accumulate_into: add the product to an existing [num_rows, dim] tensor (per-source-panel partial sums).
std::vector<torch::Tensor> panel_forward(torch::Tensor input, int64_t row_base, torch::Tensor nodePointer,
                                         torch::Tensor edgeList, torch::Tensor blockPartition,
                                         torch::Tensor edgeToColumn, torch::Tensor edgeToRow, bool x_is_tf32,
                                         c10::optional<torch::Tensor> accumulate_into) {
  TORCH_CHECK(row_base >= 0, "row_base must be >= 0");
  return {run_spmm(input, nodePointer, edgeList, nullptr, blockPartition, edgeToColumn, edgeToRow, row_base, x_is_tf32,
                   accumulate_into.has_value() ? &*accumulate_into : nullptr)};
}

// Partial product over ONE source panel (overlapped exchange): `input` holds the rows that source shipped (its
// whole panel, or the packed rows this panel references), the graph's column ids index `input`, and the product is
// added to `accumulate_into` when given.
std::vector<torch::Tensor> source_forward(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                                          torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                          torch::Tensor edgeToRow, bool x_is_tf32,
                                          c10::optional<torch::Tensor> accumulate_into) {
  return {run_spmm(input, nodePointer, edgeList, nullptr, blockPartition, edgeToColumn, edgeToRow, kDetachedRows,
                   x_is_tf32, accumulate_into.has_value() ? &*accumulate_into : nullptr)};
}

std::vector<torch::Tensor> panel_forward_AGNN(torch::Tensor input, int64_t row_base, torch::Tensor nodePointer,
                                              torch::Tensor edgeList, torch::Tensor edgeAttention,
                                              torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                              torch::Tensor edgeToRow, bool x_is_tf32) {
  TORCH_CHECK(row_base >= 0, "row_base must be >= 0");
  return {run_spmm(input, nodePointer, edgeList, &edgeAttention, blockPartition, edgeToColumn, edgeToRow, row_base,
                   x_is_tf32)};
}

std::vector<torch::Tensor> panel_forward_ef(torch::Tensor input, int64_t row_base, torch::Tensor nodePointer,
                                            torch::Tensor edgeList, torch::Tensor blockPartition,
                                            torch::Tensor edgeToColumn, torch::Tensor edgeToRow, bool x_is_tf32) {
  TORCH_CHECK(row_base >= 0, "row_base must be >= 0");
  return {run_sddmm(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, row_base, x_is_tf32)};
}

// cvt.rna.tf32 of a feature matrix, once, for callers that feed several ops or ship X between GPUs
// (pass the result with x_is_tf32 = true).  Rows are padded to a multiple of 4 floats only if needed.
torch::Tensor round_tf32(torch::Tensor input) {
  CHECK_INPUT(input);
  CHECK_F32(input);
  TORCH_CHECK(input.dim() == 2 && input.size(1) >= 1, "input must be [rows, dim]");
  TORCH_CHECK(input.size(1) % 4 == 0, "round_tf32 needs dim % 4 == 0 (16-byte rows); pass the raw matrix to the ops instead");
  c10::cuda::CUDAGuard guard(input.device());
  auto out = torch::empty_like(input);
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_round_tf32(input.data_ptr<float>(), input.size(1), out.data_ptr<float>(), out.size(1),
                                input.size(0), static_cast<int32_t>(input.size(1)), stream),
               "tcgnn_round_tf32");
  return out;
}

// ---------------------------------------------------------------------------------------------
// SGT (reference: TCGNN.cpp:172-256)
// ---------------------------------------------------------------------------------------------
void check_sgt_args(const torch::Tensor& edgeList, const torch::Tensor& nodePointer, int64_t num_nodes,
                    int64_t blockSize_h, int64_t blockSize_w, const torch::Tensor& blockPartition,
                    const torch::Tensor& edgeToColumn, const torch::Tensor& edgeToRow) {
  CHECK_CONTIGUOUS(edgeList);
  CHECK_CONTIGUOUS(nodePointer);
  CHECK_CONTIGUOUS(blockPartition);
  CHECK_CONTIGUOUS(edgeToColumn);
  CHECK_CONTIGUOUS(edgeToRow);
  CHECK_I32(edgeList);
  CHECK_I32(nodePointer);
  CHECK_I32(blockPartition);
  CHECK_I32(edgeToColumn);
  CHECK_I32(edgeToRow);
  TORCH_CHECK(num_nodes >= 0 && num_nodes <= INT32_MAX, "num_nodes out of range");
  TORCH_CHECK(blockSize_h >= 1 && blockSize_w >= 1, "block sizes must be positive");
  TORCH_CHECK(nodePointer.numel() >= num_nodes + 1, "nodePointer must have num_nodes + 1 entries");
  TORCH_CHECK(edgeToColumn.numel() >= edgeList.numel() && edgeToRow.numel() >= edgeList.numel(),
              "edgeToColumn / edgeToRow must have one entry per edge");
  TORCH_CHECK(blockPartition.numel() >= (num_nodes + blockSize_h - 1) / blockSize_h,
              "blockPartition must have ceil(num_nodes / blockSize_h) entries");
  const auto dev = edgeList.device();
  TORCH_CHECK(nodePointer.device() == dev && blockPartition.device() == dev && edgeToColumn.device() == dev &&
                  edgeToRow.device() == dev,
              "all SGT tensors must be on the same device");
}

void preprocess_impl(torch::Tensor edgeList, torch::Tensor nodePointer, int64_t num_nodes, int64_t blockSize_h,
                     int64_t blockSize_w, torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                     torch::Tensor edgeToRow, int64_t num_cols = -1, bool quiet = false) {
  if (num_cols < 0) num_cols = num_nodes;
  TORCH_CHECK(num_cols <= INT32_MAX, "num_cols out of range");
  check_sgt_args(edgeList, nodePointer, num_nodes, blockSize_h, blockSize_w, blockPartition, edgeToColumn, edgeToRow);
  int64_t tc_blocks = 0;
  if (edgeList.is_cuda()) {
    c10::cuda::CUDAGuard guard(edgeList.device());
    auto stream = c10::cuda::getCurrentCUDAStream(edgeList.get_device()).stream();
    check_status(tcgnn_sgt_cuda_panel(nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(),
                                static_cast<int32_t>(num_nodes), static_cast<int32_t>(num_cols), edgeList.numel(),
                                static_cast<int32_t>(blockSize_h),
                                static_cast<int32_t>(blockSize_w), blockPartition.data_ptr<int32_t>(),
                                edgeToColumn.data_ptr<int32_t>(), edgeToRow.data_ptr<int32_t>(), &tc_blocks, stream),
                 "tcgnn_sgt_cuda");
  } else {
    int status;
    {
      pybind11::gil_scoped_release release;
      status = tcgnn_sgt_cpu(nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(),
                             static_cast<int32_t>(num_nodes), edgeList.numel(), static_cast<int32_t>(blockSize_h),
                             static_cast<int32_t>(blockSize_w), blockPartition.data_ptr<int32_t>(),
                             edgeToColumn.data_ptr<int32_t>(), edgeToRow.data_ptr<int32_t>(), &tc_blocks, 0);
    }
    check_status(status, "tcgnn_sgt_cpu");
  }
  if (quiet) return;
  // same two lines the reference prints (TCGNN.cpp:225) so 1_log2csv.py-style log scraping keeps working
  printf("TC_Blocks:\t%lld\nExp_Edges:\t%lld\n", static_cast<long long>(tc_blocks),
         static_cast<long long>(tc_blocks * blockSize_h * blockSize_w));
  fflush(stdout);
}

void preprocess(torch::Tensor edgeList, torch::Tensor nodePointer, int64_t num_nodes, int64_t blockSize_h,
                int64_t blockSize_w, torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                torch::Tensor edgeToRow) {
  preprocess_impl(edgeList, nodePointer, num_nodes, blockSize_h, blockSize_w, blockPartition, edgeToColumn, edgeToRow);
}

void preprocess_gpu(torch::Tensor edgeList, torch::Tensor nodePointer, int64_t num_nodes, int64_t blockSize_h,
                    int64_t blockSize_w, torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                    torch::Tensor edgeToRow) {
  CHECK_CUDA(edgeList);
  CHECK_CUDA(nodePointer);
  CHECK_CUDA(blockPartition);
  CHECK_CUDA(edgeToColumn);
  CHECK_CUDA(edgeToRow);
  preprocess_impl(edgeList, nodePointer, num_nodes, blockSize_h, blockSize_w, blockPartition, edgeToColumn, edgeToRow);
}

// SGT of a row panel (sharding): num_rows rows whose column ids are global, in [0, num_cols).  Silent.
void preprocess_panel(torch::Tensor edgeList, torch::Tensor nodePointer, int64_t num_rows, int64_t num_cols,
                      int64_t blockSize_h, int64_t blockSize_w, torch::Tensor blockPartition,
                      torch::Tensor edgeToColumn, torch::Tensor edgeToRow) {
  preprocess_impl(edgeList, nodePointer, num_rows, blockSize_h, blockSize_w, blockPartition, edgeToColumn, edgeToRow,
                  num_cols, true);
}

std::vector<int64_t> plan_info(torch::Tensor nodePointer, torch::Tensor edgeList, torch::Tensor blockPartition,
                               torch::Tensor edgeToColumn, torch::Tensor edgeToRow) {
  c10::cuda::CUDAGuard guard(nodePointer.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  std::vector<int64_t> info(8, 0);
  check_status(tcgnn_plan_info(plan->plan, info.data()), "tcgnn_plan_info");
  return info;
}

// Fused "round + push": writes cvt.rna.tf32(input) to a raw device address -- local memory, a P2P-mapped
// pointer into a peer GPU's gathered matrix, or (multicast = true) an NVSwitch multicast address.
void round_tf32_into(torch::Tensor input, int64_t out_ptr, int64_t ldo, bool multicast) {
  CHECK_INPUT(input);
  CHECK_F32(input);
  TORCH_CHECK(input.dim() == 2 && input.size(1) >= 1, "input must be [rows, dim]");
  if (input.size(0) == 0) return;
  c10::cuda::CUDAGuard guard(input.device());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  float* out = reinterpret_cast<float*>(static_cast<uintptr_t>(out_ptr));
  const int st = multicast ? tcgnn_round_tf32_multicast(input.data_ptr<float>(), input.size(1), out, ldo, input.size(0),
                                                        static_cast<int32_t>(input.size(1)), stream)
                           : tcgnn_round_tf32(input.data_ptr<float>(), input.size(1), out, ldo, input.size(0),
                                              static_cast<int32_t>(input.size(1)), stream);
  check_status(st, "tcgnn_round_tf32");
}

// Operators with HOST feature / result tensors (tcgnn_*_f32_host): X and the results are CPU tensors (pinned for full
// speed), the graph tensors are CUDA tensors.  With sync = false the caller must synchronise the current stream (or
// an event recorded on it) before reading the results.
void check_host_graph(const torch::Tensor& x_host, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                      const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                      const torch::Tensor& edgeToRow) {
  TORCH_CHECK(!x_host.is_cuda() && x_host.is_contiguous() && x_host.scalar_type() == torch::kFloat32 &&
                  x_host.dim() == 2,
              "x_host must be a contiguous float32 CPU tensor [num_nodes, dim]");
  CHECK_INPUT(nodePointer);
  CHECK_INPUT(edgeList);
  CHECK_INPUT(blockPartition);
  CHECK_INPUT(edgeToColumn);
  CHECK_INPUT(edgeToRow);
  CHECK_I32(nodePointer);
  CHECK_I32(edgeList);
  CHECK_I32(blockPartition);
  CHECK_I32(edgeToColumn);
  CHECK_I32(edgeToRow);
  TORCH_CHECK(x_host.size(0) == nodePointer.size(0) - 1, "x_host has ", x_host.size(0), " rows but the graph has ",
              nodePointer.size(0) - 1, " nodes");
}

torch::Tensor host_result(c10::optional<torch::Tensor> given, std::vector<int64_t> shape, const torch::Tensor& like,
                          const char* name) {
  torch::Tensor t = given.has_value() ? *given : torch::empty(shape, like.options().pinned_memory(true));
  TORCH_CHECK(!t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kFloat32 && t.sizes().vec() == shape, name,
              " must be a contiguous float32 CPU tensor of the result's shape");
  return t;
}

torch::Tensor forward_host(torch::Tensor x_host, torch::Tensor nodePointer, torch::Tensor edgeList,
                           torch::Tensor blockPartition, torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                           c10::optional<torch::Tensor> y_host_opt, bool sync) {
  check_host_graph(x_host, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  torch::Tensor y_host = host_result(y_host_opt, {x_host.size(0), x_host.size(1)}, x_host, "y_host");
  c10::cuda::CUDAGuard guard(nodePointer.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  auto stream = c10::cuda::getCurrentCUDAStream(nodePointer.get_device()).stream();
  check_status(tcgnn_spmm_f32_host(plan->plan, x_host.data_ptr<float>(), x_host.size(1), nullptr,
                                   y_host.data_ptr<float>(), y_host.size(1), static_cast<int32_t>(x_host.size(1)),
                                   stream),
               "tcgnn_spmm_f32_host");
  if (sync) C10_CUDA_CHECK(cudaStreamSynchronize(stream));
  return y_host;
}

torch::Tensor forward_ef_host(torch::Tensor x_host, torch::Tensor nodePointer, torch::Tensor edgeList,
                              torch::Tensor blockPartition, torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                              c10::optional<torch::Tensor> e_host_opt, bool sync) {
  check_host_graph(x_host, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  torch::Tensor e_host = host_result(e_host_opt, {edgeList.size(0)}, x_host, "edge_out_host");
  c10::cuda::CUDAGuard guard(nodePointer.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  auto stream = c10::cuda::getCurrentCUDAStream(nodePointer.get_device()).stream();
  check_status(tcgnn_sddmm_f32_host(plan->plan, x_host.data_ptr<float>(), x_host.size(1), e_host.data_ptr<float>(),
                                    static_cast<int32_t>(x_host.size(1)), stream),
               "tcgnn_sddmm_f32_host");
  if (sync) C10_CUDA_CHECK(cudaStreamSynchronize(stream));
  return e_host;
}

torch::Tensor forward_AGNN_host(torch::Tensor x_host, torch::Tensor nodePointer, torch::Tensor edgeList,
                                torch::Tensor attention_w, torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                                torch::Tensor edgeToRow, c10::optional<torch::Tensor> y_host_opt, bool sync) {
  check_host_graph(x_host, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  CHECK_INPUT(attention_w);
  CHECK_F32(attention_w);
  TORCH_CHECK(attention_w.numel() >= 1, "attention_w must have at least one element");
  torch::Tensor y_host = host_result(y_host_opt, {x_host.size(0), x_host.size(1)}, x_host, "y_host");
  c10::cuda::CUDAGuard guard(nodePointer.device());
  PlanRef plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  auto stream = c10::cuda::getCurrentCUDAStream(nodePointer.get_device()).stream();
  check_status(tcgnn_agnn_f32_host(plan->plan, x_host.data_ptr<float>(), x_host.size(1), attention_w.data_ptr<float>(),
                                   y_host.data_ptr<float>(), y_host.size(1), nullptr,
                                   static_cast<int32_t>(x_host.size(1)), stream),
               "tcgnn_agnn_f32_host");
  if (sync) C10_CUDA_CHECK(cudaStreamSynchronize(stream));
  return y_host;
}

// Exchange helpers of the sharded path (sharding.py): pack referenced rows, stream-ordered flag wait.
void gather_rows(torch::Tensor src, torch::Tensor rows, torch::Tensor dst) {
  CHECK_INPUT(src);
  CHECK_INPUT(rows);
  CHECK_INPUT(dst);
  CHECK_F32(src);
  CHECK_F32(dst);
  CHECK_I32(rows);
  TORCH_CHECK(src.dim() == 2 && dst.dim() == 2 && dst.size(1) == src.size(1) && dst.size(0) >= rows.numel(),
              "gather_rows: dst must be [>= len(rows), dim]");
  c10::cuda::CUDAGuard guard(src.device());
  auto stream = c10::cuda::getCurrentCUDAStream(src.get_device()).stream();
  check_status(tcgnn_gather_rows(src.data_ptr<float>(), src.size(1), rows.data_ptr<int32_t>(), rows.numel(),
                                 dst.data_ptr<float>(), stream),
               "tcgnn_gather_rows");
}

void stream_wait_flag(torch::Tensor flags, int64_t index, int64_t value, int64_t timeout_ms, torch::Tensor error_out) {
  CHECK_INPUT(flags);
  CHECK_I32(flags);
  CHECK_INPUT(error_out);
  CHECK_I32(error_out);
  TORCH_CHECK(index >= 0 && index < flags.numel() && error_out.numel() >= 1, "stream_wait_flag: bad index");
  c10::cuda::CUDAGuard guard(flags.device());
  auto stream = c10::cuda::getCurrentCUDAStream(flags.get_device()).stream();
  check_status(tcgnn_stream_wait_flag(flags.data_ptr<int32_t>() + index, static_cast<int32_t>(value),
                                      static_cast<int32_t>(timeout_ms), error_out.data_ptr<int32_t>(), stream),
               "tcgnn_stream_wait_flag");
}

void stream_wait_flag_dev(torch::Tensor flags, int64_t index, torch::Tensor value_dev, int64_t timeout_ms,
                          torch::Tensor error_out) {
  CHECK_INPUT(flags);
  CHECK_I32(flags);
  CHECK_INPUT(value_dev);
  CHECK_I32(value_dev);
  CHECK_INPUT(error_out);
  CHECK_I32(error_out);
  TORCH_CHECK(index >= 0 && index < flags.numel() && value_dev.numel() >= 1 && error_out.numel() >= 1,
              "stream_wait_flag_dev: bad arguments");
  c10::cuda::CUDAGuard guard(flags.device());
  auto stream = c10::cuda::getCurrentCUDAStream(flags.get_device()).stream();
  check_status(tcgnn_stream_wait_flag_dev(flags.data_ptr<int32_t>() + index, value_dev.data_ptr<int32_t>(),
                                          static_cast<int32_t>(timeout_ms), error_out.data_ptr<int32_t>(), stream),
               "tcgnn_stream_wait_flag_dev");
}

// (row_ptr_t, col_idx_t, edge_map_t) of A^T on the device (tcgnn_csr_transpose)
std::vector<torch::Tensor> csr_transpose(torch::Tensor nodePointer, torch::Tensor edgeList, int64_t num_cols) {
  CHECK_INPUT(nodePointer);
  CHECK_INPUT(edgeList);
  CHECK_I32(nodePointer);
  CHECK_I32(edgeList);
  const int64_t n = nodePointer.size(0) - 1, e = edgeList.size(0);
  if (num_cols < 0) num_cols = n;
  c10::cuda::CUDAGuard guard(nodePointer.device());
  auto rp_t = torch::empty({num_cols + 1}, nodePointer.options());
  auto ci_t = torch::empty({e}, nodePointer.options());
  auto map_t = torch::empty({e}, nodePointer.options());
  auto stream = c10::cuda::getCurrentCUDAStream(nodePointer.get_device()).stream();
  check_status(tcgnn_csr_transpose(nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(), static_cast<int32_t>(n),
                                   static_cast<int32_t>(num_cols), e, rp_t.data_ptr<int32_t>(), ci_t.data_ptr<int32_t>(),
                                   map_t.data_ptr<int32_t>(), stream),
               "tcgnn_csr_transpose");
  return {rp_t, ci_t, map_t};
}

// Second phase of the balanced exchange: rows [begin, end) segments of `local` -> the same rows at every peer address.
void push_rows(torch::Tensor local, std::vector<int64_t> peer_ptrs, std::vector<int64_t> seg_begin,
               std::vector<int64_t> seg_end) {
  CHECK_INPUT(local);
  CHECK_F32(local);
  TORCH_CHECK(local.dim() == 2 && seg_begin.size() == seg_end.size(), "push_rows: bad arguments");
  std::vector<float*> peers;
  for (int64_t p : peer_ptrs) peers.push_back(reinterpret_cast<float*>(static_cast<uintptr_t>(p)));
  c10::cuda::CUDAGuard guard(local.device());
  auto stream = c10::cuda::getCurrentCUDAStream(local.get_device()).stream();
  check_status(tcgnn_push_rows(local.data_ptr<float>(), peers.data(), static_cast<int32_t>(peers.size()),
                               seg_begin.data(), seg_end.data(), static_cast<int32_t>(seg_begin.size()),
                               local.size(1), stream),
               "tcgnn_push_rows");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "TC-GNN aggregation operators (SGT + SpMM + SDDMM), B200 / sm_100a implementation";
  m.def("preprocess", &preprocess, "Preprocess Step (SGT; CPU tensors -> host threads, CUDA tensors -> device)");
  m.def("preprocess_gpu", &preprocess_gpu, "Preprocess Step (SGT, CUDA)");
  // forward computation
  m.def("forward", &spmm_forward, "TC-GNN SPMM forward (CUDA)");
  m.def("forward_ef", &sddmm_forward, "TC-GNN SDDMM forward (CUDA)");
  m.def("SDDMM_forward", &sddmm_forward, "TC-GNN SDDMM forward (CUDA) -- alias of forward_ef");
  m.def("forward_AGNN", &spmm_forward_AGNN, "TC-GNN SPMM (AGNN) forward (CUDA)");
  // backward (same kernels; the reference assumes a symmetric adjacency, gnn_conv.py:76-85)
  m.def("backward", &spmm_forward, "TC-GNN SPMM backward (CUDA)");
  m.def("backward_ef", &sddmm_forward, "TC-GNN SDDMM backward_ef (CUDA)");
  // additions
  m.def("preprocess_panel", &preprocess_panel,
        "SGT of a row panel: (edgeList, nodePointer, num_rows, num_cols, blk_h, blk_w, bp, e2c, e2r)");
  namespace py = pybind11;
  m.def("panel_forward", &panel_forward, "SpMM of a row panel: (X_all, row_base, nodePointer, edgeList, bp, e2c, e2r)",
        py::arg("input"), py::arg("row_base"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("blockPartition"),
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("x_is_tf32") = false,
        py::arg("accumulate_into") = py::none());
  m.def("source_forward", &source_forward,
        "partial SpMM over one source panel's rows: (X_src, nodePointer, edgeList, bp, e2c, e2r, x_is_tf32, accumulate_into)",
        py::arg("input"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("blockPartition"),
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("x_is_tf32") = false,
        py::arg("accumulate_into") = py::none());
  m.def("panel_forward_AGNN", &panel_forward_AGNN,
        "weighted SpMM of a row panel: (X_all, row_base, nodePointer, edgeList, edgeAttention, bp, e2c, e2r)",
        py::arg("input"), py::arg("row_base"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("edgeAttention"),
        py::arg("blockPartition"), py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("x_is_tf32") = false);
  m.def("panel_forward_ef", &panel_forward_ef, "SDDMM of a row panel: (X_all, row_base, nodePointer, edgeList, bp, e2c, e2r)",
        py::arg("input"), py::arg("row_base"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("blockPartition"),
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("x_is_tf32") = false);
  m.def("round_tf32_into", &round_tf32_into,
        "(input, out_ptr, ldo, multicast): cvt.rna.tf32(input) written to a raw (local / peer / multicast) address",
        py::arg("input"), py::arg("out_ptr"), py::arg("ldo"), py::arg("multicast") = false);
  m.def("forward_host", &forward_host,
        "SpMM with host (pinned) feature / result tensors: pipelined H2D copy, kernels, D2H copy",
        py::arg("x_host"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("blockPartition"),
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("y_host") = py::none(), py::arg("sync") = true);
  m.def("forward_ef_host", &forward_ef_host, "SDDMM with host feature / result tensors (tcgnn_sddmm_f32_host)",
        py::arg("x_host"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("blockPartition"),
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("edge_out_host") = py::none(), py::arg("sync") = true);
  m.def("forward_AGNN_host", &forward_AGNN_host, "fused AGNN with host feature / result tensors (tcgnn_agnn_f32_host)",
        py::arg("x_host"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("attention_w"),
        py::arg("blockPartition"), py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("y_host") = py::none(),
        py::arg("sync") = true);
  m.def("forward_AGNN_fused", &agnn_fused,
        "fused AGNN edge pipeline: (X, nodePointer, edgeList, attention_w, bp, e2c, e2r, want_edge_feature) -> "
        "[Y, attention (tile order), edge_feature (CSR order) or empty]",
        py::arg("input"), py::arg("nodePointer"), py::arg("edgeList"), py::arg("attention_w"),
        py::arg("blockPartition"), py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("want_edge_feature") = false);
  m.def("forward_AGNN_tile", &spmm_forward_AGNN_tile,
        "weighted SpMM with tile-ordered weights: (X, nodePointer, edgeList, att_tile, bp, e2c, e2r)");
  m.def("backward_T", &spmm_backward_T, "SpMM over the transposed graph: dX = A^T dY (directed graphs)");
  m.def("backward_T_AGNN", &spmm_backward_T_AGNN, "weighted SpMM over the transposed graph");
  m.def("csr_transpose", &csr_transpose, "(nodePointer, edgeList, num_cols=-1) -> [row_ptr_t, col_idx_t, edge_map_t]",
        py::arg("nodePointer"), py::arg("edgeList"), py::arg("num_cols") = -1);
  m.def("gather_rows", &gather_rows, "(src [n, d], rows int32 [m], dst [>= m, d]): dst[i] = src[rows[i]]");
  m.def("stream_wait_flag_dev", &stream_wait_flag_dev,
        "(flags int32, index, value_dev int32[1], timeout_ms, error_out int32[1]): like stream_wait_flag, the expected "
        "value read from device memory when the wait executes (CUDA-graph friendly)");
  m.def("stream_wait_flag", &stream_wait_flag,
        "(flags int32, index, value, timeout_ms, error_out int32[1]): block the current stream until flags[index] >= value");
  m.def("push_rows", &push_rows, "(local, peer_ptrs, seg_begin_rows, seg_end_rows): copy row segments to peers");
  m.def("round_tf32", &round_tf32, "cvt.rna.tf32 of a [rows, dim] CUDA matrix (dim % 4 == 0), for x_is_tf32 = True");
  m.def("clear_plan_cache", &clear_plan_cache, "Destroy all cached kernel plans");
  m.def("plan_info", &plan_info, "[num_nodes, num_edges, num_windows, num_tiles, plan_bytes, pairs, device, sms]");
  m.def("launch_count", [](bool reset) { return tcgnn_launch_count(reset ? 1 : 0); }, pybind11::arg("reset") = false,
        "kernels launched by this thread through the library");
  m.def("version", []() { return tcgnn_version(); });
}
