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
struct TensorKey {
  const void* ptr = nullptr;
  int64_t numel = 0;
  uint32_t version = 0;
  c10::weak_intrusive_ptr<c10::StorageImpl> storage;

  explicit TensorKey(const torch::Tensor& t)
      : ptr(t.data_ptr()),
        numel(t.numel()),
        version(t.is_inference() ? 0u : static_cast<uint32_t>(t._version())),
        storage(c10::weak_intrusive_ptr<c10::StorageImpl>(t.storage().getWeakStorageImpl())) {}

  bool matches(const torch::Tensor& t) const {
    if (storage.expired()) return false;
    if (t.data_ptr() != ptr || t.numel() != numel) return false;
    if (t.storage().unsafeGetStorageImpl() != storage._unsafe_get_target()) return false;
    const uint32_t v = t.is_inference() ? 0u : static_cast<uint32_t>(t._version());
    return v == version;
  }
  bool alive() const { return !storage.expired(); }
};

struct PlanEntry {
  std::vector<TensorKey> keys;   // nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow
  int device = 0;
  int64_t num_cols = 0;
  int64_t row_base = 0;
  tcgnn_plan* plan = nullptr;
  std::vector<torch::Tensor> keep;   // nothing kept by default (inputs are borrowed, never retained)
};

std::mutex g_cache_mu;
std::list<PlanEntry> g_cache;   // most recently used first
constexpr size_t kCacheCapacity = 8;

tcgnn_plan* get_plan(const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                     const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                     const torch::Tensor& edgeToRow, int64_t num_cols = -1, int64_t row_base = 0) {
  if (num_cols < 0) num_cols = nodePointer.size(0) - 1;
  const torch::Tensor* ts[5] = {&nodePointer, &edgeList, &blockPartition, &edgeToColumn, &edgeToRow};
  const int device = nodePointer.get_device();
  std::lock_guard<std::mutex> lock(g_cache_mu);
  for (auto it = g_cache.begin(); it != g_cache.end();) {
    bool alive = true;
    for (const auto& k : it->keys) alive = alive && k.alive();
    if (!alive) {   // the graph tensors were freed: drop the plan
      tcgnn_plan_destroy(it->plan);
      it = g_cache.erase(it);
      continue;
    }
    bool hit = it->device == device && it->num_cols == num_cols && it->row_base == row_base;
    for (int i = 0; hit && i < 5; ++i) hit = it->keys[i].matches(*ts[i]);
    if (hit) {
      g_cache.splice(g_cache.begin(), g_cache, it);
      return g_cache.front().plan;
    }
    ++it;
  }
  const int64_t num_nodes = nodePointer.size(0) - 1;
  const int64_t num_edges = edgeList.size(0);
  TORCH_CHECK(num_nodes >= 1, "nodePointer must have at least two entries");
  TORCH_CHECK(num_nodes <= INT32_MAX && num_edges <= INT32_MAX, "graph too large for int32 CSR");
  TORCH_CHECK(edgeToColumn.size(0) == num_edges && edgeToRow.size(0) == num_edges,
              "edgeToColumn / edgeToRow must have one entry per edge");
  TORCH_CHECK(blockPartition.size(0) == (num_nodes + TCGNN_BLK_H - 1) / TCGNN_BLK_H,
              "blockPartition must have ceil(num_nodes / 16) entries");
  PlanEntry entry;
  for (int i = 0; i < 5; ++i) entry.keys.emplace_back(*ts[i]);
  entry.device = device;
  entry.num_cols = num_cols;
  entry.row_base = row_base;
  TORCH_CHECK(num_cols <= INT32_MAX && row_base >= 0 && row_base + num_nodes <= num_cols,
              "row panel [", row_base, ", ", row_base + num_nodes, ") does not fit a graph of ", num_cols, " nodes");
  auto stream = c10::cuda::getCurrentCUDAStream(device).stream();
  const int status = tcgnn_plan_create_panel(
      nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(), blockPartition.data_ptr<int32_t>(),
      edgeToColumn.data_ptr<int32_t>(), edgeToRow.data_ptr<int32_t>(), static_cast<int32_t>(num_nodes),
      static_cast<int32_t>(num_cols), static_cast<int32_t>(row_base), num_edges,
      static_cast<int32_t>(blockPartition.size(0)), stream, &entry.plan);
  check_status(status, "tcgnn_plan_create");
  g_cache.push_front(std::move(entry));
  while (g_cache.size() > kCacheCapacity) {
    tcgnn_plan_destroy(g_cache.back().plan);
    g_cache.pop_back();
  }
  return g_cache.front().plan;
}

void clear_plan_cache() {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  for (auto& e : g_cache) tcgnn_plan_destroy(e.plan);
  g_cache.clear();
}

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
  if (row_base < 0) {
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
torch::Tensor run_spmm(const torch::Tensor& input, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                       const torch::Tensor* edgeAttention, const torch::Tensor& blockPartition,
                       const torch::Tensor& edgeToColumn, const torch::Tensor& edgeToRow, int64_t row_base,
                       bool x_is_tf32 = false) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, row_base);
  const float* weights = nullptr;
  if (edgeAttention != nullptr) {
    CHECK_INPUT(*edgeAttention);
    CHECK_F32(*edgeAttention);
    TORCH_CHECK(edgeAttention->device() == input.device(), "edgeAttention must be on the same device as input");
    // [n_heads, E]; like the reference kernel (TCGNN_kernel.cu:529) only head 0 is read.
    TORCH_CHECK(edgeAttention->dim() >= 1 && edgeAttention->size(-1) == edgeList.size(0),
                "edgeAttention must be [n_heads, num_edges]");
    weights = edgeAttention->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(input.device());
  tcgnn_plan* plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow,
                              row_base < 0 ? -1 : input.size(0), row_base < 0 ? 0 : row_base);
  auto output = torch::empty({nodePointer.size(0) - 1, input.size(1)}, input.options());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_spmm_f32_ex(plan, input.data_ptr<float>(), input.size(1), weights, output.data_ptr<float>(),
                                 output.size(1), static_cast<int32_t>(input.size(1)),
                                 x_is_tf32 ? TCGNN_X_IS_TF32 : 0u, stream),
               "tcgnn_spmm_f32");
  return output;
}

torch::Tensor run_sddmm(const torch::Tensor& input, const torch::Tensor& nodePointer, const torch::Tensor& edgeList,
                        const torch::Tensor& blockPartition, const torch::Tensor& edgeToColumn,
                        const torch::Tensor& edgeToRow, int64_t row_base, bool x_is_tf32 = false) {
  check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, row_base);
  c10::cuda::CUDAGuard guard(input.device());
  tcgnn_plan* plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow,
                              row_base < 0 ? -1 : input.size(0), row_base < 0 ? 0 : row_base);
  auto output = torch::empty({edgeList.size(0)}, input.options());
  auto stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
  check_status(tcgnn_sddmm_f32_ex(plan, input.data_ptr<float>(), input.size(1), output.data_ptr<float>(),
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

// Row-panel variants for 1-D destination-row sharding (new; the reference is single-GPU): `input` is the
// all-gathered feature matrix of the whole graph, the five graph tensors describe the caller's row panel
// (tcgnn_plan_create_panel), `row_base` is the global id of the panel's first row.
std::vector<torch::Tensor> panel_forward(torch::Tensor input, int64_t row_base, torch::Tensor nodePointer,
                                         torch::Tensor edgeList, torch::Tensor blockPartition,
                                         torch::Tensor edgeToColumn, torch::Tensor edgeToRow, bool x_is_tf32) {
  TORCH_CHECK(row_base >= 0, "row_base must be >= 0");
  return {run_spmm(input, nodePointer, edgeList, nullptr, blockPartition, edgeToColumn, edgeToRow, row_base, x_is_tf32)};
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
  tcgnn_plan* plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  std::vector<int64_t> info(8, 0);
  check_status(tcgnn_plan_info(plan, info.data()), "tcgnn_plan_info");
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

// SpMM with host feature / result tensors (tcgnn_spmm_f32_host): X and Y are CPU tensors (pinned for full speed),
// the graph tensors are CUDA tensors.  Returns y_host; with sync = false the caller must synchronise the current
// stream (or an event recorded on it) before reading it.
torch::Tensor forward_host(torch::Tensor x_host, torch::Tensor nodePointer, torch::Tensor edgeList,
                           torch::Tensor blockPartition, torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                           c10::optional<torch::Tensor> y_host_opt, bool sync) {
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
  const int64_t n = nodePointer.size(0) - 1;
  TORCH_CHECK(x_host.size(0) == n, "x_host has ", x_host.size(0), " rows but the graph has ", n, " nodes");
  torch::Tensor y_host = y_host_opt.has_value()
                             ? *y_host_opt
                             : torch::empty({n, x_host.size(1)}, x_host.options().pinned_memory(true));
  TORCH_CHECK(!y_host.is_cuda() && y_host.is_contiguous() && y_host.scalar_type() == torch::kFloat32 &&
                  y_host.dim() == 2 && y_host.size(0) == n && y_host.size(1) == x_host.size(1),
              "y_host must be a contiguous float32 CPU tensor [num_nodes, dim]");
  c10::cuda::CUDAGuard guard(nodePointer.device());
  tcgnn_plan* plan = get_plan(nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow);
  auto stream = c10::cuda::getCurrentCUDAStream(nodePointer.get_device()).stream();
  check_status(tcgnn_spmm_f32_host(plan, x_host.data_ptr<float>(), x_host.size(1), nullptr, y_host.data_ptr<float>(),
                                   y_host.size(1), static_cast<int32_t>(x_host.size(1)), stream),
               "tcgnn_spmm_f32_host");
  if (sync) C10_CUDA_CHECK(cudaStreamSynchronize(stream));
  return y_host;
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
        py::arg("edgeToColumn"), py::arg("edgeToRow"), py::arg("x_is_tf32") = false);
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
  m.def("push_rows", &push_rows, "(local, peer_ptrs, seg_begin_rows, seg_end_rows): copy row segments to peers");
  m.def("round_tf32", &round_tf32, "cvt.rna.tf32 of a [rows, dim] CUDA matrix (dim % 4 == 0), for x_is_tf32 = True");
  m.def("clear_plan_cache", &clear_plan_cache, "Destroy all cached kernel plans");
  m.def("plan_info", &plan_info, "[num_nodes, num_edges, num_windows, num_tiles, plan_bytes, pairs, device, sms]");
  m.def("launch_count", [](bool reset) { return tcgnn_launch_count(reset ? 1 : 0); }, pybind11::arg("reset") = false,
        "kernels launched by this thread through the library");
  m.def("version", []() { return tcgnn_version(); });
}
