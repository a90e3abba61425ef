// Operand preparation shared by SpMM and SDDMM: Xr = cvt.rna.tf32.f32(X), written once per call
// into a plan-owned scratch with a 16-byte aligned leading dimension.
//
// The reference rounds every operand element with wmma::__float_to_tf32 inside its inner loops
// (/root/reference TCGNN_conv/TCGNN_kernel.cu:436-444, :560-567, :701-709), i.e. once per USE.
// tcgen05.mma reads fp32 bit patterns from shared memory and ignores the low 13 mantissa bits
// (truncation), so feeding it raw X would bias every product.  Rounding once per ELEMENT here gives
// bit-identical operands to the reference's and lets the gather kernels move rows with cp.async
// (no register pass).  HBM-bound elementwise pass: 2 * N * D * 4 bytes, grid-stride, 128-bit accesses.
#include "plan.h"

namespace tcgnn {

constexpr int kMaxPushPeers = 16;
constexpr int kMaxPushSegs = 16;

namespace {

__global__ void __launch_bounds__(256)
tf32_round_pack_kernel(const float* __restrict__ x, int64_t ldx, int32_t dim, int64_t rows, float* __restrict__ out,
                       int64_t ldr, int vec_ok, int multimem) {
  const int32_t nvec = static_cast<int32_t>(ldr >> 2);
  const int64_t total = rows * nvec;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / nvec;
    const int32_t f = static_cast<int32_t>(i - r * nvec) * 4;
    const float* src = x + r * ldx + f;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec_ok && f + 4 <= dim) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      if (f + 0 < dim) v.x = __ldg(src + 0);
      if (f + 1 < dim) v.y = __ldg(src + 1);
      if (f + 2 < dim) v.z = __ldg(src + 2);
      if (f + 3 < dim) v.w = __ldg(src + 3);
    }
    const float4 o = tf32_rna4(v);
    float* dst = out + r * ldr + f;
    if (multimem) {
      // `out` is an NVSwitch multicast address: one store lands in the buffer of every GPU of the group
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y),
                   "f"(o.z), "f"(o.w)
                   : "memory");
    } else {
      *reinterpret_cast<float4*>(dst) = o;
    }
  }
}

}  // namespace

int round_pack_launch(tcgnn_plan* plan, const float* x, int64_t ldx, int32_t dim, int64_t ldr, cudaStream_t stream,
                      const float** xr_out) {
  const int64_t rows = plan->num_cols;
  const size_t need = static_cast<size_t>(rows) * static_cast<size_t>(ldr);
  {
    std::lock_guard<std::mutex> lock(plan->mu);
    if (plan->x_round_cap < need) {
      if (plan->x_round != nullptr) {
        // the old buffer may still be read by kernels queued on `stream`
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaFree(plan->x_round);
        plan->x_round = nullptr;
        plan->x_round_cap = 0;
        if (e != cudaSuccess) {
          set_last_error("releasing the tf32 scratch failed: %s", cudaGetErrorString(e));
          return TCGNN_ERR_CUDA;
        }
      }
      cudaError_t e = cudaMalloc(&plan->x_round, need * sizeof(float));
      if (e != cudaSuccess) {
        plan->x_round = nullptr;
        set_last_error("cudaMalloc(%zu bytes of tf32 scratch) failed: %s", need * sizeof(float), cudaGetErrorString(e));
        return TCGNN_ERR_OOM;
      }
      plan->x_round_cap = need;
    }
  }
  const int vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ldx % 4 == 0);
  const int64_t total = rows * (ldr >> 2);
  int64_t g = (total + 255) / 256;
  if (g > static_cast<int64_t>(plan->num_sms) * 16) g = static_cast<int64_t>(plan->num_sms) * 16;
  if (g < 1) g = 1;
  tf32_round_pack_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, ldx, dim, rows, plan->x_round, ldr, vec_ok, 0);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("tf32_round_pack_kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  *xr_out = plan->x_round;
  return TCGNN_OK;
}

// Second phase of the balanced exchange (sharding.py): copy row segments of the local gathered matrix to the same
// rows of every peer's copy (P2P-mapped pointers), one launch for all peers and segments.
struct PushArgs {
  float* peers[kMaxPushPeers];
  int64_t seg_begin[kMaxPushSegs];   // in 16-byte vectors from the matrix base
  int64_t seg_vecs_prefix[kMaxPushSegs + 1];
  int32_t n_peers, n_segs;
};

namespace {
__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, PushArgs a) {
  const int64_t total = a.seg_vecs_prefix[a.n_segs];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int p = blockIdx.y; p < a.n_peers; p += gridDim.y) {
    float4* dst = reinterpret_cast<float4*>(a.peers[p]);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
      int sgm = 0;
#pragma unroll
      for (int k = 1; k < kMaxPushSegs; ++k) sgm += (k < a.n_segs && i >= a.seg_vecs_prefix[k]) ? 1 : 0;
      const int64_t v = a.seg_begin[sgm] + (i - a.seg_vecs_prefix[sgm]);
      dst[v] = __ldg(s4 + v);
    }
  }
}
}  // namespace

int push_rows_launch(const float* src, float* const* peers, int32_t n_peers, const int64_t* seg_begin_rows,
                     const int64_t* seg_end_rows, int32_t n_segs, int64_t ld, cudaStream_t stream) {
  PushArgs a;
  a.n_peers = n_peers;
  a.n_segs = n_segs;
  a.seg_vecs_prefix[0] = 0;
  for (int i = 0; i < n_segs; ++i) {
    a.seg_begin[i] = seg_begin_rows[i] * (ld >> 2);
    a.seg_vecs_prefix[i + 1] = a.seg_vecs_prefix[i] + (seg_end_rows[i] - seg_begin_rows[i]) * (ld >> 2);
  }
  for (int i = 0; i < n_peers; ++i) a.peers[i] = peers[i];
  const int64_t total = a.seg_vecs_prefix[n_segs];
  if (total == 0 || n_peers == 0) return TCGNN_OK;
  int64_t gx = (total + 255) / 256;
  const int64_t cap = 148 * 8 / (n_peers > 0 ? n_peers : 1) + 1;
  if (gx > cap) gx = cap;
  push_rows_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_peers)), 256, 0, stream>>>(src, a);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("push_rows_kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  return TCGNN_OK;
}

int round_tf32_launch(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t dim, int multimem,
                      cudaStream_t stream) {
  if (rows <= 0) return TCGNN_OK;
  const int vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ldx % 4 == 0);
  const int64_t total = rows * (ldo >> 2);
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  tf32_round_pack_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, ldx, dim, rows, out, ldo, vec_ok, multimem);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("tf32_round_pack_kernel launch failed: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  return TCGNN_OK;
}

}  // namespace tcgnn
