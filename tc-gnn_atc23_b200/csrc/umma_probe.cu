// Layout probe: one CTA, caller-supplied shared-memory images and UMMA descriptors, returns the
// raw TMEM accumulator.  tests/test_gpu_umma_layouts.py uses it to pin, on real hardware, every
// shared-memory layout / descriptor encoding the production kernels rely on (MN-major 128B-swizzled
// tf32 A tiles, K-major B tiles, the 32-byte K advance inside a swizzle atom, TMEM lane mapping).
#include <vector>

#include "plan.h"

namespace tcgnn {

namespace {

constexpr int kProbeMaxA = 96 * 1024;
constexpr int kProbeMaxB = 32 * 1024;
constexpr uint32_t kProbeTmemCols = 256;   // accumulator in [0, 64), A operand (adesc == 0) from column 64

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, int a_bytes, const uint8_t* __restrict__ b_img, int b_bytes,
                  uint64_t adesc, uint64_t bdesc, uint32_t idesc, int ksteps, int a_step, int b_step,
                  float* __restrict__ d_out, int ncols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + kProbeMaxA;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(a_smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(b_smem)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<kProbeTmemCols>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool a_in_tmem = adesc == 0;   // the A image is then row-major [128][8 * ksteps] fp32
  if (a_in_tmem) {
    // thread (warp, lane) owns TMEM lane 32 * warp + lane = row m of A; k-step s sits in columns 64 + 8 s ...
    const float* arow = reinterpret_cast<const float*>(a_smem) + (warp * 32 + lane) * 8 * ksteps;
    for (int k = 0; k < ksteps; ++k) {
      uint32_t v[8];
      for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(arow[k * 8 + i]);
      tmem_st_32x32b_x8(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 64 + k * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(a_smem), b_addr = smem_u32(b_smem);
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t ad = adesc + static_cast<uint64_t>(((a_addr + k * a_step) & 0x3FFFFu) >> 4);
      const uint64_t bd = bdesc + static_cast<uint64_t>(((b_addr + k * b_step) & 0x3FFFFu) >> 4);
      if (a_in_tmem) umma_tf32_ts(tmem_base, tmem_base + 64 + k * 8, bd, idesc, k > 0 ? 1u : 0u);
      else umma_tf32(tmem_base, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * ncols + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kProbeTmemCols>(tmem_base);
  }
}

// Throughput probe: `n_mma` back-to-back MMAs from one elected lane, rotating over `n_acc` accumulators
// (acc_stride TMEM columns apart) and over `n_a` / `n_b` operand tiles (a_step / b_step bytes apart).
// Shared memory holds zeros (the result is irrelevant).  out[0] = cycles until the last MMA was issued,
// out[1] = cycles until the commit after the last MMA arrived.
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(uint64_t adesc, uint64_t bdesc, uint32_t idesc, int n_mma, int n_acc, int acc_stride, int n_a,
                  int a_step, int n_b, int b_step, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kProbeMaxA + kProbeMaxB) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + kProbeMaxA);
    const uint64_t ad0 = adesc + static_cast<uint64_t>((a_addr & 0x3FFFFu) >> 4);
    const uint64_t bd0 = bdesc + static_cast<uint64_t>((b_addr & 0x3FFFFu) >> 4);
    const long long t0 = clock64();
    if (adesc == 0) {
      // A from tensor memory: groups of 8 MMAs over n_a operand slots of 8 columns (from column 64), B tiles 512 B apart
      int ia = 0;
      for (int i = 0; i < n_mma; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_tf32_ts(tmem_base, tmem_base + 64 + ((ia + j) % n_a) * 8, bd0 + static_cast<uint64_t>((j * 512) >> 4),
                         idesc, 1u);
        }
        ia = (ia + 8) % n_a;
      }
    } else if (n_acc == 0) {
      // production pattern: groups of 8 MMAs with compile-time operand offsets (4 KB A tiles, 512 B B tiles)
      for (int i = 0; i < n_mma; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_tf32(tmem_base, ad0 + static_cast<uint64_t>((j * 4096) >> 4), bd0 + static_cast<uint64_t>((j * 512) >> 4),
                      idesc, 1u);
        }
      }
    } else {
      int ia = 0, ib = 0, ic = 0;
      for (int i = 0; i < n_mma; ++i) {
        if (elect_one())
          umma_tf32(tmem_base + ic * acc_stride, ad0 + static_cast<uint64_t>((ia * a_step) >> 4),
                    bd0 + static_cast<uint64_t>((ib * b_step) >> 4), idesc, 1u);
        if (++ia == n_a) ia = 0;
        if (++ib == n_b) ib = 0;
        if (++ic == n_acc) ic = 0;
      }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int debug_umma_bench(uint64_t adesc, uint64_t bdesc, uint32_t idesc, int32_t n_mma, int32_t n_acc, int32_t acc_stride,
                     int32_t n_a, int32_t a_step, int32_t n_b, int32_t b_step, int32_t grid, int64_t* cycles_out,
                     cudaStream_t stream) {
  if (cycles_out == nullptr || n_mma < 1 || n_acc < 0 || n_a < 1 || n_b < 1 || grid < 1 ||
      static_cast<int64_t>(n_acc) * acc_stride > 512 ||
      (adesc != 0 ? static_cast<int64_t>(n_a) * a_step > kProbeMaxA : n_a > 56) ||
      static_cast<int64_t>(n_b) * b_step > kProbeMaxB) {
    set_last_error("tcgnn_debug_umma_bench: bad argument");
    return TCGNN_ERR_INVALID_ARG;
  }
  long long* d = nullptr;
  long long h[2] = {0, 0};
  const int smem_bytes = kProbeMaxA + kProbeMaxB + 1024;
  cudaError_t e = cudaMalloc(&d, sizeof(h));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) {
    umma_bench_kernel<<<grid, 128, smem_bytes, stream>>>(adesc, bdesc, idesc, n_mma, n_acc, acc_stride, n_a, a_step,
                                                         n_b, b_step, d);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (d) cudaFree(d);
  if (e != cudaSuccess) {
    set_last_error("tcgnn_debug_umma_bench: %s", cudaGetErrorString(e));
    return TCGNN_ERR_CUDA;
  }
  cycles_out[0] = h[0];
  cycles_out[1] = h[1];
  return TCGNN_OK;
}

int debug_umma(const void* a_image, int32_t a_bytes, const void* b_image, int32_t b_bytes, uint64_t adesc,
               uint64_t bdesc, uint32_t idesc, int32_t ksteps, int32_t a_step_bytes, int32_t b_step_bytes,
               float* d_out, int32_t ncols, cudaStream_t stream) {
  if (a_image == nullptr || b_image == nullptr || d_out == nullptr || a_bytes <= 0 || b_bytes <= 0 ||
      a_bytes % 16 != 0 || b_bytes % 16 != 0 || a_bytes > kProbeMaxA || b_bytes > kProbeMaxB || ksteps < 1 ||
      (adesc == 0 && (a_bytes != 128 * 8 * 4 * ksteps || ksteps > 24)) ||
      (ncols != 16 && ncols != 32 && ncols != 64)) {
    set_last_error("tcgnn_debug_umma: bad argument");
    return TCGNN_ERR_INVALID_ARG;
  }
  uint8_t *da = nullptr, *db = nullptr;
  float* dd = nullptr;
  int status = TCGNN_ERR_CUDA;
  const int smem_bytes = kProbeMaxA + kProbeMaxB + 1024;
  cudaError_t e;
  if ((e = cudaMalloc(&da, a_bytes)) != cudaSuccess) goto done;
  if ((e = cudaMalloc(&db, b_bytes)) != cudaSuccess) goto done;
  if ((e = cudaMalloc(&dd, sizeof(float) * 128 * ncols)) != cudaSuccess) goto done;
  if ((e = cudaMemcpyAsync(da, a_image, a_bytes, cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto done;
  if ((e = cudaMemcpyAsync(db, b_image, b_bytes, cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto done;
  if ((e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)) !=
      cudaSuccess)
    goto done;
  umma_probe_kernel<<<1, 128, smem_bytes, stream>>>(da, a_bytes, db, b_bytes, adesc, bdesc, idesc, ksteps,
                                                    a_step_bytes, b_step_bytes, dd, ncols);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) goto done;
  if ((e = cudaMemcpyAsync(d_out, dd, sizeof(float) * 128 * ncols, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
    goto done;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) goto done;
  status = TCGNN_OK;
done:
  if (status != TCGNN_OK) set_last_error("tcgnn_debug_umma: %s", cudaGetErrorString(e));
  if (da) cudaFree(da);
  if (db) cudaFree(db);
  if (dd) cudaFree(dd);
  return status;
}

}  // namespace tcgnn
