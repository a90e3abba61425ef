// Do LDGSTS fills and tcgen05 operand fetches share the SM's shared-memory datapath?
//
// spmm_tc_kernel moves every gathered byte through shared memory twice: the cp.async (LDGSTS) fill and the MMA's
// operand fetch (M128 N16 K8 tf32 from shared memory = 4608 bytes per MMA; 39.2 cycles back to back = 118 B/clk).
// tools/l2_gather_bench measures the fills alone (~70 B/clk per SM), tools/umma_bench.py the MMAs alone.  This tool
// runs both at once in one CTA per SM -- `warps` gathering warps exactly as in l2_gather_bench plus one warp that
// keeps issuing MMAs whose A operands are the tiles being filled (results are garbage; only the traffic matters) --
// and reports both rates.  If the two used separate paths, both would keep their stand-alone rate; if they take turns
// on one path, time per tile = fill time + fetch time.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/build/smem_datapath_bench tools/smem_datapath_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../tc-gnn_atc23_b200/csrc/common.cuh"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

using namespace tcgnn;

constexpr int kDim = 128;
constexpr int kTileBytes = 8 * kDim * 4;

__device__ __forceinline__ int32_t row_of(uint64_t i, uint32_t rows) {
  uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return static_cast<int32_t>(__umulhi(static_cast<uint32_t>(h), rows));
}

struct Out {
  unsigned long long gather_cycles;   // max over CTAs: until the last gathering warp's copies had landed
  unsigned long long mma_cycles;      // max over CTAs: until the MMA warp's last commit arrived
  unsigned long long mmas;            // sum over CTAs
};

// mma_mode 0: no MMAs; 1: MMAs for as long as the gathers run (or `mma_fixed` of them when there are no gathers)
template <int DEPTH>
__global__ void fills_and_mma(const float* __restrict__ x, uint32_t rows, int tiles_per_warp, int gather_warps,
                              int mma_mode, int mma_fixed, Out* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ring_tiles = gather_warps * DEPTH;
  const uint32_t b_tile = smem + ring_tiles * kTileBytes;       // 512 B
  const uint32_t bars = b_tile + 512;                           // two MMA barriers
  const uint32_t done_cnt = bars + 16;                          // gathering warps that have finished
  const uint32_t tmem_slot = done_cnt + 4;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(done_cnt), "r"(0u) : "memory");
    fence_mbar_init();
  }
  if (warp == gather_warps) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const long long t0 = clock64();
  if (warp < gather_warps) {
    const uint32_t ring = smem + warp * DEPTH * kTileBytes;
    const int64_t first = (static_cast<int64_t>(blockIdx.x) * gather_warps + warp) * tiles_per_warp;
    for (int i = 0; i < tiles_per_warp; ++i) {
      const int32_t my_row = row_of(static_cast<uint64_t>(first + i) * 8 + (lane & 7), rows);
      if (i >= DEPTH) asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
      const uint32_t tile = ring + (i % DEPTH) * kTileBytes;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int32_t row = __shfl_sync(0xffffffffu, my_row, r);
        const float* src = x + static_cast<int64_t>(row) * kDim + lane * 4;
        const uint32_t dst = tile + (lane >> 3) * 1024 + sw128_base32_offset(r, lane & 7);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      uint32_t prev;
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(done_cnt) : "memory");
      if (static_cast<int>(prev) == gather_warps - 1 && tiles_per_warp > 0)
        atomicMax(&out->gather_cycles, static_cast<unsigned long long>(clock64() - t0));
    }
  } else if (mma_mode != 0) {
    constexpr uint32_t idesc = make_idesc_tf32(128, 16, true, false);
    const uint64_t adesc0 = make_smem_desc(0, 1024, 512, kSwizzle128BBase32B);
    const uint64_t bdesc = make_smem_desc(b_tile, 128, 256, kSwizzleNone);
    constexpr int kGroup = 64;   // MMAs per commit; two groups in flight keep the tensor pipe's queue full
    unsigned long long issued = 0;
    uint32_t ph[2] = {0, 0};
    int a = 0;
    for (int g = 0;; ++g) {
      if (g >= 2) {   // the group before the previous one has completed: at most two groups in flight
        mbar_wait(bars + 8 * (g & 1), ph[g & 1]);
        ph[g & 1] ^= 1u;
      }
      bool stop;
      if (tiles_per_warp > 0) {
        uint32_t d;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d) : "r"(done_cnt));
        stop = static_cast<int>(d) >= gather_warps;
      } else {
        stop = issued >= static_cast<unsigned long long>(mma_fixed);
      }
      if (stop) {
        if (g >= 1) mbar_wait(bars + 8 * ((g - 1) & 1), ph[(g - 1) & 1]);
        break;
      }
      if (elect_one()) {
        for (int jj = 0; jj < kGroup; jj += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a_addr = smem + ((a + jj + j) % ring_tiles) * kTileBytes;
            umma_tf32(tmem_base, adesc0 | static_cast<uint64_t>((a_addr & 0x3FFFFu) >> 4), bdesc, idesc, 1u);
          }
        }
        umma_commit(bars + 8 * (g & 1));
      }
      __syncwarp();
      a = (a + kGroup) % ring_tiles;
      issued += kGroup;
    }
    if (lane == 0) {
      atomicMax(&out->mma_cycles, static_cast<unsigned long long>(clock64() - t0));
      atomicAdd(&out->mmas, issued);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == gather_warps) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

__global__ void fill_kernel(float* x, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] = static_cast<float>((i * 2654435761ull) & 0xFFFF) * (1.0f / 65536.0f);
}

template <int DEPTH>
static void run(const char* what, float* x, uint32_t rows, double ws_mb, int grid, int warps, int tiles_per_warp,
                int mma_mode, int mma_fixed, Out* d_out) {
  const size_t smem = static_cast<size_t>(warps) * DEPTH * kTileBytes + 512 + 64 + 1024;
  CK(cudaFuncSetAttribute(fills_and_mma<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  Out best{};
  double best_t = 1e30;
  for (int rep = 0; rep < 4; ++rep) {   // first repetition warms the L2
    CK(cudaMemset(d_out, 0, sizeof(Out)));
    fills_and_mma<DEPTH><<<grid, (warps + 1) * 32, smem>>>(x, rows, tiles_per_warp, warps, mma_mode, mma_fixed, d_out);
    CK(cudaDeviceSynchronize());
    Out h;
    CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    const double t = static_cast<double>(h.gather_cycles ? h.gather_cycles : h.mma_cycles);
    if (rep > 0 && t < best_t) { best_t = t; best = h; }
  }
  const double tile_bytes_per_sm = static_cast<double>(warps) * tiles_per_warp * kTileBytes;
  const double fill_bpc = best.gather_cycles ? tile_bytes_per_sm / best.gather_cycles : 0.0;
  const double mma_per_sm = static_cast<double>(best.mmas) / grid;
  const double cyc_per_mma = best.mmas ? best.mma_cycles / mma_per_sm : 0.0;
  printf("{\"what\": \"%s\", \"working_set_mb\": %.1f, \"gather_warps\": %d, \"depth\": %d, \"fill_bytes_per_clk_per_sm\": %.1f, "
         "\"fill_cycles_per_tile\": %.1f, \"mma_per_sm\": %.0f, \"cycles_per_mma\": %.1f, \"fetch_bytes_per_clk_per_sm\": %.1f, "
         "\"tiles_per_sm\": %d}\n",
         what, ws_mb, warps, DEPTH, fill_bpc, fill_bpc > 0 ? kTileBytes / fill_bpc : 0.0, mma_per_sm, cyc_per_mma,
         cyc_per_mma > 0 ? 4608.0 / cyc_per_mma : 0.0, warps * tiles_per_warp);
  fflush(stdout);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount;
  float* x = nullptr;
  Out* d_out = nullptr;
  CK(cudaMalloc(&x, 128ll << 20));
  CK(cudaMalloc(&d_out, sizeof(Out)));
  for (double ws_mb : {48.0, 119.3}) {
    const int64_t rows = static_cast<int64_t>(ws_mb * 1e6 / (kDim * 4));
    fill_kernel<<<grid * 8, 256>>>(x, rows * kDim);
    CK(cudaDeviceSynchronize());
    const uint32_t r = static_cast<uint32_t>(rows);
    run<6>("MMAs alone", x, r, ws_mb, grid, 6, 0, 1, 16384, d_out);
    run<6>("fills alone", x, r, ws_mb, grid, 6, 2048, 0, 0, d_out);
    run<6>("fills + MMAs", x, r, ws_mb, grid, 6, 2048, 1, 0, d_out);
    run<4>("fills alone", x, r, ws_mb, grid, 12, 1024, 0, 0, d_out);
    run<4>("fills + MMAs", x, r, ws_mb, grid, 12, 1024, 1, 0, d_out);
  }
  return 0;
}
