// L2 -> SM gather ceiling of a B200, measured with the SpMM kernel's own access shape -- and the TMA gather4 A/B.
//
// What bounds spmm_tc_kernel is how fast 148 SMs can pull random 512-byte feature rows (D = 128 fp32) out of an
// L2-resident matrix into shared memory.  This micro-benchmark does only that (no MMA, no barriers between warps, no
// epilogue), three ways:
//   ldgsts        cp.async 16-byte copies, one warp instruction per row, 8 rows = one 4 KB tile, straight into the
//                 SWIZZLE_128B_BASE32B image tcgen05 needs for MN-major tf32 operands (what the kernel does today);
//   gather4       TMA cp.async.bulk.tensor.2d.tile::gather4 with a CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B tensor map,
//                 box {32 floats, 1}: 4 rows x 128 B per instruction, 8 instructions per tile, one issuing thread per
//                 warp, completion by mbarrier transaction bytes -- the SAME shared-memory image (verified below);
//   gather4_wide  the same with box {128 floats, 1} and no swizzle: 4 rows x 512 B per instruction, 2 per tile -- not
//                 consumable by the MMA (no swizzle), an upper bound for what the TMA engine can gather.
// Sweeps warps per CTA x tiles in flight per warp x working-set size (L2-resident to beyond L2) and prints one JSON
// line per configuration: GB/s and bytes per SM clock (whole chip).  tools/l2_gather_bench.py turns the best
// L2-resident ldgsts figure into profiles/l2_gather_peak.json, the denominator of bench.py's `l2_gather`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/build/l2_gather_bench tools/l2_gather_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int kDim = 128;           // floats per row: 512 B
constexpr int kTileBytes = 8 * kDim * 4;
// row id of slot `i` (tile * 8 + row): a hash of the global slot number, so every warp gathers its own random rows and
// nothing but the working-set size decides the L2 hit rate (a shared index ring made warps on different SMs walk the
// same rows at the same time: profiles/r02a_l2_gather_sweep_shared_index_ring.jsonl)
__device__ __forceinline__ int32_t row_of(uint64_t i, uint32_t rows) {
  uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return static_cast<int32_t>(__umulhi(static_cast<uint32_t>(h), rows));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t sw128_base32_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ ((row & 3u) << 1)) << 4);
}

// ------------------------------------------------------------------------------------------------------------------
// LDGSTS
// ------------------------------------------------------------------------------------------------------------------
template <int DEPTH>
__global__ void gather_ldgsts(const float* __restrict__ x, uint32_t rows, int tiles_per_warp,
                              unsigned long long* max_cycles, float* dump) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const uint32_t ring = smem + warp * DEPTH * kTileBytes;
  const long long t0 = clock64();
  const int64_t first = (static_cast<int64_t>(blockIdx.x) * warps + warp) * tiles_per_warp;
  for (int i = 0; i < tiles_per_warp; ++i) {
    const int32_t my_row = row_of(static_cast<uint64_t>(first + i) * 8 + (lane & 7), rows);
    if (i >= DEPTH) asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    const uint32_t tile = ring + (i % DEPTH) * kTileBytes;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int32_t row = __shfl_sync(0xffffffffu, my_row, r);
      const float* src = x + static_cast<int64_t>(row) * kDim + lane * 4;
      // lane -> 16-byte chunk `lane` of the row: feature block lane / 8 (1 KB atoms), chunk lane % 8
      const uint32_t dst = tile + (lane >> 3) * 1024 + sw128_base32_offset(r, lane & 7);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(max_cycles, static_cast<unsigned long long>(clock64() - t0));
  if (dump != nullptr && blockIdx.x == 0 && warp == 0) {   // last tile of warp 0: the image the MMA would read
    const uint32_t tile = ring + ((tiles_per_warp - 1) % DEPTH) * kTileBytes;
    for (int i = lane; i < kTileBytes / 4; i += 32) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tile + i * 4));
      dump[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA gather4
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) {   // ~2 s: a wrong transaction count must not hang the GPU
      printf("l2_gather_bench: mbarrier wait timed out (block %d)\n", blockIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                        uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(dst),
      "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

// WIDE = false: swizzled 128-byte boxes into the MMA image (8 instructions per tile); true: 512-byte boxes, row-major
template <int DEPTH, bool WIDE>
__global__ void gather_tma(const __grid_constant__ CUtensorMap map, uint32_t rows, int tiles_per_warp,
                           unsigned long long* max_cycles, float* dump) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const uint32_t ring = smem + warp * DEPTH * kTileBytes;
  const uint32_t bars = smem + warps * DEPTH * kTileBytes + warp * DEPTH * 8;
  if (lane == 0)
    for (int s = 0; s < DEPTH; ++s) mbar_init(bars + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  const int64_t first = (static_cast<int64_t>(blockIdx.x) * warps + warp) * tiles_per_warp;
  if (lane == 0) {
    for (int i = 0; i < tiles_per_warp; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mbar_wait(bars + 8 * s, ((i / DEPTH) - 1) & 1);   // the slot's previous tile has landed
      const uint64_t t8 = static_cast<uint64_t>(first + i) * 8;
      const int4 ra = make_int4(row_of(t8, rows), row_of(t8 + 1, rows), row_of(t8 + 2, rows), row_of(t8 + 3, rows));
      const int4 rb = make_int4(row_of(t8 + 4, rows), row_of(t8 + 5, rows), row_of(t8 + 6, rows), row_of(t8 + 7, rows));
      const uint32_t tile = ring + s * kTileBytes;
      mbar_expect(bars + 8 * s, kTileBytes);
      if (WIDE) {
        gather4(tile, &map, 0, ra.x, ra.y, ra.z, ra.w, bars + 8 * s);
        gather4(tile + 2048, &map, 0, rb.x, rb.y, rb.z, rb.w, bars + 8 * s);
      } else {
#pragma unroll
        for (int fb = 0; fb < 4; ++fb) {   // feature block fb: 1 KB atom pair (rows 0-3 | rows 4-7), 32 features wide
          gather4(tile + fb * 1024, &map, fb * 32, ra.x, ra.y, ra.z, ra.w, bars + 8 * s);
          gather4(tile + fb * 1024 + 512, &map, fb * 32, rb.x, rb.y, rb.z, rb.w, bars + 8 * s);
        }
      }
    }
    const int n_tail = tiles_per_warp < DEPTH ? tiles_per_warp : DEPTH;
    for (int j = 0; j < n_tail; ++j) {
      const int i = tiles_per_warp - n_tail + j;
      mbar_wait(bars + 8 * (i % DEPTH), (i / DEPTH) & 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(max_cycles, static_cast<unsigned long long>(clock64() - t0));
  if (dump != nullptr && blockIdx.x == 0 && warp == 0) {
    const uint32_t tile = ring + ((tiles_per_warp - 1) % DEPTH) * kTileBytes;
    for (int i = lane; i < kTileBytes / 4; i += 32) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tile + i * 4));
      dump[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* map, float* x, int64_t rows, bool wide) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr) return false;
  cuuint64_t dims[2] = {kDim, static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {kDim * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(wide ? kDim : 32), 1};   // gather4: 1 in the gathered dimension
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiled>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 wide ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r));
  return r == CUDA_SUCCESS;
}

__global__ void fill_kernel(float* x, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] = static_cast<float>((i * 2654435761ull) & 0xFFFF) * (1.0f / 65536.0f);
}

struct Result { double ms, gbs, bpc; };

template <typename Launch>
static Result time_it(Launch launch, unsigned long long* d_cyc, double bytes, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();   // warm-up (also brings the working set into L2 where it fits)
  CK(cudaDeviceSynchronize());
  double best_ms = 1e30;
  unsigned long long best_cyc = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemset(d_cyc, 0, sizeof(unsigned long long)));
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    if (ms < best_ms) { best_ms = ms; best_cyc = cyc; }
  }
  CK(cudaGetLastError());
  return {best_ms, bytes / (best_ms * 1e-3) / 1e9, best_cyc ? bytes / static_cast<double>(best_cyc) : 0.0};
}

template <int DEPTH>
static void run_depth(const char* mode, int warps, float* x, uint32_t idx, const CUtensorMap* map, int tiles_per_warp,
                      int grid, unsigned long long* d_cyc, double ws_mb, float* dump) {
  const size_t smem = static_cast<size_t>(warps) * DEPTH * kTileBytes + warps * DEPTH * 8 + 1024;
  if (smem > 232448) return;
  const double bytes = static_cast<double>(grid) * warps * tiles_per_warp * kTileBytes;
  Result r{};
  if (!strcmp(mode, "ldgsts")) {
    CK(cudaFuncSetAttribute(gather_ldgsts<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    r = time_it([&] { gather_ldgsts<DEPTH><<<grid, warps * 32, smem>>>(x, idx, tiles_per_warp, d_cyc, dump); }, d_cyc, bytes, 5);
  } else if (!strcmp(mode, "gather4")) {
    CK(cudaFuncSetAttribute(gather_tma<DEPTH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    r = time_it([&] { gather_tma<DEPTH, false><<<grid, warps * 32, smem>>>(*map, idx, tiles_per_warp, d_cyc, dump); }, d_cyc, bytes, 5);
  } else {
    CK(cudaFuncSetAttribute(gather_tma<DEPTH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    r = time_it([&] { gather_tma<DEPTH, true><<<grid, warps * 32, smem>>>(*map, idx, tiles_per_warp, d_cyc, dump); }, d_cyc, bytes, 5);
  }
  printf("{\"mode\": \"%s\", \"warps\": %d, \"depth\": %d, \"working_set_mb\": %.1f, \"tiles\": %.0f, \"ms\": %.4f, "
         "\"gbs\": %.1f, \"bytes_per_clk\": %.1f}\n",
         mode, warps, DEPTH, ws_mb, bytes / kTileBytes, r.ms, r.gbs, r.bpc);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int grid = prop.multiProcessorCount;
  const bool quick = argc > 1 && !strcmp(argv[1], "--quick");
  fprintf(stderr, "device: %s, %d SMs, L2 %d MB\n", prop.name, grid, prop.l2CacheSize >> 20);
  const int64_t max_rows = (512ll << 20) / (kDim * 4);
  float* x = nullptr;
  unsigned long long* d_cyc = nullptr;
  float* dump = nullptr;
  const int tiles_per_warp = quick ? 512 : 2048;
  CK(cudaMalloc(&x, max_rows * kDim * 4));
  CK(cudaMalloc(&d_cyc, 8));
  CK(cudaMalloc(&dump, 3 * kTileBytes));
  // ---- image check: the three paths gather the same 8 rows; ldgsts and gather4 must produce identical bytes
  {
    const int64_t rows = 100000;
    const uint32_t idx = static_cast<uint32_t>(rows);
    fill_kernel<<<grid * 8, 256>>>(x, rows * kDim);
    CK(cudaDeviceSynchronize());
    CUtensorMap map_s, map_w;
    const bool ok_s = make_map(&map_s, x, rows, false), ok_w = make_map(&map_w, x, rows, true);
    const size_t smem = 2 * kTileBytes + 64 + 1024;
    CK(cudaFuncSetAttribute(gather_ldgsts<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    gather_ldgsts<2><<<1, 32, smem>>>(x, idx, 3, d_cyc, dump);
    if (ok_s) {
      CK(cudaFuncSetAttribute(gather_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      gather_tma<2, false><<<1, 32, smem>>>(map_s, idx, 3, d_cyc, dump + kTileBytes / 4);
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> h(2 * kTileBytes / 4);
    CK(cudaMemcpy(h.data(), dump, 2 * kTileBytes, cudaMemcpyDeviceToHost));
    const bool same = ok_s && memcmp(h.data(), h.data() + kTileBytes / 4, kTileBytes) == 0;
    printf("{\"check\": \"gather4 (SWIZZLE_128B_ATOM_32B, box 32x1) image == ldgsts SWIZZLE_128B_BASE32B image\", "
           "\"tensor_map_ok\": %s, \"identical\": %s}\n", ok_s ? "true" : "false", same ? "true" : "false");
    fflush(stdout);
    if (!ok_w) fprintf(stderr, "wide tensor map unavailable\n");
  }
  const double ws_list_full[] = {24, 48, 96, 119.3, 192, 512};
  const double ws_list_quick[] = {48, 119.3};
  const double* ws_list = quick ? ws_list_quick : ws_list_full;
  const int n_ws = quick ? 2 : 6;
  for (int wi = 0; wi < n_ws; ++wi) {
    const int64_t rows = static_cast<int64_t>(ws_list[wi] * 1e6 / (kDim * 4));
    const uint32_t idx = static_cast<uint32_t>(rows);
    fill_kernel<<<grid * 8, 256>>>(x, rows * kDim);
    CK(cudaDeviceSynchronize());
    CUtensorMap map_s, map_w;
    const bool ok_s = make_map(&map_s, x, rows, false), ok_w = make_map(&map_w, x, rows, true);
    const double ws = rows * kDim * 4 / 1e6;
    for (int warps : {4, 6, 8, 12, 16}) {
      run_depth<2>("ldgsts", warps, x, idx, nullptr, tiles_per_warp, grid, d_cyc, ws, nullptr);
      run_depth<3>("ldgsts", warps, x, idx, nullptr, tiles_per_warp, grid, d_cyc, ws, nullptr);
      run_depth<4>("ldgsts", warps, x, idx, nullptr, tiles_per_warp, grid, d_cyc, ws, nullptr);
      run_depth<6>("ldgsts", warps, x, idx, nullptr, tiles_per_warp, grid, d_cyc, ws, nullptr);
      run_depth<8>("ldgsts", warps, x, idx, nullptr, tiles_per_warp, grid, d_cyc, ws, nullptr);
      if (ok_s) {
        run_depth<3>("gather4", warps, x, idx, &map_s, tiles_per_warp, grid, d_cyc, ws, nullptr);
        run_depth<6>("gather4", warps, x, idx, &map_s, tiles_per_warp, grid, d_cyc, ws, nullptr);
        run_depth<8>("gather4", warps, x, idx, &map_s, tiles_per_warp, grid, d_cyc, ws, nullptr);
      }
      if (ok_w) {
        run_depth<3>("gather4_wide", warps, x, idx, &map_w, tiles_per_warp, grid, d_cyc, ws, nullptr);
        run_depth<6>("gather4_wide", warps, x, idx, &map_w, tiles_per_warp, grid, d_cyc, ws, nullptr);
      }
    }
  }
  return 0;
}
