// Micro-benchmark, round 8: DRAM -> shared-memory staging throughput of one persistent CTA per SM (two 96 KB stages),
// for the ways a tile's dof lines (256 B = 64 samples each) can be fetched:
//   mode 0: 2-D TMA boxes, R rows x 256 B, row stride = ldb * 4 (dof-major array [N][ldb], today's layout)
//   mode 1: 1-D bulk copies of R * 256 B contiguous bytes (slab-major array [slab][N][64])
//   mode 2: cp.async 16 B per thread from the dof-major array (LDGSTS)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
struct Maps { CUtensorMap m[3]; };
constexpr int kLines = 384;
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Maps maps, const float* base, long long ldb, int n_tiles, int n_slabs, int mode, int rows, int halo, long long n_lines) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2];
  const uint32_t sb = smem_u32(smem), bar = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int n_units = n_tiles * n_slabs;
  // every unit is released as soon as it has landed: the producer is never blocked by consumers (pure fetch rate), two stages in flight
  if (mode != 2) {
    if (warp != 0) return;
    int i = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++i) {
      const uint32_t s = i & 1;
      if (i >= 2) mbar_wait(bar + s * 8, ((i >> 1) - 1) & 1);
      const int tile = u / n_slabs, slab = u - tile * n_slabs;
      if (lane == 0) mbar_expect_tx(bar + s * 8, kLines * 256);
      __syncwarp();
      const int dof0 = tile * (kLines - halo);  // tiles overlap by `halo` lines (re-reads served by L2)
      for (int b = lane; b < kLines / rows; b += 32) {
        const uint32_t dst = sb + s * (kLines * 256) + b * rows * 256;
        if (mode == 0) {
          const CUtensorMap* mp = &maps.m[rows == 16 ? 0 : rows == 4 ? 1 : 2];
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(mp), "r"(slab * 64), "r"(dof0 + b * rows), "r"(bar + s * 8) : "memory");
        } else {
          const float* src = base + ((size_t)slab * (size_t)n_lines + dof0 + (size_t)b * rows) * 64;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(rows * 256), "r"(bar + s * 8) : "memory");
        }
      }
    }
    // drain
    const int total = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int j = max(0, total - 2); j < total; ++j) mbar_wait(bar + (j & 1) * 8, (j >> 1) & 1);
  } else {
    int i = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++i) {
      const int tile = u / n_slabs, slab = u - tile * n_slabs;
      const int dof0 = tile * (kLines - halo);
      const uint32_t s = i & 1;
      for (int c = threadIdx.x; c < kLines * 16; c += blockDim.x) {
        const int line = c >> 4, part = c & 15;
        const float* src = base + (size_t)(dof0 + line) * ldb + slab * 64 + part * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + s * (kLines * 256) + line * 256 + part * 16), "l"(src));
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 1;");
      __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;");
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int n_slabs = 16, n_tiles = 2600;
  const long long ldb = 1024, N = (long long)n_tiles * kLines;
  float* a;
  CK(cudaMalloc(&a, (size_t)N * ldb * 4));
  CK(cudaMemset(a, 0, (size_t)N * ldb * 4));
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  Maps maps;
  const int R[3] = {16, 4, 1};
  for (int c = 0; c < 3; ++c) {
    const cuuint64_t dims[2] = {(cuuint64_t)ldb, (cuuint64_t)N}; const cuuint64_t strides[1] = {(cuuint64_t)ldb * 4};
    const cuuint32_t box[2] = {64, (cuuint32_t)R[c]}; const cuuint32_t es[2] = {1, 1};
    if (enc(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  }
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kLines * 256));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* mn[] = {"2-D TMA box (dof-major, 4 KB row stride)", "1-D bulk copy (slab-major, contiguous)", "cp.async 16 B (dof-major)"};
  for (int halo = 0; halo <= 192; halo += 192)
    for (int mode = 0; mode < 3; ++mode)
      for (int rows : {16, 4, 1}) {
        if (mode == 2 && rows != 16) continue;
        const int tiles = halo ? (int)((N - kLines) / (kLines - halo)) : n_tiles;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          k<<<148, 512, 2 * kLines * 256>>>(maps, a, ldb, tiles, n_slabs, mode, rows, halo, N);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          cudaEventElapsedTime(&ms, e0, e1);
        }
        const double bytes = (double)tiles * n_slabs * kLines * 256;
        printf("halo %3d  %-44s rows/piece %2d: %.3f ms, %.0f GB/s staged (%.1f B/clk/SM @1.95 GHz)\n", halo, mn[mode], rows, ms, bytes / ms * 1e-6, bytes / (ms * 1e-3) / 148 / 1.95e9);
      }
  return 0;
}
