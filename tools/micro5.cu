// Micro-benchmarks, round 5: cost table of the load-store pipe (cycles per warp instruction per SM, 32 warps/SM, event timed).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__constant__ int4 ctab[4096];

// MODE: 0 LDS.128 contiguous, 1 LDS.128 half-uniform, 2 LDS.128 uniform, 3 LDS.64 contiguous, 4 LDS.64 uniform,
//       5 LDS.32 contiguous, 6 LDS.32 uniform, 7 LDG.128 uniform (L1), 8 LDG.128 half-uniform (L1), 9 LDG.32 uniform (L1),
//       10 LDG.128 contiguous (L1), 11 LDC.128 uniform index, 12 LDC.128 half-uniform index, 13 LDG.64 uniform
template <int MODE>
__global__ void __launch_bounds__(512, 2) k_ls(const int4* __restrict__ g, int iters, int* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  int* si = (int*)smraw;
  for (int i = threadIdx.x; i < 20480; i += blockDim.x) si[i] = i * 2654435761u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
  const uint32_t sb = smem_u32(smraw) + warp * 4096;
  const char* gb = (const char*)g + warp * 4096;
  uint32_t lo;
  switch (MODE) {
    case 0: case 10: lo = lane * 16; break;
    case 1: case 8: case 12: lo = half * 16; break;
    case 3: lo = lane * 8; break;
    case 5: lo = lane * 4; break;
    default: lo = 0; break;
  }
  int acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint32_t o = lo + ((it & 3) << 10);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE <= 2) {
        int4 v;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sb + o + u * 32 * (MODE == 0 ? 16 : 1)));
        acc ^= v.x ^ v.y; acc ^= v.z ^ v.w;
      } else if (MODE <= 4) {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sb + o + u * (MODE == 3 ? 256 : 32)));
        acc ^= v.x ^ v.y;
      } else if (MODE <= 6) {
        int v;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(sb + o + u * (MODE == 5 ? 128 : 32)));
        acc ^= v;
      } else if (MODE == 7 || MODE == 8 || MODE == 10) {
        int4 v;
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(gb + o + u * (MODE == 10 ? 512 : 32)));
        acc ^= v.x ^ v.y; acc ^= v.z ^ v.w;
      } else if (MODE == 9) {
        int v;
        asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(gb + o + u * 32));
        acc ^= v;
      } else if (MODE == 13) {
        int2 v;
        asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(gb + o + u * 32));
        acc ^= v.x ^ v.y;
      } else if (MODE == 14) {
        int4 v;
        if (u & 1) asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(gb + lane * 16 + (o & 0xc00) + (u >> 1) * 512));
        else asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sb + lane * 16 + (o & 0xc00) + (u >> 1) * 512));
        acc ^= v.x ^ v.y; acc ^= v.z ^ v.w;
      } else {
        const int4 v = ctab[((o >> 4) + u * 2 + warp * 8) & 4095];
        acc ^= v.x ^ v.y; acc ^= v.z ^ v.w;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int* out;
  int4* g;
  CK(cudaMalloc(&out, 296 * 512 * 4));
  CK(cudaMalloc(&g, 1 << 20));
  CK(cudaMemset(g, 1, 1 << 20));
  std::vector<int4> hc(4096, make_int4(1, 2, 3, 4));
  CK(cudaMemcpyToSymbol(ctab, hc.data(), sizeof(int4) * 4096));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[] = {"LDS.128 contiguous 512B", "LDS.128 half-uniform", "LDS.128 uniform", "LDS.64 contiguous 256B", "LDS.64 uniform",
                         "LDS.32 contiguous 128B", "LDS.32 uniform", "LDG.128 uniform (L1 hit)", "LDG.128 half-uniform (L1 hit)", "LDG.32 uniform (L1 hit)",
                         "LDG.128 contiguous 512B (L1 hit)", "LDC.128 uniform index", "LDC.128 half-uniform index", "LDG.64 uniform (L1 hit)", "mix LDS.128 + LDG.128 contiguous (per instr)"};
#define RUN(MODE)                                                                                         \
  do {                                                                                                   \
    CK(cudaFuncSetAttribute(k_ls<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920));             \
    float ms = 0;                                                                                        \
    for (int rep = 0; rep < 2; ++rep) {                                                                  \
      cudaEventRecord(e0);                                                                               \
      k_ls<MODE><<<296, 512, 81920>>>(g, iters, out);                                                     \
      cudaEventRecord(e1);                                                                               \
      CK(cudaDeviceSynchronize());                                                                       \
      cudaEventElapsedTime(&ms, e0, e1);                                                                 \
    }                                                                                                    \
    double ops_per_sm = 32.0 * iters * 8;                                                                \
    printf("%-36s %.3f ms -> %.3f cycles per warp-instr per SM (@1.95 GHz)\n", names[MODE], ms, ms * 1e-3 * 1.95e9 / ops_per_sm); \
  } while (0)
  RUN(0); RUN(10); RUN(13); RUN(14);
  return 0;
}
