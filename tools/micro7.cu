// Micro-benchmark, round 7: LDS cost when lanes of different row slots read the SAME shared-memory words (broadcast).
// Question: the array delivers 128 distinct bytes per clock, the return path 256 B per clock (micro5: LDS.128 uniform = 2 cycles).
// Do two quarter-warps reading the same 128 B (or two half-warps reading the same 256 B) halve the cost of a gather?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// per-lane byte offset of the 16 B (or 8 B) the lane reads, as a function of the pattern
__device__ __forceinline__ uint32_t lane_off(int mode, int lane) {
  const int q = lane >> 3, l8 = lane & 7, h = lane >> 4, l16 = lane & 15;
  switch (mode) {
    case 0: return q * 1024 + l8 * 16;             // LDS.128: 4 quarters, 4 distinct lines (128 B each)
    case 1: return (q >> 1) * 1024 + l8 * 16;      // LDS.128: quarter pairs share a line (2 distinct 128 B)
    case 2: return l8 * 16;                        // LDS.128: all quarters the same 128 B
    case 3: return h * 1024 + l16 * 16;            // LDS.128: 2 halves, 256 B each, distinct lines
    case 4: return l16 * 16;                       // LDS.128: both halves the same 256 B
    case 5: return q * 1024 + l8 * 16 + (q & 1) * 128;  // LDS.128: 4 distinct lines, alternating 128-B halves (bank spread)
    case 6: return (q >> 1) * 1024 + l8 * 16 + (q & 1) * 128;  // LDS.128: pairs share a LINE but read different halves of it
    case 7: return h * 1024 + l16 * 8;             // LDS.64 : 2 halves, 128 B each, distinct lines
    case 8: return l16 * 8;                        // LDS.64 : both halves the same 128 B
    case 10: return q * 16;                        // LDS.128 quarter-uniform: 4 distinct 16-byte words (the forward entry read)
    case 11: return (lane >> 2) * 16;              // LDS.128: 8 distinct 16-byte words
    case 12: return h * 16;                        // LDS.128 half-uniform: 2 distinct words (the backward entry read)
    case 13: return 0;                             // LDS.128 uniform
    default: return lane * 16;                     // LDS.128 contiguous 512 B
  }
}
template <int W64>
__global__ void __launch_bounds__(512, 2) k(int mode, int iters, int* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  int* si = (int*)smraw;
  for (int i = threadIdx.x; i < 20480; i += blockDim.x) si[i] = i * 2654435761u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = smem_u32(smraw) + (warp & 7) * 256 + lane_off(mode, lane);
  int acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint32_t o = base + ((it & 3) << 12);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (W64) {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(o + u * 2048));
        acc ^= v.x ^ v.y;
      } else {
        int4 v;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(o + u * 2048));
        acc ^= v.x ^ v.y; acc ^= v.z ^ v.w;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  int* out;
  CK(cudaMalloc(&out, 296 * 512 * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[] = {"LDS.128 4 quarters x 4 distinct lines", "LDS.128 quarter pairs share a line (2 x 128 B)", "LDS.128 all quarters same 128 B",
                         "LDS.128 2 halves x 256 B distinct", "LDS.128 both halves same 256 B", "LDS.128 4 lines, alternating halves",
                         "LDS.128 pairs share a line, different halves", "LDS.64 2 halves x 128 B distinct", "LDS.64 both halves same 128 B", "LDS.128 contiguous 512 B",
                         "LDS.128 quarter-uniform (4 words)", "LDS.128 eighth-uniform (8 words)", "LDS.128 half-uniform (2 words)", "LDS.128 uniform"};
  CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920));
  CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920));
  for (int mode = 0; mode < 14; ++mode) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 7 || mode == 8) k<1><<<296, 512, 81920>>>(mode, iters, out); else k<0><<<296, 512, 81920>>>(mode, iters, out);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("%-50s %.3f ms -> %.3f cycles per warp-instr per SM (@1.95 GHz)\n", names[mode], ms, ms * 1e-3 * 1.95e9 / (32.0 * iters * 8));
  }
  return 0;
}
