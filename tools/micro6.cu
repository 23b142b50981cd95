// Micro-benchmark: per-SM throughput of the async copy engine (cp.async.bulk, L2-resident source) vs cp.async (LDGSTS).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// mode 0: cp.async.bulk pieces of `piece` bytes; mode 1: cp.async.cg 16 B per thread.  Each CTA moves `total` bytes per round.
__global__ void k(int mode, const char* __restrict__ src, size_t src_bytes, int piece, int total, int rounds, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t sb = smem_u32(sm), ba = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  uint32_t phase = 0;
  const size_t base = ((size_t)blockIdx.x * 131072) % (src_bytes - 262144);
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const char* g = src + base + (size_t)(r & 1) * 65536;
    if (mode == 0) {
      const int n = total / piece;
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(total));
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + i * piece), "l"(g + (size_t)i * piece), "r"(piece), "r"(ba) : "memory");
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(ba), "r"(phase));
      phase ^= 1;
    } else {
      for (int i = threadIdx.x; i < total / 16; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + i * 16), "l"(g + (size_t)i * 16));
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;");
      __syncthreads();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  const size_t bytes = 64u << 20;  // L2 resident
  char* src; long long* cyc;
  CK(cudaMalloc(&src, bytes)); CK(cudaMemset(src, 1, bytes)); CK(cudaMalloc(&cyc, 1024 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long h[1024];
  const int total = 96 * 1024, rounds = 200;
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
    for (int mode = 0; mode < 2; ++mode)
      for (int piece : {256, 1024, 4096, 16384}) {
        if (mode == 1 && piece != 256) continue;
        for (int rep = 0; rep < 2; ++rep) { k<<<148 * ctas_per_sm, 256, total>>>(mode, src, bytes, piece, total, rounds, cyc); CK(cudaDeviceSynchronize()); }
        cudaMemcpy(h, cyc, 148 * ctas_per_sm * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148 * ctas_per_sm; ++i) s += (double)h[i];
        s /= 148 * ctas_per_sm;
        printf("%s piece=%5d B, %d CTA/SM: %.0f cycles per 96 KB round per CTA -> %.1f B/clk per SM\n", mode ? "cp.async 16B (LDGSTS)" : "cp.async.bulk        ", piece, ctas_per_sm, s / rounds,
               (double)total * ctas_per_sm / (s / rounds));
      }
  return 0;
}
