// Micro-benchmark, round 9: cost of one tcgen05.mma with both operands in shared memory (SS), as the preconditioner
// GEMM issues it: M = 128, K = 32 bytes, canonical K-major SWIZZLE_NONE core-matrix layout (LBO = rows * 16, SBO = 128).
// One CTA per SM; one elected lane issues R accumulating MMAs back to back, commits, and the warp waits for the commit:
// cycles / R = sustained cost per MMA.  Cases: kind::tf32 N = 64 / 128 / 256, kind::f16 (bf16) N = 64 / 128 / 256.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
template <int KIND>  // 0 = tf32, 1 = f16 (bf16 inputs)
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__global__ void __launch_bounds__(128, 1) k(int N, int R, int distinct, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (__shfl_sync(0xffffffffu, warp, 0) == 1) {
    const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 1u) << 7) | ((KIND == 0 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    // A: 128 rows x 32 B -> 2 chunks of 128 * 16 B, LBO = 2048; B: N rows x 32 B, LBO = N * 16; `distinct` operand sets
    const uint32_t a_lo = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo = (((smem_u32(smem) + 32768u) & 0x3ffffu) >> 4) | (((uint32_t)N * 16u >> 4) << 16);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < R; ++r) {
        const uint32_t o = (uint32_t)(r % distinct);
        umma<KIND>(tmem, ((uint64_t)hi << 32) | (a_lo + o * (4096u >> 4)), ((uint64_t)hi << 32) | (b_lo + o * (8192u >> 4)), idesc, r != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0, out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}
int main() {
  long long* d;
  CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int R = 4096;
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {64, 128, 256})
      for (int distinct : {1, 4}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (kind == 0) k<0><<<148, 128, 96 * 1024>>>(N, R, distinct, d);
          else k<1><<<148, 128, 96 * 1024>>>(N, R, distinct, d);
          CK(cudaDeviceSynchronize());
        }
        long long h[2];
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        const double flop = 2.0 * 128 * N * (kind == 0 ? 8 : 16);
        printf("%s M=128 N=%3d K=32B SS no-swizzle, %d operand set(s): issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f flop/clk/SM = %.0f TFLOP/s at 148 SMs x 1.965 GHz\n",
               kind == 0 ? "tf32" : "bf16", N, distinct, (double)h[0] / R, (double)h[1] / R, flop / ((double)h[1] / R), flop / ((double)h[1] / R) * 148 * 1.965e9 / 1e12);
      }
  return 0;
}
