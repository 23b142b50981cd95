// Micro-benchmarks of the sm_100a primitives the residual kernels are built from (developer tool,
// not part of the product).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro micro.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

// mode: 0 contiguous 512B/warp, 1 all lanes same 16B, 2 halves each same 16B (adjacent), 3 both halves the
// same 256B line, 4 halves different lines, 5 LDS.64 contiguous 256B, 6 quarter-uniform (4 distinct 16B)
__global__ void k_lds(int mode, long long* cyc, float* sink) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = smem_u32(sm) + warp * 2048;
  uint32_t a;
  switch (mode) {
    case 0: a = base + lane * 16; break;
    case 1: a = base; break;
    case 2: a = base + (lane >> 4) * 16; break;
    case 3: a = base + (lane & 15) * 16; break;
    case 4: a = base + (lane >> 4) * 1024 + (lane & 15) * 16; break;
    case 5: a = base + lane * 8; break;
    default: a = base + (lane >> 3) * 16; break;
  }
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (mode == 5) { float2 v = lds64(a ^ (u * 256)); acc += v.x; }
      else { float4 v = lds128(a ^ (u * 512)); acc += v.x; }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

// FFMA throughput: mode 0 scalar FFMA, 1 packed fma.rn.f32x2
__global__ void k_fma(int mode, long long* cyc, float* sink) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float c0 = 1.0001f + threadIdx.x * 1e-9f, c1 = 0.9999f;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) a[u] = fmaf(a[u], c0, c1);
    }
  } else {
    unsigned long long p[8], cc0, cc1;
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc0) : "f"(c0));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc1) : "f"(c1));
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[u]) : "l"(cc0), "l"(cc1));
    }
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}

// forward-like step: 1 uniform-per-half LDS.128 (entry) + 1 LDS.128 gather + 12 FFMA.  mode 1: packed FFMA2 (+3 movs)
__global__ void k_step(int mode, long long* cyc, float* sink) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 24576; i += blockDim.x) sm[i] = (float)(i & 255) * 0.001f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
  const uint32_t sbase = smem_u32(sm);
  uint32_t ent = sbase + 65536 + warp * 1024 + half * 16;  // entries: 32B per step (2 halves)
  float4 accA = {0, 0, 0, 0}, acc1 = {0, 0, 0, 0}, acc2 = {0, 0, 0, 0};
  unsigned long long pA0 = 0, pA1 = 0, p10 = 0, p11 = 0, p20 = 0, p21 = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 e = lds128(ent + u * 32);
      uint32_t off = (__float_as_uint(e.x) & 0xff00u) + ((it * 4 + u) & 63) * 256;  // some line
      float4 x = lds128(sbase + (off & 0xffff) + (lane & 15) * 16);
      if (mode == 0) {
        accA.x = fmaf(e.y, x.x, accA.x); accA.y = fmaf(e.y, x.y, accA.y); accA.z = fmaf(e.y, x.z, accA.z); accA.w = fmaf(e.y, x.w, accA.w);
        acc1.x = fmaf(e.z, x.x, acc1.x); acc1.y = fmaf(e.z, x.y, acc1.y); acc1.z = fmaf(e.z, x.z, acc1.z); acc1.w = fmaf(e.z, x.w, acc1.w);
        acc2.x = fmaf(e.w, x.x, acc2.x); acc2.y = fmaf(e.w, x.y, acc2.y); acc2.z = fmaf(e.w, x.z, acc2.z); acc2.w = fmaf(e.w, x.w, acc2.w);
      } else {
        unsigned long long x0, x1, ca, c1, c2;
        asm("mov.b64 %0, {%1,%2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
        asm("mov.b64 %0, {%1,%2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
        asm("mov.b64 %0, {%1,%1};" : "=l"(ca) : "f"(e.y));
        asm("mov.b64 %0, {%1,%1};" : "=l"(c1) : "f"(e.z));
        asm("mov.b64 %0, {%1,%1};" : "=l"(c2) : "f"(e.w));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pA0) : "l"(ca), "l"(x0));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pA1) : "l"(ca), "l"(x1));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p10) : "l"(c1), "l"(x0));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p11) : "l"(c1), "l"(x1));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p20) : "l"(c2), "l"(x0));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p21) : "l"(c2), "l"(x1));
      }
    }
  }
  long long t1 = clock64();
  float s = accA.x + accA.y + accA.z + accA.w + acc1.x + acc1.y + acc1.z + acc1.w + acc2.x + acc2.y + acc2.z + acc2.w;
  s += (float)(pA0 ^ pA1 ^ p10 ^ p11 ^ p20 ^ p21);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}

// L2/DRAM -> shared staging throughput.  Each CTA stages `lines` pieces of `piece` bytes (row stride 4096 B,
// as in a dof-major [N][1024] float array) per round.  mode 0: cp.async 16B (LDGSTS), mode 1: cp.async.bulk.
__global__ void k_stage(int mode, const float* __restrict__ src, long long nrows, int lines, int piece, int rounds, float* sink) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t sb = smem_u32(smraw), bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int slabs = 4096 / piece;
  uint32_t phase = 0;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    // tile id: CTAs of consecutive blockIdx share rows (slab fastest)
    long long tile = ((long long)(blockIdx.x / slabs) + (long long)r * (gridDim.x / slabs));
    long long row0 = (tile * (lines * 3 / 4)) % (nrows - lines);  // 25% overlap between neighbouring tiles
    int slab = blockIdx.x % slabs;
    const char* g = (const char*)src + row0 * 4096 + (long long)slab * piece;
    if (mode == 0) {
      const int chunks = piece / 16;
      for (int i = threadIdx.x; i < lines * chunks; i += blockDim.x) {
        int l = i / chunks, c = i - l * chunks;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + l * piece + c * 16), "l"(g + (long long)l * 4096 + c * 16));
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;");
      __syncthreads();
    } else {
      if (threadIdx.x < 32) {
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(lines * piece));
        __syncwarp();
        for (int l = threadIdx.x; l < lines; l += 32)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + l * piece),
                       "l"(g + (long long)l * 4096), "r"(piece), "r"(bar_a)
                       : "memory");
      }
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a), "r"(phase));
      phase ^= 1;
    }
    acc += ((float*)smraw)[threadIdx.x];
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 4096 * sizeof(long long)));
  CK(cudaMalloc(&sink, 64));
  std::vector<long long> h(4096);
  auto report = [&](const char* name, int ctas, double ops_per_cta) {
    cudaMemcpy(h.data(), cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < ctas; ++i) s += (double)h[i];
    s /= ctas;
    printf("%-44s cycles/CTA %.0f  -> %.3f cycles per warp-instr per SM\n", name, s, s / ops_per_cta);
  };
  CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  const char* lds_names[] = {"LDS.128 contiguous 512B", "LDS.128 all lanes same 16B", "LDS.128 halves same 16B (adjacent)",
                             "LDS.128 both halves same 256B line", "LDS.128 halves different lines", "LDS.64 contiguous 256B",
                             "LDS.128 quarters uniform (4x16B)"};
  for (int warps : {8, 16}) {
    for (int m = 0; m < 7; ++m) {
      k_lds<<<148, warps * 32, 65536>>>(m, cyc, sink);
      CK(cudaDeviceSynchronize());
      char nm[128];
      snprintf(nm, sizeof nm, "%s w=%d", lds_names[m], warps);
      report(nm, 148, (double)ITERS * 8 * warps);
    }
  }
  for (int warps : {4, 8, 16}) {
    for (int m = 0; m < 2; ++m) {
      k_fma<<<148, warps * 32>>>(m, cyc, sink);
      CK(cudaDeviceSynchronize());
      char nm[128];
      snprintf(nm, sizeof nm, "%s w=%d (per 16 FMA/lane)", m ? "FFMA2" : "FFMA", warps);
      report(nm, 148, (double)ITERS * warps);
    }
  }
  for (int warps : {8, 16}) {
    for (int m = 0; m < 2; ++m) {
      k_step<<<148, warps * 32, 98304>>>(m, cyc, sink);
      CK(cudaDeviceSynchronize());
      char nm[128];
      snprintf(nm, sizeof nm, "fwd step %s w=%d (cycles/step/SM)", m ? "FFMA2" : "FFMA", warps);
      report(nm, 148, (double)ITERS * 4 * warps);
    }
  }
  // staging
  const long long nrows = 1 << 20;  // 4 GiB
  float* src;
  CK(cudaMalloc(&src, nrows * 4096));
  CK(cudaMemset(src, 0, nrows * 4096));
  CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int piece : {256, 512}) {
      int lines = 100 * 1024 / piece;
      int rounds = 64;
      int grid = 148 * 2 * 16;  // multiple of slabs
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_stage<<<grid, 256, lines * piece>>>(mode, src, nrows, lines, piece, rounds, sink);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)grid * rounds * lines * piece;
        if (rep) printf("stage mode=%s piece=%d lines=%d: %.2f ms, %.1f GB/s into smem (2 CTA/SM)\n", mode ? "bulk" : "ldgsts", piece, lines, ms, bytes / ms * 1e-6);
      }
    }
  return 0;
}
