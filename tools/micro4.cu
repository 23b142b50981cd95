// Micro-benchmarks, round 4: how to feed the per-warp entry stream (L2-resident, shared by the 16 slab CTAs of a tile).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }

struct Acc {
  u64 A0, A1, B0, B1, C0, C1;
  __device__ __forceinline__ void zero() { A0 = A1 = B0 = B1 = C0 = C1 = 0; }
  __device__ __forceinline__ void step(const int4& e, const float4& x) {
    const float a = __int_as_float(e.y), b1 = __int_as_float(e.z), b2 = __int_as_float(e.w);
    const u64 x0 = pk(x.x, x.y), x1 = pk(x.z, x.w), ca = pk(a, a), c1 = pk(b1, b1), c2 = pk(b2, b2);
    fma2(A0, ca, x0); fma2(A1, ca, x1); fma2(B0, c1, x0); fma2(B1, c1, x1); fma2(C0, c2, x0); fma2(C1, c2, x1);
  }
  __device__ __forceinline__ float sum() { return lo(A0) + hi(A0) + lo(A1) + hi(A1) + lo(B0) * hi(B0) + lo(B1) * hi(B1) + lo(C0) * hi(C0) + lo(C1) * hi(C1); }
};

constexpr int STEPS = 24;  // steps per pair (each step: 2 halves x 16 B)
// MODE 0: LDG per batch, no look-ahead.  MODE 1: register double buffer (look-ahead one batch, across pairs).
// MODE 2: per-warp shared-memory ring filled with cp.async, 16 steps (512 B) per chunk, NCH chunks in flight.
template <int MODE, int BATCH, int SHARED_L1>
__global__ void __launch_bounds__(512, 2) k_fwd(const int4* __restrict__ gstream, int pairs_per_warp, int lines, long long* cyc, float* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const uint32_t sb = smem_u32(smraw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, nw = blockDim.x >> 5;
  float* lf = (float*)smraw;
  for (int i = threadIdx.x; i < lines * 64; i += blockDim.x) lf[i] = (float)(i & 1023) * 1e-3f;
  __syncthreads();
  const int tile = SHARED_L1 ? 0 : (blockIdx.x >> 4);  // 16 slab CTAs share a tile's stream
  const int total_steps = pairs_per_warp * STEPS;
  const int4* ws = gstream + (size_t)(tile * nw + warp) * total_steps * 2;
  const uint32_t lane_off = (lane & 15) * 16;
  float total = 0.f;
  Acc acc;
  acc.zero();
  unsigned long long g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  long long t0 = clock64();
  if (MODE == 0) {
#pragma unroll 1
    for (int s = 0; s < total_steps; s += BATCH) {
      int4 e[BATCH];
      float4 x[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) e[u] = __ldg(ws + (s + u) * 2 + half);
#pragma unroll
      for (int u = 0; u < BATCH; ++u) x[u] = lds128(sb + (uint32_t)e[u].x + lane_off);
#pragma unroll
      for (int u = 0; u < BATCH; ++u) acc.step(e[u], x[u]);
      if ((s + BATCH) % STEPS == 0) { total += acc.sum(); acc.zero(); }
    }
  } else if (MODE == 1) {
    int4 e[BATCH], en[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) e[u] = __ldg(ws + u * 2 + half);
#pragma unroll 1
    for (int s = 0; s < total_steps; s += BATCH) {
      const int sn = (s + BATCH < total_steps) ? s + BATCH : 0;
#pragma unroll
      for (int u = 0; u < BATCH; ++u) en[u] = __ldg(ws + (sn + u) * 2 + half);
      float4 x[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) x[u] = lds128(sb + (uint32_t)e[u].x + lane_off);
#pragma unroll
      for (int u = 0; u < BATCH; ++u) acc.step(e[u], x[u]);
#pragma unroll
      for (int u = 0; u < BATCH; ++u) e[u] = en[u];
      if ((s + BATCH) % STEPS == 0) { total += acc.sum(); acc.zero(); }
    }
  } else {
    constexpr int NCH = 4;                 // ring slots of 512 B (16 steps)
    const uint32_t ring = sb + lines * 256 + warp * (NCH * 512);
    const int n_chunks = total_steps / 16;
    const char* src = (const char*)ws;
    // prologue: NCH-1 chunks in flight
#pragma unroll
    for (int c = 0; c < NCH - 1; ++c) {
      if (c < n_chunks) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + c * 512 + lane * 16), "l"(src + (size_t)c * 512 + lane * 16));
      asm volatile("cp.async.commit_group;");
    }
    int steps_in_pair = 0;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      {
        const int cn = c + NCH - 1;
        if (cn < n_chunks) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (cn % NCH) * 512 + lane * 16), "l"(src + (size_t)cn * 512 + lane * 16));
        asm volatile("cp.async.commit_group;");
      }
      asm volatile("cp.async.wait_group %0;" ::"n"(NCH - 1));
      __syncwarp();
      const uint32_t slot = ring + (c % NCH) * 512 + half * 16;
#pragma unroll
      for (int b = 0; b < 16; b += BATCH) {
        int4 e[BATCH];
        float4 x[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) e[u] = lds128i(slot + (b + u) * 32);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) x[u] = lds128(sb + (uint32_t)e[u].x + lane_off);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) acc.step(e[u], x[u]);
      }
      steps_in_pair += 16;
      if (steps_in_pair >= STEPS) { total += acc.sum(); acc.zero(); steps_in_pair -= STEPS; }
      __syncwarp();
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    cyc[blockIdx.x * 2] = t1 - t0;
    cyc[blockIdx.x * 2 + 1] = (long long)(g1 - g0);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = total + acc.sum();
}

int main() {
  long long* cyc;
  float* out;
  const int lines = 384, ppw = 32, W = 16, CTAS = 296 * 8;  // 8 waves of CTAs
  CK(cudaMalloc(&cyc, 2 * CTAS * sizeof(long long)));
  CK(cudaMalloc(&out, (size_t)CTAS * 512 * 4));
  std::vector<long long> h(2 * CTAS);
  const size_t n_int4 = (size_t)(CTAS / 16) * W * ppw * STEPS * 2;
  std::vector<int4> hs(n_int4);
  uint32_t rng = 12345;
  for (size_t i = 0; i < hs.size(); ++i) {
    rng = rng * 1664525u + 1013904223u;
    hs[i].x = ((rng >> 8) % lines) * 256;
    float a = 0.5f, b = 0.25f, c = -0.125f;
    hs[i].y = *(int*)&a; hs[i].z = *(int*)&b; hs[i].w = *(int*)&c;
  }
  int4* gs;
  CK(cudaMalloc(&gs, n_int4 * sizeof(int4)));
  CK(cudaMemcpy(gs, hs.data(), n_int4 * sizeof(int4), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(MODE, BATCH, SH)                                                                                     \
  do {                                                                                                      \
    size_t smem = lines * 256 + (MODE == 2 ? W * 4 * 512 : 0);                                              \
    CK(cudaFuncSetAttribute(k_fwd<MODE, BATCH, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    float ms = 0;                                                                                           \
    for (int rep = 0; rep < 2; ++rep) {                                                                     \
      cudaEventRecord(e0);                                                                                  \
      k_fwd<MODE, BATCH, SH><<<CTAS, W * 32, smem>>>(gs, ppw, lines, cyc, out);                                  \
      cudaEventRecord(e1);                                                                                  \
      CK(cudaDeviceSynchronize());                                                                          \
      cudaEventElapsedTime(&ms, e0, e1);                                                                    \
    }                                                                                                       \
    double steps = (double)CTAS * W * ppw * STEPS;                                                          \
    cudaMemcpy(h.data(), cyc, 2 * CTAS * sizeof(long long), cudaMemcpyDeviceToHost);                        \
    double sc = 0, sg = 0;                                                                                  \
    for (int i = 0; i < CTAS; ++i) { sc += (double)h[2 * i]; sg += (double)h[2 * i + 1]; }                  \
    printf("fwd mode=%d batch=%d sharedL1=%d: %.3f ms; clock64: %.3f cycles/step/SM (loop only), SM clock %.3f GHz, event-based %.3f cycles/step/SM\n", MODE, BATCH, SH, ms, \
           (sc / CTAS) / (2.0 * W * ppw * STEPS), sc / sg, ms * 1e-3 * (sc / sg) * 1e9 * 148 / steps);        \
  } while (0)
  RUN(0, 4, 0); RUN(0, 8, 0);
  RUN(1, 4, 0); RUN(1, 8, 0);
  RUN(2, 4, 0); RUN(2, 8, 0);
  RUN(0, 4, 1); RUN(1, 4, 1); RUN(2, 4, 1);
  return 0;
}
