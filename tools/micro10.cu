// Micro-benchmark, round 10 (r02): cost model of the PATCH formulation of the fused residual kernels.
//
// A patch = up to 4 velocity nodes (row pairs I[k], J[k]) + 1 single (pressure) row owned by ONE warp for 64 samples
// (2 samples per lane, packed fp32x2 arithmetic).  The warp walks the union of the patch's columns: every gathered line
// feeds all the rows of the patch that couple to it, from registers.  Per column step: one uniform 16-byte header
// {lines, kind | target mask}, 2 (forward) / 4 (backward) LDS.64 gathers, ceil(3k/4) / ceil(5k/4) uniform coefficient
// words for the k targets in the mask.  Questions: (1) cycles per patch and slab with W warps per SM (compute only),
// (2) LDGSTS staging rate of scattered 256-byte lines with 1, 2, 4 producer warps.
//
// Synthetic patch = interior vertex patch of the structured P2-P1 mesh: 19 node steps (the vertex couples to all 19,
// each of its 3 edge nodes to 9 of them), the pressure row takes all 38 velocity values; 7 pressure-column steps
// (vertex: 7, edges: 4 each).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int4 lds128(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ u64 lds64(uint32_t a) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
struct P2 { float lo, hi; };
// (lo, hi) += a * (b.lo, b.hi)
__device__ __forceinline__ void fma2s(P2& d, float a, u64 b) {
  asm("{\n.reg .b64 c, aa;\nmov.b64 c, {%0,%1};\nmov.b64 aa, {%2,%2};\nfma.rn.f32x2 c, aa, %3, c;\nmov.b64 {%0,%1}, c;\n}" : "+f"(d.lo), "+f"(d.hi) : "f"(a), "l"(b));
}
// (lo, hi) += a * b (both packed)
__device__ __forceinline__ void fma2p(P2& d, u64 a, u64 b) {
  asm("{\n.reg .b64 c;\nmov.b64 c, {%0,%1};\nfma.rn.f32x2 c, %2, %3, c;\nmov.b64 {%0,%1}, c;\n}" : "+f"(d.lo), "+f"(d.hi) : "l"(a), "l"(b));
}
// returns a * b + c with scalar a, packed b, c
__device__ __forceinline__ u64 fma2r(float a, u64 b, u64 c) {
  u64 d;
  asm("{\n.reg .b64 aa;\nmov.b64 aa, {%1,%1};\nfma.rn.f32x2 %0, aa, %2, %3;\n}" : "=l"(d) : "f"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 bc2(float a) {
  u64 r;
  asm("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(a));
  return r;
}

constexpr int kLines = 400;
constexpr int kStreamBytes = 4096;  // per warp

struct FwdAcc { P2 aI, uI, vI, aJ, uJ, vJ; };
struct BwdAcc { P2 gI, gJ, b1I, b2I, b1J, b2J; };

// ---- forward bodies ----
template <int MASK>
__device__ __forceinline__ void fwd_k1(FwdAcc (&acc)[4], uint32_t cp, u64 xI, u64 xJ) {
  constexpr int k = ((MASK >> 0) & 1) + ((MASK >> 1) & 1) + ((MASK >> 2) & 1) + ((MASK >> 3) & 1);
  constexpr int nw = (3 * k + 3) / 4;
  float c[nw * 4 + 1];
#pragma unroll
  for (int w = 0; w < nw; ++w) {
    const int4 v = lds128(cp + w * 16);
    c[4 * w] = __int_as_float(v.x); c[4 * w + 1] = __int_as_float(v.y); c[4 * w + 2] = __int_as_float(v.z); c[4 * w + 3] = __int_as_float(v.w);
  }
  int j = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if ((MASK >> t) & 1) {
      fma2s(acc[t].aI, c[j], xI); fma2s(acc[t].uI, c[j + 1], xI); fma2s(acc[t].vI, c[j + 2], xI);
      fma2s(acc[t].aJ, c[j], xJ); fma2s(acc[t].uJ, c[j + 1], xJ); fma2s(acc[t].vJ, c[j + 2], xJ);
      j += 3;
    }
}
template <int MASK>
__device__ __forceinline__ void fwd_k2(FwdAcc (&acc)[4], uint32_t cp, u64 x) {
  constexpr int k = ((MASK >> 0) & 1) + ((MASK >> 1) & 1) + ((MASK >> 2) & 1) + ((MASK >> 3) & 1);
  constexpr int nw = (2 * k + 3) / 4;
  float c[nw * 4 + 1];
#pragma unroll
  for (int w = 0; w < nw; ++w) {
    const int4 v = lds128(cp + w * 16);
    c[4 * w] = __int_as_float(v.x); c[4 * w + 1] = __int_as_float(v.y); c[4 * w + 2] = __int_as_float(v.z); c[4 * w + 3] = __int_as_float(v.w);
  }
  int j = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if ((MASK >> t) & 1) {
      fma2s(acc[t].aI, c[j], x); fma2s(acc[t].aJ, c[j + 1], x);
      j += 2;
    }
}
#define SW16(F, ...)                                                                                                   \
  switch (mask) {                                                                                                     \
    case 1: F<1>(__VA_ARGS__); break;   case 2: F<2>(__VA_ARGS__); break;   case 3: F<3>(__VA_ARGS__); break;            \
    case 4: F<4>(__VA_ARGS__); break;   case 5: F<5>(__VA_ARGS__); break;   case 6: F<6>(__VA_ARGS__); break;            \
    case 7: F<7>(__VA_ARGS__); break;   case 8: F<8>(__VA_ARGS__); break;   case 9: F<9>(__VA_ARGS__); break;            \
    case 10: F<10>(__VA_ARGS__); break; case 11: F<11>(__VA_ARGS__); break; case 12: F<12>(__VA_ARGS__); break;         \
    case 13: F<13>(__VA_ARGS__); break; case 14: F<14>(__VA_ARGS__); break; case 15: F<15>(__VA_ARGS__); break;         \
    default: break;                                                                                                   \
  }

// ---- backward bodies ----
template <int MASK>
__device__ __forceinline__ void bwd_b1(BwdAcc (&acc)[4], uint32_t cp, u64 rI, u64 rJ, u64 d1, u64 d2) {
  constexpr int k = ((MASK >> 0) & 1) + ((MASK >> 1) & 1) + ((MASK >> 2) & 1) + ((MASK >> 3) & 1);
  constexpr int nw = (5 * k + 3) / 4;
  float c[nw * 4 + 1];
#pragma unroll
  for (int w = 0; w < nw; ++w) {
    const int4 v = lds128(cp + w * 16);
    c[4 * w] = __int_as_float(v.x); c[4 * w + 1] = __int_as_float(v.y); c[4 * w + 2] = __int_as_float(v.z); c[4 * w + 3] = __int_as_float(v.w);
  }
  int j = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if ((MASK >> t) & 1) {
      u64 tt = fma2r(c[j + 1], d1, bc2(c[j]));
      tt = fma2r(c[j + 2], d2, tt);
      fma2p(acc[t].gI, rI, tt); fma2p(acc[t].gJ, rJ, tt);
      fma2s(acc[t].b1I, c[j + 3], d1); fma2s(acc[t].b2I, c[j + 4], d1);
      fma2s(acc[t].b1J, c[j + 3], d2); fma2s(acc[t].b2J, c[j + 4], d2);
      j += 5;
    }
}
template <int MASK>
__device__ __forceinline__ void bwd_b2(BwdAcc (&acc)[4], uint32_t cp, u64 r) {
  constexpr int k = ((MASK >> 0) & 1) + ((MASK >> 1) & 1) + ((MASK >> 2) & 1) + ((MASK >> 3) & 1);
  constexpr int nw = (2 * k + 3) / 4;
  float c[nw * 4 + 1];
#pragma unroll
  for (int w = 0; w < nw; ++w) {
    const int4 v = lds128(cp + w * 16);
    c[4 * w] = __int_as_float(v.x); c[4 * w + 1] = __int_as_float(v.y); c[4 * w + 2] = __int_as_float(v.z); c[4 * w + 3] = __int_as_float(v.w);
  }
  int j = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if ((MASK >> t) & 1) {
      fma2s(acc[t].gI, c[j], r); fma2s(acc[t].gJ, c[j + 1], r);
      j += 2;
    }
}

// stream: per patch [n_steps word][steps...]; step = header word {lines0, lines1, meta, coef} + coefficient words
// meta = kind | mask << 4 | n_coef_words << 8
template <int BWD, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) k_patch(const int4* __restrict__ gstream, int stream_words, int iters, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* lines = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < kLines * 64; i += blockDim.x) lines[i] = 1e-3f * (float)((i * 2654435761u) >> 20);
  int4* st = reinterpret_cast<int4*>(smem + kLines * 256) + warp * (kStreamBytes / 16);
  for (int i = lane; i < stream_words; i += 32) st[i] = gstream[i];
  __syncthreads();
  const uint32_t pool = smem_u32(smem) + lane * 8;
  const uint32_t sbase = smem_u32(st);
  P2 sacc = {0.f, 0.f};
  float total = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t sp = sbase;
    const int n_steps = lds128(sp).x;
    sp += 16;
    if (!BWD) {
      FwdAcc acc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t].aI = acc[t].uI = acc[t].vI = acc[t].aJ = acc[t].uJ = acc[t].vJ = P2{0.f, 0.f};
      int4 h = lds128(sp);
#pragma unroll 1
      for (int s = 0; s < n_steps; ++s) {
        const int4 hc = h;
        const uint32_t cp = sp + 16;
        const int kind = hc.z & 15, mask = (hc.z >> 4) & 15;
        sp = cp + ((hc.z >> 8) & 255) * 16;
        h = lds128(sp);  // next header (the stream ends with a dummy header)
        const uint32_t rot = (uint32_t)(it & 7) << 8;  // vary the lines a little from patch to patch
        if (kind == 0) {
          const u64 xI = lds64(pool + (((uint32_t)hc.x & 0xffffu) << 8) + rot), xJ = lds64(pool + (((uint32_t)hc.x >> 16) << 8) + rot);
          fma2s(sacc, __int_as_float(hc.y), xI);
          fma2s(sacc, __int_as_float(hc.w), xJ);
          SW16(fwd_k1, acc, cp, xI, xJ)
        } else {
          const u64 x = lds64(pool + (((uint32_t)hc.x & 0xffffu) << 8) + rot);
          fma2s(sacc, __int_as_float(hc.y), x);
          SW16(fwd_k2, acc, cp, x)
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        total += acc[t].aI.lo + acc[t].aI.hi + acc[t].uI.lo * acc[t].vI.hi + acc[t].uI.hi * acc[t].vI.lo + acc[t].aJ.lo + acc[t].aJ.hi + acc[t].uJ.lo * acc[t].vJ.hi + acc[t].uJ.hi * acc[t].vJ.lo;
    } else {
      BwdAcc acc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t].gI = acc[t].gJ = acc[t].b1I = acc[t].b2I = acc[t].b1J = acc[t].b2J = P2{0.f, 0.f};
      int4 h = lds128(sp);
#pragma unroll 1
      for (int s = 0; s < n_steps; ++s) {
        const int4 hc = h;
        const uint32_t cp = sp + 16;
        const int kind = hc.z & 15, mask = (hc.z >> 4) & 15;
        sp = cp + ((hc.z >> 8) & 255) * 16;
        h = lds128(sp);
        const uint32_t rot = (uint32_t)(it & 7) << 8;
        if (kind == 0) {
          const u64 rI = lds64(pool + (((uint32_t)hc.x & 0xffffu) << 8) + rot), rJ = lds64(pool + (((uint32_t)hc.x >> 16) << 8) + rot);
          const u64 d1 = lds64(pool + (((uint32_t)hc.y & 0xffffu) << 8) + rot), d2 = lds64(pool + (((uint32_t)hc.y >> 16) << 8) + rot);
          // single target: its two coefficients are the first two floats of the coefficient block
          const int4 c0 = lds128(cp);
          fma2s(sacc, __int_as_float(c0.x), rI);
          fma2s(sacc, __int_as_float(c0.y), rJ);
          SW16(bwd_b1, acc, cp + 16, rI, rJ, d1, d2)
        } else {
          const u64 r = lds64(pool + (((uint32_t)hc.x & 0xffffu) << 8) + rot);
          fma2s(sacc, __int_as_float(hc.w), r);
          SW16(bwd_b2, acc, cp, r)
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        total += acc[t].gI.lo + acc[t].gI.hi + acc[t].gJ.lo + acc[t].gJ.hi + acc[t].b1I.lo * acc[t].b2I.hi + acc[t].b1I.hi * acc[t].b2I.lo + acc[t].b1J.lo * acc[t].b2J.hi + acc[t].b1J.hi * acc[t].b2J.lo;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = total + sacc.lo + sacc.hi;
  (void)nwarps;
}


// ---------------------------------------------------------------------------------------------------------------
// Variant 2: steps sorted by target-mask CLASS (run-length encoded): one dispatch per run, fixed-size steps inside a run,
// next step's words prefetched, byte offsets in the stream (one IADD per gather).  K2/B2 steps are dense (4 targets).
//   run word   : {kind | mask << 4 | count << 8, 0, 0, 0}
//   fwd K1 step: {offI, offJ, sI, sJ} + 3k coefficients            (nw = ceil((4 + 3k) / 4) words)
//   fwd K2 step: {off, s, aI0, aJ0} {aI1, aJ1, aI2, aJ2} {aI3, aJ3, -, -}
//   bwd B1 step: {offrI, offrJ, offaI, offaJ} {sI, sJ, 5k coefficients...}   (nw = ceil((6 + 5k) / 4))
//   bwd B2 step: as K2
template <int N>
struct Words { int4 w[N]; };
template <int N>
__device__ __forceinline__ Words<N> ldw(uint32_t a) {
  Words<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.w[i] = lds128(a + 16 * i);
  return r;
}
template <int N>
__device__ __forceinline__ float wf(const Words<N>& w, int j) {
  const int4& v = w.w[j >> 2];
  return __int_as_float((j & 3) == 0 ? v.x : (j & 3) == 1 ? v.y : (j & 3) == 2 ? v.z : v.w);
}
__host__ __device__ constexpr int popc4(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

template <int MASK>
__device__ __forceinline__ void fwd_k1_run(FwdAcc (&acc)[4], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (4 + 3 * k + 3) / 4;
  Words<nw> nx = ldw<nw>(sp);
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
    const Words<nw> c = nx;
    sp += nw * 16;
    nx = ldw<nw>(sp);
    const u64 xI = lds64(pool + (uint32_t)c.w[0].x), xJ = lds64(pool + (uint32_t)c.w[0].y);
    fma2s(sacc, wf(c, 2), xI);
    fma2s(sacc, wf(c, 3), xJ);
    int j = 4;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((MASK >> t) & 1) {
        fma2s(acc[t].aI, wf(c, j), xI); fma2s(acc[t].uI, wf(c, j + 1), xI); fma2s(acc[t].vI, wf(c, j + 2), xI);
        fma2s(acc[t].aJ, wf(c, j), xJ); fma2s(acc[t].uJ, wf(c, j + 1), xJ); fma2s(acc[t].vJ, wf(c, j + 2), xJ);
        j += 3;
      }
  }
}
__device__ __forceinline__ void fwd_k2_run(FwdAcc (&acc)[4], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  Words<3> nx = ldw<3>(sp);
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
    const Words<3> c = nx;
    sp += 48;
    nx = ldw<3>(sp);
    const u64 x = lds64(pool + (uint32_t)c.w[0].x);
    fma2s(sacc, wf(c, 1), x);
#pragma unroll
    for (int t = 0; t < 4; ++t) { fma2s(acc[t].aI, wf(c, 2 + 2 * t), x); fma2s(acc[t].aJ, wf(c, 3 + 2 * t), x); }
  }
}
template <int MASK>
__device__ __forceinline__ void bwd_b1_run(BwdAcc (&acc)[4], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (6 + 5 * k + 3) / 4;
  Words<nw> nx = ldw<nw>(sp);
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
    const Words<nw> c = nx;
    sp += nw * 16;
    nx = ldw<nw>(sp);
    const u64 rI = lds64(pool + (uint32_t)c.w[0].x), rJ = lds64(pool + (uint32_t)c.w[0].y);
    const u64 d1 = lds64(pool + (uint32_t)c.w[0].z), d2 = lds64(pool + (uint32_t)c.w[0].w);
    fma2s(sacc, wf(c, 4), rI);
    fma2s(sacc, wf(c, 5), rJ);
    int j = 6;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((MASK >> t) & 1) {
        u64 tt = fma2r(wf(c, j + 1), d1, bc2(wf(c, j)));
        tt = fma2r(wf(c, j + 2), d2, tt);
        fma2p(acc[t].gI, rI, tt); fma2p(acc[t].gJ, rJ, tt);
        fma2s(acc[t].b1I, wf(c, j + 3), d1); fma2s(acc[t].b2I, wf(c, j + 4), d1);
        fma2s(acc[t].b1J, wf(c, j + 3), d2); fma2s(acc[t].b2J, wf(c, j + 4), d2);
        j += 5;
      }
  }
}
__device__ __forceinline__ void bwd_b2_run(BwdAcc (&acc)[4], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  Words<3> nx = ldw<3>(sp);
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
    const Words<3> c = nx;
    sp += 48;
    nx = ldw<3>(sp);
    const u64 x = lds64(pool + (uint32_t)c.w[0].x);
    fma2s(sacc, wf(c, 1), x);
#pragma unroll
    for (int t = 0; t < 4; ++t) { fma2s(acc[t].gI, wf(c, 2 + 2 * t), x); fma2s(acc[t].gJ, wf(c, 3 + 2 * t), x); }
  }
}
#define SWRUN(F, ...)                                                                                                  \
  switch (mask) {                                                                                                     \
    case 1: F<1>(__VA_ARGS__); break;   case 2: F<2>(__VA_ARGS__); break;   case 3: F<3>(__VA_ARGS__); break;            \
    case 4: F<4>(__VA_ARGS__); break;   case 5: F<5>(__VA_ARGS__); break;   case 6: F<6>(__VA_ARGS__); break;            \
    case 7: F<7>(__VA_ARGS__); break;   case 8: F<8>(__VA_ARGS__); break;   case 9: F<9>(__VA_ARGS__); break;            \
    case 10: F<10>(__VA_ARGS__); break; case 11: F<11>(__VA_ARGS__); break; case 12: F<12>(__VA_ARGS__); break;         \
    case 13: F<13>(__VA_ARGS__); break; case 14: F<14>(__VA_ARGS__); break; case 15: F<15>(__VA_ARGS__); break;         \
    default: break;                                                                                                   \
  }

template <int BWD, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) k_patch2(const int4* __restrict__ gstream, int stream_words, int iters, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* lines = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < kLines * 64; i += blockDim.x) lines[i] = 1e-3f * (float)((i * 2654435761u) >> 20);
  int4* st = reinterpret_cast<int4*>(smem + kLines * 256) + warp * (kStreamBytes / 16);
  for (int i = lane; i < stream_words; i += 32) st[i] = gstream[i];
  __syncthreads();
  const uint32_t pool0 = smem_u32(smem) + lane * 8;
  const uint32_t sbase = smem_u32(st);
  P2 sacc = {0.f, 0.f};
  float total = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t sp = sbase;
    const int n_runs = lds128(sp).x;
    sp += 16;
    const uint32_t pool = pool0 + ((uint32_t)(it & 7) << 8);
    if (!BWD) {
      FwdAcc acc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t].aI = acc[t].uI = acc[t].vI = acc[t].aJ = acc[t].uJ = acc[t].vJ = P2{0.f, 0.f};
#pragma unroll 1
      for (int r = 0; r < n_runs; ++r) {
        const int meta = lds128(sp).x;
        sp += 16;
        const int kind = meta & 15, mask = (meta >> 4) & 15, count = meta >> 8;
        if (kind == 0) { SWRUN(fwd_k1_run, acc, sacc, sp, count, pool) }
        else fwd_k2_run(acc, sacc, sp, count, pool);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        total += acc[t].aI.lo + acc[t].aI.hi + acc[t].uI.lo * acc[t].vI.hi + acc[t].uI.hi * acc[t].vI.lo + acc[t].aJ.lo + acc[t].aJ.hi + acc[t].uJ.lo * acc[t].vJ.hi + acc[t].uJ.hi * acc[t].vJ.lo;
    } else {
      BwdAcc acc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t].gI = acc[t].gJ = acc[t].b1I = acc[t].b2I = acc[t].b1J = acc[t].b2J = P2{0.f, 0.f};
#pragma unroll 1
      for (int r = 0; r < n_runs; ++r) {
        const int meta = lds128(sp).x;
        sp += 16;
        const int kind = meta & 15, mask = (meta >> 4) & 15, count = meta >> 8;
        if (kind == 0) { SWRUN(bwd_b1_run, acc, sacc, sp, count, pool) }
        else bwd_b2_run(acc, sacc, sp, count, pool);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        total += acc[t].gI.lo + acc[t].gI.hi + acc[t].gJ.lo + acc[t].gJ.hi + acc[t].b1I.lo * acc[t].b2I.hi + acc[t].b1I.hi * acc[t].b2I.lo + acc[t].b1J.lo * acc[t].b2J.hi + acc[t].b1J.hi * acc[t].b2J.lo;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = total + sacc.lo + sacc.hi;
}

// ---- LDGSTS staging of scattered lines by `pw` producer warps; two rounds in flight ----
__global__ void __launch_bounds__(512, 1) k_stage(const float* __restrict__ base, long long ldb, const int* __restrict__ line_dofs, int lines_per_round,
                                                 int rounds, int pw, int n_slabs) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= pw) return;
  const uint32_t sb = smem_u32(smem);
  const int slab = blockIdx.x % n_slabs;
  const int seg = blockIdx.x / n_slabs;
  for (int r = 0; r < rounds; ++r) {
    const int* ld = line_dofs + ((size_t)(seg * rounds + r) * lines_per_round);
    const uint32_t dst0 = sb + (r & 1) * (lines_per_round * 256);
    for (int l = warp * 2 + (lane >> 4); l < lines_per_round; l += pw * 2) {
      const int dof = __ldg(ld + l);
      const float* src = base + (size_t)dof * ldb + slab * 64 + (lane & 15) * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + l * 256 + (lane & 15) * 16), "l"(src));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 1;");
  }
  asm volatile("cp.async.wait_group 0;");
}

static void push(std::vector<int4>& s, int a, int b, int c, int d) { s.push_back(make_int4(a, b, c, d)); }
static int fbits(float f) { int i; memcpy(&i, &f, 4); return i; }

int main(int argc, char** argv) {
  // ---- synthetic interior vertex patch ----
  std::vector<int4> fs, bs;
  srand(1);
  auto rl = []() { return rand() % (kLines - 8); };
  {
    push(fs, 26, 0, 0, 0);
    push(bs, 26, 0, 0, 0);
    for (int i = 0; i < 19; ++i) {
      const int mask = 1 | ((i % 2 == 0) ? 2 : 0) | ((i % 2 == 1) ? 4 : 0) | ((i < 9) ? 8 : 0);
      const int k = __builtin_popcount(mask);
      int nw = (3 * k + 3) / 4;
      push(fs, rl() | (rl() << 16), fbits(0.01f), 0 | (mask << 4) | (nw << 8), fbits(0.02f));
      for (int w = 0; w < nw; ++w) push(fs, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
      nw = 1 + (5 * k + 3) / 4;
      push(bs, rl() | (rl() << 16), rl() | (rl() << 16), 0 | (mask << 4) | (nw << 8), 0);
      for (int w = 0; w < nw; ++w) push(bs, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
    }
    const int pm[7] = {15, 3, 5, 9, 7, 11, 13};  // 19 bits in total
    for (int i = 0; i < 7; ++i) {
      const int mask = pm[i], k = __builtin_popcount(mask), nw = (2 * k + 3) / 4;
      push(fs, rl(), fbits(0.01f), 1 | (mask << 4) | (nw << 8), 0);
      for (int w = 0; w < nw; ++w) push(fs, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
      push(bs, rl(), 0, 1 | (mask << 4) | (nw << 8), fbits(0.01f));
      for (int w = 0; w < nw; ++w) push(bs, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
    }
    push(fs, 0, 0, 0, 0);
    push(bs, 0, 0, 0, 0);
  }
  printf("stream words per patch: fwd %zu (%zu B), bwd %zu (%zu B)\n", fs.size(), fs.size() * 16, bs.size(), bs.size() * 16);
  if (fs.size() * 16 > kStreamBytes || bs.size() * 16 > kStreamBytes) { printf("stream too long\n"); return 1; }
  int4 *dfs, *dbs;
  float* out;
  CK(cudaMalloc(&dfs, fs.size() * 16)); CK(cudaMalloc(&dbs, bs.size() * 16)); CK(cudaMalloc(&out, 148 * 1024 * 4));
  CK(cudaMemcpy(dfs, fs.data(), fs.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbs, bs.data(), bs.size() * 16, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev_clock_khz = 0; cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  const double ghz = 1.95;
  const int iters = 2000;
  for (int bwd = 0; bwd < 2; ++bwd)
    for (int W : {8, 10, 12, 16, 20, 24}) {
      const size_t sm = (size_t)kLines * 256 + (size_t)W * kStreamBytes;
      if (sm > 232448) continue;
      float ms = 0;
#define RUNW(WW)                                                                                                     \
  if (W == WW) {                                                                                                     \
    if (bwd) CK(cudaFuncSetAttribute(k_patch<1, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));          \
    else CK(cudaFuncSetAttribute(k_patch<0, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));              \
    for (int rep = 0; rep < 2; ++rep) {                                                                              \
      cudaEventRecord(e0);                                                                                           \
      if (bwd) k_patch<1, WW><<<148, W * 32, sm>>>(dbs, (int)bs.size(), iters, out);                                  \
      else k_patch<0, WW><<<148, W * 32, sm>>>(dfs, (int)fs.size(), iters, out);                                      \
      cudaEventRecord(e1);                                                                                           \
      CK(cudaDeviceSynchronize());                                                                                   \
      cudaEventElapsedTime(&ms, e0, e1);                                                                             \
    }                                                                                                                \
  }
      RUNW(8) RUNW(10) RUNW(12) RUNW(16) RUNW(20) RUNW(24)
      const double cyc_per_patch = ms * 1e-3 * ghz * 1e9 / ((double)W * iters);
      printf("%s W=%2d: %.3f ms, %.1f SM-cycles per patch and slab, %.1f per dof and slab (9 dofs) -> cfg5 kernel estimate %.2f ms\n", bwd ? "bwd" : "fwd", W, ms,
             cyc_per_patch, cyc_per_patch / 9.0, cyc_per_patch / 9.0 * 1001334.0 * 16 / 148 / (ghz * 1e9) * 1e3);
    }
  // ---- variant 2: run-length classes ----
  {
    std::vector<int4> f2, b2;
    const int cls_mask[8] = {1, 3, 5, 9, 11, 13, 7, 15};
    const int cls_cnt[8] = {2, 3, 3, 3, 2, 2, 0, 4};
    int n_runs = 0;
    for (int c = 0; c < 8; ++c) n_runs += cls_cnt[c] > 0;
    n_runs += 1;  // the dense K2/B2 run
    push(f2, n_runs, 0, 0, 0);
    push(b2, n_runs, 0, 0, 0);
    auto off = [&]() { return rl() * 256; };
    for (int c = 0; c < 8; ++c) {
      if (cls_cnt[c] == 0) continue;
      const int k = __builtin_popcount(cls_mask[c]);
      push(f2, 0 | (cls_mask[c] << 4) | (cls_cnt[c] << 8), 0, 0, 0);
      push(b2, 0 | (cls_mask[c] << 4) | (cls_cnt[c] << 8), 0, 0, 0);
      for (int i = 0; i < cls_cnt[c]; ++i) {
        const int nwf = (4 + 3 * k + 3) / 4, nwb = (6 + 5 * k + 3) / 4;
        push(f2, off(), off(), fbits(0.01f), fbits(0.02f));
        for (int w = 1; w < nwf; ++w) push(f2, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
        push(b2, off(), off(), off(), off());
        for (int w = 1; w < nwb; ++w) push(b2, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f));
      }
    }
    push(f2, 1 | (7 << 8), 0, 0, 0);
    push(b2, 1 | (7 << 8), 0, 0, 0);
    for (int i = 0; i < 7; ++i) {
      push(f2, off(), fbits(0.01f), fbits(0.1f), fbits(0.2f)); push(f2, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f)); push(f2, fbits(0.1f), fbits(0.2f), 0, 0);
      push(b2, off(), fbits(0.01f), fbits(0.1f), fbits(0.2f)); push(b2, fbits(0.1f), fbits(0.2f), fbits(0.3f), fbits(0.4f)); push(b2, fbits(0.1f), fbits(0.2f), 0, 0);
    }
    for (int i = 0; i < 8; ++i) { push(f2, 0, 0, 0, 0); push(b2, 0, 0, 0, 0); }
    printf("variant 2 stream words per patch: fwd %zu (%zu B), bwd %zu (%zu B)\n", f2.size(), f2.size() * 16, b2.size(), b2.size() * 16);
    if (f2.size() * 16 > kStreamBytes || b2.size() * 16 > kStreamBytes) { printf("stream too long\n"); return 1; }
    int4 *df2, *db2;
    CK(cudaMalloc(&df2, f2.size() * 16)); CK(cudaMalloc(&db2, b2.size() * 16));
    CK(cudaMemcpy(df2, f2.data(), f2.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db2, b2.data(), b2.size() * 16, cudaMemcpyHostToDevice));
    for (int bwd = 0; bwd < 2; ++bwd)
      for (int W : {8, 10, 12, 16, 20}) {
        const size_t sm = (size_t)kLines * 256 + (size_t)W * kStreamBytes;
        float ms = 0;
#define RUNW2(WW)                                                                                                    \
  if (W == WW) {                                                                                                     \
    if (bwd) CK(cudaFuncSetAttribute(k_patch2<1, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));         \
    else CK(cudaFuncSetAttribute(k_patch2<0, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));             \
    for (int rep = 0; rep < 2; ++rep) {                                                                              \
      cudaEventRecord(e0);                                                                                           \
      if (bwd) k_patch2<1, WW><<<148, W * 32, sm>>>(db2, (int)b2.size(), iters, out);                                 \
      else k_patch2<0, WW><<<148, W * 32, sm>>>(df2, (int)f2.size(), iters, out);                                     \
      cudaEventRecord(e1);                                                                                           \
      CK(cudaDeviceSynchronize());                                                                                   \
      cudaEventElapsedTime(&ms, e0, e1);                                                                             \
    }                                                                                                                \
  }
        RUNW2(8) RUNW2(10) RUNW2(12) RUNW2(16) RUNW2(20)
        const double cyc_per_patch = ms * 1e-3 * ghz * 1e9 / ((double)W * iters);
        printf("v2 %s W=%2d: %.3f ms, %.1f SM-cycles per patch and slab, %.1f per dof and slab (9 dofs) -> cfg5 kernel estimate %.2f ms\n", bwd ? "bwd" : "fwd", W, ms,
               cyc_per_patch, cyc_per_patch / 9.0, cyc_per_patch / 9.0 * 1001334.0 * 16 / 148 / (ghz * 1e9) * 1e3);
      }
  }
  // ---- staging ----
  {
    const long long ldb = 1024, N = 1000000;
    float* a;
    CK(cudaMalloc(&a, (size_t)N * ldb * 4));
    CK(cudaMemset(a, 0, (size_t)N * ldb * 4));
    const int n_slabs = 16, segs = 148 / n_slabs + 1, rounds = 200;
    for (int lpr : {256, 400}) {
      // lines of a round: scattered runs of 3 dofs (interleaved u1,u2,p) around a window that advances with the round
      std::vector<int> ld((size_t)segs * rounds * lpr);
      for (int sg = 0; sg < segs; ++sg)
        for (int r = 0; r < rounds; ++r)
          for (int l = 0; l < lpr; ++l) {
            const long long basedof = ((long long)sg * rounds + r) * 160 % (N - 5000);
            ld[((size_t)sg * rounds + r) * lpr + l] = (int)(basedof + (l / 24) * 1700 % 4000 + (l % 24));
          }
      int* dld;
      CK(cudaMalloc(&dld, ld.size() * 4));
      CK(cudaMemcpy(dld, ld.data(), ld.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 400 * 256));
      for (int pw : {1, 2, 4, 16}) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          k_stage<<<148, 512, 2 * 400 * 256>>>(a, ldb, dld, lpr, rounds, pw, n_slabs);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          cudaEventElapsedTime(&ms, e0, e1);
        }
        const double bytes = 148.0 * rounds * lpr * 256;
        printf("LDGSTS staging, %d lines per round, %2d producer warps: %.3f ms, %.0f GB/s staged (%.1f B/clk/SM), %.0f cycles per round\n", lpr, pw, ms,
               bytes / ms * 1e-6, bytes / (ms * 1e-3) / 148 / (ghz * 1e9), ms * 1e-3 * ghz * 1e9 / rounds);
      }
      cudaFree(dld);
    }
  }
  return 0;
}
