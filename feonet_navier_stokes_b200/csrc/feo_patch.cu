// sm_100a kernels of the fused sparse residual, patch formulation (plan + stream format: feo_patch.h, feo_patch_plan.cpp).
//
// One persistent CTA per SM processes (segment, slab) items; within an item it walks the segment's rounds through a
// shared-memory POOL of dof lines (256 B = the 64 samples of one dof for this slab) that stay resident from round to
// round:
//   * producer warp(s): for round g + 1, after every consumer warp has released round g - 1, fetch the lines the plan
//     lists for it (16-byte cp.async copies, LDGSTS: any line to any slot, no tensor maps, zero fill past ldb) and the
//     round's operator stream (one cp.async.bulk) -- both complete on the mbarrier full[(g + 1) & 1];
//   * consumer warps: ONE warp evaluates one patch (<= 4 velocity nodes + 1 single dof) for the 64 samples, 2 samples per
//     lane, every accumulator in registers, packed fp32x2 FMAs (FFMA2).  A gathered line feeds every row (column) of the
//     patch that couples to it; the coefficients are warp-uniform LDS.64 reads of the stream.  Steps are grouped by target
//     mask (runs), so the dispatch on the mask happens once per run and the code of a run is straight-line.
// Row-/column-owned with a fixed summation order: no atomics, bit-reproducible on a given device.
// Measured on B200 (tools/micro10.cu): LDS.64 warp-uniform = 1 cycle / 8 B, LDS.64 gather = 2 cycles / 256 B, every integer
// instruction costs as much pipe time as an FFMA2 (2 cycles per warp and SM sub-partition) -- hence run-length classes
// instead of a per-step switch, and packed 16-bit line references.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "feo_patch.h"

namespace feo {
namespace {

typedef unsigned long long u64;
constexpr uint32_t kSmemMax = 232448;

struct PatchParams {
  const int32_t* seg_ptr;
  const RoundInfo* rounds;
  const LineLoad* loads;
  const uint32_t* stream;
  const float* src0;       // forward: alpha ; backward: r
  const float* src1;       // backward: alpha
  const float* fT;         // forward: load vectors
  float* outT;             // forward: rT (may be NULL) ; backward: gradT
  float* partials;         // forward: one loss partial per (CTA, consumer warp)
  const float* grad_loss;  // backward: upstream gradient (NULL = 1)
  int64_t ldb;
  int32_t B, n_slabs, n_items;
  int32_t g_div, g_mod;    // gridDim.x / n_slabs, gridDim.x % n_slabs
  int32_t warps, producers;
  uint32_t stream_cap, stream_off, bar_off;
  int32_t debug;           // FEO_DEBUG_MODE: 1 = staging only, 2 = no line loads (compute only; results are garbage)
  int32_t precond;         // forward: 1 -> r = lhs - (f - c), 0 -> r = lhs - (-f + c)
  float esign;             // backward: +1 precond branch, -1 otherwise
};

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// 16-byte asynchronous copy global -> shared; src_bytes = 0 writes zeros
__device__ __forceinline__ void ldgsts16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival when all cp.async copies this thread has issued so far have landed
__device__ __forceinline__ void ldgsts_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
struct U2 {
  uint32_t lo, hi;
};
__device__ __forceinline__ U2 lds_unit(uint32_t a) {  // one 8-byte stream unit (warp-uniform address)
  U2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.lo), "=r"(v.hi) : "r"(a));
  return v;
}
__device__ __forceinline__ u64 lds_pair(uint32_t a) {  // the lane's two samples of a staged line
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
struct P2 {
  float lo, hi;
};
// (lo, hi) += a * (b.lo, b.hi): one packed FMA, accumulator kept as two fp32 registers so that ptxas updates it in place
__device__ __forceinline__ void fma2s(P2& d, float a, u64 b) {
  asm("{\n"
      ".reg .b64 c, aa;\n"
      "mov.b64 c, {%0,%1};\n"
      "mov.b64 aa, {%2,%2};\n"
      "fma.rn.f32x2 c, aa, %3, c;\n"
      "mov.b64 {%0,%1}, c;\n"
      "}"
      : "+f"(d.lo), "+f"(d.hi)
      : "f"(a), "l"(b));
}
__device__ __forceinline__ void fma2p(P2& d, u64 a, u64 b) {
  asm("{\n"
      ".reg .b64 c;\n"
      "mov.b64 c, {%0,%1};\n"
      "fma.rn.f32x2 c, %2, %3, c;\n"
      "mov.b64 {%0,%1}, c;\n"
      "}"
      : "+f"(d.lo), "+f"(d.hi)
      : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 fma2r(float a, u64 b, u64 c) {  // a * b + c
  u64 d;
  asm("{\n"
      ".reg .b64 aa;\n"
      "mov.b64 aa, {%1,%1};\n"
      "fma.rn.f32x2 %0, aa, %2, %3;\n"
      "}"
      : "=l"(d)
      : "f"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 bc2(float a) {
  u64 r;
  asm("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ void unpk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float2 ldg2_stream(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// residual from LHS sum, load vector and convection in the reference's operation order:
// precond branch  r = LHS - (F - c) ; else  r = LHS - (-F + c)   (FEONet_steady_Navier-Stokes/train_FEONet.py:324-330, :356)
__device__ __forceinline__ float resid1(float lhs, float f, float c, bool precond) {
  return precond ? __fsub_rn(lhs, __fsub_rn(f, c)) : __fsub_rn(lhs, __fadd_rn(-f, c));
}
// c = u_i*Bu1 + u_j*Bu2 as two rounded products and one rounded add (train_FEONet.py:317-322)
__device__ __forceinline__ float conv1(float d1, float s1, float d2, float s2) { return __fadd_rn(__fmul_rn(d1, s1), __fmul_rn(d2, s2)); }

__host__ __device__ constexpr int popc4(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

// the 16-byte words of one piece of the stream (warp-uniform address)
template <int N>
struct Words {
  uint4 q[N];
  __device__ __forceinline__ uint32_t w(int j) const {
    const uint4& v = q[j >> 2];
    return (j & 3) == 0 ? v.x : (j & 3) == 1 ? v.y : (j & 3) == 2 ? v.z : v.w;
  }
  __device__ __forceinline__ float f(int j) const { return __uint_as_float(w(j)); }
};
__device__ __forceinline__ uint4 lds_word(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
template <int N>
__device__ __forceinline__ Words<N> read_words(uint32_t sp) {
  Words<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.q[i] = lds_word(sp + 16 * i);
  return r;
}

struct FwdAcc {
  P2 aI, uI, vI, aJ, uJ, vJ;
};
struct BwdAcc {
  P2 gI, gJ, b1I, b2I, b1J, b2J;
};

// ---- forward runs --------------------------------------------------------------------------------
// A run = `count` >= 1 steps of one kind and target mask; sp is left on the word after the run.
template <int MASK>
__device__ __forceinline__ void fwd_pair_run(FwdAcc (&acc)[kPatchNodes], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (4 + 3 * k + 3) / 4;
  const uint32_t end = sp + (uint32_t)count * (nw * 16);
#pragma unroll 1
  do {
    const Words<nw> c = read_words<nw>(sp);
    sp += nw * 16;
    const u64 x = lds_pair(pool + c.w(0)), y = lds_pair(pool + c.w(1));
    fma2s(sacc, c.f(2), x);
    fma2s(sacc, c.f(3), y);
    int j = 4;
#pragma unroll
    for (int t = 0; t < kPatchNodes; ++t)
      if ((MASK >> t) & 1) {
        fma2s(acc[t].aI, c.f(j), x);
        fma2s(acc[t].uI, c.f(j + 1), x);
        fma2s(acc[t].vI, c.f(j + 2), x);
        fma2s(acc[t].aJ, c.f(j), y);
        fma2s(acc[t].uJ, c.f(j + 1), y);
        fma2s(acc[t].vJ, c.f(j + 2), y);
        j += 3;
      }
  } while (sp != end);
}
template <int MASK>
__device__ __forceinline__ void fwd_plain_run(FwdAcc (&acc)[kPatchNodes], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (2 + 2 * k + 3) / 4;
  const uint32_t end = sp + (uint32_t)count * (nw * 16);
#pragma unroll 1
  do {
    const Words<nw> c = read_words<nw>(sp);
    sp += nw * 16;
    const u64 x = lds_pair(pool + c.w(0));
    fma2s(sacc, c.f(1), x);
    int j = 2;
#pragma unroll
    for (int t = 0; t < kPatchNodes; ++t)
      if ((MASK >> t) & 1) {
        fma2s(acc[t].aI, c.f(j), x);
        fma2s(acc[t].aJ, c.f(j + 1), x);
        j += 2;
      }
  } while (sp != end);
}
// ---- backward runs -------------------------------------------------------------------------------
template <int MASK>
__device__ __forceinline__ void bwd_pair_run(BwdAcc (&acc)[kPatchNodes], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (6 + 5 * k + 3) / 4;
  const uint32_t end = sp + (uint32_t)count * (nw * 16);
#pragma unroll 1
  do {
    const Words<nw> c = read_words<nw>(sp);
    sp += nw * 16;
    const u64 rI = lds_pair(pool + c.w(0)), rJ = lds_pair(pool + c.w(1));
    const u64 d1 = lds_pair(pool + c.w(2)), d2 = lds_pair(pool + c.w(3));
    fma2s(sacc, c.f(4), rI);
    fma2s(sacc, c.f(5), rJ);
    int j = 6;
#pragma unroll
    for (int t = 0; t < kPatchNodes; ++t)
      if ((MASK >> t) & 1) {
        u64 tt = fma2r(c.f(j + 1), d1, bc2(c.f(j)));
        tt = fma2r(c.f(j + 2), d2, tt);
        fma2p(acc[t].gI, rI, tt);
        fma2p(acc[t].gJ, rJ, tt);
        fma2s(acc[t].b1I, c.f(j + 3), d1);
        fma2s(acc[t].b2I, c.f(j + 4), d1);
        fma2s(acc[t].b1J, c.f(j + 3), d2);
        fma2s(acc[t].b2J, c.f(j + 4), d2);
        j += 5;
      }
  } while (sp != end);
}
template <int MASK>
__device__ __forceinline__ void bwd_plain_run(BwdAcc (&acc)[kPatchNodes], P2& sacc, uint32_t& sp, int count, uint32_t pool) {
  constexpr int k = popc4(MASK), nw = (2 + 2 * k + 3) / 4;
  const uint32_t end = sp + (uint32_t)count * (nw * 16);
#pragma unroll 1
  do {
    const Words<nw> c = read_words<nw>(sp);
    sp += nw * 16;
    const u64 x = lds_pair(pool + c.w(0));
    fma2s(sacc, c.f(1), x);
    int j = 2;
#pragma unroll
    for (int t = 0; t < kPatchNodes; ++t)
      if ((MASK >> t) & 1) {
        fma2s(acc[t].gI, c.f(j), x);
        fma2s(acc[t].gJ, c.f(j + 1), x);
        j += 2;
      }
  } while (sp != end);
}
#define FEO_MASK_SWITCH(F, ...)                                                                                              \
  switch (mask) {                                                                                                            \
    case 0: F<0>(__VA_ARGS__); break;   case 1: F<1>(__VA_ARGS__); break;   case 2: F<2>(__VA_ARGS__); break;                  \
    case 3: F<3>(__VA_ARGS__); break;   case 4: F<4>(__VA_ARGS__); break;   case 5: F<5>(__VA_ARGS__); break;                  \
    case 6: F<6>(__VA_ARGS__); break;   case 7: F<7>(__VA_ARGS__); break;   case 8: F<8>(__VA_ARGS__); break;                  \
    case 9: F<9>(__VA_ARGS__); break;   case 10: F<10>(__VA_ARGS__); break; case 11: F<11>(__VA_ARGS__); break;               \
    case 12: F<12>(__VA_ARGS__); break; case 13: F<13>(__VA_ARGS__); break; case 14: F<14>(__VA_ARGS__); break;               \
    default: F<15>(__VA_ARGS__); break;                                                                                      \
  }

struct Bars {
  uint32_t full, done;
};

// (segment, slab) of the item gridDim.x further
__device__ __forceinline__ void next_item(const PatchParams& p, int& seg, int& slab) {
  seg += p.g_div;
  slab += p.g_mod;
  if (slab >= p.n_slabs) {
    slab -= p.n_slabs;
    ++seg;
  }
}

// ---- producer warps ------------------------------------------------------------------------------
// Producer pw of P copies the lines pw * 2 + {0, 1}, (pw + P) * 2 + {0, 1}, ... of a round's load list: one LDGSTS moves two
// lines (16 lanes x 16 B each).  The list of round g + 1 rides in the stream region of round g (right after the warp
// table), so its descriptors are warp-uniform-per-half shared-memory reads; only the first round of an item reads its
// list from global memory.
template <bool BWD>
__device__ __forceinline__ void copy_line(const PatchParams& p, uint32_t sb, uint32_t ds, uint32_t slot, const float* base0, const float* base1,
                                          uint32_t src_bytes, int piece) {
  const float* src = ((BWD && (ds >> 31)) ? base1 : base0) + (int64_t)(ds & 0x7fffffffu) * p.ldb;
  ldgsts16(sb + (slot << 8) + (uint32_t)piece * 16u, src, src_bytes);
}
template <bool BWD>
__device__ __forceinline__ void produce(const PatchParams& p, uint32_t sb, const Bars& bars, int pw, int lane) {
  int seg = (int)blockIdx.x / p.n_slabs, slab = (int)blockIdx.x - seg * p.n_slabs;
  uint32_t g = 0;
  const int half = lane >> 4, piece = lane & 15;
  const uint32_t tbl_bytes = (uint32_t)((p.warps + 1) / 2) * 16u;
  const int stride = p.producers * 2;
  for (int u = (int)blockIdx.x; u < p.n_items; u += (int)gridDim.x, next_item(p, seg, slab)) {
    const int r0 = __ldg(p.seg_ptr + seg), r1 = __ldg(p.seg_ptr + seg + 1);
    const int col0 = slab * kSlab + piece * 4;
    const uint32_t src_bytes = col0 < p.ldb ? 16u : 0u;
    const float* base0 = p.src0 + (src_bytes ? col0 : 0);
    const float* base1 = (BWD ? p.src1 : p.src0) + (src_bytes ? col0 : 0);
    int4 R = __ldg(reinterpret_cast<const int4*>(p.rounds + r0));  // {load_begin, n_loads, stream_begin, n_words}
    for (int rho = r0; rho < r1; ++rho, ++g) {
      const uint32_t b = g & 1u;
      int4 Rn = R;
      if (rho + 1 < r1) Rn = __ldg(reinterpret_cast<const int4*>(p.rounds + rho + 1));
      // round g - 2 (same barrier) must be released; an item starts with an empty pool, so its first round also waits
      // for the last round of the previous item
      if (g >= 2) mbar_wait(bars.done + b * 8, ((g >> 1) - 1u) & 1u);
      const uint32_t full = bars.full + b * 8;
      if (rho == r0) {
        if (g >= 1) mbar_wait(bars.done + (b ^ 1u) * 8, ((g - 1u) >> 1) & 1u);
        if (pw == 0 && lane == 0) {
          mbar_expect_tx(full, (uint32_t)R.w * 16u);
          bulk_copy(sb + p.stream_off + b * p.stream_cap, p.stream + (size_t)R.z * 4, (uint32_t)R.w * 16u, full);
        }
        if (p.debug != 2)
          for (int l0 = pw * 2; l0 < R.y; l0 += stride * 16) {  // 16 line pairs of this producer per batch
            const int mine = l0 + (lane >> 1) * stride + (lane & 1);  // lane 2q + h holds the descriptor of pair q, line h
            uint2 d = make_uint2(0u, 0u);
            if (mine < R.y) d = __ldg(reinterpret_cast<const uint2*>(p.loads + R.x + mine));
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
              const uint32_t ds = __shfl_sync(0xffffffffu, d.x, 2 * q + half), slot = __shfl_sync(0xffffffffu, d.y, 2 * q + half);
              if (l0 + q * stride + half < R.y) copy_line<BWD>(p, sb, ds, slot, base0, base1, src_bytes, piece);
            }
          }
      } else {
        // the list is in the stream of round g - 1, which must have landed
        mbar_wait(bars.full + (b ^ 1u) * 8, ((g - 1u) >> 1) & 1u);
        if (pw == 0 && lane == 0) {
          mbar_expect_tx(full, (uint32_t)R.w * 16u);
          bulk_copy(sb + p.stream_off + b * p.stream_cap, p.stream + (size_t)R.z * 4, (uint32_t)R.w * 16u, full);
        }
        if (p.debug != 2) {
          const uint32_t list = sb + p.stream_off + (b ^ 1u) * p.stream_cap + tbl_bytes;
#pragma unroll 4
          for (int l = pw * 2 + half; l < R.y; l += stride) {
            const U2 d = lds_unit(list + (uint32_t)l * 8u);
            copy_line<BWD>(p, sb, d.lo, d.hi, base0, base1, src_bytes, piece);
          }
        }
      }
      ldgsts_arrive(full);
      R = Rn;
    }
  }
}

// ---- the kernel --------------------------------------------------------------------------------------
template <bool BWD, int NT>
__global__ void __launch_bounds__(NT, 1) residual_patch_kernel(const PatchParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sb = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Bars bars;
  bars.full = sb + p.bar_off;
  bars.done = bars.full + 16;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(bars.full + k * 8, 1u + 32u * (uint32_t)p.producers);
      mbar_init(bars.done + k * 8, (uint32_t)p.warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= p.warps) {
    produce<BWD>(p, sb, bars, warp - p.warps, lane);
    return;
  }
  const uint32_t pool = sb + (uint32_t)lane * 8u;
  const bool precond = p.precond != 0;
  double dsum = 0.0;
  float g2 = 2.0f;
  if (BWD && p.grad_loss != nullptr) g2 = 2.0f * __ldg(p.grad_loss);
  int seg = (int)blockIdx.x / p.n_slabs, slab = (int)blockIdx.x - seg * p.n_slabs;
  uint32_t g = 0;
  for (int u = (int)blockIdx.x; u < p.n_items; u += (int)gridDim.x, next_item(p, seg, slab)) {
    const int n_rounds = __ldg(p.seg_ptr + seg + 1) - __ldg(p.seg_ptr + seg);
    const int b0 = slab * kSlab + lane * 2;  // this lane's two samples
    const bool in_ld = b0 < p.ldb;
    for (int rr = 0; rr < n_rounds; ++rr, ++g) {
      const uint32_t b = g & 1u;
      mbar_wait(bars.full + b * 8, (g >> 1) & 1u);
      if (p.debug != 1) {
        const uint32_t sbuf = sb + p.stream_off + b * p.stream_cap;
        const U2 tbl = lds_unit(sbuf + (uint32_t)warp * 8u);
        uint32_t sp = sbuf + tbl.lo;
        float lsum = 0.f;
        for (uint32_t ip = 0; ip < tbl.hi; ++ip) {
          // ---- patch header ----
          const Words<BWD ? 1 : kPatchHeaderWords> H = read_words<BWD ? 1 : kPatchHeaderWords>(sp);
          sp += kPatchHeaderWords * 16;
          const int n_runs = (int)(H.w(0) & 0xfffu);
          const uint32_t hp = sp - kPatchHeaderWords * 16;  // the epilogue reads the header again (its registers are not kept)
          P2 sacc = {0.f, 0.f};
          if (!BWD) {
            // the load vectors of the patch's rows are requested now and used in the epilogue
            float2 fI[kPatchNodes], fJ[kPatchNodes], fS = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < kPatchNodes; ++t) {
              fI[t] = fJ[t] = make_float2(0.f, 0.f);
              const int dI = (int)H.w(4 + 4 * t), dJ = (int)H.w(5 + 4 * t);
              if (dI >= 0 && in_ld) {
                fI[t] = ldg2_stream(p.fT + (int64_t)dI * p.ldb + b0);
                fJ[t] = ldg2_stream(p.fT + (int64_t)dJ * p.ldb + b0);
              }
            }
            if (((H.w(0) >> 12) & 1u) && in_ld) fS = ldg2_stream(p.fT + (int64_t)(int)H.w(1) * p.ldb + b0);
            FwdAcc acc[kPatchNodes];
#pragma unroll
            for (int t = 0; t < kPatchNodes; ++t) acc[t].aI = acc[t].uI = acc[t].vI = acc[t].aJ = acc[t].uJ = acc[t].vJ = P2{0.f, 0.f};
#pragma unroll 1
            for (int r = 0; r < n_runs; ++r) {
              const uint32_t meta = lds_unit(sp).lo;
              sp += 16;
              const int mask = (int)((meta >> 4) & 15u), count = (int)(meta >> 8);
              if ((meta & 15u) == 0u) {
                FEO_MASK_SWITCH(fwd_pair_run, acc, sacc, sp, count, pool)
              } else {
                FEO_MASK_SWITCH(fwd_plain_run, acc, sacc, sp, count, pool)
              }
            }
            // ---- epilogue: r = A - (F - c) / A - (-F + c), loss partial ----
#pragma unroll
            for (int t = 0; t < kPatchNodes; ++t) {
              const uint4 hw = lds_word(hp + 16 * (1 + t));  // {dof I, dof J, off(own I), off(own J)}
              const int dI = (int)hw.x, dJ = (int)hw.y;
              if (dI < 0) continue;
              float d1[2], d2[2];
              unpk(lds_pair(pool + hw.z), d1[0], d1[1]);
              unpk(lds_pair(pool + hw.w), d2[0], d2[1]);
              float2 rI, rJ;
              rI.x = resid1(acc[t].aI.lo, fI[t].x, conv1(d1[0], acc[t].uI.lo, d2[0], acc[t].vI.lo), precond);
              rI.y = resid1(acc[t].aI.hi, fI[t].y, conv1(d1[1], acc[t].uI.hi, d2[1], acc[t].vI.hi), precond);
              rJ.x = resid1(acc[t].aJ.lo, fJ[t].x, conv1(d1[0], acc[t].uJ.lo, d2[0], acc[t].vJ.lo), precond);
              rJ.y = resid1(acc[t].aJ.hi, fJ[t].y, conv1(d1[1], acc[t].uJ.hi, d2[1], acc[t].vJ.hi), precond);
              if (b0 < p.B) lsum = __fmaf_rn(rJ.x, rJ.x, __fmaf_rn(rI.x, rI.x, lsum));
              if (b0 + 1 < p.B) lsum = __fmaf_rn(rJ.y, rJ.y, __fmaf_rn(rI.y, rI.y, lsum));
              if (p.outT != nullptr && in_ld) {
                *reinterpret_cast<float2*>(p.outT + (int64_t)dI * p.ldb + b0) = rI;
                *reinterpret_cast<float2*>(p.outT + (int64_t)dJ * p.ldb + b0) = rJ;
              }
            }
            const U2 h0 = lds_unit(hp);
            if ((h0.lo >> 12) & 1u) {
              const int ds = (int)h0.hi;
              float2 r;
              r.x = resid1(sacc.lo, fS.x, 0.f, precond);
              r.y = resid1(sacc.hi, fS.y, 0.f, precond);
              if (b0 < p.B) lsum = __fmaf_rn(r.x, r.x, lsum);
              if (b0 + 1 < p.B) lsum = __fmaf_rn(r.y, r.y, lsum);
              if (p.outT != nullptr && in_ld) *reinterpret_cast<float2*>(p.outT + (int64_t)ds * p.ldb + b0) = r;
            }
          } else {
            BwdAcc acc[kPatchNodes];
#pragma unroll
            for (int t = 0; t < kPatchNodes; ++t) acc[t].gI = acc[t].gJ = acc[t].b1I = acc[t].b2I = acc[t].b1J = acc[t].b2J = P2{0.f, 0.f};
#pragma unroll 1
            for (int r = 0; r < n_runs; ++r) {
              const uint32_t meta = lds_unit(sp).lo;
              sp += 16;
              const int mask = (int)((meta >> 4) & 15u), count = (int)(meta >> 8);
              if ((meta & 15u) == 0u) {
                FEO_MASK_SWITCH(bwd_pair_run, acc, sacc, sp, count, pool)
              } else {
                FEO_MASK_SWITCH(bwd_plain_run, acc, sacc, sp, count, pool)
              }
            }
            // ---- epilogue: E-term of the own rows, scale by 2 g ----
#pragma unroll
            for (int t = 0; t < kPatchNodes; ++t) {
              const uint4 hw = lds_word(hp + 16 * (1 + t));
              const int dI = (int)hw.x, dJ = (int)hw.y;
              if (dI < 0) continue;
              float rI[2], rJ[2];
              unpk(lds_pair(pool + hw.z), rI[0], rI[1]);
              unpk(lds_pair(pool + hw.w), rJ[0], rJ[1]);
              float2 oI, oJ;
              oI.x = acc[t].gI.lo + p.esign * (acc[t].b1I.lo * rI[0] + acc[t].b1J.lo * rJ[0]);
              oI.y = acc[t].gI.hi + p.esign * (acc[t].b1I.hi * rI[1] + acc[t].b1J.hi * rJ[1]);
              oJ.x = acc[t].gJ.lo + p.esign * (acc[t].b2I.lo * rI[0] + acc[t].b2J.lo * rJ[0]);
              oJ.y = acc[t].gJ.hi + p.esign * (acc[t].b2I.hi * rI[1] + acc[t].b2J.hi * rJ[1]);
              if (in_ld) {
                *reinterpret_cast<float2*>(p.outT + (int64_t)dI * p.ldb + b0) = make_float2(oI.x * g2, oI.y * g2);
                *reinterpret_cast<float2*>(p.outT + (int64_t)dJ * p.ldb + b0) = make_float2(oJ.x * g2, oJ.y * g2);
              }
            }
            const U2 h0 = lds_unit(hp);
            if (((h0.lo >> 12) & 1u) && in_ld)
              *reinterpret_cast<float2*>(p.outT + (int64_t)(int)h0.hi * p.ldb + b0) = make_float2(sacc.lo * g2, sacc.hi * g2);
          }
        }
        dsum += (double)lsum;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars.done + b * 8);
    }
  }
  if (!BWD) {
    dsum = warp_sum(dsum);
    if (lane == 0) p.partials[(size_t)blockIdx.x * p.warps + warp] = (float)dsum;
  }
}

int check_layout(const void* ptr, int64_t ld, int32_t B, const char* what) {
  if (ptr == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " is NULL");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " not 16-byte aligned");
  if (ld % 4 != 0 || ld < ((B + 3) / 4) * 4)
    return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + ": ldb must be a multiple of 4 and >= ceil4(B)");
  return FEO_OK;
}

int debug_mode() {
  const char* s = std::getenv("FEO_DEBUG_MODE");
  return s != nullptr ? atoi(s) : 0;
}

template <bool BWD>
int launch(const DevPatchPlan& P, PatchParams p, int grid, cudaStream_t st) {
  const int nt = (P.warps + P.producers) * 32;
  p.stream_off = (uint32_t)P.pool_lines * kLineBytes;
  p.stream_off = (p.stream_off + 1023u) / 1024u * 1024u;
  p.bar_off = p.stream_off + 2u * (uint32_t)P.stream_cap;
  const uint32_t total = p.bar_off + 64u;
  if (total > kSmemMax) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan needs more shared memory than an SM has");
#define FEO_LAUNCH(NT)                                                                                                   \
  do {                                                                                                                   \
    FEO_CUDA_CHECK(cudaFuncSetAttribute(residual_patch_kernel<BWD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total)); \
    residual_patch_kernel<BWD, NT><<<(unsigned)grid, (unsigned)nt, total, st>>>(p);                                       \
  } while (0)
  // the register file is split over the four SM sub-partitions: 3 / 4 / 5 / 6 warps each -> 168 / 128 / 96 / 80 registers
  if (nt <= 384)
    FEO_LAUNCH(384);
  else if (nt <= 512)
    FEO_LAUNCH(512);
  else if (nt <= 640)
    FEO_LAUNCH(640);
  else if (nt <= 768)
    FEO_LAUNCH(768);
  else
    return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan has more warps than the kernels are built for");
#undef FEO_LAUNCH
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

}  // namespace

int launch_patch_fwd(const feo_operator* op, const DevPatchPlan& P, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                     float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  if (int rc = check_layout(fT, ldb, B, "fT")) return rc;
  if (rT != nullptr)
    if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (loss_out == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "loss_out is NULL");
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)P.n_segments * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "too many work items");
  int sms = 1;
  if (int rc = sm_count(&sms)) return rc;
  const int grid = (int)std::min<int64_t>(count, sms);
  const size_t n_partials = (size_t)grid * P.warps;
  if (ws == nullptr || ws_bytes < n_partials * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  PatchParams p{};
  p.seg_ptr = P.seg_ptr;
  p.rounds = P.rounds;
  p.loads = P.loads;
  p.stream = P.stream;
  p.src0 = alphaT;
  p.src1 = alphaT;
  p.fT = fT;
  p.outT = rT;
  p.partials = (float*)ws;
  p.ldb = ldb;
  p.B = B;
  p.n_slabs = n_slabs;
  p.n_items = (int32_t)count;
  p.g_div = grid / n_slabs;
  p.g_mod = grid % n_slabs;
  p.warps = P.warps;
  p.producers = P.producers;
  p.stream_cap = (uint32_t)P.stream_cap;
  p.debug = debug_mode();
  p.precond = op->ns_branch;
  if (int rc = launch<false>(P, p, grid, st)) return rc;
  return finalize_loss((float*)ws, (int)n_partials, 1.0f, loss_out, st);
}

int launch_patch_bwd(const feo_operator* op, const DevPatchPlan& P, const float* alphaT, const float* rT, const float* grad_loss,
                     float* gradT, int64_t ldb, int32_t B, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (int rc = check_layout(gradT, ldb, B, "gradT")) return rc;
  if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)P.n_segments * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "too many work items");
  int sms = 1;
  if (int rc = sm_count(&sms)) return rc;
  const int grid = (int)std::min<int64_t>(count, sms);
  PatchParams p{};
  p.seg_ptr = P.seg_ptr;
  p.rounds = P.rounds;
  p.loads = P.loads;
  p.stream = P.stream;
  p.src0 = rT;
  p.src1 = alphaT;
  p.outT = gradT;
  p.grad_loss = grad_loss;
  p.ldb = ldb;
  p.B = B;
  p.n_slabs = n_slabs;
  p.n_items = (int32_t)count;
  p.g_div = grid / n_slabs;
  p.g_mod = grid % n_slabs;
  p.warps = P.warps;
  p.producers = P.producers;
  p.stream_cap = (uint32_t)P.stream_cap;
  p.debug = debug_mode();
  p.precond = op->ns_branch;
  p.esign = op->ns_branch ? 1.0f : -1.0f;
  return launch<true>(P, p, grid, st);
}

}  // namespace feo
