// sm_100a kernels of the fused sparse residual, lattice formulation (plan: feo_lattice.h, feo_lattice_plan.cpp; the cell
// bodies are generated: tools/gen_lattice_stencil.py -> feo_lattice_gen.inc).
//
// One persistent CTA per SM sweeps STRIPS of W cells upwards through the lattice, one slab of 64 samples at a time:
//   * shared memory = a ring of R lattice rows; a row slot holds the contiguous dof run of the strip's nodes (2W + 3 lattice
//     columns: the strip and its halo) as 256-byte lines (64 samples of one dof); backward: the r run and the alpha run;
//   * the producer warp stages the two lattice rows that enter the 5-row window of the next step(s) with ONE 2-D TMA box per
//     row and source array (cp.async.bulk.tensor, completion on the step's `full` mbarrier); rows above / below the lattice
//     are fetched from out-of-bounds coordinates, i.e. zero filled;
//   * consumer warp w evaluates cell (strip * W + w, cj) for the 64 samples, 2 samples per lane: 45 (forward) / 83 (backward)
//     conflict-free LDS.64 gathers feed all 9 rows / columns of the cell; every accumulator lives in registers; the
//     arithmetic is packed fp32 (FFMA2) whose coefficient operand comes straight from the constant bank (kernel parameters:
//     one table per cell class, class 0 -- the interior -- at compile-time offsets, the boundary classes indexed);
//   * row-/column-owned with a fixed summation order: no atomics, bit-reproducible on a given device.
// Work = (strip, slab, cj) steps, cut into gridDim.x equal contiguous chunks.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "feo_lattice.h"
#include "feo_lattice_gen.inc"

namespace feo {
int make_box_map(const float* base, int64_t ldb, int32_t n, int32_t box_rows, CUtensorMap* out);  // feo_tiled.cu
int sm_count(int* out);                                                                            // feo_tiled.cu

namespace {

typedef unsigned long long u64;
constexpr uint32_t kSmemMax = 232448;
constexpr int kStagesMax = 4;

struct LatMaps {
  CUtensorMap m[2][2];  // [source array][row parity]: box = 64 samples x (5W + 8 | 4W + 6) dofs
  CUtensorMap f[2];     // forward: load vectors, [row parity]: box = 64 samples x (5W | 4W) dofs -- the strip's own rows (L2 prefetch)
};

template <int NCOEF>
struct LatParams {
  const float* fT;         // forward: load vectors
  float* outT;             // forward: rT (may be NULL) ; backward: gradT
  float* partials;         // forward: one loss partial per (CTA, consumer warp)
  const float* grad_loss;  // backward: upstream gradient (NULL = 1)
  int64_t ldb;
  int32_t B, n_slabs;
  int32_t n, nc, N;        // mesh cells per side, cell origins per side, dofs
  int32_t W, n_strips;     // consumer warps = cells per strip, strips
  int32_t total_steps;     // n_strips * n_slabs * nc
  int32_t R, D, K;         // ring rows (7 or 9); a step's new rows reuse the slots step - D released, D = (R - 3) / 2; K = D + 1 barrier stages
  int32_t he, ho;          // lines of an even / odd lattice row run
  uint32_t slot_bytes, bar_off;
  int32_t debug;           // FEO_DEBUG_MODE: 1 = staging only, 2 = compute only (results are garbage)
  int32_t sync_mask;       // bit 0: release a step's oldest rows early, bit 1: wait for its newest rows late (developer knob FEO_LAT_SYNC)
  int32_t precond;         // forward: 1 -> r = lhs - (f - c), 0 -> r = lhs - (-f + c)
  float esign;             // backward: +1 precond branch, -1 otherwise
  uint8_t cat_cls[25];     // class of a cell by the boundary-layer categories of (cj, ci): lat_cat(cj) * 5 + lat_cat(ci)
  uint8_t exist[kLatMaxClasses];
  float tab[kLatMaxClasses * NCOEF];
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
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ u64 lds64(uint32_t a) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// acc += c * x on two packed fp32 lanes, c a scalar (a constant-bank / uniform-register operand of FFMA2)
__device__ __forceinline__ void fmac(u64& acc, float c, u64 x) { acc = fma2(pk(c, c), x, acc); }
__device__ __forceinline__ u64 ldg_pair_stream(const float* p) {  // two samples of a load-vector line, read once
  u64 v;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_pair(float* p, u64 v) { asm volatile("st.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Epilogue arithmetic on sample pairs (FEONet_steady_Navier-Stokes/train_FEONet.py:317-330, :356):
//   c = u_i Bu1 + u_j Bu2       -> fl(fma(u_i, Bu1, fl(u_j Bu2)))   (ptxas contracts the packed multiply-add: one rounding less
//                                                                     than the reference's two products and an add)
//   r = LHS - (F - c)  (precond branch)  /  LHS - (-F + c) = LHS + (F - c)  (else): one rounded F - c, one rounded LHS -/+ it
__device__ __forceinline__ u64 conv2(u64 d1, u64 s1, u64 d2, u64 s2) { return fma2(d1, s1, mul2(d2, s2)); }
__device__ __forceinline__ u64 resid2(u64 lhs, u64 f, u64 c, u64 neg_sign) { return fma2(neg_sign, sub2(f, c), lhs); }

// ---- the walk over a CTA's chunk of (column, cj) steps; column = (strip, slab) ---------------------------
struct Walk {
  int32_t s, s_end, strip, slab, cj, local;
  int32_t st, ph;  // barrier stage of the step (step index mod K) and its phase parity
  template <typename P>
  __device__ __forceinline__ void start(const P& p) {
    s = (int32_t)(((int64_t)blockIdx.x * p.total_steps) / (int64_t)gridDim.x);
    s_end = (int32_t)(((int64_t)(blockIdx.x + 1) * p.total_steps) / (int64_t)gridDim.x);
    const int32_t col = s / p.nc;
    cj = s - col * p.nc;
    strip = col / p.n_slabs;
    slab = col - strip * p.n_slabs;
    local = 0;
    st = 0;
    ph = 0;
  }
  __device__ __forceinline__ bool done() const { return s >= s_end; }
  template <typename P>
  __device__ __forceinline__ void next(const P& p) {
    ++s;
    ++cj;
    ++local;
    if (cj == p.nc) {
      cj = 0;
      local = 0;
      if (++slab == p.n_slabs) {
        slab = 0;
        ++strip;
      }
    }
    if (++st == p.K) {
      st = 0;
      ph ^= 1;
    }
  }
};

// dof of u1 at node (0, y) in the interleaved lattice numbering; rows off the lattice map past the end (TMA zero fill)
template <typename P>
__device__ __forceinline__ int32_t row_dof0(const P& p, int y) {
  if (y < 0 || y > 2 * p.n) return p.N + 4096;
  return (y >> 1) * (9 * p.n + 5) + (y & 1) * (5 * p.n + 3);
}

// ---- producer warp ---------------------------------------------------------------------------------------
// Step g of a CTA receives its rows on full[g % K] and releases them on done[g % K].  A step's two new rows go to the ring
// slots that step g - D released; consumers release the two oldest rows of a step EARLY (after part A of the body, see
// fwd_cell / bwd_cell) unless the step ends a segment, and wait for the new rows LATE (before part C), so a fetch has about
// two step times of slack with a 7-row ring.
__device__ __forceinline__ void tma_prefetch_l2(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
template <bool BWD, typename P>
__device__ __forceinline__ void produce(const LatMaps& maps, const P& p, uint32_t sb, uint32_t full, uint32_t done, int lane) {
  if (p.debug == 2) return;
  Walk w;
  w.start(p);
  const uint32_t bytes_e = (uint32_t)p.he * kLineBytes * (BWD ? 2u : 1u), bytes_o = (uint32_t)p.ho * kLineBytes * (BWD ? 2u : 1u);
  int32_t st_prev = 0, ph_prev = 0;
  uint32_t slot0 = 0;  // ring slot of the window's first row = (2 * local) mod R
  for (int32_t g = 0; !w.done(); ++g, st_prev = w.st, ph_prev = w.ph, w.next(p)) {
    if (w.local == 0) {
      slot0 = 0;
      if (g > 0) mbar_wait(done + (uint32_t)st_prev * 8, (uint32_t)ph_prev);  // the new segment overwrites the whole ring
    } else {
      slot0 += 2;
      if (slot0 >= (uint32_t)p.R) slot0 -= (uint32_t)p.R;
      if (w.local >= p.D) {
        // step g - D = g + 1 - K used the stage of step g + 1 one phase earlier
        int32_t st_n = w.st + 1, ph_n = w.ph;
        if (st_n == p.K) {
          st_n = 0;
          ph_n ^= 1;
        }
        mbar_wait(done + (uint32_t)st_n * 8, (uint32_t)ph_n ^ 1u);
      }
    }
    if (lane == 0) {
      const int32_t c0 = w.slab * kSlab, ci0 = w.strip * p.W;
      // the segment's first step brings its whole window (rows 2cj-2 .. 2cj+2 -> slots 0..4), later steps the two new rows
      const int first = w.local == 0 ? 0 : 3;
      const uint32_t bar = full + (uint32_t)w.st * 8;
      // row 2cj - 2 + i has the parity of i: rows 0..4 = three even and two odd ones, rows 3..4 = one of each
      mbar_expect_tx(bar, first == 0 ? 3u * bytes_e + 2u * bytes_o : bytes_e + bytes_o);
      for (int i = first; i < 5; ++i) {
        const int y = 2 * w.cj - 2 + i, odd = i & 1;
        uint32_t slot = slot0 + (uint32_t)i;
        if (slot >= (uint32_t)p.R) slot -= (uint32_t)p.R;
        const int32_t d0 = row_dof0(p, y) + (ci0 - 1) * (odd ? 4 : 5);
        const uint32_t dst = sb + slot * p.slot_bytes;
        tma_box(dst, &maps.m[0][odd], c0, d0, bar);
        if (BWD) tma_box(dst + (uint32_t)p.he * kLineBytes, &maps.m[1][odd], c0, d0, bar);
      }
      if (!BWD) {  // the load vectors of the step's own rows are wanted in its epilogue: bring them to L2 now
        tma_prefetch_l2(&maps.f[0], c0, row_dof0(p, 2 * w.cj) + 5 * ci0);
        tma_prefetch_l2(&maps.f[1], c0, row_dof0(p, 2 * w.cj + 1) + 4 * ci0);
      }
    }
    __syncwarp();
  }
}

// ---- cell bodies -------------------------------------------------------------------------------------------
// byte offset of the line of (node dx, comp) within the row run, relative to the warp's cell (dx = -2 .. 2)
__device__ __forceinline__ constexpr uint32_t lat_off(int dx, int dy, int comp) {
  return (uint32_t)(lat_pos(dx + 2, dy & 1) + comp) * (uint32_t)kLineBytes;
}
// targets 0..3 = V (0,0), H (1,0), T (0,1), D (1,1)
__host__ __device__ constexpr int tgt_x(int t) { return t & 1; }
__host__ __device__ constexpr int tgt_y(int t) { return t >> 1; }

#define BEGIN {
#define END }
#define C(i) (FAST ? p.tab[(i)] : p.tab[cbase + (i)])

// what a consumer warp does between the parts of a cell body
struct StepSync {
  uint32_t full_bar, full_ph, done_bar;
  bool wait_late;      // wait for the step's new rows before part C (else they were waited for at the start)
  bool release_early;  // release the step's two oldest rows after part A (else at the end of the step)
  __device__ __forceinline__ void after_a(int lane) const {
    if (release_early) {
      __syncwarp();
      if (lane == 0) mbar_arrive(done_bar);
    }
  }
  __device__ __forceinline__ void before_c() const {
    if (wait_late) mbar_wait(full_bar, full_ph);
  }
};

// forward: r = A a -/+ (F - c) for the 9 rows of a cell, loss partial
template <bool FAST, bool ELEM, typename P>
__device__ __forceinline__ void fwd_cell(const P& p, const StepSync& sy, int lane, const uint32_t (&Bx)[5], int cbase, uint32_t ex,
                                         int64_t oE, int64_t oO, int b0, u64 neg_sign, float& lsum) {
  // element offsets of the cell's rows in the dof-major arrays (sample b0 included): V = (oE, oE + ldb), P = oE + 2 ldb,
  // H = oE + 3 ldb, T = oO, D = oO + 2 ldb
  const bool in_ld = b0 < p.ldb, in_b = b0 < p.B, in_b1 = b0 + 1 < p.B;
  const int64_t offs[kLatTargets] = {oE, oE + 3 * p.ldb, oO, oO + 2 * p.ldb};
  u64 fv[kLatTargets][2], fp = 0ull;
#pragma unroll
  for (int t = 0; t < kLatTargets; ++t) fv[t][0] = fv[t][1] = 0ull;
  if (in_ld) {
#pragma unroll
    for (int t = 0; t < kLatTargets; ++t)
      if (FAST || ((ex >> t) & 1u)) {
        fv[t][0] = ldg_pair_stream(p.fT + offs[t]);
        fv[t][1] = ldg_pair_stream(p.fT + offs[t] + p.ldb);
      }
    fp = ldg_pair_stream(p.fT + oE + 2 * p.ldb);
  }
  u64 acc[3][kLatTargets][2], sacc = 0ull;
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int t = 0; t < kLatTargets; ++t) acc[m][t][0] = acc[m][t][1] = 0ull;
#define LDX(v, dx, dy, comp) const u64 v = lds64(Bx[(dy) + 2] + lat_off(dx, dy, comp));
#define FV(mat, t, i) fmac(acc[mat][t][0], C(i), xI); fmac(acc[mat][t][1], C(i), xJ);
#define FSI(i) fmac(sacc, C(i), xI);
#define FSJ(i) fmac(sacc, C(i), xJ);
#define FP(t, tc, i) fmac(acc[0][t][tc], C(i), xP);
#define FSP(i) fmac(sacc, C(i), xP);
  // ELEM: the same rows as a matrix-free element walk in gather form (one FMA per element-level entry; A/B variant)
  if constexpr (ELEM) {
    FEO_LAT_FWDE_BODY_A
  } else {
    FEO_LAT_FWD_BODY_A
  }
  sy.after_a(lane);
  if constexpr (ELEM) {
    FEO_LAT_FWDE_BODY_B
  } else {
    FEO_LAT_FWD_BODY_B
  }
  sy.before_c();
  if constexpr (ELEM) {
    FEO_LAT_FWDE_BODY_C
  } else {
    FEO_LAT_FWD_BODY_C
  }
#undef LDX
#undef FV
#undef FSI
#undef FSJ
#undef FP
#undef FSP
  auto finish = [&](u64 r, int64_t off) {
    float r0, r1;
    unpk(r, r0, r1);
    if (in_b) lsum = fmaf(r0, r0, lsum);
    if (in_b1) lsum = fmaf(r1, r1, lsum);
    if (p.outT != nullptr && in_b) stg_pair(p.outT + off, r);
  };
#pragma unroll
  for (int t = 0; t < kLatTargets; ++t) {
    if (!FAST && !((ex >> t) & 1u)) continue;
    const u64 d1 = lds64(Bx[tgt_y(t) + 2] + lat_off(tgt_x(t), tgt_y(t), 0)), d2 = lds64(Bx[tgt_y(t) + 2] + lat_off(tgt_x(t), tgt_y(t), 1));
#pragma unroll
    for (int c = 0; c < 2; ++c)
      finish(resid2(acc[0][t][c], fv[t][c], conv2(d1, acc[1][t][c], d2, acc[2][t][c]), neg_sign), offs[t] + (c ? p.ldb : 0));
  }
  finish(resid2(sacc, fp, 0ull, neg_sign), oE + 2 * p.ldb);
}

// backward: grad = 2 g [A^T r + s (B1^T (d1 r) + B2^T (d2 r) + E-term)] for the 9 columns of a cell
template <bool FAST, typename P>
__device__ __forceinline__ void bwd_cell(const P& p, const StepSync& sy, int lane, const uint32_t (&Br)[5], const uint32_t (&Ba)[5], int cbase,
                                         uint32_t ex, int64_t oE, int64_t oO, int b0, float g2) {
  const int64_t offs[kLatTargets] = {oE, oE + 3 * p.ldb, oO, oO + 2 * p.ldb};
  u64 g[kLatTargets][2], bu[3][kLatTargets][2], sacc = 0ull;  // bu[1] = Bu1, bu[2] = Bu2 of the own rows ([0] unused)
#pragma unroll
  for (int t = 0; t < kLatTargets; ++t) {
    g[t][0] = g[t][1] = 0ull;
    bu[1][t][0] = bu[1][t][1] = bu[2][t][0] = bu[2][t][1] = 0ull;
  }
#define LDR(v, dx, dy, comp) const u64 v = lds64(Br[(dy) + 2] + lat_off(dx, dy, comp));
#define LDA(v, dx, dy, comp) const u64 v = lds64(Ba[(dy) + 2] + lat_off(dx, dy, comp));
#define BTA(t, ia) fmac(g[t][0], C(ia), rI); fmac(g[t][1], C(ia), rJ);
#define BTB(t, ia, ib1, ib2)                                        \
  {                                                                 \
    u64 T = 0ull;                                                   \
    if ((ia) >= 0) {                                                \
      const float ca = C((ia) >= 0 ? (ia) : 0);                     \
      T = pk(ca, ca);                                               \
    }                                                               \
    if ((ib1) >= 0) fmac(T, C((ib1) >= 0 ? (ib1) : 0), d1);         \
    if ((ib2) >= 0) fmac(T, C((ib2) >= 0 ? (ib2) : 0), d2);         \
    g[t][0] = fma2(rI, T, g[t][0]);                                 \
    g[t][1] = fma2(rJ, T, g[t][1]);                                 \
  }
#define BF(mat, t, i) fmac(bu[mat][t][0], C(i), d1); fmac(bu[mat][t][1], C(i), d2);
#define BSI(i) fmac(sacc, C(i), rI);
#define BSJ(i) fmac(sacc, C(i), rJ);
#define BP(t, tc, i) fmac(g[t][tc], C(i), rP);
#define BSP(i) fmac(sacc, C(i), rP);
  FEO_LAT_BWD_BODY_A
  sy.after_a(lane);
  FEO_LAT_BWD_BODY_B
  sy.before_c();
  FEO_LAT_BWD_BODY_C
#undef LDR
#undef LDA
#undef BTA
#undef BTB
#undef BF
#undef BSI
#undef BSJ
#undef BP
#undef BSP
  const bool st = b0 < p.B;
  const u64 es2 = pk(p.esign, p.esign), g22 = pk(g2, g2);
#pragma unroll
  for (int t = 0; t < kLatTargets; ++t) {
    if (!FAST && !((ex >> t) & 1u)) continue;
    const u64 rI = lds64(Br[tgt_y(t) + 2] + lat_off(tgt_x(t), tgt_y(t), 0)), rJ = lds64(Br[tgt_y(t) + 2] + lat_off(tgt_x(t), tgt_y(t), 1));
    // E-term of the own rows: g_I += e (Bu1_I rI + Bu1_J rJ), g_J += e (Bu2_I rI + Bu2_J rJ); scale by 2 g
    const u64 oI = mul2(fma2(es2, fma2(bu[1][t][1], rJ, mul2(bu[1][t][0], rI)), g[t][0]), g22);
    const u64 oJ = mul2(fma2(es2, fma2(bu[2][t][1], rJ, mul2(bu[2][t][0], rI)), g[t][1]), g22);
    if (st) {
      stg_pair(p.outT + offs[t], oI);
      stg_pair(p.outT + offs[t] + p.ldb, oJ);
    }
  }
  if (st) stg_pair(p.outT + oE + 2 * p.ldb, mul2(sacc, g22));
}
#undef C
#undef BEGIN
#undef END

// ---- the kernel ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int lat_ncoef(bool bwd, bool elem) { return bwd ? FEO_LAT_BWD_NCOEF : (elem ? FEO_LAT_FWDE_NCOEF : FEO_LAT_FWD_NCOEF); }

template <bool BWD, int NT, bool ELEM>
__global__ void __launch_bounds__(NT, 1)
    residual_lattice_kernel(const __grid_constant__ LatMaps maps, const __grid_constant__ LatParams<lat_ncoef(BWD, ELEM)> p) {
  constexpr int NCOEF = lat_ncoef(BWD, ELEM);
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sb = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full = sb + p.bar_off, done = full + kStagesMax * 8;
  if (threadIdx.x == 0) {
    for (int k = 0; k < p.K; ++k) {
      mbar_init(full + k * 8, 1);
      mbar_init(done + k * 8, (uint32_t)p.W);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == p.W) {
    produce<BWD>(maps, p, sb, full, done, lane);
    return;
  }
  if (warp > p.W) return;
  const bool precond = p.precond != 0;
  float g2 = 2.0f;
  if (BWD && p.grad_loss != nullptr) g2 = 2.0f * __ldg(p.grad_loss);
  double dsum = 0.0;
  const uint32_t lane_base = sb + (uint32_t)lane * 8u;
  const uint32_t cell_off[2] = {(uint32_t)warp * 5u * kLineBytes, (uint32_t)warp * 4u * kLineBytes};
  const uint32_t a_off = (uint32_t)p.he * kLineBytes;
  Walk w;
  w.start(p);
  uint32_t slot0 = 0;
  const u64 neg_sign = precond ? pk(-1.f, -1.f) : pk(1.f, 1.f);
  const int64_t row_pair = (int64_t)(9 * p.n + 5) * p.ldb;  // dofs of an (even, odd) lattice row pair = one cell row further up
  int64_t oE = 0, oO = 0;
  while (!w.done()) {
    const int32_t ci = w.strip * p.W + warp, cj = w.cj, slab = w.slab;
    const bool first = w.local == 0;
    StepSync sy;
    sy.full_bar = full + (uint32_t)w.st * 8;
    sy.full_ph = (uint32_t)w.ph;
    sy.done_bar = done + (uint32_t)w.st * 8;
    if (first) {
      slot0 = 0;
    } else {
      slot0 += 2;
      if (slot0 >= (uint32_t)p.R) slot0 -= (uint32_t)p.R;
    }
    w.next(p);
    const bool last = w.done() || w.local == 0;  // the step ends a segment: the producer may overwrite the whole ring after it
    const bool active = p.debug != 1 && ci < p.nc;
    sy.wait_late = p.debug != 2 && !first && active && (p.sync_mask & 2);
    sy.release_early = !last && active && (p.sync_mask & 1);
    if (p.debug != 2 && !sy.wait_late) mbar_wait(sy.full_bar, sy.full_ph);
    if (active) {
      uint32_t Bx[5], Ba[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        uint32_t slot = slot0 + (uint32_t)i;
        if (slot >= (uint32_t)p.R) slot -= (uint32_t)p.R;
        Bx[i] = lane_base + slot * p.slot_bytes + cell_off[i & 1];
        Ba[i] = Bx[i] + a_off;
      }
      const int b0 = slab * kSlab + lane * 2;
      if (first) {
        oE = (int64_t)(cj * (9 * p.n + 5) + 5 * ci) * p.ldb + b0;
        oO = (int64_t)(cj * (9 * p.n + 5) + (5 * p.n + 3) + 4 * ci) * p.ldb + b0;
      } else {
        oE += row_pair;
        oO += row_pair;
      }
      const uint32_t cls = p.cat_cls[lat_cat(cj, p.nc) * 5 + lat_cat(ci, p.nc)];
      if (!BWD) {
        float lsum = 0.f;
        if (cls == 0)
          fwd_cell<true, ELEM>(p, sy, lane, Bx, 0, 15u, oE, oO, b0, neg_sign, lsum);
        else
          fwd_cell<false, ELEM>(p, sy, lane, Bx, (int)cls * NCOEF, p.exist[cls], oE, oO, b0, neg_sign, lsum);
        dsum += (double)lsum;
      } else {
        if (cls == 0)
          bwd_cell<true>(p, sy, lane, Bx, Ba, 0, 15u, oE, oO, b0, g2);
        else
          bwd_cell<false>(p, sy, lane, Bx, Ba, (int)cls * NCOEF, p.exist[cls], oE, oO, b0, g2);
      }
    }
    if (!sy.release_early) {
      __syncwarp();
      if (lane == 0) mbar_arrive(sy.done_bar);
    }
  }
  if (!BWD) {
    dsum = warp_sum(dsum);
    if (lane == 0) p.partials[(size_t)blockIdx.x * p.W + warp] = (float)dsum;
  }
}

int check_layout(const void* ptr, int64_t ld, int32_t B, const char* what) {
  if (ptr == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " is NULL");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " not 16-byte aligned");
  if (ld % 4 != 0 || ld < ((B + 3) / 4) * 4)
    return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + ": ldb must be a multiple of 4 and >= ceil4(B)");
  return FEO_OK;
}

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s != nullptr ? atoi(s) : dflt;
}

// strip width: as many consumer warps as the register file / the ring allow, preferring widths that fill the last strip
int pick_width(int nc, int w_max) {
  int best = std::min(w_max, nc), best_waste = 1 << 30;
  for (int w = std::min(w_max, nc); w >= std::max(1, std::min(w_max, nc) - 3); --w) {
    const int waste = (nc + w - 1) / w * w - nc;
    if (waste * 100 < best_waste * 100 - 2 * nc) {  // accept a narrower strip only if it saves > 2 % of the cells
      best = w;
      best_waste = waste;
    }
  }
  return best;
}

template <bool BWD, bool ELEM>
int launch(const feo_operator* op, const DevLatticePlan& L, const float* src0, const float* src1, const float* fT, float* outT,
           float* partials, size_t partial_cap, const float* grad_loss, int64_t ldb, int32_t B, int* n_partials, cudaStream_t st) {
  constexpr int NCOEF = lat_ncoef(BWD, ELEM);
  const int dir = BWD ? 1 : (ELEM ? 2 : 0);
  LatParams<NCOEF> p;  // 20-30 KB of kernel parameters (the class tables)
  p.fT = fT;
  p.outT = outT;
  p.partials = partials;
  p.grad_loss = grad_loss;
  p.ldb = ldb;
  p.B = B;
  p.n_slabs = (B + kSlab - 1) / kSlab;
  p.n = L.n;
  p.nc = L.nc;
  p.N = op->n;
  const int w_max = BWD ? 13 : 15;
  // backward: a 6-row ring of (r, alpha) runs fits 12 cells (13 warps x 128 registers); forward: 15 cells fit a 9-row ring
  p.W = env_int(BWD ? "FEO_LAT_W_BWD" : "FEO_LAT_W_FWD", pick_width(L.nc, BWD ? 12 : 15));
  if (p.W < 1 || p.W > w_max) return fail(FEO_ERR_INVALID_ARGUMENT, "lattice kernel: strip width out of range");
  p.n_strips = (L.nc + p.W - 1) / p.W;
  const int64_t total = (int64_t)p.n_strips * p.n_slabs * L.nc;
  if (total >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "too many work items");
  p.total_steps = (int32_t)total;
  p.he = 5 * p.W + 8;
  p.ho = 4 * p.W + 6;
  p.slot_bytes = (uint32_t)p.he * kLineBytes * (BWD ? 2u : 1u);
  p.R = env_int(BWD ? "FEO_LAT_R_BWD" : "FEO_LAT_R_FWD", BWD ? 6 : 9);
  if (p.R < 6 || p.R > 9) return fail(FEO_ERR_INVALID_ARGUMENT, "lattice kernel: ring rows must be 6 .. 9");
  while (p.R > 6 && (uint64_t)p.R * p.slot_bytes + 64 > kSmemMax) p.R -= 1;
  p.D = (p.R - 3) / 2;  // rows 2L + 3 - R, 2L + 4 - R belong to the early-released part of steps >= L - D
  p.K = p.D + 1;
  p.bar_off = (uint32_t)p.R * p.slot_bytes;
  const uint32_t smem = p.bar_off + 2 * kStagesMax * 8;
  if (smem > kSmemMax) return fail(FEO_ERR_INVALID_ARGUMENT, "lattice kernel: the row ring does not fit shared memory");
  p.debug = env_int("FEO_DEBUG_MODE", 0);
  p.sync_mask = env_int(BWD ? "FEO_LAT_SYNC_BWD" : "FEO_LAT_SYNC_FWD", 3);
  if (p.R % 2 == 0) p.sync_mask |= 1;  // an even ring has no spare slot: the slots a fetch needs are free only after the early release
  p.precond = op->ns_branch;
  p.esign = op->ns_branch ? 1.0f : -1.0f;
  if (L.n_classes[dir] > kLatMaxClasses || L.tab[dir].size() != (size_t)L.n_classes[dir] * NCOEF)
    return fail(FEO_ERR_INVALID_ARGUMENT, "lattice plan: class tables do not match the kernels");
  std::copy(L.cat_cls[dir], L.cat_cls[dir] + 25, p.cat_cls);
  std::fill(p.exist, p.exist + kLatMaxClasses, (uint8_t)0);
  std::copy(L.exist[dir].begin(), L.exist[dir].end(), p.exist);
  std::copy(L.tab[dir].begin(), L.tab[dir].end(), p.tab);
  int sms = 1;
  if (int rc = sm_count(&sms)) return rc;
  const int grid = (int)std::min<int64_t>(total, sms);
  if (n_partials != nullptr) {
    *n_partials = grid * p.W;
    if ((size_t)*n_partials > partial_cap) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  }
  LatMaps maps;
  for (int a = 0; a < 2; ++a) {
    const float* base = a == 0 ? src0 : src1;
    if (int rc = make_box_map(base, ldb, op->n, p.he, &maps.m[a][0])) return rc;
    if (int rc = make_box_map(base, ldb, op->n, p.ho, &maps.m[a][1])) return rc;
  }
  maps.f[0] = maps.f[1] = maps.m[0][0];
  if (!BWD) {
    if (int rc = make_box_map(fT, ldb, op->n, 5 * p.W, &maps.f[0])) return rc;
    if (int rc = make_box_map(fT, ldb, op->n, 4 * p.W, &maps.f[1])) return rc;
  }
  // the register file (64 K) is split over the launched warps: up to 12 / 13 / 14 / 16 warps -> 168 / 152 / 144 / 128 registers
  const unsigned nt = (unsigned)(p.W + 1) * 32;
#define FEO_LAT_LAUNCH(NT)                                                                                          \
  do {                                                                                                              \
    auto kern = residual_lattice_kernel<BWD, NT, ELEM>;                                                                   \
    FEO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    kern<<<(unsigned)grid, nt, smem, st>>>(maps, p);                                                                \
  } while (0)
  if constexpr (BWD) {
    if (nt <= 384)
      FEO_LAT_LAUNCH(384);
    else if (nt <= 416)
      FEO_LAT_LAUNCH(416);
    else
      FEO_LAT_LAUNCH(448);
  } else {
    FEO_LAT_LAUNCH(512);
  }
#undef FEO_LAT_LAUNCH
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

}  // namespace

int lattice_fwd_warps() { return 15; }

int launch_lattice_fwd(const feo_operator* op, const DevLatticePlan& L, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                       float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  if (int rc = check_layout(fT, ldb, B, "fT")) return rc;
  if (rT != nullptr)
    if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (loss_out == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "loss_out is NULL");
  if (ws == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  int n_partials = 0;
  if (L.element_walk) {
    if (int rc = launch<false, true>(op, L, alphaT, alphaT, fT, rT, (float*)ws, ws_bytes / sizeof(float), nullptr, ldb, B, &n_partials, st)) return rc;
  } else {
    if (int rc = launch<false, false>(op, L, alphaT, alphaT, fT, rT, (float*)ws, ws_bytes / sizeof(float), nullptr, ldb, B, &n_partials, st)) return rc;
  }
  return finalize_loss((float*)ws, n_partials, 1.0f, loss_out, st);
}

int launch_lattice_bwd(const feo_operator* op, const DevLatticePlan& L, const float* alphaT, const float* rT, const float* grad_loss,
                       float* gradT, int64_t ldb, int32_t B, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (int rc = check_layout(gradT, ldb, B, "gradT")) return rc;
  const float* a = L.has_conv ? alphaT : rT;  // linear operators: the alpha tables are all zero, any finite source will do
  if (int rc = check_layout(a, ldb, B, "alphaT")) return rc;
  return launch<true, false>(op, L, rT, a, nullptr, gradT, nullptr, 0, grad_loss, ldb, B, nullptr, st);
}

}  // namespace feo
