// sm_100a kernels of the fused sparse residual (forward: r, loss; backward: grad alpha).
//
// A work UNIT = one tile of operator rows (forward) / columns (backward) x one slab of 64 samples
// (feo_internal.h, feo_tiles.cpp).  The kernels are PERSISTENT: one CTA per SM walks the units
// blockIdx.x, blockIdx.x + gridDim.x, ... through a two-stage shared-memory pipeline:
//   1. a producer warp stages the dof lines of the NEXT unit (64 samples x 4 B each) with 2-D TMA boxes
//      (cp.async.bulk.tensor) straight from the dof-major batch arrays while the consumer warps work on
//      the current one (full/empty mbarriers per stage; a warp may run one unit ahead of the slowest);
//   2. every consumer warp streams its private list of 16-byte operator words through a 2 x 512 B shared
//      memory ring filled by 1-D bulk copies (cp.async.bulk) that complete on per-slot mbarriers; the
//      ring runs continuously across units, so the first words of the next unit are already in flight
//      when the current one ends;
//   3. gathers are conflict-free LDS.128 from the staged lines, the arithmetic is packed fp32
//      (fma.rn.f32x2: two IEEE fp32 FMAs per instruction).
// The load-store pipe (128 B/clk of shared-memory data per SM) and the exposed shared-memory round trips of the
// 15 consumer warps are what bound these kernels, so the mapping minimises the gathers and batches them:
// forward = one velocity row PAIR (pair quads) or one row (row / A-quads) per quarter-warp, 2 x 4 samples per
// lane; backward = one column PAIR per half-warp, 4 samples per lane, so that the gathers of r[I], r[J],
// alpha[I], alpha[J] of a neighbour node serve both columns; several steps' loads per loop iteration.
// Everything is row-/column-owned with a fixed summation order: no atomics, bit-reproducible.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "feo_internal.h"
#include "feo_patch.h"
#include "feo_lattice.h"

namespace feo {
namespace {

typedef unsigned long long u64;
constexpr uint32_t kSmemMax = 232448;  // opt-in dynamic shared memory per CTA (227 KB)

struct TensorMaps {
  CUtensorMap m[2][kBoxClasses];  // [source array][box class]
};

struct TiledParams {
  const int32_t* tile_box_ptr;
  const int32_t* tile_lines;
  const StageBox* boxes;
  const WarpRange* warp_range;
  const int4* stream;
  const float* fT;         // forward: load vectors
  float* outT;             // forward: rT (may be NULL) ; backward: gradT
  float* partials;         // forward: one loss partial per (CTA, consumer warp)
  const float* grad_loss;  // backward: upstream gradient (NULL = 1)
  int64_t ldb;
  int32_t B, n_slabs;
  int32_t n_units;    // tiles x slabs
  int32_t g_div, g_mod;  // gridDim.x / n_slabs, gridDim.x % n_slabs: (tile, slab) of a CTA's next unit without a division
  int32_t warps;      // consumer warps (the tile plan's warp count); warp `warps` is the producer
  uint32_t buf_bytes; // size of one line stage
  int32_t n_stages;   // line stages in shared memory (2 or 3)
  int32_t debug;      // developer timing modes (FEO_DEBUG_MODE): 1 = stage lines only, 2 = compute only (no staging)
  int32_t precond;    // forward: 1 -> r = lhs - (f - c), 0 -> r = lhs - (-f + c)
  float esign;        // backward: +1 precond branch, -1 otherwise
  uint32_t ring_off;  // byte offset of the warp rings in dynamic shared memory
  uint32_t bar_off;   // byte offset of the mbarriers
};

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ int4 lds_word(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// 16 bytes of a staged line as two packed sample pairs
__device__ __forceinline__ void lds_pairs(uint32_t a, u64& p0, u64& p1) {
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(p0), "=l"(p1) : "r"(a));
}
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// d = a * b + d on two packed fp32 lanes (each an IEEE fused multiply-add)
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 fma2r(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 bc(float a) { return pk(a, a); }
// Accumulator pair kept as two fp32 registers: (lo, hi) += a * (b.lo, b.hi) as ONE packed FMA.  Passing the
// accumulator as scalars (packed only inside the asm block) lets ptxas update it in place; with 64-bit
// loop-carried accumulators it parks results in the dying gather registers and copies them back.
struct P2 {
  float lo, hi;
};
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
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The per-warp operator stream: a ring of kRingChunks slots of kChunkWords 16-byte words in shared
// memory (base aligned to the ring size), filled by 1-D bulk copies that complete on per-slot
// mbarriers.  The ring is continuous over the units of the persistent CTA: chunk g (counted over all
// units) lives in slot g % kRingChunks, every unit's words start at a fresh chunk.  The reader ENTERS
// chunk g when it is about to read its first word; chunk g - 1 is then fully consumed and its slot is
// refilled with chunk g + kRingChunks - 1, which may already belong to a later unit.
constexpr uint32_t kChunkBytes = kChunkWords * 16;
constexpr uint32_t kRingBytes = kRingChunks * kChunkBytes;
struct Stream {
  uint32_t base, bar;  // shared addresses of the slots / their mbarriers
  uint32_t entered;    // chunks entered by the reader so far
  uint32_t issued;     // chunks issued by the feeder so far
  const int4* src;     // next chunk to issue
  int32_t left;        // chunks left to issue in the feeder's current unit
  int32_t u_next;      // the unit the feeder moves to next, its tile and slab
  int32_t t_next, s_next;
  int2 nxt;            // that unit's WarpRange {begin, n_words}, fetched one unit ahead

  __device__ __forceinline__ void fetch_range(const TiledParams& p, int warp) {
    if (u_next < p.n_units) nxt = __ldg(reinterpret_cast<const int2*>(p.warp_range + (size_t)t_next * p.warps + warp));
  }
  // all lanes keep the bookkeeping, lane 0 issues the copy
  __device__ __forceinline__ void issue_one(const TiledParams& p, int warp, int lane) {
    while (left == 0) {
      if (u_next >= p.n_units) return;
      src = p.stream + nxt.x;
      left = (nxt.y + kChunkWords - 1) / kChunkWords;
      u_next += (int32_t)gridDim.x;
      t_next += p.g_div;
      s_next += p.g_mod;
      if (s_next >= p.n_slabs) {
        s_next -= p.n_slabs;
        ++t_next;
      }
      fetch_range(p, warp);
    }
    if (lane == 0) {
      const uint32_t slot = issued & (kRingChunks - 1);
      mbar_expect_tx(bar + slot * 8, kChunkBytes);
      bulk_copy(base + slot * kChunkBytes, src, kChunkBytes, bar + slot * 8);
    }
    src += kChunkWords;
    --left;
    ++issued;
  }
  __device__ __forceinline__ void start(const TiledParams& p, uint32_t base_, uint32_t bar_, int warp, int lane) {
    base = base_;
    bar = bar_;
    entered = issued = 0;
    src = nullptr;
    left = 0;
    u_next = (int32_t)blockIdx.x;
    t_next = u_next / p.n_slabs;
    s_next = u_next - t_next * p.n_slabs;
    nxt = make_int2(0, 0);
    fetch_range(p, warp);
    for (int c = 0; c < kRingChunks; ++c) issue_one(p, warp, lane);
  }
  __device__ __forceinline__ void enter(const TiledParams& p, int warp, int lane) {
    const uint32_t g = entered++;
    mbar_wait(bar + (g & (kRingChunks - 1)) * 8, (g / kRingChunks) & 1u);
    if (g >= 1) {
      __syncwarp();
      issue_one(p, warp, lane);
    }
  }
};

// (lo, hi) += (a.lo, a.hi) * (b.lo, b.hi)
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

// residual from LHS sum, load vector and convection, mirroring the reference's operation order:
// precond branch  r = LHS - (F - c) ; else  r = LHS - (-F + c)
// (FEONet_steady_Navier-Stokes/train_FEONet.py:324-330, :356)
__device__ __forceinline__ float resid1(float lhs, float f, float c, bool precond) {
  return precond ? __fsub_rn(lhs, __fsub_rn(f, c)) : __fsub_rn(lhs, __fadd_rn(-f, c));
}
// c = u_i*Bu1 + u_j*Bu2 as two rounded products and one rounded add (train_FEONet.py:317-322)
__device__ __forceinline__ float conv1(float d1, float s1, float d2, float s2) {
  return __fadd_rn(__fmul_rn(d1, s1), __fmul_rn(d2, s2));
}

// Shared-memory map of a persistent CTA: [stage 0 lines]...[stage n-1 lines][warp rings][mbarriers]
//   mbarriers: full[4], empty[4] (n_stages used), then kRingChunks per consumer warp
struct Bars {
  uint32_t full, empty, rings;
};
// Unit i of a CTA lives in stage i % n_stages; its barriers are in phase (i / n_stages) & 1.
struct StageCursor {
  uint32_t s = 0, ph = 0;
  __device__ __forceinline__ void next(int32_t n_stages) {
    if (++s == (uint32_t)n_stages) {
      s = 0;
      ph ^= 1u;
    }
  }
};
__device__ __forceinline__ Bars setup_barriers(const TiledParams& p, uint32_t sb, int warp, int lane) {
  Bars b;
  b.full = sb + p.bar_off;
  b.empty = b.full + 32;
  b.rings = b.full + 64;
  if (threadIdx.x == 0)
    for (int k = 0; k < p.n_stages; ++k) {
      mbar_init(b.full + k * 8, 1);
      mbar_init(b.empty + k * 8, (uint32_t)p.warps);
    }
  if (warp < p.warps && lane == 0)
    for (int c = 0; c < kRingChunks; ++c) mbar_init(b.rings + (uint32_t)warp * (kRingChunks * 8) + c * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  return b;
}

// (tile, slab) of the unit gridDim.x further
__device__ __forceinline__ void next_unit(const TiledParams& p, int& tile, int& slab) {
  tile += p.g_div;
  slab += p.g_mod;
  if (slab >= p.n_slabs) {
    slab -= p.n_slabs;
    ++tile;
  }
}

// The producer warp: stages the lines of unit i into stage i % n_stages as soon as every consumer warp has
// released the unit that used the stage before.  The staging boxes of the NEXT unit are fetched into
// registers (three per lane) right after the current unit's copies are issued, so that no dependent global
// load sits between the release of a stage and the first TMA request into it.
constexpr int kBoxPrefetch = 3;
struct BoxBatch {
  int32_t b0, b1;
  uint32_t bytes;
  StageBox bx[kBoxPrefetch];
  __device__ __forceinline__ void fetch(const TiledParams& p, int tile, int lane) {
    b0 = p.tile_box_ptr[tile];
    b1 = p.tile_box_ptr[tile + 1];
    bytes = (uint32_t)p.tile_lines[tile] * kLineBytes;
#pragma unroll
    for (int k = 0; k < kBoxPrefetch; ++k) {
      const int b = b0 + lane + 32 * k;
      if (b < b1) bx[k] = p.boxes[b];
    }
  }
};
__device__ __forceinline__ void produce_lines(const TensorMaps& maps, const TiledParams& p, uint32_t sb, const Bars& bars, int lane) {
  if (p.debug == 2) return;
  int tile = (int)blockIdx.x / p.n_slabs, slab = (int)blockIdx.x - tile * p.n_slabs;
  BoxBatch cur_boxes;
  cur_boxes.fetch(p, tile, lane);
  StageCursor cur;
  int i = 0;
  for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x, ++i, cur.next(p.n_stages)) {
    const uint32_t s = cur.s;
    if (i >= p.n_stages) mbar_wait(bars.empty + s * 8, cur.ph ^ 1u);  // the unit that used the stage before is released
    if (lane == 0) mbar_expect_tx(bars.full + s * 8, cur_boxes.bytes);
    __syncwarp();
    const uint32_t dst = sb + s * p.buf_bytes;
    const int c0 = slab * kSlab;
#pragma unroll
    for (int k = 0; k < kBoxPrefetch; ++k)
      if (cur_boxes.b0 + lane + 32 * k < cur_boxes.b1) {
        const StageBox bx = cur_boxes.bx[k];
        tma_box(dst + (uint32_t)bx.line0 * kLineBytes, &maps.m[bx.src][bx.cls], c0, bx.dof0, bars.full + s * 8);
      }
    for (int b = cur_boxes.b0 + lane + 32 * kBoxPrefetch; b < cur_boxes.b1; b += 32) {
      const StageBox bx = p.boxes[b];
      tma_box(dst + (uint32_t)bx.line0 * kLineBytes, &maps.m[bx.src][bx.cls], c0, bx.dof0, bars.full + s * 8);
    }
    next_unit(p, tile, slab);
    if (u + (int)gridDim.x < p.n_units) cur_boxes.fetch(p, tile, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// forward: r = A a -/+ (F - c), loss partial = sum r^2
// ---------------------------------------------------------------------------------------------
template <int MAXW>
__global__ void __launch_bounds__((MAXW + 1) * 32, 1) residual_fwd_tiled(const __grid_constant__ TensorMaps maps, const TiledParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sb = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Bars bars = setup_barriers(p, sb, warp, lane);
  if (warp == p.warps) {
    produce_lines(maps, p, sb, bars, lane);
    return;
  }
  Stream ring;
  if (p.debug == 1) {  // staging only: wait for every unit's lines, touch nothing
    StageCursor cur;
    for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x, cur.next(p.n_stages)) {
      mbar_wait(bars.full + cur.s * 8, cur.ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(bars.empty + cur.s * 8);
    }
    return;
  }
  ring.start(p, sb + p.ring_off + (uint32_t)warp * kRingBytes, bars.rings + (uint32_t)warp * (kRingChunks * 8), warp, lane);

  const int q = lane >> 3;  // this lane's row slot in the quad
  // a lane owns samples [4l, 4l+4) and [32+4l, 32+4l+4), l = lane & 7: each LDS.128 of a quarter-warp then
  // reads 128 contiguous bytes of one line (conflict-free), the second one 128 B further
  const uint32_t lane_off = (uint32_t)(lane & 7) * 16;
  const uint32_t rq = ring.base + (uint32_t)q * 16;  // this quarter's word within a 64-byte unit
  const bool precond = p.precond != 0;
  const WarpRange* my_ranges = p.warp_range + warp;
  double dsum = 0.0;

  int tile_n = (int)blockIdx.x / p.n_slabs, slab_n = (int)blockIdx.x - tile_n * p.n_slabs;  // the unit after the current one
  int32_t n_words_next = __ldg(&my_ranges[(size_t)tile_n * p.warps].n_words);
  StageCursor cur;
  for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x, cur.next(p.n_stages)) {
    const int32_t n_words = n_words_next;
    const int slab = slab_n;
    next_unit(p, tile_n, slab_n);
    if (u + (int)gridDim.x < p.n_units) n_words_next = __ldg(&my_ranges[(size_t)tile_n * p.warps].n_words);
    const uint32_t lines = sb + cur.s * p.buf_bytes + lane_off;
    const int b0 = slab * kSlab + (lane & 7) * 4;
    if (p.debug != 2) mbar_wait(bars.full + cur.s * 8, cur.ph);
    uint32_t ptr = (ring.entered & (kRingChunks - 1)) * kChunkBytes;  // the unit's words start at a fresh chunk
    float lsum = 0.f;

    // stream = quads: [header unit][spare unit][n_steps step units], n_steps even; units are 64 B, so
    // a pair of units never straddles a 512-byte chunk
    int units_left = n_words / 4;
    while (units_left > 0) {
      if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
      const int4 hdr = lds_word(rq + ptr);
      ptr = (ptr + 128) & (kRingBytes - 1);
      if (hdr.w & 2) {
        // ---- pair quad: each quarter owns the two velocity rows (I[k], J[k]) of a node ----------------------
        // [header][spare][nX X-steps (2 units)][nP P-steps (unit each, nP even)][nS S-steps (unit each)][pad if nS odd]
        const int rowI = hdr.x, rowJ = hdr.y;
        const int nS = (hdr.w >> 8) & 255, nP = (hdr.w >> 16) & 255, nX = (int)((uint32_t)hdr.w >> 24);
        units_left -= 2 + 2 * nX + nP + nS + (nS & 1);
        float4 fI[2], fJ[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          fI[k] = fJ[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rowI >= 0 && b0 + 32 * k < p.ldb) {
            fI[k] = ldg4_stream(p.fT + (int64_t)rowI * p.ldb + b0 + 32 * k);
            fJ[k] = ldg4_stream(p.fT + (int64_t)rowJ * p.ldb + b0 + 32 * k);
          }
        }
        P2 aI[4], uI[4], vI[4], aJ[4], uJ[4], vJ[4];  // A, B1, B2 sums of row I and of row J
#pragma unroll
        for (int k = 0; k < 4; ++k) aI[k] = uI[k] = vI[k] = aJ[k] = uJ[k] = vJ[k] = P2{0.f, 0.f};
        // X-steps: one column feeds both rows with its own coefficients {off, aI, b1I, b2I} {aJ, b1J, b2J, 0}
        for (int s = 0; s < nX;) {
          if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
          const int seg = min(nX - s, (int)((kChunkBytes - (ptr & (kChunkBytes - 1))) / 128));
          s += seg;
          const uint32_t rp = rq + ptr;
          ptr = (ptr + (uint32_t)seg * 128) & (kRingBytes - 1);
#pragma unroll 1
          for (int t = 0; t < seg; ++t) {
            const int4 e0 = lds_word(rp + (uint32_t)t * 128);
            const int4 e1 = lds_word(rp + (uint32_t)t * 128 + 64);
            u64 x[4];
            lds_pairs(lines + (uint32_t)e0.x, x[0], x[1]);
            lds_pairs(lines + (uint32_t)e0.x + 128, x[2], x[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(aI[k], __int_as_float(e0.y), x[k]);
              fma2s(uI[k], __int_as_float(e0.z), x[k]);
              fma2s(vI[k], __int_as_float(e0.w), x[k]);
              fma2s(aJ[k], __int_as_float(e1.x), x[k]);
              fma2s(uJ[k], __int_as_float(e1.y), x[k]);
              fma2s(vJ[k], __int_as_float(e1.z), x[k]);
            }
          }
        }
        // P-steps, two per iteration: one column with b1 = b2 = 0 feeds both rows {off, aI, aJ, 0}
        for (int s = 0; s < nP;) {
          if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
          const int seg = min(nP - s, (int)((kChunkBytes - (ptr & (kChunkBytes - 1))) / 64));
          s += seg;
          const uint32_t rp = rq + ptr;
          ptr = (ptr + (uint32_t)seg * 64) & (kRingBytes - 1);
#pragma unroll 1
          for (int t = 0; t < seg; t += 2) {
            const int4 e0 = lds_word(rp + (uint32_t)t * 64);
            const int4 e1 = lds_word(rp + (uint32_t)t * 64 + 64);
            u64 x0[4], x1[4];
            lds_pairs(lines + (uint32_t)e0.x, x0[0], x0[1]);
            lds_pairs(lines + (uint32_t)e0.x + 128, x0[2], x0[3]);
            lds_pairs(lines + (uint32_t)e1.x, x1[0], x1[1]);
            lds_pairs(lines + (uint32_t)e1.x + 128, x1[2], x1[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(aI[k], __int_as_float(e0.y), x0[k]);
              fma2s(aJ[k], __int_as_float(e0.z), x0[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(aI[k], __int_as_float(e1.y), x1[k]);
              fma2s(aJ[k], __int_as_float(e1.z), x1[k]);
            }
          }
        }
        // S-steps: the columns (I[m], J[m]) of a neighbour node with the same coefficients for both rows
        // {line(I[m]) | line(J[m]) << 16, a, b1, b2}: row I takes alpha[I[m]], row J takes alpha[J[m]]
        for (int s = 0; s < nS;) {
          if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
          const int seg = min(nS - s, (int)((kChunkBytes - (ptr & (kChunkBytes - 1))) / 64));
          s += seg;
          const uint32_t rp = rq + ptr;
          ptr = (ptr + (uint32_t)seg * 64) & (kRingBytes - 1);
#pragma unroll 1
          for (int t = 0; t < seg; ++t) {
            const int4 e = lds_word(rp + (uint32_t)t * 64);
            const uint32_t oI = ((uint32_t)e.x & 0xffffu) * kLineBytes, oJ = ((uint32_t)e.x >> 16) * kLineBytes;
            u64 xI[4], xJ[4];
            lds_pairs(lines + oI, xI[0], xI[1]);
            lds_pairs(lines + oI + 128, xI[2], xI[3]);
            lds_pairs(lines + oJ, xJ[0], xJ[1]);
            lds_pairs(lines + oJ + 128, xJ[2], xJ[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(aI[k], __int_as_float(e.y), xI[k]);
              fma2s(uI[k], __int_as_float(e.z), xI[k]);
              fma2s(vI[k], __int_as_float(e.w), xI[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(aJ[k], __int_as_float(e.y), xJ[k]);
              fma2s(uJ[k], __int_as_float(e.z), xJ[k]);
              fma2s(vJ[k], __int_as_float(e.w), xJ[k]);
            }
          }
        }
        if (nS & 1) ptr = (ptr + 64) & (kRingBytes - 1);  // pad unit: items stay 128-byte aligned
        if (rowI >= 0) {
          u64 d1[4], d2[4];
          const uint32_t li = ((uint32_t)hdr.z & 0xffffu) * kLineBytes, lj = ((uint32_t)hdr.z >> 16) * kLineBytes;
          lds_pairs(lines + li, d1[0], d1[1]);
          lds_pairs(lines + li + 128, d1[2], d1[3]);
          lds_pairs(lines + lj, d2[0], d2[1]);
          lds_pairs(lines + lj + 128, d2[2], d2[3]);
          float cI[8], cJ[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float u0, u1, v0, v1;
            unpk(d1[k], u0, u1);
            unpk(d2[k], v0, v1);
            cI[2 * k] = conv1(u0, uI[k].lo, v0, vI[k].lo);
            cI[2 * k + 1] = conv1(u1, uI[k].hi, v1, vI[k].hi);
            cJ[2 * k] = conv1(u0, uJ[k].lo, v0, vJ[k].lo);
            cJ[2 * k + 1] = conv1(u1, uJ[k].hi, v1, vJ[k].hi);
          }
          float* rI = p.outT != nullptr ? p.outT + (int64_t)rowI * p.ldb + b0 : nullptr;
          float* rJ = p.outT != nullptr ? p.outT + (int64_t)rowJ * p.ldb + b0 : nullptr;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int b = b0 + 32 * k;
            if (b < p.ldb) {
              float4 r, q4;
              r.x = resid1(aI[2 * k].lo, fI[k].x, cI[4 * k + 0], precond);
              r.y = resid1(aI[2 * k].hi, fI[k].y, cI[4 * k + 1], precond);
              r.z = resid1(aI[2 * k + 1].lo, fI[k].z, cI[4 * k + 2], precond);
              r.w = resid1(aI[2 * k + 1].hi, fI[k].w, cI[4 * k + 3], precond);
              q4.x = resid1(aJ[2 * k].lo, fJ[k].x, cJ[4 * k + 0], precond);
              q4.y = resid1(aJ[2 * k].hi, fJ[k].y, cJ[4 * k + 1], precond);
              q4.z = resid1(aJ[2 * k + 1].lo, fJ[k].z, cJ[4 * k + 2], precond);
              q4.w = resid1(aJ[2 * k + 1].hi, fJ[k].w, cJ[4 * k + 3], precond);
              if (b + 0 < p.B) lsum = fmaf(q4.x, q4.x, fmaf(r.x, r.x, lsum));
              if (b + 1 < p.B) lsum = fmaf(q4.y, q4.y, fmaf(r.y, r.y, lsum));
              if (b + 2 < p.B) lsum = fmaf(q4.z, q4.z, fmaf(r.z, r.z, lsum));
              if (b + 3 < p.B) lsum = fmaf(q4.w, q4.w, fmaf(r.w, r.w, lsum));
              if (rI != nullptr && b < p.B) {
                *reinterpret_cast<float4*>(rI + 32 * k) = r;
                *reinterpret_cast<float4*>(rJ + 32 * k) = q4;
              }
            }
          }
        }
        continue;
      }
      const int n_steps = hdr.y;
      const int row = hdr.x;
      units_left -= 2 + ((hdr.w & 4) ? n_steps / 2 : n_steps);
      // the load vector of this row is fetched now and consumed in the epilogue (hides the DRAM latency)
      float4 fv[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        fv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= 0 && b0 + 32 * k < p.ldb) fv[k] = ldg4_stream(p.fT + (int64_t)row * p.ldb + b0 + 32 * k);
      }
      P2 accA[4], acc1[4], acc2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) accA[k] = acc1[k] = acc2[k] = P2{0.f, 0.f};
      // step pairs, in segments that end at a chunk boundary so that the inner loop carries no ring logic.
      // (A software-pipelined variant of this loop -- gathers of pair j + 1 in flight during the FMAs of pair j --
      // was measured SLOWER, 5.0-5.7 ms vs 4.06 ms at cfg5: the kernel is bound by the load-store pipe, not by
      // the exposed round trips, and the extra control flow costs issue slots.)
      if (hdr.w & 4) {
        // A-quad: rows without B1/B2 entries (pressure rows), two steps per word {off, a, off, a}; four steps
        // (two words, four gathered lines) per iteration
        for (int s = 0; s < n_steps;) {
          if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
          const int seg = min((n_steps - s) / 2, (int)((kChunkBytes - (ptr & (kChunkBytes - 1))) / 64));  // packed units
          s += 2 * seg;
          const uint32_t rp = rq + ptr;
          ptr = (ptr + (uint32_t)seg * 64) & (kRingBytes - 1);
#pragma unroll 1
          for (int t = 0; t < seg; t += 2) {
            const int4 e0 = lds_word(rp + (uint32_t)t * 64);
            const int4 e1 = lds_word(rp + (uint32_t)t * 64 + 64);
            u64 x[4][4];
            lds_pairs(lines + (uint32_t)e0.x, x[0][0], x[0][1]);
            lds_pairs(lines + (uint32_t)e0.x + 128, x[0][2], x[0][3]);
            lds_pairs(lines + (uint32_t)e0.z, x[1][0], x[1][1]);
            lds_pairs(lines + (uint32_t)e0.z + 128, x[1][2], x[1][3]);
            lds_pairs(lines + (uint32_t)e1.x, x[2][0], x[2][1]);
            lds_pairs(lines + (uint32_t)e1.x + 128, x[2][2], x[2][3]);
            lds_pairs(lines + (uint32_t)e1.z, x[3][0], x[3][1]);
            lds_pairs(lines + (uint32_t)e1.z + 128, x[3][2], x[3][3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              fma2s(accA[k], __int_as_float(e0.y), x[0][k]);
              fma2s(accA[k], __int_as_float(e0.w), x[1][k]);
              fma2s(accA[k], __int_as_float(e1.y), x[2][k]);
              fma2s(accA[k], __int_as_float(e1.w), x[3][k]);
            }
          }
        }
      }
      for (int s = (hdr.w & 4) ? n_steps : 0; s < n_steps;) {
        if ((ptr & (kChunkBytes - 1)) == 0) ring.enter(p, warp, lane);
        const int seg = min(n_steps - s, (int)((kChunkBytes - (ptr & (kChunkBytes - 1))) / 64));
        s += seg;
        const uint32_t rp = rq + ptr;
        ptr = (ptr + (uint32_t)seg * 64) & (kRingBytes - 1);
#pragma unroll 1
        for (int t = 0; t < seg; t += 2) {
          const int4 e0 = lds_word(rp + (uint32_t)t * 64);
          const int4 e1 = lds_word(rp + (uint32_t)t * 64 + 64);
          u64 x0[4], x1[4];
          lds_pairs(lines + (uint32_t)e0.x, x0[0], x0[1]);
          lds_pairs(lines + (uint32_t)e0.x + 128, x0[2], x0[3]);
          lds_pairs(lines + (uint32_t)e1.x, x1[0], x1[1]);
          lds_pairs(lines + (uint32_t)e1.x + 128, x1[2], x1[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            fma2s(accA[k], __int_as_float(e0.y), x0[k]);
            fma2s(acc1[k], __int_as_float(e0.z), x0[k]);
            fma2s(acc2[k], __int_as_float(e0.w), x0[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            fma2s(accA[k], __int_as_float(e1.y), x1[k]);
            fma2s(acc1[k], __int_as_float(e1.z), x1[k]);
            fma2s(acc2[k], __int_as_float(e1.w), x1[k]);
          }
        }
      }
      // epilogue: convection product, load vector, residual, loss, store
      if (row >= 0) {
        float lhs[8], c[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          lhs[2 * k] = accA[k].lo;
          lhs[2 * k + 1] = accA[k].hi;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = 0.f;
        if (hdr.w & 1) {
          u64 d1[4], d2[4];
          const uint32_t li = ((uint32_t)hdr.z & 0xffffu) * kLineBytes, lj = ((uint32_t)hdr.z >> 16) * kLineBytes;
          lds_pairs(lines + li, d1[0], d1[1]);
          lds_pairs(lines + li + 128, d1[2], d1[3]);
          lds_pairs(lines + lj, d2[0], d2[1]);
          lds_pairs(lines + lj + 128, d2[2], d2[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float u0, u1, v0, v1;
            unpk(d1[k], u0, u1);
            unpk(d2[k], v0, v1);
            c[2 * k] = conv1(u0, acc1[k].lo, v0, acc2[k].lo);
            c[2 * k + 1] = conv1(u1, acc1[k].hi, v1, acc2[k].hi);
          }
        }
        float* rrow = p.outT != nullptr ? p.outT + (int64_t)row * p.ldb + b0 : nullptr;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int b = b0 + 32 * k;
          if (b < p.ldb) {
            const float4 f = fv[k];
            float4 r;
            r.x = resid1(lhs[4 * k + 0], f.x, c[4 * k + 0], precond);
            r.y = resid1(lhs[4 * k + 1], f.y, c[4 * k + 1], precond);
            r.z = resid1(lhs[4 * k + 2], f.z, c[4 * k + 2], precond);
            r.w = resid1(lhs[4 * k + 3], f.w, c[4 * k + 3], precond);
            if (b + 0 < p.B) lsum = fmaf(r.x, r.x, lsum);
            if (b + 1 < p.B) lsum = fmaf(r.y, r.y, lsum);
            if (b + 2 < p.B) lsum = fmaf(r.z, r.z, lsum);
            if (b + 3 < p.B) lsum = fmaf(r.w, r.w, lsum);
            if (rrow != nullptr && b < p.B) *reinterpret_cast<float4*>(rrow + 32 * k) = r;
          }
        }
      }
    }
    dsum += (double)lsum;  // per-unit fp32 partial, accumulated over the units in fp64
    __syncwarp();
    if (lane == 0) mbar_arrive(bars.empty + cur.s * 8);  // this warp is done with the stage
  }
  // fixed-order reduction: lanes -> warp; the warp partials are summed in fp64 by finalize_loss_kernel
  dsum = warp_sum(dsum);
  if (lane == 0) p.partials[(size_t)blockIdx.x * p.warps + warp] = (float)dsum;
}

// ---------------------------------------------------------------------------------------------
// backward: grad = 2 g [A^T r + s (B1^T (d1 r) + B2^T (d2 r) + E-term)], column-pair owned
// ---------------------------------------------------------------------------------------------
template <int MAXW>
__global__ void __launch_bounds__((MAXW + 1) * 32, 1) residual_bwd_tiled(const __grid_constant__ TensorMaps maps, const TiledParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sb = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Bars bars = setup_barriers(p, sb, warp, lane);
  if (warp == p.warps) {
    produce_lines(maps, p, sb, bars, lane);
    return;
  }
  Stream ring;
  if (p.debug == 1) {  // staging only: wait for every unit's lines, touch nothing
    StageCursor cur;
    for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x, cur.next(p.n_stages)) {
      mbar_wait(bars.full + cur.s * 8, cur.ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(bars.empty + cur.s * 8);
    }
    return;
  }
  ring.start(p, sb + p.ring_off + (uint32_t)warp * kRingBytes, bars.rings + (uint32_t)warp * (kRingChunks * 8), warp, lane);

  const int h = lane >> 4;                            // this lane's pair slot in the duo
  const uint32_t lane_off = (uint32_t)(lane & 15) * 16;  // 4 samples = 16 B of every line
  const uint32_t rh = ring.base + (uint32_t)h * 16;   // this half's word within a word pair
  const float g2 = 2.0f * (p.grad_loss != nullptr ? __ldg(p.grad_loss) : 1.0f);
  const WarpRange* my_ranges = p.warp_range + warp;

  int tile_n = (int)blockIdx.x / p.n_slabs, slab_n = (int)blockIdx.x - tile_n * p.n_slabs;  // the unit after the current one
  int32_t n_words_next = __ldg(&my_ranges[(size_t)tile_n * p.warps].n_words);
  StageCursor cur;
  for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x, cur.next(p.n_stages)) {
    const int32_t n_words = n_words_next;
    const int slab = slab_n;
    next_unit(p, tile_n, slab_n);
    if (u + (int)gridDim.x < p.n_units) n_words_next = __ldg(&my_ranges[(size_t)tile_n * p.warps].n_words);
    const uint32_t lines = sb + cur.s * p.buf_bytes + lane_off;
    const int b0 = slab * kSlab + (lane & 15) * 4;
    if (p.debug != 2) mbar_wait(bars.full + cur.s * 8, cur.ph);

    // The stream is walked by linear word position `pos` (counted over all units: the unit starts at a
    // fresh chunk); pieces (header 4 words, S-step 4, V-step 6, P-step 2, A-step 2, X-step 4) never straddle a 32-word chunk:
    // when the next piece does not fit, both the plan builder and this reader skip to the next chunk.
    int pos = (int)(ring.entered * kChunkWords);
    const int pos_end = pos + n_words;
    auto place = [&](int len) {  // returns the ring address of the piece, entering a new chunk if needed
      if ((pos & (kChunkWords - 1)) + len > kChunkWords) pos = (pos + kChunkWords - 1) & ~(kChunkWords - 1);
      if ((pos & (kChunkWords - 1)) == 0) ring.enter(p, warp, lane);
      return rh + (((uint32_t)pos & (kRingChunks * kChunkWords - 1)) << 4);
    };
    while (pos < pos_end) {
      const uint32_t ha = place(4);
      const int4 h0 = lds_word(ha);
      const int4 h1 = lds_word(ha + 32);
      pos += 4;
      const int nV = h0.z, nA = h0.w, nX = h1.x, nS = (int)((uint32_t)h1.w & 0xffffu), nP = (int)((uint32_t)h1.w >> 16);
      P2 accI[2], accJ[2], bu1I[2], bu2I[2], bu1J[2], bu2J[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) accI[k] = accJ[k] = bu1I[k] = bu2I[k] = bu1J[k] = bu2J[k] = P2{0.f, 0.f};
      // symmetric V-steps: both columns share a, b1, b2 (one effective coefficient t) and f1, f2
      for (int v = 0; v < nS;) {
        const uint32_t wa = place(4);
        const int cnt = min(nS - v, (kChunkWords - (pos & (kChunkWords - 1))) / 4);
        v += cnt;
        pos += 4 * cnt;
        // two steps per iteration: the kernel is bound by exposed shared-memory round trips at 15 warps per SM, so
        // each trip (word loads, then gathers) carries the loads of two steps
        struct SWords {
          int4 w0, w1;
        };
        struct SLines {
          u64 d1[2], d2[2], rI[2], rJ[2];
        };
        auto s_words = [&](int t) {
          SWords w;
          w.w0 = lds_word(wa + (uint32_t)t * 64);
          w.w1 = lds_word(wa + (uint32_t)t * 64 + 32);
          return w;
        };
        auto s_gather = [&](const SWords& w) {
          SLines x;
          lds_pairs(lines + ((uint32_t)w.w0.y & 0xffffu) * kLineBytes, x.d1[0], x.d1[1]);
          lds_pairs(lines + ((uint32_t)w.w0.y >> 16) * kLineBytes, x.d2[0], x.d2[1]);
          lds_pairs(lines + ((uint32_t)w.w0.x & 0xffffu) * kLineBytes, x.rI[0], x.rI[1]);
          lds_pairs(lines + ((uint32_t)w.w0.x >> 16) * kLineBytes, x.rJ[0], x.rJ[1]);
          return x;
        };
        auto s_fma = [&](const SWords& w, const SLines& x) {
          const u64 a = bc(__int_as_float(w.w0.z)), b1 = bc(__int_as_float(w.w0.w)), b2 = bc(__int_as_float(w.w1.x));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const u64 tt = fma2r(b2, x.d2[k], fma2r(b1, x.d1[k], a));
            fma2p(accI[k], tt, x.rI[k]);
            fma2p(accJ[k], tt, x.rJ[k]);
            fma2s(bu1I[k], __int_as_float(w.w1.y), x.d1[k]);
            fma2s(bu2I[k], __int_as_float(w.w1.z), x.d1[k]);
            fma2s(bu1J[k], __int_as_float(w.w1.y), x.d2[k]);
            fma2s(bu2J[k], __int_as_float(w.w1.z), x.d2[k]);
          }
        };
        int t = 0;
#pragma unroll 1
        for (; t + 1 < cnt; t += 2) {
          const SWords wA = s_words(t), wB = s_words(t + 1);
          const SLines xA = s_gather(wA), xB = s_gather(wB);
          s_fma(wA, xA);
          s_fma(wB, xB);
        }
        if (t < cnt) {
          const SWords wA = s_words(t);
          const SLines xA = s_gather(wA);
          s_fma(wA, xA);
        }
      }
      for (int v = 0; v < nV;) {
        const uint32_t wa = place(6);
        const int cnt = min(nV - v, (kChunkWords - (pos & (kChunkWords - 1))) / 6);
        v += cnt;
        pos += 6 * cnt;
#pragma unroll 1
        for (int t = 0; t < cnt; ++t) {
          const int4 w0 = lds_word(wa + (uint32_t)t * 96);
          const int4 w1 = lds_word(wa + (uint32_t)t * 96 + 32);
          const int4 w2 = lds_word(wa + (uint32_t)t * 96 + 64);
          u64 d1[2], d2[2], rI[2], rJ[2];
          lds_pairs(lines + ((uint32_t)w0.y & 0xffffu) * kLineBytes, d1[0], d1[1]);
          lds_pairs(lines + ((uint32_t)w0.y >> 16) * kLineBytes, d2[0], d2[1]);
          lds_pairs(lines + ((uint32_t)w0.x & 0xffffu) * kLineBytes, rI[0], rI[1]);
          lds_pairs(lines + ((uint32_t)w0.x >> 16) * kLineBytes, rJ[0], rJ[1]);
          const u64 aI = bc(__int_as_float(w0.z)), b1I = bc(__int_as_float(w0.w)), b2I = bc(__int_as_float(w1.x));
          const u64 aJ = bc(__int_as_float(w1.y)), b1J = bc(__int_as_float(w1.z)), b2J = bc(__int_as_float(w1.w));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const u64 tI = fma2r(b2I, d2[k], fma2r(b1I, d1[k], aI));
            const u64 tJ = fma2r(b2J, d2[k], fma2r(b1J, d1[k], aJ));
            fma2p(accI[k], tI, rI[k]);
            fma2p(accJ[k], tJ, rJ[k]);
            fma2s(bu1I[k], __int_as_float(w2.x), d1[k]);
            fma2s(bu2I[k], __int_as_float(w2.y), d1[k]);
            fma2s(bu1J[k], __int_as_float(w2.z), d2[k]);
            fma2s(bu2J[k], __int_as_float(w2.w), d2[k]);
          }
        }
      }
      // plain steps whose two columns read the same source row: one gather; four steps per iteration
      for (int a = 0; a < nP;) {
        const uint32_t wa = place(2);
        const int cnt = min(nP - a, (kChunkWords - (pos & (kChunkWords - 1))) / 2);
        a += cnt;
        pos += 2 * cnt;
        int t = 0;
#pragma unroll 1
        for (; t + 3 < cnt; t += 4) {
          int4 w[4];
          u64 rX[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = lds_word(wa + (uint32_t)(t + j) * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) lds_pairs(lines + (uint32_t)w[j].x * kLineBytes, rX[j][0], rX[j][1]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              fma2s(accI[k], __int_as_float(w[j].y), rX[j][k]);
              fma2s(accJ[k], __int_as_float(w[j].z), rX[j][k]);
            }
        }
#pragma unroll 1
        for (; t < cnt; ++t) {
          const int4 w0 = lds_word(wa + (uint32_t)t * 32);
          u64 rX[2];
          lds_pairs(lines + (uint32_t)w0.x * kLineBytes, rX[0], rX[1]);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            fma2s(accI[k], __int_as_float(w0.y), rX[k]);
            fma2s(accJ[k], __int_as_float(w0.z), rX[k]);
          }
        }
      }
      for (int a = 0; a < nA;) {
        const uint32_t wa = place(2);
        const int cnt = min(nA - a, (kChunkWords - (pos & (kChunkWords - 1))) / 2);
        a += cnt;
        pos += 2 * cnt;
        int t = 0;
#pragma unroll 1
        for (; t + 1 < cnt; t += 2) {
          int4 w[2];
          u64 rI[2][2], rJ[2][2];
#pragma unroll
          for (int j = 0; j < 2; ++j) w[j] = lds_word(wa + (uint32_t)(t + j) * 32);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            lds_pairs(lines + ((uint32_t)w[j].x & 0xffffu) * kLineBytes, rI[j][0], rI[j][1]);
            lds_pairs(lines + ((uint32_t)w[j].x >> 16) * kLineBytes, rJ[j][0], rJ[j][1]);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              fma2s(accI[k], __int_as_float(w[j].y), rI[j][k]);
              fma2s(accJ[k], __int_as_float(w[j].z), rJ[j][k]);
            }
        }
        if (t < cnt) {
          const int4 w0 = lds_word(wa + (uint32_t)t * 32);
          u64 rI[2], rJ[2];
          lds_pairs(lines + ((uint32_t)w0.x & 0xffffu) * kLineBytes, rI[0], rI[1]);
          lds_pairs(lines + ((uint32_t)w0.x >> 16) * kLineBytes, rJ[0], rJ[1]);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            fma2s(accI[k], __int_as_float(w0.y), rI[k]);
            fma2s(accJ[k], __int_as_float(w0.z), rJ[k]);
          }
        }
      }
#pragma unroll 1
      for (int x = 0; x < nX; ++x) {
        const uint32_t wa = place(4);
        pos += 4;
        const int4 w0 = lds_word(wa);
        const int4 w1 = lds_word(wa + 32);
        u64 xv[2];
        lds_pairs(lines + (uint32_t)w0.x * kLineBytes, xv[0], xv[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          fma2s(bu1I[k], __int_as_float(w0.y), xv[k]);
          fma2s(bu2I[k], __int_as_float(w0.z), xv[k]);
          fma2s(bu1J[k], __int_as_float(w0.w), xv[k]);
          fma2s(bu2J[k], __int_as_float(w1.x), xv[k]);
        }
      }
      // epilogue: E-term, scale, store
      float oI[4] = {accI[0].lo, accI[0].hi, accI[1].lo, accI[1].hi};
      float oJ[4] = {accJ[0].lo, accJ[0].hi, accJ[1].lo, accJ[1].hi};
      if (h1.z & 1) {
        u64 pI[2], pJ[2];
        lds_pairs(lines + ((uint32_t)h1.y & 0xffffu) * kLineBytes, pI[0], pI[1]);
        lds_pairs(lines + ((uint32_t)h1.y >> 16) * kLineBytes, pJ[0], pJ[1]);
        float ri[4], rj[4];
        unpk(pI[0], ri[0], ri[1]);
        unpk(pI[1], ri[2], ri[3]);
        unpk(pJ[0], rj[0], rj[1]);
        unpk(pJ[1], rj[2], rj[3]);
        const float s1i[4] = {bu1I[0].lo, bu1I[0].hi, bu1I[1].lo, bu1I[1].hi}, s2i[4] = {bu2I[0].lo, bu2I[0].hi, bu2I[1].lo, bu2I[1].hi};
        const float s1j[4] = {bu1J[0].lo, bu1J[0].hi, bu1J[1].lo, bu1J[1].hi}, s2j[4] = {bu2J[0].lo, bu2J[0].hi, bu2J[1].lo, bu2J[1].hi};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          oI[k] = fmaf(p.esign, fmaf(s1j[k], rj[k], s1i[k] * ri[k]), oI[k]);
          oJ[k] = fmaf(p.esign, fmaf(s2j[k], rj[k], s2i[k] * ri[k]), oJ[k]);
        }
      }
      if (b0 < p.B) {
        const int cI = h0.x, cJ = h0.y;
        if (cI >= 0) *reinterpret_cast<float4*>(p.outT + (int64_t)cI * p.ldb + b0) = make_float4(oI[0] * g2, oI[1] * g2, oI[2] * g2, oI[3] * g2);
        if (cJ >= 0) *reinterpret_cast<float4*>(p.outT + (int64_t)cJ * p.ldb + b0) = make_float4(oJ[0] * g2, oJ[1] * g2, oJ[2] * g2, oJ[3] * g2);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars.empty + cur.s * 8);  // this warp is done with the stage
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FEO_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr)
      return fail(FEO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  *out = fn;
  return FEO_OK;
}

// one map per box class over the dof-major array base[n][ldb]: box = kSlab samples x kBoxRows[cls] dofs
int make_maps(const float* base, int64_t ldb, int32_t n, CUtensorMap out[kBoxClasses]) {
  EncodeTiledFn enc;
  if (int rc = get_encode(&enc)) return rc;
  // the driver entry point needs the primary context current on THIS thread (autograd runs the backward
  // on its own threads, where no runtime call may have bound it yet)
  int dev = 0;
  FEO_CUDA_CHECK(cudaGetDevice(&dev));
  FEO_CUDA_CHECK(cudaSetDevice(dev));
  for (int cls = 0; cls < kBoxClasses; ++cls) {
    const cuuint64_t dims[2] = {(cuuint64_t)ldb, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)ldb * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kSlab, (cuuint32_t)kBoxRows[cls]};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&out[cls], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FEO_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  }
  return FEO_OK;
}

int check_layout(const void* p, int64_t ld, int32_t B, const char* what) {
  if (p == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " is NULL");
  if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " not 16-byte aligned");
  if (ld % 4 != 0 || ld < ((B + 3) / 4) * 4)
    return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + ": ldb must be a multiple of 4 and >= ceil4(B)");
  return FEO_OK;
}

struct SmemLayout {
  uint32_t buf_bytes, ring_off, bar_off, total;
};
SmemLayout smem_layout(const DevTilePlan& T) {
  SmemLayout L;
  L.buf_bytes = ((uint32_t)T.max_lines * kLineBytes + 1023u) / 1024u * 1024u;
  L.ring_off = (uint32_t)T.stages * L.buf_bytes;
  L.bar_off = L.ring_off + (uint32_t)T.warps * kRingBytes;
  L.total = L.bar_off + 64 + (uint32_t)T.warps * (kRingChunks * 8);
  return L;
}

int debug_mode() {
  const char* s = std::getenv("FEO_DEBUG_MODE");
  return s != nullptr ? atoi(s) : 0;
}

}  // namespace
int sm_count(int* out) {
  static int cached[64] = {0};  // per device ordinal: handles of several devices may live in one process
  int dev = 0, n = 0;
  FEO_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || cached[dev] == 0) {
    FEO_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    n = n > 0 ? n : 1;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    *out = n;
    return FEO_OK;
  }
  *out = cached[dev];
  return FEO_OK;
}
// one map over the dof-major array base[n][ldb] with a box of kSlab samples x box_rows dofs (lattice plan: one box per lattice row)
int make_box_map(const float* base, int64_t ldb, int32_t n, int32_t box_rows, CUtensorMap* out) {
  EncodeTiledFn enc;
  if (int rc = get_encode(&enc)) return rc;
  int dev = 0;
  FEO_CUDA_CHECK(cudaGetDevice(&dev));
  FEO_CUDA_CHECK(cudaSetDevice(dev));
  if (box_rows < 1 || box_rows > 256) return fail(FEO_ERR_INVALID_ARGUMENT, "TMA box height out of range");
  const cuuint64_t dims[2] = {(cuuint64_t)ldb, (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)ldb * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kSlab, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FEO_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return FEO_OK;
}
// a contiguous fp32 array seen as rows of `row_floats` (<= 256) floats, box = box_rows whole rows: bulk movement of pre-tiled
// blocks through the tensor-map path (the CTA-pair loads of feo_dense_tc.cu need a tensor map to count on the peer's mbarrier)
int make_row_map(const float* base, int32_t row_floats, int64_t rows, int32_t box_rows, CUtensorMap* out) {
  EncodeTiledFn enc;
  if (int rc = get_encode(&enc)) return rc;
  if (row_floats < 4 || row_floats > 256 || box_rows < 1 || box_rows > 256 || rows < 1) return fail(FEO_ERR_INVALID_ARGUMENT, "TMA row map out of range");
  const cuuint64_t dims[2] = {(cuuint64_t)row_floats, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_floats * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)row_floats, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FEO_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return FEO_OK;
}
namespace {

// persistent launch: one CTA per SM (or per unit when there are fewer units), `warps` consumer warps + 1 producer warp
template <typename Kernel>
int launch_persistent(Kernel k15, Kernel k19, const DevTilePlan& T, const SmemLayout& L, int grid, const TensorMaps& maps,
                      const TiledParams& p, cudaStream_t st) {
  if (T.warps > 19) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan has more warps than the kernels are built for");
  if (L.total > kSmemMax) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan needs more shared memory than an SM has");
  Kernel k = T.warps <= 15 ? k15 : k19;  // 16 warps x 128 registers or 20 warps x 96
  FEO_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  k<<<(unsigned)grid, (unsigned)(T.warps + 1) * 32, L.total, st>>>(maps, p);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

}  // namespace

size_t fused_partials_needed(int32_t warps) { return (size_t)1024 * (size_t)(warps > 0 ? warps : 1); }  // <= 1024 SMs

int launch_residual_fwd(const feo_operator* op, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                        float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (op->lattice.present) return launch_lattice_fwd(op, op->lattice, alphaT, fT, ldb, B, loss_out, rT, ws, ws_bytes, st);
  if (op->patch_f.present) return launch_patch_fwd(op, op->patch_f, alphaT, fT, ldb, B, loss_out, rT, ws, ws_bytes, st);
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  if (int rc = check_layout(fT, ldb, B, "fT")) return rc;
  if (rT != nullptr)
    if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (loss_out == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "loss_out is NULL");
  const DevTilePlan& T = op->tiles_f;
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)T.n_tiles * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "too many work units");
  int sms = 1;
  if (int rc = sm_count(&sms)) return rc;
  const int grid = (int)std::min<int64_t>(count, sms);
  const size_t n_partials = (size_t)grid * T.warps;
  if (ws == nullptr || ws_bytes < n_partials * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  TensorMaps maps;
  if (int rc = make_maps(alphaT, ldb, op->n, maps.m[0])) return rc;
  for (int c = 0; c < kBoxClasses; ++c) maps.m[1][c] = maps.m[0][c];
  const SmemLayout L = smem_layout(T);
  TiledParams p{T.tile_box_ptr, T.tile_lines, T.boxes, T.warp_range, reinterpret_cast<const int4*>(T.stream), fT, rT, (float*)ws,
                nullptr, ldb, B, n_slabs, (int32_t)count, grid / n_slabs, grid % n_slabs, T.warps, L.buf_bytes, T.stages, debug_mode(), op->ns_branch, 0.f, L.ring_off,
                L.bar_off};
  if (int rc = launch_persistent(residual_fwd_tiled<15>, residual_fwd_tiled<19>, T, L, grid, maps, p, st)) return rc;
  return finalize_loss((float*)ws, (int)n_partials, 1.0f, loss_out, st);
}

int launch_residual_bwd(const feo_operator* op, const float* alphaT, const float* rT, const float* grad_loss,
                        float* gradT, int64_t ldb, int32_t B, cudaStream_t st) {
  if (op->lattice.present) return launch_lattice_bwd(op, op->lattice, alphaT, rT, grad_loss, gradT, ldb, B, st);
  if (op->patch_b.present) return launch_patch_bwd(op, op->patch_b, alphaT, rT, grad_loss, gradT, ldb, B, st);
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (int rc = check_layout(gradT, ldb, B, "gradT")) return rc;
  if (op->has_conv)
    if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  const DevTilePlan& T = op->tiles_b;
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)T.n_tiles * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "too many work units");
  int sms = 1;
  if (int rc = sm_count(&sms)) return rc;
  const int grid = (int)std::min<int64_t>(count, sms);
  TensorMaps maps;
  if (int rc = make_maps(rT, ldb, op->n, maps.m[0])) return rc;
  if (int rc = make_maps(op->has_conv ? alphaT : rT, ldb, op->n, maps.m[1])) return rc;
  const SmemLayout L = smem_layout(T);
  TiledParams p{T.tile_box_ptr, T.tile_lines, T.boxes, T.warp_range, reinterpret_cast<const int4*>(T.stream), nullptr, gradT, nullptr,
                grad_loss, ldb, B, n_slabs, (int32_t)count, grid / n_slabs, grid % n_slabs, T.warps, L.buf_bytes, T.stages, debug_mode(),
                op->ns_branch, op->ns_branch ? 1.0f : -1.0f, L.ring_off, L.bar_off};
  return launch_persistent(residual_bwd_tiled<15>, residual_bwd_tiled<19>, T, L, grid, maps, p, st);
}

}  // namespace feo
