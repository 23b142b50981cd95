// sm_100a kernels of the fused sparse residual (forward: r, loss; backward: grad alpha).
//
// One CTA = one tile of operator rows (forward) / columns (backward) x one slab of 64 samples
// (feo_internal.h, feo_tiles.cpp).  Data path of a CTA:
//   1. the tile's dof lines (64 samples x 4 B each) are staged in shared memory by 2-D TMA boxes
//      (cp.async.bulk.tensor) straight from the dof-major batch arrays -- no registers, no LSU slots;
//   2. every warp streams its private list of 16-byte operator words through a 2 x 512 B shared
//      memory ring filled by 1-D bulk copies (cp.async.bulk) that complete on per-slot mbarriers;
//   3. gathers are conflict-free LDS.128 from the staged lines, the arithmetic is packed fp32
//      (fma.rn.f32x2: two IEEE fp32 FMAs per instruction).
// The load-store pipe (128 B/clk of shared-memory data per SM) is what bounds these kernels, so the
// mapping minimises its use: forward = one row per quarter-warp, 2 x 4 samples per lane (10 LSU cycles per
// 4 row entries); backward = one column PAIR per half-warp, 4 samples per lane, so that the gathers of
// r[I], r[J], alpha[I], alpha[J] of a neighbour node serve both columns.
// Everything is row-/column-owned with a fixed summation order: no atomics, bit-reproducible.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "feo_internal.h"

namespace feo {
namespace {

typedef unsigned long long u64;
constexpr int kMaxWarps = 10;  // warps per CTA the kernels are compiled for (2 CTAs per SM, <= 102 registers)

struct TensorMaps {
  CUtensorMap m[2][3];  // [source array][box class]
};

struct TiledParams {
  const int32_t* tile_box_ptr;
  const int32_t* tile_lines;
  const StageBox* boxes;
  const WarpRange* warp_range;
  const int4* stream;
  const float* fT;         // forward: load vectors
  float* outT;             // forward: rT (may be NULL) ; backward: gradT
  float* partials;         // forward: one loss partial per CTA
  const float* grad_loss;  // backward: upstream gradient (NULL = 1)
  int64_t ldb;
  int32_t B, n_slabs;
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

// The per-warp operator stream: a ring of kRingChunks slots of kChunkWords 16-byte words in shared
// memory (base aligned to the ring size), filled by 1-D bulk copies that complete on per-slot
// mbarriers.  Readers keep a byte offset `ptr` into the ring; stream items never straddle a chunk
// pair boundary the reader does not check (see the kernels), so the bookkeeping is one test per item.
constexpr uint32_t kChunkBytes = kChunkWords * 16;
constexpr uint32_t kRingBytes = kRingChunks * kChunkBytes;
struct Ring {
  uint32_t base, bar;  // shared addresses of the slots / their mbarriers
  const int4* src;     // the warp's words in global memory
  int32_t n_chunks;    // chunks of this warp's stream
  int32_t chunk;       // chunks entered so far
  uint32_t ptr;        // ring byte offset of the next word to read

  __device__ __forceinline__ void issue(int32_t c) const {
    const uint32_t slot = (uint32_t)c & (kRingChunks - 1);
    mbar_expect_tx(bar + slot * 8, kChunkBytes);
    bulk_copy(base + slot * kChunkBytes, src + (size_t)c * kChunkWords, kChunkBytes, bar + slot * 8);
  }
  __device__ __forceinline__ void start(uint32_t base_, uint32_t bar_, const int4* src_, int32_t n_words, int lane) {
    base = base_;
    bar = bar_;
    src = src_;
    n_chunks = (n_words + kChunkWords - 1) / kChunkWords;
    chunk = 0;
    ptr = 0;
    if (lane == 0)
      for (int32_t c = 0; c < kRingChunks && c < n_chunks; ++c) issue(c);
  }
  // The reader is about to read the first word of the next chunk: wait until it has landed; the chunk
  // before it is fully consumed, so its slot is refilled with the chunk kRingChunks - 1 ahead.
  __device__ __forceinline__ void enter(int lane) {
    const int32_t c = chunk++;
    mbar_wait(bar + ((uint32_t)c & (kRingChunks - 1)) * 8, ((uint32_t)c / kRingChunks) & 1u);
    if (c >= 1) {
      __syncwarp();
      const int32_t nc = c + kRingChunks - 1;
      if (nc < n_chunks && lane == 0) issue(nc);
    }
  }
  __device__ __forceinline__ bool at_chunk_start() const { return (ptr & (kChunkBytes - 1)) == 0; }
  __device__ __forceinline__ void skip(uint32_t bytes) { ptr = (ptr + bytes) & (kRingBytes - 1); }
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

// Shared prologue: barriers, line staging, ring start.  Returns when the tile's lines have landed.
template <typename RingT>
__device__ __forceinline__ int32_t stage_tile(const TensorMaps& maps, const TiledParams& p, uint32_t sb, int tile, int slab, int warp,
                                              int lane, RingT& ring) {
  const uint32_t bar_lines = sb + p.bar_off;
  const uint32_t bar_ring = bar_lines + 8 + (uint32_t)warp * (kRingChunks * 8);
  if (threadIdx.x == 0) {
    mbar_init(bar_lines, 1);
    mbar_expect_tx(bar_lines, (uint32_t)p.tile_lines[tile] * kLineBytes);
  }
  if (lane == 0)
    for (int c = 0; c < kRingChunks; ++c) mbar_init(bar_ring + c * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int b0 = p.tile_box_ptr[tile], b1 = p.tile_box_ptr[tile + 1];
  for (int b = b0 + (int)threadIdx.x; b < b1; b += (int)blockDim.x) {
    const StageBox bx = p.boxes[b];
    tma_box(sb + (uint32_t)bx.line0 * kLineBytes, &maps.m[bx.src][bx.cls], slab * kSlab, bx.dof0, bar_lines);
  }
  const WarpRange wr = p.warp_range[(size_t)tile * (blockDim.x >> 5) + warp];
  ring.start(sb + p.ring_off + (uint32_t)warp * kRingBytes, bar_ring, p.stream + wr.begin, wr.n_words, lane);
  mbar_wait(bar_lines, 0);
  return wr.n_words;
}

// ---------------------------------------------------------------------------------------------
// forward: r = A a -/+ (F - c), loss partial = sum r^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxWarps * 32, 2) residual_fwd_tiled(const __grid_constant__ TensorMaps maps, const TiledParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ float s_part[32];
  const uint32_t sb = smem_u32(smem);
  const int tile = blockIdx.x / p.n_slabs, slab = blockIdx.x - tile * p.n_slabs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Ring ring;
  const int ring_units = stage_tile(maps, p, sb, tile, slab, warp, lane, ring) / 4;

  const int q = lane >> 3;  // this lane's row slot in the quad
  // a lane owns samples [4l, 4l+4) and [32+4l, 32+4l+4), l = lane & 7: each LDS.128 of a quarter-warp then
  // reads 128 contiguous bytes of one line (conflict-free), the second one 128 B further
  const uint32_t lines = sb + (uint32_t)(lane & 7) * 16;
  const uint32_t rq = ring.base + (uint32_t)q * 16;  // this quarter's word within a 64-byte unit
  const int b0 = slab * kSlab + (lane & 7) * 4;
  const bool precond = p.precond != 0;
  float lsum = 0.f;

  // stream = quads: [header unit][spare unit][n_steps step units], n_steps even; units are 64 B, so
  // a pair of units never straddles a 512-byte chunk
  int units_left = ring_units;
  while (units_left > 0) {
    if (ring.at_chunk_start()) ring.enter(lane);
    const int4 hdr = lds_word(rq + ring.ptr);
    ring.skip(128);
    const int n_steps = hdr.y;
    const int row = hdr.x;
    units_left -= 2 + n_steps;
    // the load vector of this row is fetched now and consumed in the epilogue (hides the DRAM latency)
    float4 fv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      fv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row >= 0 && b0 + 32 * k < p.ldb) fv[k] = ldg4_stream(p.fT + (int64_t)row * p.ldb + b0 + 32 * k);
    }
    P2 accA[4], acc1[4], acc2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) accA[i] = acc1[i] = acc2[i] = P2{0.f, 0.f};
    // step pairs, in segments that end at a chunk boundary so that the inner loop carries no ring logic
    for (int s = 0; s < n_steps;) {
      if (ring.at_chunk_start()) ring.enter(lane);
      const int seg = min(n_steps - s, (int)((kChunkBytes - (ring.ptr & (kChunkBytes - 1))) / 64));
      s += seg;
      const uint32_t rp = rq + ring.ptr;
      ring.skip((uint32_t)seg * 64);
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
        for (int i = 0; i < 4; ++i) {
          fma2s(accA[i], __int_as_float(e0.y), x0[i]);
          fma2s(acc1[i], __int_as_float(e0.z), x0[i]);
          fma2s(acc2[i], __int_as_float(e0.w), x0[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fma2s(accA[i], __int_as_float(e1.y), x1[i]);
          fma2s(acc1[i], __int_as_float(e1.z), x1[i]);
          fma2s(acc2[i], __int_as_float(e1.w), x1[i]);
        }
      }
    }
    // epilogue: convection product, load vector, residual, loss, store
    if (row >= 0) {
      float lhs[8], c[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lhs[2 * i] = accA[i].lo;
        lhs[2 * i + 1] = accA[i].hi;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = 0.f;
      if (hdr.w & 1) {
        u64 d1[4], d2[4];
        const uint32_t li = ((uint32_t)hdr.z & 0xffffu) * kLineBytes, lj = ((uint32_t)hdr.z >> 16) * kLineBytes;
        lds_pairs(lines + li, d1[0], d1[1]);
        lds_pairs(lines + li + 128, d1[2], d1[3]);
        lds_pairs(lines + lj, d2[0], d2[1]);
        lds_pairs(lines + lj + 128, d2[2], d2[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float u0, u1, v0, v1;
          unpk(d1[i], u0, u1);
          unpk(d2[i], v0, v1);
          c[2 * i] = conv1(u0, acc1[i].lo, v0, acc2[i].lo);
          c[2 * i + 1] = conv1(u1, acc1[i].hi, v1, acc2[i].hi);
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
  // fixed-order block reduction of the loss partial
  lsum = warp_sum(lsum);
  if (lane == 0) s_part[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) t += s_part[w];
    p.partials[blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// backward: grad = 2 g [A^T r + s (B1^T (d1 r) + B2^T (d2 r) + E-term)], column-pair owned
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxWarps * 32, 2) residual_bwd_tiled(const __grid_constant__ TensorMaps maps, const TiledParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sb = smem_u32(smem);
  const int tile = blockIdx.x / p.n_slabs, slab = blockIdx.x - tile * p.n_slabs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Ring ring;
  const int n_words = stage_tile(maps, p, sb, tile, slab, warp, lane, ring);

  const int h = lane >> 4;                                   // this lane's pair slot in the duo
  const uint32_t lines = sb + (uint32_t)(lane & 15) * 16;    // 4 samples = 16 B of every line
  const uint32_t rh = ring.base + (uint32_t)h * 16;          // this half's word within a word pair
  const int b0 = slab * kSlab + (lane & 15) * 4;
  const float g2 = 2.0f * (p.grad_loss != nullptr ? __ldg(p.grad_loss) : 1.0f);

  // The stream is walked by linear word position `pos`; pieces (header 4 words, V-step 6, A-step 2,
  // X-step 4) never straddle a 32-word chunk: when the next piece does not fit, both the plan builder
  // and this reader skip to the next chunk.
  int pos = 0;
  auto place = [&](int len) {  // returns the ring address of the piece, entering a new chunk if needed
    if ((pos & (kChunkWords - 1)) + len > kChunkWords) pos = (pos + kChunkWords - 1) & ~(kChunkWords - 1);
    if ((pos & (kChunkWords - 1)) == 0) ring.enter(lane);
    return rh + (((uint32_t)pos & (kRingChunks * kChunkWords - 1)) << 4);
  };
  while (pos < n_words) {
    const uint32_t ha = place(4);
    const int4 h0 = lds_word(ha);
    const int4 h1 = lds_word(ha + 32);
    pos += 4;
    const int nV = h0.z, nA = h0.w, nX = h1.x;
    P2 accI[2], accJ[2], bu1I[2], bu2I[2], bu1J[2], bu2J[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) accI[i] = accJ[i] = bu1I[i] = bu2I[i] = bu1J[i] = bu2J[i] = P2{0.f, 0.f};
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
        for (int i = 0; i < 2; ++i) {
          const u64 tI = fma2r(b2I, d2[i], fma2r(b1I, d1[i], aI));
          const u64 tJ = fma2r(b2J, d2[i], fma2r(b1J, d1[i], aJ));
          fma2p(accI[i], tI, rI[i]);
          fma2p(accJ[i], tJ, rJ[i]);
          fma2s(bu1I[i], __int_as_float(w2.x), d1[i]);
          fma2s(bu2I[i], __int_as_float(w2.y), d1[i]);
          fma2s(bu1J[i], __int_as_float(w2.z), d2[i]);
          fma2s(bu2J[i], __int_as_float(w2.w), d2[i]);
        }
      }
    }
    for (int a = 0; a < nA;) {
      const uint32_t wa = place(2);
      const int cnt = min(nA - a, (kChunkWords - (pos & (kChunkWords - 1))) / 2);
      a += cnt;
      pos += 2 * cnt;
#pragma unroll 1
      for (int t = 0; t < cnt; ++t) {
        const int4 w0 = lds_word(wa + (uint32_t)t * 32);
        u64 rI[2], rJ[2];
        lds_pairs(lines + ((uint32_t)w0.x & 0xffffu) * kLineBytes, rI[0], rI[1]);
        lds_pairs(lines + ((uint32_t)w0.x >> 16) * kLineBytes, rJ[0], rJ[1]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          fma2s(accI[i], __int_as_float(w0.y), rI[i]);
          fma2s(accJ[i], __int_as_float(w0.z), rJ[i]);
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
      for (int i = 0; i < 2; ++i) {
        fma2s(bu1I[i], __int_as_float(w0.y), xv[i]);
        fma2s(bu2I[i], __int_as_float(w0.z), xv[i]);
        fma2s(bu1J[i], __int_as_float(w0.w), xv[i]);
        fma2s(bu2J[i], __int_as_float(w1.x), xv[i]);
      }
    }
    // epilogue: E-term, scale, store
    float oI[4] = {accI[0].lo, accI[0].hi, accI[1].lo, accI[1].hi};
    float oJ[4] = {accJ[0].lo, accJ[0].hi, accJ[1].lo, accJ[1].hi};
    if (h1.z & 1) {
      const float4 rI = *reinterpret_cast<const float4*>(smem + ((uint32_t)h1.y & 0xffffu) * kLineBytes + (lane & 15) * 16);
      const float4 rJ = *reinterpret_cast<const float4*>(smem + ((uint32_t)h1.y >> 16) * kLineBytes + (lane & 15) * 16);
      const float ri[4] = {rI.x, rI.y, rI.z, rI.w}, rj[4] = {rJ.x, rJ.y, rJ.z, rJ.w};
      const float s1i[4] = {bu1I[0].lo, bu1I[0].hi, bu1I[1].lo, bu1I[1].hi}, s2i[4] = {bu2I[0].lo, bu2I[0].hi, bu2I[1].lo, bu2I[1].hi};
      const float s1j[4] = {bu1J[0].lo, bu1J[0].hi, bu1J[1].lo, bu1J[1].hi}, s2j[4] = {bu2J[0].lo, bu2J[0].hi, bu2J[1].lo, bu2J[1].hi};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        oI[i] = fmaf(p.esign, fmaf(s1j[i], rj[i], s1i[i] * ri[i]), oI[i]);
        oJ[i] = fmaf(p.esign, fmaf(s2j[i], rj[i], s2i[i] * ri[i]), oJ[i]);
      }
    }
    if (b0 < p.B) {
      const int cI = h0.x, cJ = h0.y;
      if (cI >= 0) *reinterpret_cast<float4*>(p.outT + (int64_t)cI * p.ldb + b0) = make_float4(oI[0] * g2, oI[1] * g2, oI[2] * g2, oI[3] * g2);
      if (cJ >= 0) *reinterpret_cast<float4*>(p.outT + (int64_t)cJ * p.ldb + b0) = make_float4(oJ[0] * g2, oJ[1] * g2, oJ[2] * g2, oJ[3] * g2);
    }
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
int make_maps(const float* base, int64_t ldb, int32_t n, CUtensorMap out[3]) {
  EncodeTiledFn enc;
  if (int rc = get_encode(&enc)) return rc;
  // the driver entry point needs the primary context current on THIS thread (autograd runs the backward
  // on its own threads, where no runtime call may have bound it yet)
  int dev = 0;
  FEO_CUDA_CHECK(cudaGetDevice(&dev));
  FEO_CUDA_CHECK(cudaSetDevice(dev));
  for (int cls = 0; cls < 3; ++cls) {
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
  uint32_t ring_off, bar_off, total;
};
SmemLayout smem_layout(const DevTilePlan& T) {
  SmemLayout L;
  L.ring_off = ((uint32_t)T.max_lines * kLineBytes + kRingBytes - 1) / kRingBytes * kRingBytes;  // rings are size aligned
  L.bar_off = L.ring_off + (uint32_t)T.warps * kRingBytes;
  L.total = L.bar_off + 8 + (uint32_t)T.warps * (kRingChunks * 8) + 8;
  return L;
}

}  // namespace

int launch_residual_fwd(const feo_operator* op, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                        float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  if (int rc = check_layout(fT, ldb, B, "fT")) return rc;
  if (rT != nullptr)
    if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (loss_out == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "loss_out is NULL");
  const DevTilePlan& T = op->tiles_f;
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)T.n_tiles * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "grid too large");
  if (ws == nullptr || ws_bytes < (size_t)count * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  TensorMaps maps;
  if (int rc = make_maps(alphaT, ldb, op->n, maps.m[0])) return rc;
  for (int c = 0; c < 3; ++c) maps.m[1][c] = maps.m[0][c];
  const SmemLayout L = smem_layout(T);
  if (T.warps > kMaxWarps) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan has more warps than the kernels are built for");
  TiledParams p{T.tile_box_ptr, T.tile_lines, T.boxes, T.warp_range, reinterpret_cast<const int4*>(T.stream), fT, rT, (float*)ws,
                nullptr, ldb, B, n_slabs, op->ns_branch, 0.f, L.ring_off, L.bar_off};
  FEO_CUDA_CHECK(cudaFuncSetAttribute(residual_fwd_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  residual_fwd_tiled<<<(unsigned)count, T.warps * 32, L.total, st>>>(maps, p);
  FEO_CUDA_CHECK(cudaGetLastError());
  return finalize_loss((float*)ws, (int)count, 1.0f, loss_out, st);
}

int launch_residual_bwd(const feo_operator* op, const float* alphaT, const float* rT, const float* grad_loss,
                        float* gradT, int64_t ldb, int32_t B, cudaStream_t st) {
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(rT, ldb, B, "rT")) return rc;
  if (int rc = check_layout(gradT, ldb, B, "gradT")) return rc;
  if (op->has_conv)
    if (int rc = check_layout(alphaT, ldb, B, "alphaT")) return rc;
  const DevTilePlan& T = op->tiles_b;
  const int32_t n_slabs = (B + kSlab - 1) / kSlab;
  const int64_t count = (int64_t)T.n_tiles * n_slabs;
  if (count >= ((int64_t)1 << 31)) return fail(FEO_ERR_UNSUPPORTED, "grid too large");
  TensorMaps maps;
  if (int rc = make_maps(rT, ldb, op->n, maps.m[0])) return rc;
  if (int rc = make_maps(op->has_conv ? alphaT : rT, ldb, op->n, maps.m[1])) return rc;
  const SmemLayout L = smem_layout(T);
  if (T.warps > kMaxWarps) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan has more warps than the kernels are built for");
  TiledParams p{T.tile_box_ptr, T.tile_lines, T.boxes, T.warp_range, reinterpret_cast<const int4*>(T.stream), nullptr, gradT, nullptr,
                grad_loss, ldb, B, n_slabs, op->ns_branch, op->ns_branch ? 1.0f : -1.0f, L.ring_off, L.bar_off};
  FEO_CUDA_CHECK(cudaFuncSetAttribute(residual_bwd_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  residual_bwd_tiled<<<(unsigned)count, T.warps * 32, L.total, st>>>(maps, p);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

}  // namespace feo
