// Dense (preconditioned) operator apply on the sm_100a tensor cores:
//   CT[n x ldb] = scale * D[n x n] XT[n x ldb]  (- sub)  (+ sum of squares)
// replaces  torch.matmul(A @ P, alpha^T)  of the reference's preconditioned weak forms
// (FEONet_Stokes_square/train_FEONet.py:264,299; steady NS :325,:363; time-dependent :347,:403).
//
// tcgen05.mma kind::tf32 with the error-compensated three-product split ("3xTF32"): every fp32 operand is
// written to shared memory as hi = rna_tf32(x) and lo = rna_tf32(x - hi), and the accumulator in tensor
// memory receives  lo*hi + hi*lo + hi*hi.  The dropped lo*lo term and the rounding of lo are O(2^-22)
// relative per product and unbiased, i.e. fp32-grade, which plain TF32 (2^-11) is not: the north-star
// tolerance on the loss is 1e-5 relative.
//
// One CTA = one 128 x 64 tile of CT (UMMA M = 128, N = 64, K = 8; 128-column tiles selectable), k-blocks of 16.
//  * D operand: split ONCE at set-up (dense_split_tiles: per (row tile, k-block) a 16 KB block [hi | lo] already in
//    the shared-memory operand layout), so a stage is one cp.async.bulk (UBLKCP) onto an mbarrier; SA = 3 stages, the
//    copy of k-block kb + SA - 2 is issued as soon as the MMAs of kb - 2 have released its stage.
//  * XT operand (the activations: split at run time; sample-contiguous, i.e. MN-major): through registers --
//    coalesced row reads, cvt.rna.tf32 split, STS.128 straight into the K-major core-matrix layout, which also
//    transposes it; two stages, two register sets (k-blocks kb + 1 and kb + 2 in flight).
// Warp roles (288 threads): eight loader / epilogue warps and one ISSUER warp.  Loaders: wait until the MMAs of
// k-block kb - 2 have released XT stage kb % 2 (commit mbarrier), split + STS, fence.proxy.async, arrive on the
// stage's "written" mbarrier, issue the loads of kb + 2.  Issuer (warp-uniform code, incremental stage / phase /
// descriptor state on the uniform datapath, one elected lane issues): bulk copy of the D stage SA - 2 k-blocks
// ahead, wait for "written" and for the D stage, six tcgen05.mma, tcgen05.commit onto the commit mbarrier.  With one
// thread doing both jobs behind a bar.sync, ~480 cycles of MMA issue per k-block (ELECT / R2UR broadcast loops around
// every UTCHMMA under a `tid == 0` guard, descriptors rebuilt from byte addresses) sat on top of ~500 cycles of
// loader work.  64 KB of shared memory and 64 TMEM columns per CTA: three CTAs per SM.
// Optional (FEO_DENSE_CLUSTER=2): two column tiles form a cluster, each CTA fetches one half of every D stage and
// multicasts it, commits are multicast to both CTAs' barriers (stages are released together).
//
// The tensor core aligns and TRUNCATES when it adds a K = 8 product group to the fp32 accumulator, which biases a
// long accumulation towards zero (measured: 3e-6 of |D||x| at n = 2549, against 1e-7 for the split itself).  The
// accumulator is therefore drained every `flush` k-blocks (default 4 = 64 k, 24 accumulations; 7 % of the run time at
// N = 2549) into fp32 registers with round-to-nearest adds -- the epilogue's own TMEM mapping, 32 registers per
// thread -- and restarted with accumulate = 0 once every loader thread has arrived on the "drained" mbarrier.
//
// Shared-memory operand layout (canonical K-major, SWIZZLE_NONE): element (r, k) of a [128 x 16] tile lives at
//   (k / 4) * 2048 + r * 16 + (k % 4) * 4   bytes,
// i.e. 8 x 16-byte core matrices, 128 B each, SBO (next 8 rows) = 128 B, LBO (next 4 k) = 2048 B.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "feo_internal.h"

namespace feo {
int make_row_map(const float* base, int32_t row_floats, int64_t rows, int32_t box_rows, CUtensorMap* out);  // feo_tiled.cu
namespace {

constexpr int TBM = 128, TBK = 16;                      // CT tile rows, k-block; tile columns BN = 128 or 64 (template)
constexpr int kTcThreads = 256;                         // loader / epilogue threads (8 warps)
constexpr int kTcBlock = kTcThreads + 32;                // + the issuer warp
constexpr uint32_t kABytes = TBM * TBK * 4;              // 8 KB: A_hi or A_lo of a stage
constexpr int kTcStages = 2;                             // XT operand stages (register pass)
__host__ __device__ constexpr uint32_t b_stage_bytes(int BN) { return 2 * (uint32_t)BN * TBK * 4; }  // B_hi, B_lo
__host__ __device__ constexpr uint32_t smem_bytes_for(int BN, int SA) { return SA * 2 * kABytes + kTcStages * b_stage_bytes(BN); }
constexpr int kFlushDefault = 4;                         // k-blocks between drains of the TMEM accumulator (FEO_DENSE_FLUSH)
constexpr uint32_t kLboA = TBM * 16;                     // 2048 B between the k-chunks (4 k each) of an A tile; BN * 16 for B
constexpr uint32_t kSbo = 128;                           // 8-row groups are contiguous

// instruction descriptor of tcgen05.mma kind::tf32 (cute::UMMA::InstrDescriptor bit layout):
// c_format F32 (bit 4), a/b_format TF32 = 2 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc(int BN) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, LBO, SBO (all >> 4), version 1 (sm_100), no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// one lane of a converged warp (CUTLASS' elect_one_sync): the compiler keeps the guarded code on the uniform datapath,
// where tcgen05.mma / cp.async.bulk take their descriptors from -- a per-thread `tid == 0` guard makes it broadcast
// every operand with an ELECT / R2UR loop instead (measured: ~110 cycles per MMA issued)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// the same 1-D bulk copy delivered to the same CTA-relative address (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void bulk_copy_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier receives one arrival when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the arrival goes to the same mbarrier of every CTA in `mask` (the cluster's CTAs release shared operand stages together)
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// ---- CTA-pair (cta_group::2) helpers --------------------------------------------------------------------------------
// the address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking: has the phase completed?
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquires what a peer CTA released
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // a shared::cta address with this bit cleared names the same offset in the pair's leader
// the same load delivered to the same offset of every CTA in `mask` (cluster ranks), counted on the full barrier of each
// destination's pair leader
__device__ __forceinline__ void tma_box_pair_multicast(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar & kPeerBitMask), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_mask(uint32_t bar, uint16_t mask) {  // one arrival on the same mbarrier of every CTA in `mask`
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// 2-D tensor-map load of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction is counted on the mbarrier at
// `leader_bar` (a shared::cluster address: the leader's barrier) -- the hardware hand-over that needs no software relay
__device__ __forceinline__ void tma_box_pair(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
               : "memory");
}
// one M = 256 MMA over the pair: each CTA contributes its own 128 rows of A and its half (N / 2 rows) of the B tile from the
// same shared-memory offsets, and receives its 128 accumulator rows in its own tensor memory; issued by the leader CTA only
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // one arrival on the same mbarrier of both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// x = hi + lo (+ O(2^-22 |x|)), both exactly representable in tf32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
__device__ __forceinline__ void store_split(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4& v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

template <int BN, int SA, int CL>  // CT tile columns; stages of the D operand ring (its copies run SA - 2 k-blocks ahead);
                                   // CL = 2: two column tiles form a cluster and share every D stage (each CTA fetches one half, multicast)
__global__ void __launch_bounds__(kTcBlock, BN == 128 ? 2 : 3) dense_apply_tc_kernel(const float* __restrict__ Dsplit, int32_t n,
                                                                    const float* __restrict__ XT, float* __restrict__ CT,
                                                                    int64_t ldb, int32_t B, float scale,
                                                                    const float* __restrict__ scale_dev,
                                                                    const float* __restrict__ sub,
                                                                    float* __restrict__ partials, int32_t flush, int32_t debug) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar[kTcStages];   // commit barriers: MMAs of a k-block done (by k-block parity)
  __shared__ __align__(8) uint64_t s_full[SA];          // D operand stage landed
  __shared__ __align__(8) uint64_t s_bfull[kTcStages];  // XT stage written by all loader threads
  __shared__ __align__(8) uint64_t s_drained;           // accumulator chunk read back by all loader threads
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[kTcThreads / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = __shfl_sync(0xffffffffu, warp, 0) == kTcThreads / 32;  // warp-uniform by construction: keeps the issue code uniform
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  constexpr uint32_t kTmemCols = BN;                      // fp32 accumulator columns (power of two >= 32)
  constexpr uint32_t kBBytes = (uint32_t)BN * TBK * 4;    // B_hi or B_lo of a stage
  constexpr uint32_t kAStage = 2 * kABytes;               // [hi | lo]
  constexpr uint32_t kBStage = b_stage_bytes(BN);
  constexpr uint32_t kBBase = SA * kAStage;
  constexpr int kAhead = SA - 2;                          // the stage of k-block kb + kAhead was last read by kb - 2
  constexpr uint32_t kLboB = BN * 16;
  constexpr uint32_t kIdesc = instr_desc(BN);
  constexpr int CB = BN / 64;                             // k-chunks of the B tile per thread
  constexpr int WC = BN / 2;                              // accumulator columns per warp

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kTcStages; ++s) mbar_init(smem_u32(&s_bar[s]), CL);  // one commit per CTA of the cluster
    for (int s = 0; s < SA; ++s) mbar_init(smem_u32(&s_full[s]), 1);
    for (int s = 0; s < kTcStages; ++s) mbar_init(smem_u32(&s_bfull[s]), kTcThreads);
    mbar_init(smem_u32(&s_drained), kTcThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync();  // the peer's mbarriers exist before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  const int nkb = (n + TBK - 1) / TBK;
  // D operand: block (row tile, kb) of Dsplit -> A stage kb % SA, one bulk copy
  const float* a_src = Dsplit + (size_t)blockIdx.y * nkb * (kAStage / 4);

  // XT operand: thread -> (sample column, CB k-chunks): four coalesced row reads of XT per chunk
  const int b_c = tid % BN, b_q = (tid / BN) * CB;      // k-chunks b_q .. b_q + CB - 1
  const bool b_col_ok = n0 + b_c < ldb;
  const float* b_src = XT + n0 + b_c;
  // two register sets: the loads of k-blocks kb + 1 and kb + 2 are in flight while kb is split and multiplied
  float4 bv0[CB], bv1[CB];
  auto load_block = [&](int k0, float4 (&bv)[CB]) {
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      const int k = k0 + (b_q + j) * 4;
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = (b_col_ok && k + i < n) ? __ldg(b_src + (int64_t)(k + i) * ldb) : 0.f;
      bv[j] = make_float4(t[0], t[1], t[2], t[3]);
    }
  };

  // warp w owns TMEM lanes 32 (w % 4) .. +31 (its quarter) and columns WC (w / 4) .. +WC-1; lane = row of CT
  const uint32_t t_own = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * WC);
  float acc[WC];
#pragma unroll
  for (int i = 0; i < WC; ++i) acc[i] = 0.f;

  // ---- issuer warp: operator stages (bulk copies) and the MMAs --------------------------------------------------
  if (is_issuer) {
    // Everything below is warp-uniform and kept incremental (stage indices, phases, descriptor low words advance by
    // additions): the uniform datapath that feeds UTCHMMA / UBLKCP is narrow, and rebuilding descriptors from byte
    // addresses (shift, mask, or; `% SA`) cost more cycles per k-block than the six MMAs take to run.
    const uint32_t half = CL == 1 ? 0u : cluster_ctarank() * kABytes;        // cluster: this CTA's half of a D stage
    const uint32_t copy_bytes = CL == 1 ? kAStage : kABytes;
    auto copy_a = [&](int stage, int kb) {  // D block kb of this row tile -> stage
      const uint32_t bar = smem_u32(&s_full[0]) + (uint32_t)stage * 8;
      mbar_expect_tx(bar, kAStage);
      const uint32_t dst = smem_u32(smem) + (uint32_t)stage * kAStage + half;
      const float* src = a_src + (size_t)kb * (kAStage / 4) + half / 4;
      if (CL == 1) bulk_copy(dst, src, copy_bytes, bar);
      else bulk_copy_multicast(dst, src, copy_bytes, bar, (uint16_t)3);
    };
    if (elect_one())
      for (int kb = 0; kb < kAhead && kb < nkb; ++kb) copy_a(kb, kb);  // kAhead < SA: stage = kb
    int c_stage = kAhead;  // stage of the next copy (k-block kAhead)
    // descriptor words: lo = (address >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14
    constexpr uint32_t kDescHi = (kSbo >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((kLboA >> 4) << 16);
    const uint32_t b_lo0 = (((smem_u32(smem) + kBBase) & 0x3ffffu) >> 4) | ((kLboB >> 4) << 16);
    auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; };
    int a_stage = 0, a_phase = 0, chunk_pos = 0, drain_phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      const uint32_t bar_s = smem_u32(&s_bar[0]) + (uint32_t)s * 8;
      // the MMAs of k-block kb - 2 have completed (use j of a commit barrier waits for commit j - 1):
      // A stage (kb + kAhead) % SA is free
      if (kb >= kTcStages) mbar_wait(bar_s, ((kb >> 1) - 1) & 1);
      if (kb + kAhead < nkb) {
        if (elect_one()) copy_a(c_stage, kb + kAhead);
        if (++c_stage == SA) c_stage = 0;
      }
      mbar_wait(smem_u32(&s_bfull[0]) + (uint32_t)s * 8, (kb >> 1) & 1);    // the loaders have written XT stage s
      mbar_wait(smem_u32(&s_full[0]) + (uint32_t)a_stage * 8, a_phase);     // the D stage has landed
      const bool first = chunk_pos == 0;
      if (first && kb > 0) {
        mbar_wait(smem_u32(&s_drained), drain_phase);                       // accumulator read back by everyone
        drain_phase ^= 1;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t la = a_lo0 + (uint32_t)a_stage * (kAStage >> 4), lb = b_lo0 + (uint32_t)s * (kBStage >> 4);
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          const uint64_t a_hi = desc(la + ks * (2 * kLboA >> 4)), a_lo = desc(la + (kABytes >> 4) + ks * (2 * kLboA >> 4));
          const uint64_t b_hi = desc(lb + ks * (2 * kLboB >> 4)), b_lo = desc(lb + (kBBytes >> 4) + ks * (2 * kLboB >> 4));
          if (debug < 2) umma_tf32(tmem, a_lo, b_hi, kIdesc, !(first && ks == 0));  // first product of a chunk overwrites
          if (debug < 1) umma_tf32(tmem, a_hi, b_lo, kIdesc, 1);
          if (debug < 1) umma_tf32(tmem, a_hi, b_hi, kIdesc, 1);
        }
        if (CL == 1) umma_commit(bar_s);
        else umma_commit_multicast(bar_s, (uint16_t)3);
      }
      __syncwarp();
      if (++a_stage == SA) a_stage = 0, a_phase ^= 1;
      if (++chunk_pos == flush) chunk_pos = 0;
    }
  } else {
    // ---- loader warps: XT stages through registers, accumulator drains -----------------------------------------
    auto k_block = [&](int kb, float4 (&bv)[CB]) {
      const int s = kb & 1;
      uint8_t* stage_b = smem + kBBase + s * kBStage;
      // XT stage s was last read by the MMAs of k-block kb - 2
      if (kb >= kTcStages) mbar_wait(smem_u32(&s_bar[s]), ((kb >> 1) - 1) & 1);
#pragma unroll
      for (int j = 0; j < CB; ++j)
        store_split(stage_b, stage_b + kBBytes, (uint32_t)(b_q + j) * kLboB + (uint32_t)b_c * 16, bv[j]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
      mbar_arrive(smem_u32(&s_bfull[s]));
      // this register set is free again; the loads go out AFTER the proxy fence, which otherwise holds the thread until
      // they have returned from L2
      if (kb + 2 < nkb) load_block((kb + 2) * TBK, bv);
      if ((kb + 1) % flush == 0 || kb + 1 == nkb) {
        // drain: this k-block's commit covers every MMA issued so far
        mbar_wait(smem_u32(&s_bar[s]), (kb >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int j = 0; j < WC / 16; ++j) {
          uint32_t v[16];
          tmem_ld16(t_own + (uint32_t)(j * 16), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[j * 16 + i] += __uint_as_float(v[i]);
        }
        // orders these TMEM reads before the next chunk's overwrite: the issuer waits for all 256 arrivals
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(smem_u32(&s_drained));
      }
    };
    load_block(0, bv0);
    if (nkb > 1) load_block(TBK, bv1);
    for (int kb = 0; kb < nkb; kb += 2) {
      k_block(kb, bv0);
      if (kb + 1 < nkb) k_block(kb + 1, bv1);
    }
  }

  // epilogue
  const float sc = scale * (scale_dev != nullptr ? __ldg(scale_dev) : 1.0f);
  const int m = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * WC;
  float lsum = 0.f;
#pragma unroll
  for (int q = 0; q < WC / 4; ++q) {
    const int c = n0 + cbase + q * 4;
    if (!is_issuer && m < n && c < ldb) {
      float4 o = make_float4(sc * acc[q * 4 + 0], sc * acc[q * 4 + 1], sc * acc[q * 4 + 2], sc * acc[q * 4 + 3]);
      if (sub != nullptr) {
        const float4 sv = __ldg(reinterpret_cast<const float4*>(sub + (int64_t)m * ldb + c));
        o.x -= sv.x;
        o.y -= sv.y;
        o.z -= sv.z;
        o.w -= sv.w;
      }
      if (c + 0 < B) lsum = fmaf(o.x, o.x, lsum);
      if (c + 1 < B) lsum = fmaf(o.y, o.y, lsum);
      if (c + 2 < B) lsum = fmaf(o.z, o.z, lsum);
      if (c + 3 < B) lsum = fmaf(o.w, o.w, lsum);
      *reinterpret_cast<float4*>(CT + (int64_t)m * ldb + c) = o;
    }
  }
  // fixed-order CTA reduction of the loss partial
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0 && !is_issuer) s_part[warp] = lsum;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && partials != nullptr) {
    float t = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) t += s_part[w];
    partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
  if (CL > 1) cluster_sync();  // no CTA leaves while its peer may still signal it
}

// ---------------------------------------------------------------------------------------------------------------------
// Second generation (round 2): the activations are pre-split too.
//
// The first kernel is paced by its XT operand pipeline (global -> registers -> cvt.rna -> STS by eight warps: a run without
// any MMA takes 0.144 of its 0.162 ms at N = 2549, B = 1024).  Here a bandwidth-bound PRE-PASS (dense_split_x_kernel: XT is
// 10 MB, its split 20 MB) writes, per (column tile, k-block), one block [hi | lo] already in the K-major core-matrix layout,
// so that BOTH operands of a stage arrive by bulk copies (UBLKCP) on one transaction barrier and no warp touches operand
// data.  The accumulator ping-pongs between two TMEM regions: while the eight epilogue warps drain chunk c (tcgen05.ld,
// round-to-nearest adds in registers: the truncation drain) the issuer already accumulates chunk c + 1 into the other one.
//   barriers: full[S]       stage landed (24 KB of transactions: D block 16 KB + X block 8 KB at BN = 64)
//             mma_done[2]   per k-block parity: MMAs of k-block kb completed -> its stage may be refilled (issuer only)
//             chunk_done[2] per accumulator: every MMA of the chunk completed -> drain it (epilogue warps)
//             drained[2]    per accumulator: read back by all 256 epilogue threads -> the issuer may overwrite it
// ---------------------------------------------------------------------------------------------------------------------
template <int BN, int HALVES>
__global__ void __launch_bounds__(256) dense_split_x_kernel(const float* __restrict__ XT, int32_t n, int64_t ldb, float* __restrict__ Xs,
                                                            int32_t nkb) {
  // block (column tile ct, k-block kb) -> Xs + (ct * nkb + kb) * [hi BN x 16 | lo BN x 16], element (c, k) at
  // (k / 4) * BN * 16 + c * 16 + (k % 4) * 4 bytes of its half; HALVES = 2 (CTA pairs): the block is two such [hi | lo] blocks
  // of BN / 2 columns each, one per CTA of the pair
  constexpr int W = BN / HALVES;  // columns of one [hi | lo] sub-block
  const int ct = blockIdx.x, kb = blockIdx.y;
  uint8_t* blk = reinterpret_cast<uint8_t*>(Xs) + ((size_t)ct * nkb + kb) * b_stage_bytes(BN);
  constexpr uint32_t kHalf = (uint32_t)W * TBK * 4;
  for (int it = threadIdx.x; it < BN * (TBK / 4); it += 256) {
    const int c = it % BN, q = it / BN;  // consecutive threads read consecutive samples of one row of XT
    const int64_t col = (int64_t)ct * BN + c;
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = kb * TBK + q * 4 + i;
      t[i] = (k < n && col < ldb) ? __ldg(XT + (int64_t)k * ldb + col) : 0.f;
    }
    uint8_t* sub = blk + (size_t)(c / W) * 2 * kHalf;
    store_split(sub, sub + kHalf, (uint32_t)q * (W * 16) + (uint32_t)(c % W) * 16, make_float4(t[0], t[1], t[2], t[3]));
  }
}

template <int BN, int S, int CL>
__global__ void __launch_bounds__(kTcBlock, (BN <= 64 || (BN == 128 && S == 3)) ? 2 : 1) dense_apply_tc2_kernel(const float* __restrict__ Dsplit, const float* __restrict__ Xsplit, int32_t n,
                                                                     float* __restrict__ CT, int64_t ldb, int32_t B, float scale,
                                                                     const float* __restrict__ scale_dev, const float* __restrict__ sub,
                                                                     float* __restrict__ partials, int32_t flush, int32_t debug, int32_t gap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_full[S];
  __shared__ __align__(8) uint64_t s_empty[S];  // per stage: the MMAs that read it have completed
  __shared__ __align__(8) uint64_t s_chunk[2];
  __shared__ __align__(8) uint64_t s_drained[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[kTcThreads / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = __shfl_sync(0xffffffffu, warp, 0) == kTcThreads / 32;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  constexpr uint32_t kTmemCols = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);  // two accumulators of BN columns (allocations are powers of two)
  constexpr uint32_t kBBytes = (uint32_t)BN * TBK * 4;
  constexpr uint32_t kAStage = 2 * kABytes, kBStage = b_stage_bytes(BN), kStage = kAStage + kBStage;
  // copies run kAhead = S - gap k-blocks ahead; a stage is refilled once the commit of the k-block `gap` back has arrived: the
  // arrival of a tcgen05.commit takes long enough that a gap of 2 (round 2's first version) paced the whole pipeline
  const int kAhead = S - gap;
  constexpr uint32_t kLboB = BN * 16;
  constexpr uint32_t kIdesc = instr_desc(BN);
  constexpr int WC = BN / 2;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_chunk[s]), 1);
      mbar_init(smem_u32(&s_drained[s]), kTcThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int nkb = (n + TBK - 1) / TBK;
  const int n_chunks = (nkb + flush - 1) / flush;
  const float* a_src = Dsplit + (size_t)blockIdx.y * nkb * (kAStage / 4);
  const float* x_src = Xsplit + (size_t)blockIdx.x * nkb * (kBStage / 4);

  const uint32_t t_own = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * WC);
  float acc[WC];
#pragma unroll
  for (int i = 0; i < WC; ++i) acc[i] = 0.f;

  if (is_issuer) {
    const uint32_t half = CL == 1 ? 0u : cluster_ctarank() * kABytes;
    auto copy_stage = [&](int stage, int kb) {
      const uint32_t bar = smem_u32(&s_full[0]) + (uint32_t)stage * 8;
      mbar_expect_tx(bar, kStage);
      const uint32_t dst = smem_u32(smem) + (uint32_t)stage * kStage;
      const float* srca = a_src + (size_t)kb * (kAStage / 4) + half / 4;
      if (CL == 1) bulk_copy(dst, srca, kAStage, bar);
      else bulk_copy_multicast(dst + half, srca, kABytes, bar, (uint16_t)3);
      bulk_copy(dst + kAStage, x_src + (size_t)kb * (kBStage / 4), kBStage, bar);
    };
    if (elect_one())
      for (int kb = 0; kb < kAhead && kb < nkb; ++kb) copy_stage(kb, kb);
    int c_stage = kAhead % S;
    constexpr uint32_t kDescHi = (kSbo >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((kLboA >> 4) << 16);
    const uint32_t b_lo0 = (((smem_u32(smem) + kAStage) & 0x3ffffu) >> 4) | ((kLboB >> 4) << 16);
    auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; };
    int stage = 0, phase = 0, chunk_pos = 0, chunk = 0;
    int e_stage = 0, e_phase = 0;  // the stage this iteration refills: last read by k-block kb - gap
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb >= gap) {
        mbar_wait(smem_u32(&s_empty[0]) + (uint32_t)e_stage * 8, e_phase);
        if (++e_stage == S) e_stage = 0, e_phase ^= 1;
      }
      if (kb + kAhead < nkb) {
        if (elect_one()) copy_stage(c_stage, kb + kAhead);
        if (++c_stage == S) c_stage = 0;
      }
      mbar_wait(smem_u32(&s_full[0]) + (uint32_t)stage * 8, phase);
      const bool first = chunk_pos == 0;
      const int buf = chunk & 1;
      if (first && chunk >= 2) mbar_wait(smem_u32(&s_drained[0]) + (uint32_t)buf * 8, ((chunk >> 1) - 1) & 1);  // accumulator free again
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool last_of_chunk = chunk_pos == flush - 1 || kb == nkb - 1;
      if (elect_one()) {
        const uint32_t la = a_lo0 + (uint32_t)stage * (kStage >> 4), lb = b_lo0 + (uint32_t)stage * (kStage >> 4);
        const uint32_t td = tmem + (uint32_t)buf * BN;
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          const uint64_t a_hi = desc(la + ks * (2 * kLboA >> 4)), a_lo = desc(la + (kABytes >> 4) + ks * (2 * kLboA >> 4));
          const uint64_t b_hi = desc(lb + ks * (2 * kLboB >> 4)), b_lo = desc(lb + (kBBytes >> 4) + ks * (2 * kLboB >> 4));
          if (debug < 2) umma_tf32(td, a_lo, b_hi, kIdesc, !(first && ks == 0));
          if (debug < 1) umma_tf32(td, a_hi, b_lo, kIdesc, 1);
          if (debug < 1) umma_tf32(td, a_hi, b_hi, kIdesc, 1);
        }
        const uint32_t bar_e = smem_u32(&s_empty[0]) + (uint32_t)stage * 8;
        if (CL == 1) umma_commit(bar_e);
        else umma_commit_multicast(bar_e, (uint16_t)3);
        if (last_of_chunk) umma_commit(smem_u32(&s_chunk[0]) + (uint32_t)buf * 8);
      }
      __syncwarp();
      if (++stage == S) stage = 0, phase ^= 1;
      if (last_of_chunk) chunk_pos = 0, ++chunk;
      else ++chunk_pos;
    }
  } else {
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(&s_chunk[0]) + (uint32_t)buf * 8, (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < WC / 16; ++j) {
        uint32_t v[16];
        tmem_ld16(t_own + (uint32_t)(buf * BN + j * 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[j * 16 + i] += __uint_as_float(v[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&s_drained[0]) + (uint32_t)buf * 8);
    }
  }

  // epilogue (as in the first kernel)
  const float sc = scale * (scale_dev != nullptr ? __ldg(scale_dev) : 1.0f);
  const int m = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * WC;
  float lsum = 0.f;
#pragma unroll
  for (int q = 0; q < WC / 4; ++q) {
    const int c = n0 + cbase + q * 4;
    if (!is_issuer && m < n && c < ldb) {
      float4 o = make_float4(sc * acc[q * 4 + 0], sc * acc[q * 4 + 1], sc * acc[q * 4 + 2], sc * acc[q * 4 + 3]);
      if (sub != nullptr) {
        const float4 sv = __ldg(reinterpret_cast<const float4*>(sub + (int64_t)m * ldb + c));
        o.x -= sv.x;
        o.y -= sv.y;
        o.z -= sv.z;
        o.w -= sv.w;
      }
      if (c + 0 < B) lsum = fmaf(o.x, o.x, lsum);
      if (c + 1 < B) lsum = fmaf(o.y, o.y, lsum);
      if (c + 2 < B) lsum = fmaf(o.z, o.z, lsum);
      if (c + 3 < B) lsum = fmaf(o.w, o.w, lsum);
      *reinterpret_cast<float4*>(CT + (int64_t)m * ldb + c) = o;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0 && !is_issuer) s_part[warp] = lsum;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && partials != nullptr) {
    float t = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) t += s_part[w];
    partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
  if (CL > 1) cluster_sync();
}

// ---------------------------------------------------------------------------------------------------------------------
// Third generation: CTA pairs (tcgen05 cta_group::2).
//
// The second kernel is paced by what each SM has to take in: its 128-row operator panel AND its whole activation tile, both as
// [hi | lo] -- 36 KB per k-block at 160 columns, ~59 B/clk for the whole run (a run without MMAs is barely faster).  Here two
// CTAs on the SMs of one TPC (cluster 2 x 1: two row tiles of the same column tile) run ONE M = 256 MMA per product: each
// holds its own 128 operator rows and only HALF of the activation tile (the tensor cores of the pair read both halves), so an
// SM takes in 16 + 10 instead of 16 + 20 KB per k-block and reads 39 instead of 54 KB of operands per k-block from its shared
// memory.  The leader CTA (rank 0) issues the MMAs and commits to the mbarriers of both CTAs; both fill their own stages with
// tensor-map loads (cp.async.bulk.tensor .cta_group::2) whose transactions are counted on the LEADER's full barrier -- a first
// version relayed "my stage has landed" by a remote mbarrier arrive per k-block and lost 60 % to the cluster-scope handshakes.
//   barriers: full[S]      (leader) both stages landed (transaction bytes of both CTAs)
//             empty[S]     MMAs of the k-block that read a stage done, in both CTAs (multicast commit) -> stage free
//             chunk_done[2] accumulator complete, in both CTAs            drained[2]  (leader) read back by all 512 epilogue threads
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t instr_desc_pair(int BN) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }

constexpr uint32_t kRowBytes = 1024;  // the pre-split arrays as rows of 256 floats for the tensor-map loads of the pair kernel
struct PairMaps {
  CUtensorMap a;  // the pre-split operator: a box of 16 rows = one 16 KB [hi | lo] stage of a row tile
  CUtensorMap x;  // the pre-split activations: a box of BN / 16 rows = one CTA's [hi | lo] half of a stage
  CUtensorMap ah; // the operator again with a box of 8 rows: the hi or the lo half of a stage (clusters of two pairs)
};

template <int BN, int S, int CM, int G>  // CM = 2: clusters of two pairs (two column tiles) that share every operator stage by multicast;
                                         // G: k-blocks per stage (one full / empty hand-over, one commit per G k-blocks)
__global__ void __launch_bounds__(kTcBlock, 1) dense_apply_tc3_kernel(const __grid_constant__ PairMaps maps, int32_t n,
                                                                     int32_t row_tiles, float* __restrict__ CT, int64_t ldb, int32_t B, float scale,
                                                                     const float* __restrict__ scale_dev, const float* __restrict__ sub,
                                                                     float* __restrict__ partials, int32_t flush, int32_t debug, int32_t gap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_full[S];
  __shared__ __align__(8) uint64_t s_empty[S];
  // accumulators in tensor memory: three where 3 BN <= 512 columns -- a drain (commit multicast, tcgen05.ld, local barrier, remote
  // arrival at the leader) takes about as long as the MMAs of one chunk, so with two the leader waited for it at every chunk
  constexpr int NACC = 3 * BN <= 512 ? 3 : 2;
  __shared__ __align__(8) uint64_t s_chunk[NACC];
  __shared__ __align__(8) uint64_t s_drained[NACC];        // (leader) both CTAs have read the accumulator back: one arrival per CTA
  __shared__ __align__(8) uint64_t s_drained_local[NACC];  // this CTA's 256 epilogue threads have
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[kTcThreads / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = __shfl_sync(0xffffffffu, warp, 0) == kTcThreads / 32;
  const uint32_t crank = cluster_ctarank();  // cluster (2, CM, 1), x fastest: pairs are the ranks (2 cy, 2 cy + 1)
  const uint32_t rank = crank & 1u, cy = crank >> 1;  // rank in the pair: the two row tiles 2p, 2p + 1 of one column tile
  const uint32_t pair_leader = crank & ~1u;
  const int row_tile = blockIdx.x, col_tile = blockIdx.y;  // pairs are formed along x (a 2-CTA kernel needs an even cluster x)
  const bool leader = rank == 0;
  const int m0 = row_tile * TBM, n0 = col_tile * BN;
  constexpr int HB = BN / 2;                                                   // activation columns held by one CTA
  constexpr uint32_t kTmemCols = NACC * BN <= 128 ? 128 : (NACC * BN <= 256 ? 256 : 512);  // NACC accumulators of BN columns
  constexpr uint32_t kXHalf = 2u * HB * TBK * 4;                              // this CTA's [hi | lo] activation sub-block
  constexpr uint32_t kXBytes = (uint32_t)HB * TBK * 4;                        // hi or lo of it
  constexpr uint32_t kAStage = 2 * kABytes, kStage = G * (kAStage + kXHalf);  // a stage: G operator blocks, then G activation half-blocks
  static_assert(G == 1 || CM == 1, "multi-k-block stages are built for single pairs");
  const int kAhead = S - gap;  // copies run S - gap k-blocks ahead, a stage is refilled after the commit of k-block kb - gap
  constexpr uint32_t kLboB = HB * 16;
  constexpr uint32_t kIdesc = instr_desc_pair(BN);
  constexpr int WC = BN / 2;

  if (warp == 0) {  // the same warp of both CTAs: a pair allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), CM);  // one commit per pair of the cluster
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(smem_u32(&s_chunk[s]), 1);
      mbar_init(smem_u32(&s_drained[s]), 2);
      mbar_init(smem_u32(&s_drained_local[s]), kTcThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();  // both CTAs' mbarriers and the pair's tensor memory exist
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int nkb = (n + TBK - 1) / TBK;
  const int nsb = (nkb + G - 1) / G;                 // stages' worth of k-blocks ("super-blocks"); the loops below count these
  const int flush_sb = max(1, flush / G);
  const int n_chunks = (nsb + flush_sb - 1) / flush_sb;
  const int a_tile = min(row_tile, row_tiles - 1);  // an odd count of row tiles: the padding CTA re-reads the last one, stores nothing
  // row coordinates (rows of 256 bytes) of this CTA's first operator stage and first activation half-stage
  const int32_t a_row0 = a_tile * nkb * (int32_t)(kAStage / kRowBytes);
  const int32_t x_row0 = col_tile * nkb * (int32_t)(b_stage_bytes(BN) / kRowBytes) + (int32_t)rank * (int32_t)(kXHalf / kRowBytes);

  const uint32_t t_own = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * WC);
  float acc[WC];
#pragma unroll
  for (int i = 0; i < WC; ++i) acc[i] = 0.f;

  if (is_issuer) {
    // both CTAs' loads count on the full barrier of the pair's LEADER, which expects the bytes of both stages.  CM = 2: the two
    // column tiles of the cluster read the same operator rows -- the cy = 0 CTA fetches the hi half of its row tile's stage, the
    // cy = 1 CTA the lo half, each multicast to both (half the operator reads from L2)
    const uint16_t col_mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));  // this row tile's CTAs in both column tiles
    auto copy_stage = [&](int stage, int sb) {
      const uint32_t bar = smem_u32(&s_full[0]) + (uint32_t)stage * 8;
      if (leader) mbar_expect_tx(bar, 2 * kStage);
      const uint32_t dst = smem_u32(smem) + (uint32_t)stage * kStage;
      const int32_t a_row = a_row0 + sb * G * (int32_t)(kAStage / kRowBytes);
      if (CM == 1) {  // G consecutive operator blocks of the row tile: one box (rows past the array are zero-filled and counted)
        tma_box_pair(dst, &maps.a, 0, a_row, bar & kPeerBitMask);
      } else {
        const uint32_t half = cy * kABytes;
        tma_box_pair_multicast(dst + half, &maps.ah, 0, a_row + (int32_t)(half / kRowBytes), bar, col_mask);
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
        tma_box_pair(dst + G * kAStage + g * kXHalf, &maps.x, 0, x_row0 + (sb * G + g) * (int32_t)(b_stage_bytes(BN) / kRowBytes), bar & kPeerBitMask);
    };
    if (elect_one())
      for (int sb = 0; sb < kAhead && sb < nsb; ++sb) copy_stage(sb, sb);
    int c_stage = kAhead % S;
    constexpr uint32_t kDescHi = (kSbo >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((kLboA >> 4) << 16);
    const uint32_t b_lo0 = (((smem_u32(smem) + G * kAStage) & 0x3ffffu) >> 4) | ((kLboB >> 4) << 16);
    auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; };
    int stage = 0, phase = 0, chunk_pos = 0, chunk = 0;
    int e_stage = 0, e_phase = 0;            // the stage the copy of this iteration refills, last read by k-block kb - gap
    for (int kb = 0; kb < nsb; ++kb) {  // kb counts super-blocks of G k-blocks here
      // One "empty" mbarrier PER STAGE (the commit of k-block kb arrives on empty[kb % S] of both CTAs): the leader may run
      // several k-blocks ahead of the peer's loop, and a barrier shared by alternating k-blocks could then complete two
      // phases between two looks of the peer.
      if (kb >= gap) {
        mbar_wait(smem_u32(&s_empty[0]) + (uint32_t)e_stage * 8, e_phase);  // MMAs of kb - gap done, pair-wide
        if (++e_stage == S) e_stage = 0, e_phase ^= 1;
      }
      if (kb + kAhead < nsb) {
        if (elect_one()) copy_stage(c_stage, kb + kAhead);
        if (++c_stage == S) c_stage = 0;
      }
      const bool first = chunk_pos == 0;
      const int buf = chunk % NACC;
      const bool last_of_chunk = chunk_pos == flush_sb - 1 || kb == nsb - 1;
      if (leader) {
        mbar_wait(smem_u32(&s_full[0]) + (uint32_t)stage * 8, phase);  // both CTAs' stages have landed
        if (first && chunk >= NACC) mbar_wait_cluster(smem_u32(&s_drained[0]) + (uint32_t)buf * 8, ((chunk / NACC) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t la = a_lo0 + (uint32_t)stage * (kStage >> 4), lb = b_lo0 + (uint32_t)stage * (kStage >> 4);
          const uint32_t td = tmem + (uint32_t)buf * BN;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (kb * G + g < nkb) {  // the last stage may hold fewer k-blocks
#pragma unroll
              for (int ks = 0; ks < TBK / 8; ++ks) {
                const uint32_t lag = la + g * (kAStage >> 4) + ks * (2 * kLboA >> 4), lbg = lb + g * (kXHalf >> 4) + ks * (2 * kLboB >> 4);
                const uint64_t a_hi = desc(lag), a_lo = desc(lag + (kABytes >> 4));
                const uint64_t b_hi = desc(lbg), b_lo = desc(lbg + (kXBytes >> 4));
                if (debug < 2) umma_tf32_pair(td, a_lo, b_hi, kIdesc, !(first && g == 0 && ks == 0));
                if (debug < 1) umma_tf32_pair(td, a_hi, b_lo, kIdesc, 1);
                if (debug < 1) umma_tf32_pair(td, a_hi, b_hi, kIdesc, 1);
              }
            }
          }
          umma_commit_pair_mask(smem_u32(&s_empty[0]) + (uint32_t)stage * 8, (uint16_t)((1u << (2 * CM)) - 1));  // every CTA of the cluster
          if (last_of_chunk) umma_commit_pair_mask(smem_u32(&s_chunk[0]) + (uint32_t)buf * 8, (uint16_t)(3u << pair_leader));  // this pair
        }
      }
      __syncwarp();
      if (++stage == S) stage = 0, phase ^= 1;
      if (last_of_chunk) chunk_pos = 0, ++chunk;
      else ++chunk_pos;
    }
  } else {
    const uint32_t drained_of_leader = mapa(smem_u32(&s_drained[0]), pair_leader);
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c % NACC;
      mbar_wait(smem_u32(&s_chunk[0]) + (uint32_t)buf * 8, (c / NACC) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < WC / 16; ++j) {
        uint32_t v[16];
        tmem_ld16(t_own + (uint32_t)(buf * BN + j * 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[j * 16 + i] += __uint_as_float(v[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      // 256 local arrivals, then ONE remote arrival per CTA on the leader's barrier (512 cluster-scope releases per chunk cost more)
      mbar_arrive(smem_u32(&s_drained_local[0]) + (uint32_t)buf * 8);
      if (tid == 0) {
        mbar_wait(smem_u32(&s_drained_local[0]) + (uint32_t)buf * 8, (c / NACC) & 1);
        mbar_arrive_cluster(drained_of_leader + (uint32_t)buf * 8);
      }
    }
  }

  // epilogue (as in the first kernel)
  const float sc = scale * (scale_dev != nullptr ? __ldg(scale_dev) : 1.0f);
  const int m = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * WC;
  float lsum = 0.f;
#pragma unroll
  for (int q = 0; q < WC / 4; ++q) {
    const int c = n0 + cbase + q * 4;
    if (!is_issuer && m < n && c < ldb) {
      float4 o = make_float4(sc * acc[q * 4 + 0], sc * acc[q * 4 + 1], sc * acc[q * 4 + 2], sc * acc[q * 4 + 3]);
      if (sub != nullptr) {
        const float4 sv = __ldg(reinterpret_cast<const float4*>(sub + (int64_t)m * ldb + c));
        o.x -= sv.x;
        o.y -= sv.y;
        o.z -= sv.z;
        o.w -= sv.w;
      }
      if (c + 0 < B) lsum = fmaf(o.x, o.x, lsum);
      if (c + 1 < B) lsum = fmaf(o.y, o.y, lsum);
      if (c + 2 < B) lsum = fmaf(o.z, o.z, lsum);
      if (c + 3 < B) lsum = fmaf(o.w, o.w, lsum);
      *reinterpret_cast<float4*>(CT + (int64_t)m * ldb + c) = o;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0 && !is_issuer) s_part[warp] = lsum;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && partials != nullptr) {
    float t = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) t += s_part[w];
    partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
  cluster_sync();  // neither CTA frees the pair's tensor memory or leaves while the other may still use or signal it
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  const int v = e != nullptr ? atoi(e) : dflt;
  return v > 0 ? v : dflt;
}
template <int BN, int SA, int CL>
int launch_tc(dim3 grid, const float* Dsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
              const float* scale_dev, const float* sub, float* partials, int flush, cudaStream_t st) {
  static bool configured = false;
  static const int debug = env_int("FEO_DENSE_DEBUG", 0);  // developer timing: 1 = one product of three, 2 = no MMAs (results are garbage)
  const int smem_bytes = (int)smem_bytes_for(BN, SA);
  if (!configured) {
    FEO_CUDA_CHECK(cudaFuncSetAttribute(dense_apply_tc_kernel<BN, SA, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcBlock);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  FEO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dense_apply_tc_kernel<BN, SA, CL>, Dsplit, n, XT, CT, ldb, B, scale, scale_dev, sub, partials,
                                    (int32_t)flush, (int32_t)debug));
  return FEO_OK;
}
template <int BN, int S, int CL>
int launch_tc2(dim3 grid, const float* Dsplit, float* Xsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
               const float* scale_dev, const float* sub, float* partials, int flush, cudaStream_t st) {
  static bool configured = false;
  static const int debug = env_int("FEO_DENSE_DEBUG", 0);
  const int smem_bytes = S * (int)(2 * kABytes + b_stage_bytes(BN));
  if (!configured) {
    FEO_CUDA_CHECK(cudaFuncSetAttribute(dense_apply_tc2_kernel<BN, S, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int nkb = (n + TBK - 1) / TBK;
  dense_split_x_kernel<BN, 1><<<dim3(grid.x, (unsigned)nkb), 256, 0, st>>>(XT, n, ldb, Xsplit, nkb);
  FEO_CUDA_CHECK(cudaGetLastError());
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcBlock);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  static const int gap_env = env_int("FEO_DENSE_GAP", 0);
  const int gap = gap_env >= 2 && gap_env < S ? gap_env : std::max(2, S / 2);
  FEO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dense_apply_tc2_kernel<BN, S, CL>, Dsplit, (const float*)Xsplit, n, CT, ldb, B, scale, scale_dev, sub,
                                    partials, (int32_t)flush, (int32_t)debug, (int32_t)gap));
  return FEO_OK;
}
template <int BN, int S, int CM, int G>
int launch_tc3(dim3 grid, const float* Dsplit, float* Xsplit, int32_t n, int32_t row_tiles, const float* XT, float* CT, int64_t ldb, int32_t B,
               float scale, const float* scale_dev, const float* sub, float* partials, int flush, cudaStream_t st) {
  static bool configured = false;
  static const int debug = env_int("FEO_DENSE_DEBUG", 0);
  const int smem_bytes = S * G * (int)(2 * kABytes + b_stage_bytes(BN) / 2);
  if (!configured) {
    FEO_CUDA_CHECK(cudaFuncSetAttribute(dense_apply_tc3_kernel<BN, S, CM, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int nkb = (n + TBK - 1) / TBK;
  dense_split_x_kernel<BN, 2><<<dim3(grid.y, (unsigned)nkb), 256, 0, st>>>(XT, n, ldb, Xsplit, nkb);
  FEO_CUDA_CHECK(cudaGetLastError());
  // the pre-split arrays as rows of 256 floats (1 KB): stages are boxes of whole rows
  PairMaps maps;
  const int64_t a_rows = (int64_t)row_tiles * nkb * (2 * kABytes / kRowBytes), x_rows = (int64_t)grid.y * nkb * (b_stage_bytes(BN) / kRowBytes);
  if (a_rows > INT32_MAX || x_rows > INT32_MAX) return fail(FEO_ERR_UNSUPPORTED, "dense_apply: operand too large for the pair kernel");
  if (int rc = make_row_map(Dsplit, kRowBytes / 4, a_rows, (int32_t)(G * 2 * kABytes / kRowBytes), &maps.a)) return rc;
  if (int rc = make_row_map(Dsplit, kRowBytes / 4, a_rows, (int32_t)(kABytes / kRowBytes), &maps.ah)) return rc;
  if (int rc = make_row_map(Xsplit, kRowBytes / 4, x_rows, (int32_t)(b_stage_bytes(BN) / 2 / kRowBytes), &maps.x)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcBlock);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;   // the two row tiles of a pair
  attr[0].val.clusterDim.y = CM;  // column tiles that share the operator stages
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int gap_env = env_int("FEO_DENSE_GAP", 0);
  const int gap = gap_env >= 1 && gap_env < S ? gap_env : std::max(1, S / 2);
  FEO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dense_apply_tc3_kernel<BN, S, CM, G>, maps, n, row_tiles, CT, ldb, B, scale, scale_dev, sub, partials, (int32_t)flush,
                                    (int32_t)debug, (int32_t)gap));
  return FEO_OK;
}
// cvt.rna.tf32.f32 on the host: round to nearest, ties away from zero, to 10 mantissa bits
float rna_tf32(float x) {
  uint32_t u = f2u(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;  // inf / nan
  u = (u + 0x1000u) & 0xffffe000u;
  float r;
  __builtin_memcpy(&r, &u, 4);
  return r;
}
}  // namespace

std::vector<float> dense_split_tiles(const float* src, int32_t n, bool transposed) {
  const int64_t row_tiles = (n + TBM - 1) / TBM, nkb = (n + TBK - 1) / TBK;
  const size_t block = 2 * kABytes / 4;  // floats per (row tile, k-block): [hi | lo]
  std::vector<float> out((size_t)row_tiles * nkb * block, 0.f);
  for (int32_t r = 0; r < n; ++r)
    for (int32_t k = 0; k < n; ++k) {
      const float x = transposed ? src[(size_t)k * n + r] : src[(size_t)r * n + k];
      const float hi = rna_tf32(x), lo = rna_tf32(x - hi);
      const int32_t rr = r % TBM, kk = k % TBK;
      const size_t at = ((size_t)(r / TBM) * nkb + k / TBK) * block + (size_t)(kk / 4) * (kLboA / 4) + (size_t)rr * 4 + kk % 4;
      out[at] = hi;
      out[at + kABytes / 4] = lo;
    }
  return out;
}

// bytes of the pre-split activations of the second-generation kernel: one [hi | lo] block per (column tile, k-block), column
// tiles padded to the cluster size; the largest over the tile widths the launcher may pick (a multiple of 1 KB)
size_t dense_xsplit_bytes(int32_t n, int64_t cols) {
  const int64_t c4 = (cols + 3) / 4 * 4, nkb = (n + TBK - 1) / TBK;
  size_t need = 0;
  for (int bn : {64, 128, 160, 192}) need = std::max(need, (size_t)((c4 + bn - 1) / bn + 1) * (size_t)nkb * b_stage_bytes(bn));
  return need;
}

int sm_count(int* out);  // feo_tiled.cu

int launch_dense_tc(const float* Dsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B,
                    float scale, const float* scale_dev, const float* sub, float* partials, int* count_out,
                    float* xsplit, size_t xsplit_bytes, cudaStream_t st) {
  static const int flush = env_int("FEO_DENSE_FLUSH", kFlushDefault);
  static const int bn_env = env_int("FEO_DENSE_BN", 0);
  if (Dsplit == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "dense operator not present in this handle");
  const int64_t cols = (B + 3) / 4 * 4;
  const int64_t row_tiles = (n + TBM - 1) / TBM;
  // FEO_DENSE_CLUSTER=2: pairs of column tiles share every operator stage through cluster multicast (each CTA fetches one
  // half).  Measured equal to the plain launch (0.164 vs 0.162 ms at N = 2549, B = 1024; 0.90 vs 0.89 ms at B = 8192):
  // the crossbar traffic it halves is not what paces the kernel, so the plain launch stays the default.
  static const int cl_env = env_int("FEO_DENSE_CLUSTER", 0);
  const int cl = cl_env == 2 ? 2 : 1;
  static const int sa_env = env_int("FEO_DENSE_ASTAGES", 0);
  // second generation (pre-split activations, ping-pong accumulators): the default whenever the caller's workspace holds
  // the split (feo_workspace_bytes accounts for it); FEO_DENSE_GEN=1 keeps the first kernel
  static const int gen_env = env_int("FEO_DENSE_GEN", 2);
  if (gen_env != 1 && xsplit != nullptr && xsplit_bytes >= dense_xsplit_bytes(n, cols)) {
    // Tile width and kernel generation: an M = 128 MMA takes the same time at N = 64 as at N = 128 (tools/micro9.cu) and
    // proportionally longer above, and the tiles of one SM share its tensor pipe, so the run time goes as
    //   waves * max(BN, 128),   waves = ceil(CTAs / SMs);
    // a wave of CTA pairs (third generation; row tiles rounded up to an even count) measured ~0.8 of a wave of single CTAs.
    // N = 2549, B = 1024: pairs at 160 columns (140 CTAs, one wave) 0.077 ms, singles 0.103; 128 columns two waves, 64 columns
    // three half-rate waves.  N = 2680, B = 1000: 21 row tiles -> 154 paired CTAs = two waves, singles (147) stay at one.
    // Ties go to the narrower tile (more SMs share the operand traffic).  FEO_DENSE_BN / FEO_DENSE_GEN=2|3 force either.
    int sms = 148;
    if (int rc = sm_count(&sms)) return rc;
    const int64_t rt2 = (row_tiles + 1) / 2 * 2;
    static const bool gen_forced = std::getenv("FEO_DENSE_GEN") != nullptr;
    int bn = 64;
    bool pairs_fit = false;
    int64_t best = -1;
    for (int cand : {64, 128, 160}) {
      if (bn_env != 0 && cand != bn_env) continue;
      const int64_t ct = (cols + cand - 1) / cand;
      const int64_t c2 = (row_tiles * ct + sms - 1) / sms * std::max(cand, 128) * 10;
      const int64_t c3 = (rt2 * ct + sms - 1) / sms * std::max(cand, 128) * 8;
      if ((!gen_forced || gen_env == 2) && (best < 0 || c2 < best)) best = c2, bn = cand, pairs_fit = false;
      if (cand >= 128 && cl == 1 && (!gen_forced || gen_env == 3) && (best < 0 || c3 < best)) best = c3, bn = cand, pairs_fit = true;
    }
    if (bn_env == 192) bn = 192, pairs_fit = cl == 1;  // 192-column tiles exist for the pair kernel only
    const unsigned col_tiles = (unsigned)((cols + bn - 1) / bn);
    dim3 grid((col_tiles + cl - 1) / cl * cl, (unsigned)row_tiles);
    *count_out = (int)(grid.x * grid.y);
    static const int cm_env = env_int("FEO_DENSE_CM", 0);
    if (bn >= 128 && cl == 1 && pairs_fit) {
      // CTA pairs: the row tiles are paired (an odd count gets a padding CTA); FEO_DENSE_CM=2: two column tiles per cluster
      // share the operator stages by multicast (an odd count of column tiles gets a padding column)
      const int cm = cm_env == 2 ? 2 : 1;
      const unsigned ct3 = (col_tiles + cm - 1) / cm * cm;
      dim3 grid3((unsigned)rt2, ct3);  // x: row tiles (paired), y: column tiles
      *count_out = (int)(grid3.x * grid3.y);
      // FEO_DENSE_KG: k-blocks per stage (1 / 2 / 4); the stage count shrinks accordingly (S * G stages' worth of k-blocks <= 8)
      static const int kg_env = env_int("FEO_DENSE_KG", 0);
      const int kg = (kg_env == 1 || kg_env == 2 || kg_env == 4) && cm == 1 ? kg_env : (cm == 1 ? 2 : 1);
      const int s_max = (bn == 192 ? 7 : 8) / kg;
      const int s3 = sa_env >= 2 && sa_env <= s_max ? sa_env : s_max;
#define FEO_TC3_CASE(BN_, S_, CM_, G_)                    \
  if (bn == BN_ && s3 == S_ && cm == CM_ && kg == G_)     \
  return launch_tc3<BN_, S_, CM_, G_>(grid3, Dsplit, xsplit, n, (int32_t)row_tiles, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush, st)
      FEO_TC3_CASE(128, 4, 1, 1);
      FEO_TC3_CASE(128, 8, 1, 1);
      FEO_TC3_CASE(160, 4, 1, 1);
      FEO_TC3_CASE(160, 6, 1, 1);
      FEO_TC3_CASE(160, 8, 1, 1);
      FEO_TC3_CASE(192, 7, 1, 1);
      FEO_TC3_CASE(128, 4, 1, 2);
      FEO_TC3_CASE(160, 4, 1, 2);
      FEO_TC3_CASE(160, 3, 1, 2);
      FEO_TC3_CASE(192, 3, 1, 2);
      FEO_TC3_CASE(128, 2, 1, 4);
      FEO_TC3_CASE(160, 2, 1, 4);
      FEO_TC3_CASE(128, 8, 2, 1);
      FEO_TC3_CASE(160, 8, 2, 1);
      FEO_TC3_CASE(192, 7, 2, 1);
      FEO_TC3_CASE(192, 5, 2, 1);
#undef FEO_TC3_CASE
      return fail(FEO_ERR_INVALID_ARGUMENT, "dense_apply: no third-generation kernel for this tile configuration");
    }
    // stages: 24 / 32 / 36 KB each; two CTAs per SM up to 128 columns (four stages at 64, three at 128), one at 160 (six)
    const int s_dflt = bn == 64 ? 4 : (bn == 128 ? 3 : 6);
    const int s2 = sa_env >= 3 && sa_env <= 6 ? sa_env : s_dflt;
#define FEO_TC2_CASE(BN_, S_, CL_) \
  if (bn == BN_ && s2 == S_ && cl == CL_) \
  return launch_tc2<BN_, S_, CL_>(grid, Dsplit, xsplit, n, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush, st)
    FEO_TC2_CASE(64, 3, 1);
    FEO_TC2_CASE(64, 4, 1);
    FEO_TC2_CASE(64, 5, 1);
    FEO_TC2_CASE(64, 6, 1);
    FEO_TC2_CASE(64, 4, 2);
    FEO_TC2_CASE(128, 3, 1);
    FEO_TC2_CASE(128, 4, 1);
    FEO_TC2_CASE(128, 5, 1);
    FEO_TC2_CASE(128, 6, 1);
    FEO_TC2_CASE(128, 3, 2);
    FEO_TC2_CASE(160, 4, 1);
    FEO_TC2_CASE(160, 5, 1);
    FEO_TC2_CASE(160, 6, 1);
    FEO_TC2_CASE(160, 6, 2);
#undef FEO_TC2_CASE
    return fail(FEO_ERR_INVALID_ARGUMENT, "dense_apply: no second-generation kernel for this tile configuration");
  }
  // first generation: 64-column tiles put three CTAs on every SM and were the fastest at every measured size (N = 387 .. 2549,
  // B = 1024 .. 8192); 128-column tiles (half the operator re-reads, 128 registers) stay selectable for experiments
  int bn = 64;
  if (bn_env == 64 || bn_env == 128) bn = bn_env;
  // cluster launches need a grid that is a multiple of the cluster: a padding column tile runs the protocol and stores nothing
  const unsigned col_tiles = (unsigned)((cols + bn - 1) / bn);
  dim3 grid((col_tiles + cl - 1) / cl * cl, (unsigned)row_tiles);
  *count_out = (int)(grid.x * grid.y);
  const int sa = sa_env >= 3 && sa_env <= 5 ? sa_env : 3;
#define FEO_TC_CASE(BN_, SA_, CL_) \
  if (bn == BN_ && sa == SA_ && cl == CL_) \
  return launch_tc<BN_, SA_, CL_>(grid, Dsplit, n, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush, st)
  FEO_TC_CASE(64, 3, 2);
  FEO_TC_CASE(64, 4, 2);
  FEO_TC_CASE(64, 5, 2);
  FEO_TC_CASE(128, 3, 2);
  FEO_TC_CASE(64, 3, 1);
  FEO_TC_CASE(64, 4, 1);
  FEO_TC_CASE(128, 3, 1);
#undef FEO_TC_CASE
  return fail(FEO_ERR_INVALID_ARGUMENT, "dense_apply: no kernel for this tile configuration");
}

}  // namespace feo
