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
// One CTA = one 128 x 128 tile of CT (UMMA M = 128, N = 128, K = 8), k-blocks of 16, two shared-memory
// stages.  The operands come straight from global memory through registers (the split needs a register
// pass anyway, and XT is sample-contiguous, i.e. MN-major: the register pass also transposes it into the
// K-major core-matrix layout), so there is no TMA here; the stage hand-over is
//   generic stores -> fence.proxy.async -> bar.sync -> one thread issues 6 MMAs -> tcgen05.commit -> mbarrier
// and the loads of the next k-block are in flight while the tensor core works.  64 KB of shared memory and
// 128 TMEM columns per CTA: two CTAs per SM overlap each other's load latency.
//
// The tensor core aligns and TRUNCATES when it adds a K = 8 product group to the fp32 accumulator, which biases a
// long accumulation towards zero (measured: 3e-6 of |D||x| at n = 2549, against 1e-7 for the split itself).  The
// accumulator is therefore drained every `flush` k-blocks (default 4 = 64 k, 24 accumulations; costs nothing measurable) into fp32 registers with round-to-nearest adds
// -- 64 registers per thread, the epilogue's own TMEM mapping -- and restarted with accumulate = 0.
//
// Shared-memory operand layout (canonical K-major, SWIZZLE_NONE): element (r, k) of a [128 x 16] tile lives at
//   (k / 4) * 2048 + r * 16 + (k % 4) * 4   bytes,
// i.e. 8 x 16-byte core matrices, 128 B each, SBO (next 8 rows) = 128 B, LBO (next 4 k) = 2048 B.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "feo_internal.h"

namespace feo {
namespace {

constexpr int TBM = 128, TBK = 16;                      // CT tile rows, k-block; tile columns BN = 128 or 64 (template)
constexpr int kTcThreads = 256;
constexpr uint32_t kABytes = TBM * TBK * 4;              // 8 KB: A_hi or A_lo of a stage
__host__ __device__ constexpr uint32_t stage_bytes(int BN) { return 2 * kABytes + 2 * (uint32_t)BN * TBK * 4; }  // A_hi, A_lo, B_hi, B_lo
constexpr int kTcStages = 2;
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

template <int BN>
__global__ void __launch_bounds__(kTcThreads, BN == 128 ? 2 : 3) dense_apply_tc_kernel(const float* __restrict__ D, int32_t n, int32_t ldd,
                                                                    const float* __restrict__ XT, float* __restrict__ CT,
                                                                    int64_t ldb, int32_t B, float scale,
                                                                    const float* __restrict__ scale_dev,
                                                                    const float* __restrict__ sub,
                                                                    float* __restrict__ partials, int32_t flush) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar[kTcStages];
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[kTcThreads / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  constexpr uint32_t kTmemCols = BN;                      // fp32 accumulator columns (power of two >= 32)
  constexpr uint32_t kBBytes = (uint32_t)BN * TBK * 4;    // B_hi or B_lo of a stage
  constexpr uint32_t kStage = stage_bytes(BN);
  constexpr uint32_t kLboB = BN * 16;
  constexpr uint32_t kIdesc = instr_desc(BN);
  constexpr int CB = BN / 64;                             // k-chunks of the B tile per thread
  constexpr int WC = BN / 2;                              // accumulator columns per warp

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kTcStages; ++s) mbar_init(smem_u32(&s_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  // loader coordinates.  A tile: thread -> (row, pair of k-chunks) = 32 contiguous bytes of a row of D;
  // B tile: thread -> (sample column, CB k-chunks): four coalesced row reads of XT per chunk.
  const int a_r = tid & 127, a_c = (tid >> 7) * 2;      // k-chunks a_c, a_c + 1
  const int b_c = tid % BN, b_q = (tid / BN) * CB;      // k-chunks b_q .. b_q + CB - 1
  const bool a_row_ok = m0 + a_r < n;
  const bool b_col_ok = n0 + b_c < ldb;
  const float* a_src = D + (int64_t)(m0 + a_r) * ldd + a_c * 4;
  const float* b_src = XT + n0 + b_c;

  // two register sets: the loads of k-blocks kb + 1 and kb + 2 are in flight while kb is split and multiplied
  float4 av0[2], bv0[CB], av1[2], bv1[CB];
  auto load_block = [&](int k0, float4 (&av)[2], float4 (&bv)[CB]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + (a_c + j) * 4;
      av[j] = (a_row_ok && k < ldd) ? __ldg(reinterpret_cast<const float4*>(a_src + k0 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
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

  const int nkb = (n + TBK - 1) / TBK;
  auto k_block = [&](int kb, float4 (&av)[2], float4 (&bv)[CB]) {
    const int s = kb & 1;
    uint8_t* stage = smem + s * kStage;
    // the MMAs that read this stage two k-blocks ago have completed (use j waits for commit j - 1)
    if (kb >= kTcStages) mbar_wait(smem_u32(&s_bar[s]), ((kb >> 1) - 1) & 1);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      store_split(stage, stage + kABytes, (uint32_t)(a_c + j) * kLboA + (uint32_t)a_r * 16, av[j]);
#pragma unroll
    for (int j = 0; j < CB; ++j)
      store_split(stage + 2 * kABytes, stage + 2 * kABytes + kBBytes, (uint32_t)(b_q + j) * kLboB + (uint32_t)b_c * 16, bv[j]);
    if (kb + 2 < nkb) load_block((kb + 2) * TBK, av, bv);  // this register set is free again
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(stage);
#pragma unroll
      for (int ks = 0; ks < TBK / 8; ++ks) {
        const uint64_t a_hi = smem_desc(base + ks * 2 * kLboA, kLboA), a_lo = smem_desc(base + kABytes + ks * 2 * kLboA, kLboA);
        const uint64_t b_hi = smem_desc(base + 2 * kABytes + ks * 2 * kLboB, kLboB);
        const uint64_t b_lo = smem_desc(base + 2 * kABytes + kBBytes + ks * 2 * kLboB, kLboB);
        umma_tf32(tmem, a_lo, b_hi, kIdesc, ((kb % flush) | ks) != 0);  // first product of a chunk overwrites
        umma_tf32(tmem, a_hi, b_lo, kIdesc, 1);
        umma_tf32(tmem, a_hi, b_hi, kIdesc, 1);
      }
      umma_commit(smem_u32(&s_bar[s]));
    }
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
      // orders these TMEM reads before the next chunk's overwrite (next k-block: bar.sync, then the issuing thread's fence)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  };
  load_block(0, av0, bv0);
  if (nkb > 1) load_block(TBK, av1, bv1);
  for (int kb = 0; kb < nkb; kb += 2) {
    k_block(kb, av0, bv0);
    if (kb + 1 < nkb) k_block(kb + 1, av1, bv1);
  }

  // epilogue
  const float sc = scale * (scale_dev != nullptr ? __ldg(scale_dev) : 1.0f);
  const int m = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * WC;
  float lsum = 0.f;
#pragma unroll
  for (int q = 0; q < WC / 4; ++q) {
    const int c = n0 + cbase + q * 4;
    if (m < n && c < ldb) {
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
  if (lane == 0) s_part[warp] = lsum;
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
}

template <int BN>
int launch_tc(dim3 grid, const float* D, int32_t n, int32_t ldd, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
              const float* scale_dev, const float* sub, float* partials, int flush, cudaStream_t st) {
  static bool configured = false;
  const int smem_bytes = kTcStages * stage_bytes(BN);
  if (!configured) {
    FEO_CUDA_CHECK(cudaFuncSetAttribute(dense_apply_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  dense_apply_tc_kernel<BN><<<grid, kTcThreads, smem_bytes, st>>>(D, n, ldd, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}
int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  const int v = e != nullptr ? atoi(e) : dflt;
  return v > 0 ? v : dflt;
}
}  // namespace

int launch_dense_tc(const float* D, int32_t n, int32_t ldd, const float* XT, float* CT, int64_t ldb, int32_t B,
                    float scale, const float* scale_dev, const float* sub, float* partials, int* count_out,
                    cudaStream_t st) {
  static const int flush = env_int("FEO_DENSE_FLUSH", kFlushDefault);
  static const int bn_env = env_int("FEO_DENSE_BN", 0);
  const int64_t cols = (B + 3) / 4 * 4;
  const int64_t row_tiles = (n + TBM - 1) / TBM;
  // 128-column tiles halve the reads of D per flop; 64-column tiles put several CTAs on every SM, which is what hides
  // the global-load latency of a k-block while the problem is small
  int bn = row_tiles * ((cols + 127) / 128) >= 4 * 148 ? 128 : 64;
  if (bn_env == 64 || bn_env == 128) bn = bn_env;
  dim3 grid((unsigned)((cols + bn - 1) / bn), (unsigned)row_tiles);
  *count_out = (int)(grid.x * grid.y);
  if (bn == 128) return launch_tc<128>(grid, D, n, ldd, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush, st);
  return launch_tc<64>(grid, D, n, ldd, XT, CT, ldb, B, scale, scale_dev, sub, partials, flush, st);
}

}  // namespace feo
