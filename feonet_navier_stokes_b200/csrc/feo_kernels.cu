// sm_100a kernels around the fused residual path (feo_tiled.cu): generic sparse applies for the
// materialised (LHS, RHS) API, the time-dependent sequence kernels, the dense operator GEMM, layout
// transposes and the loss reductions.
//
// Data layout: every batch of coefficient vectors is dof-major, XT[d*ldb + b] (see feonet_b200.h).
// In the generic kernels a warp owns one operator row for 128 consecutive samples: lane l holds
// samples 4l..4l+3 as a float4, so every gather `XT[col*ldb + b]` is one fully coalesced 512-byte
// request.  Everything is row-owned: no atomics, fixed summation order, bit-reproducible.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "feo_internal.h"

namespace feo {
namespace {

constexpr int kWarpSamples = 128;  // samples per warp (float4 per lane)
constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming load: read once, do not keep in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void fma4(float4& acc, float a, const float4& x) {
  acc.x = fmaf(a, x.x, acc.x);
  acc.y = fmaf(a, x.y, acc.y);
  acc.z = fmaf(a, x.z, acc.z);
  acc.w = fmaf(a, x.w, acc.w);
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fixed-order block reduction; thread 0 writes the CTA partial.
__device__ __forceinline__ void block_partial(float v, float* partials, int slot) {
  __shared__ float s_part[32];
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) t += s_part[w];
    partials[slot] = t;
  }
}

__global__ void finalize_loss_kernel(const float* __restrict__ partials, int count, float scale,
                                     float* __restrict__ loss_out) {
  __shared__ double s[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += (double)partials[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = (float)(s[0] * (double)scale);
}

// ---------------------------------------------------------------------------------------------
// generic sparse apply: YT[r] = scale * sum_k val_k XT[col_k] (+ YT[r])
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) spmm_kernel(const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ col,
                                                        const float* __restrict__ val, int32_t n,
                                                        const float* __restrict__ XT, float* __restrict__ YT,
                                                        int64_t ldb, int32_t B, float scale, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int r = blockIdx.x * nwarps + warp;
  if (r >= n) return;
  const int b = blockIdx.y * kWarpSamples + lane * 4;
  const bool active = b < B;
  const int bb = active ? b : 0;
  float4 acc = zero4();
  const int kb = __ldg(rowptr + r), ke = __ldg(rowptr + r + 1);
#pragma unroll 4
  for (int k = kb; k < ke; ++k) {
    const int c = __ldg(col + k);
    const float v = __ldg(val + k);
    fma4(acc, v, ldg4(XT + (int64_t)c * ldb + bb));
  }
  if (!active) return;
  float* y = YT + (int64_t)r * ldb + b;
  float4 out = make_float4(scale * acc.x, scale * acc.y, scale * acc.z, scale * acc.w);
  if (accumulate) {
    const float4 o = *reinterpret_cast<const float4*>(y);
    out.x += o.x;
    out.y += o.y;
    out.z += o.z;
    out.w += o.w;
  }
  stg4(y, out);
}

// ---------------------------------------------------------------------------------------------
// time-dependent Stokes residual (FEONet_time_dep_Stokes/train_FEONet.py:343-362, :398-400), pseudo-samples j = b*T + t
// contiguous in memory (dof-major [n, ldj]):
//   forward   r[i, j] = sum_c M[i,c] x[c, j]  -  ( sum_c S[i,c] prev[c, j] + dt F[i, b] ),   prev[c, j] = t > 0 ? x[c, j-1] : u0[c, b]
//   backward  g[c, j] = 2/T gl ( sum_i M[i,c] r[i, j]  -  [t < T-1] sum_i S[i,c] r[i, j+1] )
// One warp = one row x 256 pseudo-samples (two tiles of 128), a lane owns FOUR CONSECUTIVE j per tile: per column of the
// union pattern (M, S share it; one 16-byte word {col, m, s}) ONE 128-bit gather per tile feeds both sums -- the neighbouring
// time level of three of the four elements is already in the lane's registers, the fourth comes from the neighbouring lane
// by shuffle (lane 0 / 31 load it), and the t = 0 elements (one in T) read u0.  The row's words are fetched by the lanes in
// one coalesced load and handed round by shuffle, and every load of two columns (gathers, lane-edge elements, initial
// conditions) is issued before the first use: the first version (two passes of scalar gathers, a dependent word load per
// entry) was bound by exposed L2 latency.
// The order of every row sum is that of the separate CSR rows (absent coefficients are exact zero terms).
// ---------------------------------------------------------------------------------------------
constexpr int kSeqTiles = 1;   // tiles of 128 pseudo-samples per warp
#ifndef FEO_SEQ_BATCH
#define FEO_SEQ_BATCH 2         // columns whose loads are in flight together
#endif
#ifndef FEO_SEQ_MINB
#define FEO_SEQ_MINB 4          // resident blocks the register allocation aims at
#endif
template <bool BACKWARD>
__global__ void __launch_bounds__(kThreads, FEO_SEQ_MINB) seq_kernel(const int32_t* __restrict__ rowptr, const SeqEnt* __restrict__ ent, int32_t n,
                                                       const float* __restrict__ XT, const float* __restrict__ u0T,
                                                       const float* __restrict__ fT, float dt, int64_t ldj,
                                                       int64_t ldb, int32_t B, int32_t T,
                                                       const float* __restrict__ grad_loss, float* __restrict__ outT,
                                                       float* __restrict__ partials) {
  constexpr int NT = kSeqTiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int r = blockIdx.x * nwarps + warp;
  const int J = B * T;  // the launcher checks that B * T + 256 fits 32 bits: one 32-bit division per lane and tile below
  float lsum = 0.f;
  if (r < n) {  // warp-uniform
    int j0[NT];
    bool valid[NT], ok[NT][4], edge[NT][4];
    int bs[NT][4], e1[NT], eb[NT];  // forward, T >= 4: the lane's only t = 0 element of the tile (if any) and its sample
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      j0[q] = (blockIdx.y * NT + q) * kWarpSamples + lane * 4;
      valid[q] = j0[q] < ldj;  // ldj % 4 == 0: the four elements exist in memory (columns >= J are padding)
      e1[q] = -1, eb[q] = 0;
      const int jc = j0[q] < J ? j0[q] : 0;
      int b = jc / T, t = jc - b * T;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ok[q][i] = j0[q] + i < J;
        bs[q][i] = ok[q][i] ? b : 0;
        // forward: the previous level of t = 0 is the initial condition; backward: the last level has no successor
        edge[q][i] = BACKWARD ? (!ok[q][i] || t == T - 1) : (ok[q][i] && t == 0);
        if (!BACKWARD && edge[q][i]) e1[q] = i, eb[q] = b;
        if (++t == T) t = 0, ++b;
      }
    }
    const bool one_edge = T >= 4;  // warp-uniform
    float fv[NT][4];              // forward: the load vector of the row, fetched before the sums
#pragma unroll
    for (int q = 0; q < NT; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) fv[q][i] = (!BACKWARD && ok[q][i]) ? __ldg(fT + (int64_t)r * ldb + bs[q][i]) : 0.f;
    float acc[NT][4], acc2[NT][4];
#pragma unroll
    for (int q = 0; q < NT; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][i] = 0.f, acc2[q][i] = 0.f;
    const int kb = __ldg(rowptr + r), ke = __ldg(rowptr + r + 1);
    const int4* words = reinterpret_cast<const int4*>(ent);

    struct Col {
      int c;
      float m, s;
      float4 x[NT];
      float nb[NT], u[NT];
    };
    auto fetch = [&](const int4& mine, int src, Col& k) {
      k.c = __shfl_sync(0xffffffffu, mine.x, src);
      k.m = __int_as_float(__shfl_sync(0xffffffffu, mine.y, src));
      k.s = __int_as_float(__shfl_sync(0xffffffffu, mine.z, src));
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const float* row = XT + (int64_t)k.c * ldj + j0[q];
        k.x[q] = valid[q] ? __ldg(reinterpret_cast<const float4*>(row)) : make_float4(0.f, 0.f, 0.f, 0.f);
        // the element beyond the lane's four that the shuffle cannot deliver
        if (!BACKWARD) k.nb[q] = (lane == 0 && valid[q] && j0[q] > 0) ? __ldg(row - 1) : 0.f;
        else k.nb[q] = (lane == 31 && j0[q] + 4 < J) ? __ldg(row + 4) : 0.f;
        k.u[q] = (!BACKWARD && one_edge && e1[q] >= 0 && k.s != 0.f) ? __ldg(u0T + (int64_t)k.c * ldb + eb[q]) : 0.f;
      }
    };
    auto apply = [&](const Col& k) {
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const float4 x = k.x[q];
        float y[4];  // the neighbouring time level
        if (!BACKWARD) {
          const float up = __shfl_up_sync(0xffffffffu, x.w, 1);
          y[0] = lane == 0 ? k.nb[q] : up, y[1] = x.x, y[2] = x.y, y[3] = x.z;
          if (one_edge) {
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = e1[q] == i ? k.u[q] : y[i];
          } else if (k.s != 0.f) {  // warp-uniform: columns outside S need no initial condition
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (edge[q][i]) y[i] = __ldg(u0T + (int64_t)k.c * ldb + bs[q][i]);
          }
        } else {
          const float dn = __shfl_down_sync(0xffffffffu, x.x, 1);
          y[0] = x.y, y[1] = x.z, y[2] = x.w, y[3] = lane == 31 ? k.nb[q] : dn;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (edge[q][i]) y[i] = 0.f;
        }
        acc[q][0] = fmaf(k.m, x.x, acc[q][0]), acc[q][1] = fmaf(k.m, x.y, acc[q][1]);
        acc[q][2] = fmaf(k.m, x.z, acc[q][2]), acc[q][3] = fmaf(k.m, x.w, acc[q][3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[q][i] = fmaf(k.s, y[i], acc2[q][i]);
      }
    };
    for (int k0 = kb; k0 < ke; k0 += 32) {  // 32 words per coalesced fetch
      const int cnt = min(32, ke - k0);
      const int4 mine = lane < cnt ? __ldg(words + k0 + lane) : make_int4(0, 0, 0, 0);
      int k = 0;
      for (; k + FEO_SEQ_BATCH <= cnt; k += FEO_SEQ_BATCH) {  // the loads of FEO_SEQ_BATCH columns in flight
        Col c[FEO_SEQ_BATCH];
#pragma unroll
        for (int u = 0; u < FEO_SEQ_BATCH; ++u) fetch(mine, k + u, c[u]);
#pragma unroll
        for (int u = 0; u < FEO_SEQ_BATCH; ++u) apply(c[u]);
      }
      for (; k < cnt; ++k) {
        Col c0;
        fetch(mine, k, c0);
        apply(c0);
      }
    }
    const float g = BACKWARD ? (2.0f / (float)T) * (grad_loss != nullptr ? __ldg(grad_loss) : 1.0f) : 0.f;
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!BACKWARD) {
          // RHS_t = prev S^T + dt F ; r = LHS - RHS (FEONet_time_dep_Stokes/train_FEONet.py:357, :398)
          const float rhs = fmaf(dt, fv[q][i], acc2[q][i]);
          o[i] = ok[q][i] ? acc[q][i] - rhs : 0.f;
          lsum = fmaf(o[i], o[i], lsum);
        } else {
          o[i] = ok[q][i] ? g * (acc[q][i] - acc2[q][i]) : 0.f;
        }
      }
      if (valid[q]) *reinterpret_cast<float4*>(outT + (int64_t)r * ldj + j0[q]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  if (!BACKWARD) block_partial(lsum, partials, blockIdx.y * gridDim.x + blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// sum (x-y)^2 over [n][0..B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) sq_diff_kernel(const float* __restrict__ xT, const float* __restrict__ yT,
                                                           int32_t n, int64_t ldb, int32_t B,
                                                           float* __restrict__ partials) {
  float lsum = 0.f;
  for (int r = blockIdx.x; r < n; r += gridDim.x)
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      float d = xT[(int64_t)r * ldb + b];
      if (yT != nullptr) d -= yT[(int64_t)r * ldb + b];
      lsum = fmaf(d, d, lsum);
    }
  block_partial(lsum, partials, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// tiled transpose; optionally the destination rows are scattered through a map (assemble_u_init, dof permutation on the way
// in) or the source rows gathered through one (dof permutation on the way out): either way whole 256-byte runs move
// ---------------------------------------------------------------------------------------------
template <bool SMAP, bool DMAP>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int64_t src_ld,
                                                        float* __restrict__ dst, int64_t dst_ld, int32_t rows,
                                                        int32_t cols, const int32_t* __restrict__ dst_row_map,
                                                        const int32_t* __restrict__ src_row_map) {
  __shared__ float tile[64][65];
  __shared__ int32_t s_map[64];  // the tile's 64 mapped rows, looked up once
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  if (SMAP) {
    if (threadIdx.x < 64) s_map[threadIdx.x] = r0 + threadIdx.x < rows ? __ldg(src_row_map + r0 + threadIdx.x) : 0;
    __syncthreads();
  }
  float v[16];  // all sixteen loads of a thread in flight before the first shared-memory store
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = ty + 4 * k, r = r0 + i, c = c0 + tx;
    const int64_t srow = SMAP ? (int64_t)s_map[i] : (int64_t)r;
    v[k] = (r < rows && c < cols) ? __ldg(src + srow * src_ld + c) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) tile[ty + 4 * k][tx] = v[k];
  if (SMAP && DMAP) __syncthreads();  // the gather above has finished with s_map
  if (DMAP && threadIdx.x < 64) s_map[threadIdx.x] = c0 + threadIdx.x < cols ? __ldg(dst_row_map + c0 + threadIdx.x) : 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = ty + 4 * k, c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) {
      const int64_t drow = DMAP ? (int64_t)s_map[i] : (int64_t)c;
      dst[drow * dst_ld + r] = tile[tx][i];
    }
  }
}

// The same transpose for TALL sources (dof-major -> row-major: many source rows, few columns): tiles of 128 source rows x 32
// columns, so that the destination -- rows of the caller's [B, N] tensor, which start at arbitrary 4-byte alignment -- is
// written in runs of 512 bytes instead of 256 (half as many partially written sectors at the run ends).
template <bool SMAP>
__global__ void __launch_bounds__(256) transpose_tall_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst,
                                                             int64_t dst_ld, int32_t rows, int32_t cols,
                                                             const int32_t* __restrict__ src_row_map) {
  __shared__ float tile[128][33];
  __shared__ int32_t s_map[128];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 128;
  if (SMAP) {
    if (threadIdx.x < 128) s_map[threadIdx.x] = r0 + threadIdx.x < rows ? __ldg(src_row_map + r0 + threadIdx.x) : 0;
    __syncthreads();
  }
  {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = ty + 8 * k, r = r0 + i, c = c0 + tx;
      const int64_t srow = SMAP ? (int64_t)s_map[i] : (int64_t)r;
      v[k] = (r < rows && c < cols) ? __ldg(src + srow * src_ld + c) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) tile[ty + 8 * k][tx] = v[k];
  }
  __syncthreads();
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;  // 128 x 2
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int j = ty + 2 * k, c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows) dst[(int64_t)c * dst_ld + r] = tile[tx][j];
  }
}

int check_layout(const void* p, int64_t ld, int32_t B, const char* what) {
  if (p == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " is NULL");
  if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + " not 16-byte aligned");
  if (ld % 4 != 0 || ld < ((B + 3) / 4) * 4)
    return fail(FEO_ERR_INVALID_ARGUMENT, std::string(what) + ": ldb must be a multiple of 4 and >= ceil4(B)");
  return FEO_OK;
}

}  // namespace

size_t loss_partials_needed(int32_t n, int32_t fused_warps, int64_t cols) {
  // the fused forward writes one partial per (persistent CTA, consumer warp), the generic kernels one per (8 rows, 128 samples)
  const int64_t fused = (int64_t)fused_partials_needed(fused_warps);
  const int64_t generic = (int64_t)((n + 7) / 8) * ((cols + kWarpSamples - 1) / kWarpSamples);
  return (size_t)(std::max<int64_t>(std::max(fused, generic), 1024)) * sizeof(float);
}

int finalize_loss(float* partials, int count, float scale, float* loss_out, cudaStream_t st) {
  finalize_loss_kernel<<<1, 1024, 0, st>>>(partials, count, scale, loss_out);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}
static int finalize(float* partials, int count, float scale, float* loss_out, cudaStream_t st) {
  return finalize_loss(partials, count, scale, loss_out, st);
}

int launch_transpose(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                     const int32_t* dst_row_map, const int32_t* src_row_map, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return FEO_OK;
  if (src == nullptr || dst == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "transpose: NULL pointer");
  if (dst_row_map == nullptr && rows >= 4 * (int64_t)cols && rows >= 1024) {  // tall source: long destination runs
    dim3 gt((cols + 31) / 32, (rows + 127) / 128);
    if (src_row_map != nullptr) transpose_tall_kernel<true><<<gt, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, src_row_map);
    else transpose_tall_kernel<false><<<gt, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, src_row_map);
    FEO_CUDA_CHECK(cudaGetLastError());
    return FEO_OK;
  }
  dim3 grid((cols + 63) / 64, (rows + 63) / 64);
  if (src_row_map != nullptr && dst_row_map != nullptr)
    transpose_kernel<true, true><<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, dst_row_map, src_row_map);
  else if (src_row_map != nullptr)
    transpose_kernel<true, false><<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, dst_row_map, src_row_map);
  else if (dst_row_map != nullptr)
    transpose_kernel<false, true><<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, dst_row_map, src_row_map);
  else
    transpose_kernel<false, false><<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols, dst_row_map, src_row_map);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

int launch_spmm(const DevCsr& K, int32_t n, const float* XT, float* YT, int64_t ldb, int32_t B, float scale,
                int32_t accumulate, cudaStream_t st) {
  if (!K.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "spmm: matrix not present in this operator");
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(XT, ldb, B, "XT")) return rc;
  if (int rc = check_layout(YT, ldb, B, "YT")) return rc;
  dim3 grid((n + 7) / 8, (B + kWarpSamples - 1) / kWarpSamples);
  spmm_kernel<<<grid, kThreads, 0, st>>>(K.rowptr, K.col, K.val, n, XT, YT, ldb, B, scale, accumulate);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

int launch_seq(const DevSeqPlan& P, int32_t n, bool backward, const float* XT, const float* u0T,
               const float* fT, float dt, int64_t ldj, int64_t ldb, int32_t B, int32_t T, const float* grad_loss,
               float* outT, float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!P.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "sequence path needs S and A");
  if (B <= 0 || T <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B and T must be positive");
  const int64_t J = (int64_t)B * T;
  if (XT == nullptr || outT == nullptr || ldj < J || ldj % 4 != 0) return fail(FEO_ERR_INVALID_ARGUMENT, "seq: bad XT/outT/ldj");
  if (ldj > (int64_t)INT32_MAX - 2 * kSeqTiles * kWarpSamples) return fail(FEO_ERR_UNSUPPORTED, "seq: B * T too large for one launch");
  if ((reinterpret_cast<uintptr_t>(XT) | reinterpret_cast<uintptr_t>(outT)) % 16 != 0)
    return fail(FEO_ERR_INVALID_ARGUMENT, "seq: XT/outT must be 16-byte aligned");
  const int by = (int)((J + kSeqTiles * kWarpSamples - 1) / (kSeqTiles * kWarpSamples));
  dim3 grid((n + 7) / 8, by);
  if (!backward) {
    if (u0T == nullptr || fT == nullptr || ldb < B || loss_out == nullptr)
      return fail(FEO_ERR_INVALID_ARGUMENT, "seq fwd: bad u0T/fT/ldb/loss_out");
    const int count = grid.x * grid.y;
    if (ws == nullptr || ws_bytes < (size_t)count * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
    seq_kernel<false><<<grid, kThreads, 0, st>>>(P.rowptr, P.ent, n, XT, u0T, fT, dt, ldj, ldb, B, T, nullptr, outT, (float*)ws);
    FEO_CUDA_CHECK(cudaGetLastError());
    return finalize((float*)ws, count, 1.0f / (float)T, loss_out, st);
  }
  seq_kernel<true><<<grid, kThreads, 0, st>>>(P.rowptr, P.ent, n, XT, nullptr, nullptr, dt, ldj, ldb, B, T, grad_loss, outT, nullptr);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

// ---------------------------------------------------------------------------------------------
// input synthesis of `closure` (FEONet_steady_Navier-Stokes/train_FEONet.py:337-345): the sin/cos forcing of every
// sample on the resol x resol grid cartesian_prod(linspace(-1, 1, resol)), one kernel instead of ~10 eager ops
//   out[b][0][i][j] = m0 sin(n0 x_i + n1 y_j),  out[b][1][i][j] = m1 cos(n2 x_i + n3 y_j)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sincos_grid_kernel(const float* __restrict__ coeff, int32_t B, int32_t resol,
                                                          float* __restrict__ out) {
  const int b = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || cell >= resol * resol) return;
  const float* c = coeff + (size_t)b * 6;
  const int i = cell / resol, j = cell - i * resol;
  // torch.linspace(-1, 1, n): start + k * step for the first half, end - (n - 1 - k) * step for the second
  const float step = 2.0f / (float)(resol - 1);
  auto lin = [&](int k) { return k < resol / 2 ? -1.0f + step * (float)k : 1.0f - step * (float)(resol - 1 - k); };
  const float x = resol > 1 ? lin(i) : -1.0f, y = resol > 1 ? lin(j) : -1.0f;
  const size_t base = (size_t)b * 2 * resol * resol + cell;
  out[base] = c[0] * sinf(__fadd_rn(__fmul_rn(c[2], x), __fmul_rn(c[3], y)));
  out[base + (size_t)resol * resol] = c[1] * cosf(__fadd_rn(__fmul_rn(c[4], x), __fmul_rn(c[5], y)));
}

int launch_sincos_grid(const float* coeff, int32_t B, int32_t resol, float* out, cudaStream_t st) {
  if (coeff == nullptr || out == nullptr || B <= 0 || resol <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "sincos_grid: bad arguments");
  if (B > 65535) return fail(FEO_ERR_UNSUPPORTED, "sincos_grid: batch too large for one launch");
  dim3 grid((resol * resol + 255) / 256, B);
  sincos_grid_kernel<<<grid, 256, 0, st>>>(coeff, B, resol, out);
  FEO_CUDA_CHECK(cudaGetLastError());
  return FEO_OK;
}

int launch_sq_diff_sum(const float* xT, const float* yT, int32_t n, int64_t ldb, int32_t B, float scale,
                       float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (xT == nullptr || loss_out == nullptr || n <= 0 || B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "sq_diff_sum: bad arguments");
  const int blocks = std::min(n, 1024);
  if (ws == nullptr || ws_bytes < (size_t)blocks * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
  sq_diff_kernel<<<blocks, kThreads, 0, st>>>(xT, yT, n, ldb, B, (float*)ws);
  FEO_CUDA_CHECK(cudaGetLastError());
  return finalize((float*)ws, blocks, scale, loss_out, st);
}

}  // namespace feo

// ---------------------------------------------------------------------------------------------
// dense operator path: CT[n x ldb] = scale * D[n x n] XT[n x ldb]  (- sub)  (+ loss)
// fp32 SIMT GEMM, 64x64x16 tiles, 4x4 register blocking.  D is stored with ld = ceil4(n) and zero
// padding, so its float4 loads are aligned for any n (387, 813, 2549 ... are not multiples of 4).
// The preconditioned operators are 83-93 % dense and tiny (2 B n^2 <= 13
// GFLOP), so exact fp32 FMA keeps the 1e-5 loss tolerance without an error-compensated split.
// ---------------------------------------------------------------------------------------------
namespace feo {
namespace {
constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) dense_apply_kernel(const float* __restrict__ D, int32_t n, int32_t ldd,
                                                          const float* __restrict__ XT, float* __restrict__ CT,
                                                          int64_t ldb, int32_t B, float scale,
                                                          const float* __restrict__ scale_dev,
                                                          const float* __restrict__ sub, float* __restrict__ partials) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int ty = tid >> 4, tx = tid & 15;
  // loader coordinates
  const int a_r = tid >> 2, a_k = (tid & 3) * 4;   // A tile: 64 rows x 16 k, float4 along k
  const int b_k = tid >> 4, b_c = (tid & 15) * 4;  // B tile: 16 k x 64 cols, float4 along cols
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < n; k0 += GBK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + a_r < n && k0 + a_k < ldd) av = __ldg(reinterpret_cast<const float4*>(D + (int64_t)(m0 + a_r) * ldd + k0 + a_k));
    if (k0 + b_k < n && n0 + b_c < ldb) bv = __ldg(reinterpret_cast<const float4*>(XT + (int64_t)(k0 + b_k) * ldb + n0 + b_c));
    __syncthreads();  // previous tile fully consumed
    As[a_k + 0][a_r] = av.x;
    As[a_k + 1][a_r] = av.y;
    As[a_k + 2][a_r] = av.z;
    As[a_k + 3][a_r] = av.w;
    *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const float sc = scale * (scale_dev != nullptr ? __ldg(scale_dev) : 1.0f);
  float lsum = 0.f;
  const int c = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < n && c < ldb) {
      float4 o = make_float4(sc * acc[i][0], sc * acc[i][1], sc * acc[i][2], sc * acc[i][3]);
      if (sub != nullptr) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(sub + (int64_t)m * ldb + c));
        o.x -= s.x;
        o.y -= s.y;
        o.z -= s.z;
        o.w -= s.w;
      }
      if (c + 0 < B) lsum = fmaf(o.x, o.x, lsum);
      if (c + 1 < B) lsum = fmaf(o.y, o.y, lsum);
      if (c + 2 < B) lsum = fmaf(o.z, o.z, lsum);
      if (c + 3 < B) lsum = fmaf(o.w, o.w, lsum);
      *reinterpret_cast<float4*>(CT + (int64_t)m * ldb + c) = o;
    }
  }
  if (partials != nullptr) block_partial(lsum, partials, blockIdx.y * gridDim.x + blockIdx.x);
}
}  // namespace

int launch_dense(const float* D, const float* Dsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
                 const float* scale_dev, const float* sub, float* loss_out, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  if (D == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "dense operator not present in this handle");
  if (B <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "B must be positive");
  if (int rc = check_layout(XT, ldb, B, "XT")) return rc;
  if (int rc = check_layout(CT, ldb, B, "CT")) return rc;
  if (sub != nullptr)
    if (int rc = check_layout(sub, ldb, B, "sub")) return rc;
  const int32_t ldd = (n + 3) / 4 * 4;
  dim3 grid((int)((((B + 3) / 4 * 4) + GBN - 1) / GBN), (n + GBM - 1) / GBM);
  int count = grid.x * grid.y;
  float* partials = nullptr;
  if (loss_out != nullptr) {
    if (ws == nullptr || ws_bytes < (size_t)count * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
    partials = (float*)ws;
  }
  // product path: tcgen05 3xTF32 (feo_dense_tc.cu); FEO_DENSE_SIMT=1 selects the fp32 FMA kernel (developer comparison)
  static const bool simt = [] {
    const char* e = std::getenv("FEO_DENSE_SIMT");
    return e != nullptr && atoi(e) != 0;
  }();
  if (!simt) {
    // the tensor-core kernel writes one partial per 128 x 64 tile, the column tiles padded to the cluster size
    const size_t tc_count = (size_t)(((B + 3) / 4 * 4 + 63) / 64 + 1) * (size_t)((n + 127) / 128);
    if (loss_out != nullptr && ws_bytes < tc_count * sizeof(float)) return fail(FEO_ERR_INVALID_ARGUMENT, "workspace too small");
    // the workspace: loss partials first, then (1 KB aligned) the pre-split activations of the second-generation kernel
    const size_t part_bytes = (tc_count * sizeof(float) + 1023) / 1024 * 1024;
    float* xsplit = nullptr;
    size_t xsplit_bytes = 0;
    if (ws != nullptr && ws_bytes > part_bytes) {
      xsplit = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + part_bytes);
      xsplit_bytes = ws_bytes - part_bytes;
    }
    if (int rc = launch_dense_tc(Dsplit, n, XT, CT, ldb, B, scale, scale_dev, sub, partials, &count, xsplit, xsplit_bytes, st)) return rc;
  } else {
    dense_apply_kernel<<<grid, 256, 0, st>>>(D, n, ldd, XT, CT, ldb, B, scale, scale_dev, sub, partials);
    FEO_CUDA_CHECK(cudaGetLastError());
  }
  if (loss_out != nullptr) return finalize(partials, count, 1.0f, loss_out, st);
  return FEO_OK;
}
}  // namespace feo
