// Internal structures of libfeonet_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "feonet_b200.h"

namespace feo {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define FEO_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::feo::fail(FEO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)

// ---- host-side canonical CSR ------------------------------------------------------------------
struct HostCsr {
  int32_t n = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<float> val;
  bool present() const { return !rowptr.empty(); }
  int64_t nnz() const { return (int64_t)col.size(); }
};

// One fused forward entry: r-row gathers alpha[col] once and feeds three accumulators.
struct FwdEntry {  // 16 B, read as int4
  int32_t col;
  float a, b1, b2;
};
struct FwdEntryLin {  // 8 B, read as int2 (no convection)
  int32_t col;
  float a;
};
// Backward, column-owned. Type A: only the linear operator contributes.
struct BwdEntryA {  // 8 B
  int32_t row;
  float a;
};
// Type B: row h also contributes through s*(d1_h*B1[h,c] + d2_h*B2[h,c]); pi/pj = dofs supplying d1/d2.
struct BwdEntryB {  // 32 B, read as 2 x int4
  int32_t row, pi, pj, pad0;
  float a, b1s, b2s, pad1;
};

// The walk order shared by forward and backward: blobs of units, a unit = a velocity pair
// (I[k], J[k]) or a single dof.  Slots are rows (fwd) / columns (bwd) in processing order.
struct HostPlan {
  int32_t n = 0, n_u = 0;
  bool has_conv = false;
  int32_t ns_branch = 0;
  std::vector<int32_t> pi, pj, kind;  // per dof: partner dofs (-1 if none); kind 0 none, 1 I-row, 2 J-row
  std::vector<int32_t> blob_uptr;     // [n_blobs+1] unit offsets
  std::vector<int32_t> unit_ptr;      // [n_units+1] slot offsets
  std::vector<int32_t> slot_row;      // [n_slots]
  std::vector<int32_t> fptr;          // [n_slots+1]
  std::vector<FwdEntry> fent;         // has_conv
  std::vector<FwdEntryLin> fent_lin;  // !has_conv
  std::vector<int32_t> bptrA, bptrB;  // [n_slots+1]
  std::vector<BwdEntryA> bentA;
  std::vector<BwdEntryB> bentB;
  int64_t nnz_union = 0;
  int64_t n_bent_real = 0;  // backward entries before batch padding
  int32_t max_row_nnz = 0;
  int32_t max_blob_fent = 0, max_blob_bentA = 0, max_blob_bentB = 0;
};

// entry streams are padded per row/column to these multiples (= the kernels' load-batch sizes)
constexpr int kPadF = 4, kPadBA = 4, kPadBB = 2;

struct PlanTuning {
  int32_t blob_rows = 64;      // target rows per blob
  int32_t blob_max_ent = 2048; // cap on fused entries per blob (shared-memory staging budget)
};
PlanTuning tuning_from_env();

int canonicalize(const feo_csr& in, int32_t n, const char* name, HostCsr* out);
HostCsr transpose(const HostCsr& a);
HostCsr axpy(const HostCsr& s, float dt, const HostCsr& a);  // S + dt*A
int build_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
               const int32_t* idx_j, int32_t ns_branch, const PlanTuning& tune, HostPlan* plan);

// ---- tile plan for the shared-memory-staged kernels (feo_tiles.cpp / feo_tiled.cu) -----------------
constexpr int kTileSamples = 64;                           // samples per CTA = floats per staged dof line
constexpr int kTileBatchF = 4, kTileBatchB = 2, kTileBatchA = 4;  // steps per load batch
struct TilePlan {
  bool backward = false, has_conv = false;
  int32_t n_tiles = 0, max_lines = 0, max_pairs = 0;
  std::vector<int32_t> tile_line_ptr, tile_pair_ptr;   // [n_tiles+1]
  std::vector<int32_t> line_dof, line_src;             // staged line -> dof id, source (fwd: alpha; bwd: 0 = r, 1 = alpha)
  std::vector<int32_t> pair_a, pair_b;                 // dofs owned by the two half-warps (-1: idle half)
  std::vector<int32_t> pair_li, pair_lj, pair_vel;     // fwd: local lines of alpha[pi], alpha[pj]; velocity flag
  std::vector<int32_t> pair_step_ptr, pair_stepA_ptr;  // [n_pairs+1] step offsets (bwd: convective / plain lists)
  std::vector<FwdEntry> steps_f;                       // 2 per step (half A, half B); col = local line
  std::vector<BwdEntryB> steps_b;
  std::vector<BwdEntryA> steps_a;
};
int build_tile_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                    const int32_t* idx_j, int32_t ns_branch, bool backward, int32_t max_lines, int32_t max_pairs,
                    TilePlan* out);

struct DevTilePlan {
  int32_t n_tiles = 0, max_lines = 0, max_pairs = 0;
  int32_t *tile_line_ptr = nullptr, *tile_pair_ptr = nullptr, *line_dof = nullptr, *line_src = nullptr,
          *pair_a = nullptr, *pair_b = nullptr, *pair_li = nullptr, *pair_lj = nullptr, *pair_vel = nullptr,
          *pair_step_ptr = nullptr, *pair_stepA_ptr = nullptr;
  FwdEntry* steps_f = nullptr;
  BwdEntryB* steps_b = nullptr;
  BwdEntryA* steps_a = nullptr;
};

// ---- device-side operator ---------------------------------------------------------------------
struct DevCsr {
  int32_t* rowptr = nullptr;
  int32_t* col = nullptr;
  float* val = nullptr;
  int64_t nnz = 0;
  bool present() const { return rowptr != nullptr; }
};

}  // namespace feo

struct feo_operator {
  int32_t n = 0, n_u = 0;
  bool has_conv = false, has_seq = false;
  int32_t ns_branch = 0;
  float dt = 0.f;
  feo::DevCsr csr[5], csrT[5];
  int32_t *idx_i = nullptr, *idx_j = nullptr;
  // fused walk
  int32_t n_blobs = 0, n_units = 0, n_slots = 0;
  int32_t *blob_uptr = nullptr, *unit_ptr = nullptr, *slot_row = nullptr, *slot_pi = nullptr, *slot_pj = nullptr,
          *slot_kind = nullptr;
  int32_t* fptr = nullptr;
  void* fent = nullptr;
  int32_t *bptrA = nullptr, *bptrB = nullptr;
  feo::BwdEntryA* bentA = nullptr;
  feo::BwdEntryB* bentB = nullptr;
  int32_t max_blob_fent = 0, max_blob_bentA = 0, max_blob_bentB = 0;
  // shared-memory-staged walk (default path)
  feo::DevTilePlan tiles_f, tiles_b;
  bool use_tiled = true;
  // dense
  float *dM = nullptr, *dMT = nullptr, *dP = nullptr;
  // bookkeeping
  int64_t nnz[5] = {0, 0, 0, 0, 0};
  int64_t nnz_union = 0;
  int32_t max_row_nnz = 0;
  int64_t device_bytes = 0;
  std::vector<void*> allocations;
};

namespace feo {
// kernel launchers (feo_kernels.cu / feo_gemm.cu)
int launch_transpose(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                     const int32_t* dst_row_map, cudaStream_t st);
int launch_residual_fwd(const feo_operator* op, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                        float* loss_out, float* rT, float* eT, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_residual_bwd(const feo_operator* op, const float* alphaT, const float* rT, const float* eT,
                        const float* grad_loss, float* gradT, int64_t ldb, int32_t B, cudaStream_t st);
int launch_spmm(const DevCsr& K, int32_t n, const float* XT, float* YT, int64_t ldb, int32_t B, float scale,
                int32_t accumulate, cudaStream_t st);
int launch_seq(const DevCsr& M, const DevCsr& S, int32_t n, bool backward, const float* XT, const float* u0T,
               const float* fT, float dt, int64_t ldj, int64_t ldb, int32_t B, int32_t T, const float* grad_loss,
               float* outT, float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_sq_diff_sum(const float* xT, const float* yT, int32_t n, int64_t ldb, int32_t B, float scale,
                       float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_dense(const float* D, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
                 const float* scale_dev, const float* sub, float* loss_out, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t loss_partials_needed(int32_t n, int32_t n_blobs, int64_t cols);
int launch_residual_fwd_tiled(const feo_operator* op, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                              float* loss_out, float* rT, float* eT, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_residual_bwd_tiled(const feo_operator* op, const float* alphaT, const float* rT, const float* eT,
                              const float* grad_loss, float* gradT, int64_t ldb, int32_t B, cudaStream_t st);
}  // namespace feo
