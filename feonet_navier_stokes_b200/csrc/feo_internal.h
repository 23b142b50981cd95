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
const std::string& last_error();
#define FEO_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::feo::fail(FEO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)

// ---- host-side canonical CSR (feo_host.cpp) -----------------------------------------------------
struct HostCsr {
  int32_t n = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<float> val;
  bool present() const { return !rowptr.empty(); }
  int64_t nnz() const { return (int64_t)col.size(); }
};
int canonicalize(const feo_csr& in, int32_t n, const char* name, HostCsr* out);
HostCsr transpose(const HostCsr& a);
HostCsr axpy(const HostCsr& s, float dt, const HostCsr& a);  // S + dt*A

// ---- time-dependent path: union pattern of M = S + dt A and S, one 16-byte word per entry (feo_kernels.cu: seq_kernel) ----
struct alignas(16) SeqEnt {
  int32_t col;
  float m, s;  // coefficient of M (applied to x[j]) and of S (applied to the neighbouring time level x[j -/+ 1])
  int32_t pad;
};
struct SeqPlan {
  std::vector<int32_t> rowptr;
  std::vector<SeqEnt> ent;
};
SeqPlan build_seq_plan(const HostCsr& m, const HostCsr& s);  // rows merged by ascending column

// ---- union pattern of A, B1, B2 with the pair structure of idx_sol (feo_tiles.cpp), shared by the tile and patch planners ----
struct UEnt {
  int32_t col;
  float a, b1, b2;
};

struct Front {
  int32_t n = 0;
  bool conv = false;
  float sgn = 1.f;
  std::vector<int32_t> pi, pj, kind, mate;   // per dof: partners, 0 none / 1 I-dof / 2 J-dof, the other dof of the pair
  std::vector<int32_t> smate;                // single dofs matched two by two (they share source rows), -1: alone
  std::vector<int32_t> unit_of, unit_first;  // units: a velocity pair (I[k], J[k]), two matched single dofs, or a single dof
  std::vector<int32_t> ptr;                  // union rows
  std::vector<UEnt> ent;
  std::vector<int32_t> tptr, trow, tsrc;     // transposed union: per column the source rows + index into ent
  int32_t max_row_nnz = 0;

  int unit_rows(int32_t u, int32_t o[2]) const {
    const int32_t r = unit_first[u];
    o[0] = r;
    if (kind[r] == 1) {
      o[1] = mate[r];
      return 2;
    }
    if (kind[r] == 0 && smate[r] >= 0) {
      o[1] = smate[r];
      return 2;
    }
    return 1;
  }
  // entry (h, c) takes part in the convective term: row h is a velocity row and B1 or B2 is stored there
  bool is_conv(int32_t h, const UEnt& e) const { return conv && kind[h] != 0 && (e.b1 != 0.f || e.b2 != 0.f); }
};

int build_front(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                const int32_t* idx_j, int32_t ns_branch, bool match_singles, Front* out);

// ---- tile plan of the fused residual kernels (feo_tiles.cpp / feo_tiled.cu) -----------------------
// A work unit = one TILE of operator rows (forward) / columns (backward) for one SLAB of 64 consecutive
// samples.  A persistent CTA stages the "lines" the tile touches -- line = the 64 samples of one dof,
// 256 B -- in shared memory with 2-D TMA boxes (two stages: the next unit loads while the current one is
// processed), and its consumer warps walk private STREAMS of 16-byte words that are fed through
// shared-memory rings by 1-D bulk copies.
//   forward : a warp processes QUADS, one velocity row pair (pair quad) or one row (row quad, A-quad) per
//             quarter-warp, 8 samples per lane
//   backward: a warp processes DUOS of column pairs, one pair (e.g. the velocity pair (I[k], J[k]))
//             per half-warp, 4 samples per lane
constexpr int kSlab = 64;                  // samples per work unit
constexpr int kLineBytes = kSlab * 4;      // one staged dof line
constexpr int kChunkWords = 32;            // stream words (16 B) per bulk copy: 512 B
constexpr int kRingChunks = 2;             // chunks per warp ring (power of two)
constexpr int kBoxClasses = 5;
constexpr int kBoxRows[kBoxClasses] = {16, 8, 4, 2, 1};  // TMA box heights (dofs) available for staging

struct Word16 {
  uint32_t w[4];
};
inline uint32_t f2u(float f) {
  uint32_t u;
  __builtin_memcpy(&u, &f, 4);
  return u;
}
inline float u2f(uint32_t u) {
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
}

// One staging copy: box class `cls` (kBoxRows[cls] consecutive dofs from dof0) -> shared lines line0...
struct StageBox {  // 8 B
  int32_t dof0;
  uint16_t line0;
  uint8_t cls;
  uint8_t src;  // forward: 0 = alpha.  backward: 0 = r, 1 = alpha
};
struct WarpRange {  // 8 B
  int32_t begin;   // first stream word of the warp (multiple of kChunkWords)
  int32_t n_words; // words the warp consumes (the region is zero padded to a multiple of kChunkWords)
};

// Stream formats (all offsets are LINE indices within the tile; byte offset = index * 256).
//
// forward, per quad, quarter q reads word 4*unit + q; [header unit][spare unit][n_steps step units], n_steps even:
//   header unit : {row (-1: idle), n_steps, line(alpha[pi]) | line(alpha[pj]) << 16, flags (bit 0: velocity row)}
//   step unit   : {line(col) * 256, a, b1, b2}
// forward A-quad (flags bit 2): rows whose entries all have b1 = b2 = 0; n_steps is a multiple of 4
//   packed unit : {off(step 2j), a, off(step 2j + 1), a}
// forward pair quad (flags bit 1), quarter q owns the velocity rows (I[k], J[k]) of a node:
//   header unit : {row I, row J (-1: idle), line(alpha[I]) | line(alpha[J]) << 16, 3 | nS << 8 | nP << 16 | nX << 24}
//   spare unit, then nX X-steps (2 units), nP P-steps (nP even), nS S-steps, one pad unit if nS is odd
//   X-step : {off(col), aI, b1I, b2I} {aJ, b1J, b2J, 0}   one column feeds both rows
//   P-step : {off(col), aI, aJ, 0}                         the same with b1 = b2 = 0 (e.g. pressure columns)
//   S-step : {line(I[m]) | line(J[m]) << 16, a, b1, b2}    row I takes alpha[I[m]], row J takes alpha[J[m]], same coefficients
// backward, per duo, half h reads word 2*k + h; [header][S-steps][V-steps][P-steps][A-steps][X-steps]:
//   header      : w0 = {cI, cJ (-1: idle), nV, nA}   w1 = {nX, line(r[cI]) | line(r[cJ]) << 16, flags (bit 0: E-term), nS | nP << 16}
//   V-step (3)  : w0 = {line(r[hI]) | line(r[hJ]) << 16, line(alpha[pi]) | line(alpha[pj]) << 16, aI, b1I}
//                 w1 = {b2I, aJ, b1J, b2J}   w2 = {f1I, f2I, f1J, f2J}
//       accI += r[hI] * (aI + b1I d1 + b2I d2)   accJ += r[hJ] * (aJ + b1J d1 + b2J d2)      (b* carry the branch sign)
//       Bu1[cI] += f1I d1   Bu2[cI] += f2I d1    Bu1[cJ] += f1J d2   Bu2[cJ] += f2J d2
//   S-step (2)  : a V-step whose two columns share the coefficients (the usual case: A's velocity block is
//                 diag(K, K), B1 = diag(Dx, Dx), B2 = diag(Dy, Dy)):  w0 as above with a, b1;  w1 = {b2, f1, f2, 0}
//   P-step (1)  : {line(r[h]), aI, aJ, 0}      both columns read the same source row h (one gather)
//   A-step (1)  : {line(r[hI]) | line(r[hJ]) << 16, aI, aJ, 0}
//   X-step (2)  : w0 = {line(alpha[x]), c1I, c2I, c1J}  w1 = {c2J, 0, 0, 0}: Bu1[cI] += c1I x, Bu2[cI] += c2I x, Bu1[cJ] += c1J x, ...
struct TileTuning {
  int32_t max_lines = 414;  // staged lines per tile (shared-memory budget: 2 stages x lines x 256 B), fillers included
  int32_t fill_reserve_pct = 6;  // share of max_lines kept free for fillers while a tile grows (when fill_gap > 0)
  int32_t fill_gap = 0;     // runs of needed dofs separated by <= fill_gap unneeded dofs are staged as one run
  int32_t warps = 15;       // consumer warps per CTA (+ 1 producer warp)
  int32_t stages = 2;       // line stages in shared memory (2 or 3)
  bool pair_rows = true;    // forward: walk the two velocity rows of a node together (pair quads)
  bool pack_a_rows = true;  // forward: rows without B1/B2 entries use two steps per stream word (A-quads)
  bool match_singles = true;  // backward: pair single (pressure) dofs that share source rows, once, before tiling
};
TileTuning tile_tuning_from_env(bool backward);

struct TilePlan {
  bool backward = false, has_conv = false;
  int32_t n = 0, warps = 0, stages = 2, n_tiles = 0, max_lines = 0;
  std::vector<int32_t> tile_box_ptr;  // [n_tiles+1]
  std::vector<int32_t> tile_lines;    // [n_tiles]
  std::vector<StageBox> boxes;
  std::vector<WarpRange> warp_range;  // [n_tiles * warps]
  std::vector<Word16> stream;
  // statistics
  int64_t total_lines = 0, n_items = 0, real_entries = 0, slot_entries = 0;
};
// ns_branch: 1 -> r = A a - F + c, 0 -> r = A a + F - c (only meaningful with convection)
int build_tile_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                    const int32_t* idx_j, int32_t ns_branch, bool backward, const TileTuning& tune, TilePlan* out);
// fp64 host replay of a tile plan for one sample, decoding boxes and streams exactly as the kernels do.
// forward: in0 = alpha, in1 = f -> out = r ;  backward: in0 = r, in1 = alpha -> out = grad / (2 g)
int replay_tile_plan(const TilePlan& T, int32_t ns_branch, const double* in0, const double* in1, double* out);

struct DevTilePlan {
  int32_t n_tiles = 0, max_lines = 0, warps = 0, stages = 2;
  int32_t* tile_box_ptr = nullptr;
  int32_t* tile_lines = nullptr;
  StageBox* boxes = nullptr;
  WarpRange* warp_range = nullptr;
  Word16* stream = nullptr;
};

// device copy of a patch plan (feo_patch.h): the second-generation plan of the fused residual kernels
struct LineLoad;
struct RoundInfo;
struct DevPatchPlan {
  bool present = false;
  int32_t warps = 0, producers = 1, pool_lines = 0, stream_cap = 0, n_segments = 0, n_rounds = 0;
  int32_t* seg_ptr = nullptr;
  RoundInfo* rounds = nullptr;
  LineLoad* loads = nullptr;
  uint32_t* stream = nullptr;  // 32-bit slots, rounds start on 16-byte words
};

constexpr int kLatTables = 3;  // forward, backward, forward as an element walk (feo_lattice.h)
// device side of a lattice plan (feo_lattice.h): the third-generation plan for structured P2-P1 lattices
struct DevLatticePlan {
  bool present = false, has_conv = false;
  bool element_walk = false;             // developer knob FEO_LATTICE_ELEMENT=1: the forward kernel runs table 2 (A/B measurement)
  int32_t n = 0, nc = 0;
  uint8_t cat_cls[kLatTables][25];       // class of a cell by boundary-layer category (feo_lattice.h)
  int32_t n_classes[kLatTables] = {0, 0, 0};
  std::vector<uint8_t> exist[kLatTables];  // host copies: the class tables travel as kernel parameters
  std::vector<float> tab[kLatTables];
};

// ---- device-side operator ---------------------------------------------------------------------
struct DevSeqPlan {
  int32_t* rowptr = nullptr;
  SeqEnt* ent = nullptr;
  bool present() const { return rowptr != nullptr; }
};

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
  bool has_conv = false, has_seq = false, has_sparse = false;
  int32_t ns_branch = 0;
  float dt = 0.f;
  feo::DevCsr csr[5], csrT[5];
  feo::DevSeqPlan seq_f, seq_b;  // union (M, S) rows / union (M^T, S^T) rows of the time-dependent residual
  int32_t *idx_i = nullptr, *idx_j = nullptr;
  feo::DevTilePlan tiles_f, tiles_b;
  feo::DevPatchPlan patch_f, patch_b;  // used instead of the tile plans when both are present
  feo::DevLatticePlan lattice;         // used instead of either when present
  // dense
  float *dM = nullptr, *dMT = nullptr, *dP = nullptr;
  float *dMs = nullptr, *dMTs = nullptr, *dPs = nullptr;  // the same matrices pre-split into TF32 hi/lo operand tiles (feo_dense_tc.cu)
  // bookkeeping
  int64_t nnz[5] = {0, 0, 0, 0, 0};
  int64_t nnz_union = 0;
  int32_t max_row_nnz = 0;
  int64_t device_bytes = 0;
  std::vector<void*> allocations;
};

namespace feo {
// kernel launchers (feo_kernels.cu / feo_tiled.cu)
int launch_transpose(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                     const int32_t* dst_row_map, const int32_t* src_row_map, cudaStream_t st);
int launch_spmm(const DevCsr& K, int32_t n, const float* XT, float* YT, int64_t ldb, int32_t B, float scale,
                int32_t accumulate, cudaStream_t st);
int launch_seq(const DevSeqPlan& P, int32_t n, bool backward, const float* XT, const float* u0T,
               const float* fT, float dt, int64_t ldj, int64_t ldb, int32_t B, int32_t T, const float* grad_loss,
               float* outT, float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_sq_diff_sum(const float* xT, const float* yT, int32_t n, int64_t ldb, int32_t B, float scale,
                       float* loss_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_dense(const float* D, const float* Dsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
                 const float* scale_dev, const float* sub, float* loss_out, void* ws, size_t ws_bytes,
                 cudaStream_t st);
// tcgen05 3xTF32 tile GEMM behind launch_dense (feo_dense_tc.cu); writes one loss partial per CT tile.
// Dsplit = dense_split_tiles(D): per (128-row tile, 16-column k-block) one 16 KB block [hi | lo] in the kernel's
// shared-memory operand layout, so that a stage of the D operand is ONE bulk copy.
std::vector<float> dense_split_tiles(const float* src, int32_t n, bool transposed);
int launch_dense_tc(const float* Dsplit, int32_t n, const float* XT, float* CT, int64_t ldb, int32_t B,
                    float scale, const float* scale_dev, const float* sub, float* partials, int* count_out,
                    float* xsplit, size_t xsplit_bytes, cudaStream_t st);
size_t dense_xsplit_bytes(int32_t n, int64_t cols);  // scratch of the pre-split activations (second-generation dense kernel)
int launch_sincos_grid(const float* coeff, int32_t B, int32_t resol, float* out, cudaStream_t st);
size_t loss_partials_needed(int32_t n, int32_t fused_warps, int64_t cols);
size_t fused_partials_needed(int32_t warps);  // floats: one loss partial per (persistent CTA, consumer warp)
int finalize_loss(float* partials, int count, float scale, float* loss_out, cudaStream_t st);
int launch_residual_fwd(const feo_operator* op, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                        float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_residual_bwd(const feo_operator* op, const float* alphaT, const float* rT, const float* grad_loss,
                        float* gradT, int64_t ldb, int32_t B, cudaStream_t st);
}  // namespace feo
