// Patch plan of the fused residual kernels, second generation (feo_patch_plan.cpp / feo_patch.cu).
//
// Why: the tile plan (feo_tiles.cpp) gathers every operator entry's alpha line from shared memory once per ROW it feeds;
// its kernels are bound by the load-store pipe (profiles/r01_ncu_summary.md).  On a P2-P1 mesh the 9 neighbours of an
// edge node lie inside the 19 neighbours of either end vertex, so a PATCH = one hub node + up to three nodes whose
// columns the hub already needs (+ one single row, e.g. the pressure dof of the vertex) can be evaluated from ONE gather
// per column: 45 gathered lines feed 9 rows instead of 149 (forward), 83 instead of 243 (backward).  Nothing here knows
// about meshes: patches are found greedily from the CSR patterns, nodes are the positional pairs (I[k], J[k]).
//
// Work decomposition
//   patch   : <= 4 velocity nodes (row pairs forward, column pairs backward) + <= 1 single dof; ONE warp owns it for one
//             slab of 64 samples (2 samples per lane, packed fp32x2 arithmetic, every accumulator in registers)
//   round   : <= W patches, one per consumer warp, that are evaluated at the same time by one CTA
//   segment : a run of consecutive rounds; a persistent CTA processes (segment, slab) items
// Shared memory = a POOL of dof lines (256 B = 64 samples of one dof) + two round-stream buffers.  Lines stay resident
// from one round to the next: the producer warps fetch, with 16-byte cp.async copies, only the lines round g + 1 needs
// and round g does not hold, into slots that round g - 1 has released (the plan assigns the slots).  The same
// mbarrier (full[g & 1]) also tracks the bulk copy of the round's operator stream.
//
// Operator stream of a round (32-bit slots grouped in 16-byte words; every piece starts on a word; "off" = byte offset of
// a line in the pool = slot * 256):
//   table   : per consumer warp {byte offset of its first patch in the region, number of patches}
//   patch   : W0 = {n_runs | has_single << 12, dof of the single (-1: none), 0, 0}
//             4 x {dof I, dof J (-1: slot unused), off(own line I), off(own line J)}
//             runs...
//   run     : {kind | mask << 4 | count << 8, 0, 0, 0} followed by `count` >= 1 steps of the same kind and target mask
//             (mask bit t = node slot t takes part; k = popcount(mask))
//   forward steps (lines are alpha lines):
//     kind 0 (pair)  : [off(I[m]), off(J[m]), sI, sJ, (a, b1, b2) x k]                              -> ceil((4 + 3k) / 4) words
//         x = alpha[I[m]], y = alpha[J[m]]:  A_I[t] += a x, U_I[t] += b1 x, V_I[t] += b2 x, A_J[t] += a y, ...; S += sI x + sJ y
//     kind 1 (plain) : [off(c), s, (aI, aJ) x k]                                                    -> ceil((2 + 2k) / 4) words
//         x = alpha[c]:  A_I[t] += aI x, A_J[t] += aJ x;  S += s x
//   backward steps (r lines and alpha lines share the pool):
//     kind 0 (pair)  : [off(r[I m]), off(r[J m]), off(alpha[I m]), off(alpha[J m]), sI, sJ, (a, b1s, b2s, f1, f2) x k]
//                                                                                                     -> ceil((6 + 5k) / 4) words
//         T = a + b1s d1 + b2s d2;  G_I[t] += rI T, G_J[t] += rJ T;  Bu1_I[t] += f1 d1, Bu2_I[t] += f2 d1, Bu1_J[t] += f1 d2,
//         Bu2_J[t] += f2 d2;  S += sI rI + sJ rJ            (b1s, b2s carry the branch sign)
//     kind 1 (plain) : [off(r[h]), s, (aI, aJ) x k]
// Epilogues: forward  r = A - (F - c) or A - (-F + c), c = fl(fl(d1 U) + fl(d2 V)), loss partial, store r;
//            backward G_I += e (Bu1_I rI + Bu1_J rJ), G_J += e (Bu2_I rI + Bu2_J rJ), store 2 g G.
//
// The plan is APPLICABLE when every convective entry (a velocity row with B1 or B2 != 0 at that column) couples the same
// component of two nodes and its mirror entry in the partner row carries the same coefficients -- A's velocity block
// diag(K, K), B1 = diag(Dx, Dx), B2 = diag(Dy, Dy), which is what assemble_fenics.py emits; identity Dirichlet rows
// qualify.  Anything else (cross-component convection, non-colocated pairings) keeps the tile plan: a set-up choice
// between two device code paths, never a CPU fallback.
#pragma once
#include "feo_internal.h"

namespace feo {

constexpr int kPatchNodes = 4;
constexpr int kPatchHeaderWords = 1 + kPatchNodes;

struct LineLoad {    // 8 B
  uint32_t dof_src;  // dof | src << 31 (backward: src 0 = r, 1 = alpha; forward: 0 = alpha)
  uint32_t slot;     // pool slot the line goes to
};
struct RoundInfo {   // 16 B
  int32_t load_begin, n_loads;     // LineLoad range of the lines this round adds to the pool
  int32_t stream_begin, n_words;   // 16-byte words of the round's stream region
};

struct PatchTuning {
  int32_t warps = 20;         // consumer warps per CTA = patches per round
  int32_t producers = 1;      // producer warps
  int32_t seg_rounds = 48;    // rounds per segment
  int32_t stream_cap = 0;     // bytes of one round-stream buffer (0: warps * 1536 / 2048, rounded up to 1 KB)
  int32_t pool_lines = 0;     // 0: whatever 227 KB leave after the stream buffers
  int32_t merge_pad = 2;        // run classes of a patch are merged while that pads at most this many target slots per merge
  int32_t round_fill_pct = 64;  // lines one round may hold, in percent of the pool
};
PatchTuning patch_tuning_from_env(bool backward);

struct PatchPlan {
  bool backward = false, applicable = false;
  std::string why_not;          // when not applicable
  int32_t n = 0, warps = 0, producers = 1, pool_lines = 0, stream_cap = 0;
  std::vector<int32_t> seg_ptr; // [n_segments + 1] into rounds
  std::vector<RoundInfo> rounds;
  std::vector<LineLoad> loads;
  std::vector<uint32_t> stream;  // 32-bit slots
  // statistics
  int64_t n_patches = 0, n_nodes = 0, n_singles = 0, n_steps = 0, n_gathers = 0, real_entries = 0, slot_entries = 0, max_union_lines = 0;
  int32_t n_segments() const { return (int32_t)seg_ptr.size() - 1; }
};

int build_patch_plan(const Front& F, bool backward, const PatchTuning& tune, PatchPlan* out);
// fp64 host replay of a patch plan for one sample: simulates the pool (slots, residency) and decodes the streams exactly as
// the kernels do.  forward: in0 = alpha, in1 = f -> out = r ;  backward: in0 = r, in1 = alpha -> out = grad / (2 g)
int replay_patch_plan(const PatchPlan& P, bool has_conv, int32_t ns_branch, const double* in0, const double* in1, double* out);

int launch_patch_fwd(const feo_operator* op, const DevPatchPlan& P, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                     float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_patch_bwd(const feo_operator* op, const DevPatchPlan& P, const float* alphaT, const float* rT, const float* grad_loss,
                     float* gradT, int64_t ldb, int32_t B, cudaStream_t st);
int sm_count(int* out);

}  // namespace feo
