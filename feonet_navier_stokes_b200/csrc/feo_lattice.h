// Lattice plan of the fused residual kernels, third generation (feo_lattice_plan.cpp / feo_lattice.cu).
//
// Why: the tile plan gathers one dof line from shared memory per operator entry and reads a 16-byte coefficient word per
// entry; both go through the load-store pipe, which bounds its kernels at 0.57 / 0.33 of the HBM roofline
// (profiles/r01_ncu_summary.md).  On a STRUCTURED right-diagonal P2-P1 mesh (the reference's `RectangleMesh` set-ups,
// FEONet_steady_Navier-Stokes/assemble_fenics.py:50-56, and BASELINE.json configs[4]) in lattice-lexicographic dof order
//   * the rows of a CELL -- vertex node V = (2ci, 2cj), its edge nodes H, T, D and the pressure dof of V: 9 dofs -- couple
//     only to the 5 x 5 node window around V, so 45 gathered lines feed 9 rows (5 per dof instead of 14.6) forward and 83
//     feed 9 columns backward (9.2 per dof instead of 27);
//   * every cell of a CLASS (interior; edges; corners; the layers next to Dirichlet rows) has the same coefficients, so they
//     are kernel parameters: constant-bank operands of the FMAs, no shared-memory reads and no operator stream at all;
//   * the dofs of a lattice row are one contiguous run, so a CTA sweeping a strip of cells upwards stages TWO boxes per step
//     (the two lattice rows that enter the window) instead of ~50 per tile.
// The pattern of possible couplings is generated (tools/gen_lattice_stencil.py -> feo_lattice_gen.inc); the planner fills
// one coefficient table per cell class from the CSR matrices handed to feo_op_create and VERIFIES that they are fully
// explained by the pattern and the lattice numbering -- identity Dirichlet rows, truncated boundary stencils and natural
// boundaries are just classes.  Anything else (unstructured meshes, other dof orders, cross-component forms, non-colocated
// (I, J) pairings) keeps the tile plan: a set-up choice between device code paths, never a CPU fallback.
#pragma once
#include "feo_internal.h"

namespace feo {

constexpr int kLatMaxClasses = 32;   // coefficient tables that fit the kernel parameters
constexpr int kLatTargets = 4;       // V, H, T, D

struct LatticePlan {
  bool applicable = false;
  std::string why_not;
  bool has_conv = false;
  int32_t n = 0;    // cells per side of the mesh (the lattice has 2n + 1 nodes per side, n + 1 cell origins)
  int32_t nc = 0;   // n + 1
  int32_t N = 0;
  std::vector<int32_t> row_dof0;  // [2n + 5]: dof of u1 at node (0, y) for y = -2 .. 2n + 2 (index y + 2); rows off the lattice: N + 4096
  // per table (0 forward, 1 backward, 2 forward as a matrix-free element walk: the A/B variant of DESIGN.md section 3.5)
  int32_t n_classes[kLatTables] = {0, 0, 0};
  int32_t n_coef[kLatTables] = {0, 0, 0};
  std::vector<uint8_t> cls[kLatTables];    // [nc * nc] class of cell (ci, cj) at cj * nc + ci; class 0 = the most frequent complete one
  uint8_t cat_cls[kLatTables][25];         // the same by boundary-layer categories: cls = cat_cls[lat_cat(cj) * 5 + lat_cat(ci)] (verified)
  std::vector<uint8_t> exist[kLatTables];  // [n_classes] bit t: target node t exists (bit 0 is always set, the pressure dof exists with V)
  std::vector<float> tab[kLatTables];      // [n_classes][n_coef]
  int64_t real_entries = 0;
};

// ns_branch as in build_tile_plan (the backward tables carry the branch sign)
int build_lattice_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i, const int32_t* idx_j,
                       int32_t ns_branch, LatticePlan* out);
// fp64 host replay for one sample: runs the generated cell bodies over the class tables exactly as the kernels do.
// table 0 / 2 (forward): in0 = alpha, in1 = f -> out = r ;  table 1 (backward): in0 = r, in1 = alpha -> out = grad / (2 g)
int replay_lattice_plan(const LatticePlan& L, int table, int32_t ns_branch, const double* in0, const double* in1, double* out);

int launch_lattice_fwd(const feo_operator* op, const DevLatticePlan& L, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                       float* loss_out, float* rT, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_lattice_bwd(const feo_operator* op, const DevLatticePlan& L, const float* alphaT, const float* rT, const float* grad_loss,
                       float* gradT, int64_t ldb, int32_t B, cudaStream_t st);
int lattice_fwd_warps();  // consumer warps of the forward kernel (loss partials per CTA)

// lattice geometry shared by the planner, the replay and the kernels: position (in lines) of node x within the staged run
// of a lattice row; even rows hold (u1, u2, p) at even x and (u1, u2) at odd x, odd rows (u1, u2) everywhere
// boundary-layer category of a cell index: 0, 1 = the first two, 3, 4 = the last two, 2 = everything between.  A cell's class
// depends on (lat_cat(cj), lat_cat(ci)) only: boundary conditions change the rows of boundary dofs, which reach two layers
// of cells through the transposed (backward) stencils.
__host__ __device__ constexpr int lat_cat(int i, int nc) { return i < 2 ? i : (i >= nc - 2 ? 4 - (nc - 1 - i) : 2); }
__host__ __device__ constexpr int lat_pos(int x, int odd_row) { return odd_row ? 2 * x : (x >> 1) * 5 + (x & 1) * 3; }

}  // namespace feo
