// Host planner and fp64 replay of the lattice plan (feo_lattice.h): recognises the structured right-diagonal P2-P1 lattice in
// lexicographic "interleaved" dof order from the CSR matrices and idx_sol alone, fills one coefficient table per cell class
// and verifies that the matrices are fully explained by the generated pattern.  Pure host code.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "feo_lattice.h"
#include "feo_lattice_gen.inc"

namespace feo {
namespace {

struct Desc {
  int8_t mat, rdx, rdy, rc, cdx, cdy, cc, is_signed, twin, div, first;
};
#define FEO_DESC_ROW(i, mat, rdx, rdy, rc, cdx, cdy, cc, sg, tw, dv, fi) {mat, rdx, rdy, rc, cdx, cdy, cc, sg, tw, dv, fi},
const Desc kFwdDesc[] = {FEO_LAT_FWD_DESC(FEO_DESC_ROW)};
const Desc kBwdDesc[] = {FEO_LAT_BWD_DESC(FEO_DESC_ROW)};
const Desc kFwdElemDesc[] = {FEO_LAT_FWDE_DESC(FEO_DESC_ROW)};
#undef FEO_DESC_ROW
static_assert(sizeof(kFwdDesc) / sizeof(Desc) == FEO_LAT_FWD_NCOEF, "forward table layout");
static_assert(sizeof(kBwdDesc) / sizeof(Desc) == FEO_LAT_BWD_NCOEF, "backward table layout");
static_assert(sizeof(kFwdElemDesc) / sizeof(Desc) == FEO_LAT_FWDE_NCOEF, "element-walk forward table layout");
const Desc* const kDesc[kLatTables] = {kFwdDesc, kBwdDesc, kFwdElemDesc};
constexpr int kNCoef[kLatTables] = {FEO_LAT_FWD_NCOEF, FEO_LAT_BWD_NCOEF, FEO_LAT_FWDE_NCOEF};
constexpr int kTx[kLatTargets] = {0, 1, 0, 1}, kTy[kLatTargets] = {0, 0, 1, 1};

struct Geometry {
  int32_t n = 0, m = 0, N = 0;
  const int32_t* row0 = nullptr;  // row_dof0 + 2
  // dof of (x, y, comp) or -1 off the lattice; comp 2 = pressure (even x, even y only)
  int32_t dof(int x, int y, int comp) const {
    if (x < 0 || y < 0 || x >= m || y >= m) return -1;
    if (comp == 2 && ((x | y) & 1)) return -1;
    return row0[y] + lat_pos(x, y & 1) + comp;
  }
};

float lookup(const HostCsr& M, int32_t r, int32_t c) {
  if (!M.present() || r < 0 || c < 0) return 0.f;
  const int32_t* b = M.col.data() + M.rowptr[r];
  const int32_t* e = M.col.data() + M.rowptr[r + 1];
  const int32_t* it = std::lower_bound(b, e, c);
  return (it != e && *it == c) ? M.val[(size_t)(it - M.col.data())] : 0.f;
}

int not_applicable(LatticePlan* L, const std::string& why) {
  L->applicable = false;
  L->why_not = why;
  return FEO_OK;
}

}  // namespace

int build_lattice_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i, const int32_t* idx_j,
                       int32_t ns_branch, LatticePlan* L) {
  *L = LatticePlan();
  if (!A.present()) return not_applicable(L, "no sparse A");
  const int32_t N = A.n;
  // ---- lattice size from the dof counts: n_u = (2n + 1)^2, n_p = (n + 1)^2 ----
  // (a linear operator may come without idx_sol: then N = 9 n^2 + 10 n + 3 fixes n, and the pairing is the lattice's own)
  const bool have_idx = n_u > 0 && idx_i != nullptr && idx_j != nullptr;
  if (!have_idx && (B1.present() || B2.present())) return not_applicable(L, "convective operator without idx_sol");
  int32_t m;
  if (have_idx) {
    m = (int32_t)std::llround(std::sqrt((double)n_u));
  } else {
    const int64_t nn = (int64_t)std::llround((-10.0 + std::sqrt(100.0 - 36.0 * (3.0 - (double)N))) / 18.0);
    m = (int32_t)(2 * nn + 1);
    n_u = m * m;
  }
  if (n_u <= 0 || (int64_t)m * m != n_u || m < 5 || (m & 1) == 0) return not_applicable(L, "n_u is not an odd square");
  const int32_t n = (m - 1) / 2, nc = n + 1;
  if ((int64_t)N != 2 * (int64_t)n_u + (int64_t)nc * nc) return not_applicable(L, "dof count does not match a P2-P1 lattice");
  // ---- the numbering the kernels assume: per lattice row one contiguous run, (u1, u2[, p]) per node ----
  L->n = n;
  L->nc = nc;
  L->N = N;
  L->row_dof0.assign((size_t)m + 4, N + 4096);
  {
    int32_t cur = 0;
    for (int y = 0; y < m; ++y) {
      L->row_dof0[(size_t)y + 2] = cur;
      cur += (y & 1) ? 2 * m : 5 * n + 3;
    }
    if (cur != N) return not_applicable(L, "internal: lattice numbering does not add up");
  }
  Geometry G;
  G.n = n;
  G.m = m;
  G.N = N;
  G.row0 = L->row_dof0.data() + 2;
  // every (I[k], J[k]) must be the (u1, u2) pair of one lattice node, every node exactly once
  {
    std::vector<int8_t> role((size_t)N, 2);  // 0: u1 of a node, 1: u2, 2: pressure
    for (int y = 0; y < m; ++y)
      for (int x = 0; x < m; ++x) {
        role[(size_t)G.dof(x, y, 0)] = 0;
        role[(size_t)G.dof(x, y, 1)] = 1;
      }
    std::vector<uint8_t> seen((size_t)N, 0);
    for (int32_t k = 0; have_idx && k < n_u; ++k) {
      const int32_t i = idx_i[k], j = idx_j[k];
      if (i < 0 || i >= N || role[(size_t)i] != 0 || j != i + 1 || seen[(size_t)i])
        return not_applicable(L, "idx_sol is not the interleaved lattice numbering");
      seen[(size_t)i] = 1;
    }
  }
  const bool conv = B1.present() || B2.present();
  L->has_conv = conv && (B1.nnz() > 0 || B2.nnz() > 0);
  const float sgn = ns_branch ? 1.f : -1.f;
  const HostCsr* mats[3] = {&A, &B1, &B2};

  // ---- one table per cell and direction; classes by (existence mask, table) ----
  int64_t matched[kLatTables][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int dir = 0; dir < kLatTables; ++dir) {
    const Desc* D = kDesc[dir];
    const int nd = kNCoef[dir];
    L->n_coef[dir] = nd;
    std::map<std::vector<uint32_t>, int32_t> ids;
    std::vector<std::vector<uint32_t>> keys;
    std::vector<int64_t> freq;
    std::vector<int32_t> cell_cls((size_t)nc * nc);
    std::vector<uint32_t> key((size_t)nd + 1);
    for (int cj = 0; cj < nc; ++cj)
      for (int ci = 0; ci < nc; ++ci) {
        const int ox = 2 * ci, oy = 2 * cj;
        uint32_t ex = 0;
        for (int t = 0; t < kLatTargets; ++t)
          if (G.dof(ox + kTx[t], oy + kTy[t], 0) >= 0) ex |= 1u << t;
        key[0] = ex;
        for (int i = 0; i < nd; ++i) {
          const Desc& d = D[i];
          const int32_t r = G.dof(ox + d.rdx, oy + d.rdy, d.rc), c = G.dof(ox + d.cdx, oy + d.cdy, d.cc);
          float v = lookup(*mats[d.mat], r, c);
          if (d.twin) {
            const float w = lookup(*mats[d.mat], G.dof(ox + d.rdx, oy + d.rdy, 1), G.dof(ox + d.cdx, oy + d.cdy, 1));
            if (w != v) return not_applicable(L, "the two velocity components of a node pair carry different coefficients");
            if (v != 0.f && d.first) matched[dir][d.mat] += 1;  // the (J, J) entry
          }
          if (v != 0.f && d.first) matched[dir][d.mat] += 1;
          if (d.is_signed) v *= sgn;
          if (d.div > 1) v /= (float)d.div;  // element walk: `div` statements share the assembled entry
          key[(size_t)i + 1] = f2u(v == 0.f ? 0.f : v);  // -0 and +0 are one class
        }
        auto it = ids.find(key);
        int32_t id;
        if (it == ids.end()) {
          id = (int32_t)keys.size();
          ids.emplace(key, id);
          keys.push_back(key);
          freq.push_back(0);
        } else {
          id = it->second;
        }
        freq[(size_t)id]++;
        cell_cls[(size_t)cj * nc + ci] = id;
      }
    const int32_t ncls = (int32_t)keys.size();
    if (ncls > kLatMaxClasses) return not_applicable(L, "more cell classes than the kernel parameters hold (non-uniform mesh?)");
    // class 0 = the most frequent class of complete cells (the interior, unless the mesh is tiny): the kernels' fast path
    std::vector<int32_t> order((size_t)ncls), rank((size_t)ncls);
    for (int32_t i = 0; i < ncls; ++i) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
      const bool fa = keys[(size_t)a][0] == 15u, fb = keys[(size_t)b][0] == 15u;
      return fa != fb ? fa : freq[(size_t)a] > freq[(size_t)b];
    });
    for (int32_t i = 0; i < ncls; ++i) rank[(size_t)order[(size_t)i]] = i;
    L->n_classes[dir] = ncls;
    L->exist[dir].assign((size_t)ncls, 0);
    L->tab[dir].assign((size_t)ncls * nd, 0.f);
    for (int32_t i = 0; i < ncls; ++i) {
      const auto& k = keys[(size_t)order[(size_t)i]];
      L->exist[dir][(size_t)i] = (uint8_t)k[0];
      for (int j = 0; j < nd; ++j) L->tab[dir][(size_t)i * nd + j] = u2f(k[(size_t)j + 1]);
    }
    if (L->exist[dir][0] != 15) return not_applicable(L, "no complete cell");
    L->cls[dir].resize((size_t)nc * nc);
    for (size_t i = 0; i < cell_cls.size(); ++i) L->cls[dir][i] = (uint8_t)rank[(size_t)cell_cls[i]];
    // the kernels find a cell's class from its boundary-layer categories: that must reproduce the map
    std::memset(L->cat_cls[dir], 0xff, 25);
    for (int cj = 0; cj < nc; ++cj)
      for (int ci = 0; ci < nc; ++ci) {
        uint8_t& slot = L->cat_cls[dir][lat_cat(cj, nc) * 5 + lat_cat(ci, nc)];
        const uint8_t c = L->cls[dir][(size_t)cj * nc + ci];
        if (slot == 0xff) slot = c;
        if (slot != c) return not_applicable(L, "cell classes are not a function of the boundary layer (non-uniform mesh?)");
      }
    for (int i = 0; i < 25; ++i)
      if (L->cat_cls[dir][i] == 0xff) L->cat_cls[dir][i] = 0;
  }
  // ---- coverage: every stored entry of A, and of the velocity rows of B1 / B2, is held by exactly one table slot.
  // Forward slots hold distinct entries of the cell's own rows, so equal counts mean full coverage; the backward tables hold
  // each entry of A once (columns of the cell) and each convective entry twice (T-terms by column, E-terms by row).
  int64_t want[3] = {A.nnz(), 0, 0};
  {
    std::vector<uint8_t> is_vel((size_t)N, 0);
    for (int32_t k = 0; have_idx && k < n_u; ++k) is_vel[(size_t)idx_i[k]] = is_vel[(size_t)idx_j[k]] = 1;
    for (int mi = 1; mi < 3; ++mi)
      if (mats[mi]->present())
        for (int32_t r = 0; r < N; ++r)
          if (is_vel[(size_t)r]) want[mi] += mats[mi]->rowptr[r + 1] - mats[mi]->rowptr[r];
  }
  for (int mi = 0; mi < 3; ++mi) {
    if (matched[0][mi] != want[mi] || matched[2][mi] != want[mi]) return not_applicable(L, "matrix entries outside the lattice stencil (forward)");
    if (matched[1][mi] != (mi == 0 ? want[mi] : 2 * want[mi])) return not_applicable(L, "matrix entries outside the lattice stencil (backward)");
  }
  L->real_entries = want[0] + want[1] + want[2];
  L->applicable = true;
  return FEO_OK;
}

// ---- fp64 replay: the generated bodies over doubles --------------------------------------------------
int replay_lattice_plan(const LatticePlan& L, int table, int32_t ns_branch, const double* in0, const double* in1, double* out) {
  if (table < 0 || table >= kLatTables) return fail(FEO_ERR_INVALID_ARGUMENT, "lattice replay: unknown table");
  const bool backward = table == 1;
  if (!L.applicable) return fail(FEO_ERR_UNSUPPORTED, "lattice plan not applicable: " + L.why_not);
  Geometry G;
  G.n = L.n;
  G.m = 2 * L.n + 1;
  G.N = L.N;
  G.row0 = L.row_dof0.data() + 2;
  const int dir = table;
  const int nd = L.n_coef[dir];
  const bool precond = ns_branch != 0;
  const double esign = precond ? 1.0 : -1.0;
  for (int32_t i = 0; i < L.N; ++i) out[i] = 0.0;
  for (int cj = 0; cj < L.nc; ++cj)
    for (int ci = 0; ci < L.nc; ++ci) {
      const int cls = L.cls[dir][(size_t)cj * L.nc + ci];
      const float* tab = L.tab[dir].data() + (size_t)cls * nd;
      const uint32_t ex = L.exist[dir][(size_t)cls];
      const int ox = 2 * ci, oy = 2 * cj;
      auto val = [&](const double* src, int dx, int dy, int comp) -> double {
        const int32_t d = G.dof(ox + dx, oy + dy, comp);
        return d >= 0 ? src[d] : 0.0;
      };
#define C(i) ((double)tab[i])
#define BEGIN {
#define END }
      if (!backward) {
        double acc[3][kLatTargets][2] = {}, sacc = 0.0;
#define LDX(v, dx, dy, comp) const double v = val(in0, dx, dy, comp);
#define FV(mat, t, i) acc[mat][t][0] += C(i) * xI; acc[mat][t][1] += C(i) * xJ;
#define FSI(i) sacc += C(i) * xI;
#define FSJ(i) sacc += C(i) * xJ;
#define FP(t, tc, i) acc[0][t][tc] += C(i) * xP;
#define FSP(i) sacc += C(i) * xP;
        if (table == 2) {
          FEO_LAT_FWDE_BODY_A FEO_LAT_FWDE_BODY_B FEO_LAT_FWDE_BODY_C
        } else {
          FEO_LAT_FWD_BODY_A FEO_LAT_FWD_BODY_B FEO_LAT_FWD_BODY_C
        }
#undef LDX
#undef FV
#undef FSI
#undef FSJ
#undef FP
#undef FSP
        for (int t = 0; t < kLatTargets; ++t) {
          if (!((ex >> t) & 1u)) continue;
          const int32_t dI = G.dof(ox + kTx[t], oy + kTy[t], 0), dJ = dI + 1;
          const double d1 = in0[dI], d2 = in0[dJ];
          for (int c = 0; c < 2; ++c) {
            const int32_t row = c ? dJ : dI;
            const double conv = d1 * acc[1][t][c] + d2 * acc[2][t][c];
            out[row] = precond ? acc[0][t][c] - (in1[row] - conv) : acc[0][t][c] - (-in1[row] + conv);
          }
        }
        const int32_t dP = G.dof(ox, oy, 2);
        out[dP] = precond ? sacc - in1[dP] : sacc + in1[dP];
      } else {
        double g[kLatTargets][2] = {}, bu[3][kLatTargets][2] = {}, sacc = 0.0;  // bu[1] = Bu1, bu[2] = Bu2 of the own rows
#define LDR(v, dx, dy, comp) const double v = val(in0, dx, dy, comp);
#define LDA(v, dx, dy, comp) const double v = val(in1, dx, dy, comp);
#define BTA(t, ia) g[t][0] += C(ia) * rI; g[t][1] += C(ia) * rJ;
#define BTB(t, ia, ib1, ib2)                                                                        \
  {                                                                                                 \
    const double T = ((ia) >= 0 ? C((ia) >= 0 ? (ia) : 0) : 0.0) + ((ib1) >= 0 ? C((ib1) >= 0 ? (ib1) : 0) * d1 : 0.0) + \
                     ((ib2) >= 0 ? C((ib2) >= 0 ? (ib2) : 0) * d2 : 0.0);                           \
    g[t][0] += rI * T;                                                                              \
    g[t][1] += rJ * T;                                                                              \
  }
#define BF(mat, t, i) bu[mat][t][0] += C(i) * d1; bu[mat][t][1] += C(i) * d2;
#define BSI(i) sacc += C(i) * rI;
#define BSJ(i) sacc += C(i) * rJ;
#define BP(t, tc, i) g[t][tc] += C(i) * rP;
#define BSP(i) sacc += C(i) * rP;
        FEO_LAT_BWD_BODY_A FEO_LAT_BWD_BODY_B FEO_LAT_BWD_BODY_C
#undef LDR
#undef LDA
#undef BTA
#undef BTB
#undef BF
#undef BSI
#undef BSJ
#undef BP
#undef BSP
        for (int t = 0; t < kLatTargets; ++t) {
          if (!((ex >> t) & 1u)) continue;
          const int32_t dI = G.dof(ox + kTx[t], oy + kTy[t], 0), dJ = dI + 1;
          const double rI = in0[dI], rJ = in0[dJ];
          out[dI] = g[t][0] + esign * (bu[1][t][0] * rI + bu[1][t][1] * rJ);
          out[dJ] = g[t][1] + esign * (bu[2][t][0] * rI + bu[2][t][1] * rJ);
        }
        out[G.dof(ox, oy, 2)] = sacc;
      }
#undef C
#undef BEGIN
#undef END
    }
  return FEO_OK;
}

}  // namespace feo
