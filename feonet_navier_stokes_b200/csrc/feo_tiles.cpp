// Host-side tile plan of the fused residual kernels (feo_tiled.cu).  Pure host code; the CPU test
// suite checks it through replay_tile_plan (feo_debug_tile_replay), which decodes the staging boxes
// and the per-warp streams exactly as the kernels do.
//
// A tile = a compact set of operator rows (forward) / columns (backward), grown by BFS over the union
// pattern of A, B1, B2, plus the dof "lines" it has to stage in shared memory (one line = the 64
// samples of one dof in the dof-major batch layout).  Lines are sorted by dof so that runs of
// consecutive dofs become a handful of 2-D TMA boxes.  Nothing here assumes a mesh: I, J are opaque
// index lists, the matrices arbitrary CSR (quirks 3, 10).
//
// forward : rows are sorted by length and grouped four by four into QUADS (one row per quarter-warp).
// backward: columns are handled in PAIRS -- the velocity pair (I[k], J[k]) or two single dofs -- so
//   that one gather of r[I[m]], r[J[m]], alpha[I[m]], alpha[J[m]] serves both columns; two pairs of
//   similar length form a DUO (one pair per half-warp).  The backward recomputes Bu1, Bu2 of its own
//   rows (the E-term of SURVEY.md Appendix A.2) from the alpha lines it gathers anyway, so the forward
//   saves nothing but r.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>

#include "feo_internal.h"

namespace feo {

TileTuning tile_tuning_from_env(bool backward) {
  TileTuning t;
  // one persistent CTA per SM: 2 line stages + one 1 KB ring per consumer warp must fit 227 KB of shared memory.
  // Measured at cfg5 (tools/time_kernels.py): forward 19 warps x 374 lines 3.74 ms (15 x 384: 3.95), backward 15 x 384
  // 6.71 ms (19 x 374: 7.02); gap filling of the staging runs (fill_gap > 0) did not pay at any setting.
  t.warps = 15;
  t.max_lines = 414;  // with 2 x 512 B rings per warp; 4-chunk rings and 384 lines: backward 6.10 ms instead of 5.89 ms
  if (const char* s = std::getenv(backward ? "FEO_TILE_LINES_BWD" : "FEO_TILE_LINES_FWD")) t.max_lines = atoi(s);
  if (const char* s = std::getenv(backward ? "FEO_TILE_WARPS_BWD" : "FEO_TILE_WARPS_FWD")) t.warps = atoi(s);
  if (const char* s = std::getenv("FEO_TILE_PAIR_ROWS")) t.pair_rows = atoi(s) != 0;
  if (const char* s = std::getenv("FEO_TILE_PACK_A_ROWS")) t.pack_a_rows = atoi(s) != 0;
  if (const char* s = std::getenv("FEO_TILE_MATCH_SINGLES")) t.match_singles = atoi(s) != 0;
  if (const char* s = std::getenv("FEO_TILE_FILL_GAP")) t.fill_gap = std::min(std::max(atoi(s), 0), 8);
  if (const char* s = std::getenv("FEO_TILE_FILL_RESERVE")) t.fill_reserve_pct = std::min(std::max(atoi(s), 0), 50);
  if (const char* s = std::getenv(backward ? "FEO_TILE_STAGES_BWD" : "FEO_TILE_STAGES_FWD")) t.stages = atoi(s);
  t.stages = std::min(std::max(t.stages, 2), 3);
  t.warps = std::min(std::max(t.warps, 1), 19);
  const int32_t budget = ((232448 - t.warps * (kRingChunks * kChunkWords * 16 + kRingChunks * 8) - 1024) / t.stages - 1024) / kLineBytes;  // stages are 1 KB aligned
  t.max_lines = std::min(std::max(t.max_lines, 32), budget);
  return t;
}

int build_front(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                const int32_t* idx_j, int32_t ns_branch, bool match_singles, Front* out) {
  Front& F = *out;
  const int32_t n = A.n;
  F.n = n;
  if (B1.present() != B2.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "B1 and B2 must be given together");
  F.conv = B1.present() && B2.present() && n_u > 0;
  if (F.conv && (idx_i == nullptr || idx_j == nullptr)) return fail(FEO_ERR_INVALID_ARGUMENT, "convection needs idx_i/idx_j");
  F.sgn = ns_branch ? 1.0f : -1.0f;
  F.pi.assign(n, -1);
  F.pj.assign(n, -1);
  F.kind.assign(n, 0);
  F.mate.assign(n, -1);
  if (F.conv)
    for (int32_t k = 0; k < n_u; ++k) {
      const int32_t i = idx_i[k], j = idx_j[k];
      if (i < 0 || i >= n || j < 0 || j >= n) return fail(FEO_ERR_INVALID_ARGUMENT, "idx_sol entry out of range");
      if (i == j || F.kind[i] != 0 || F.kind[j] != 0)
        return fail(FEO_ERR_UNSUPPORTED, "idx_sol[0]/idx_sol[1] must be duplicate-free and disjoint");
      F.kind[i] = 1;
      F.kind[j] = 2;
      F.pi[i] = F.pi[j] = i;
      F.pj[i] = F.pj[j] = j;
      F.mate[i] = j;
      F.mate[j] = i;
    }
  // union pattern (rows sorted by column) and its transpose
  F.ptr.assign(n + 1, 0);
  for (int32_t r = 0; r < n; ++r) {
    int32_t i = A.rowptr[r], ie = A.rowptr[r + 1];
    int32_t j = F.conv ? B1.rowptr[r] : 0, je = F.conv ? B1.rowptr[r + 1] : 0;
    int32_t k = F.conv ? B2.rowptr[r] : 0, ke = F.conv ? B2.rowptr[r + 1] : 0;
    while (i < ie || j < je || k < ke) {
      const int32_t ca = i < ie ? A.col[i] : INT32_MAX, c1 = j < je ? B1.col[j] : INT32_MAX, c2 = k < ke ? B2.col[k] : INT32_MAX;
      const int32_t c = std::min(ca, std::min(c1, c2));
      UEnt e{c, 0.f, 0.f, 0.f};
      if (ca == c) e.a = A.val[i++];
      if (c1 == c) e.b1 = B1.val[j++];
      if (c2 == c) e.b2 = B2.val[k++];
      F.ent.push_back(e);
    }
    F.ptr[r + 1] = (int32_t)F.ent.size();
    F.max_row_nnz = std::max(F.max_row_nnz, F.ptr[r + 1] - F.ptr[r]);
  }
  F.tptr.assign(n + 1, 0);
  for (const UEnt& e : F.ent) F.tptr[e.col + 1]++;
  for (int32_t c = 0; c < n; ++c) F.tptr[c + 1] += F.tptr[c];
  F.trow.resize(F.ent.size());
  F.tsrc.resize(F.ent.size());
  std::vector<int32_t> cur(F.tptr.begin(), F.tptr.end() - 1);
  for (int32_t r = 0; r < n; ++r)
    for (int32_t k = F.ptr[r]; k < F.ptr[r + 1]; ++k) {
      const int32_t p = cur[F.ent[k].col]++;
      F.trow[p] = r;
      F.tsrc[p] = k;
    }
  // Single dofs (pressure) are matched two by two, once and for all, with the single dof that shares the most
  // source rows: the backward then gathers a shared row once for both columns (P-steps), and because the partner
  // does not depend on the tiling, neither does the summation order.
  F.smate.assign(n, -1);
  if (match_singles) {
    std::vector<int32_t> cnt(n, 0), touched;
    for (int32_t a = 0; a < n; ++a) {
      if (F.kind[a] != 0 || F.smate[a] >= 0) continue;
      touched.clear();
      int32_t n_src = 0;
      for (int32_t p = F.tptr[a]; p < F.tptr[a + 1]; ++p) {
        if (F.ent[F.tsrc[p]].a == 0.f) continue;
        ++n_src;
        const int32_t h = F.trow[p];
        for (int32_t k = F.ptr[h]; k < F.ptr[h + 1]; ++k) {
          const int32_t c = F.ent[k].col;
          if (c != a && F.kind[c] == 0 && F.smate[c] < 0 && F.ent[k].a != 0.f && cnt[c]++ == 0) touched.push_back(c);
        }
      }
      int32_t best = -1;
      for (int32_t c : touched) {
        if (best < 0 || cnt[c] > cnt[best] || (cnt[c] == cnt[best] && c < best)) best = c;
      }
      if (best >= 0 && 4 * cnt[best] >= n_src) {
        F.smate[a] = best;
        F.smate[best] = a;
      }
      for (int32_t c : touched) cnt[c] = 0;
    }
  }
  F.unit_of.assign(n, -1);
  for (int32_t r = 0; r < n; ++r) {
    if (F.unit_of[r] >= 0) continue;
    const int32_t u = (int32_t)F.unit_first.size();
    if (F.kind[r] == 0) {
      F.unit_first.push_back(r);
      F.unit_of[r] = u;
      if (F.smate[r] >= 0) F.unit_of[F.smate[r]] = u;
    } else {
      F.unit_first.push_back(F.pi[r]);
      F.unit_of[F.pi[r]] = F.unit_of[F.pj[r]] = u;
    }
  }
  return FEO_OK;
}

namespace {

// ---- backward pair description -------------------------------------------------------------------
struct VStep {
  int32_t hI, hJ;    // source rows feeding column cI / cJ (-1: none)
  int32_t kpi, kpj;  // dofs whose alpha lines are d1, d2
  float aI, b1I, b2I, aJ, b1J, b2J, f1I, f2I, f1J, f2J;
};
struct AStep {
  int32_t hI, hJ;
  float aI, aJ;
};
struct XStep {
  int32_t x;
  float c1I, c2I, c1J, c2J;
};
struct PairItem {
  int32_t cI = -1, cJ = -1;
  bool vel = false;
  std::vector<VStep> s;  // symmetric V-steps: both columns use the same a, b1, b2, f1, f2 (two words instead of three)
  std::vector<VStep> v;
  std::vector<AStep> p;  // plain steps whose two columns read the SAME source row (one gather)
  std::vector<AStep> a;
  std::vector<XStep> x;
};
// load-store pipe cycles of a duo step (tools/micro5.cu, micro7.cu): gathers 4 each, half-uniform words 2 each
constexpr int64_t kCostS = 20, kCostV = 22, kCostP = 6, kCostA = 10, kCostX = 8, kCostDuo = 24;

struct TEnt {
  int32_t h;
  float a, b1s, b2s;
};
struct FEnt {
  int32_t x;
  float b1, b2;
  bool used;
};

void build_pair(const Front& F, int32_t cI, int32_t cJ, bool fixed_partner, PairItem* out, int64_t* real_entries) {
  PairItem& P = *out;
  P.cI = cI;
  P.cJ = cJ;
  P.vel = F.kind[cI] != 0;
  const int32_t cols[2] = {cI, cJ};
  std::vector<TEnt> conv[2], plain[2];
  std::vector<FEnt> fwd[2];
  for (int s = 0; s < 2; ++s) {
    const int32_t c = cols[s];
    if (c < 0) continue;
    for (int32_t p = F.tptr[c]; p < F.tptr[c + 1]; ++p) {
      const UEnt& e = F.ent[F.tsrc[p]];
      const int32_t h = F.trow[p];
      if (F.is_conv(h, e))
        conv[s].push_back(TEnt{h, e.a, F.sgn * e.b1, F.sgn * e.b2});
      else if (e.a != 0.f)
        plain[s].push_back(TEnt{h, e.a, 0.f, 0.f});
      else
        continue;  // B1/B2 stored on a non-velocity row: no effect on the residual
      ++*real_entries;
    }
    if (F.kind[c] != 0)
      for (int32_t k = F.ptr[c]; k < F.ptr[c + 1]; ++k)
        if (F.is_conv(c, F.ent[k])) fwd[s].push_back(FEnt{F.ent[k].col, F.ent[k].b1, F.ent[k].b2, false});
  }
  // convective entries: those of cI and cJ whose source rows share (pi, pj) ride in one V-step
  std::map<std::pair<int32_t, int32_t>, std::pair<std::vector<int32_t>, std::vector<int32_t>>> groups;
  for (int s = 0; s < 2; ++s)
    for (int32_t t = 0; t < (int32_t)conv[s].size(); ++t) {
      const int32_t h = conv[s][t].h;
      auto& g = groups[{F.pi[h], F.pj[h]}];
      (s == 0 ? g.first : g.second).push_back(t);
    }
  auto take = [](std::vector<FEnt>& list, int32_t x, float* b1, float* b2) {
    for (FEnt& f : list)
      if (!f.used && f.x == x) {
        *b1 = f.b1;
        *b2 = f.b2;
        f.used = true;
        return;
      }
  };
  for (auto& kv : groups) {
    const size_t m = std::max(kv.second.first.size(), kv.second.second.size());
    for (size_t t = 0; t < m; ++t) {
      VStep s{-1, -1, kv.first.first, kv.first.second, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (t < kv.second.first.size()) {
        const TEnt& e = conv[0][kv.second.first[t]];
        s.hI = e.h;
        s.aI = e.a;
        s.b1I = e.b1s;
        s.b2I = e.b2s;
      }
      if (t < kv.second.second.size()) {
        const TEnt& e = conv[1][kv.second.second[t]];
        s.hJ = e.h;
        s.aJ = e.a;
        s.b1J = e.b1s;
        s.b2J = e.b2s;
      }
      if (P.vel) {
        take(fwd[0], s.kpi, &s.f1I, &s.f2I);
        take(fwd[1], s.kpj, &s.f1J, &s.f2J);
      }
      const bool sym = s.hI >= 0 && s.hJ >= 0 && f2u(s.aI) == f2u(s.aJ) && f2u(s.b1I) == f2u(s.b1J) && f2u(s.b2I) == f2u(s.b2J) &&
                       f2u(s.f1I) == f2u(s.f1J) && f2u(s.f2I) == f2u(s.f2J);
      (sym ? P.s : P.v).push_back(s);
    }
  }
  // forward entries of the own rows that no V-step gathers: extra steps
  if (P.vel) {
    std::map<int32_t, XStep> extra;
    for (int s = 0; s < 2; ++s)
      for (const FEnt& f : fwd[s])
        if (!f.used) {
          XStep& x = extra.emplace(f.x, XStep{f.x, 0.f, 0.f, 0.f, 0.f}).first->second;
          if (s == 0) {
            x.c1I += f.b1;
            x.c2I += f.b2;
          } else {
            x.c1J += f.b1;
            x.c2J += f.b2;
          }
        }
    for (auto& kv : extra) P.x.push_back(kv.second);
  }
  // plain entries: each column walks its source rows in increasing order.  A pair whose partner never changes (a
  // velocity pair, two matched single dofs) first takes the source rows BOTH columns read -- one gather serves the
  // two of them (P-steps, e.g. the pressure rows of the divergence block) -- then the rest; two single dofs that
  // merely share a pair slot in this tile are zipped position by position whatever their partner is, so that the
  // summation order, hence the result bits, do not depend on the tiling.
  {
    auto by_h = [](const TEnt& x, const TEnt& y) { return x.h < y.h; };
    std::stable_sort(plain[0].begin(), plain[0].end(), by_h);
    std::stable_sort(plain[1].begin(), plain[1].end(), by_h);
    std::vector<TEnt> rest[2];
    if (fixed_partner) {
      size_t i = 0, j = 0;
      while (i < plain[0].size() && j < plain[1].size()) {
        if (plain[0][i].h == plain[1][j].h) {
          P.p.push_back(AStep{plain[0][i].h, plain[1][j].h, plain[0][i].a, plain[1][j].a});
          ++i;
          ++j;
        } else if (plain[0][i].h < plain[1][j].h) {
          rest[0].push_back(plain[0][i++]);
        } else {
          rest[1].push_back(plain[1][j++]);
        }
      }
      rest[0].insert(rest[0].end(), plain[0].begin() + i, plain[0].end());
      rest[1].insert(rest[1].end(), plain[1].begin() + j, plain[1].end());
    } else {
      rest[0] = plain[0];
      rest[1] = plain[1];
    }
    const size_t m = std::max(rest[0].size(), rest[1].size());
    for (size_t t = 0; t < m; ++t) {
      AStep s{-1, -1, 0.f, 0.f};
      if (t < rest[0].size()) {
        s.hI = rest[0][t].h;
        s.aI = rest[0][t].a;
      }
      if (t < rest[1].size()) {
        s.hJ = rest[1][t].h;
        s.aJ = rest[1][t].a;
      }
      P.a.push_back(s);
    }
  }
}

// ---- forward pair description -----------------------------------------------------------------------
// The two velocity rows (I[k], J[k]) of a node are walked together: a neighbour node whose columns (I[m], J[m])
// carry the same coefficients in both rows is ONE stream word (S-step); any other column feeds both rows from one
// gather (P-step when its B1/B2 coefficients vanish, e.g. pressure columns; X-step otherwise).
struct FwdSym {
  int32_t colI, colJ;
  float a, b1, b2;
};
struct FwdCol {
  int32_t col;
  float aI, b1I, b2I, aJ, b1J, b2J;
  int32_t n_entries;
};
struct FwdPair {
  int32_t rI = -1, rJ = -1;
  std::vector<FwdSym> s;
  std::vector<FwdCol> p, x;
};

void build_fwd_pair(const Front& F, int32_t rI, int32_t rJ, FwdPair* out) {
  FwdPair& P = *out;
  P.rI = rI;
  P.rJ = rJ;
  std::map<int32_t, FwdCol> cols;  // by column, increasing
  for (int side = 0; side < 2; ++side) {
    const int32_t r = side == 0 ? rI : rJ;
    for (int32_t k = F.ptr[r]; k < F.ptr[r + 1]; ++k) {
      const UEnt& e = F.ent[k];
      FwdCol& c = cols.emplace(e.col, FwdCol{e.col, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0}).first->second;
      if (side == 0) {
        c.aI = e.a;
        c.b1I = e.b1;
        c.b2I = e.b2;
      } else {
        c.aJ = e.a;
        c.b1J = e.b1;
        c.b2J = e.b2;
      }
      ++c.n_entries;
    }
  }
  auto same = [](float x, float y) { return f2u(x) == f2u(y); };
  auto zero3 = [](float a, float b, float c) { return f2u(a) == 0u && f2u(b) == 0u && f2u(c) == 0u; };
  // S-steps: column I[m] is stored by row I alone, J[m] by row J alone, with the same coefficients
  for (auto& kv : cols) {
    FwdCol& c = kv.second;
    if (F.kind[c.col] != 1 || c.n_entries != 1 || !zero3(c.aJ, c.b1J, c.b2J)) continue;
    auto jt = cols.find(F.mate[c.col]);
    if (jt == cols.end()) continue;
    FwdCol& d = jt->second;
    if (d.n_entries == 1 && zero3(d.aI, d.b1I, d.b2I) && same(c.aI, d.aJ) && same(c.b1I, d.b1J) && same(c.b2I, d.b2J)) {
      P.s.push_back(FwdSym{c.col, d.col, c.aI, c.b1I, c.b2I});
      c.n_entries = d.n_entries = -1;
    }
  }
  for (auto& kv : cols) {
    const FwdCol& c = kv.second;
    if (c.n_entries < 0) continue;  // part of an S-step
    if (f2u(c.b1I) == 0u && f2u(c.b2I) == 0u && f2u(c.b1J) == 0u && f2u(c.b2J) == 0u)
      P.p.push_back(c);
    else
      P.x.push_back(c);
  }
}

Word16 mk(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return Word16{{a, b, c, d}}; }

}  // namespace

int build_tile_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                    const int32_t* idx_j, int32_t ns_branch, bool backward, const TileTuning& tune, TilePlan* out) {
  Front F;
  if (int rc = build_front(A, B1, B2, n_u, idx_i, idx_j, ns_branch, backward && tune.match_singles, &F)) return rc;
  const int32_t n = F.n;
  const int32_t W = tune.warps;
  TilePlan& T = *out;
  T = TilePlan();
  T.backward = backward;
  T.has_conv = F.conv;
  T.n = n;
  T.warps = W;
  T.stages = tune.stages;
  const int32_t n_units = (int32_t)F.unit_first.size();

  // lines a row (forward) / column (backward) needs.  key = dof (forward) or src * n + dof (backward;
  // src 0 = r, 1 = alpha): sorting by key sorts by (src, dof)
  auto for_each_line = [&](int32_t d, auto&& fn) {
    if (!backward) {
      for (int32_t k = F.ptr[d]; k < F.ptr[d + 1]; ++k) fn((int64_t)F.ent[k].col);
      if (F.kind[d] != 0) {
        fn((int64_t)F.pi[d]);
        fn((int64_t)F.pj[d]);
      }
    } else {
      for (int32_t p = F.tptr[d]; p < F.tptr[d + 1]; ++p) {
        const int32_t h = F.trow[p];
        const UEnt& e = F.ent[F.tsrc[p]];
        const bool cv = F.is_conv(h, e);
        if (!cv && e.a == 0.f) continue;
        fn((int64_t)h);
        if (cv) {
          fn((int64_t)n + F.pi[h]);
          fn((int64_t)n + F.pj[h]);
        }
      }
      if (F.kind[d] != 0) {  // E-term: the residual lines of the own pair and the alpha lines of the row's B entries
        fn((int64_t)F.pi[d]);
        fn((int64_t)F.pj[d]);
        for (int32_t k = F.ptr[d]; k < F.ptr[d + 1]; ++k)
          if (F.is_conv(d, F.ent[k])) fn((int64_t)n + F.ent[k].col);
      }
    }
  };

  // ---- grow tiles --------------------------------------------------------------------------------
  // the needed lines of a tile may take `grow_lines`; the rest of the budget is left for filler lines (see below)
  const int32_t grow_lines = tune.fill_gap > 0 ? tune.max_lines - tune.max_lines * tune.fill_reserve_pct / 100 : tune.max_lines;
  std::vector<char> seen(n_units, 0);
  std::vector<int32_t> stamp((backward ? 2 : 1) * (size_t)n, -1);
  std::deque<int32_t> frontier, q;
  int32_t next_unseen = 0, placed = 0;
  std::vector<std::vector<int32_t>> tile_units;
  std::vector<int64_t> fresh;
  while (placed < n_units) {
    const int32_t tid = (int32_t)tile_units.size();
    tile_units.emplace_back();
    int32_t lines = 0;
    q.clear();
    while (true) {
      if (q.empty()) {
        int32_t seed = -1;
        if (tile_units.back().empty()) {
          while (!frontier.empty() && seed < 0) {
            const int32_t u = frontier.front();
            frontier.pop_front();
            if (!seen[u]) seed = u;
          }
        }
        if (seed < 0 && tile_units.back().empty()) {
          while (next_unseen < n_units && seen[next_unseen]) ++next_unseen;
          if (next_unseen < n_units) seed = next_unseen;
        }
        if (seed < 0) break;
        seen[seed] = 1;
        q.push_back(seed);
      }
      const int32_t u = q.front();
      int32_t rr[2];
      const int32_t nr = F.unit_rows(u, rr);
      fresh.clear();
      for (int32_t t = 0; t < nr; ++t)
        for_each_line(rr[t], [&](int64_t key) {
          if (stamp[key] != tid) {
            stamp[key] = tid;
            fresh.push_back(key);
          }
        });
      const int32_t add = (int32_t)fresh.size();
      if (!tile_units.back().empty() && lines + add > grow_lines) {
        for (int64_t key : fresh) stamp[key] = -1;  // not staged after all
        break;
      }
      if (add > grow_lines)
        return fail(FEO_ERR_UNSUPPORTED, "a row needs more dof lines than a tile can stage; use the dense operator path");
      q.pop_front();
      tile_units.back().push_back(u);
      ++placed;
      lines += add;
      for (int32_t t = 0; t < nr; ++t) {
        if (!backward) {
          for (int32_t k = F.ptr[rr[t]]; k < F.ptr[rr[t] + 1]; ++k) {
            const int32_t v = F.unit_of[F.ent[k].col];
            if (!seen[v]) {
              seen[v] = 1;
              q.push_back(v);
            }
          }
        } else {
          for (int32_t p = F.tptr[rr[t]]; p < F.tptr[rr[t] + 1]; ++p) {
            const int32_t v = F.unit_of[F.trow[p]];
            if (!seen[v]) {
              seen[v] = 1;
              q.push_back(v);
            }
          }
        }
      }
    }
    // whatever is still queued seeds the following tiles (keeps consecutive tiles adjacent)
    for (int32_t u : q) {
      seen[u] = 0;
      frontier.push_back(u);
    }
    if (tile_units.back().empty()) tile_units.pop_back();
  }

  // ---- emit tiles --------------------------------------------------------------------------------
  T.tile_box_ptr.assign(1, 0);
  std::vector<int64_t> keys;
  for (const auto& units : tile_units) {
    // 1. the tile's lines, sorted by (src, dof); runs of consecutive dofs become TMA boxes
    keys.clear();
    for (int32_t u : units) {
      int32_t rr[2];
      const int32_t nr = F.unit_rows(u, rr);
      for (int32_t t = 0; t < nr; ++t) for_each_line(rr[t], [&](int64_t key) { keys.push_back(key); });
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    // The copy engine spends about the same time on a box whatever its height (~50 cycles per SM, tools/micro8.cu),
    // so the number of boxes is what matters: runs separated by a few unneeded dofs are merged by staging the
    // dofs in between as well (smallest gaps first), and run lengths are rounded up to one box where that
    // costs at most a few extra lines.  Fillers are staged but never referenced by the streams.
    {
      struct Run {
        int64_t start, len;
      };
      std::vector<Run> runs;
      for (size_t i = 0; i < keys.size();) {
        size_t j = i + 1;
        while (j < keys.size() && keys[j] == keys[j - 1] + 1 && (keys[j] >= n) == (keys[i] >= n)) ++j;
        runs.push_back(Run{keys[i], (int64_t)(j - i)});
        i = j;
      }
      int64_t budget = (int64_t)tune.max_lines - (int64_t)keys.size();
      auto same_src = [&](int64_t a, int64_t b) { return (a >= n) == (b >= n); };
      for (int64_t g = 1; g <= tune.fill_gap && budget > 0; ++g) {
        std::vector<Run> merged;
        for (const Run& r : runs) {
          if (!merged.empty()) {
            Run& m = merged.back();
            const int64_t gap = r.start - (m.start + m.len);
            if (gap == g && gap <= budget && same_src(m.start, r.start)) {
              m.len += gap + r.len;
              budget -= gap;
              continue;
            }
          }
          merged.push_back(r);
        }
        runs.swap(merged);
      }
      for (size_t i = 0; i < runs.size(); ++i) {
        Run& r = runs[i];
        const int64_t rem = r.len % 16;
        if (rem == 0) continue;
        int64_t c = 1;
        while (c < rem) c *= 2;
        const int64_t junk = c - rem, src_end = r.start >= n ? 2 * (int64_t)n : (int64_t)n;
        const int64_t room = (i + 1 < runs.size() && same_src(runs[i + 1].start, r.start) ? runs[i + 1].start : src_end) - (r.start + r.len);
        if (junk > 0 && junk <= (c >= 16 ? 3 : c >= 8 ? 2 : 1) && junk <= budget && junk <= room) {
          r.len += junk;
          budget -= junk;
        }
      }
      keys.clear();
      for (const Run& r : runs)
        for (int64_t k = 0; k < r.len; ++k) keys.push_back(r.start + k);
    }
    if (keys.size() > 65535) return fail(FEO_ERR_UNSUPPORTED, "tile stages too many lines");
    T.max_lines = std::max<int32_t>(T.max_lines, (int32_t)keys.size());
    T.tile_lines.push_back((int32_t)keys.size());
    T.total_lines += (int64_t)keys.size();
    for (size_t i = 0; i < keys.size();) {
      size_t j = i + 1;
      while (j < keys.size() && keys[j] == keys[j - 1] + 1 && (keys[j] >= n) == (keys[i] >= n)) ++j;
      size_t len = j - i, pos = i;
      for (int cls = 0; cls < kBoxClasses; ++cls)
        while (len >= (size_t)kBoxRows[cls]) {
          const int64_t key = keys[pos];
          T.boxes.push_back(StageBox{(int32_t)(key >= n ? key - n : key), (uint16_t)pos, (uint8_t)cls, (uint8_t)(key >= n ? 1 : 0)});
          pos += kBoxRows[cls];
          len -= kBoxRows[cls];
        }
      i = j;
    }
    if (std::getenv("FEO_PLAN_DEBUG") && T.tile_lines.size() % 4000 == 0) {
      long c[kBoxClasses] = {0, 0, 0, 0, 0};
      for (const StageBox& b : T.boxes) c[b.cls]++;
      fprintf(stderr, "[plan] boxes 16/8/4/2/1 rows: %ld %ld %ld %ld %ld over %ld tiles, %ld lines\n", c[0], c[1], c[2], c[3], c[4], (long)T.tile_lines.size(), (long)T.total_lines);
    }
    T.tile_box_ptr.push_back((int32_t)T.boxes.size());
    auto LINE = [&](int32_t dof, int32_t src) -> uint32_t {
      const int64_t key = (int64_t)src * n + dof;
      return (uint32_t)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
    };

    // 2. work items and their streams, one list of words per item
    std::vector<std::vector<Word16>> item_words;
    std::vector<std::vector<int32_t>> item_parts;  // sizes (words) of the pieces that must not straddle a chunk
    std::vector<int64_t> item_cost;
    if (!backward) {
      std::vector<int32_t> rows;
      std::vector<FwdPair> fpairs;
      for (int32_t u : units) {
        int32_t rr[2];
        const int32_t nr = F.unit_rows(u, rr);
        if (nr == 2 && F.conv && tune.pair_rows && F.kind[rr[0]] != 0) {
          fpairs.emplace_back();
          build_fwd_pair(F, rr[0], rr[1], &fpairs.back());
        } else {
          for (int32_t t = 0; t < nr; ++t) rows.push_back(rr[t]);
        }
      }
      // pair quads: four velocity pairs of similar shape, one per quarter-warp
      std::stable_sort(fpairs.begin(), fpairs.end(), [](const FwdPair& x, const FwdPair& y) {
        if (x.s.size() != y.s.size()) return x.s.size() > y.s.size();
        if (x.p.size() != y.p.size()) return x.p.size() > y.p.size();
        return x.x.size() > y.x.size();
      });
      for (size_t g = 0; g < fpairs.size(); g += 4) {
        const FwdPair* q4[4];
        size_t nS = 0, nP = 0, nX = 0;
        for (int qd = 0; qd < 4; ++qd) {
          q4[qd] = g + qd < fpairs.size() ? &fpairs[g + qd] : nullptr;
          if (q4[qd] != nullptr) {
            nS = std::max(nS, q4[qd]->s.size());
            nP = std::max(nP, q4[qd]->p.size());
            nX = std::max(nX, q4[qd]->x.size());
          }
        }
        nP = (nP + 1) / 2 * 2;  // P-steps are consumed two by two
        if (nS > 255 || nP > 255 || nX > 255) return fail(FEO_ERR_UNSUPPORTED, "a velocity row pair has too many entries for the fused forward plan");
        std::vector<Word16> w;
        for (int qd = 0; qd < 4; ++qd) {
          const FwdPair* P = q4[qd];
          const uint32_t flags = 3u | ((uint32_t)nS << 8) | ((uint32_t)nP << 16) | ((uint32_t)nX << 24);
          if (P != nullptr)
            w.push_back(mk((uint32_t)P->rI, (uint32_t)P->rJ, LINE(F.pi[P->rI], 0) | (LINE(F.pj[P->rI], 0) << 16), flags));
          else
            w.push_back(mk(0xffffffffu, 0xffffffffu, 0u, flags));
        }
        for (int qd = 0; qd < 4; ++qd) w.push_back(mk(0u, 0u, 0u, 0u));  // spare unit
        for (size_t st = 0; st < nX; ++st) {
          for (int half = 0; half < 2; ++half)
            for (int qd = 0; qd < 4; ++qd) {
              const FwdPair* P = q4[qd];
              if (P != nullptr && st < P->x.size()) {
                const FwdCol& c = P->x[st];
                w.push_back(half == 0 ? mk(LINE(c.col, 0) * kLineBytes, f2u(c.aI), f2u(c.b1I), f2u(c.b2I))
                                      : mk(f2u(c.aJ), f2u(c.b1J), f2u(c.b2J), 0u));
                if (half == 0) T.real_entries += c.n_entries;
              } else {
                w.push_back(mk(0u, 0u, 0u, 0u));
              }
              if (half == 0) T.slot_entries += 2;
            }
        }
        for (size_t st = 0; st < nP; ++st)
          for (int qd = 0; qd < 4; ++qd) {
            const FwdPair* P = q4[qd];
            if (P != nullptr && st < P->p.size()) {
              const FwdCol& c = P->p[st];
              w.push_back(mk(LINE(c.col, 0) * kLineBytes, f2u(c.aI), f2u(c.aJ), 0u));
              T.real_entries += c.n_entries;
            } else {
              w.push_back(mk(0u, 0u, 0u, 0u));
            }
            T.slot_entries += 2;
          }
        for (size_t st = 0; st < nS; ++st)
          for (int qd = 0; qd < 4; ++qd) {
            const FwdPair* P = q4[qd];
            if (P != nullptr && st < P->s.size()) {
              const FwdSym& c = P->s[st];
              w.push_back(mk(LINE(c.colI, 0) | (LINE(c.colJ, 0) << 16), f2u(c.a), f2u(c.b1), f2u(c.b2)));
              T.real_entries += 2;
            } else {
              w.push_back(mk(0u, 0u, 0u, 0u));
            }
            T.slot_entries += 2;
          }
        if (nS & 1)
          for (int qd = 0; qd < 4; ++qd) w.push_back(mk(0u, 0u, 0u, 0u));  // pad unit: items stay 8-word aligned
        // load-store pipe cycles per quad step: X 4 + 8, P 2 + 8, S 2 + 16; header, epilogue gathers and stores
        item_cost.push_back(12 * (int64_t)nX + 10 * (int64_t)nP + 18 * (int64_t)nS + 40);
        item_parts.emplace_back(w.size() / 4, 4);
        item_words.push_back(std::move(w));
      }
      // rows whose entries all have b1 = b2 = 0 (pressure rows; every row of an operator without convection) go
      // into A-quads: two steps per stream word, one accumulator per sample
      auto a_only = [&](int32_t r) {
        for (int32_t k = F.ptr[r]; k < F.ptr[r + 1]; ++k)
          if (f2u(F.ent[k].b1) != 0u || f2u(F.ent[k].b2) != 0u) return false;
        return true;
      };
      std::vector<int32_t> rows_a, rows_g;
      for (int32_t r : rows) (tune.pack_a_rows && a_only(r) ? rows_a : rows_g).push_back(r);
      auto by_len = [&](int32_t x, int32_t y) { return F.ptr[x + 1] - F.ptr[x] > F.ptr[y + 1] - F.ptr[y]; };
      std::stable_sort(rows_a.begin(), rows_a.end(), by_len);
      std::stable_sort(rows_g.begin(), rows_g.end(), by_len);
      for (int packed = 1; packed >= 0; --packed) {
        const std::vector<int32_t>& rws = packed ? rows_a : rows_g;
        for (size_t g = 0; g < rws.size(); g += 4) {
          int32_t r4[4], n_steps = 0;
          for (int qd = 0; qd < 4; ++qd) {
            r4[qd] = g + qd < rws.size() ? rws[g + qd] : -1;
            if (r4[qd] >= 0) n_steps = std::max(n_steps, F.ptr[r4[qd] + 1] - F.ptr[r4[qd]]);
          }
          const int32_t group = packed ? 4 : 2;  // steps consumed per loop iteration
          n_steps = (n_steps + group - 1) / group * group;
          std::vector<Word16> w;
          for (int qd = 0; qd < 4; ++qd) {
            const int32_t r = r4[qd];
            const bool vel = r >= 0 && F.kind[r] != 0;
            const uint32_t offs = vel ? (LINE(F.pi[r], 0) | (LINE(F.pj[r], 0) << 16)) : 0u;
            w.push_back(mk((uint32_t)r, (uint32_t)n_steps, offs, packed ? 4u : (vel ? 1u : 0u)));  // A-quads have no convection
          }
          for (int qd = 0; qd < 4; ++qd) w.push_back(mk(0u, 0u, 0u, 0u));  // spare unit: keeps step groups 128-byte aligned
          auto ENT = [&](int32_t r, int32_t st) -> const UEnt* { return r >= 0 && F.ptr[r] + st < F.ptr[r + 1] ? &F.ent[F.ptr[r] + st] : nullptr; };
          if (packed) {
            for (int32_t st = 0; st < n_steps; st += 2)
              for (int qd = 0; qd < 4; ++qd) {
                const UEnt *e0 = ENT(r4[qd], st), *e1 = ENT(r4[qd], st + 1);
                w.push_back(mk(e0 ? LINE(e0->col, 0) * kLineBytes : 0u, e0 ? f2u(e0->a) : 0u, e1 ? LINE(e1->col, 0) * kLineBytes : 0u,
                               e1 ? f2u(e1->a) : 0u));
                T.real_entries += (e0 ? 1 : 0) + (e1 ? 1 : 0);
                T.slot_entries += 2;
              }
            item_cost.push_back(9 * (int64_t)n_steps + 12);
            item_parts.emplace_back((size_t)(1 + n_steps / 4), 8);  // header pair, groups of two packed units: 8 words each
          } else {
            for (int32_t st = 0; st < n_steps; ++st)
              for (int qd = 0; qd < 4; ++qd) {
                const UEnt* e = ENT(r4[qd], st);
                if (e != nullptr) {
                  w.push_back(mk(LINE(e->col, 0) * kLineBytes, f2u(e->a), f2u(e->b1), f2u(e->b2)));
                  ++T.real_entries;
                } else {
                  w.push_back(mk(0u, 0u, 0u, 0u));  // padding: line 0 of the tile with zero coefficients
                }
                ++T.slot_entries;
              }
            item_cost.push_back(10 * (int64_t)n_steps + 12);
            item_parts.emplace_back((size_t)(1 + n_steps / 2), 8);  // header pair, step pairs: 8 words each
          }
          item_words.push_back(std::move(w));
        }
      }
    } else {
      std::vector<PairItem> pairs;
      int32_t pending = -1;
      for (int32_t u : units) {
        int32_t rr[2];
        if (F.unit_rows(u, rr) == 2) {
          pairs.emplace_back();
          build_pair(F, rr[0], rr[1], true, &pairs.back(), &T.real_entries);
        } else if (pending < 0) {
          pending = rr[0];
        } else {
          pairs.emplace_back();
          build_pair(F, pending, rr[0], false, &pairs.back(), &T.real_entries);
          pending = -1;
        }
      }
      if (pending >= 0) {
        pairs.emplace_back();
        build_pair(F, pending, -1, false, &pairs.back(), &T.real_entries);
      }
      std::stable_sort(pairs.begin(), pairs.end(), [](const PairItem& x, const PairItem& y) {
        if (x.s.size() != y.s.size()) return x.s.size() > y.s.size();
        if (x.v.size() != y.v.size()) return x.v.size() > y.v.size();
        if (x.p.size() != y.p.size()) return x.p.size() > y.p.size();
        if (x.a.size() != y.a.size()) return x.a.size() > y.a.size();
        return x.x.size() > y.x.size();
      });
      const PairItem idle;
      for (size_t g = 0; g < pairs.size(); g += 2) {
        const PairItem* pp[2] = {&pairs[g], g + 1 < pairs.size() ? &pairs[g + 1] : &idle};
        const uint32_t nS = (uint32_t)std::max(pp[0]->s.size(), pp[1]->s.size());
        const uint32_t nV = (uint32_t)std::max(pp[0]->v.size(), pp[1]->v.size());
        const uint32_t nP = (uint32_t)std::max(pp[0]->p.size(), pp[1]->p.size());
        const uint32_t nA = (uint32_t)std::max(pp[0]->a.size(), pp[1]->a.size());
        const uint32_t nX = (uint32_t)std::max(pp[0]->x.size(), pp[1]->x.size());
        if (nS > 65535 || nP > 65535) return fail(FEO_ERR_UNSUPPORTED, "a column has too many entries for the fused backward plan");
        std::vector<Word16> w;
        for (int h = 0; h < 2; ++h) w.push_back(mk((uint32_t)pp[h]->cI, (uint32_t)pp[h]->cJ, nV, nA));
        for (int h = 0; h < 2; ++h) {
          const PairItem& P = *pp[h];
          const uint32_t own = P.vel ? (LINE(P.cI, 0) | (LINE(P.cJ, 0) << 16)) : 0u;
          w.push_back(mk(nX, own, P.vel ? 1u : 0u, nS | (nP << 16)));
        }
        auto r_lines = [&](int32_t hI, int32_t hJ) { return LINE(hI >= 0 ? hI : hJ, 0) | (LINE(hJ >= 0 ? hJ : hI, 0) << 16); };
        for (uint32_t s = 0; s < nS; ++s) {
          Word16 ws[2][2];
          for (int h = 0; h < 2; ++h) {
            ws[h][0] = ws[h][1] = mk(0u, 0u, 0u, 0u);
            if (s < pp[h]->s.size()) {
              const VStep& v = pp[h]->s[s];
              ws[h][0] = mk(r_lines(v.hI, v.hJ), LINE(v.kpi, 1) | (LINE(v.kpj, 1) << 16), f2u(v.aI), f2u(v.b1I));
              ws[h][1] = mk(f2u(v.b2I), f2u(v.f1I), f2u(v.f2I), 0u);
            }
            T.slot_entries += 2;
          }
          for (int k = 0; k < 2; ++k)
            for (int h = 0; h < 2; ++h) w.push_back(ws[h][k]);
        }
        for (uint32_t s = 0; s < nV; ++s) {
          Word16 ws[2][3];
          for (int h = 0; h < 2; ++h) {
            ws[h][0] = ws[h][1] = ws[h][2] = mk(0u, 0u, 0u, 0u);
            if (s < pp[h]->v.size()) {
              const VStep& v = pp[h]->v[s];
              ws[h][0] = mk(r_lines(v.hI, v.hJ), LINE(v.kpi, 1) | (LINE(v.kpj, 1) << 16), f2u(v.aI), f2u(v.b1I));
              ws[h][1] = mk(f2u(v.b2I), f2u(v.aJ), f2u(v.b1J), f2u(v.b2J));
              ws[h][2] = mk(f2u(v.f1I), f2u(v.f2I), f2u(v.f1J), f2u(v.f2J));
            }
            T.slot_entries += 2;
          }
          for (int k = 0; k < 3; ++k)
            for (int h = 0; h < 2; ++h) w.push_back(ws[h][k]);
        }
        for (uint32_t s = 0; s < nP; ++s)
          for (int h = 0; h < 2; ++h) {
            Word16 x = mk(0u, 0u, 0u, 0u);
            if (s < pp[h]->p.size()) {
              const AStep& a = pp[h]->p[s];
              x = mk(LINE(a.hI, 0), f2u(a.aI), f2u(a.aJ), 0u);
            }
            T.slot_entries += 2;
            w.push_back(x);
          }
        for (uint32_t s = 0; s < nA; ++s)
          for (int h = 0; h < 2; ++h) {
            Word16 x = mk(0u, 0u, 0u, 0u);
            if (s < pp[h]->a.size()) {
              const AStep& a = pp[h]->a[s];
              x = mk(r_lines(a.hI, a.hJ), f2u(a.aI), f2u(a.aJ), 0u);
            }
            T.slot_entries += 2;
            w.push_back(x);
          }
        for (uint32_t s = 0; s < nX; ++s) {
          Word16 ws[2][2];
          for (int h = 0; h < 2; ++h) {
            ws[h][0] = ws[h][1] = mk(0u, 0u, 0u, 0u);
            if (s < pp[h]->x.size()) {
              const XStep& x = pp[h]->x[s];
              ws[h][0] = mk(LINE(x.x, 1), f2u(x.c1I), f2u(x.c2I), f2u(x.c1J));
              ws[h][1] = mk(f2u(x.c2J), 0u, 0u, 0u);
            }
          }
          for (int k = 0; k < 2; ++k)
            for (int h = 0; h < 2; ++h) w.push_back(ws[h][k]);
        }
        item_cost.push_back(kCostS * nS + kCostV * nV + kCostP * nP + kCostA * nA + kCostX * nX + kCostDuo);
        std::vector<int32_t> parts(1, 4);
        parts.insert(parts.end(), nS, 4);
        parts.insert(parts.end(), nV, 6);
        parts.insert(parts.end(), nP, 2);
        parts.insert(parts.end(), nA, 2);
        parts.insert(parts.end(), nX, 4);
        item_parts.push_back(std::move(parts));
        item_words.push_back(std::move(w));
      }
    }
    T.n_items += (int64_t)item_words.size();

    // 3. longest-processing-time assignment of the items to the warps; a warp's items are contiguous in the stream
    std::vector<int32_t> order(item_words.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return item_cost[x] > item_cost[y]; });
    // Ties go to the first warp of a scan whose start and direction change from tile to tile: a consumer warp
    // may run one unit ahead of the slowest one, so the warps that get the larger share must not be the
    // same in consecutive units.
    std::vector<int64_t> load(W, 0);
    std::vector<std::vector<int32_t>> mine(W);
    const uint32_t tile_no = (uint32_t)T.tile_lines.size() - 1u;
    const uint32_t hsh = tile_no * 2654435761u;
    const int32_t w0 = (int32_t)((hsh >> 8) % (uint32_t)W), dir = (hsh >> 7) & 1u ? 1 : W - 1;
    for (int32_t it : order) {
      int32_t best = w0;
      for (int32_t k = 1, w = (w0 + dir) % W; k < W; ++k, w = (w + dir) % W)
        if (load[w] < load[best]) best = w;
      load[best] += item_cost[it];
      mine[best].push_back(it);
    }
    if (std::getenv("FEO_PLAN_DEBUG")) {
      static double sum_max = 0, sum_mean = 0; static long nt = 0;
      int64_t mx = 0, tot = 0; for (int32_t w = 0; w < W; ++w) { mx = std::max(mx, load[w]); tot += load[w]; }
      sum_max += (double)mx; sum_mean += (double)tot / W; ++nt;
      if (nt % 2000 == 0) fprintf(stderr, "[plan] tiles %ld items/tile %.1f balance mean/max = %.3f\n", nt, (double)item_words.size(), sum_mean / sum_max);
    }
    for (int32_t w = 0; w < W; ++w) {
      WarpRange R{(int32_t)T.stream.size(), 0};
      for (int32_t it : mine[w]) {
        // no piece straddles a chunk: the reader applies the same rule (skip to the next chunk when a piece does not fit)
        size_t at = 0;
        for (int32_t len : item_parts[it]) {
          while ((T.stream.size() - (size_t)R.begin) % kChunkWords + (size_t)len > (size_t)kChunkWords) T.stream.push_back(mk(0u, 0u, 0u, 0u));
          T.stream.insert(T.stream.end(), item_words[it].begin() + at, item_words[it].begin() + at + len);
          at += (size_t)len;
        }
        R.n_words = (int32_t)(T.stream.size() - (size_t)R.begin);
      }
      while (T.stream.size() % kChunkWords != 0) T.stream.push_back(mk(0u, 0u, 0u, 0u));
      T.warp_range.push_back(R);
    }
    if (T.stream.size() >= (size_t)INT32_MAX / 2) return fail(FEO_ERR_UNSUPPORTED, "operator stream too large");
  }
  T.n_tiles = (int32_t)T.tile_lines.size();
  return FEO_OK;
}

// ---------------------------------------------------------------------------------------------
// fp64 replay: decodes boxes and streams exactly as feo_tiled.cu does
// ---------------------------------------------------------------------------------------------
int replay_tile_plan(const TilePlan& T, int32_t ns_branch, const double* in0, const double* in1, double* out) {
  const bool precond = T.has_conv ? ns_branch != 0 : true;
  const double esign = precond ? 1.0 : -1.0;
  const int32_t W = T.warps;
  std::vector<double> smem;
  std::vector<char> valid;
  for (int32_t t = 0; t < T.n_tiles; ++t) {
    const int32_t nl = T.tile_lines[t];
    smem.assign(nl, 0.0);
    valid.assign(nl, 0);
    for (int32_t b = T.tile_box_ptr[t]; b < T.tile_box_ptr[t + 1]; ++b) {
      const StageBox& bx = T.boxes[b];
      for (int32_t k = 0; k < kBoxRows[bx.cls]; ++k) {
        const int32_t l = bx.line0 + k, dof = bx.dof0 + k;
        if (l >= nl || dof >= T.n || valid[l]) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: staging boxes overlap or overflow");
        valid[l] = 1;
        smem[l] = !T.backward ? in0[dof] : (bx.src ? in1[dof] : in0[dof]);
      }
    }
    for (int32_t l = 0; l < nl; ++l)
      if (!valid[l]) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: a staged line is not covered by a box");
    bool bad = false;
    auto S = [&](uint32_t line) -> double {
      if ((int32_t)line >= nl) {
        bad = true;
        return 0.0;
      }
      return smem[line];
    };
    for (int32_t w = 0; w < W; ++w) {
      const WarpRange& R = T.warp_range[(size_t)t * W + w];
      if (R.begin % kChunkWords != 0) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: warp stream not chunk aligned");
      const Word16* const s0 = T.stream.data() + R.begin;
      const Word16* s = s0;
      const Word16* const end = s + R.n_words;
      auto fit = [&](int32_t len) {  // skip to the next chunk when the next piece does not fit (as the kernels do)
        const int32_t at = (int32_t)(s - s0) % kChunkWords;
        if (at + len > kChunkWords) s += kChunkWords - at;
      };
      while (s < end) {
        if (!T.backward && (s[0].w[3] & 2u)) {  // pair quad
          const uint32_t fl = s[0].w[3];
          const uint32_t nS = (fl >> 8) & 255u, nP = (fl >> 16) & 255u, nX = fl >> 24;
          for (int qd = 0; qd < 4; ++qd) {
            const Word16& H = s[qd];
            if (H.w[3] != fl) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: pair quad step counts differ");
            double aI = 0, uI = 0, vI = 0, aJ = 0, uJ = 0, vJ = 0;
            const Word16* q = s + 8;
            for (uint32_t st = 0; st < nX; ++st, q += 8) {
              const Word16 &e0 = q[qd], &e1 = q[4 + qd];
              if (e0.w[0] % kLineBytes != 0) bad = true;
              const double x = S(e0.w[0] / kLineBytes);
              aI += (double)u2f(e0.w[1]) * x;
              uI += (double)u2f(e0.w[2]) * x;
              vI += (double)u2f(e0.w[3]) * x;
              aJ += (double)u2f(e1.w[0]) * x;
              uJ += (double)u2f(e1.w[1]) * x;
              vJ += (double)u2f(e1.w[2]) * x;
            }
            for (uint32_t st = 0; st < nP; ++st, q += 4) {
              const Word16& e = q[qd];
              if (e.w[0] % kLineBytes != 0) bad = true;
              const double x = S(e.w[0] / kLineBytes);
              aI += (double)u2f(e.w[1]) * x;
              aJ += (double)u2f(e.w[2]) * x;
            }
            for (uint32_t st = 0; st < nS; ++st, q += 4) {
              const Word16& e = q[qd];
              const double xI = S(e.w[0] & 0xffffu), xJ = S(e.w[0] >> 16);
              aI += (double)u2f(e.w[1]) * xI;
              uI += (double)u2f(e.w[2]) * xI;
              vI += (double)u2f(e.w[3]) * xI;
              aJ += (double)u2f(e.w[1]) * xJ;
              uJ += (double)u2f(e.w[2]) * xJ;
              vJ += (double)u2f(e.w[3]) * xJ;
            }
            const int32_t rI = (int32_t)H.w[0], rJ = (int32_t)H.w[1];
            if (rI < 0) continue;
            const double d1 = S(H.w[2] & 0xffffu), d2 = S(H.w[2] >> 16);
            const double cI = d1 * uI + d2 * vI, cJ = d1 * uJ + d2 * vJ;
            out[rI] = precond ? aI - (in1[rI] - cI) : aI - (-in1[rI] + cI);
            out[rJ] = precond ? aJ - (in1[rJ] - cJ) : aJ - (-in1[rJ] + cJ);
          }
          s += 8 + 8 * (size_t)nX + 4 * (size_t)nP + 4 * (size_t)nS + ((nS & 1u) ? 4 : 0);
        } else if (!T.backward && (s[0].w[3] & 4u)) {  // A-quad: two steps per word, coefficients of A only
          const int32_t n_steps = (int32_t)s[0].w[1];
          for (int qd = 0; qd < 4; ++qd) {
            const Word16& H = s[qd];
            if ((int32_t)H.w[1] != n_steps || !(H.w[3] & 4u)) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: quad step counts differ");
            double accA = 0;
            for (int32_t st = 0; st < n_steps; ++st) {
              const Word16& e = s[8 + 4 * (st / 2) + qd];
              const uint32_t off = e.w[2 * (st & 1)];
              if (off % kLineBytes != 0) bad = true;
              accA += (double)u2f(e.w[2 * (st & 1) + 1]) * S(off / kLineBytes);
            }
            const int32_t row = (int32_t)H.w[0];
            if (row < 0) continue;
            const double f = in1[row];
            out[row] = precond ? accA - (f - 0.0) : accA - (-f + 0.0);
          }
          s += 8 + 2 * (size_t)n_steps;
        } else if (!T.backward) {
          const int32_t n_steps = (int32_t)s[0].w[1];
          for (int qd = 0; qd < 4; ++qd) {
            const Word16& H = s[qd];
            if ((int32_t)H.w[1] != n_steps) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: quad step counts differ");
            double accA = 0, acc1 = 0, acc2 = 0;
            for (int32_t st = 0; st < n_steps; ++st) {
              const Word16& e = s[8 + 4 * st + qd];
              if (e.w[0] % kLineBytes != 0) bad = true;
              const double x = S(e.w[0] / kLineBytes);
              accA += (double)u2f(e.w[1]) * x;
              acc1 += (double)u2f(e.w[2]) * x;
              acc2 += (double)u2f(e.w[3]) * x;
            }
            const int32_t row = (int32_t)H.w[0];
            if (row < 0) continue;
            const double c = (H.w[3] & 1u) ? S(H.w[2] & 0xffffu) * acc1 + S(H.w[2] >> 16) * acc2 : 0.0;
            const double f = in1[row];
            out[row] = precond ? accA - (f - c) : accA - (-f + c);
          }
          s += 8 + 4 * (size_t)n_steps;
        } else {
          fit(4);
          const uint32_t nV = s[0].w[2], nA = s[0].w[3], nX = s[2].w[0], nS = s[2].w[3] & 0xffffu, nP = s[2].w[3] >> 16;
          for (int h = 0; h < 2; ++h) {
            const Word16 &H0 = s[h], &H1 = s[2 + h];
            if (H0.w[2] != nV || H0.w[3] != nA || H1.w[0] != nX || H1.w[3] != s[2].w[3])
              return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: duo step counts differ");
            double accI = 0, accJ = 0, bu1I = 0, bu2I = 0, bu1J = 0, bu2J = 0;
            const Word16* p = s + 4;
            auto fitp = [&](int32_t len) {
              const int32_t at = (int32_t)(p - s0) % kChunkWords;
              if (at + len > kChunkWords) p += kChunkWords - at;
            };
            for (uint32_t v = 0; v < nS; ++v, p += 4) {
              fitp(4);
              const Word16 &w0 = p[h], &w1 = p[2 + h];
              const double rI = S(w0.w[0] & 0xffffu), rJ = S(w0.w[0] >> 16), d1 = S(w0.w[1] & 0xffffu), d2 = S(w0.w[1] >> 16);
              const double t = (double)u2f(w0.w[2]) + (double)u2f(w0.w[3]) * d1 + (double)u2f(w1.w[0]) * d2;
              accI += rI * t;
              accJ += rJ * t;
              bu1I += (double)u2f(w1.w[1]) * d1;
              bu2I += (double)u2f(w1.w[2]) * d1;
              bu1J += (double)u2f(w1.w[1]) * d2;
              bu2J += (double)u2f(w1.w[2]) * d2;
            }
            for (uint32_t v = 0; v < nV; ++v, p += 6) {
              fitp(6);
              const Word16 &w0 = p[h], &w1 = p[2 + h], &w2 = p[4 + h];
              const double rI = S(w0.w[0] & 0xffffu), rJ = S(w0.w[0] >> 16), d1 = S(w0.w[1] & 0xffffu), d2 = S(w0.w[1] >> 16);
              accI += rI * ((double)u2f(w0.w[2]) + (double)u2f(w0.w[3]) * d1 + (double)u2f(w1.w[0]) * d2);
              accJ += rJ * ((double)u2f(w1.w[1]) + (double)u2f(w1.w[2]) * d1 + (double)u2f(w1.w[3]) * d2);
              bu1I += (double)u2f(w2.w[0]) * d1;
              bu2I += (double)u2f(w2.w[1]) * d1;
              bu1J += (double)u2f(w2.w[2]) * d2;
              bu2J += (double)u2f(w2.w[3]) * d2;
            }
            for (uint32_t a = 0; a < nP; ++a, p += 2) {
              fitp(2);
              const Word16& w0 = p[h];
              const double rx = S(w0.w[0]);
              accI += (double)u2f(w0.w[1]) * rx;
              accJ += (double)u2f(w0.w[2]) * rx;
            }
            for (uint32_t a = 0; a < nA; ++a, p += 2) {
              fitp(2);
              const Word16& w0 = p[h];
              accI += (double)u2f(w0.w[1]) * S(w0.w[0] & 0xffffu);
              accJ += (double)u2f(w0.w[2]) * S(w0.w[0] >> 16);
            }
            for (uint32_t x = 0; x < nX; ++x, p += 4) {
              fitp(4);
              const Word16 &w0 = p[h], &w1 = p[2 + h];
              const double xv = S(w0.w[0]);
              bu1I += (double)u2f(w0.w[1]) * xv;
              bu2I += (double)u2f(w0.w[2]) * xv;
              bu1J += (double)u2f(w0.w[3]) * xv;
              bu2J += (double)u2f(w1.w[0]) * xv;
            }
            if (H1.w[2] & 1u) {
              const double rI = S(H1.w[1] & 0xffffu), rJ = S(H1.w[1] >> 16);
              accI += esign * (bu1I * rI + bu1J * rJ);
              accJ += esign * (bu2I * rI + bu2J * rJ);
            }
            const int32_t cI = (int32_t)H0.w[0], cJ = (int32_t)H0.w[1];
            if (cI >= 0) out[cI] = accI;
            if (cJ >= 0) out[cJ] = accJ;
            if (h == 1) s = p;  // both halves walk the same pieces
          }
        }
      }
      if (s != end) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: warp stream boundaries inconsistent");
    }
    if (bad) return fail(FEO_ERR_INVALID_ARGUMENT, "tile plan: stream references a line outside the tile");
  }
  return FEO_OK;
}

}  // namespace feo
