// Host-side tile plan for the shared-memory-staged residual kernels (feo_tiled.cu).
//
// A tile = a compact set of operator rows (forward) / columns (backward) grown by BFS over the
// union pattern, plus the list of dof "lines" it has to stage: in the dof-major layout one line is
// the 64 consecutive samples of one dof (256 contiguous bytes), fetched with one bulk-async copy.
// A warp owns a PAIR of rows -- a velocity pair (I[l], J[l]) or two single dofs -- one per
// half-warp, and walks "steps": step k holds one entry for each half.  Entries of the two rows are
// aligned by the unit (velocity pair / single dof) of their column so that lines needed by both
// halves in the same step (pressure columns in the forward, alpha[I[k]], alpha[J[k]] in the
// backward) are one shared-memory broadcast.  Pure host code; checked on CPU via the debug hooks.
#include <algorithm>
#include <deque>
#include <unordered_map>

#include "feo_internal.h"

namespace feo {

namespace {

struct UEnt {
  int32_t col;
  float a, b1, b2;
};

struct Graph {
  int32_t n = 0;
  std::vector<int32_t> ptr;   // union rows
  std::vector<UEnt> ent;
  std::vector<int32_t> tptr;  // transposed union: for column c, source rows + index into ent
  std::vector<int32_t> trow, tsrc;
};

void build_graph(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, bool conv, Graph* g) {
  const int32_t n = A.n;
  g->n = n;
  g->ptr.assign(n + 1, 0);
  for (int32_t r = 0; r < n; ++r) {
    int32_t i = A.rowptr[r], ie = A.rowptr[r + 1];
    int32_t j = conv ? B1.rowptr[r] : 0, je = conv ? B1.rowptr[r + 1] : 0;
    int32_t k = conv ? B2.rowptr[r] : 0, ke = conv ? B2.rowptr[r + 1] : 0;
    while (i < ie || j < je || k < ke) {
      int32_t ca = i < ie ? A.col[i] : INT32_MAX, c1 = j < je ? B1.col[j] : INT32_MAX,
              c2 = k < ke ? B2.col[k] : INT32_MAX;
      int32_t c = std::min(ca, std::min(c1, c2));
      UEnt e{c, 0.f, 0.f, 0.f};
      if (ca == c) e.a = A.val[i++];
      if (c1 == c) e.b1 = B1.val[j++];
      if (c2 == c) e.b2 = B2.val[k++];
      g->ent.push_back(e);
    }
    g->ptr[r + 1] = (int32_t)g->ent.size();
  }
  g->tptr.assign(n + 1, 0);
  for (const UEnt& e : g->ent) g->tptr[e.col + 1]++;
  for (int32_t i = 0; i < n; ++i) g->tptr[i + 1] += g->tptr[i];
  g->trow.resize(g->ent.size());
  g->tsrc.resize(g->ent.size());
  std::vector<int32_t> cur(g->tptr.begin(), g->tptr.end() - 1);
  for (int32_t r = 0; r < n; ++r)
    for (int32_t k = g->ptr[r]; k < g->ptr[r + 1]; ++k) {
      int32_t p = cur[g->ent[k].col]++;
      g->trow[p] = r;
      g->tsrc[p] = k;
    }
}

struct Units {
  std::vector<int32_t> unit_of, first, mate;  // per dof unit id; per unit first dof; per dof partner (-1)
  int32_t rows(int32_t u, const std::vector<int32_t>& kind, int32_t out[2]) const {
    int32_t r = first[u];
    out[0] = r;
    if (kind[r] == 1) {
      out[1] = mate[r];
      return 2;
    }
    return 1;
  }
};

}  // namespace

int build_tile_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
                    const int32_t* idx_j, int32_t ns_branch, bool backward, int32_t max_lines, int32_t max_pairs,
                    TilePlan* out) {
  const int32_t n = A.n;
  TilePlan& T = *out;
  T = TilePlan();
  T.backward = backward;
  const bool conv = B1.present() && B2.present() && n_u > 0;
  T.has_conv = conv;
  std::vector<int32_t> pi(n, -1), pj(n, -1), kind(n, 0);
  Units U;
  U.mate.assign(n, -1);
  if (conv)
    for (int32_t k = 0; k < n_u; ++k) {
      int32_t i = idx_i[k], j = idx_j[k];
      if (i < 0 || i >= n || j < 0 || j >= n) return fail(FEO_ERR_INVALID_ARGUMENT, "idx_sol entry out of range");
      if (i == j || kind[i] != 0 || kind[j] != 0)
        return fail(FEO_ERR_UNSUPPORTED, "idx_sol[0]/idx_sol[1] must be duplicate-free and disjoint");
      kind[i] = 1;
      kind[j] = 2;
      pi[i] = pi[j] = i;
      pj[i] = pj[j] = j;
      U.mate[i] = j;
      U.mate[j] = i;
    }
  U.unit_of.assign(n, -1);
  for (int32_t r = 0; r < n; ++r) {
    if (U.unit_of[r] >= 0) continue;
    int32_t u = (int32_t)U.first.size();
    if (kind[r] == 0) {
      U.first.push_back(r);
      U.unit_of[r] = u;
    } else {
      U.first.push_back(pi[r]);
      U.unit_of[pi[r]] = U.unit_of[pj[r]] = u;
    }
  }
  const int32_t n_units = (int32_t)U.first.size();
  Graph G;
  build_graph(A, B1, B2, conv, &G);
  const float sgn = ns_branch ? 1.0f : -1.0f;

  // lines a dof row/column needs: forward = union columns (+ its partners); backward = source rows of
  // the transposed pattern as r-lines (tag 0) and, for convective entries, alpha[pi], alpha[pj] (tag 1)
  auto is_conv_entry = [&](int32_t h, const UEnt& e) { return conv && kind[h] != 0 && (e.b1 != 0.f || e.b2 != 0.f); };
  auto for_each_line = [&](int32_t d, auto&& fn) {
    if (!backward) {
      for (int32_t k = G.ptr[d]; k < G.ptr[d + 1]; ++k) fn(2 * (int64_t)G.ent[k].col);
      if (kind[d] != 0) {
        fn(2 * (int64_t)pi[d]);
        fn(2 * (int64_t)pj[d]);
      }
    } else {
      for (int32_t p = G.tptr[d]; p < G.tptr[d + 1]; ++p) {
        int32_t h = G.trow[p];
        const UEnt& e = G.ent[G.tsrc[p]];
        bool cv = is_conv_entry(h, e);
        if (!cv && e.a == 0.f) continue;
        fn(2 * (int64_t)h);
        if (cv) {
          fn(2 * (int64_t)pi[h] + 1);
          fn(2 * (int64_t)pj[h] + 1);
        }
      }
    }
  };

  // ---- grow tiles ------------------------------------------------------------------------------
  std::vector<char> seen(n_units, 0);
  std::vector<int32_t> stamp(2 * (size_t)n, -1);
  std::deque<int32_t> frontier, q;
  int32_t next_unseen = 0, placed = 0;
  std::vector<std::vector<int32_t>> tile_units;
  while (placed < n_units) {
    const int32_t tid = (int32_t)tile_units.size();
    tile_units.emplace_back();
    int32_t lines = 0, rows_in = 0;
    q.clear();
    while (true) {
      if (q.empty()) {
        int32_t seed = -1;
        while (!frontier.empty() && seed < 0) {
          int32_t c = frontier.front();
          frontier.pop_front();
          if (!seen[c]) seed = c;
        }
        if (seed < 0) {
          while (next_unseen < n_units && seen[next_unseen]) ++next_unseen;
          if (next_unseen >= n_units) break;
          seed = next_unseen;
        }
        seen[seed] = 1;
        q.push_back(seed);
      }
      int32_t u = q.front();
      int32_t rr[2];
      int32_t nr = U.rows(u, kind, rr);
      int32_t add = 0;
      for (int32_t t = 0; t < nr; ++t)
        for_each_line(rr[t], [&](int64_t key) {
          if (stamp[key] != tid) {
            stamp[key] = tid;
            ++add;
          }
        });
      // (stamps of a unit that does not fit are harmless: the tile is closed and the next tile has a new id)
      if (rows_in > 0 && (lines + add > max_lines || rows_in + nr > 2 * max_pairs)) break;
      if (add > max_lines) return fail(FEO_ERR_UNSUPPORTED, "a row needs more dof lines than a tile can stage; use the dense path");
      q.pop_front();
      tile_units.back().push_back(u);
      ++placed;
      lines += add;
      rows_in += nr;
      for (int32_t t = 0; t < nr; ++t) {
        if (!backward) {
          for (int32_t k = G.ptr[rr[t]]; k < G.ptr[rr[t] + 1]; ++k) {
            int32_t v = U.unit_of[G.ent[k].col];
            if (!seen[v]) {
              seen[v] = 1;
              q.push_back(v);
            }
          }
        } else {
          for (int32_t p = G.tptr[rr[t]]; p < G.tptr[rr[t] + 1]; ++p) {
            int32_t v = U.unit_of[G.trow[p]];
            if (!seen[v]) {
              seen[v] = 1;
              q.push_back(v);
            }
          }
        }
      }
    }
    for (int32_t v : q) {
      seen[v] = 0;
      frontier.push_back(v);
    }
    if (tile_units.back().empty()) tile_units.pop_back();
  }

  // ---- emit tiles ------------------------------------------------------------------------------
  T.tile_line_ptr.assign(1, 0);
  T.tile_pair_ptr.assign(1, 0);
  T.pair_step_ptr.assign(1, 0);
  T.pair_stepA_ptr.assign(1, 0);
  std::unordered_map<int64_t, int32_t> lmap;
  std::vector<int64_t> keys;
  for (const auto& units : tile_units) {
    // 1. the tile's lines, sorted by (array, dof) so the bulk copies walk memory in address order
    keys.clear();
    lmap.clear();
    for (int32_t u : units) {
      int32_t rr[2];
      int32_t nr = U.rows(u, kind, rr);
      for (int32_t t = 0; t < nr; ++t) for_each_line(rr[t], [&](int64_t key) { if (lmap.emplace(key, 0).second) keys.push_back(key); });
    }
    std::sort(keys.begin(), keys.end(), [](int64_t x, int64_t y) { return (x & 1) != (y & 1) ? (x & 1) < (y & 1) : x < y; });
    for (size_t i = 0; i < keys.size(); ++i) {
      lmap[keys[i]] = (int32_t)i;
      T.line_dof.push_back((int32_t)(keys[i] >> 1));
      T.line_src.push_back((int32_t)(keys[i] & 1));
    }
    T.max_lines = std::max<int32_t>(T.max_lines, (int32_t)keys.size());
    T.tile_line_ptr.push_back((int32_t)T.line_dof.size());
    auto L = [&](int32_t dof, int32_t tag) { return lmap.at(2 * (int64_t)dof + tag); };

    // 2. pairs: velocity pairs as they are, single dofs two by two
    std::vector<std::pair<int32_t, int32_t>> pairs;
    int32_t pending = -1;
    for (int32_t u : units) {
      int32_t rr[2];
      if (U.rows(u, kind, rr) == 2) {
        pairs.emplace_back(rr[0], rr[1]);
      } else if (pending < 0) {
        pending = rr[0];
      } else {
        pairs.emplace_back(pending, rr[0]);
        pending = -1;
      }
    }
    if (pending >= 0) pairs.emplace_back(pending, -1);

    // 3. steps
    struct Item {
      int32_t unit, dof;  // unit + dof of the "other side" (column in fwd, source row in bwd)
      UEnt e;
    };
    for (auto [da, db] : pairs) {
      T.pair_a.push_back(da);
      T.pair_b.push_back(db);
      const bool vel = kind[da] != 0;
      T.pair_li.push_back(vel && !backward ? L(pi[da], 0) : -1);
      T.pair_lj.push_back(vel && !backward ? L(pj[da], 0) : -1);
      T.pair_vel.push_back(vel ? 1 : 0);
      std::vector<Item> it[2];
      const int32_t dofs[2] = {da, db};
      for (int h = 0; h < 2; ++h) {
        int32_t d = dofs[h];
        if (d < 0) continue;
        if (!backward) {
          for (int32_t k = G.ptr[d]; k < G.ptr[d + 1]; ++k) it[h].push_back(Item{U.unit_of[G.ent[k].col], G.ent[k].col, G.ent[k]});
        } else {
          for (int32_t p = G.tptr[d]; p < G.tptr[d + 1]; ++p) {
            const UEnt& e = G.ent[G.tsrc[p]];
            int32_t hrow = G.trow[p];
            if (!is_conv_entry(hrow, e) && e.a == 0.f) continue;
            it[h].push_back(Item{U.unit_of[hrow], hrow, e});
          }
        }
        std::stable_sort(it[h].begin(), it[h].end(), [](const Item& x, const Item& y) { return x.unit != y.unit ? x.unit < y.unit : x.dof < y.dof; });
      }
      // merge by unit; in the backward, convective (3-gather) entries and plain entries go to separate lists
      for (int pass = 0; pass < (backward ? 2 : 1); ++pass) {
        size_t ia = 0, ib = 0;
        int32_t nsteps = 0;
        auto want = [&](const Item& x) { return !backward || (pass == 0) == is_conv_entry(x.dof, x.e); };
        auto emit = [&](const Item* xa, const Item* xb) {
          const Item* xs[2] = {xa, xb};
          for (int h = 0; h < 2; ++h) {
            const Item* x = xs[h] ? xs[h] : xs[1 - h];  // padding half re-reads the other half's line(s): a broadcast
            const bool real = xs[h] != nullptr;
            if (!backward) {
              T.steps_f.push_back(FwdEntry{L(x->dof, 0), real ? x->e.a : 0.f, real ? x->e.b1 : 0.f, real ? x->e.b2 : 0.f});
            } else if (pass == 0) {
              T.steps_b.push_back(BwdEntryB{L(x->dof, 0), L(pi[x->dof], 1), L(pj[x->dof], 1), 0, real ? x->e.a : 0.f,
                                            real ? sgn * x->e.b1 : 0.f, real ? sgn * x->e.b2 : 0.f, 0.f});
            } else {
              T.steps_a.push_back(BwdEntryA{L(x->dof, 0), real ? x->e.a : 0.f});
            }
          }
          ++nsteps;
        };
        while (true) {
          while (ia < it[0].size() && !want(it[0][ia])) ++ia;
          while (ib < it[1].size() && !want(it[1][ib])) ++ib;
          const bool ha = ia < it[0].size(), hb = ib < it[1].size();
          if (!ha && !hb) break;
          if (ha && hb && it[0][ia].unit == it[1][ib].unit) {
            emit(&it[0][ia], &it[1][ib]);
            ++ia;
            ++ib;
          } else if (ha && (!hb || it[0][ia].unit < it[1][ib].unit)) {
            emit(&it[0][ia], nullptr);
            ++ia;
          } else {
            emit(nullptr, &it[1][ib]);
            ++ib;
          }
        }
        // pad to the kernel's batch size by repeating the last step with zero coefficients
        const int32_t batch = !backward ? kTileBatchF : (pass == 0 ? kTileBatchB : kTileBatchA);
        for (; nsteps % batch != 0; ++nsteps) {
          if (!backward) {
            FwdEntry za = T.steps_f[T.steps_f.size() - 2], zb = T.steps_f[T.steps_f.size() - 1];
            za.a = za.b1 = za.b2 = zb.a = zb.b1 = zb.b2 = 0.f;
            T.steps_f.push_back(za);
            T.steps_f.push_back(zb);
          } else if (pass == 0) {
            BwdEntryB za = T.steps_b[T.steps_b.size() - 2], zb = T.steps_b[T.steps_b.size() - 1];
            za.a = za.b1s = za.b2s = zb.a = zb.b1s = zb.b2s = 0.f;
            T.steps_b.push_back(za);
            T.steps_b.push_back(zb);
          } else {
            BwdEntryA za = T.steps_a[T.steps_a.size() - 2], zb = T.steps_a[T.steps_a.size() - 1];
            za.a = zb.a = 0.f;
            T.steps_a.push_back(za);
            T.steps_a.push_back(zb);
          }
        }
        if (!backward)
          T.pair_step_ptr.push_back((int32_t)(T.steps_f.size() / 2));
        else if (pass == 0)
          T.pair_step_ptr.push_back((int32_t)(T.steps_b.size() / 2));
        else
          T.pair_stepA_ptr.push_back((int32_t)(T.steps_a.size() / 2));
      }
    }
    T.max_pairs = std::max<int32_t>(T.max_pairs, (int32_t)pairs.size());
    T.tile_pair_ptr.push_back((int32_t)T.pair_a.size());
  }
  T.n_tiles = (int32_t)T.tile_line_ptr.size() - 1;
  return FEO_OK;
}

}  // namespace feo
