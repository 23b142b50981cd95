// Host-side set-up: canonical CSR, transposes, the fused union pattern and the locality "blobs"
// the residual kernels walk.  Pure host code (no CUDA calls) so it can be checked on a CPU box.
//
// Semantics follow the reference's formula, not the PDE (SURVEY.md section 8a "quirks"):
//   * entries are kept iff value != 0 after the fp32 cast (quirk 10);
//   * (idx_i[k], idx_j[k]) are paired positionally and treated as opaque ints (quirk 4);
//   * Dirichlet identity rows inside A/B1/B2 are ordinary CSR rows here (quirk 3).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <queue>

#include "feo_internal.h"

namespace feo {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
const std::string& last_error() { return g_last_error; }

PlanTuning tuning_from_env() {
  PlanTuning t;
  if (const char* s = std::getenv("FEO_BLOB_ROWS")) t.blob_rows = std::max(2, atoi(s));
  if (const char* s = std::getenv("FEO_BLOB_MAX_ENT")) t.blob_max_ent = std::max(64, atoi(s));
  return t;
}

int canonicalize(const feo_csr& in, int32_t n, const char* name, HostCsr* out) {
  out->n = n;
  out->rowptr.clear();
  out->col.clear();
  out->val.clear();
  if (in.rowptr == nullptr) return FEO_OK;
  if (in.col == nullptr || in.val == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(name) + ": col/val NULL");
  if (in.rowptr[0] != 0) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(name) + ": rowptr[0] != 0");
  out->rowptr.assign(n + 1, 0);
  std::vector<std::pair<int32_t, float>> row;
  for (int32_t r = 0; r < n; ++r) {
    int32_t b = in.rowptr[r], e = in.rowptr[r + 1];
    if (e < b) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(name) + ": rowptr not monotone");
    row.clear();
    for (int32_t k = b; k < e; ++k) {
      int32_t c = in.col[k];
      if (c < 0 || c >= n) return fail(FEO_ERR_INVALID_ARGUMENT, std::string(name) + ": column index out of range");
      row.emplace_back(c, in.val[k]);
    }
    std::stable_sort(row.begin(), row.end(), [](auto& x, auto& y) { return x.first < y.first; });
    for (size_t k = 0; k < row.size();) {
      int32_t c = row[k].first;
      float v = 0.f;
      while (k < row.size() && row[k].first == c) v += row[k++].second;
      if (v != 0.0f) {  // threshold 0, not an epsilon (quirk 10)
        out->col.push_back(c);
        out->val.push_back(v);
      }
    }
    out->rowptr[r + 1] = (int32_t)out->col.size();
  }
  return FEO_OK;
}

HostCsr transpose(const HostCsr& a) {
  HostCsr t;
  if (!a.present()) return t;
  t.n = a.n;
  t.rowptr.assign(a.n + 1, 0);
  for (int32_t c : a.col) t.rowptr[c + 1]++;
  for (int32_t i = 0; i < a.n; ++i) t.rowptr[i + 1] += t.rowptr[i];
  t.col.resize(a.col.size());
  t.val.resize(a.val.size());
  std::vector<int32_t> cur(t.rowptr.begin(), t.rowptr.end() - 1);
  for (int32_t r = 0; r < a.n; ++r)
    for (int32_t k = a.rowptr[r]; k < a.rowptr[r + 1]; ++k) {
      int32_t p = cur[a.col[k]]++;
      t.col[p] = r;
      t.val[p] = a.val[k];
    }
  return t;
}

HostCsr axpy(const HostCsr& s, float dt, const HostCsr& a) {
  // M = S + dt*A evaluated in fp32 like `S_mat + dt * A_mat` (FEONet_time_dep_Stokes/train_FEONet.py:345)
  HostCsr m;
  m.n = s.n;
  m.rowptr.assign(s.n + 1, 0);
  for (int32_t r = 0; r < s.n; ++r) {
    int32_t i = s.rowptr[r], ie = s.rowptr[r + 1], j = a.rowptr[r], je = a.rowptr[r + 1];
    while (i < ie || j < je) {
      int32_t cs = i < ie ? s.col[i] : INT32_MAX, ca = j < je ? a.col[j] : INT32_MAX;
      int32_t c = std::min(cs, ca);
      float v = 0.f;
      if (cs == c) v = s.val[i++];
      if (ca == c) v = v + dt * a.val[j++];
      if (v != 0.0f) {
        m.col.push_back(c);
        m.val.push_back(v);
      }
    }
    m.rowptr[r + 1] = (int32_t)m.col.size();
  }
  return m;
}

namespace {
struct UEnt {
  int32_t col;
  float a, b1, b2;
};

// union of row r of A, B1, B2 (each sorted by column)
void union_row(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t r, std::vector<UEnt>* out) {
  out->clear();
  int32_t i = A.rowptr[r], ie = A.rowptr[r + 1];
  int32_t j = B1.present() ? B1.rowptr[r] : 0, je = B1.present() ? B1.rowptr[r + 1] : 0;
  int32_t k = B2.present() ? B2.rowptr[r] : 0, ke = B2.present() ? B2.rowptr[r + 1] : 0;
  while (i < ie || j < je || k < ke) {
    int32_t ca = i < ie ? A.col[i] : INT32_MAX, c1 = j < je ? B1.col[j] : INT32_MAX,
            c2 = k < ke ? B2.col[k] : INT32_MAX;
    int32_t c = std::min(ca, std::min(c1, c2));
    UEnt e{c, 0.f, 0.f, 0.f};
    if (ca == c) e.a = A.val[i++];
    if (c1 == c) e.b1 = B1.val[j++];
    if (c2 == c) e.b2 = B2.val[k++];
    out->push_back(e);
  }
}
}  // namespace

int build_plan(const HostCsr& A, const HostCsr& B1, const HostCsr& B2, int32_t n_u, const int32_t* idx_i,
               const int32_t* idx_j, int32_t ns_branch, const PlanTuning& tune, HostPlan* plan) {
  const int32_t n = A.n;
  HostPlan& P = *plan;
  P = HostPlan();
  P.n = n;
  P.n_u = n_u;
  P.ns_branch = ns_branch;
  P.has_conv = B1.present() && B2.present() && n_u > 0;
  if (B1.present() != B2.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "B1 and B2 must be given together");
  if (P.has_conv && (idx_i == nullptr || idx_j == nullptr))
    return fail(FEO_ERR_INVALID_ARGUMENT, "convection needs idx_i/idx_j");

  // partner lookups from the opaque index lists
  P.pi.assign(n, -1);
  P.pj.assign(n, -1);
  P.kind.assign(n, 0);
  std::vector<int32_t> mate(n, -1);
  if (P.has_conv) {
    for (int32_t k = 0; k < n_u; ++k) {
      int32_t i = idx_i[k], j = idx_j[k];
      if (i < 0 || i >= n || j < 0 || j >= n) return fail(FEO_ERR_INVALID_ARGUMENT, "idx_sol entry out of range");
      if (i == j || P.kind[i] != 0 || P.kind[j] != 0)
        return fail(FEO_ERR_UNSUPPORTED, "idx_sol[0]/idx_sol[1] must be duplicate-free and disjoint");
      P.kind[i] = 1;
      P.kind[j] = 2;
      P.pi[i] = P.pi[j] = i;
      P.pj[i] = P.pj[j] = j;
      mate[i] = j;
      mate[j] = i;
    }
  }

  // units: a velocity pair (I[k] then J[k]) or a single dof
  std::vector<int32_t> unit_of(n, -1), unit_first;  // unit -> first dof (I[k] for pairs)
  for (int32_t r = 0; r < n; ++r) {
    if (unit_of[r] >= 0) continue;
    int32_t u = (int32_t)unit_first.size();
    if (P.kind[r] == 0) {
      unit_first.push_back(r);
      unit_of[r] = u;
    } else {
      int32_t i = P.pi[r], j = P.pj[r];
      unit_first.push_back(i);
      unit_of[i] = unit_of[j] = u;
    }
  }
  const int32_t n_units = (int32_t)unit_first.size();
  auto unit_rows = [&](int32_t u, int32_t out[2]) {
    int32_t r = unit_first[u];
    out[0] = r;
    if (P.kind[r] == 1) {
      out[1] = mate[r];
      return 2;
    }
    return 1;
  };

  // union rows (cached: needed by blob growth and by the entry streams)
  std::vector<int32_t> uptr(n + 1, 0);
  std::vector<UEnt> uent;
  {
    std::vector<UEnt> row;
    const HostCsr none;
    for (int32_t r = 0; r < n; ++r) {
      union_row(A, P.has_conv ? B1 : none, P.has_conv ? B2 : none, r, &row);
      uent.insert(uent.end(), row.begin(), row.end());
      uptr[r + 1] = (int32_t)uent.size();
      P.max_row_nnz = std::max<int32_t>(P.max_row_nnz, (int32_t)row.size());
    }
    P.nnz_union = (int64_t)uent.size();
  }
  if (2 * (int64_t)(P.max_row_nnz + kPadF) > tune.blob_max_ent)
    return fail(FEO_ERR_UNSUPPORTED, "a row is too dense for the sparse path; use the dense operator path");

  // blobs: grow compact neighbourhoods by BFS over the union pattern so that the rows a CTA walks
  // share most of their columns (L1/L2 reuse); next seeds come from the previous frontier.
  std::vector<char> seen(n_units, 0);
  std::vector<int32_t> order;
  order.reserve(n_units);
  P.blob_uptr.push_back(0);
  std::deque<int32_t> frontier;  // candidate seeds
  int32_t next_unseen = 0;
  std::deque<int32_t> q;
  while ((int32_t)order.size() < n_units) {
    int32_t rows_in_blob = 0;
    int64_t ent_in_blob = 0;
    q.clear();
    bool full = false;
    while (!full) {
      if (q.empty()) {  // (re)seed: oldest frontier candidate, else first unseen dof in index order
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
      int32_t nr = unit_rows(u, rr);
      int64_t ue = 0;
      for (int32_t t = 0; t < nr; ++t) ue += (uptr[rr[t] + 1] - uptr[rr[t]] + kPadF - 1) / kPadF * kPadF;
      if (rows_in_blob > 0 && (rows_in_blob + nr > tune.blob_rows || ent_in_blob + ue > tune.blob_max_ent)) {
        full = true;
        break;
      }
      q.pop_front();
      order.push_back(u);
      rows_in_blob += nr;
      ent_in_blob += ue;
      for (int32_t t = 0; t < nr; ++t)
        for (int32_t k = uptr[rr[t]]; k < uptr[rr[t] + 1]; ++k) {
          int32_t v = unit_of[uent[k].col];
          if (!seen[v]) {
            seen[v] = 1;
            q.push_back(v);
          }
        }
    }
    // whatever is left in the queue was marked seen but not placed: hand it to the frontier
    for (int32_t v : q) {
      seen[v] = 0;
      frontier.push_back(v);
    }
    if ((int32_t)order.size() > P.blob_uptr.back()) P.blob_uptr.push_back((int32_t)order.size());
  }

  // slots in walk order
  P.unit_ptr.assign(1, 0);
  for (int32_t u : order) {
    int32_t rr[2];
    int32_t nr = unit_rows(u, rr);
    for (int32_t t = 0; t < nr; ++t) P.slot_row.push_back(rr[t]);
    P.unit_ptr.push_back((int32_t)P.slot_row.size());
  }
  const int32_t n_slots = (int32_t)P.slot_row.size();
  if (n_slots != n) return fail(FEO_ERR_INVALID_ARGUMENT, "internal: slot count != n");

  // forward stream; every row is padded to a multiple of kPadF with zero-valued entries that repeat
  // the last column (a harmless L1-hit gather) so the kernels run fixed-size load batches
  P.fptr.assign(1, 0);
  for (int32_t s = 0; s < n_slots; ++s) {
    int32_t r = P.slot_row[s];
    int32_t cnt = 0, last = 0;
    for (int32_t k = uptr[r]; k < uptr[r + 1]; ++k, ++cnt) {
      const UEnt& e = uent[k];
      last = e.col;
      if (P.has_conv)
        P.fent.push_back(FwdEntry{e.col, e.a, e.b1, e.b2});
      else
        P.fent_lin.push_back(FwdEntryLin{e.col, e.a});
    }
    for (; cnt % kPadF != 0; ++cnt) {
      if (P.has_conv)
        P.fent.push_back(FwdEntry{last, 0.f, 0.f, 0.f});
      else
        P.fent_lin.push_back(FwdEntryLin{last, 0.f});
    }
    P.fptr.push_back(P.has_conv ? (int32_t)P.fent.size() : (int32_t)P.fent_lin.size());
  }

  // backward stream: column-owned lists from the transposed union pattern
  {
    std::vector<int32_t> tptr(n + 1, 0);
    for (const UEnt& e : uent) tptr[e.col + 1]++;
    for (int32_t i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
    std::vector<int32_t> trow(uent.size());
    std::vector<int32_t> tsrc(uent.size());
    std::vector<int32_t> cur(tptr.begin(), tptr.end() - 1);
    for (int32_t r = 0; r < n; ++r)
      for (int32_t k = uptr[r]; k < uptr[r + 1]; ++k) {
        int32_t p = cur[uent[k].col]++;
        trow[p] = r;
        tsrc[p] = k;
      }
    const float s = ns_branch ? 1.0f : -1.0f;
    P.bptrA.assign(1, 0);
    P.bptrB.assign(1, 0);
    for (int32_t sl = 0; sl < n_slots; ++sl) {
      int32_t c = P.slot_row[sl];
      int32_t cntA = 0, cntB = 0;
      for (int32_t p = tptr[c]; p < tptr[c + 1]; ++p) {
        int32_t h = trow[p];
        const UEnt& e = uent[tsrc[p]];
        bool conv = P.has_conv && P.kind[h] != 0 && (e.b1 != 0.f || e.b2 != 0.f);
        if (conv) {
          P.bentB.push_back(BwdEntryB{h, P.pi[h], P.pj[h], 0, e.a, s * e.b1, s * e.b2, 0.f});
          ++cntB;
        } else if (e.a != 0.f) {
          P.bentA.push_back(BwdEntryA{h, e.a});
          ++cntA;
        }
      }
      P.n_bent_real += cntA + cntB;
      for (; cntA % kPadBA != 0; ++cntA) P.bentA.push_back(BwdEntryA{P.bentA.back().row, 0.f});
      for (; cntB % kPadBB != 0; ++cntB) {
        BwdEntryB z = P.bentB.back();
        z.a = z.b1s = z.b2s = 0.f;
        P.bentB.push_back(z);
      }
      P.bptrA.push_back((int32_t)P.bentA.size());
      P.bptrB.push_back((int32_t)P.bentB.size());
    }
  }

  // per-blob staging sizes
  for (size_t b = 0; b + 1 < P.blob_uptr.size(); ++b) {
    int32_t s0 = P.unit_ptr[P.blob_uptr[b]], s1 = P.unit_ptr[P.blob_uptr[b + 1]];
    P.max_blob_fent = std::max(P.max_blob_fent, P.fptr[s1] - P.fptr[s0]);
    P.max_blob_bentA = std::max(P.max_blob_bentA, P.bptrA[s1] - P.bptrA[s0]);
    P.max_blob_bentB = std::max(P.max_blob_bentB, P.bptrB[s1] - P.bptrB[s0]);
  }
  return FEO_OK;
}

}  // namespace feo
