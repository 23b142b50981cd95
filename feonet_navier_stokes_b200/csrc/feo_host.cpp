// Host-side CSR utilities and error plumbing.  Pure host code (no CUDA calls).
//
// Semantics follow the reference's formula, not the PDE (SURVEY.md section 8a "quirks"):
//   * entries are kept iff value != 0 after the fp32 cast (quirk 10);
//   * Dirichlet identity rows inside A/B1/B2 are ordinary CSR rows here (quirk 3).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "feo_internal.h"

namespace feo {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
const std::string& last_error() { return g_last_error; }

int canonicalize(const feo_csr& in, int32_t n, const char* name, HostCsr* out) {
  out->n = n;
  out->rowptr.clear();
  out->col.clear();
  out->val.clear();
  if (in.rowptr == nullptr) return FEO_OK;
  if ((in.col == nullptr || in.val == nullptr) && in.rowptr[n] != 0)  // an all-zero matrix may come with empty arrays
    return fail(FEO_ERR_INVALID_ARGUMENT, std::string(name) + ": col/val NULL");
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

SeqPlan build_seq_plan(const HostCsr& m, const HostCsr& s) {
  // One word per column of the union pattern, ascending: the kernel adds the M terms and the S terms of a row in the order
  // of their own CSR rows (an absent coefficient is an exact fma with 0), as two separate spmm passes would.
  SeqPlan p;
  p.rowptr.assign(m.n + 1, 0);
  for (int32_t r = 0; r < m.n; ++r) {
    int32_t i = m.rowptr[r], ie = m.rowptr[r + 1], j = s.rowptr[r], je = s.rowptr[r + 1];
    while (i < ie || j < je) {
      const int32_t cm = i < ie ? m.col[i] : INT32_MAX, cs = j < je ? s.col[j] : INT32_MAX;
      SeqEnt e{std::min(cm, cs), 0.f, 0.f, 0};
      if (cm == e.col) e.m = m.val[i++];
      if (cs == e.col) e.s = s.val[j++];
      p.ent.push_back(e);
    }
    p.rowptr[r + 1] = (int32_t)p.ent.size();
  }
  return p;
}

}  // namespace feo
