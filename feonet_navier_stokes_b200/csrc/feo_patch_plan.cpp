// Host-side patch plan of the fused residual kernels (see feo_patch.h for the decomposition and the stream format).
// Pure host code; the CPU test-suite checks it through replay_patch_plan (feo_debug_patch_replay), which simulates the
// line pool and decodes the streams exactly as feo_patch.cu does.
#include "feo_patch.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <limits>
#include <map>

namespace feo {

PatchTuning patch_tuning_from_env(bool backward) {
  PatchTuning t;
  // The register file is split over the four SM sub-partitions, so the warp count (consumers + producers) should be a
  // multiple of 4: 20 warps -> 96 registers (forward), 12 warps -> 168 registers (backward: the pool holds the lines of about 10 patches next to
  // their predecessors').  A round holds one patch per consumer warp.
  t.warps = backward ? 11 : 19;
  t.producers = backward ? 5 : 1;  // backward: 11 + 5 = 16 warps (128 registers); its rounds fetch ~280 lines each
  if (const char* s = std::getenv(backward ? "FEO_PATCH_WARPS_BWD" : "FEO_PATCH_WARPS_FWD")) t.warps = atoi(s);
  if (const char* s = std::getenv(backward ? "FEO_PATCH_PRODUCERS_BWD" : "FEO_PATCH_PRODUCERS_FWD")) t.producers = atoi(s);
  if (const char* s = std::getenv("FEO_PATCH_SEG_ROUNDS")) t.seg_rounds = atoi(s);
  if (const char* s = std::getenv("FEO_PATCH_MERGE_PAD")) t.merge_pad = std::max(atoi(s), 0);
  if (const char* s = std::getenv("FEO_PATCH_ROUND_FILL")) t.round_fill_pct = std::min(std::max(atoi(s), 10), 100);
  if (const char* s = std::getenv(backward ? "FEO_PATCH_STREAM_CAP_BWD" : "FEO_PATCH_STREAM_CAP_FWD")) t.stream_cap = atoi(s);
  if (const char* s = std::getenv(backward ? "FEO_PATCH_POOL_BWD" : "FEO_PATCH_POOL_FWD")) t.pool_lines = atoi(s);
  t.warps = std::min(std::max(t.warps, 1), 23);
  t.producers = std::min(std::max(t.producers, 1), 8);
  t.seg_rounds = std::max(t.seg_rounds, 2);
  if (t.stream_cap <= 0) t.stream_cap = t.warps * (backward ? 2048 : 1536) + 4096;  // + the load list of the next round
  t.stream_cap = (t.stream_cap + 1023) / 1024 * 1024;
  const int32_t budget = (232448 - 2 * t.stream_cap - 1024) / kLineBytes;
  if (t.pool_lines <= 0 || t.pool_lines > budget) t.pool_lines = budget;
  t.pool_lines = std::min(t.pool_lines, 1023);  // slots are 10-bit fields of 16-bit line references
  return t;
}

namespace {

// lines one round may hold (its successor needs room in the pool for the lines the two do not share)
inline int32_t round_lines_cap(const PatchTuning& t) { return std::max<int32_t>(t.pool_lines * t.round_fill_pct / 100, 1); }

inline bool same_bits(float x, float y) { return f2u(x) == f2u(y); }

struct PairRel {  // node m feeds the target node with the same coefficients in both components
  int32_t m;
  float c[5];  // forward: a, b1, b2 ; backward: a, b1s, b2s, f1, f2
};
struct PlainRel {  // one line feeds (target I, target J) with plain coefficients
  int32_t dof;
  float aI, aJ;
};
struct NodeRels {
  std::vector<PairRel> pair;   // sorted by m
  std::vector<PlainRel> plain; // sorted by dof
};
struct SPair {
  int32_t m;
  float sI, sJ;
};
struct SPlain {
  int32_t dof;
  float s;
};
struct SingleRels {
  std::vector<SPair> pair;
  std::vector<SPlain> plain;
};

struct Topo {
  int32_t n = 0, n_nodes = 0;
  std::vector<int32_t> node_of;  // per dof, -1 for singles
  std::vector<int32_t> nI, nJ;   // dofs of a node
  std::vector<int32_t> singles;  // dofs with kind 0
  std::vector<int32_t> single_id;  // per dof
};

Topo make_topo(const Front& F) {
  Topo T;
  T.n = F.n;
  T.node_of.assign(F.n, -1);
  T.single_id.assign(F.n, -1);
  for (int32_t d = 0; d < F.n; ++d) {
    if (F.kind[d] == 1) {
      T.node_of[d] = T.node_of[F.mate[d]] = T.n_nodes++;
      T.nI.push_back(d);
      T.nJ.push_back(F.mate[d]);
    } else if (F.kind[d] == 0) {
      T.single_id[d] = (int32_t)T.singles.size();
      T.singles.push_back(d);
    }
  }
  return T;
}

struct ColAcc {
  bool hasI = false, hasJ = false, used = false;
  float I[3] = {0, 0, 0}, J[3] = {0, 0, 0};
};

// forward relations of node k (rows nI[k], nJ[k]); returns false (with why) when an entry does not fit the patch model
bool fwd_node_rels(const Front& F, const Topo& T, int32_t k, NodeRels* out, int64_t* real, std::string* why) {
  std::map<int32_t, ColAcc> cols;
  const int32_t rows[2] = {T.nI[k], T.nJ[k]};
  for (int side = 0; side < 2; ++side)
    for (int32_t p = F.ptr[rows[side]]; p < F.ptr[rows[side] + 1]; ++p) {
      const UEnt& e = F.ent[p];
      ColAcc& c = cols[e.col];
      float* dst = side == 0 ? c.I : c.J;
      (side == 0 ? c.hasI : c.hasJ) = true;
      dst[0] = e.a;
      dst[1] = e.b1;
      dst[2] = e.b2;
      ++*real;
    }
  for (auto& kv : cols) {
    ColAcc& c = kv.second;
    if (c.used || F.kind[kv.first] != 1 || !c.hasI || c.hasJ) continue;
    auto jt = cols.find(F.mate[kv.first]);
    if (jt == cols.end()) continue;
    ColAcc& d = jt->second;
    if (d.used || !d.hasJ || d.hasI) continue;
    if (!same_bits(c.I[0], d.J[0]) || !same_bits(c.I[1], d.J[1]) || !same_bits(c.I[2], d.J[2])) continue;
    out->pair.push_back(PairRel{T.node_of[kv.first], {c.I[0], c.I[1], c.I[2], 0.f, 0.f}});
    c.used = d.used = true;
  }
  for (auto& kv : cols) {
    const ColAcc& c = kv.second;
    if (c.used) continue;
    if (c.I[1] != 0.f || c.I[2] != 0.f || c.J[1] != 0.f || c.J[2] != 0.f) {
      *why = "a convective entry is not mirrored in the partner row (cross-component or asymmetric B1/B2)";
      return false;
    }
    if (c.I[0] != 0.f || c.J[0] != 0.f) out->plain.push_back(PlainRel{kv.first, c.I[0], c.J[0]});
  }
  std::sort(out->pair.begin(), out->pair.end(), [](const PairRel& x, const PairRel& y) { return x.m < y.m; });
  return true;
}

void fwd_single_rels(const Front& F, const Topo& T, int32_t s, SingleRels* out, int64_t* real) {
  std::map<int32_t, SPair> by_node;
  for (int32_t p = F.ptr[s]; p < F.ptr[s + 1]; ++p) {
    const UEnt& e = F.ent[p];
    if (e.a == 0.f) continue;  // B1/B2 stored on a non-velocity row have no effect on the residual
    ++*real;
    if (F.kind[e.col] == 0) {
      out->plain.push_back(SPlain{e.col, e.a});
    } else {
      SPair& sp = by_node.emplace(T.node_of[e.col], SPair{T.node_of[e.col], 0.f, 0.f}).first->second;
      (F.kind[e.col] == 1 ? sp.sI : sp.sJ) = e.a;
    }
  }
  for (auto& kv : by_node) {
    const SPair& sp = kv.second;
    if (sp.sI != 0.f && sp.sJ != 0.f)
      out->pair.push_back(sp);
    else
      out->plain.push_back(sp.sI != 0.f ? SPlain{T.nI[sp.m], sp.sI} : SPlain{T.nJ[sp.m], sp.sJ});
  }
  std::sort(out->plain.begin(), out->plain.end(), [](const SPlain& x, const SPlain& y) { return x.dof < y.dof; });
}

// backward relations of the column pair of node k
bool bwd_node_rels(const Front& F, const Topo& T, int32_t k, NodeRels* out, int64_t* real, std::string* why) {
  std::map<int32_t, ColAcc> src;  // by source row
  const int32_t cols[2] = {T.nI[k], T.nJ[k]};
  for (int side = 0; side < 2; ++side)
    for (int32_t p = F.tptr[cols[side]]; p < F.tptr[cols[side] + 1]; ++p) {
      const int32_t h = F.trow[p];
      const UEnt& e = F.ent[F.tsrc[p]];
      const bool cv = F.is_conv(h, e);
      if (!cv && e.a == 0.f) continue;
      ColAcc& c = src[h];
      float* dst = side == 0 ? c.I : c.J;
      (side == 0 ? c.hasI : c.hasJ) = true;
      dst[0] = e.a;
      dst[1] = cv ? F.sgn * e.b1 : 0.f;
      dst[2] = cv ? F.sgn * e.b2 : 0.f;
      ++*real;
    }
  std::map<int32_t, PairRel> pairs;
  for (auto& kv : src) {
    ColAcc& c = kv.second;
    if (c.used || F.kind[kv.first] != 1 || !c.hasI || c.hasJ) continue;
    auto jt = src.find(F.mate[kv.first]);
    if (jt == src.end()) continue;
    ColAcc& d = jt->second;
    if (d.used || !d.hasJ || d.hasI) continue;
    if (!same_bits(c.I[0], d.J[0]) || !same_bits(c.I[1], d.J[1]) || !same_bits(c.I[2], d.J[2])) continue;
    const int32_t m = T.node_of[kv.first];
    pairs[m] = PairRel{m, {c.I[0], c.I[1], c.I[2], 0.f, 0.f}};
    c.used = d.used = true;
  }
  for (auto& kv : src) {
    const ColAcc& c = kv.second;
    if (c.used) continue;
    if (c.I[1] != 0.f || c.I[2] != 0.f || c.J[1] != 0.f || c.J[2] != 0.f) {
      *why = "a convective entry is not mirrored in the partner column (cross-component or asymmetric B1/B2)";
      return false;
    }
    out->plain.push_back(PlainRel{kv.first, c.I[0], c.J[0]});
  }
  // E-term: the convective entries of the node's own rows, Bu1[cI] = sum B1[cI, I m] alpha[I m], Bu1[cJ] = sum B1[cJ, J m] alpha[J m]
  std::map<int32_t, std::pair<const UEnt*, const UEnt*>> fw;
  for (int side = 0; side < 2; ++side)
    for (int32_t p = F.ptr[cols[side]]; p < F.ptr[cols[side] + 1]; ++p) {
      const UEnt& e = F.ent[p];
      if (!F.is_conv(cols[side], e)) continue;
      if (F.kind[e.col] != (side == 0 ? 1 : 2)) {
        *why = "a convective entry couples different components (or a non-velocity column)";
        return false;
      }
      auto& slot = fw[T.node_of[e.col]];
      (side == 0 ? slot.first : slot.second) = &e;
    }
  for (auto& kv : fw) {
    const UEnt *eI = kv.second.first, *eJ = kv.second.second;
    if (eI == nullptr || eJ == nullptr || !same_bits(eI->b1, eJ->b1) || !same_bits(eI->b2, eJ->b2)) {
      *why = "the B1/B2 entries of a node's two rows differ";
      return false;
    }
    PairRel& pr = pairs.emplace(kv.first, PairRel{kv.first, {0.f, 0.f, 0.f, 0.f, 0.f}}).first->second;
    pr.c[3] = eI->b1;
    pr.c[4] = eI->b2;
  }
  for (auto& kv : pairs) out->pair.push_back(kv.second);
  return true;
}

bool bwd_single_rels(const Front& F, const Topo& T, int32_t q, SingleRels* out, int64_t* real, std::string* why) {
  std::map<int32_t, SPair> by_node;
  for (int32_t p = F.tptr[q]; p < F.tptr[q + 1]; ++p) {
    const int32_t h = F.trow[p];
    const UEnt& e = F.ent[F.tsrc[p]];
    if (F.is_conv(h, e)) {
      *why = "B1/B2 entry in a non-velocity column";
      return false;
    }
    if (e.a == 0.f) continue;
    ++*real;
    if (F.kind[h] == 0) {
      out->plain.push_back(SPlain{h, e.a});
    } else {
      SPair& sp = by_node.emplace(T.node_of[h], SPair{T.node_of[h], 0.f, 0.f}).first->second;
      (F.kind[h] == 1 ? sp.sI : sp.sJ) = e.a;
    }
  }
  for (auto& kv : by_node) {
    const SPair& sp = kv.second;
    if (sp.sI != 0.f && sp.sJ != 0.f)
      out->pair.push_back(sp);
    else
      out->plain.push_back(sp.sI != 0.f ? SPlain{T.nI[sp.m], sp.sI} : SPlain{T.nJ[sp.m], sp.sJ});
  }
  std::sort(out->plain.begin(), out->plain.end(), [](const SPlain& x, const SPlain& y) { return x.dof < y.dof; });
  return true;
}

struct Patch {
  int32_t nodes[kPatchNodes] = {-1, -1, -1, -1};
  int32_t n_nodes = 0;
  int32_t single = -1;  // dof
  // steps
  struct PairStep {
    int32_t m;
    uint32_t mask;
    float c[kPatchNodes][5];
    float sI, sJ;
  };
  struct PlainStep {
    int32_t dof;
    uint32_t mask;
    float aI[kPatchNodes], aJ[kPatchNodes];
    float s;
  };
  std::vector<PairStep> pair;
  std::vector<PlainStep> plain;
  std::vector<int64_t> keys;  // lines (src * n + dof), sorted unique
  int32_t units = 0;          // stream size in 16-byte words
};

inline int popc4(uint32_t m) { return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u)); }
// step sizes in 16-byte words (feo_patch.h)
inline int pair_words(bool backward, int k) { return backward ? (6 + 5 * k + 3) / 4 : (4 + 3 * k + 3) / 4; }
inline int plain_words(int k) { return (2 + 2 * k + 3) / 4; }

}  // namespace

int build_patch_plan(const Front& F, bool backward, const PatchTuning& tune, PatchPlan* out) {
  PatchPlan& P = *out;
  P = PatchPlan();
  P.backward = backward;
  P.n = F.n;
  P.warps = tune.warps;
  P.producers = tune.producers;
  P.pool_lines = tune.pool_lines;
  P.stream_cap = tune.stream_cap;
  if (!F.conv) {
    P.why_not = "operator has no convective term";
    return FEO_OK;
  }
  const Topo T = make_topo(F);
  const int32_t n = F.n, NN = T.n_nodes;

  // ---- relations ----------------------------------------------------------------------------------------------------
  std::vector<NodeRels> rel(NN);
  for (int32_t k = 0; k < NN; ++k)
    if (!(backward ? bwd_node_rels(F, T, k, &rel[k], &P.real_entries, &P.why_not) : fwd_node_rels(F, T, k, &rel[k], &P.real_entries, &P.why_not)))
      return FEO_OK;
  std::vector<SingleRels> srel(T.singles.size());
  for (size_t i = 0; i < T.singles.size(); ++i) {
    if (backward) {
      if (!bwd_single_rels(F, T, T.singles[i], &srel[i], &P.real_entries, &P.why_not)) return FEO_OK;
    } else {
      fwd_single_rels(F, T, T.singles[i], &srel[i], &P.real_entries);
    }
  }

  // ---- patches: a hub node + up to three nodes whose source nodes the hub needs anyway --------------------------------
  std::vector<int32_t> order(NN);
  for (int32_t k = 0; k < NN; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return rel[x].pair.size() > rel[y].pair.size(); });
  std::vector<int32_t> patch_of(NN, -1), stamp(NN, -1);
  std::vector<Patch> patches;
  std::vector<int32_t> small;  // single-node patches with next to no work, merged four by four below
  for (int32_t h : order) {
    if (patch_of[h] >= 0) continue;
    for (const PairRel& r : rel[h].pair) stamp[r.m] = h;
    struct Cand {
      int32_t m, inter, extra;
    };
    std::vector<Cand> cand;
    for (const PairRel& r : rel[h].pair) {
      const int32_t m = r.m;
      if (m == h || patch_of[m] >= 0) continue;
      int32_t inter = 0;
      for (const PairRel& q : rel[m].pair) inter += stamp[q.m] == h;
      const int32_t extra = (int32_t)rel[m].pair.size() - inter;
      if (extra <= inter / 4) cand.push_back(Cand{m, inter, extra});
    }
    std::stable_sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) {
      if (x.extra != y.extra) return x.extra < y.extra;
      return x.inter > y.inter;
    });
    if (cand.empty() && rel[h].pair.size() <= 2) {
      small.push_back(h);
      patch_of[h] = -2;
      continue;
    }
    Patch pt;
    pt.nodes[pt.n_nodes++] = h;
    for (const Cand& c : cand) {
      if (pt.n_nodes == kPatchNodes) break;
      pt.nodes[pt.n_nodes++] = c.m;
    }
    for (int t = 0; t < pt.n_nodes; ++t) patch_of[pt.nodes[t]] = (int32_t)patches.size();
    patches.push_back(std::move(pt));
  }
  for (size_t i = 0; i < small.size(); i += kPatchNodes) {
    Patch pt;
    for (size_t j = i; j < std::min(small.size(), i + kPatchNodes); ++j) {
      pt.nodes[pt.n_nodes++] = small[j];
      patch_of[small[j]] = (int32_t)patches.size();
    }
    patches.push_back(std::move(pt));
  }
  // singles join the patch that already gathers most of their source nodes
  {
    std::vector<int32_t> cnt(patches.size(), 0), touched;
    std::vector<std::vector<int32_t>> step_nodes;  // lazily: nodes a patch gathers
    std::vector<int32_t> mark(NN, -1);
    for (size_t i = 0; i < T.singles.size(); ++i) {
      const SingleRels& sr = srel[i];
      int32_t best = -1, best_cnt = 0;
      if (!sr.pair.empty()) {
        // candidate patches: those owning one of the source nodes
        touched.clear();
        for (const SPair& sp : sr.pair) {
          const int32_t pc = patch_of[sp.m];
          if (pc >= 0 && patches[pc].single < 0 && cnt[pc]++ == 0) touched.push_back(pc);
        }
        for (int32_t pc : touched) {
          // overlap of the single's source nodes with the nodes the patch gathers
          for (int t = 0; t < patches[pc].n_nodes; ++t)
            for (const PairRel& r : rel[patches[pc].nodes[t]].pair) mark[r.m] = pc;
          int32_t ov = 0;
          for (const SPair& sp : sr.pair) ov += mark[sp.m] == pc;
          if (ov > best_cnt || (ov == best_cnt && best >= 0 && pc < best)) {
            best = pc;
            best_cnt = ov;
          }
          for (int t = 0; t < patches[pc].n_nodes; ++t)
            for (const PairRel& r : rel[patches[pc].nodes[t]].pair) mark[r.m] = -1;
          cnt[pc] = 0;
        }
      }
      if (best >= 0 && 2 * best_cnt >= (int32_t)sr.pair.size()) {
        patches[best].single = T.singles[i];
      } else {
        Patch pt;
        pt.single = T.singles[i];
        patches.push_back(std::move(pt));
      }
    }
  }
  const int32_t NP = (int32_t)patches.size();
  P.n_patches = NP;
  P.n_nodes = NN;
  P.n_singles = (int64_t)T.singles.size();

  // ---- steps and lines of every patch ---------------------------------------------------------------------------------
  auto key_of = [&](int32_t dof, int32_t src) { return (int64_t)src * n + dof; };
  for (Patch& pt : patches) {
    std::map<int32_t, Patch::PairStep> ps;
    std::map<int32_t, Patch::PlainStep> pl;
    auto pair_step = [&](int32_t m) -> Patch::PairStep& {
      auto it = ps.find(m);
      if (it == ps.end()) {
        Patch::PairStep s{};
        s.m = m;
        it = ps.emplace(m, s).first;
      }
      return it->second;
    };
    auto plain_step = [&](int32_t dof) -> Patch::PlainStep& {
      auto it = pl.find(dof);
      if (it == pl.end()) {
        Patch::PlainStep s{};
        s.dof = dof;
        it = pl.emplace(dof, s).first;
      }
      return it->second;
    };
    for (int t = 0; t < pt.n_nodes; ++t) {
      const NodeRels& R = rel[pt.nodes[t]];
      for (const PairRel& r : R.pair) {
        Patch::PairStep& s = pair_step(r.m);
        s.mask |= 1u << t;
        for (int j = 0; j < 5; ++j) s.c[t][j] = r.c[j];
      }
      for (const PlainRel& r : R.plain) {
        Patch::PlainStep& s = plain_step(r.dof);
        s.mask |= 1u << t;
        s.aI[t] = r.aI;
        s.aJ[t] = r.aJ;
      }
    }
    if (pt.single >= 0) {
      const SingleRels& sr = srel[T.single_id[pt.single]];
      for (const SPair& sp : sr.pair) {
        Patch::PairStep& s = pair_step(sp.m);
        s.sI = sp.sI;
        s.sJ = sp.sJ;
      }
      for (const SPlain& sp : sr.plain) plain_step(sp.dof).s = sp.s;
    }
    for (auto& kv : ps) pt.pair.push_back(kv.second);
    for (auto& kv : pl) pt.plain.push_back(kv.second);
    // Runs cost latency (header read, dispatch), so classes are merged while that pads at most `merge_pad` target slots:
    // the cheapest pair of classes is merged into the union of their masks (padding slots carry zero coefficients).
    auto merge_classes = [&](auto& steps) {
      while (tune.merge_pad > 0) {
        int cnt[16] = {0};
        for (const auto& st : steps) cnt[st.mask & 15u]++;
        int best_a = -1, best_b = -1, best_cost = 1 << 30;
        for (int a = 0; a < 16; ++a)
          for (int b = a + 1; b < 16; ++b) {
            if (cnt[a] == 0 || cnt[b] == 0) continue;
            const int u = a | b;
            int cost = (popc4(u) - popc4(a)) * cnt[a] + (popc4(u) - popc4(b)) * cnt[b];
            if (u != a && u != b && cnt[u] > 0) cost += 0;  // joins an existing class
            if (cost < best_cost) {
              best_cost = cost;
              best_a = a;
              best_b = b;
            }
          }
        if (best_a < 0 || best_cost > tune.merge_pad) break;
        const uint32_t u = (uint32_t)(best_a | best_b);
        for (auto& st : steps)
          if ((int)st.mask == best_a || (int)st.mask == best_b) st.mask = u;
      }
    };
    merge_classes(pt.pair);
    merge_classes(pt.plain);
    // runs: steps sorted by (mask, source)
    std::stable_sort(pt.pair.begin(), pt.pair.end(), [](const Patch::PairStep& x, const Patch::PairStep& y) { return x.mask < y.mask; });
    std::stable_sort(pt.plain.begin(), pt.plain.end(), [](const Patch::PlainStep& x, const Patch::PlainStep& y) { return x.mask < y.mask; });
    // lines
    for (const Patch::PairStep& s : pt.pair) {
      pt.keys.push_back(key_of(T.nI[s.m], 0));
      pt.keys.push_back(key_of(T.nJ[s.m], 0));
      if (backward) {
        pt.keys.push_back(key_of(T.nI[s.m], 1));
        pt.keys.push_back(key_of(T.nJ[s.m], 1));
      }
      P.n_gathers += backward ? 4 : 2;
      P.slot_entries += 2 * popc4(s.mask);
    }
    for (const Patch::PlainStep& s : pt.plain) {
      pt.keys.push_back(key_of(s.dof, 0));
      P.n_gathers += 1;
      P.slot_entries += 2 * popc4(s.mask);
    }
    for (int t = 0; t < pt.n_nodes; ++t) {  // own lines of the epilogue: forward alpha[I], alpha[J]; backward r[I], r[J]
      pt.keys.push_back(key_of(T.nI[pt.nodes[t]], 0));
      pt.keys.push_back(key_of(T.nJ[pt.nodes[t]], 0));
    }
    std::sort(pt.keys.begin(), pt.keys.end());
    pt.keys.erase(std::unique(pt.keys.begin(), pt.keys.end()), pt.keys.end());
    // stream size
    int32_t u = kPatchHeaderWords;
    uint32_t last = 0xffffffffu;
    for (const Patch::PairStep& s : pt.pair) {
      if (s.mask != last) ++u;
      last = s.mask;
      u += pair_words(backward, popc4(s.mask));
    }
    last = 0xffffffffu;
    for (const Patch::PlainStep& s : pt.plain) {
      if (s.mask != last) ++u;
      last = s.mask;
      u += plain_words(popc4(s.mask));
    }
    pt.units = u;
    P.n_steps += (int64_t)pt.pair.size() + (int64_t)pt.plain.size();
    if ((int32_t)pt.keys.size() > tune.pool_lines / 2)
      return fail(FEO_ERR_UNSUPPORTED, "a patch needs more dof lines than the pool can hold; use the tile plan");
  }

  // ---- patch adjacency (for the order in which rounds grow) -------------------------------------------------------------
  // patch -> patches owning the nodes it gathers
  auto for_each_neighbour = [&](int32_t p, auto&& fn) {
    for (const Patch::PairStep& s : patches[p].pair) {
      const int32_t q = patch_of[s.m];
      if (q >= 0 && q != p) fn(q);
    }
  };
  // singles-only patches have no node of their own: link them through the nodes they gather (one direction is enough for BFS
  // from the node side if we also record the reverse)
  std::vector<std::vector<int32_t>> extra_adj(NP);
  for (int32_t p = 0; p < NP; ++p)
    if (patches[p].n_nodes == 0)
      for (const Patch::PairStep& s : patches[p].pair) {
        const int32_t q = patch_of[s.m];
        if (q >= 0) extra_adj[q].push_back(p);
      }

  // ---- rounds: compact groups of <= W patches grown by BFS; consecutive rounds are adjacent so that lines stay resident ----
  const int32_t W = tune.warps, C = tune.pool_lines;
  const int32_t cap_units = tune.stream_cap / 16 - (W + 1) / 2 - 1 - (round_lines_cap(tune) + 1) / 2;  // table, load list of the next round
  // a round may not fill the pool on its own: its successor needs room for the lines the two do not share
  const int32_t round_lines = round_lines_cap(tune);
  std::vector<int32_t> last_round((backward ? 2 : 1) * (size_t)n, -1000);
  std::vector<char> seen(NP, 0), placed(NP, 0);
  std::deque<int32_t> frontier, q;
  std::vector<std::vector<int32_t>> round_patches;
  std::vector<int32_t> round_seg_first;  // per round: 1 when it starts a segment
  int32_t next_unseen = 0, n_placed = 0, seg_first = 0, prev_size = 0;
  while (n_placed < NP) {
    const int32_t rho = (int32_t)round_patches.size();
    if (rho - seg_first >= tune.seg_rounds) seg_first = rho;
    bool cold = rho == seg_first;
    round_patches.emplace_back();
    round_seg_first.push_back(cold ? 1 : 0);
    std::vector<int32_t>& cur = round_patches.back();
    int32_t union_cnt = cold ? 0 : prev_size, cur_size = 0, units = 0;
    q.clear();
    while ((int32_t)cur.size() < W) {
      if (q.empty()) {
        int32_t seed = -1;
        while (!frontier.empty() && seed < 0) {
          const int32_t u = frontier.front();
          frontier.pop_front();
          if (!placed[u]) seed = u;
        }
        if (seed < 0) {
          while (next_unseen < NP && (placed[next_unseen] || seen[next_unseen])) ++next_unseen;
          if (next_unseen < NP) seed = next_unseen;
        }
        if (seed < 0) {  // everything left is queued somewhere: scan
          for (int32_t u = 0; u < NP && seed < 0; ++u)
            if (!placed[u]) seed = u;
        }
        if (seed < 0) break;
        seen[seed] = 1;
        q.push_back(seed);
      }
      const int32_t u = q.front();
      const Patch& pt = patches[u];
      int32_t add_union = 0, add_cur = 0;
      for (int64_t key : pt.keys) {
        const int32_t lr = last_round[key];
        if (lr == rho) continue;
        ++add_cur;
        if (!(lr == rho - 1 && !cold)) ++add_union;
      }
      if (!cur.empty() && (union_cnt + add_union > C || cur_size + add_cur > round_lines || units + pt.units > cap_units)) break;
      if (union_cnt + add_union > C && !cold) {
        // not even one patch fits next to the previous round: start a new segment here (the pool is drained first)
        cold = true;
        seg_first = rho;
        round_seg_first.back() = 1;
        union_cnt = 0;
        continue;
      }
      if (union_cnt + add_union > C || pt.units > cap_units)
        return fail(FEO_ERR_UNSUPPORTED, "a patch does not fit the line pool; use the tile plan");
      q.pop_front();
      for (int64_t key : pt.keys) last_round[key] = rho;
      union_cnt += add_union;
      cur_size += add_cur;
      units += pt.units;
      cur.push_back(u);
      placed[u] = 1;
      ++n_placed;
      auto visit = [&](int32_t v) {
        if (!seen[v] && !placed[v]) {
          seen[v] = 1;
          q.push_back(v);
        }
      };
      for_each_neighbour(u, visit);
      for (int32_t v : extra_adj[u]) visit(v);
    }
    // what is still queued seeds the next rounds, most recent first (the next round starts next to this one)
    for (auto it = q.rbegin(); it != q.rend(); ++it) {
      seen[*it] = 0;
      frontier.push_front(*it);
    }
    prev_size = cur_size;
    P.max_union_lines = std::max<int64_t>(P.max_union_lines, union_cnt);
    if (cur.empty()) {
      round_patches.pop_back();
      round_seg_first.pop_back();
    }
  }

  // ---- slots, loads and streams -------------------------------------------------------------------------------------------
  const int32_t NR = (int32_t)round_patches.size();
  std::vector<int32_t> key_slot((backward ? 2 : 1) * (size_t)n, -1);
  std::vector<int64_t> slot_key(C, -1);
  std::vector<int32_t> slot_last(C, -1000);
  std::vector<int32_t> free_slots;
  std::vector<std::vector<int32_t>> touched(NR);
  std::vector<int64_t> keys;
  std::vector<std::vector<uint32_t>> round_tbl, round_body;
  int32_t seg_begin = 0;
  for (int32_t rho = 0; rho < NR; ++rho) {
    if (round_seg_first[rho]) {
      P.seg_ptr.push_back(rho);
      seg_begin = rho;
      free_slots.clear();
      for (int32_t s = C - 1; s >= 0; --s) {
        free_slots.push_back(s);
        slot_last[s] = -1000;
        slot_key[s] = -1;
      }
    } else if (rho - 2 >= seg_begin) {
      for (int32_t s : touched[rho - 2])
        if (slot_last[s] == rho - 2) free_slots.push_back(s);
    }
    keys.clear();
    for (int32_t u : round_patches[rho]) keys.insert(keys.end(), patches[u].keys.begin(), patches[u].keys.end());
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    RoundInfo R{(int32_t)P.loads.size(), 0, 0, 0};
    for (int64_t key : keys) {
      int32_t s = key_slot[key];
      const bool resident = s >= 0 && slot_key[s] == key && slot_last[s] == rho - 1 && rho - 1 >= seg_begin;
      if (!resident) {
        if (free_slots.empty()) return fail(FEO_ERR_UNSUPPORTED, "patch plan: line pool exhausted");
        s = free_slots.back();
        free_slots.pop_back();
        slot_key[s] = key;
        key_slot[key] = s;
        const int32_t src = key >= n ? 1 : 0;
        P.loads.push_back(LineLoad{(uint32_t)(key - (int64_t)src * n) | ((uint32_t)src << 31), (uint32_t)s});
      }
      slot_last[s] = rho;
      touched[rho].push_back(s);
    }
    R.n_loads = (int32_t)P.loads.size() - R.load_begin;
    auto LINE = [&](int32_t dof, int32_t src) -> uint32_t { return (uint32_t)key_slot[key_of(dof, src)]; };
    // stream region: table, patches (32-bit slots, pieces aligned to 16-byte words)
    // (the region of round rho also carries the load list of round rho + 1, which is known one iteration later: the
    // patch part is kept per round and the regions are assembled after the loop)
    round_tbl.emplace_back((size_t)((W + 1) / 2) * 4, 0u);
    round_body.emplace_back();
    std::vector<uint32_t>& TB = round_tbl.back();
    std::vector<uint32_t>& S = round_body.back();
    const size_t base = 0;
    int32_t w = 0;
    auto pad16 = [&]() {
      while (S.size() & 3u) S.push_back(0u);
    };
    for (int32_t u : round_patches[rho]) {
      const Patch& pt = patches[u];
      TB[2 * (size_t)w] = (uint32_t)(S.size() * 4);  // relative to the start of the patch part; rebased below
      TB[2 * (size_t)w + 1] = 1u;
      ++w;
      // runs
      uint32_t n_runs = 0, last = 0xffffffffu;
      for (const Patch::PairStep& s : pt.pair) {
        n_runs += s.mask != last;
        last = s.mask;
      }
      last = 0xffffffffu;
      for (const Patch::PlainStep& s : pt.plain) {
        n_runs += s.mask != last;
        last = s.mask;
      }
      if (n_runs >= 4096) return fail(FEO_ERR_UNSUPPORTED, "patch with too many runs");
      S.insert(S.end(), {n_runs | (pt.single >= 0 ? 1u << 12 : 0u), (uint32_t)pt.single, 0u, 0u});
      for (int t = 0; t < kPatchNodes; ++t) {
        if (t < pt.n_nodes)
          S.insert(S.end(), {(uint32_t)T.nI[pt.nodes[t]], (uint32_t)T.nJ[pt.nodes[t]], LINE(T.nI[pt.nodes[t]], 0) << 8, LINE(T.nJ[pt.nodes[t]], 0) << 8});
        else
          S.insert(S.end(), {0xffffffffu, 0xffffffffu, 0u, 0u});
      }
      for (size_t i = 0; i < pt.pair.size();) {
        size_t j = i;
        while (j < pt.pair.size() && pt.pair[j].mask == pt.pair[i].mask) ++j;
        if (j - i >= (1u << 20)) return fail(FEO_ERR_UNSUPPORTED, "run too long");
        S.insert(S.end(), {0u | (pt.pair[i].mask << 4) | ((uint32_t)(j - i) << 8), 0u, 0u, 0u});
        for (; i < j; ++i) {
          const Patch::PairStep& s = pt.pair[i];
          S.push_back(LINE(T.nI[s.m], 0) << 8);
          S.push_back(LINE(T.nJ[s.m], 0) << 8);
          if (backward) {
            S.push_back(LINE(T.nI[s.m], 1) << 8);
            S.push_back(LINE(T.nJ[s.m], 1) << 8);
          }
          S.push_back(f2u(s.sI));
          S.push_back(f2u(s.sJ));
          for (int t = 0; t < kPatchNodes; ++t)
            if ((s.mask >> t) & 1u)
              for (int c = 0; c < (backward ? 5 : 3); ++c) S.push_back(f2u(s.c[t][c]));
          pad16();
        }
      }
      for (size_t i = 0; i < pt.plain.size();) {
        size_t j = i;
        while (j < pt.plain.size() && pt.plain[j].mask == pt.plain[i].mask) ++j;
        S.insert(S.end(), {1u | (pt.plain[i].mask << 4) | ((uint32_t)(j - i) << 8), 0u, 0u, 0u});
        for (; i < j; ++i) {
          const Patch::PlainStep& s = pt.plain[i];
          S.push_back(LINE(s.dof, 0) << 8);
          S.push_back(f2u(s.s));
          for (int t = 0; t < kPatchNodes; ++t)
            if ((s.mask >> t) & 1u) {
              S.push_back(f2u(s.aI[t]));
              S.push_back(f2u(s.aJ[t]));
            }
          pad16();
        }
      }
    }
    (void)base;
    P.rounds.push_back(R);
    if (P.stream.size() / 4 >= (size_t)std::numeric_limits<int32_t>::max()) return fail(FEO_ERR_UNSUPPORTED, "operator stream too large");
  }
  P.seg_ptr.push_back(NR);
  // assemble the regions: [table][load list of the next round of the segment][patches]
  for (int32_t rho = 0; rho < NR; ++rho) {
    RoundInfo& R = P.rounds[rho];
    const bool has_next = rho + 1 < NR && !round_seg_first[rho + 1];
    const int32_t n_next = has_next ? P.rounds[rho + 1].n_loads : 0;
    const size_t base = P.stream.size();
    std::vector<uint32_t>& TB = round_tbl[rho];
    const size_t list_slots = ((size_t)n_next * 2 + 3) / 4 * 4;
    const uint32_t body_off = (uint32_t)((TB.size() + list_slots) * 4);
    for (int32_t w = 0; w < W; ++w)
      if (TB[2 * w + 1] != 0u) TB[2 * w] += body_off;
    P.stream.insert(P.stream.end(), TB.begin(), TB.end());
    for (int32_t i = 0; i < n_next; ++i) {
      const LineLoad& L = P.loads[P.rounds[rho + 1].load_begin + i];
      P.stream.push_back(L.dof_src);
      P.stream.push_back(L.slot);
    }
    while ((P.stream.size() - base) & 3u) P.stream.push_back(0u);
    P.stream.insert(P.stream.end(), round_body[rho].begin(), round_body[rho].end());
    R.stream_begin = (int32_t)(base / 4);
    R.n_words = (int32_t)((P.stream.size() - base) / 4);
    if (R.n_words * 16 > tune.stream_cap) return fail(FEO_ERR_UNSUPPORTED, "patch plan: round stream exceeds its buffer");
    if (P.stream.size() / 4 >= (size_t)std::numeric_limits<int32_t>::max()) return fail(FEO_ERR_UNSUPPORTED, "operator stream too large");
    std::vector<uint32_t>().swap(round_body[rho]);
  }
  P.applicable = true;
  if (std::getenv("FEO_PLAN_DEBUG")) {
    long hist[5] = {0, 0, 0, 0, 0}, with_single = 0;
    for (const Patch& pt : patches) {
      hist[pt.n_nodes]++;
      with_single += pt.single >= 0;
    }
    long n_pair = 0, n_plain = 0, n_runs = 0, k_pair = 0, k_plain = 0;
    for (const Patch& pt : patches) {
      uint32_t last = 0xffffffffu;
      for (const auto& s : pt.pair) { n_runs += s.mask != last; last = s.mask; k_pair += popc4(s.mask); }
      last = 0xffffffffu;
      for (const auto& s : pt.plain) { n_runs += s.mask != last; last = s.mask; k_plain += popc4(s.mask); }
      n_pair += (long)pt.pair.size();
      n_plain += (long)pt.plain.size();
    }
    fprintf(stderr, "[patch plan %s] per patch: %.1f pair steps (%.1f targets), %.1f plain steps (%.1f targets), %.1f runs\n", backward ? "bwd" : "fwd",
            (double)n_pair / NP, (double)k_pair / NP, (double)n_plain / NP, (double)k_plain / NP, (double)n_runs / NP);
    fprintf(stderr,
            "[patch plan %s] nodes %d singles %zu patches %d (by nodes 0..4: %ld %ld %ld %ld %ld; with single %ld) rounds %d segments %d "
            "loads/dof %.2f gathers/dof %.2f stream %.1f MB (%.0f B/patch) max union %ld of %d\n",
            backward ? "bwd" : "fwd", NN, T.singles.size(), NP, hist[0], hist[1], hist[2], hist[3], hist[4], with_single, NR, P.n_segments(),
            (double)P.loads.size() / n, (double)P.n_gathers / n, P.stream.size() * 4e-6, P.stream.size() * 4.0 / NP, (long)P.max_union_lines, C);
  }
  return FEO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// fp64 replay
// ---------------------------------------------------------------------------------------------------------------
int replay_patch_plan(const PatchPlan& P, bool has_conv, int32_t ns_branch, const double* in0, const double* in1, double* out) {
  if (!P.applicable) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan not applicable: " + P.why_not);
  const bool precond = has_conv ? ns_branch != 0 : true;
  const double esign = precond ? 1.0 : -1.0;
  const bool bw = P.backward;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<double> pool(P.pool_lines, nan);
  // the first round of a segment takes its load list from P.loads; every other round from the stream region of its
  // predecessor (right after the warp table), as the producer warps do
  auto apply_loads = [&](int32_t rho, bool first) -> int {
    const RoundInfo& R = P.rounds[rho];
    const uint32_t* list = nullptr;
    if (!first) {
      const RoundInfo& Q = P.rounds[rho - 1];
      const size_t tbl = (size_t)((P.warps + 1) / 2) * 4;
      if ((size_t)Q.n_words * 4 < tbl + (size_t)R.n_loads * 2) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: embedded load list out of bounds");
      list = P.stream.data() + (size_t)Q.stream_begin * 4 + tbl;
    }
    for (int32_t i = 0; i < R.n_loads; ++i) {
      const LineLoad& L0 = P.loads[R.load_begin + i];
      const uint32_t ds = list ? list[2 * i] : L0.dof_src, slot = list ? list[2 * i + 1] : L0.slot;
      const uint32_t dof = ds & 0x7fffffffu, src = ds >> 31;
      if ((int32_t)slot >= P.pool_lines || (int32_t)dof >= P.n) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: load out of range");
      pool[slot] = bw ? (src ? in1[dof] : in0[dof]) : in0[dof];
    }
    return FEO_OK;
  };
  for (int32_t sg = 0; sg < P.n_segments(); ++sg) {
    const int32_t r0 = P.seg_ptr[sg], r1 = P.seg_ptr[sg + 1];
    std::fill(pool.begin(), pool.end(), nan);  // an item starts with nothing resident
    if (int rc = apply_loads(r0, true)) return rc;
    for (int32_t rho = r0; rho < r1; ++rho) {
      // the loads of the next round may land while this one runs
      if (rho + 1 < r1)
        if (int rc = apply_loads(rho + 1, false)) return rc;
      const RoundInfo& R = P.rounds[rho];
      if (R.n_words * 16 > P.stream_cap) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: round stream larger than its buffer");
      const uint32_t* reg = P.stream.data() + (size_t)R.stream_begin * 4;
      const uint32_t* const end = reg + (size_t)R.n_words * 4;
      bool bad = false;
      auto S = [&](uint32_t off) -> double {  // byte offset of a line in the pool
        if ((off & 255u) != 0u || (int32_t)(off >> 8) >= P.pool_lines) {
          bad = true;
          return 0.0;
        }
        return pool[off >> 8];
      };
      for (int32_t w = 0; w < P.warps; ++w) {
        const uint32_t off = reg[2 * w], np = reg[2 * w + 1];
        if (np == 0) continue;
        if (off % 16 != 0 || (int32_t)(off / 16) >= R.n_words) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: bad warp table");
        const uint32_t* s = reg + off / 4;
        for (uint32_t ip = 0; ip < np; ++ip) {
          if (s + 4 * kPatchHeaderWords > end) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: patch header out of bounds");
          const uint32_t n_runs = s[0] & 0xfffu;
          const bool has_single = (s[0] >> 12) & 1u;
          const int32_t single = (int32_t)s[1];
          int32_t dI[kPatchNodes], dJ[kPatchNodes];
          uint32_t ownI[kPatchNodes], ownJ[kPatchNodes];
          for (int t = 0; t < kPatchNodes; ++t) {
            dI[t] = (int32_t)s[4 + 4 * t];
            dJ[t] = (int32_t)s[5 + 4 * t];
            ownI[t] = s[6 + 4 * t];
            ownJ[t] = s[7 + 4 * t];
          }
          s += 4 * kPatchHeaderWords;
          double A[kPatchNodes][6];  // forward: aI uI vI aJ uJ vJ ; backward: gI gJ b1I b2I b1J b2J
          for (auto& a : A)
            for (double& v : a) v = 0.0;
          double SA = 0.0;
          for (uint32_t r = 0; r < n_runs; ++r) {
            if (s + 4 > end) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: run header out of bounds");
            const uint32_t meta = s[0];
            s += 4;
            const uint32_t kind = meta & 15u, mask = (meta >> 4) & 15u, count = meta >> 8;
            const int k = popc4(mask);
            if (kind > 1 || count == 0) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: bad run header");
            for (uint32_t i = 0; i < count; ++i) {
              const int nw = kind == 0 ? pair_words(bw, k) : plain_words(k);
              if (s + 4 * nw > end) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: step out of bounds");
              auto FL = [&](int j) { return (double)u2f(s[j]); };
              if (kind == 0 && !bw) {
                const double x = S(s[0]), y = S(s[1]);
                SA += FL(2) * x + FL(3) * y;
                int j = 4;
                for (int t = 0; t < kPatchNodes; ++t)
                  if ((mask >> t) & 1u) {
                    A[t][0] += FL(j) * x;
                    A[t][1] += FL(j + 1) * x;
                    A[t][2] += FL(j + 2) * x;
                    A[t][3] += FL(j) * y;
                    A[t][4] += FL(j + 1) * y;
                    A[t][5] += FL(j + 2) * y;
                    j += 3;
                  }
              } else if (kind == 0) {
                const double rI = S(s[0]), rJ = S(s[1]), d1 = S(s[2]), d2 = S(s[3]);
                SA += FL(4) * rI + FL(5) * rJ;
                int j = 6;
                for (int t = 0; t < kPatchNodes; ++t)
                  if ((mask >> t) & 1u) {
                    const double tt = FL(j) + FL(j + 1) * d1 + FL(j + 2) * d2;
                    A[t][0] += rI * tt;
                    A[t][1] += rJ * tt;
                    A[t][2] += FL(j + 3) * d1;
                    A[t][3] += FL(j + 4) * d1;
                    A[t][4] += FL(j + 3) * d2;
                    A[t][5] += FL(j + 4) * d2;
                    j += 5;
                  }
              } else {
                const double x = S(s[0]);
                SA += FL(1) * x;
                int j = 2;
                for (int t = 0; t < kPatchNodes; ++t)
                  if ((mask >> t) & 1u) {
                    A[t][0] += FL(j) * x;
                    A[t][bw ? 1 : 3] += FL(j + 1) * x;
                    j += 2;
                  }
              }
              s += 4 * nw;
            }
          }
          for (int t = 0; t < kPatchNodes; ++t) {
            if (dI[t] < 0) continue;
            if (dI[t] >= P.n || dJ[t] < 0 || dJ[t] >= P.n) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: node dof out of range");
            const double o1 = S(ownI[t]), o2 = S(ownJ[t]);
            if (!bw) {
              const double cI = o1 * A[t][1] + o2 * A[t][2], cJ = o1 * A[t][4] + o2 * A[t][5];
              out[dI[t]] = precond ? A[t][0] - (in1[dI[t]] - cI) : A[t][0] - (-in1[dI[t]] + cI);
              out[dJ[t]] = precond ? A[t][3] - (in1[dJ[t]] - cJ) : A[t][3] - (-in1[dJ[t]] + cJ);
            } else {
              out[dI[t]] = A[t][0] + esign * (A[t][2] * o1 + A[t][4] * o2);
              out[dJ[t]] = A[t][1] + esign * (A[t][3] * o1 + A[t][5] * o2);
            }
          }
          if (has_single) {
            if (single < 0 || single >= P.n) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: single dof out of range");
            if (!bw)
              out[single] = precond ? SA - (in1[single] - 0.0) : SA - (-in1[single] + 0.0);
            else
              out[single] = SA;
          }
        }
      }
      if (bad) return fail(FEO_ERR_INVALID_ARGUMENT, "patch plan: stream references a slot outside the pool");
    }
  }
  return FEO_OK;
}

}  // namespace feo
