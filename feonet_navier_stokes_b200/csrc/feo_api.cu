// extern "C" entry points of libfeonet_b200.so (declared in include/feonet_b200.h).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "feo_internal.h"
#include "feo_patch.h"
#include "feo_lattice.h"

namespace feo {
namespace {
template <typename T>
int upload(feo_operator* op, const std::vector<T>& host, T** dev) {
  *dev = nullptr;
  size_t bytes = std::max<size_t>(host.size() * sizeof(T), 16);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(FEO_ERR_OUT_OF_MEMORY, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  op->allocations.push_back(p);
  op->device_bytes += (int64_t)bytes;
  if (!host.empty()) FEO_CUDA_CHECK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = reinterpret_cast<T*>(p);
  return FEO_OK;
}

int upload_csr(feo_operator* op, const HostCsr& h, DevCsr* d) {
  if (!h.present()) return FEO_OK;
  if (int rc = upload(op, h.rowptr, &d->rowptr)) return rc;
  if (int rc = upload(op, h.col, &d->col)) return rc;
  if (int rc = upload(op, h.val, &d->val)) return rc;
  d->nnz = h.nnz();
  return FEO_OK;
}

// host [n,n] row-major -> device [n, ceil4(n)] zero padded (optionally transposed)
int upload_dense(feo_operator* op, const float* src, int32_t n, bool transposed, float** dev) {
  const int32_t ld = (n + 3) / 4 * 4;
  std::vector<float> buf((size_t)n * ld, 0.f);
  for (int32_t r = 0; r < n; ++r)
    for (int32_t c = 0; c < n; ++c) buf[(size_t)r * ld + c] = transposed ? src[(size_t)c * n + r] : src[(size_t)r * n + c];
  return upload(op, buf, dev);
}

int check_handle(feo_handle_t h) {
  if (h == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "operator handle is NULL");
  return FEO_OK;
}

int prepare(const feo_operator_desc* desc, HostCsr* A, HostCsr* B1, HostCsr* B2, HostCsr* S) {
  if (desc == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "desc is NULL");
  if (desc->abi_version != FEO_ABI_VERSION) return fail(FEO_ERR_INVALID_ARGUMENT, "ABI version mismatch");
  if (desc->n <= 0) return fail(FEO_ERR_INVALID_ARGUMENT, "n must be positive");
  if (int rc = canonicalize(desc->A, desc->n, "A", A)) return rc;
  if (int rc = canonicalize(desc->B1, desc->n, "B1", B1)) return rc;
  if (int rc = canonicalize(desc->B2, desc->n, "B2", B2)) return rc;
  if (int rc = canonicalize(desc->S, desc->n, "S", S)) return rc;
  if (!A->present() && desc->dense_m == nullptr && desc->dense_p == nullptr)
    return fail(FEO_ERR_INVALID_ARGUMENT, "operator needs A (CSR) or a dense matrix");
  if (desc->n_u < 0 || (desc->n_u > 0 && (desc->idx_i == nullptr || desc->idx_j == nullptr)))
    return fail(FEO_ERR_INVALID_ARGUMENT, "idx_i/idx_j missing");
  return FEO_OK;
}
}  // namespace
}  // namespace feo

using namespace feo;

extern "C" {

int feo_abi_version(void) { return FEO_ABI_VERSION; }
const char* feo_last_error_string(void) { return feo::last_error().c_str(); }

int feo_op_create(const feo_operator_desc* desc, feo_handle_t* out) {
  if (out == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  HostCsr A, B1, B2, S;
  if (int rc = prepare(desc, &A, &B1, &B2, &S)) return rc;
  feo_operator* op = new (std::nothrow) feo_operator();
  if (op == nullptr) return fail(FEO_ERR_OUT_OF_MEMORY, "host allocation failed");
  auto bail = [&](int rc) {
    feo_op_destroy(op);
    return rc;
  };
  op->n = desc->n;
  op->n_u = desc->n_u;
  op->ns_branch = desc->ns_precond_branch ? 1 : 0;
  op->dt = desc->dt;
  int rc;
  const HostCsr* mats[4] = {&A, &B1, &B2, &S};
  for (int m = 0; m < 4; ++m) {
    if ((rc = upload_csr(op, *mats[m], &op->csr[m]))) return bail(rc);
    if ((rc = upload_csr(op, transpose(*mats[m]), &op->csrT[m]))) return bail(rc);
    op->nnz[m] = mats[m]->nnz();
  }
  if (S.present() && A.present()) {  // time-dependent operator M = S + dt*A
    HostCsr M = axpy(S, desc->dt, A);
    if ((rc = upload_csr(op, M, &op->csr[FEO_MAT_M]))) return bail(rc);
    if ((rc = upload_csr(op, transpose(M), &op->csrT[FEO_MAT_M]))) return bail(rc);
    op->nnz[FEO_MAT_M] = M.nnz();
    const SeqPlan pf = build_seq_plan(M, S), pb = build_seq_plan(transpose(M), transpose(S));
    if ((rc = upload(op, pf.rowptr, &op->seq_f.rowptr)) || (rc = upload(op, pf.ent, &op->seq_f.ent))) return bail(rc);
    if ((rc = upload(op, pb.rowptr, &op->seq_b.rowptr)) || (rc = upload(op, pb.ent, &op->seq_b.ent))) return bail(rc);
    op->has_seq = true;
  }
  if (desc->n_u > 0) {
    std::vector<int32_t> ii(desc->idx_i, desc->idx_i + desc->n_u), jj(desc->idx_j, desc->idx_j + desc->n_u);
    for (int32_t k = 0; k < desc->n_u; ++k)
      if (ii[k] < 0 || ii[k] >= desc->n || jj[k] < 0 || jj[k] >= desc->n)
        return bail(fail(FEO_ERR_INVALID_ARGUMENT, "idx_sol entry out of range"));
    if ((rc = upload(op, ii, &op->idx_i))) return bail(rc);
    if ((rc = upload(op, jj, &op->idx_j))) return bail(rc);
  }
  if (A.present()) {
    // Plan of the fused residual kernels.  FEO_PLAN = auto (default): the lattice plan (feo_lattice.h) when the operator is a
    // structured P2-P1 lattice it models, else the tile plan; tile / patch / lattice force one of them (patch and lattice
    // fail when the operator does not fit).  All are device code paths.
    const char* plan_env = std::getenv("FEO_PLAN");
    const std::string want = plan_env != nullptr ? plan_env : "auto";
    bool use_lattice = false;
    if (want == "auto" || want == "lattice") {
      LatticePlan lp;
      const int32_t branch = (op->ns_branch || !(B1.present() || B2.present())) ? 1 : 0;
      if ((rc = build_lattice_plan(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, branch, &lp))) return bail(rc);
      use_lattice = lp.applicable;
      if (!use_lattice && want == "lattice")
        return bail(fail(FEO_ERR_UNSUPPORTED, "FEO_PLAN=lattice but the operator is not a lattice the plan models: " + lp.why_not));
      if (use_lattice) {
        DevLatticePlan& D = op->lattice;
        D.n = lp.n;
        D.nc = lp.nc;
        D.has_conv = lp.has_conv;
        const char* ew = std::getenv("FEO_LATTICE_ELEMENT");
        D.element_walk = ew != nullptr && atoi(ew) != 0;
        for (int dir = 0; dir < kLatTables; ++dir) {
          D.n_classes[dir] = lp.n_classes[dir];
          D.exist[dir] = lp.exist[dir];
          D.tab[dir] = lp.tab[dir];
          std::memcpy(D.cat_cls[dir], lp.cat_cls[dir], 25);
        }
        D.present = true;
        op->has_conv = lp.has_conv;
        op->nnz_union = lp.real_entries;
      }
    }
    bool use_patch = false;
    if (want == "patch") {  // never chosen automatically: slower than the tile plan at cfg5
      Front F;
      if ((rc = build_front(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, op->ns_branch, false, &F))) return bail(rc);
      PatchPlan pf, pb;
      int rf = build_patch_plan(F, false, patch_tuning_from_env(false), &pf);
      int rb = rf == FEO_OK && pf.applicable ? build_patch_plan(F, true, patch_tuning_from_env(true), &pb) : rf;
      use_patch = rf == FEO_OK && rb == FEO_OK && pf.applicable && pb.applicable;
      if (!use_patch && want == "patch")
        return bail(fail(FEO_ERR_UNSUPPORTED, "FEO_PLAN=patch but the operator does not fit the patch plan: " +
                                                  (rf != FEO_OK || rb != FEO_OK ? last_error() : (pf.applicable ? pb.why_not : pf.why_not))));
      if (use_patch) {
        for (int bw = 0; bw < 2; ++bw) {
          const PatchPlan& H = bw ? pb : pf;
          DevPatchPlan& D = bw ? op->patch_b : op->patch_f;
          D.warps = H.warps;
          D.producers = H.producers;
          D.pool_lines = H.pool_lines;
          D.stream_cap = H.stream_cap;
          D.n_segments = H.n_segments();
          D.n_rounds = (int32_t)H.rounds.size();
          if ((rc = upload(op, H.seg_ptr, &D.seg_ptr))) return bail(rc);
          if ((rc = upload(op, H.rounds, &D.rounds))) return bail(rc);
          if ((rc = upload(op, H.loads, &D.loads))) return bail(rc);
          if ((rc = upload(op, H.stream, &D.stream))) return bail(rc);
          D.present = true;
        }
        op->has_conv = F.conv;
        op->nnz_union = pf.real_entries;
      }
    }
    // tile plans of the fused residual kernels (forward: row-owned, backward: column-pair-owned)
    for (int bw = 0; bw < 2 && !use_patch && !use_lattice; ++bw) {
      TilePlan plan;
      if ((rc = build_tile_plan(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, op->ns_branch, bw != 0,
                                tile_tuning_from_env(bw != 0), &plan)))
        return bail(rc);
      DevTilePlan& D = bw ? op->tiles_b : op->tiles_f;
      D.n_tiles = plan.n_tiles;
      D.max_lines = plan.max_lines;
      D.warps = plan.warps;
      D.stages = plan.stages;
      if ((rc = upload(op, plan.tile_box_ptr, &D.tile_box_ptr))) return bail(rc);
      if ((rc = upload(op, plan.tile_lines, &D.tile_lines))) return bail(rc);
      if ((rc = upload(op, plan.boxes, &D.boxes))) return bail(rc);
      if ((rc = upload(op, plan.warp_range, &D.warp_range))) return bail(rc);
      if ((rc = upload(op, plan.stream, &D.stream))) return bail(rc);
      if (!bw) {
        op->has_conv = plan.has_conv;
        op->nnz_union = plan.real_entries;
      }
    }
    if (!op->has_conv) op->ns_branch = 1;  // linear Stokes: r = A a - F (FEONet_Stokes_square/train_FEONet.py:264-270)
    op->has_sparse = true;
    for (int32_t r = 0; r < A.n; ++r) op->max_row_nnz = std::max(op->max_row_nnz, A.rowptr[r + 1] - A.rowptr[r]);
  }
  if (desc->dense_m != nullptr) {
    if ((rc = upload_dense(op, desc->dense_m, desc->n, false, &op->dM))) return bail(rc);
    if ((rc = upload_dense(op, desc->dense_m, desc->n, true, &op->dMT))) return bail(rc);
    if ((rc = upload(op, dense_split_tiles(desc->dense_m, desc->n, false), &op->dMs))) return bail(rc);
    if ((rc = upload(op, dense_split_tiles(desc->dense_m, desc->n, true), &op->dMTs))) return bail(rc);
  }
  if (desc->dense_p != nullptr) {
    if ((rc = upload_dense(op, desc->dense_p, desc->n, false, &op->dP))) return bail(rc);
    if ((rc = upload(op, dense_split_tiles(desc->dense_p, desc->n, false), &op->dPs))) return bail(rc);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(FEO_ERR_CUDA, "device synchronize failed after upload"));
  *out = op;
  return FEO_OK;
}

int feo_op_destroy(feo_handle_t h) {
  if (h == nullptr) return FEO_OK;
  for (void* p : h->allocations) cudaFree(p);
  delete h;
  return FEO_OK;
}

int feo_op_get_info(feo_handle_t h, feo_op_info* info) {
  if (int rc = check_handle(h)) return rc;
  if (info == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "info is NULL");
  std::memset(info, 0, sizeof(*info));
  info->n = h->n;
  info->n_u = h->n_u;
  info->has_conv = h->has_conv;
  info->has_seq = h->has_seq;
  info->has_dense_m = h->dM != nullptr;
  info->has_dense_p = h->dP != nullptr;
  info->nnz_a = h->nnz[0];
  info->nnz_b1 = h->nnz[1];
  info->nnz_b2 = h->nnz[2];
  info->nnz_s = h->nnz[3];
  info->nnz_union = h->nnz_union;
  // patch plan: rounds; lattice plan: cells
  info->n_tiles_fwd = h->lattice.present ? h->lattice.nc * h->lattice.nc : h->patch_f.present ? h->patch_f.n_rounds : h->tiles_f.n_tiles;
  info->n_tiles_bwd = h->lattice.present ? h->lattice.nc * h->lattice.nc : h->patch_b.present ? h->patch_b.n_rounds : h->tiles_b.n_tiles;
  info->max_row_nnz = h->max_row_nnz;
  info->device_bytes = h->device_bytes;
  return FEO_OK;
}

size_t feo_workspace_bytes(feo_handle_t h, int32_t B, int32_t T) {
  if (h == nullptr || B <= 0) return 0;
  if (T < 1) T = 1;
  size_t need = loss_partials_needed(h->n, std::max(std::max(h->tiles_f.warps, h->patch_f.warps), h->lattice.present ? lattice_fwd_warps() : 0),
                                     (int64_t)B * T);
  if (h->dMs != nullptr || h->dPs != nullptr) {
    // dense applies: loss partials of the tile grid (1 KB aligned), then the pre-split activations (feo_dense_tc.cu)
    const int64_t cols = ((int64_t)B * T + 3) / 4 * 4;
    const size_t tc_count = (size_t)((cols + 63) / 64 + 1) * (size_t)((h->n + 127) / 128);
    need = std::max(need, (tc_count * sizeof(float) + 1023) / 1024 * 1024 + dense_xsplit_bytes(h->n, (int64_t)B * T));
  }
  return need;
}

int feo_transpose(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                  const int32_t* dst_row_map, void* stream) {
  return launch_transpose(src, src_ld, dst, dst_ld, rows, cols, dst_row_map, nullptr, (cudaStream_t)stream);
}

int feo_transpose_gather(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                         const int32_t* src_row_map, void* stream) {
  return launch_transpose(src, src_ld, dst, dst_ld, rows, cols, nullptr, src_row_map, (cudaStream_t)stream);
}

int feo_op_plan(feo_handle_t h) {
  if (h == nullptr || !h->has_sparse) return FEO_PLAN_NONE;
  return h->lattice.present ? FEO_PLAN_LATTICE : h->patch_f.present ? FEO_PLAN_PATCH : FEO_PLAN_TILE;
}

int feo_residual_fwd(feo_handle_t h, const float* alphaT, const float* fT, int64_t ldb, int32_t B, float* loss_out,
                     float* rT, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!h->has_sparse) return fail(FEO_ERR_UNSUPPORTED, "operator has no sparse A: use feo_dense_apply");
  return launch_residual_fwd(h, alphaT, fT, ldb, B, loss_out, rT, workspace, workspace_bytes, (cudaStream_t)stream);
}

int feo_residual_bwd(feo_handle_t h, const float* alphaT, const float* rT, const float* grad_loss, float* gradT, int64_t ldb,
                     int32_t B, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!h->has_sparse) return fail(FEO_ERR_UNSUPPORTED, "operator has no sparse A: use feo_dense_apply");
  return launch_residual_bwd(h, alphaT, rT, grad_loss, gradT, ldb, B, (cudaStream_t)stream);
}

int feo_spmm(feo_handle_t h, int32_t which, int32_t transpose, const float* XT, float* YT, int64_t ldb, int32_t B,
             float scale, int32_t accumulate, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (which < 0 || which > FEO_MAT_M) return fail(FEO_ERR_INVALID_ARGUMENT, "spmm: unknown matrix id");
  const DevCsr& K = transpose ? h->csrT[which] : h->csr[which];
  return launch_spmm(K, h->n, XT, YT, ldb, B, scale, accumulate, (cudaStream_t)stream);
}

int feo_dense_apply(feo_handle_t h, int32_t which, const float* XT, float* CT, int64_t ldb, int32_t B, float scale,
                    const float* scale_dev, const float* sub, float* loss_out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  if (int rc = check_handle(h)) return rc;
  const float* D = which == FEO_DENSE_M ? h->dM : which == FEO_DENSE_MT ? h->dMT : which == FEO_DENSE_P ? h->dP : nullptr;
  if (which < 0 || which > FEO_DENSE_P) return fail(FEO_ERR_INVALID_ARGUMENT, "dense_apply: unknown matrix id");
  const float* Ds = which == FEO_DENSE_M ? h->dMs : which == FEO_DENSE_MT ? h->dMTs : which == FEO_DENSE_P ? h->dPs : nullptr;
  return launch_dense(D, Ds, h->n, XT, CT, ldb, B, scale, scale_dev, sub, loss_out, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

int feo_seq_fwd(feo_handle_t h, const float* predT, const float* u0T, const float* fT, int64_t ldj, int64_t ldb,
                int32_t B, int32_t T, float* loss_out, float* rT, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  return launch_seq(h->seq_f, h->n, false, predT, u0T, fT, h->dt, ldj, ldb, B, T, nullptr, rT,
                    loss_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int feo_seq_bwd(feo_handle_t h, const float* rT, const float* grad_loss, float* gradT, int64_t ldj, int32_t B, int32_t T,
                void* stream) {
  if (int rc = check_handle(h)) return rc;
  return launch_seq(h->seq_b, h->n, true, rT, nullptr, nullptr, h->dt, ldj, 0, B, T,
                    grad_loss, gradT, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int feo_assemble_u_init(feo_handle_t h, const float* init_x, const float* init_y, float* u0T, int64_t ldb, int32_t B,
                        void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (h->n_u <= 0 || h->idx_i == nullptr) return fail(FEO_ERR_INVALID_ARGUMENT, "operator has no idx_sol");
  if (init_x == nullptr || init_y == nullptr || u0T == nullptr || ldb < B || B <= 0)
    return fail(FEO_ERR_INVALID_ARGUMENT, "assemble_u_init: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  FEO_CUDA_CHECK(cudaMemsetAsync(u0T, 0, (size_t)h->n * ldb * sizeof(float), st));
  if (int rc = launch_transpose(init_x, h->n_u, u0T, ldb, B, h->n_u, h->idx_i, nullptr, st)) return rc;
  return launch_transpose(init_y, h->n_u, u0T, ldb, B, h->n_u, h->idx_j, nullptr, st);
}

int feo_sincos_forcing_grid(const float* coeff_f, int32_t B, int32_t resol_in, float* value_f, void* stream) {
  return launch_sincos_grid(coeff_f, B, resol_in, value_f, (cudaStream_t)stream);
}

int feo_sq_diff_sum(const float* xT, const float* yT, int32_t n, int64_t ldb, int32_t B, float scale, float* loss_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return launch_sq_diff_sum(xT, yT, n, ldb, B, scale, loss_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- test hook (host only; exercised by the CPU test-suite, never by the product path) ----------
// Builds the tile plan of the fused residual kernels on the host and replays its staging boxes and
// per-warp streams in fp64 for ONE sample, decoding them exactly as the kernels do, so the CPU suite
// can compare the set-up code (union pattern, partner lookups, sign folding, pair/duo streams) with
// the oracle without a GPU.  forward: in0 = alpha, in1 = f, out = r; backward: in0 = r, in1 = alpha,
// out = grad / (2 g).  max_lines / warps <= 0: defaults.
// stats[0..7] = {n_tiles, max_lines, total_lines, n_boxes, n_items, real_entries, slot_entries, stream_words}.
int feo_debug_tile_replay(const feo_operator_desc* desc, int32_t backward, int32_t max_lines, int32_t warps,
                          const double* in0, const double* in1, double* out, int64_t* stats) {
  HostCsr A, B1, B2, S;
  if (int rc = prepare(desc, &A, &B1, &B2, &S)) return rc;
  if (!A.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "tile replay needs A");
  TileTuning tune = tile_tuning_from_env(backward != 0);
  if (max_lines > 0) tune.max_lines = max_lines;
  if (warps > 0) tune.warps = warps;
  TilePlan T;
  const int32_t branch = desc->ns_precond_branch ? 1 : 0;
  if (int rc = build_tile_plan(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, branch, backward != 0, tune, &T)) return rc;
  if (stats != nullptr) {
    stats[0] = T.n_tiles;
    stats[1] = T.max_lines;
    stats[2] = T.total_lines;
    stats[3] = (int64_t)T.boxes.size();
    stats[4] = T.n_items;
    stats[5] = T.real_entries;
    stats[6] = T.slot_entries;
    stats[7] = (int64_t)T.stream.size();
  }
  if (in0 == nullptr || in1 == nullptr || out == nullptr) return FEO_OK;
  return replay_tile_plan(T, branch, in0, in1, out);
}

int feo_debug_patch_replay(const feo_operator_desc* desc, int32_t backward, int32_t warps, int32_t pool_lines, int32_t seg_rounds,
                           const double* in0, const double* in1, double* out, int64_t* stats) {
  HostCsr A, B1, B2, S;
  if (int rc = prepare(desc, &A, &B1, &B2, &S)) return rc;
  if (!A.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "patch replay needs A");
  const int32_t branch = desc->ns_precond_branch ? 1 : 0;
  Front F;
  if (int rc = build_front(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, branch, false, &F)) return rc;
  PatchTuning tune = patch_tuning_from_env(backward != 0);
  if (warps > 0) tune.warps = warps;
  if (pool_lines > 0) tune.pool_lines = pool_lines;
  if (seg_rounds > 0) tune.seg_rounds = seg_rounds;
  PatchPlan P;
  if (int rc = build_patch_plan(F, backward != 0, tune, &P)) return rc;
  if (stats != nullptr) {
    stats[0] = P.applicable ? 1 : 0;
    stats[1] = P.n_patches;
    stats[2] = (int64_t)P.rounds.size();
    stats[3] = P.applicable ? P.n_segments() : 0;
    stats[4] = (int64_t)P.loads.size();
    stats[5] = P.n_gathers;
    stats[6] = P.real_entries;
    stats[7] = P.slot_entries;
    stats[8] = (int64_t)P.stream.size() / 4;
    stats[9] = P.max_union_lines;
  }
  if (!P.applicable) return fail(FEO_ERR_UNSUPPORTED, "patch plan not applicable: " + P.why_not);
  if (in0 == nullptr || in1 == nullptr || out == nullptr) return FEO_OK;
  return replay_patch_plan(P, F.conv, branch, in0, in1, out);
}

int feo_debug_lattice_replay(const feo_operator_desc* desc, int32_t backward, const double* in0, const double* in1, double* out,
                             int64_t* stats) {
  HostCsr A, B1, B2, S;
  if (int rc = prepare(desc, &A, &B1, &B2, &S)) return rc;
  if (!A.present()) return fail(FEO_ERR_INVALID_ARGUMENT, "lattice replay needs A");
  const int32_t branch = (desc->ns_precond_branch || !(B1.present() || B2.present())) ? 1 : 0;
  LatticePlan P;
  if (int rc = build_lattice_plan(A, B1, B2, desc->n_u, desc->idx_i, desc->idx_j, branch, &P)) return rc;
  if (stats != nullptr) {
    stats[0] = P.applicable ? 1 : 0;
    stats[1] = P.n;
    stats[2] = P.n_classes[0];
    stats[3] = P.n_classes[1];
    stats[4] = P.n_coef[0];
    stats[5] = P.n_coef[1];
  }
  if (!P.applicable) return fail(FEO_ERR_UNSUPPORTED, "lattice plan not applicable: " + P.why_not);
  if (in0 == nullptr || in1 == nullptr || out == nullptr) return FEO_OK;
  return replay_lattice_plan(P, backward, branch, in0, in1, out);  // backward = table: 0 forward, 1 backward, 2 element-walk forward
}

}  // extern "C"

int64_t feo_debug_dense_split_replay(const float* dense, int32_t n, int32_t transposed, const double* x, double* out_hi,
                                     double* out_lo) {
  if (dense == nullptr || n <= 0) return feo::fail(FEO_ERR_INVALID_ARGUMENT, "dense_split_replay: bad arguments");
  const std::vector<float> t = feo::dense_split_tiles(dense, n, transposed != 0);
  if (x != nullptr && out_hi != nullptr && out_lo != nullptr) {
    // decode exactly as dense_apply_tc_kernel addresses a stage: block (row tile, k-block) = [hi 8 KB | lo 8 KB],
    // element (r, k) at (k / 4) * 2048 + r * 16 + (k % 4) * 4 bytes
    const int64_t nkb = (n + 15) / 16;
    for (int32_t r = 0; r < n; ++r) {
      double hi = 0.0, lo = 0.0;
      for (int64_t k = 0; k < nkb * 16; ++k) {
        const size_t at = ((size_t)(r / 128) * nkb + k / 16) * 4096 + (size_t)((k % 16) / 4) * 512 + (size_t)(r % 128) * 4 + k % 4;
        const double xv = k < n ? x[k] : 1.0;  // padding columns must hold zeros: x = 1 there exposes a leak
        hi += (double)t[at] * xv;
        lo += (double)t[at + 2048] * xv;
      }
      out_hi[r] = hi;
      out_lo[r] = lo;
    }
  }
  return (int64_t)t.size();
}
