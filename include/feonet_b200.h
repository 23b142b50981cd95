/*
 * feonet_b200.h -- C ABI of the B200-native FEONet residual-loss library (libfeonet_b200.so).
 *
 * The reference (haltmayermarc/FEONet_Navier_Stokes) has no FFI: its hot path is the set of
 * Python functions `weak_form` / `closure` / `weak_form_sequence` / `assemble_u_init` inside
 * the four `train_FEONet.py` scripts, built from torch eager ops.  Each entry point below names the reference
 * lines it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types; every call returns an int status
 *     (FEO_OK or a negative FEO_ERR_*), never throws; `feo_last_error_string()` explains.
 *   - `feo_op_create` takes HOST arrays (one-off set-up) and owns the device copies it makes.
 *     Every other pointer is a DEVICE pointer supplied by the caller (torch-owned memory); the
 *     library never allocates per-step memory.  All compute calls are asynchronous on `stream`
 *     (a cudaStream_t passed as void*), e.g. torch.cuda.current_stream().cuda_stream.
 *   - Handles are not thread-safe; one handle per device.
 *
 * Device layout ("dof-major"): a batch of coefficient vectors alpha[B,N] is stored as
 *   XT[d * ldb + b],  d = dof in [0,N), b = sample in [0,B),  ldb % 4 == 0, ldb >= B, base 16-B aligned,
 * i.e. the transpose of the reference's row-major [B,N] tensor (in torch: a [B,N] tensor with
 * strides (1, ldb)).  `feo_transpose` converts either way.  Sequences [B,T,N] use pseudo-samples
 * j = b*T + t:  XT[d * ldj + j].
 */
#ifndef FEONET_B200_H
#define FEONET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEO_OK 0
#define FEO_ERR_INVALID_ARGUMENT (-1)
#define FEO_ERR_CUDA (-2)
#define FEO_ERR_UNSUPPORTED (-3)
#define FEO_ERR_OUT_OF_MEMORY (-4)

#define FEO_ABI_VERSION 2

/* Which stored matrix a generic apply uses. */
enum feo_matrix_id { FEO_MAT_A = 0, FEO_MAT_B1 = 1, FEO_MAT_B2 = 2, FEO_MAT_S = 3, FEO_MAT_M = 4 /* S + dt*A */ };
enum feo_plan_id { FEO_PLAN_NONE = 0, FEO_PLAN_TILE = 1, FEO_PLAN_PATCH = 2, FEO_PLAN_LATTICE = 3 };

/* Dense operators kept by a handle. */
enum feo_dense_id { FEO_DENSE_M = 0 /* LHS operator, e.g. A@P */, FEO_DENSE_MT = 1, FEO_DENSE_P = 2 /* preconditioner */ };

/* Host-side CSR matrix (int32 indices, fp32 values). rowptr==NULL means "absent". */
typedef struct feo_csr {
  const int32_t* rowptr; /* [n+1] */
  const int32_t* col;    /* [nnz], need not be sorted; duplicates are summed */
  const float* val;      /* [nnz] */
} feo_csr;

/*
 * Operator description.  Replaces the module-level operator state of the reference scripts:
 *   A, B1, B2           FEONet_steady_Navier-Stokes/train_FEONet.py:82-85, 293-295 (dense there)
 *   matrix              FEONet_Stokes_square/train_FEONet.py:255 ; hole :258
 *   S, A, dt            FEONet_time_dep_Stokes/train_FEONet.py:316-318, DT
 *   idx_sol             IDX_SOL = mesh['idx_sol'] (steady NS :81) -- opaque int lists I, J
 *   PRECOND / precond   steady NS :137-168 ; Stokes_square :123-143
 */
typedef struct feo_operator_desc {
  int32_t abi_version; /* FEO_ABI_VERSION */
  int32_t n;           /* NUM_PTS */
  feo_csr A, B1, B2, S;
  int32_t n_u;          /* len(idx_sol[0]) == len(idx_sol[1]); 0 if no index lists */
  const int32_t* idx_i; /* idx_sol[0] */
  const int32_t* idx_j; /* idx_sol[1] */
  /* steady-NS sign branch (FEONet_steady_Navier-Stokes/train_FEONet.py:324-330):
   *   1: r = M a - F + c   (DO_PRECOND)      0: r = M a + F - c   (else)            */
  int32_t ns_precond_branch;
  float dt;                /* time-dependent variant: M = S + dt*A; 0 otherwise */
  const float* dense_m;    /* host [n,n] row-major LHS operator (A@P, (S+dt A)@P) or NULL */
  const float* dense_p;    /* host [n,n] row-major preconditioner for the output map or NULL */
} feo_operator_desc;

typedef struct feo_operator* feo_handle_t;

typedef struct feo_op_info {
  int32_t n, n_u, has_conv, has_seq, has_dense_m, has_dense_p;
  int64_t nnz_a, nnz_b1, nnz_b2, nnz_s, nnz_union; /* stored (value != 0) entries */
  int32_t n_tiles_fwd, n_tiles_bwd, max_row_nnz; /* tiles of the fused forward / backward kernels */
  int64_t device_bytes; /* device memory owned by the handle */
} feo_op_info;

/* Library / set-up ------------------------------------------------------------------------- */
int feo_abi_version(void);
const char* feo_last_error_string(void);

/* Builds device CSR (+ transposes), the fused union pattern {col, a, b1, b2}, partner lookups
 * from (idx_i, idx_j) and the tile plans (staging boxes + per-warp operator streams) the fused
 * kernels walk.  One-off; synchronous. */
int feo_op_create(const feo_operator_desc* desc, feo_handle_t* out);
int feo_op_destroy(feo_handle_t h);
int feo_op_get_info(feo_handle_t h, feo_op_info* info);

/* Bytes of scratch `feo_*_fwd` needs for a batch of B samples (T pseudo-steps, 1 if none). */
size_t feo_workspace_bytes(feo_handle_t h, int32_t B, int32_t T);

/* Layout -------------------------------------------------------------------------------------
 * dst[c * dst_ld + r] = src[r * src_ld + c] for r < rows, c < cols (fp32).  Converts the
 * reference's row-major [B,N] tensors to dof-major and back.  If dst_row_map != NULL the
 * destination row is dst_row_map[c] (used by feo_assemble_u_init). */
int feo_transpose(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                  const int32_t* dst_row_map, void* stream);
/* The same with the SOURCE row gathered through a map: dst[c * dst_ld + r] = src[src_row_map[r] * src_ld + c].
 * feo_transpose's dst_row_map (way in) and this (way out) fold a dof permutation into the layout pass the row-major
 * boundary runs anyway -- how an operator in FEniCS' dof order reaches the lattice kernels (reorder.py). */
int feo_transpose_gather(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int32_t rows, int32_t cols,
                         const int32_t* src_row_map, void* stream);
/* Which plan the fused residual kernels of this handle run: FEO_PLAN_* (none when the handle has no sparse A). */
int feo_op_plan(feo_handle_t h);

/* Fused sparse residual loss -----------------------------------------------------------------
 * Replaces weak_form + the per-dof loss loop of closure:
 *   steady NS     FEONet_steady_Navier-Stokes/train_FEONet.py:301-332 and :351-360
 *   linear Stokes FEONet_Stokes_square/train_FEONet.py:261-271 and :290-296 (DO_PRECOND False),
 *                 FEONet-square-with-hole/train_FEONet.py:264-274 and :293-299
 * r = A a -/+ (F - c), loss = sum r^2.  c (convection) is present iff the handle has B1,B2,I,J.
 *   alphaT, fT  [N][ldb] dof-major inputs
 *   loss_out    device fp32 scalar
 *   rT          [N][ldb] residual, saved for backward (NULL: loss only)
 *   workspace   >= feo_workspace_bytes(h, B, 1) bytes
 * Samples are processed in slabs of 64; rows of alphaT beyond B (up to ldb) may hold anything.
 */
int feo_residual_fwd(feo_handle_t h, const float* alphaT, const float* fT, int64_t ldb, int32_t B,
                     float* loss_out, float* rT, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above (autograd of steady NS :463 / Stokes :396):
 *   gradT = 2 * (*grad_loss) * [A^T r + s(B1^T(d1.r) + B2^T(d2.r) + e)],  grad_loss NULL means 1.
 * e (the derivative through the advecting velocity) is recomputed from
 * alphaT and rT; alphaT may be NULL when the handle has no convection. */
int feo_residual_bwd(feo_handle_t h, const float* alphaT, const float* rT, const float* grad_loss,
                     float* gradT, int64_t ldb, int32_t B, void* stream);

/* Generic sparse apply ------------------------------------------------------------------------
 * YT = scale * op(K) XT (+ YT if accumulate), K = stored matrix `which`, op = transpose?K^T:K.
 * Used for the materialised (LHS, RHS) tensors `weak_form` returns and their VJPs
 * (steady NS :308-309,:325,:329 ; Stokes :264-267). */
int feo_spmm(feo_handle_t h, int32_t which, int32_t transpose, const float* XT, float* YT, int64_t ldb,
             int32_t B, float scale, int32_t accumulate, void* stream);

/* Dense operator path (preconditioned variants) ---------------------------------------------
 * CT = scale * D XT  with D = stored dense matrix `which` ([n,n]); epilogue:
 *   sub != NULL : CT = scale * D XT - sub          (residual r = (A P) a - F, Stokes :264)
 *   loss_out != NULL : also loss = sum CT^2         (Stokes :290-296)
 *   scale_dev != NULL : multiplies scale by *scale_dev (backward: 2 * grad_loss * M^T r)
 * FEO_DENSE_P gives the output map u = P a of closure (Stokes :298-301, steady NS :362-363). */
int feo_dense_apply(feo_handle_t h, int32_t which, const float* XT, float* CT, int64_t ldb, int32_t B,
                    float scale, const float* scale_dev, const float* sub, float* loss_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Time-dependent Stokes ------------------------------------------------------------------------
 * Replaces weak_form_sequence + the loss of closure (FEONet_time_dep_Stokes/train_FEONet.py:343-362, :398-400)
 * for the un-preconditioned operator M = S + dt*A (sparse):
 *   r[:,t,:] = M x_t - S prev_t - dt F,  prev_0 = u_init, prev_t = x_{t-1};  loss = (1/T) sum r^2
 *   predT [N][ldj] with pseudo-sample j = b*T + t ; u0T, fT [N][ldb] ; rT [N][ldj]. */
int feo_seq_fwd(feo_handle_t h, const float* predT, const float* u0T, const float* fT, int64_t ldj, int64_t ldb,
                int32_t B, int32_t T, float* loss_out, float* rT, void* workspace, size_t workspace_bytes,
                void* stream);
/* gradT[:,t] = (2/T) * (*grad_loss) * [M^T r_t - S^T r_{t+1}] (second term absent for t = T-1). */
int feo_seq_bwd(feo_handle_t h, const float* rT, const float* grad_loss, float* gradT, int64_t ldj, int32_t B,
                int32_t T, void* stream);

/* assemble_u_init (FEONet_time_dep_Stokes/train_FEONet.py:323-335): u0T[I[k]][b] = init_x[b][k],
 * u0T[J[k]][b] = init_y[b][k], zero elsewhere.  init_x/init_y are row-major [B, n_u]. */
int feo_assemble_u_init(feo_handle_t h, const float* init_x, const float* init_y, float* u0T, int64_t ldb,
                        int32_t B, void* stream);

/* Input synthesis of `closure` (FEONet_steady_Navier-Stokes/train_FEONet.py:337-345, FEONet_Stokes_square
 * :277-283): value_f[b] = [m0 sin(n0 x + n1 y), m1 cos(n2 x + n3 y)] on cartesian_prod(linspace(-1, 1, resol_in)),
 * coeff_f row-major [B, 6] = (m0, m1, n0, n1, n2, n3), value_f row-major [B, 2, resol_in, resol_in]. */
int feo_sincos_forcing_grid(const float* coeff_f, int32_t B, int32_t resol_in, float* value_f, void* stream);

/* Elementwise sum-of-squares of (x - y) over an [n][ldb] dof-major pair, first B columns:
 * the loss block of closure applied to materialised (LHS, RHS). y may be NULL. */
int feo_sq_diff_sum(const float* xT, const float* yT, int32_t n, int64_t ldb, int32_t B, float scale,
                    float* loss_out, void* workspace, size_t workspace_bytes, void* stream);

/* Test hook (HOST ONLY, no CUDA calls) ----------------------------------------------------------
 * Exercised by the CPU test-suite to validate the set-up code without a GPU; never called by the
 * product path.  Builds the tile plan of the fused residual kernels and replays its staging boxes and
 * per-warp streams in fp64 for one sample, decoding them as the kernels do.  forward: in0 = alpha,
 * in1 = f, out = r; backward: in0 = r, in1 = alpha, out = grad / (2 g).  in0 == NULL: statistics only.
 * stats[0..7] = {n_tiles, max_lines, total_lines, n_boxes, n_items, real_entries, slot_entries, stream_words}. */
int feo_debug_tile_replay(const feo_operator_desc* desc, int32_t backward, int32_t max_lines, int32_t warps,
                          const double* in0, const double* in1, double* out, int64_t* stats);

/* Test hook (HOST ONLY): builds the PATCH plan of the fused residual kernels (second-generation plan: one warp evaluates a
 * patch of up to 4 velocity nodes + 1 single dof from one gather per column; lines stay resident in a shared-memory
 * pool from round to round) and replays its line loads, pool slots and operator streams in fp64 for one sample, as the
 * kernels decode them.  Arguments as feo_debug_tile_replay; warps / pool_lines / seg_rounds <= 0: defaults.
 * Returns FEO_ERR_UNSUPPORTED when the operator does not fit the patch model (the tile plan is used then).
 * stats[0..9] = {applicable, n_patches, n_rounds, n_segments, n_loads, n_gathers, real_entries, slot_entries,
 *               stream_units, max_union_lines}. */
int feo_debug_patch_replay(const feo_operator_desc* desc, int32_t backward, int32_t warps, int32_t pool_lines,
                           int32_t seg_rounds, const double* in0, const double* in1, double* out, int64_t* stats);

/* Test hook (HOST ONLY): builds the LATTICE plan of the fused residual kernels (third-generation plan for structured
 * right-diagonal P2-P1 meshes in lattice order: one warp evaluates a cell -- a vertex node, its three edge nodes and the
 * vertex's pressure dof -- from one gather per window line, coefficients per cell class as kernel parameters) and replays
 * the generated cell bodies over the class tables in fp64 for one sample.  Arguments as feo_debug_tile_replay.
 * Returns FEO_ERR_UNSUPPORTED when the operator is not such a lattice (the tile plan is used then).
 * stats[0..5] = {applicable, n (cells per side), classes forward, classes backward, coefficients forward, coefficients backward}. */
int feo_debug_lattice_replay(const feo_operator_desc* desc, int32_t backward, const double* in0, const double* in1,
                             double* out, int64_t* stats);

/* Test hook (HOST ONLY): splits a dense [n,n] row-major operator into the TF32 hi/lo operand tiles of the
 * tensor-core apply (feo_dense_apply) and replays them as the kernel reads them -- per 128-row tile and 16-column
 * k-block, K-major core matrices -- for one vector x: out_hi = sum hi*x, out_lo = sum lo*x in fp64, so that
 * out_hi + out_lo = D x (or D^T x) up to the 2^-22 split error.  Returns the number of floats of the tiled array. */
int64_t feo_debug_dense_split_replay(const float* dense, int32_t n, int32_t transposed, const double* x,
                                     double* out_hi, double* out_lo);

#ifdef __cplusplus
}
#endif
#endif /* FEONET_B200_H */
