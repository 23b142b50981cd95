"""Reference-compatible training API: `weak_form`, `closure`, `weak_form_sequence`,
`assemble_u_init`, `rel_L2_error` with the signatures of the reference's `train_FEONet.py`.

The reference functions read module globals (DO_PRECOND, PRECOND, IDX_SOL, NUM_PTS, device,
FORCE, gparams['model'], DT).  Each class below holds that state as
attributes and exposes bound methods with the reference signatures, so a training script swaps

    weak_form, closure = feo.weak_form, feo.closure

and keeps its epoch loop unchanged.  Matrices passed per call are accepted for
signature compatibility; the device operator is built once from them and cached by identity.

`weak_form` returns materialised (LHS, RHS) with autograd support (sparse/dense applies run in
our kernels, the index arithmetic is literally the reference's).  `closure` is the hot path: the
network forward stays in PyTorch, the residual loss and its backward are one fused kernel each.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .functional import (DenseFn, DenseResidualLossFn, NsDenseResidualLossFn, ResidualLossFn, SeqDenseResidualLossFn, SeqResidualLossFn, SpmmFn, _TransposeCache,
                         precond_output)
from .operator import FEOperator, _dense_host, _is_identity, to_host_csr


def rel_L2_error(pred: torch.Tensor, true: torch.Tensor) -> torch.Tensor:
    """FEONet_steady_Navier-Stokes/train_FEONet.py:368-369."""
    return (torch.sum((true - pred) ** 2, dim=-1) / torch.sum(true ** 2, dim=-1)) ** 0.5


def sincos_forcing_grid(coeff: torch.Tensor, resol_in: int) -> torch.Tensor:
    """Input synthesis of `closure` (steady NS :337-345; Stokes :277-283): one fused kernel
    (`feo_sincos_forcing_grid`) on CUDA tensors; the reference's eager formula elsewhere (tests, CPU set-up)."""
    if coeff.is_cuda:
        import ctypes as C

        lib = L.load_library()
        c = coeff.detach().to(torch.float32).contiguous()
        out = torch.empty((c.shape[0], 2, resol_in, resol_in), dtype=torch.float32, device=c.device)
        with torch.cuda.device(c.device):
            L.check(lib.feo_sincos_forcing_grid(C.c_void_p(c.data_ptr()), c.shape[0], int(resol_in), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
    device = coeff.device
    m0, m1, n0, n1, n2, n3 = (coeff[:, [k]] for k in range(6))
    grid_x = torch.linspace(-1, 1, resol_in)
    g = torch.cartesian_prod(grid_x, grid_x).to(device)
    x, y = g[:, 0], g[:, 1]
    f = torch.stack([m0 * torch.sin(n0 * x + n1 * y), m1 * torch.cos(n2 * x + n3 * y)], dim=1)
    return f.reshape(-1, 2, resol_in, resol_in)


def _fold_dense(A, P) -> np.ndarray:
    """M = A @ P evaluated once at set-up (the reference recomputes it every call: steady NS :325,
    Stokes :264 even per sample).  fp64 accumulate, rounded once to fp32."""
    csr = to_host_csr(A)
    import scipy.sparse as sp

    n = csr[0].shape[0] - 1
    A64 = sp.csr_matrix((csr[2].astype(np.float64), csr[1], csr[0]), shape=(n, n))
    return np.asarray(A64 @ _dense_host(P).astype(np.float64), dtype=np.float32)


class _Base:
    def __init__(self, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        self._fcache = _TransposeCache()
        self._op: Optional[FEOperator] = None
        self._op_key = None
        self._op_refs = ()  # the keyed objects themselves: ids / data pointers cannot be recycled while they are held

    def _key(self, *objs):
        """Cache key of the device operator.  Identity (`is`) of the matrix objects plus, for tensors, their in-place version
        counter; the objects are kept alive in `_op_refs`, so a recycled id() or data_ptr() can never alias a stale operator.
        (numpy / scipy matrices edited in place are not detected -- build a new `SteadyNavierStokes` / `LinearStokes` then.)"""
        same = len(objs) == len(self._op_refs) and all(a is b for a, b in zip(objs, self._op_refs))
        self._op_refs_new = objs
        versions = tuple(o._version if isinstance(o, torch.Tensor) else 0 for o in objs)
        return (same, versions)

    def _key_matches(self, key) -> bool:
        return self._op is not None and key[0] and self._op_key is not None and key[1] == self._op_key[1]

    def _key_commit(self, key):
        self._op_refs = self._op_refs_new
        self._op_key = (True, key[1])

    def _run_model(self, model, coeff_f, value_f, resol_in):
        """Network forward exactly as the reference's `closure` dispatches it."""
        if self.model_name in ("Net2D", "UNetWithHead"):
            if self.force == "grf" and value_f is not None:
                return model(value_f.reshape(-1, 2, resol_in, resol_in))
            return model(sincos_forcing_grid(coeff_f, resol_in))
        return model(coeff_f).unsqueeze(1)


class LinearStokes(_Base):
    """Linear Stokes (square / square-with-hole): FEONet_Stokes_square/train_FEONet.py:261-301,
    FEONet-square-with-hole/train_FEONet.py:264-309.

    State mirrored from the reference's globals: DO_PRECOND, NUM_PTS, gparams['model'], FORCE."""

    def __init__(self, matrix=None, precond=None, do_precond: bool = False, model_name: str = "FCNN",
                 force: str = "sincos", hole_signature: bool = False, device=None, idx_sol=None, dof_positions=None):
        """idx_sol, dof_positions (optional; the reference's npz carries both, its Stokes `weak_form` uses neither): with them a
        structured-mesh operator in FEniCS' dof order is renumbered internally so that the un-preconditioned residual runs in the
        lattice kernels (reorder.py); tensors keep the caller's numbering."""
        super().__init__(device)
        self.idx_sol, self.dof_positions = idx_sol, dof_positions
        self.DO_PRECOND = bool(do_precond)
        self.model_name, self.force, self.hole_signature = model_name, force, hole_signature
        if matrix is not None:
            self._operator(matrix, precond)

    def _operator(self, matrix, precond) -> FEOperator:
        key = self._key(matrix, precond)
        if not self._key_matches(key):
            n = matrix.shape[0]
            if self.DO_PRECOND:
                if precond is None:
                    raise ValueError("do_precond is set but no preconditioner was given (precond=None)")
                P = _dense_host(precond)
                if _is_identity(P):
                    self._op = FEOperator(n, A=matrix, device=self.device)
                else:
                    self._op = FEOperator(n, A=matrix, dense_m=_fold_dense(matrix, P), dense_p=P, device=self.device)
            else:
                perm = None
                if self.dof_positions is not None and self.idx_sol is not None:
                    from .reorder import is_identity, lattice_permutation

                    perm = lattice_permutation(self.idx_sol, np.asarray(self.dof_positions))
                    if is_identity(perm):
                        perm = None
                # the renumbered operator carries idx_sol so that the planner can pair the velocity dofs
                self._op = FEOperator(n, A=matrix, idx_sol=self.idx_sol if perm is not None else None, dof_perm=perm, device=self.device)
                if perm is not None and self._op.plan != "lattice":
                    self._op = FEOperator(n, A=matrix, device=self.device)
            self._key_commit(key)
        return self._op

    @property
    def operator(self) -> FEOperator:
        return self._op

    def weak_form(self, coeff_u, load_vec_f, matrix, precond):
        op = self._operator(matrix, precond)
        u = coeff_u.squeeze(1) if coeff_u.dim() == 3 else coeff_u
        LHS = DenseFn.apply(u, op) if op.has_dense_m else SpmmFn.apply(u, op, L.FEO_MAT_A)
        return LHS, load_vec_f.to(self.device)

    def residual_loss(self, coeff_u, load_vec_f, matrix, precond):
        op = self._operator(matrix, precond)
        u = coeff_u.squeeze(1) if coeff_u.dim() == 3 else coeff_u
        F = load_vec_f.to(self.device)
        fn = DenseResidualLossFn if op.has_dense_m else ResidualLossFn
        return fn.apply(u, F, op, self._fcache)

    def closure(self, model, coeff_f, *args):
        # square: closure(model, coeff_f, load_vec_f, matrix, precond, resol_in)
        # hole:   closure(model, coeff_f, value_f, load_vec_f, matrix, precond, resol_in)
        if self.hole_signature or len(args) == 5:
            value_f, load_vec_f, matrix, precond, resol_in = args
        else:
            value_f = None
            load_vec_f, matrix, precond, resol_in = args
        pred = self._run_model(model, coeff_f, value_f, resol_in)
        loss = self.residual_loss(pred, load_vec_f, matrix, precond)
        op = self._op
        if self.DO_PRECOND and op.has_dense_p:
            return loss, precond_output(op, pred)
        return loss, pred


class SteadyNavierStokes(_Base):
    """Steady Navier-Stokes: FEONet_steady_Navier-Stokes/train_FEONet.py:301-365.

    State mirrored from the reference's globals: DO_PRECOND, PRECOND, IDX_SOL, NUM_PTS, FORCE,
    gparams['model']."""

    def __init__(self, A=None, B1=None, B2=None, idx_sol=None, do_precond: bool = False, precond=None,
                 model_name: str = "FCNN", force: str = "sincos", device=None, dof_positions=None):
        """dof_positions (optional): the [N, 2] coordinates of every global dof -- the `p` array of the reference's npz
        (`W.tabulate_dof_coordinates()`, assemble_fenics.py:125, :138).  With them an operator assembled on a structured mesh
        in FEniCS' own dof order is renumbered internally so that the lattice kernels apply (reorder.py); tensors keep the
        caller's numbering."""
        super().__init__(device)
        self.dof_positions = dof_positions
        self.DO_PRECOND = bool(do_precond)
        self.PRECOND = precond
        self.IDX_SOL = idx_sol
        self.model_name, self.force = model_name, force
        self._identity_precond = True
        self._op_conv = None  # (0, B1, B2): the convective part of the residual when PRECOND is a genuinely dense matrix
        if A is not None:
            self._operator(A, B1, B2, idx_sol)

    def _operator(self, A, B1, B2, idx_sol) -> FEOperator:
        key = self._key(A, B1, B2, self.PRECOND)
        if not self._key_matches(key):
            n = A.shape[0]
            kw = dict(A=A, B1=B1, B2=B2, idx_sol=idx_sol, ns_precond_branch=self.DO_PRECOND, device=self.device)
            self._identity_precond = True
            if self.DO_PRECOND and self.PRECOND is not None:
                P = _dense_host(self.PRECOND)
                if not _is_identity(P):  # the shipped script always uses eye(N) (:142, quirk 8)
                    self._identity_precond = False
                    kw.update(dense_m=_fold_dense(A, P), dense_p=P)
                    # the convective part of the dense-P residual runs through the fused sparse kernels of (0, B1, B2)
                    import scipy.sparse as sp

                    self._op_conv = FEOperator(n, A=sp.csr_matrix((n, n), dtype=np.float32), B1=B1, B2=B2, idx_sol=idx_sol,
                                               ns_precond_branch=True, device=self.device)
            perm = None
            if self.dof_positions is not None and self._identity_precond and idx_sol is not None:
                from .reorder import is_identity, lattice_permutation

                perm = lattice_permutation(idx_sol, np.asarray(self.dof_positions))
                if is_identity(perm):
                    perm = None
            self._op = FEOperator(n, dof_perm=perm, **kw)
            if perm is not None and self._op.plan != "lattice":  # the renumbering bought nothing: keep the caller's order
                self._op = FEOperator(n, **kw)
            self._key_commit(key)
        return self._op

    @property
    def operator(self) -> FEOperator:
        return self._op

    def weak_form(self, coeff_u, load_vec_f, A, B1, B2, idx_sol):
        op = self._operator(A, B1, B2, idx_sol)
        u_batch = coeff_u.squeeze(1) if coeff_u.dim() == 3 else coeff_u
        i, j, _ = idx_sol
        i = torch.as_tensor(np.asarray(i, dtype=np.int64), device=u_batch.device)
        j = torch.as_tensor(np.asarray(j, dtype=np.int64), device=u_batch.device)
        Bu1 = SpmmFn.apply(u_batch, op, L.FEO_MAT_B1)
        Bu2 = SpmmFn.apply(u_batch, op, L.FEO_MAT_B2)
        convection = torch.zeros_like(u_batch, memory_format=torch.contiguous_format)
        convection[:, i] += u_batch[:, i] * Bu1[:, i]
        convection[:, j] += u_batch[:, i] * Bu1[:, j]
        convection[:, i] += u_batch[:, j] * Bu2[:, i]
        convection[:, j] += u_batch[:, j] * Bu2[:, j]
        F = load_vec_f.to(self.device)
        if self.DO_PRECOND:
            LHS = DenseFn.apply(u_batch, op) if not self._identity_precond else SpmmFn.apply(u_batch, op, L.FEO_MAT_A)
            return LHS, F - convection
        return SpmmFn.apply(u_batch, op, L.FEO_MAT_A), -F + convection

    def residual_loss(self, coeff_u, load_vec_f, A, B1, B2, idx_sol):
        op = self._operator(A, B1, B2, idx_sol)
        u = coeff_u.squeeze(1) if coeff_u.dim() == 3 else coeff_u
        if not self._identity_precond:  # dense P with convection: sparse convective kernels + one tensor-core apply
            return NsDenseResidualLossFn.apply(u, load_vec_f.to(self.device), op, self._op_conv, self._fcache)
        return ResidualLossFn.apply(u, load_vec_f.to(self.device), op, self._fcache)

    def closure(self, model, coeff_f, f_values, load_vec_f, A, B1, B2, resol_in):
        pred = self._run_model(model, coeff_f, f_values, resol_in)
        loss = self.residual_loss(pred, load_vec_f, A, B1, B2, self.IDX_SOL)
        if self.DO_PRECOND and not self._identity_precond:
            return loss, precond_output(self._op, pred)
        return loss, pred  # P = I: (P @ pred^T)^T == pred


class TimeDependentStokes(_Base):
    """Time-dependent Stokes: FEONet_time_dep_Stokes/train_FEONet.py:323-406.

    State mirrored from the reference's globals: DO_PRECOND, IDX_SOL, NUM_PTS, DT, gparams['model']."""

    def __init__(self, S_mat=None, A_mat=None, idx_sol=None, dt: float = 0.1, do_precond: bool = False, precond=None,
                 model_name: str = "RNN", device=None):
        super().__init__(device)
        self.DO_PRECOND, self.DT, self.IDX_SOL = bool(do_precond), float(dt), idx_sol
        self.model_name = model_name
        if S_mat is not None:
            self._operator(S_mat, A_mat, precond, self.DT)

    def _operator(self, S_mat, A_mat, precond, dt) -> FEOperator:
        key = self._key(S_mat, A_mat, precond)
        key = (key[0] and getattr(self, "_op_dt", None) == float(dt), key[1])
        if not self._key_matches(key):
            n = S_mat.shape[0]
            kw = dict(A=A_mat, S=S_mat, idx_sol=self.IDX_SOL, dt=float(dt), device=self.device)
            self._op_dt = float(dt)
            if self.DO_PRECOND:
                if precond is None:
                    raise ValueError("do_precond is set but no preconditioner was given (precond=None)")
                P = _dense_host(precond)
                import scipy.sparse as sp

                s, a = to_host_csr(S_mat), to_host_csr(A_mat)
                S32 = sp.csr_matrix((s[2], s[1], s[0]), shape=(n, n))
                A32 = sp.csr_matrix((a[2], a[1], a[0]), shape=(n, n))
                sysm = (S32 + np.float32(dt) * A32).astype(np.float64)
                kw.update(dense_m=np.asarray(sysm @ P.astype(np.float64), dtype=np.float32), dense_p=P)
            self._op = FEOperator(n, **kw)
            self._key_commit(key)
        return self._op

    @property
    def operator(self) -> FEOperator:
        return self._op

    def assemble_u_init(self, init_x, init_y, idx_sol=None, num_pts=None, device=None):
        op = self._op
        if op is None:
            raise L.FeoError("assemble_u_init needs the operator (construct TimeDependentStokes with S_mat, A_mat)")
        u0T = op.assemble_u_init(init_x.to(self.device), init_y.to(self.device))
        return op.from_dof_major(u0T, init_x.shape[0])

    def weak_form_sequence(self, pred_seq, load_vec_f, S_mat, A_mat, precond, dt, u_init, do_precond):
        self.DO_PRECOND = bool(do_precond)
        op = self._operator(S_mat, A_mat, precond, dt)
        B, T, N = pred_seq.shape
        flat = pred_seq.reshape(B * T, N)
        LHS = (DenseFn.apply(flat, op) if op.has_dense_m else SpmmFn.apply(flat, op, L.FEO_MAT_M)).reshape(B, T, N)
        prev = torch.cat([u_init.unsqueeze(1), pred_seq[:, :-1, :]], dim=1).reshape(B * T, N)
        RHS = SpmmFn.apply(prev, op, L.FEO_MAT_S).reshape(B, T, N) + dt * load_vec_f.to(self.device).unsqueeze(1)
        return LHS, RHS

    def residual_loss(self, pred_seq, load_vec_f, S_mat, A_mat, precond, dt, u_init):
        op = self._operator(S_mat, A_mat, precond, dt)
        if op.has_dense_m:  # preconditioned: tensor-core apply fused with the subtraction of S prev + dt F and the reduction
            return SeqDenseResidualLossFn.apply(pred_seq, u_init, load_vec_f.to(self.device), op, float(dt))
        return SeqResidualLossFn.apply(pred_seq, u_init, load_vec_f.to(self.device), op)

    def closure(self, model, coeffs_init, init_value_x, init_value_y, load_vec_f, S_mat, A_mat, p, precond, dt, seq_len):
        op = self._operator(S_mat, A_mat, precond, self.DT)
        u_init = self.assemble_u_init(init_value_x, init_value_y)
        if self.model_name == "RNN":
            pred_seq = model(u_init, seq_len=seq_len)
        elif self.model_name == "UNet1D":
            pred_seq = model(torch.cat([u_init.unsqueeze(1), p], dim=1), seq_len=seq_len)
        else:
            pred_seq = model(coeffs_init, seq_len=seq_len)
        loss = self.residual_loss(pred_seq, load_vec_f, S_mat, A_mat, precond, self.DT, u_init)
        if self.DO_PRECOND and op.has_dense_p:
            return loss, precond_output(op, pred_seq)
        return loss, pred_seq
