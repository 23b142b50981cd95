"""TEST INFRASTRUCTURE -- CPU restatement (numpy/scipy) of the FEONet residual-loss hot path.

This is the parity ORACLE.  It is not part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
it, and only as the checker / the CPU baseline.  The product path
(`feonet_navier_stokes_b200`) never routes through this module and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pins are outputs of the reference's own functions
(`weak_form`, `closure`, `weak_form_sequence`, `assemble_u_init`), AST-extracted from
`/root/reference/*/train_FEONet.py` and executed unmodified in the authoring container by
`oracle/make_golden.py`; results are committed under `tests/golden/` and
`tests/test_oracle_golden.py` checks every function below against them.

Every function cites the reference lines it restates.  Matrices may be dense ndarrays or
scipy.sparse matrices (needed at ~1M dofs where the reference's dense N x N storage is
infeasible); `dtype` selects float32 (the reference's arithmetic) or float64 (a tighter
yardstick for the same formula).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

Array = np.ndarray


def _mat(K, dtype):
    if K is None:
        return None
    if sp.issparse(K):
        return K.astype(dtype).tocsr()
    return np.asarray(K, dtype=dtype)


def _apply(X: Array, K) -> Array:
    """X @ K.T for dense or sparse K (rows of X are samples)."""
    if sp.issparse(K):
        return np.asarray((K @ X.T).T)
    return X @ K.T


def _apply_t(X: Array, K) -> Array:
    """X @ K for dense or sparse K."""
    if sp.issparse(K):
        return np.asarray((K.T @ X.T).T)
    return X @ K


def _fold(M, P, dtype):
    """Operator fold M @ P (reference recomputes `A @ PRECOND` every call:
    FEONet_steady_Navier-Stokes/train_FEONet.py:325, FEONet_Stokes_square/train_FEONet.py:264)."""
    if P is None:
        return M
    P = np.asarray(P, dtype=dtype)
    if sp.issparse(M):
        return np.asarray(M @ P)
    return M @ P


# ----------------------------------------------------------------------------------------------
# A.1 linear Stokes (square / hole)
# ----------------------------------------------------------------------------------------------
def stokes_weak_form(coeff_u: Array, load_vec_f: Array, matrix, precond=None, do_precond: bool = False,
                     dtype=np.float32) -> Tuple[Array, Array]:
    """`weak_form` of FEONet_Stokes_square/train_FEONet.py:261-271 (identical at
    FEONet-square-with-hole/train_FEONet.py:264-274): LHS[b] = (matrix @ precond) @ u_b, RHS = F.
    coeff_u is [B,1,N] or [B,N]."""
    u = np.asarray(coeff_u, dtype=dtype).reshape(coeff_u.shape[0], -1)
    M = _fold(_mat(matrix, dtype), precond if do_precond else None, dtype)
    return _apply(u, M), np.asarray(load_vec_f, dtype=dtype)


def residual_loss(LHS: Array, RHS: Array) -> float:
    """Loss block of `closure` (FEONet_Stokes_square/train_FEONet.py:290-296,
    FEONet_steady_Navier-Stokes/train_FEONet.py:354-360): per-dof MSELoss(reduction='sum') over the
    batch, then summed over dofs == sum_{b,i} (LHS-RHS)^2."""
    d = LHS - RHS
    per_dof = np.sum(d * d, axis=0, dtype=d.dtype)
    return float(np.sum(per_dof, dtype=d.dtype))


def stokes_loss_and_grad(coeff_u, load_vec_f, matrix, precond=None, do_precond=False, dtype=np.float32):
    """loss and d loss / d coeff_u for A.1: grad = 2 (LHS-F) @ M."""
    u = np.asarray(coeff_u, dtype=dtype).reshape(coeff_u.shape[0], -1)
    M = _fold(_mat(matrix, dtype), precond if do_precond else None, dtype)
    r = _apply(u, M) - np.asarray(load_vec_f, dtype=dtype)
    return residual_loss(r, np.zeros_like(r)), 2.0 * _apply_t(r, M), r


def precond_output(coeff_u: Array, precond, do_precond: bool, dtype=np.float32) -> Array:
    """Second output of `closure`: (precond @ pred^T)^T (FEONet_Stokes_square/train_FEONet.py:298-301,
    FEONet_steady_Navier-Stokes/train_FEONet.py:362-365, time-dep :402-404)."""
    u = np.asarray(coeff_u, dtype=dtype)
    if not do_precond:
        return u
    return u @ np.asarray(precond, dtype=dtype).T


# ----------------------------------------------------------------------------------------------
# A.2 steady Navier-Stokes
# ----------------------------------------------------------------------------------------------
def ns_convection(u: Array, B1, B2, I: Sequence[int], J: Sequence[int]):
    """Nodal-product convection of FEONet_steady_Navier-Stokes/train_FEONet.py:308-322.
    Returns (convection, Bu1, Bu2)."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    Bu1 = _apply(u, B1)  # :308
    Bu2 = _apply(u, B2)  # :309
    conv = np.zeros_like(u)  # :314
    conv[:, I] += u[:, I] * Bu1[:, I]  # :317
    conv[:, J] += u[:, I] * Bu1[:, J]  # :318
    conv[:, I] += u[:, J] * Bu2[:, I]  # :321
    conv[:, J] += u[:, J] * Bu2[:, J]  # :322
    return conv, Bu1, Bu2


def ns_weak_form(coeff_u, load_vec_f, A, B1, B2, I, J, do_precond: bool, precond=None, dtype=np.float32):
    """`weak_form` of FEONet_steady_Navier-Stokes/train_FEONet.py:301-332, both branches:
    precond (:324-326)  LHS = u (A P)^T, RHS = F - conv ;  else (:328-330) LHS = u A^T, RHS = -F + conv."""
    u = np.asarray(coeff_u, dtype=dtype).reshape(coeff_u.shape[0], -1)
    F = np.asarray(load_vec_f, dtype=dtype)
    A_, B1_, B2_ = _mat(A, dtype), _mat(B1, dtype), _mat(B2, dtype)
    conv, _, _ = ns_convection(u, B1_, B2_, I, J)
    if do_precond:
        M = _fold(A_, precond, dtype)
        return _apply(u, M), F - conv
    return _apply(u, A_), -F + conv


def ns_loss_and_grad(coeff_u, load_vec_f, A, B1, B2, I, J, do_precond: bool, precond=None, dtype=np.float32):
    """Loss and its gradient for A.2 (equal to autograd through
    FEONet_steady_Navier-Stokes/train_FEONet.py:301-360):
    G = 2r ; grad = G M + s[(d1.G) B1 + (d2.G) B2 + E1^T(Bu1.G) + E2^T(Bu2.G)], s=+1 precond / -1 else."""
    u = np.asarray(coeff_u, dtype=dtype).reshape(coeff_u.shape[0], -1)
    F = np.asarray(load_vec_f, dtype=dtype)
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    A_, B1_, B2_ = _mat(A, dtype), _mat(B1, dtype), _mat(B2, dtype)
    conv, Bu1, Bu2 = ns_convection(u, B1_, B2_, I, J)
    if do_precond:
        M, s = _fold(A_, precond, dtype), 1.0
        r = _apply(u, M) - (F - conv)
    else:
        M, s = A_, -1.0
        r = _apply(u, M) - (-F + conv)
    loss = residual_loss(r, np.zeros_like(r))
    G = 2.0 * r
    d1 = np.zeros_like(u)
    d2 = np.zeros_like(u)
    d1[:, I] = u[:, I]
    d1[:, J] = u[:, I]
    d2[:, I] = u[:, J]
    d2[:, J] = u[:, J]
    grad = _apply_t(G, M) + s * (_apply_t(d1 * G, B1_) + _apply_t(d2 * G, B2_))
    w1, w2 = Bu1 * G, Bu2 * G
    grad[:, I] += s * (w1[:, I] + w1[:, J])
    grad[:, J] += s * (w2[:, I] + w2[:, J])
    return loss, grad.astype(dtype), r


# ----------------------------------------------------------------------------------------------
# A.3 time-dependent Stokes
# ----------------------------------------------------------------------------------------------
def assemble_u_init(init_x: Array, init_y: Array, I, J, num_pts: int, dtype=np.float32) -> Array:
    """`assemble_u_init` of FEONet_time_dep_Stokes/train_FEONet.py:323-335."""
    init_x = np.asarray(init_x, dtype=dtype).reshape(init_x.shape[0], -1)
    init_y = np.asarray(init_y, dtype=dtype).reshape(init_y.shape[0], -1)
    u0 = np.zeros((init_x.shape[0], num_pts), dtype=dtype)
    u0[:, np.asarray(I, dtype=np.int64)] = init_x
    u0[:, np.asarray(J, dtype=np.int64)] = init_y
    return u0


def seq_weak_form(pred_seq, load_vec_f, S_mat, A_mat, precond, dt, u_init, do_precond, dtype=np.float32):
    """`weak_form_sequence` of FEONet_time_dep_Stokes/train_FEONet.py:343-362:
    M=(S+dt A)[P]; LHS = pred M^T; RHS_t = prev_t S^T + dt F, prev_0 = u_init, prev_t = pred[:,t-1]."""
    X = np.asarray(pred_seq, dtype=dtype)
    B, T, N = X.shape
    F = np.asarray(load_vec_f, dtype=dtype)
    S_, A_ = _mat(S_mat, dtype), _mat(A_mat, dtype)
    sysm = S_ + dtype(dt) * A_
    M = _fold(sysm.tocsr() if sp.issparse(sysm) else sysm, precond if do_precond else None, dtype)
    LHS = _apply(X.reshape(B * T, N), M).reshape(B, T, N)
    prev = np.concatenate([np.asarray(u_init, dtype=dtype)[:, None, :], X[:, :-1, :]], axis=1)
    RHS = _apply(prev.reshape(B * T, N), S_).reshape(B, T, N) + dtype(dt) * F[:, None, :]
    return LHS, RHS, M, S_


def seq_loss_and_grad(pred_seq, load_vec_f, S_mat, A_mat, precond, dt, u_init, do_precond, dtype=np.float32):
    """Loss of FEONet_time_dep_Stokes/train_FEONet.py:398-400, `(resid**2).sum(dim=(0,2)).mean()`,
    and its gradient g_t = (2/T)[r_t M - r_{t+1} S] (u_prev is not detached)."""
    LHS, RHS, M, S_ = seq_weak_form(pred_seq, load_vec_f, S_mat, A_mat, precond, dt, u_init, do_precond, dtype)
    r = LHS - RHS
    B, T, N = r.shape
    per_t = np.sum(r * r, axis=(0, 2), dtype=r.dtype)
    loss = float(np.mean(per_t, dtype=r.dtype))
    g = _apply_t(r.reshape(B * T, N), M).reshape(B, T, N)
    gs = _apply_t(r.reshape(B * T, N), S_).reshape(B, T, N)
    g[:, :-1, :] -= gs[:, 1:, :]
    return loss, (dtype(2.0 / T) * g).astype(dtype), r


def rel_L2_error(pred: Array, true: Array) -> Array:
    """`rel_L2_error` (FEONet_steady_Navier-Stokes/train_FEONet.py:368-369)."""
    return (np.sum((true - pred) ** 2, axis=-1) / np.sum(true ** 2, axis=-1)) ** 0.5


def sincos_forcing_grid(coeff_f: Array, resol_in: int, dtype=np.float32) -> Array:
    """Input synthesis in `closure` (FEONet_steady_Navier-Stokes/train_FEONet.py:337-345):
    value_f[b] = [m0 sin(n0 x + n1 y), m1 cos(n2 x + n3 y)] on cartesian_prod(linspace(-1,1,r))."""
    c = np.asarray(coeff_f, dtype=dtype)
    g = np.linspace(-1, 1, resol_in, dtype=dtype)
    x = np.repeat(g, resol_in)[None, :]
    y = np.tile(g, resol_in)[None, :]
    f1 = c[:, [0]] * np.sin(c[:, [2]] * x + c[:, [3]] * y)
    f2 = c[:, [1]] * np.cos(c[:, [4]] * x + c[:, [5]] * y)
    return np.stack([f1, f2], axis=1).reshape(-1, 2, resol_in, resol_in).astype(dtype)


# ----------------------------------------------------------------------------------------------
# Multi-threaded CPU port used ONLY as bench.py's cpu_baseline / --impl reference leg
# ----------------------------------------------------------------------------------------------
class TorchCpuSteadyNS:
    """The A.2 formula (ns_loss_and_grad above, i.e. FEONet_steady_Navier-Stokes/train_FEONet.py:301-360
    and its autograd) on the host cores with torch CPU sparse-CSR matrices, fp32.

    The reference itself stores A, B1, B2 as dense N x N tensors (:293-295), which is infeasible at
    ~1M dofs (4 TB per matrix); this is the same arithmetic with the zeros skipped, so it is a
    generous CPU baseline ("port"), not the reference's dense timing."""

    def __init__(self, A, B1, B2, I, J, do_precond: bool, threads: Optional[int] = None):
        import torch

        if threads:
            torch.set_num_threads(int(threads))
        self.torch = torch
        self.threads = torch.get_num_threads()

        def csr(K):
            K = sp.csr_matrix(K).astype(np.float32)
            return torch.sparse_csr_tensor(torch.from_numpy(K.indptr.astype(np.int64)),
                                           torch.from_numpy(K.indices.astype(np.int64)),
                                           torch.from_numpy(K.data), size=K.shape)

        self.A, self.B1, self.B2 = csr(A), csr(B1), csr(B2)
        self.AT, self.B1T, self.B2T = csr(sp.csr_matrix(A).T), csr(sp.csr_matrix(B1).T), csr(sp.csr_matrix(B2).T)
        self.I = torch.as_tensor(np.asarray(I, dtype=np.int64))
        self.J = torch.as_tensor(np.asarray(J, dtype=np.int64))
        self.do_precond = bool(do_precond)

    def loss_and_grad(self, alpha, F):
        """alpha, F: [B,N] float32 (numpy or torch). Returns (loss float, grad [B,N] torch)."""
        torch = self.torch
        X = torch.as_tensor(alpha).t().contiguous()  # [N,B]
        Ft = torch.as_tensor(F).t()
        I, J = self.I, self.J
        Bu1, Bu2, LHS = self.B1 @ X, self.B2 @ X, self.A @ X
        conv = torch.zeros_like(X)
        conv[I] = X[I] * Bu1[I] + X[J] * Bu2[I]
        conv[J] = X[I] * Bu1[J] + X[J] * Bu2[J]
        if self.do_precond:
            r, s = LHS - (Ft - conv), 1.0
        else:
            r, s = LHS - (-Ft + conv), -1.0
        loss = float((r * r).sum())
        G = 2.0 * r
        d1, d2 = torch.zeros_like(X), torch.zeros_like(X)
        d1[I], d1[J], d2[I], d2[J] = X[I], X[I], X[J], X[J]
        grad = self.AT @ G + s * (self.B1T @ (d1 * G) + self.B2T @ (d2 * G))
        w1, w2 = Bu1 * G, Bu2 * G
        grad[I] += s * (w1[I] + w1[J])
        grad[J] += s * (w2[I] + w2[J])
        return loss, grad.t()


# ----------------------------------------------------------------------------------------------
# The reference's OWN execution plan on the host cores (bench.py --configs): dense operators, eager torch ops, the
# per-dof Python loss loop into a CPU tensor, autograd backward -- timed beside the GPU path at cfg1-4.  Restated
# (not imported: /root/reference does not exist on the GPU box); tests/test_oracle_golden.py pins it to the goldens the
# reference's unmodified functions produced, tools/time_reference_here.py times both side by side in the authoring container.
# ----------------------------------------------------------------------------------------------
class TorchReferenceLoops:
    """Step = loss forward + `loss.backward()` to d loss / d alpha, exactly as the reference schedules it."""

    def __init__(self, threads: Optional[int] = None):
        import torch

        if threads:
            torch.set_num_threads(int(threads))
        self.torch = torch
        self.threads = torch.get_num_threads()
        self.mse_sum = torch.nn.MSELoss(reduction="sum")

    def _per_dof_loss(self, lhs, rhs):
        """FEONet_steady_Navier-Stokes/train_FEONet.py:354-360 (same block in every variant): one MSE(sum) per dof,
        written into a CPU tensor by a Python loop, then summed."""
        torch = self.torch
        n = lhs.shape[1]
        per_dof = torch.zeros((n,))
        for d in range(n):
            per_dof[d] = self.mse_sum(lhs[:, d], rhs[:, d])
        return torch.sum(per_dof)

    def steady_ns_step(self, alpha, F, A, B1, B2, I, J, do_precond: bool, precond=None):
        """FEONet_steady_Navier-Stokes/train_FEONet.py:301-332 + :351-360 + :463; alpha, F [B,N], dense fp32 matrices."""
        torch = self.torch
        a = torch.as_tensor(alpha).clone().unsqueeze(1).requires_grad_(True)  # [B,1,N] as the networks emit it
        Ft, At, B1t, B2t = (torch.as_tensor(x) for x in (F, A, B1, B2))
        I, J = list(map(int, I)), list(map(int, J))
        u = a.squeeze(1)
        s1, s2 = u @ B1t.T, u @ B2t.T
        conv = torch.zeros_like(u)
        conv[:, I] += u[:, I] * s1[:, I]
        conv[:, J] += u[:, I] * s1[:, J]
        conv[:, I] += u[:, J] * s2[:, I]
        conv[:, J] += u[:, J] * s2[:, J]
        if do_precond:
            Pt = torch.eye(At.shape[0]) if precond is None else torch.as_tensor(precond)
            lhs, rhs = u @ (At @ Pt).T, Ft - conv  # the N^3 fold runs on every call (:325)
        else:
            lhs, rhs = u @ At.T, -Ft + conv
        loss = self._per_dof_loss(lhs, rhs)
        loss.backward()
        return float(loss), a.grad.squeeze(1)

    def linear_stokes_step(self, alpha, F, matrix, precond, do_precond: bool):
        """FEONet_Stokes_square/train_FEONet.py:261-271 + :290-296 (hole variant :264-274): one (matrix @ precond).mm per
        SAMPLE inside a list comprehension."""
        torch = self.torch
        a = torch.as_tensor(alpha).clone().unsqueeze(1).requires_grad_(True)
        Ft, Mt = torch.as_tensor(F), torch.as_tensor(matrix)
        cols = a.transpose(1, 2)  # [B,N,1]
        if do_precond:
            Pt = torch.as_tensor(precond)
            lhs = torch.stack([(Mt @ Pt).mm(c) for c in cols])
        else:
            lhs = torch.stack([Mt.mm(c) for c in cols])
        lhs = torch.sum(lhs, dim=-1)
        loss = self._per_dof_loss(lhs, Ft)
        loss.backward()
        return float(loss), a.grad.squeeze(1)
