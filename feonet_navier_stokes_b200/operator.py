"""Device-resident FEM operator handle (thin wrapper over the C ABI).

The reference keeps A, B1, B2, S and the preconditioner as dense N x N fp32 tensors and
applies them with dense `@`.  Here they are converted ONCE to CSR
(entries kept iff value != 0 after the fp32 cast -- quirk 10) and handed to
`feo_op_create`, which builds the fused walk plan on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def ceil4(x: int) -> int:
    return (int(x) + 3) // 4 * 4


def to_host_csr(K) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """dense torch/numpy, scipy.sparse or torch sparse-CSR -> (rowptr i32, col i32, val f32)."""
    if K is None:
        return None
    try:
        import scipy.sparse as sp
    except Exception:  # pragma: no cover
        sp = None
    if sp is not None and sp.issparse(K):
        K = K.tocsr()
        val = K.data.astype(np.float32)
        keep = val != 0
        if not keep.all():
            K = sp.csr_matrix((val, K.indices, K.indptr), shape=K.shape)
            K.eliminate_zeros()
            val = K.data
        return (np.ascontiguousarray(K.indptr, dtype=np.int32), np.ascontiguousarray(K.indices, dtype=np.int32),
                np.ascontiguousarray(val, dtype=np.float32))
    if isinstance(K, torch.Tensor):
        if K.layout == torch.sparse_csr:
            return (K.crow_indices().cpu().numpy().astype(np.int32), K.col_indices().cpu().numpy().astype(np.int32),
                    K.values().cpu().numpy().astype(np.float32))
        K = K.detach().to("cpu").numpy()
    K = np.asarray(K).astype(np.float32)  # same fp64 -> fp32 cast as `.float()` in the reference
    if K.ndim != 2 or K.shape[0] != K.shape[1]:
        raise ValueError("operator matrices must be square 2-D")
    rows, cols = np.nonzero(K)
    rowptr = np.zeros(K.shape[0] + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=K.shape[0]), out=rowptr[1:])
    return rowptr, cols.astype(np.int32), np.ascontiguousarray(K[rows, cols], dtype=np.float32)


def _dense_host(M) -> Optional[np.ndarray]:
    if M is None:
        return None
    if isinstance(M, torch.Tensor):
        M = M.detach().to("cpu").numpy()
    return np.ascontiguousarray(np.asarray(M), dtype=np.float32)


def _is_identity(P: np.ndarray) -> bool:
    n = P.shape[0]
    return P.shape == (n, n) and np.count_nonzero(P) == n and bool(np.all(np.diagonal(P) == 1.0))


def build_desc(n, A=None, B1=None, B2=None, S=None, idx_i=None, idx_j=None, ns_precond_branch=False, dt=0.0,
               dense_m=None, dense_p=None):
    """Return (FeoOperatorDesc, keepalive) for feo_op_create / the debug hooks."""
    keep = []
    desc = L.FeoOperatorDesc()
    desc.abi_version = L.FEO_ABI_VERSION
    desc.n = int(n)

    def csr_field(K):
        c = L.FeoCsr()
        t = to_host_csr(K)
        if t is not None:
            if t[0].shape[0] != n + 1:
                raise ValueError("matrix size does not match n")
            keep.extend(t)
            c.rowptr = t[0].ctypes.data_as(L.i32p)
            c.col = t[1].ctypes.data_as(L.i32p)
            c.val = t[2].ctypes.data_as(L.f32p)
        return c

    desc.A, desc.B1, desc.B2, desc.S = csr_field(A), csr_field(B1), csr_field(B2), csr_field(S)
    if idx_i is not None:
        ii = np.ascontiguousarray(np.asarray(idx_i, dtype=np.int64).astype(np.int32))
        jj = np.ascontiguousarray(np.asarray(idx_j, dtype=np.int64).astype(np.int32))
        if ii.shape != jj.shape:
            raise ValueError("idx_sol[0] and idx_sol[1] must have equal length")
        keep += [ii, jj]
        desc.n_u = int(ii.size)
        desc.idx_i = ii.ctypes.data_as(L.i32p)
        desc.idx_j = jj.ctypes.data_as(L.i32p)
    desc.ns_precond_branch = 1 if ns_precond_branch else 0
    desc.dt = float(dt)
    for name, M in (("dense_m", dense_m), ("dense_p", dense_p)):
        Mh = _dense_host(M)
        if Mh is not None:
            if Mh.shape != (n, n):
                raise ValueError(f"{name} must be [n,n]")
            keep.append(Mh)
            setattr(desc, name, Mh.ctypes.data_as(L.f32p))
    return desc, keep


class FEOperator:
    """One operator handle per device.  All tensor arguments are dof-major storages [N, ldb]
    (fp32, contiguous, ldb % 4 == 0) living on `self.device`."""

    def __init__(self, n: int, A=None, B1=None, B2=None, S=None, idx_sol: Optional[Sequence] = None,
                 ns_precond_branch: bool = False, dt: float = 0.0, dense_m=None, dense_p=None,
                 device: Optional[torch.device] = None, dof_perm=None):
        """dof_perm (optional, reorder.py): new_of_old[d] = internal position of the caller's dof d.  Every matrix and idx_sol
        are renumbered at set-up; tensors cross `to_dof_major` / `from_dof_major` in the CALLER's numbering and the layout
        passes apply the permutation (row-major tensors: inside the transpose they need anyway; dof-major tensors: one row
        gather).  Everything between those two calls -- every kernel -- works in the internal numbering."""
        if not torch.cuda.is_available():
            raise L.FeoError("FEOperator needs a CUDA device: the FEM residual path has no CPU fallback")
        self.lib = L.load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n = int(n)
        idx_i = idx_j = None
        if idx_sol is not None:
            idx_i, idx_j = idx_sol[0], idx_sol[1]
        self.perm_new_of_old = self.perm_old_of_new = None
        if dof_perm is not None:
            from . import reorder as R

            pn = np.asarray(dof_perm, dtype=np.int64)
            if pn.shape != (self.n,) or not np.array_equal(np.sort(pn), np.arange(self.n)):
                raise ValueError("dof_perm must be a permutation of range(n)")
            if not R.is_identity(pn):
                A, B1, B2, S = (None if K is None else R.permute_csr(K, pn) for K in (A, B1, B2, S))
                dense_m, dense_p = (None if M is None else R.permute_dense(M, pn) for M in (dense_m, dense_p))
                if idx_i is not None:
                    idx_i, idx_j = pn[np.asarray(idx_i, dtype=np.int64)], pn[np.asarray(idx_j, dtype=np.int64)]
                self.perm_new_of_old = torch.tensor(pn, dtype=torch.int32, device=self.device)
                self.perm_old_of_new = torch.tensor(np.argsort(pn), dtype=torch.int32, device=self.device)
        desc, keep = build_desc(n, A, B1, B2, S, idx_i, idx_j, ns_precond_branch, dt, dense_m, dense_p)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.feo_op_create(C.byref(desc), C.byref(self._handle)))
        del keep
        info = L.FeoOpInfo()
        L.check(self.lib.feo_op_get_info(self._handle, C.byref(info)))
        self.info = info
        self.has_conv = bool(info.has_conv)
        self.has_seq = bool(info.has_seq)
        self.has_dense_m = bool(info.has_dense_m)
        self.has_dense_p = bool(info.has_dense_p)
        self.has_sparse = info.n_tiles_fwd > 0
        self.ns_precond_branch = bool(ns_precond_branch)
        self.plan = {0: "none", 1: "tile", 2: "patch", 3: "lattice"}[int(self.lib.feo_op_plan(self._handle))]
        self._ws = None
        self.launches = 0  # kernels launched through this handle (bench.py's gpu_launches)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self.lib.feo_op_destroy(h)
            except Exception:  # pragma: no cover
                pass
            self._handle = C.c_void_p()

    # -- helpers ------------------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        # the stream of THIS handle's device: the caller's current device may be another one
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, fn, *args) -> None:
        """Every launch runs with the handle's device current (the library sizes grids, encodes tensor maps and
        launches on the current device) and on that device's current stream."""
        with torch.cuda.device(self.device):
            L.check(fn(*args))

    @staticmethod
    def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _check(self, t: torch.Tensor, rows: Optional[int] = None):
        assert t.dtype == torch.float32 and t.is_cuda and t.dim() == 2 and t.is_contiguous(), "need fp32 CUDA [N, ldb]"
        assert t.shape[1] % 4 == 0 and (rows is None or t.shape[0] == rows)

    def workspace(self, B: int, T: int = 1) -> torch.Tensor:
        need = int(self.lib.feo_workspace_bytes(self._handle, B, T))
        need = max(need, 4096 * 4)
        if self._ws is None or self._ws.numel() * 4 < need:
            self._ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=self.device)
        return self._ws

    def new(self, ldb: int, rows: Optional[int] = None) -> torch.Tensor:
        return torch.empty((self.n if rows is None else rows, ldb), dtype=torch.float32, device=self.device)

    # -- layout -------------------------------------------------------------------------------
    def to_dof_major(self, x: torch.Tensor, ldb: Optional[int] = None) -> torch.Tensor:
        """row-major [B,N] (any strides) -> dof-major storage [N, ldb]; zero-copy when x already is
        a dof-major view (stride(0)==1, stride(1)%4==0, 16-B aligned).  With a dof permutation the rows of the result are in
        the INTERNAL numbering: scattered by the transpose itself, or gathered row by row from a dof-major input."""
        assert x.dim() == 2 and x.dtype == torch.float32 and x.is_cuda
        B, N = x.shape
        want = ceil4(B) if ldb is None else ldb
        if (x.stride(0) == 1 or B == 1) and x.stride(1) % 4 == 0 and x.stride(1) >= ceil4(B) and x.data_ptr() % 16 == 0 \
                and (ldb is None or x.stride(1) == ldb) and x.untyped_storage().nbytes() - x.storage_offset() * 4 >= N * x.stride(1) * 4:
            xT = torch.as_strided(x, (N, x.stride(1)), (x.stride(1), 1))
            if self.perm_old_of_new is None:
                return xT
            return xT.index_select(0, self.perm_old_of_new)  # whole 4*ldb-byte rows move: a coalesced gather pass
        if x.stride(1) != 1:
            x = x.contiguous()
        out = torch.empty((N, want), dtype=torch.float32, device=x.device)
        self._call(self.lib.feo_transpose, self._p(x), x.stride(0), self._p(out), want, B, N, self._p(self.perm_new_of_old), self._stream())
        self.launches += 1
        return out

    def from_dof_major(self, xT: torch.Tensor, B: int, contiguous: bool = False) -> torch.Tensor:
        """dof-major storage [N, ldb] (internal numbering) -> [B,N] tensor in the caller's numbering; a strided view unless
        `contiguous`."""
        if not contiguous:
            if self.perm_new_of_old is not None:
                xT = xT.index_select(0, self.perm_new_of_old)
            return xT[:, :B].t()
        N, ldb = xT.shape
        out = torch.empty((B, N), dtype=torch.float32, device=xT.device)
        if self.perm_new_of_old is None:
            self._call(self.lib.feo_transpose, self._p(xT), ldb, self._p(out), N, N, B, None, self._stream())
        else:  # out[b, d] = xT[new_of_old[d], b]
            self._call(self.lib.feo_transpose_gather, self._p(xT), ldb, self._p(out), N, N, B, self._p(self.perm_new_of_old), self._stream())
        self.launches += 1
        return out

    # -- fused sparse residual ----------------------------------------------------------------
    def residual_fwd(self, aT: torch.Tensor, fT: torch.Tensor, B: int, save: bool = True):
        self._check(aT, self.n)
        self._check(fT, self.n)
        ldb = aT.shape[1]
        assert fT.shape[1] == ldb
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        rT = self.new(ldb) if save else None
        ws = self.workspace(B)
        self._call(self.lib.feo_residual_fwd, self._handle, self._p(aT), self._p(fT), ldb, B, self._p(loss), self._p(rT),
                                          self._p(ws), ws.numel() * 4, self._stream())
        self.launches += 2
        return loss, rT

    def residual_bwd(self, aT, rT, B: int, grad_loss: Optional[torch.Tensor] = None, out=None) -> torch.Tensor:
        ldb = rT.shape[1]
        gT = self.new(ldb) if out is None else out
        self._call(self.lib.feo_residual_bwd, self._handle, self._p(aT), self._p(rT), self._p(grad_loss),
                                          self._p(gT), ldb, B, self._stream())
        self.launches += 1
        return gT

    # -- generic applies ----------------------------------------------------------------------
    def spmm(self, which: int, transpose: bool, xT: torch.Tensor, B: int, scale: float = 1.0, out=None,
             accumulate: bool = False) -> torch.Tensor:
        self._check(xT, self.n)
        yT = self.new(xT.shape[1]) if out is None else out
        self._call(self.lib.feo_spmm, self._handle, which, int(transpose), self._p(xT), self._p(yT), xT.shape[1], B,
                                  float(scale), int(accumulate), self._stream())
        self.launches += 1
        return yT

    def dense_apply(self, which: int, xT: torch.Tensor, B: int, scale: float = 1.0, scale_dev=None, sub=None,
                    want_loss: bool = False):
        self._check(xT, self.n)
        ldb = xT.shape[1]
        cT = self.new(ldb)
        loss = torch.empty((), dtype=torch.float32, device=self.device) if want_loss else None
        ws = self.workspace(B)
        self._call(self.lib.feo_dense_apply, self._handle, which, self._p(xT), self._p(cT), ldb, B, float(scale),
                                         self._p(scale_dev), self._p(sub), self._p(loss), self._p(ws), ws.numel() * 4,
                                         self._stream())
        self.launches += 2 if want_loss else 1
        return (cT, loss) if want_loss else cT

    def sq_diff_sum(self, xT, yT, B: int, scale: float = 1.0) -> torch.Tensor:
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        ws = self.workspace(B)
        self._call(self.lib.feo_sq_diff_sum, self._p(xT), self._p(yT), xT.shape[0], xT.shape[1], B, float(scale),
                                         self._p(loss), self._p(ws), ws.numel() * 4, self._stream())
        self.launches += 2
        return loss

    # -- time-dependent -----------------------------------------------------------------------
    def seq_fwd(self, pT, u0T, fT, B: int, T: int):
        ldj, ldb = pT.shape[1], u0T.shape[1]
        assert fT.shape[1] == ldb
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        rT = self.new(ldj)
        ws = self.workspace(B, T)
        self._call(self.lib.feo_seq_fwd, self._handle, self._p(pT), self._p(u0T), self._p(fT), ldj, ldb, B, T, self._p(loss),
                                     self._p(rT), self._p(ws), ws.numel() * 4, self._stream())
        self.launches += 2
        return loss, rT

    def seq_bwd(self, rT, B: int, T: int, grad_loss=None) -> torch.Tensor:
        gT = self.new(rT.shape[1])
        self._call(self.lib.feo_seq_bwd, self._handle, self._p(rT), self._p(grad_loss), self._p(gT), rT.shape[1], B, T,
                                     self._stream())
        self.launches += 1
        return gT

    def assemble_u_init(self, init_x: torch.Tensor, init_y: torch.Tensor) -> torch.Tensor:
        """[B, n_u] x2 -> dof-major u0 storage [N, ceil4(B)]."""
        B = init_x.shape[0]
        init_x = init_x.reshape(B, -1).contiguous().float()
        init_y = init_y.reshape(B, -1).contiguous().float()
        u0T = self.new(ceil4(B))
        self._call(self.lib.feo_assemble_u_init, self._handle, self._p(init_x), self._p(init_y), self._p(u0T), u0T.shape[1], B,
                                             self._stream())
        self.launches += 3
        return u0T
