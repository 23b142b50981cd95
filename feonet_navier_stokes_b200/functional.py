"""torch.autograd glue over the C ABI.

Tensors cross this layer as ordinary `[B, N]` torch tensors.  The device-native layout is
dof-major (`strides == (1, ldb)`, see include/feonet_b200.h); a tensor that already has it is
used in place (zero copy), anything else goes through `feo_transpose`.  Gradients come back in
the layout the input arrived in.  `dof_major_empty/zeros` create native tensors; `LinearT`
(network.py) makes a network emit them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L
from .operator import FEOperator, ceil4


def dof_major_empty(B: int, N: int, device, ldb: Optional[int] = None) -> torch.Tensor:
    """A [B,N] fp32 tensor whose memory is dof-major ([N, ldb] storage)."""
    ldb = ceil4(B) if ldb is None else ldb
    return torch.empty((N, ldb), dtype=torch.float32, device=device)[:, :B].t()


def dof_major_zeros(B: int, N: int, device, ldb: Optional[int] = None) -> torch.Tensor:
    ldb = ceil4(B) if ldb is None else ldb
    return torch.zeros((N, ldb), dtype=torch.float32, device=device)[:, :B].t()


def to_dof_major_tensor(x: torch.Tensor) -> torch.Tensor:
    """Copy a [B,N] tensor into dof-major memory (pure torch; for data set-up, e.g. load vectors)."""
    out = dof_major_empty(x.shape[0], x.shape[1], x.device)
    out.copy_(x)
    return out


def is_dof_major(x: torch.Tensor) -> bool:
    return x.dim() == 2 and (x.stride(0) == 1 or x.shape[0] == 1) and x.stride(1) % 4 == 0 \
        and x.stride(1) >= ceil4(x.shape[0]) and x.data_ptr() % 16 == 0


def _prep(op: FEOperator, x: torch.Tensor, ldb: Optional[int] = None) -> torch.Tensor:
    if x.dtype != torch.float32:
        x = x.float()
    return op.to_dof_major(x, ldb)


class _TransposeCache:
    """Load vectors are data: in full-batch training the same tensor OBJECT comes back every epoch, so its
    dof-major copy is cached.  The key is the identity of the tensor (a strong reference is kept, so its storage
    cannot be recycled for another tensor while the entry lives) plus its version counter: a loop that makes a
    fresh device tensor per step (`v.to(device)` on a shuffled DataLoader batch, as the reference epoch loop does)
    never hits, whatever address the caching allocator hands back."""

    def __init__(self):
        self.src = None
        self.version = -1
        self.ldb = -1
        self.val = None

    def get(self, op: FEOperator, f: torch.Tensor, ldb: int) -> torch.Tensor:
        if f.is_cuda and torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture (graphs.py): the layout pass must be IN the graph -- a replay sees whatever the caller has
            # copied into the static load-vector buffer since, which a copy cached at capture time would not
            return _prep(op, f.detach(), ldb)
        if self.src is not f or self.version != f._version or self.ldb != ldb:
            self.val = _prep(op, f.detach(), ldb)
            self.src, self.version, self.ldb = f, f._version, ldb
        return self.val


class ResidualLossFn(torch.autograd.Function):
    """loss = sum (A a -/+ (F - c))^2, fused sparse forward/backward (feo_residual_fwd/bwd).

    Reference: weak_form + loss loop of closure, FEONet_steady_Navier-Stokes/train_FEONet.py:301-360,
    FEONet_Stokes_square/train_FEONet.py:261-296 (un-preconditioned)."""

    @staticmethod
    def forward(ctx, alpha: torch.Tensor, F: torch.Tensor, op: FEOperator, fcache: Optional[_TransposeCache]):
        B, N = alpha.shape
        native = is_dof_major(alpha)
        aT = _prep(op, alpha.detach())
        ldb = aT.shape[1]
        fT = fcache.get(op, F, ldb) if fcache is not None else _prep(op, F.detach(), ldb)
        need = bool(ctx.needs_input_grad[0])
        loss, rT = op.residual_fwd(aT, fT, B, save=need)
        ctx.op, ctx.B, ctx.native = op, B, native
        if need:
            ctx.save_for_backward(aT, rT)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        aT, rT = ctx.saved_tensors
        op: FEOperator = ctx.op
        g = grad_out.detach().to(torch.float32).contiguous()
        gT = op.residual_bwd(aT, rT, ctx.B, grad_loss=g)
        return op.from_dof_major(gT, ctx.B, contiguous=not ctx.native), None, None, None


class DenseResidualLossFn(torch.autograd.Function):
    """loss = sum (M a - F)^2 with a dense operator M = A @ P (preconditioned linear Stokes,
    FEONet_Stokes_square/train_FEONet.py:264, :290-296)."""

    @staticmethod
    def forward(ctx, alpha, F, op: FEOperator, fcache):
        B, N = alpha.shape
        native = is_dof_major(alpha)
        aT = _prep(op, alpha.detach())
        ldb = aT.shape[1]
        fT = fcache.get(op, F, ldb) if fcache is not None else _prep(op, F.detach(), ldb)
        rT, loss = op.dense_apply(L.FEO_DENSE_M, aT, B, sub=fT, want_loss=True)
        ctx.op, ctx.B, ctx.native = op, B, native
        ctx.save_for_backward(rT)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (rT,) = ctx.saved_tensors
        op: FEOperator = ctx.op
        g = grad_out.detach().to(torch.float32).contiguous()
        gT = op.dense_apply(L.FEO_DENSE_MT, rT, ctx.B, scale=2.0, scale_dev=g)
        return op.from_dof_major(gT, ctx.B, contiguous=not ctx.native), None, None, None


class SpmmFn(torch.autograd.Function):
    """Y = X @ K^T for a stored sparse matrix K (the reference's `u_batch @ K.T`,
    FEONet_steady_Navier-Stokes/train_FEONet.py:308-309, :329); VJP = G @ K."""

    @staticmethod
    def forward(ctx, x, op: FEOperator, which: int):
        B = x.shape[0]
        native = is_dof_major(x)
        yT = op.spmm(which, False, _prep(op, x.detach()), B)
        ctx.op, ctx.which, ctx.B, ctx.native = op, which, B, native
        return op.from_dof_major(yT, B, contiguous=not native)

    @staticmethod
    def backward(ctx, g):
        op: FEOperator = ctx.op
        gT = op.spmm(ctx.which, True, _prep(op, g), ctx.B)
        return op.from_dof_major(gT, ctx.B, contiguous=not ctx.native), None, None


class DenseFn(torch.autograd.Function):
    """Y = X @ M^T for the stored dense LHS operator (VJP through the stored M^T)."""

    @staticmethod
    def forward(ctx, x, op: FEOperator):
        B = x.shape[0]
        native = is_dof_major(x)
        yT = op.dense_apply(L.FEO_DENSE_M, _prep(op, x.detach()), B)
        ctx.op, ctx.B, ctx.native = op, B, native
        return op.from_dof_major(yT, B, contiguous=not native)

    @staticmethod
    def backward(ctx, g):
        op: FEOperator = ctx.op
        gT = op.dense_apply(L.FEO_DENSE_MT, _prep(op, g), ctx.B)
        return op.from_dof_major(gT, ctx.B, contiguous=not ctx.native), None


class SeqResidualLossFn(torch.autograd.Function):
    """loss = (1/T) sum_t |(S+dt A) x_t - S prev_t - dt F|^2, un-preconditioned sparse form of
    FEONet_time_dep_Stokes/train_FEONet.py:343-362, :398-400. pred_seq [B,T,N], u_init/F [B,N]."""

    @staticmethod
    def forward(ctx, pred_seq, u_init, F, op: FEOperator):
        B, T, N = pred_seq.shape
        p2 = pred_seq.detach().reshape(B * T, N)
        pT = _prep(op, p2)
        u0T = _prep(op, u_init.detach())
        fT = _prep(op, F.detach(), u0T.shape[1])
        loss, rT = op.seq_fwd(pT, u0T, fT, B, T)
        ctx.op, ctx.B, ctx.T = op, B, T
        ctx.save_for_backward(rT)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (rT,) = ctx.saved_tensors
        op: FEOperator = ctx.op
        g = grad_out.detach().to(torch.float32).contiguous()
        gT = op.seq_bwd(rT, ctx.B, ctx.T, grad_loss=g)
        grad = op.from_dof_major(gT, ctx.B * ctx.T, contiguous=True).reshape(ctx.B, ctx.T, -1)
        return grad, None, None, None


class NsDenseResidualLossFn(torch.autograd.Function):
    """Steady Navier-Stokes with a genuinely dense preconditioner (FEONet_steady_Navier-Stokes/train_FEONet.py:324-326 with
    PRECOND != I): r = alpha (A P)^T - (F - c), c from the raw alpha.  The convective part comes from the fused sparse kernels of
    an operator that holds B1, B2 and an EMPTY A (its residual is c - F), the linear part from one tensor-core apply that
    subtracts F - c and reduces the squares; backward = M^T r (dense) + the sparse kernels' convective terms.  No eager index
    arithmetic, no materialised LHS / RHS."""

    @staticmethod
    def forward(ctx, alpha, F, op: FEOperator, op_conv: FEOperator, fcache):
        B, N = alpha.shape
        native = is_dof_major(alpha)
        aT = _prep(op, alpha.detach())
        ldb = aT.shape[1]
        fT = fcache.get(op, F, ldb) if fcache is not None else _prep(op, F.detach(), ldb)
        _, cT = op_conv.residual_fwd(aT, fT, B)  # 0 - (F - c)
        rT, loss = op.dense_apply(L.FEO_DENSE_M, aT, B, sub=cT.neg_(), want_loss=True)
        ctx.op, ctx.op_conv, ctx.B, ctx.native = op, op_conv, B, native
        ctx.save_for_backward(aT, rT)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        aT, rT = ctx.saved_tensors
        op: FEOperator = ctx.op
        g = grad_out.detach().to(torch.float32).contiguous()
        gT = op.dense_apply(L.FEO_DENSE_MT, rT, ctx.B, scale=2.0, scale_dev=g)
        gT.add_(ctx.op_conv.residual_bwd(aT, rT, ctx.B, grad_loss=g))
        return op.from_dof_major(gT, ctx.B, contiguous=not ctx.native), None, None, None, None


class SeqDenseResidualLossFn(torch.autograd.Function):
    """loss = (1/T) sum_t |M x_t - S prev_t - dt F|^2 with the dense preconditioned system matrix M = (S + dt A) P
    (FEONet_time_dep_Stokes/train_FEONet.py:343-362 with do_precond, :398-400).  The right-hand side S prev + dt F is formed
    once (sparse apply), then ONE tensor-core apply evaluates M x, subtracts it and reduces the squares (feo_dense_apply with
    `sub` and `loss_out`): no materialised LHS, no eager reduction.  Backward: g_t = (2/T) [M^T r_t - S^T r_{t+1}], r_T = 0."""

    @staticmethod
    def forward(ctx, pred_seq, u_init, F, op: FEOperator, dt: float):
        B, T, N = pred_seq.shape
        BT = B * T
        pT = _prep(op, pred_seq.detach().reshape(BT, N))  # pseudo-samples j = b * T + t
        ld = pT.shape[1]
        u0T = _prep(op, u_init.detach())
        fT = _prep(op, F.detach(), u0T.shape[1])
        prevT = op.new(ld)
        pv, xv = prevT[:, :BT].view(N, B, T), pT[:, :BT].view(N, B, T)
        pv[:, :, 1:] = xv[:, :, :-1]
        pv[:, :, 0] = u0T[:, :B]
        if ld > BT:
            prevT[:, BT:] = 0.0
        rhsT = op.spmm(L.FEO_MAT_S, False, prevT, BT)
        rhsT[:, :BT].view(N, B, T).add_(fT[:, :B].unsqueeze(2), alpha=float(dt))
        rT, loss = op.dense_apply(L.FEO_DENSE_M, pT, BT, sub=rhsT, want_loss=True)
        ctx.op, ctx.B, ctx.T = op, B, T
        ctx.save_for_backward(rT)
        return loss / T

    @staticmethod
    def backward(ctx, grad_out):
        (rT,) = ctx.saved_tensors
        op: FEOperator = ctx.op
        B, T = ctx.B, ctx.T
        BT, N = B * T, rT.shape[0]
        g = grad_out.detach().to(torch.float32).contiguous()
        gT = op.dense_apply(L.FEO_DENSE_MT, rT, BT, scale=2.0 / T, scale_dev=g)
        nextT = torch.zeros_like(rT)
        nextT[:, :BT].view(N, B, T)[:, :, :-1] = rT[:, :BT].view(N, B, T)[:, :, 1:]
        sT = op.spmm(L.FEO_MAT_S, True, nextT, BT)
        gT.add_(sT * (g * (-2.0 / T)))
        grad = op.from_dof_major(gT, BT, contiguous=True).reshape(B, T, -1)
        return grad, None, None, None, None


def precond_output(op: FEOperator, pred: torch.Tensor) -> torch.Tensor:
    """u = (P @ pred^T)^T, the second output of `closure`.  Only ever used detached in the reference
    (FEONet_Stokes_square/train_FEONet.py:401, steady NS :478 -- quirk 9), so no graph is kept."""
    shape = pred.shape
    x = pred.detach().reshape(-1, shape[-1])
    B = x.shape[0]
    uT = op.dense_apply(L.FEO_DENSE_P, _prep(op, x), B)
    return op.from_dof_major(uT, B, contiguous=True).reshape(shape)
