"""SPAI preconditioner set-up on the device.

The reference builds `precond_{ne}_{bc}.npy` with `spai(A, m)` (`FEONet_Stokes_square/train_FEONet.py:104-121`): m
steps of the minimal-residual iteration  M <- M + a (I - A M),  a = <G, A G> / ||A G||_F^2,  started from
M = 2 / ||A A^T||_1 * A, all in dense scipy/numpy on the host -- 10 to 40 minutes for its meshes
(`test.ipynb#c4`: 20 000 steps at 31 it/s for N = 387).  It is one-off set-up, three dense N x N products per step:
here the same recurrence runs in fp64 on the GPU through `torch.matmul` (plain library GEMMs; nothing to fuse).
"""
from __future__ import annotations

import numpy as np
import torch


def spai_device(A, m: int, device=None, dtype=torch.float64, start: str = "onenormest") -> torch.Tensor:
    """m SPAI steps on `device`; returns the dense preconditioner [N, N] (same dtype) on that device.

    start="onenormest" (default) scales the initial guess exactly as the reference does, with scipy's 1-norm ESTIMATE of
    A A^T (one scalar, evaluated on the host from the device product; the reference's own dependency); start="exact" uses the
    exact 1-norm.  The two coincide unless the estimator misses the maximal column (it is exact for every operator in
    tests/)."""
    dev = torch.device(device if device is not None else "cuda")
    if hasattr(A, "todense"):
        A = np.asarray(A.todense())
    A = torch.as_tensor(np.asarray(A) if not isinstance(A, torch.Tensor) else A, dtype=dtype, device=dev)
    n = A.shape[0]
    eye = torch.eye(n, dtype=dtype, device=dev)
    AAt = A @ A.T
    if start == "onenormest":
        from scipy.sparse.linalg import onenormest

        norm1 = float(onenormest(AAt.cpu().numpy()))
    elif start == "exact":
        norm1 = float(torch.linalg.matrix_norm(AAt, ord=1))
    else:
        raise ValueError("start must be 'onenormest' or 'exact'")
    M = (2.0 / norm1) * A
    for _ in range(int(m)):
        G = eye - A @ M
        AG = A @ G
        M = M + (torch.sum(G * AG) / torch.sum(AG * AG)) * G
    return M
