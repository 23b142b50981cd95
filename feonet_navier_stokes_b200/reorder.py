"""Dof permutations: how an operator in ANY dof order reaches the lattice kernels.

The lattice plan (csrc/feo_lattice_plan.cpp) recognises a structured right-diagonal P2-P1 mesh only in the lattice-
lexicographic interleaved numbering -- P2 nodes row by row, (u1, u2[, p]) of a node adjacent.  FEniCS numbers a mixed space its
own way, and the reference stores the coordinates of every global dof next to the operators (`p = W.tabulate_dof_coordinates()`,
`idx_sol = [dofs(u1), dofs(u2), dofs(p)]`: FEONet_steady_Navier-Stokes/assemble_fenics.py:121-142, saved :350-353).  From those
two arrays `lattice_permutation` derives the renumbering; `FEOperator(dof_perm=...)` applies it to the matrices at set-up and
folds it into the row-major <-> dof-major layout passes (`feo_transpose` / `feo_transpose_gather`), which the reference-facing
boundary runs anyway: no extra pass at run time.  The planner still verifies every stored entry against the generated stencil,
so a wrong guess only means the tile plan is used.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def lattice_permutation(idx_sol: Sequence, pos: np.ndarray, tol: float = 1e-6) -> Optional[np.ndarray]:
    """new_of_old[d] = position of global dof d in the interleaved lattice-lexicographic numbering, or None when the dofs do
    not sit on a full (2n+1) x (2n+1) half-step lattice with one (u1, u2) pair per node and one pressure dof per vertex.

    idx_sol = (I, J, K) global dof ids of u1, u2, p; pos [N, 2] coordinates of every global dof."""
    I, J, K = (np.asarray(list(x), dtype=np.int64) for x in idx_sol[:3])
    pos = np.asarray(pos, dtype=np.float64)
    N = I.size + J.size + K.size
    if pos.ndim != 2 or pos.shape[0] != N or pos.shape[1] < 2 or I.size != J.size or I.size == 0:
        return None
    m = int(round(np.sqrt(I.size)))
    if m * m != I.size or m < 3 or m % 2 == 0:
        return None
    n = (m - 1) // 2
    if K.size != (n + 1) ** 2:
        return None
    lo, hi = pos[I].min(axis=0)[:2], pos[I].max(axis=0)[:2]
    h2 = (hi - lo) / (m - 1)
    if np.any(h2 <= 0):
        return None

    def lattice(ids):
        q = (pos[ids, :2] - lo) / h2
        r = np.rint(q)
        if np.abs(q - r).max() > tol * m or r.min() < 0 or r.max() > m - 1:
            return None
        return r.astype(np.int64)

    li, lj, lk = lattice(I), lattice(J), lattice(K)
    if li is None or lj is None or lk is None or not np.array_equal(li, lj):  # (I[k], J[k]) must be colocated
        return None
    if np.any(lk % 2):  # pressure dofs live on vertices
        return None
    node_u = li[:, 1] * m + li[:, 0]
    node_p = lk[:, 1] * m + lk[:, 0]
    if np.unique(node_u).size != I.size or np.unique(node_p).size != K.size:
        return None
    has_p = np.zeros(m * m, dtype=np.int64)
    has_p[node_p] = 1
    start = np.concatenate([[0], np.cumsum(2 + has_p)[:-1]])
    new_of_old = np.full(N, -1, dtype=np.int64)
    new_of_old[I] = start[node_u]
    new_of_old[J] = start[node_u] + 1
    new_of_old[K] = start[node_p] + 2
    if new_of_old.min() < 0 or np.unique(new_of_old).size != N:
        return None
    return new_of_old


def is_identity(perm: Optional[np.ndarray]) -> bool:
    return perm is None or bool(np.array_equal(np.asarray(perm), np.arange(len(perm))))


def permute_csr(K, new_of_old: np.ndarray):
    """P K P^T as scipy CSR with sorted columns: entry (r, c) moves to (new_of_old[r], new_of_old[c])."""
    import scipy.sparse as sp

    from .operator import to_host_csr

    t = to_host_csr(K)
    n = t[0].shape[0] - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(t[0]))
    out = sp.csr_matrix((t[2], (new_of_old[rows], new_of_old[t[1]])), shape=(n, n))
    out.sort_indices()
    return out


def permute_dense(M, new_of_old: np.ndarray) -> np.ndarray:
    """P M P^T for a dense [N, N] matrix."""
    from .operator import _dense_host

    Mh = _dense_host(M)
    old_of_new = np.argsort(new_of_old)
    return np.ascontiguousarray(Mh[np.ix_(old_of_new, old_of_new)])
