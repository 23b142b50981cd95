"""Offline P2-P1 (Taylor-Hood) assembler used as a TEST / BENCH FIXTURE.

The reference builds its operators with legacy FEniCS (`dolfin`, `mshr`), which is not
available offline.  This module re-creates the same discrete operators with numpy/scipy so
that the hot path has realistic inputs:

* forms: `FEONet_Stokes_square/assemble_fenics.py:59-60` (mu<grad v,grad u> - p div v - q div u),
  `FEONet_steady_Navier-Stokes/assemble_fenics.py:87-98` (A, B1, B2),
  `FEONet-square-with-hole/assemble_fenics.py:88-90` (symmetric-gradient form, +q div u),
  `FEONet_time_dep_Stokes/assemble_fenics.py:108-110,122` (mass S, +q div u);
* Dirichlet rows: `bc.apply(K)` == zero the row, 1.0 on the diagonal, applied to every matrix
  including B1, B2 and S (`FEONet_steady_Navier-Stokes/assemble_fenics.py:104-117`);
* npz schema: SURVEY.md Appendix C.

Known-answer pins (reference notebooks): n=6 channel_flow => N=387, ne=72,
cond_2(A) = 167.32636402645198 (`FEONet_Stokes_square/test.ipynb#c3`); A min/max =
-0.13333/1.0 and B1 (no BC) min/max = -/+0.2667/n
(`FEONet_steady_Navier-Stokes/compare_ordering_nonlinear.ipynb#c13,c15`).

Nothing here is on the product path: the CUDA library consumes the CSR arrays this module
(or a real reference npz) produces.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.sparse as sp

# 6-point, degree-4 Dunavant rule on the reference triangle (weights sum to 1).
_QA, _QWA = 0.445948490915965, 0.223381589678011
_QB, _QWB = 0.091576213509771, 0.109951743655322
_QUAD_L = np.array(
    [
        [1 - 2 * _QA, _QA, _QA],
        [_QA, 1 - 2 * _QA, _QA],
        [_QA, _QA, 1 - 2 * _QA],
        [1 - 2 * _QB, _QB, _QB],
        [_QB, 1 - 2 * _QB, _QB],
        [_QB, _QB, 1 - 2 * _QB],
    ]
)
_QUAD_W = np.array([_QWA] * 3 + [_QWB] * 3)


def _p2_basis(lam: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """P2 basis values [q,6] and d(phi)/d(lambda_i) [q,6,3] at barycentric points lam [q,3].

    Local node order: 3 vertices, then the midpoints of the edges opposite vertex 0, 1, 2.
    """
    l0, l1, l2 = lam[:, 0], lam[:, 1], lam[:, 2]
    z = np.zeros_like(l0)
    phi = np.stack(
        [l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1],
        axis=1,
    )
    dphi = np.stack(
        [
            np.stack([4 * l0 - 1, z, z], axis=1),
            np.stack([z, 4 * l1 - 1, z], axis=1),
            np.stack([z, z, 4 * l2 - 1], axis=1),
            np.stack([z, 4 * l2, 4 * l1], axis=1),
            np.stack([4 * l2, z, 4 * l0], axis=1),
            np.stack([4 * l1, 4 * l0, z], axis=1),
        ],
        axis=1,
    )
    return phi, dphi


@dataclass
class P2P1Mesh:
    """Triangulation with P2 velocity nodes and P1 pressure nodes."""

    p2_xy: np.ndarray  # [n_u, 2] coordinates of P2 nodes
    p1_xy: np.ndarray  # [n_p, 2] coordinates of P1 nodes (mesh vertices)
    tri_p2: np.ndarray  # [ne, 6] P2 node ids per element (local order of _p2_basis)
    tri_p1: np.ndarray  # [ne, 3] P1 node ids per element
    p1_to_p2: np.ndarray  # [n_p] P2 node id colocated with each P1 node
    n_struct: int = 0  # cells per side for structured meshes, 0 otherwise

    @property
    def ne(self) -> int:
        return int(self.tri_p2.shape[0])

    @property
    def n_u(self) -> int:
        return int(self.p2_xy.shape[0])

    @property
    def n_p(self) -> int:
        return int(self.p1_xy.shape[0])


def structured_mesh(n: int, lo: float = 0.0, hi: float = 1.0) -> P2P1Mesh:
    """`RectangleMesh(Point(lo,lo), Point(hi,hi), n, n)` with the default "right" diagonal.

    P2 nodes live on the (2n+1)^2 half-step lattice in lexicographic (y-major) order, so
    n_u=(2n+1)^2, n_p=(n+1)^2, ne=2n^2 (SURVEY.md section 8 size table).
    """
    m = 2 * n + 1
    fx, fy = np.meshgrid(np.arange(m), np.arange(m), indexing="xy")
    h2 = (hi - lo) / (2 * n)
    p2_xy = np.stack([lo + fx.ravel() * h2, lo + fy.ravel() * h2], axis=1)

    def nid(ix, iy):  # lattice -> P2 node id
        return iy * m + ix

    ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    ci, cj = ci.ravel(), cj.ravel()
    # lattice coordinates of the 4 cell corners
    v00 = (2 * ci, 2 * cj)
    v10 = (2 * ci + 2, 2 * cj)
    v11 = (2 * ci + 2, 2 * cj + 2)
    v01 = (2 * ci, 2 * cj + 2)

    def tri(a, b, c):
        mid = lambda p, q: ((p[0] + q[0]) // 2, (p[1] + q[1]) // 2)
        return np.stack(
            [nid(*a), nid(*b), nid(*c), nid(*mid(b, c)), nid(*mid(a, c)), nid(*mid(a, b))], axis=1
        )

    lower = tri(v00, v10, v11)
    upper = tri(v00, v11, v01)  # counter-clockwise
    tri_p2 = np.empty((2 * n * n, 6), dtype=np.int64)
    tri_p2[0::2] = lower
    tri_p2[1::2] = upper

    # P1 nodes = lattice points with both coordinates even, lexicographic
    pm = n + 1
    vx, vy = np.meshgrid(np.arange(pm), np.arange(pm), indexing="xy")
    p1_to_p2 = nid(2 * vx.ravel(), 2 * vy.ravel())
    p2_to_p1 = -np.ones(m * m, dtype=np.int64)
    p2_to_p1[p1_to_p2] = np.arange(pm * pm)
    tri_p1 = p2_to_p1[tri_p2[:, :3]]
    return P2P1Mesh(p2_xy, p2_xy[p1_to_p2], tri_p2, tri_p1, p1_to_p2, n_struct=n)


def unstructured_mesh(vertices: np.ndarray, triangles: np.ndarray) -> P2P1Mesh:
    """P2/P1 node sets for an arbitrary conforming triangulation (vertices [nv,2], triangles [ne,3])."""
    vertices = np.asarray(vertices, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64).copy()
    # orient counter-clockwise
    a, b, c = vertices[t[:, 0]], vertices[t[:, 1]], vertices[t[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    flip = area2 < 0
    t[flip, 1], t[flip, 2] = t[flip, 2].copy(), t[flip, 1].copy()
    nv = vertices.shape[0]
    # local edges opposite vertex 0,1,2
    e_loc = np.stack([t[:, [1, 2]], t[:, [0, 2]], t[:, [0, 1]]], axis=1)  # [ne,3,2]
    e_sorted = np.sort(e_loc, axis=2).reshape(-1, 2)
    uniq, inv = np.unique(e_sorted, axis=0, return_inverse=True)
    mids = 0.5 * (vertices[uniq[:, 0]] + vertices[uniq[:, 1]])
    p2_xy = np.concatenate([vertices, mids], axis=0)
    tri_p2 = np.concatenate([t, nv + inv.reshape(-1, 3)], axis=1)
    return P2P1Mesh(p2_xy, vertices.copy(), tri_p2, t, np.arange(nv), n_struct=0)


def square_with_hole_mesh(n_side: int = 10, radius: float = 0.5, seed: int = 0) -> P2P1Mesh:
    """Delaunay stand-in for mshr's `Rectangle((-1,-1),(1,1)) - Circle((0,0),0.5)`
    (`FEONet-square-with-hole/assemble_fenics.py:48-50`)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    t = np.linspace(-1.0, 1.0, n_side + 1)
    g = np.stack(np.meshgrid(t, t, indexing="xy"), axis=-1).reshape(-1, 2)
    h = 2.0 / n_side
    interior = (np.abs(g[:, 0]) < 1 - 1e-12) & (np.abs(g[:, 1]) < 1 - 1e-12)
    g = g + (rng.random(g.shape) - 0.5) * 0.15 * h * interior[:, None]
    keep = np.hypot(g[:, 0], g[:, 1]) > radius + 0.35 * h
    nc = max(12, int(round(2 * np.pi * radius / h)))
    ang = np.linspace(0, 2 * np.pi, nc, endpoint=False)
    circ = radius * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    pts = np.concatenate([g[keep], circ], axis=0)
    tri = Delaunay(pts).simplices
    cen = pts[tri].mean(axis=1)
    tri = tri[np.hypot(cen[:, 0], cen[:, 1]) > radius * np.cos(np.pi / nc) - 1e-9]
    used = np.unique(tri)
    remap = -np.ones(pts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    return unstructured_mesh(pts[used], remap[tri])


# ----------------------------------------------------------------------------------------------
# element integrals
# ----------------------------------------------------------------------------------------------
def _element_integrals(mesh: P2P1Mesh, chunk: int = 65536):
    """Yield per-chunk element matrices.

    Kdd[e,d,d',a,b] = int d_d phi_a d_d' phi_b ; D[e,d,a,b] = int phi_a d_d phi_b ;
    Q[e,d,a,j] = int d_d phi_a psi_j ; Mass[e,a,b] = int phi_a phi_b.
    """
    phi, dphi_dl = _p2_basis(_QUAD_L)  # [q,6], [q,6,3]
    psi = _QUAD_L  # P1 basis = barycentric coordinates [q,3]
    xy = mesh.p2_xy
    for s in range(0, mesh.ne, chunk):
        t = mesh.tri_p2[s : s + chunk]
        p0, p1, p2 = xy[t[:, 0]], xy[t[:, 1]], xy[t[:, 2]]
        area2 = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
        gl = np.empty((t.shape[0], 3, 2))
        gl[:, 0, 0], gl[:, 0, 1] = p1[:, 1] - p2[:, 1], p2[:, 0] - p1[:, 0]
        gl[:, 1, 0], gl[:, 1, 1] = p2[:, 1] - p0[:, 1], p0[:, 0] - p2[:, 0]
        gl[:, 2, 0], gl[:, 2, 1] = p0[:, 1] - p1[:, 1], p1[:, 0] - p0[:, 0]
        gl /= area2[:, None, None]
        area = 0.5 * np.abs(area2)
        dphi = np.einsum("qai,eid->eqad", dphi_dl, gl)  # [e,q,6,2]
        w = _QUAD_W[None, :] * area[:, None]  # [e,q]
        Kdd = np.einsum("eq,eqac,eqbd->ecdab", w, dphi, dphi)
        D = np.einsum("eq,qa,eqbd->edab", w, phi, dphi)
        Q = np.einsum("eq,eqad,qj->edaj", w, dphi, psi)
        Mass = np.einsum("eq,qa,qb->eab", w, phi, phi)
        yield slice(s, s + t.shape[0]), Kdd, D, Q, Mass


def _coo(rows, cols, vals, n, drop_rel: float = 1e-13):
    """Sum element contributions; couplings that cancel to round-off noise (|v| < drop_rel*max|v|,
    i.e. mathematically zero entries of the structured stencil) are removed so the stored
    pattern is the true one."""
    K = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n)).tocsr()
    if K.nnz:
        K.data[np.abs(K.data) < drop_rel * np.abs(K.data).max()] = 0.0
        K.eliminate_zeros()
    K.sort_indices()
    return K


@dataclass
class StokesOperators:
    """Discrete operators in the reference's npz conventions."""

    mesh: P2P1Mesh
    N: int
    idx_u1: np.ndarray
    idx_u2: np.ndarray
    idx_p: np.ndarray
    A: sp.csr_matrix
    B1: Optional[sp.csr_matrix] = None
    B2: Optional[sp.csr_matrix] = None
    S: Optional[sp.csr_matrix] = None
    bc_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    bc_vals: np.ndarray = field(default_factory=lambda: np.zeros(0))
    pos: Optional[np.ndarray] = None  # [N,2] dof coordinates (npz key "p")
    A_nobc: Optional[sp.csr_matrix] = None
    B1_nobc: Optional[sp.csr_matrix] = None

    @property
    def idx_sol(self) -> np.ndarray:
        """Object array of three python int lists, exactly what `np.load(...)['idx_sol']` yields
        (`FEONet_steady_Navier-Stokes/assemble_fenics.py:142-143`).  Built once (it is a module-level
        global in the reference): converting ~1e6 ints to python lists costs milliseconds."""
        cached = self.__dict__.get("_idx_sol")
        if cached is None:
            cached = np.empty(3, dtype=object)
            cached[0], cached[1], cached[2] = self.idx_u1.tolist(), self.idx_u2.tolist(), self.idx_p.tolist()
            self.__dict__["_idx_sol"] = cached
        return cached

    def load_vector_sincos(self, coeff: np.ndarray) -> np.ndarray:
        """L_a = int f.v for f=(m0 sin(n0 x+n1 y), m1 cos(n2 x+n3 y)); Dirichlet rows hold the BC
        value (`FEONet_Stokes_square/assemble_fenics.py:123-131`). coeff [B,6] -> [B,N]."""
        coeff = np.atleast_2d(coeff)
        mesh = self.mesh
        phi, _ = _p2_basis(_QUAD_L)
        xy = mesh.p2_xy
        out = np.zeros((coeff.shape[0], self.N))
        t = mesh.tri_p2
        p0, p1, p2 = xy[t[:, 0]], xy[t[:, 1]], xy[t[:, 2]]
        area = 0.5 * np.abs(
            (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
        )
        xq = np.einsum("qi,eid->eqd", _QUAD_L, np.stack([p0, p1, p2], axis=1))  # [e,q,2]
        w = _QUAD_W[None, :] * area[:, None]
        for b in range(coeff.shape[0]):
            m0, m1, n0, n1, n2, n3 = coeff[b]
            f1 = m0 * np.sin(n0 * xq[..., 0] + n1 * xq[..., 1])
            f2 = m1 * np.cos(n2 * xq[..., 0] + n3 * xq[..., 1])
            l1 = np.einsum("eq,eq,qa->ea", w, f1, phi)
            l2 = np.einsum("eq,eq,qa->ea", w, f2, phi)
            np.add.at(out[b], self.idx_u1[t], l1)
            np.add.at(out[b], self.idx_u2[t], l2)
        out[:, self.bc_dofs] = self.bc_vals[None, :]
        return out


def dof_numbering(mesh: P2P1Mesh, ordering: str = "blocked"):
    """Global dof ids for (u1, u2, p).

    "blocked": [u1 | u2 | p]; "interleaved": per P2 node (u1,u2[,p at vertices]), closer to what
    FEniCS emits for a mixed space.  The hot path treats idx_sol as opaque int lists either way
    (SURVEY.md section 8a quirk 4).
    """
    n_u, n_p = mesh.n_u, mesh.n_p
    if ordering == "blocked":
        return np.arange(n_u), n_u + np.arange(n_u), 2 * n_u + np.arange(n_p)
    if ordering == "interleaved":
        has_p = np.zeros(n_u, dtype=np.int64)
        has_p[mesh.p1_to_p2] = 1
        width = 2 + has_p
        start = np.concatenate([[0], np.cumsum(width)[:-1]])
        return start, start + 1, start[mesh.p1_to_p2] + 2
    raise ValueError(f"unknown ordering {ordering!r}")


def apply_dirichlet(K: sp.csr_matrix, dofs: np.ndarray) -> sp.csr_matrix:
    """dolfin's `bc.apply(K)`: zero the rows, put 1.0 on the diagonal (non-symmetric)."""
    if dofs.size == 0:
        return K.tocsr()
    n = K.shape[0]
    keep = np.ones(n)
    keep[dofs] = 0.0
    ident = np.zeros(n)
    ident[dofs] = 1.0
    out = (sp.diags(keep) @ K + sp.diags(ident)).tocsr()
    out.eliminate_zeros()
    out.sort_indices()
    return out


def channel_flow_bc(mesh: P2P1Mesh, idx_u1, idx_u2, idx_p, lo=0.0, hi=1.0, p_in=8.0):
    """walls y in {lo,hi}: u=(0,0); inflow x=lo: p=8; outflow x=hi: p=0
    (`FEONet_Stokes_square/assemble_fenics.py:46-54`)."""
    eps = 1e-12
    wall = (np.abs(mesh.p2_xy[:, 1] - lo) < eps) | (np.abs(mesh.p2_xy[:, 1] - hi) < eps)
    pin = np.abs(mesh.p1_xy[:, 0] - lo) < eps
    pout = np.abs(mesh.p1_xy[:, 0] - hi) < eps
    dofs = np.concatenate([idx_p[pin], idx_p[pout], idx_u1[wall], idx_u2[wall]])
    vals = np.concatenate([np.full(pin.sum(), p_in), np.zeros(pout.sum()), np.zeros(2 * wall.sum())])
    # a later bc overrides an earlier one on shared dofs, as in the reference's bc list order
    _, last = np.unique(dofs[::-1], return_index=True)
    sel = np.sort(dofs.size - 1 - last)
    return dofs[sel], vals[sel]


def lower_bc(mesh: P2P1Mesh, idx_u1, idx_u2, idx_p, lo=0.0):
    """y=lo: u=(3+1.7 sin(2 pi x), 0) (`FEONet_Stokes_square/assemble_fenics.py:40-45`)."""
    low = np.abs(mesh.p2_xy[:, 1] - lo) < 1e-12
    x = mesh.p2_xy[low, 0]
    dofs = np.concatenate([idx_u1[low], idx_u2[low]])
    vals = np.concatenate([3.0 + 1.7 * np.sin(2 * np.pi * x), np.zeros(low.sum())])
    return dofs, vals


def assemble_operators(
    mesh: P2P1Mesh,
    mu: float = 0.1,
    bc: str = "channel_flow",
    ordering: str = "blocked",
    form: str = "grad",
    q_sign: float = -1.0,
    with_convection: bool = False,
    with_mass: bool = False,
    domain=(0.0, 1.0),
    keep_nobc: bool = False,
) -> StokesOperators:
    """Assemble A (and B1,B2,S) with identity Dirichlet rows.

    form="grad":    mu<grad v,grad u> - p div v + q_sign * q div u
    form="symgrad": 0.5 mu<grad v+grad v^T, grad u+grad u^T> - p div v + q_sign * q div u
    """
    iu1, iu2, ip = dof_numbering(mesh, ordering)
    N = 2 * mesh.n_u + mesh.n_p
    t2, t1 = mesh.tri_p2, mesh.tri_p1
    blocks: Dict[str, list] = {k: [] for k in ("A", "B1", "B2", "S")}

    def add(name, rows, cols, vals):
        blocks[name].append((rows.ravel(), cols.ravel(), vals.ravel()))

    for sl, Kdd, D, Q, Mass in _element_integrals(mesh):
        g2, g1 = t2[sl], t1[sl]
        r_u1 = np.broadcast_to(iu1[g2][:, :, None], (g2.shape[0], 6, 6))
        c_u1 = np.broadcast_to(iu1[g2][:, None, :], (g2.shape[0], 6, 6))
        r_u2 = np.broadcast_to(iu2[g2][:, :, None], (g2.shape[0], 6, 6))
        c_u2 = np.broadcast_to(iu2[g2][:, None, :], (g2.shape[0], 6, 6))
        lap = Kdd[:, 0, 0] + Kdd[:, 1, 1]
        if form == "grad":
            add("A", r_u1, c_u1, mu * lap)
            add("A", r_u2, c_u2, mu * lap)
        elif form == "symgrad":
            add("A", r_u1, c_u1, mu * (lap + Kdd[:, 0, 0]))
            add("A", r_u1, c_u2, mu * Kdd[:, 1, 0])  # d_y phi_a d_x phi_b
            add("A", r_u2, c_u1, mu * Kdd[:, 0, 1])
            add("A", r_u2, c_u2, mu * (lap + Kdd[:, 1, 1]))
        else:
            raise ValueError(form)
        # velocity rows x pressure cols: -int p d_d phi_a
        ru1p = np.broadcast_to(iu1[g2][:, :, None], (g2.shape[0], 6, 3))
        ru2p = np.broadcast_to(iu2[g2][:, :, None], (g2.shape[0], 6, 3))
        cp = np.broadcast_to(ip[g1][:, None, :], (g2.shape[0], 6, 3))
        add("A", ru1p, cp, -Q[:, 0])
        add("A", ru2p, cp, -Q[:, 1])
        # pressure rows x velocity cols: q_sign * int psi_j d_d phi_b
        add("A", cp, ru1p, q_sign * Q[:, 0])
        add("A", cp, ru2p, q_sign * Q[:, 1])
        if with_convection:
            add("B1", r_u1, c_u1, D[:, 0])
            add("B1", r_u2, c_u2, D[:, 0])
            add("B2", r_u1, c_u1, D[:, 1])
            add("B2", r_u2, c_u2, D[:, 1])
        if with_mass:
            add("S", r_u1, c_u1, Mass)
            add("S", r_u2, c_u2, Mass)

    def build(name):
        if not blocks[name]:
            return None
        r = np.concatenate([b[0] for b in blocks[name]])
        c = np.concatenate([b[1] for b in blocks[name]])
        v = np.concatenate([b[2] for b in blocks[name]])
        return _coo(r, c, v, N)

    A0, B10, B20, S0 = build("A"), build("B1"), build("B2"), build("S")

    lo, hi = domain
    if bc == "channel_flow":
        dofs, vals = channel_flow_bc(mesh, iu1, iu2, ip, lo, hi)
    elif bc == "lower":
        dofs, vals = lower_bc(mesh, iu1, iu2, ip, lo)
    elif bc in (None, "none"):
        dofs, vals = np.zeros(0, dtype=np.int64), np.zeros(0)
    else:
        raise ValueError(bc)

    pos = np.zeros((N, 2))
    pos[iu1], pos[iu2], pos[ip] = mesh.p2_xy, mesh.p2_xy, mesh.p1_xy
    fix = lambda K: None if K is None else apply_dirichlet(K, dofs)
    return StokesOperators(
        mesh=mesh, N=N, idx_u1=iu1, idx_u2=iu2, idx_p=ip,
        A=fix(A0), B1=fix(B10), B2=fix(B20), S=fix(S0),
        bc_dofs=dofs, bc_vals=vals, pos=pos,
        A_nobc=A0 if keep_nobc else None, B1_nobc=B10 if keep_nobc else None,
    )


def spai(A: np.ndarray, m: int, start: str = "onenormest") -> np.ndarray:
    """Minimal-residual sparse-approximate-inverse iteration, dense restatement of the reference's
    `spai` (`FEONet_Stokes_square/train_FEONet.py:104-121`): M <- M + a (I - A M), started from 2 / ||A A^T||_1 * A with the
    reference's 1-norm estimate (scipy `onenormest`) or, start="exact", the exact norm."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if start == "onenormest":
        from scipy.sparse.linalg import onenormest

        norm1 = float(onenormest(A @ A.T))
    else:
        norm1 = float(np.linalg.norm(A @ A.T, 1))
    M = (2.0 / norm1) * A
    eye = np.eye(n)
    for _ in range(m):
        G = eye - A @ M
        AG = A @ G
        M = M + (np.sum(G * AG) / np.sum(AG * AG)) * G
    return M


# convenience constructors for the BASELINE.json configs (SURVEY.md section 8 size table) ----------
def config_operators(name: str, n: Optional[int] = None, ordering: str = "blocked") -> StokesOperators:
    if name == "stokes_square":  # cfg1: n=6 -> N=387
        return assemble_operators(structured_mesh(n or 6), mu=0.1, bc="channel_flow", ordering=ordering)
    if name == "steady_ns":  # cfg3: n=15 -> N=2178 ; cfg5: n=333 -> N=1001334
        return assemble_operators(structured_mesh(n or 15), mu=0.1, bc="channel_flow", ordering=ordering,
                                  with_convection=True)
    if name == "time_dep":  # cfg4: n=10 -> N=1003, mu=1, +q div u
        return assemble_operators(structured_mesh(n or 10), mu=1.0, bc="channel_flow", ordering=ordering,
                                  q_sign=+1.0, with_mass=True)
    if name == "hole":  # cfg2 stand-in
        mesh = square_with_hole_mesh(n or 10)
        return assemble_operators(mesh, mu=0.1, bc="channel_flow", ordering=ordering, form="symgrad",
                                  q_sign=+1.0, domain=(-1.0, 1.0))
    raise ValueError(name)
