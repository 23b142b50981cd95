"""CPU: the C-ABI library loads, exports every declared symbol, and its host-side set-up code
(union pattern, partner lookups, tiles, staging boxes, forward/backward operator streams) agrees with the oracle.
No CUDA call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from feonet_navier_stokes_b200 import _lib as L
from feonet_navier_stokes_b200.fixtures import config_operators
from feonet_navier_stokes_b200.operator import build_desc, to_host_csr
from oracle import feonet_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = L.load_library()
    header = open(os.path.join(ROOT, "include", "feonet_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|size_t|const char\*)\s+(feo_\w+)\(", header, flags=re.M))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.feo_abi_version() == L.FEO_ABI_VERSION


def test_to_host_csr_threshold_zero():
    K = np.array([[1.0, 0.0, 1e-30], [0.0, 0.0, 0.0], [1e-60, 2.0, -3.0]])  # 1e-60 -> 0 in fp32
    rowptr, col, val = to_host_csr(K)
    assert rowptr.tolist() == [0, 2, 2, 4] and col.tolist() == [0, 2, 1, 2]
    assert val.dtype == np.float32 and val[1] == np.float32(1e-30)


def _replay(lib, desc, backward, in0, in1, n, max_lines=0, warps=0):
    out = np.full(n, np.nan)
    stats = (C.c_int64 * 8)()
    L.check(lib.feo_debug_tile_replay(C.byref(desc), int(backward), max_lines, warps, in0.ctypes.data_as(L.f64p),
                                      in1.ctypes.data_as(L.f64p), out.ctypes.data_as(L.f64p), stats))
    return out, list(stats)


@pytest.mark.parametrize("n,ordering,branch,max_lines,warps", [
    (3, "blocked", 1, 0, 0), (3, "interleaved", 0, 0, 0), (5, "interleaved", 1, 96, 3), (6, "blocked", 0, 96, 16),
    (8, "interleaved", 1, 128, 5)])
def test_tile_replay_matches_oracle_ns(n, ordering, branch, max_lines, warps):
    """The staging boxes and the per-warp streams of the fused kernels, decoded on the host as the kernels
    decode them, reproduce the oracle's residual and gradient (fp64, one sample)."""
    lib = L.load_library()
    op = config_operators("steady_ns", n, ordering=ordering)
    desc, keep = build_desc(op.N, op.A, op.B1, op.B2, None, op.idx_u1, op.idx_u2, bool(branch))
    rng = np.random.default_rng(n)
    alpha = rng.standard_normal(op.N)
    f = rng.standard_normal(op.N)
    A32, B132, B232 = (K.astype(np.float32).astype(np.float64) for K in (op.A, op.B1, op.B2))
    lo, go, ro = orc.ns_loss_and_grad(alpha[None], f[None], A32, B132, B232, op.idx_u1, op.idx_u2, bool(branch), dtype=np.float64)
    r, st_f = _replay(lib, desc, False, alpha, f, op.N, max_lines, warps)
    U = abs(op.A.astype(np.float32)) + abs(op.B1.astype(np.float32)) + abs(op.B2.astype(np.float32))
    assert st_f[5] == U.nnz  # every stored entry of the union pattern appears exactly once
    assert np.allclose(r, ro[0], rtol=1e-11, atol=1e-12)
    assert abs(np.sum(r * r) - lo) < 1e-11 * abs(lo)
    g, st_b = _replay(lib, desc, True, ro[0], alpha, op.N, max_lines, warps)
    assert np.allclose(2.0 * g, go[0], rtol=1e-10, atol=1e-11)
    if max_lines:
        assert st_f[1] <= max_lines and st_b[1] <= max_lines and st_f[0] > 1


def test_tile_replay_noncolocated_pairs_and_cross_terms():
    """I, J are opaque lists (quirk 4): shuffle the pairing, and add B1/B2 entries that couple the two
    components and the pressure columns so that the generic V-/X-step paths are exercised."""
    lib = L.load_library()
    op = config_operators("steady_ns", 4, ordering="interleaved")
    rng = np.random.default_rng(7)
    idx_i = np.asarray(op.idx_u1).copy()
    idx_j = np.asarray(op.idx_u2).copy()
    perm = rng.permutation(len(idx_j))
    idx_j = idx_j[perm]  # pairs are no longer the two components of one node
    import scipy.sparse as sp
    noise = sp.random(op.N, op.N, density=0.01, random_state=3, data_rvs=rng.standard_normal).tocsr()
    B1 = (op.B1 + noise).tocsr()
    B2 = (op.B2 + noise.T).tocsr()
    desc, keep = build_desc(op.N, op.A, B1, B2, None, idx_i, idx_j, True)
    alpha, f = rng.standard_normal(op.N), rng.standard_normal(op.N)
    A32, B132, B232 = (K.astype(np.float32).astype(np.float64) for K in (op.A, B1, B2))
    lo, go, ro = orc.ns_loss_and_grad(alpha[None], f[None], A32, B132, B232, idx_i, idx_j, True, dtype=np.float64)
    r, _ = _replay(lib, desc, False, alpha, f, op.N, 96, 4)
    assert np.allclose(r, ro[0], rtol=1e-11, atol=1e-12)
    g, _ = _replay(lib, desc, True, ro[0], alpha, op.N, 200, 4)
    assert np.allclose(2.0 * g, go[0], rtol=1e-10, atol=1e-11)


def test_tile_replay_linear_and_small_tiles():
    lib = L.load_library()
    op = config_operators("hole", 5)
    desc, keep = build_desc(op.N, op.A)
    rng = np.random.default_rng(0)
    alpha, f = rng.standard_normal(op.N), rng.standard_normal(op.N)
    A32 = op.A.astype(np.float32).astype(np.float64)
    lo, go, _ = orc.stokes_loss_and_grad(alpha[None], f[None], A32, dtype=np.float64)
    r, st = _replay(lib, desc, False, alpha, f, op.N, 48, 2)
    assert st[0] > 4 and np.allclose(r, A32 @ alpha - f, rtol=1e-11, atol=1e-12)
    g, _ = _replay(lib, desc, True, r, alpha, op.N, 48, 2)
    assert np.allclose(2.0 * g, go[0], rtol=1e-10, atol=1e-11)


def test_bad_inputs_are_rejected():
    lib = L.load_library()
    op = config_operators("steady_ns", 2)
    bad_i = op.idx_u1.copy()
    bad_i[0] = op.idx_u2[0]  # I and J overlap
    desc, keep = build_desc(op.N, op.A, op.B1, op.B2, None, bad_i, op.idx_u2, True)
    assert lib.feo_debug_tile_replay(C.byref(desc), 0, 0, 0, None, None, None, None) == -3
    assert b"disjoint" in lib.feo_last_error_string()
    desc, keep = build_desc(op.N, op.A, op.B1, None)
    assert lib.feo_debug_tile_replay(C.byref(desc), 0, 0, 0, None, None, None, None) == -1
    # a row that touches more lines than a tile can stage is refused, not mis-computed
    desc, keep = build_desc(op.N, np.ones((op.N, op.N)))
    assert lib.feo_debug_tile_replay(C.byref(desc), 0, 32, 0, None, None, None, None) == -3


@pytest.mark.parametrize("n,transposed", [(5, 0), (72, 1), (130, 0), (387, 1)])
def test_dense_split_tiles_replay(n, transposed):
    """Set-up of the tensor-core preconditioner GEMM (feo_dense_tc.cu): the operator is split into TF32 hi/lo halves
    and laid out per (128-row tile, 16-column k-block) in the kernel's shared-memory operand layout.  Decoded the way
    the kernel addresses it: hi + lo reproduces D x to the split error (2^-22 per entry), hi alone only to TF32
    (2^-11), both halves are exact TF32 numbers, and the zero padding of ragged tiles holds zeros."""
    lib = L.load_library()
    rng = np.random.default_rng(n)
    D = rng.standard_normal((n, n)).astype(np.float32)
    D[rng.random((n, n)) < 0.2] = 0.0
    x = rng.standard_normal(n)
    hi, lo = np.empty(n), np.empty(n)
    size = lib.feo_debug_dense_split_replay(D.ctypes.data_as(L.f32p), n, transposed, x.ctypes.data_as(L.f64p),
                                            hi.ctypes.data_as(L.f64p), lo.ctypes.data_as(L.f64p))
    assert size == ((n + 127) // 128) * ((n + 15) // 16) * 4096
    M = (D.T if transposed else D).astype(np.float64)
    bound = np.abs(M) @ np.abs(x)
    assert (np.abs(hi + lo - M @ x) <= 2.0 ** -21 * bound + 1e-300).all()
    assert np.abs(hi - M @ x).max() > 2.0 ** -16 * bound.max() or n < 8  # the lo half matters
    assert (np.abs(hi - M @ x) <= 2.0 ** -11 * bound + 1e-300).all()
    # statistics-only call and argument errors
    assert lib.feo_debug_dense_split_replay(D.ctypes.data_as(L.f32p), n, 0, None, None, None) == size
    assert lib.feo_debug_dense_split_replay(None, n, 0, None, None, None) < 0


def _lattice_replay(lib, desc, backward, in0, in1, n):
    out = np.full(n, np.nan)
    stats = (C.c_int64 * 8)()
    rc = lib.feo_debug_lattice_replay(C.byref(desc), int(backward), in0.ctypes.data_as(L.f64p), in1.ctypes.data_as(L.f64p),
                                      out.ctypes.data_as(L.f64p), stats)
    return rc, out, list(stats)


@pytest.mark.parametrize("n,branch,variant", [(2, 1, "steady_ns"), (3, 0, "steady_ns"), (6, 1, "steady_ns"), (9, 0, "steady_ns"),
                                              (5, 1, "stokes_square")])
def test_lattice_replay_matches_oracle(n, branch, variant):
    """Lattice plan (feo_lattice.h): the class tables filled from the CSR and the GENERATED cell bodies (the same macro lists
    the CUDA kernels compile), run in fp64 on the host, reproduce the oracle's residual and gradient -- interior, edge and
    corner cells, identity Dirichlet rows, both sign branches, and the linear Stokes operator (no B1 / B2)."""
    lib = L.load_library()
    op = config_operators(variant, n, ordering="interleaved")
    has_b = op.B1 is not None
    # a linear operator may come without idx_sol (FEOperator(N, A=...)): the lattice is then recognised from N and A alone
    desc, keep = build_desc(op.N, op.A, op.B1, op.B2, None, op.idx_u1 if has_b else None, op.idx_u2 if has_b else None, bool(branch))
    rng = np.random.default_rng(100 + n)
    alpha = rng.standard_normal(op.N)
    f = rng.standard_normal(op.N)
    A32 = op.A.astype(np.float32).astype(np.float64)
    if has_b:
        B132, B232 = (K.astype(np.float32).astype(np.float64) for K in (op.B1, op.B2))
        lo, go, ro = orc.ns_loss_and_grad(alpha[None], f[None], A32, B132, B232, op.idx_u1, op.idx_u2, bool(branch), dtype=np.float64)
        ro, go = ro[0], go[0]
    else:
        ro = A32 @ alpha - f
        go = 2.0 * (A32.T @ ro)
    rc, r, st = _lattice_replay(lib, desc, False, alpha, f, op.N)
    assert rc == 0 and st[0] == 1 and st[1] == n, (rc, st, lib.feo_last_error_string())
    assert st[2] <= 9 and st[3] <= 32
    assert np.allclose(r, ro, rtol=1e-11, atol=1e-12)
    rc, g, _ = _lattice_replay(lib, desc, True, ro, alpha, op.N)
    assert rc == 0
    assert np.allclose(2.0 * g, go, rtol=1e-10, atol=1e-11)
    # table 2: the forward rows as a matrix-free element walk (every assembled entry split over the elements that sum to it;
    # the A/B variant of DESIGN.md section 3.5) -- same residual up to the fp32 rounding of the split coefficients
    rc, r2, _ = _lattice_replay(lib, desc, 2, alpha, f, op.N)
    assert rc == 0
    assert np.allclose(r2, ro, rtol=2e-6, atol=2e-6)


def test_lattice_plan_rejects_what_it_does_not_model():
    """Blocked dof order, shuffled (I, J) pairings, an unstructured mesh and a perturbed entry outside the stencil are not
    lattices: the planner says so (the tile plan is used for them), it never mis-evaluates."""
    lib = L.load_library()
    x = np.zeros(1)

    def applicable(op, idx_i=None, idx_j=None, A=None):
        desc, keep = build_desc(op.N, op.A if A is None else A, op.B1, op.B2, None, op.idx_u1 if idx_i is None else idx_i,
                                op.idx_u2 if idx_j is None else idx_j, True)
        stats = (C.c_int64 * 8)()
        rc = lib.feo_debug_lattice_replay(C.byref(desc), 0, None, None, None, stats)
        return rc == 0 and stats[0] == 1

    good = config_operators("steady_ns", 4, ordering="interleaved")
    assert applicable(good)
    assert not applicable(config_operators("steady_ns", 4, ordering="blocked"))
    rng = np.random.default_rng(3)
    assert not applicable(good, idx_j=np.asarray(good.idx_u2)[rng.permutation(len(good.idx_u2))])
    assert not applicable(config_operators("hole", None, ordering="interleaved"))
    A = good.A.tolil(copy=True)
    A[int(good.idx_u1[0]), int(good.idx_u1[-1])] = 0.5  # couples two far-away nodes
    assert not applicable(good, A=A.tocsr())
    A = good.A.tolil(copy=True)
    k = len(good.idx_u1) // 2
    A[int(good.idx_u1[k]), int(good.idx_u2[k])] = 0.25  # cross-component entry
    assert not applicable(good, A=A.tocsr())


def test_lattice_permutation_from_dof_coordinates():
    """reorder.lattice_permutation: from idx_sol and the dof coordinates the reference stores (`p`), the renumbering that
    turns a blocked (or arbitrarily relabelled) structured operator into exactly the interleaved one the lattice plan models;
    anything that is not a full structured lattice is refused."""
    import scipy.sparse as sp

    from feonet_navier_stokes_b200.fixtures import config_operators
    from feonet_navier_stokes_b200.reorder import is_identity, lattice_permutation, permute_csr, permute_dense

    fb = config_operators("steady_ns", 5, ordering="blocked")
    fi = config_operators("steady_ns", 5, ordering="interleaved")
    perm = lattice_permutation(fb.idx_sol, fb.pos)
    assert perm is not None and not is_identity(perm)
    for Kb, Ki in ((fb.A, fi.A), (fb.B1, fi.B1), (fb.B2, fi.B2)):
        P, Q = permute_csr(Kb, perm), sp.csr_matrix(Ki).astype(np.float32)
        Q.eliminate_zeros()
        Q.sort_indices()
        assert np.array_equal(P.indptr, Q.indptr) and np.array_equal(P.indices, Q.indices) and np.array_equal(P.data, Q.data)
    assert np.array_equal(perm[np.asarray(fb.idx_u1)], np.asarray(fi.idx_u1)) and np.array_equal(perm[np.asarray(fb.idx_p)], np.asarray(fi.idx_p))
    assert is_identity(lattice_permutation(fi.idx_sol, fi.pos))
    D = np.arange(fb.N * fb.N, dtype=np.float32).reshape(fb.N, fb.N)
    assert np.array_equal(permute_dense(D, perm)[perm[3], perm[7]], D[3, 7])
    # refusals: an unstructured mesh, a non-colocated pairing, coordinates off the lattice
    fh = config_operators("hole", 6)
    assert lattice_permutation(fh.idx_sol, fh.pos) is None
    J = np.asarray(fb.idx_u2).copy()
    J[[0, 1]] = J[[1, 0]]
    assert lattice_permutation([fb.idx_u1, J, fb.idx_p], fb.pos) is None
    pos = fb.pos.copy()
    pos[3, 0] += 0.01
    assert lattice_permutation(fb.idx_sol, pos) is None


def test_numa_binding_is_a_no_op_without_a_visible_topology():
    """parallel.bind_to_gpu_numa_node leaves the affinity alone when there is no GPU / no sysfs topology to read."""
    import os

    from feonet_navier_stokes_b200.parallel import bind_to_gpu_numa_node

    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node("cuda:0") is None or isinstance(bind_to_gpu_numa_node("cuda:0"), int)
    import torch

    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
