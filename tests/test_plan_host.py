"""CPU: the C-ABI library loads, exports every declared symbol, and its host-side set-up code
(union pattern, partner lookups, blobs, forward/backward entry streams) agrees with the oracle.
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
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(feo_\w+)\(", header, flags=re.M))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.feo_abi_version() == L.FEO_ABI_VERSION


def test_to_host_csr_threshold_zero():
    K = np.array([[1.0, 0.0, 1e-30], [0.0, 0.0, 0.0], [1e-60, 2.0, -3.0]])  # 1e-60 -> 0 in fp32
    rowptr, col, val = to_host_csr(K)
    assert rowptr.tolist() == [0, 2, 2, 4] and col.tolist() == [0, 2, 1, 2]
    assert val.dtype == np.float32 and val[1] == np.float32(1e-30)


@pytest.mark.parametrize("n,ordering,branch", [(3, "blocked", 1), (3, "interleaved", 0), (5, "interleaved", 1), (6, "blocked", 0)])
def test_plan_replay_matches_oracle_ns(n, ordering, branch):
    lib = L.load_library()
    op = config_operators("steady_ns", n, ordering=ordering)
    desc, keep = build_desc(op.N, op.A, op.B1, op.B2, None, op.idx_u1, op.idx_u2, bool(branch))
    stats = (C.c_int64 * 8)()
    L.check(lib.feo_debug_plan_check(C.byref(desc), stats))
    U = abs(op.A.astype(np.float32)) + abs(op.B1.astype(np.float32)) + abs(op.B2.astype(np.float32))
    assert stats[2] == U.nnz and stats[7] == 1 and stats[1] == op.mesh.n_u + op.mesh.n_p
    rng = np.random.default_rng(n)
    alpha = rng.standard_normal(op.N)
    f = rng.standard_normal(op.N)
    r = np.zeros(op.N)
    g = np.zeros(op.N)
    loss = C.c_double()
    L.check(lib.feo_debug_plan_replay(C.byref(desc), alpha.ctypes.data_as(L.f64p), f.ctypes.data_as(L.f64p),
                                      r.ctypes.data_as(L.f64p), g.ctypes.data_as(L.f64p), C.byref(loss)))
    A32, B132, B232 = (K.astype(np.float32).astype(np.float64) for K in (op.A, op.B1, op.B2))
    lo, go, ro = orc.ns_loss_and_grad(alpha[None], f[None], A32, B132, B232, op.idx_u1, op.idx_u2, bool(branch), dtype=np.float64)
    assert abs(loss.value - lo) < 1e-11 * abs(lo)
    assert np.allclose(r, ro[0], rtol=1e-11, atol=1e-12) and np.allclose(g, go[0], rtol=1e-10, atol=1e-11)


def test_plan_replay_linear_and_small_blobs(monkeypatch):
    lib = L.load_library()
    monkeypatch.setenv("FEO_BLOB_ROWS", "7")
    op = config_operators("hole", 5)
    desc, keep = build_desc(op.N, op.A)
    stats = (C.c_int64 * 8)()
    L.check(lib.feo_debug_plan_check(C.byref(desc), stats))
    assert stats[7] == 0 and stats[0] >= op.N // 7 and stats[1] == op.N
    rng = np.random.default_rng(0)
    alpha, f = rng.standard_normal(op.N), rng.standard_normal(op.N)
    r, g, loss = np.zeros(op.N), np.zeros(op.N), C.c_double()
    L.check(lib.feo_debug_plan_replay(C.byref(desc), alpha.ctypes.data_as(L.f64p), f.ctypes.data_as(L.f64p),
                                      r.ctypes.data_as(L.f64p), g.ctypes.data_as(L.f64p), C.byref(loss)))
    A32 = op.A.astype(np.float32).astype(np.float64)
    lo, go, _ = orc.stokes_loss_and_grad(alpha[None], f[None], A32, dtype=np.float64)
    assert abs(loss.value - lo) < 1e-11 * abs(lo) and np.allclose(g, go[0], rtol=1e-10, atol=1e-11)


def test_bad_inputs_are_rejected():
    lib = L.load_library()
    op = config_operators("steady_ns", 2)
    bad_i = op.idx_u1.copy()
    bad_i[0] = op.idx_u2[0]  # I and J overlap
    desc, keep = build_desc(op.N, op.A, op.B1, op.B2, None, bad_i, op.idx_u2, True)
    assert lib.feo_debug_plan_check(C.byref(desc), None) == -3
    assert b"disjoint" in lib.feo_last_error_string()
    desc, keep = build_desc(op.N, op.A, op.B1, None)
    assert lib.feo_debug_plan_check(C.byref(desc), None) == -1
