"""CPU: the oracle restatement reproduces the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference functions)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_cases
from oracle import feonet_oracle as orc

# fp32 restatement vs fp32 reference: only summation order differs
RTOL_LOSS, RTOL_VEC = 2e-6, 2e-5


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def _relerr(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize("name", golden_cases("ns_"))
def test_steady_ns(name):
    g = _load(name)
    dp = bool(g["do_precond"])
    LHS, RHS = orc.ns_weak_form(g["alpha"], g["F"], g["A"], g["B1"], g["B2"], g["idx_u1"], g["idx_u2"], dp, g["P"])
    assert _relerr(LHS, g["LHS"]) < RTOL_VEC and _relerr(RHS, g["RHS"]) < RTOL_VEC
    loss, grad, _ = orc.ns_loss_and_grad(g["alpha"], g["F"], g["A"], g["B1"], g["B2"], g["idx_u1"], g["idx_u2"], dp, g["P"])
    assert abs(loss - float(g["loss"])) <= RTOL_LOSS * abs(float(g["loss"]))
    assert _relerr(grad, g["grad"]) < RTOL_VEC
    # fp64 evaluation of the same formula is also within fp32 round-off of the reference
    loss64, grad64, _ = orc.ns_loss_and_grad(g["alpha"], g["F"], g["A"], g["B1"], g["B2"], g["idx_u1"], g["idx_u2"], dp, g["P"], dtype=np.float64)
    assert abs(loss64 - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert _relerr(grad64, g["grad"]) < 1e-4
    u = orc.precond_output(g["alpha"][:, None, :], g["P"], dp)
    assert _relerr(u, g["u_pred"]) < RTOL_VEC


@pytest.mark.parametrize("name", golden_cases("stokes_") + golden_cases("hole_"))
def test_linear_stokes(name):
    g = _load(name)
    dp = bool(g["do_precond"])
    LHS, RHS = orc.stokes_weak_form(g["alpha"][:, None, :], g["F"], g["A"], g["P"], dp)
    assert _relerr(LHS, g["LHS"]) < RTOL_VEC and np.array_equal(RHS, g["RHS"])
    loss, grad, _ = orc.stokes_loss_and_grad(g["alpha"], g["F"], g["A"], g["P"], dp)
    assert abs(loss - float(g["loss"])) <= RTOL_LOSS * abs(float(g["loss"]))
    assert _relerr(grad, g["grad"]) < RTOL_VEC
    assert _relerr(orc.precond_output(g["alpha"][:, None, :], g["P"], dp), g["u_pred"]) < RTOL_VEC


@pytest.mark.parametrize("name", golden_cases("timedep_"))
def test_time_dep(name):
    g = _load(name)
    dp = bool(g["do_precond"])
    N = g["A"].shape[0]
    u0 = orc.assemble_u_init(g["init_x"], g["init_y"], g["idx_u1"], g["idx_u2"], N)
    assert np.array_equal(u0, g["u_init"])  # index scatter: bit-exact
    LHS, RHS, _, _ = orc.seq_weak_form(g["pred"], g["F"], g["S"], g["A"], g["P"], float(g["dt"]), u0, dp)
    assert _relerr(LHS, g["LHS"]) < RTOL_VEC and _relerr(RHS, g["RHS"]) < RTOL_VEC
    loss, grad, _ = orc.seq_loss_and_grad(g["pred"], g["F"], g["S"], g["A"], g["P"], float(g["dt"]), u0, dp)
    assert abs(loss - float(g["loss"])) <= RTOL_LOSS * abs(float(g["loss"]))
    assert _relerr(grad, g["grad"]) < RTOL_VEC
    assert _relerr(orc.precond_output(g["pred"], g["P"], dp), g["u_pred"]) < RTOL_VEC


def test_sparse_matches_dense():
    """The scipy-CSR form used at ~1M dofs equals the dense form."""
    import scipy.sparse as sp

    g = _load("ns_noprecond_n4")
    dense = orc.ns_loss_and_grad(g["alpha"], g["F"], g["A"], g["B1"], g["B2"], g["idx_u1"], g["idx_u2"], False, dtype=np.float64)
    csr = orc.ns_loss_and_grad(g["alpha"], g["F"], sp.csr_matrix(g["A"]), sp.csr_matrix(g["B1"]), sp.csr_matrix(g["B2"]),
                               g["idx_u1"], g["idx_u2"], False, dtype=np.float64)
    assert abs(dense[0] - csr[0]) < 1e-12 * abs(dense[0]) and _relerr(csr[1], dense[1]) < 1e-12


@pytest.mark.parametrize("name", golden_cases("ns_") + golden_cases("stokes_") + golden_cases("hole_"))
def test_reference_loops_restatement_is_bit_equal(name):
    """oracle.TorchReferenceLoops (the reference's own execution plan, timed by bench.py --configs on the GPU box's host)
    runs the same torch ops in the same order as the unmodified reference functions: loss and gradient of the goldens are
    reproduced BIT FOR BIT (same torch build, same thread count as when the goldens were made is not required: the per-dof
    loss loop and the dense GEMMs are deterministic on one host; a tolerance of one ulp-scale relative error is allowed for
    GEMM blocking differences across CPUs)."""
    g = _load(name)
    dp = bool(g["do_precond"])
    loops = orc.TorchReferenceLoops()
    if name.startswith("ns_"):
        loss, grad = loops.steady_ns_step(g["alpha"], g["F"], g["A"], g["B1"], g["B2"], g["idx_u1"], g["idx_u2"], dp, g["P"] if dp else None)
    else:
        loss, grad = loops.linear_stokes_step(g["alpha"], g["F"], g["A"], g["P"], dp)
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert _relerr(grad.numpy(), g["grad"]) < 1e-6


@pytest.mark.parametrize("name", golden_cases("spai_"))
def test_spai_restatement_matches_the_reference(name):
    """fixtures.spai (host restatement, onenormest start) against the output of the reference's own `spai`
    (FEONet_Stokes_square/train_FEONet.py:104-121), AST-extracted and run by oracle/make_golden.py."""
    from feonet_navier_stokes_b200.fixtures import spai

    g = _load(name)
    M = spai(g["A"], int(g["m"]))
    assert _relerr(M, g["M"]) < 1e-12
    assert abs(np.linalg.norm(np.eye(M.shape[0]) - g["A"] @ M) - float(g["residual"])) < 1e-10
