"""GPU parity tests of the lattice plan (feo_lattice.h): the plan feo_op_create picks for structured right-diagonal P2-P1
operators in lattice order -- BASELINE.json configs[4] and the reference's RectangleMesh set-ups -- against the fp64 oracle
(FEONet_steady_Navier-Stokes/train_FEONet.py:301-365) and against the tile plan on the same inputs.
Tolerances as north_star: loss 1e-5, gradient 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import feonet_oracle as orc

pytestmark = pytest.mark.gpu
LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def feo():
    import feonet_navier_stokes_b200 as f

    assert torch.cuda.is_available()
    f.load_library(build_if_missing=False)
    return f


def _run(feo, fx, alpha, F, branch, native=True):
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=bool(branch), device=dev)
    a = torch.tensor(alpha, device=dev)
    a = feo.to_dof_major_tensor(a) if native else a.unsqueeze(1)
    a.requires_grad_(True)
    loss = ns.residual_loss(a, torch.tensor(F, device=dev), fx.A, fx.B1, fx.B2, fx.idx_sol)
    (g,) = torch.autograd.grad(loss, a)
    return loss.item(), g.reshape(alpha.shape).cpu().numpy(), ns


@pytest.mark.parametrize("n,B,branch", [(4, 1, 1), (4, 65, 0), (7, 2, 0), (7, 63, 1), (12, 130, 1), (20, 37, 0), (40, 200, 1), (33, 64, 0)])
def test_lattice_vs_oracle_and_tile_plan(feo, monkeypatch, n, B, branch):
    """Ragged batches (B = 1, 2, 37, 63, 65, 130: lanes past B, slabs past ldb), meshes from one strip (n = 4: 5 cells per
    side, every cell a boundary class) to several strips with partly filled last strips (n = 20, 33, 40), both sign branches."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", n, ordering="interleaved")
    rng = np.random.default_rng(1000 * n + B)
    alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, bool(branch), dtype=np.float64)
    monkeypatch.setenv("FEO_PLAN", "lattice")  # fails loudly if the operator were not recognised
    l1, g1, ns = _run(feo, fx, alpha, F, branch)
    assert ns.operator.info.n_tiles_fwd == (n + 1) ** 2  # the lattice plan reports its cells
    assert abs(l1 - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(g1, go) < GRAD_RTOL and _relmax(g1, go) < GRAD_RTOL
    l1b, g1b, _ = _run(feo, fx, alpha, F, branch, native=False)  # the reference's row-major [B,1,N] boundary
    assert l1b == l1 and np.array_equal(g1b, g1)
    monkeypatch.setenv("FEO_PLAN", "tile")
    l2, g2, ns2 = _run(feo, fx, alpha, F, branch)
    assert ns2.operator.info.n_tiles_fwd != (n + 1) ** 2 or n < 3
    assert abs(l1 - l2) <= 2e-6 * abs(lo) and _rel(g1, g2) < 2e-6  # two device code paths, same formula: fp32 round-off apart


def test_lattice_is_bit_reproducible_and_width_independent(feo, monkeypatch):
    """Row- / column-owned arithmetic with a fixed order: the same bits on every launch, for every strip width and ring depth."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 18, ordering="interleaved")
    rng = np.random.default_rng(5)
    alpha = (0.3 * rng.standard_normal((96, fx.N))).astype(np.float32)
    F = rng.standard_normal((96, fx.N)).astype(np.float32)
    monkeypatch.setenv("FEO_PLAN", "lattice")
    l0, g0, _ = _run(feo, fx, alpha, F, 1)
    l1, g1, _ = _run(feo, fx, alpha, F, 1)
    assert l0 == l1 and np.array_equal(g0, g1)
    for wf, wb, rf in (("5", "4", "7"), ("9", "7", "9"), ("15", "11", "7")):
        monkeypatch.setenv("FEO_LAT_W_FWD", wf)
        monkeypatch.setenv("FEO_LAT_W_BWD", wb)
        monkeypatch.setenv("FEO_LAT_R_FWD", rf)
        l2, g2, _ = _run(feo, fx, alpha, F, 1)
        assert np.array_equal(g0, g2)  # per-cell arithmetic does not depend on the strip decomposition
        assert abs(l2 - l0) <= 1e-6 * abs(l0)  # the loss partials are grouped per (CTA, warp)


def test_lattice_linear_stokes_without_idx_sol(feo, monkeypatch):
    """A linear operator handed over without idx_sol (FEOperator(N, A=...)): recognised from N and A alone; r = A a - F,
    grad = 2 A^T r (FEONet_Stokes_square/train_FEONet.py:261-301 without a preconditioner)."""
    from feonet_navier_stokes_b200.fixtures import config_operators
    from feonet_navier_stokes_b200.operator import FEOperator

    fx = config_operators("stokes_square", 16, ordering="interleaved")
    monkeypatch.setenv("FEO_PLAN", "lattice")
    dev = torch.device("cuda")
    op = FEOperator(fx.N, A=fx.A, device=dev)
    assert op.info.n_tiles_fwd == 17 ** 2
    rng = np.random.default_rng(9)
    B = 70
    alpha = rng.standard_normal((B, fx.N)).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    ldb = 72
    aT = torch.zeros(fx.N, ldb, device=dev)
    fT = torch.zeros(fx.N, ldb, device=dev)
    aT[:, :B] = torch.tensor(alpha, device=dev).t()
    fT[:, :B] = torch.tensor(F, device=dev).t()
    loss, rT = op.residual_fwd(aT, fT, B)
    gT = torch.empty_like(aT)
    op.residual_bwd(aT, rT, B, out=gT)
    A64 = fx.A.astype(np.float32).astype(np.float64)
    r = (A64 @ alpha.astype(np.float64).T).T - F
    lo = float((r * r).sum())
    go = 2.0 * (A64.T @ r.T).T
    assert abs(loss.item() - lo) <= LOSS_RTOL * lo
    assert _rel(rT[:, :B].t().cpu().numpy(), r) < 1e-5
    assert _rel(gT[:, :B].t().cpu().numpy(), go) < GRAD_RTOL


def test_element_walk_variant_matches(feo, monkeypatch):
    """FEO_LATTICE_ELEMENT=1: the forward kernel evaluates the rows as a matrix-free element walk in gather form (one FMA per
    element-level entry, DESIGN.md section 3.5: the measured alternative to the assembled stencil).  Same loss / residual."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 14, ordering="interleaved")
    rng = np.random.default_rng(77)
    alpha = (0.3 * rng.standard_normal((70, fx.N))).astype(np.float32)
    F = rng.standard_normal((70, fx.N)).astype(np.float32)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, True, dtype=np.float64)
    monkeypatch.setenv("FEO_PLAN", "lattice")
    monkeypatch.setenv("FEO_LATTICE_ELEMENT", "1")
    l1, g1, _ = _run(feo, fx, alpha, F, 1)
    assert abs(l1 - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(g1, go) < GRAD_RTOL


@pytest.mark.parametrize("n,B,branch", [(6, 9, 1), (13, 70, 0)])
def test_any_dof_order_reaches_the_lattice_kernels(feo, n, B, branch):
    """An operator in another dof numbering (blocked [u1 | u2 | p], and a random relabelling of it -- FEniCS' mixed-space order
    is neither) with the dof coordinates the reference stores (`p`, assemble_fenics.py:125, :138): renumbered at set-up
    (reorder.py), permutation folded into the layout passes.  Results in the CALLER's numbering: equal to the oracle on the
    caller's operator and, bit for bit, to the interleaved operator's results on the permuted inputs."""
    from feonet_navier_stokes_b200.fixtures import config_operators
    from feonet_navier_stokes_b200.reorder import lattice_permutation, permute_csr

    dev = torch.device("cuda")
    rng = np.random.default_rng(7 * n + B)
    fb = config_operators("steady_ns", n, ordering="blocked")
    fi = config_operators("steady_ns", n, ordering="interleaved")
    N = fb.N
    shuffle = rng.permutation(N)  # new label of blocked dof d
    cases = [(fb.A, fb.B1, fb.B2, fb.idx_sol, fb.pos)]
    inv = np.argsort(shuffle)
    cases.append((permute_csr(fb.A, shuffle), permute_csr(fb.B1, shuffle), permute_csr(fb.B2, shuffle),
                  [shuffle[np.asarray(ix)] for ix in fb.idx_sol], fb.pos[inv]))
    for A, B1, B2, idx_sol, pos in cases:
        perm = lattice_permutation(idx_sol, pos)
        assert perm is not None
        alpha = (0.3 * rng.standard_normal((B, N))).astype(np.float32)
        F = rng.standard_normal((B, N)).astype(np.float32)
        lo, go, _ = orc.ns_loss_and_grad(alpha, F, A, B1, B2, np.asarray(idx_sol[0]), np.asarray(idx_sol[1]), bool(branch), dtype=np.float64)
        ns = feo.SteadyNavierStokes(A, B1, B2, idx_sol, do_precond=bool(branch), device=dev, dof_positions=pos)
        assert ns.operator.plan == "lattice"
        plain = feo.SteadyNavierStokes(A, B1, B2, idx_sol, do_precond=bool(branch), device=dev)
        assert plain.operator.plan == "tile"
        # the interleaved operator on the permuted inputs: the same kernels on the same numbers
        old_of_new = np.argsort(perm)
        li, gi, _ = _run(feo, fi, alpha[:, old_of_new], F[:, old_of_new], branch)
        for native in (False, True):
            a = torch.tensor(alpha, device=dev)
            a = feo.to_dof_major_tensor(a) if native else a.unsqueeze(1)
            a.requires_grad_(True)
            loss = ns.residual_loss(a, torch.tensor(F, device=dev), A, B1, B2, idx_sol)
            (g,) = torch.autograd.grad(loss, a)
            g = g.reshape(B, N).cpu().numpy()
            assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
            assert _rel(g, go) < GRAD_RTOL and _relmax(g, go) < GRAD_RTOL
            assert loss.item() == li and np.array_equal(g[:, old_of_new], gi)
        # the materialised weak_form outputs go through the same layout passes
        a = torch.tensor(alpha, device=dev).unsqueeze(1)
        lhs, rhs = ns.weak_form(a, torch.tensor(F, device=dev), A, B1, B2, idx_sol)
        lhs0, rhs0 = plain.weak_form(a, torch.tensor(F, device=dev), A, B1, B2, idx_sol)
        assert torch.allclose(lhs, lhs0, rtol=1e-5, atol=1e-5) and torch.allclose(rhs, rhs0, rtol=1e-5, atol=1e-5)


def test_linear_stokes_in_another_dof_order_reaches_the_lattice_kernels(feo):
    """The linear Stokes operator (FEONet_Stokes_square/train_FEONet.py:261-271) in blocked order with idx_sol and the dof
    coordinates: renumbered into the lattice plan, results in the caller's numbering against the oracle."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    dev = torch.device("cuda")
    fb = config_operators("stokes_square", 9, ordering="blocked")
    rng = np.random.default_rng(5)
    B = 45
    alpha = (0.3 * rng.standard_normal((B, fb.N))).astype(np.float32)
    F = rng.standard_normal((B, fb.N)).astype(np.float32)
    lo, go, _ = orc.stokes_loss_and_grad(alpha, F, fb.A, None, False, dtype=np.float64)
    st = feo.LinearStokes(fb.A, None, do_precond=False, device=dev, idx_sol=fb.idx_sol, dof_positions=fb.pos)
    assert st.operator.plan == "lattice"
    assert feo.LinearStokes(fb.A, None, do_precond=False, device=dev).operator.plan == "tile"
    a = torch.tensor(alpha, device=dev).unsqueeze(1).requires_grad_(True)
    loss = st.residual_loss(a, torch.tensor(F, device=dev), fb.A, None)
    (g,) = torch.autograd.grad(loss, a)
    assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(g.squeeze(1).cpu().numpy(), go) < GRAD_RTOL
