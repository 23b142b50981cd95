"""GPU parity tests: the CUDA path (through the C ABI) vs the reference's own outputs
(tests/golden) and vs the oracle on seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: fp32 loss within 1e-5 relative,
gradients within 1e-4 relative; integer/index work (layout transposes, assemble_u_init) bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, golden_cases
from oracle import feonet_oracle as orc

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _idx_sol(g):
    out = np.empty(3, dtype=object)
    out[0], out[1], out[2] = g["idx_u1"].tolist(), g["idx_u2"].tolist(), g["idx_p"].tolist()
    return out


@pytest.fixture(scope="module")
def feo():
    import feonet_navier_stokes_b200 as f

    assert torch.cuda.is_available()
    f.load_library(build_if_missing=False)  # the prebuilt in-tree .so must be the thing that runs
    return f


class Leaf(torch.nn.Module):
    def __init__(self, alpha):
        super().__init__()
        self.alpha = torch.nn.Parameter(alpha.clone())

    def forward(self, *a, **k):
        return self.alpha


# ------------------------------------------------------------------------------------------------
# golden vectors = outputs of the unmodified reference functions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases("ns_"))
def test_golden_steady_ns(feo, name):
    g = _load(name)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)  # noqa: E731
    A, B1, B2, P = t(g["A"]), t(g["B1"]), t(g["B2"]), t(g["P"])
    idx_sol = _idx_sol(g)
    ns = feo.SteadyNavierStokes(A, B1, B2, idx_sol, do_precond=bool(g["do_precond"]), precond=P, model_name="FCNN", device=dev)
    model = Leaf(t(g["alpha"]))
    loss, u_pred = ns.closure(model, torch.zeros(g["alpha"].shape[0], 6, device=dev), None, t(g["F"]), A, B1, B2, 8)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _rel(model.alpha.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL
    assert _relmax(model.alpha.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL
    assert _rel(u_pred.detach().cpu().numpy().reshape(g["u_pred"].shape), g["u_pred"]) < 1e-5
    # materialised weak_form, with autograd through it
    a2 = t(g["alpha"]).requires_grad_(True)
    LHS, RHS = ns.weak_form(a2.unsqueeze(1), t(g["F"]), A, B1, B2, idx_sol)
    assert _rel(LHS.detach().cpu().numpy(), g["LHS"]) < 1e-5 and _rel(RHS.detach().cpu().numpy(), g["RHS"]) < 1e-5
    torch.sum((LHS - RHS) ** 2).backward()
    assert _rel(a2.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL


@pytest.mark.parametrize("name", golden_cases("stokes_") + golden_cases("hole_"))
def test_golden_linear_stokes(feo, name):
    g = _load(name)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)  # noqa: E731
    A, P = t(g["A"]), t(g["P"])
    hole = name.startswith("hole_")
    st = feo.LinearStokes(A, P, do_precond=bool(g["do_precond"]), model_name="FCNN", hole_signature=hole, device=dev)
    model = Leaf(t(g["alpha"]))
    coeff = torch.zeros(g["alpha"].shape[0], 6, device=dev)
    args = (None, t(g["F"]), A, P, 8) if hole else (t(g["F"]), A, P, 8)
    loss, u_pred = st.closure(model, coeff, *args)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _rel(model.alpha.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL
    assert _relmax(model.alpha.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL
    assert _rel(u_pred.detach().cpu().numpy().reshape(g["u_pred"].shape), g["u_pred"]) < 1e-5
    LHS, RHS = st.weak_form(t(g["alpha"]).unsqueeze(1), t(g["F"]), A, P)
    assert _rel(LHS.cpu().numpy(), g["LHS"]) < 1e-5 and np.array_equal(RHS.cpu().numpy(), g["RHS"])


@pytest.mark.parametrize("name", golden_cases("timedep_"))
def test_golden_time_dep(feo, name):
    g = _load(name)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)  # noqa: E731
    S, A, P, dt = t(g["S"]), t(g["A"]), t(g["P"]), float(g["dt"])
    td = feo.TimeDependentStokes(S, A, _idx_sol(g), dt=dt, do_precond=bool(g["do_precond"]), precond=P, model_name="RNN", device=dev)

    class SeqLeaf(Leaf):
        def forward(self, u_init, seq_len=None):
            return self.alpha

    model = SeqLeaf(t(g["pred"]))
    ix, iy = t(g["init_x"]).unsqueeze(1), t(g["init_y"]).unsqueeze(1)
    u0 = td.assemble_u_init(ix, iy)
    assert np.array_equal(u0.cpu().numpy(), g["u_init"])  # index scatter: bit-exact
    T = g["pred"].shape[1]
    loss, out = td.closure(model, None, ix, iy, t(g["F"]), S, A, None, P, dt, T)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _rel(model.alpha.grad.cpu().numpy(), g["grad"]) < GRAD_RTOL
    assert _rel(out.detach().cpu().numpy(), g["u_pred"]) < 1e-5
    LHS, RHS = td.weak_form_sequence(t(g["pred"]), t(g["F"]), S, A, P, dt, u0, bool(g["do_precond"]))
    assert _rel(LHS.cpu().numpy(), g["LHS"]) < 1e-5 and _rel(RHS.cpu().numpy(), g["RHS"]) < 1e-5


# ------------------------------------------------------------------------------------------------
# one training-step evaluation against the reference's own closure + network + autograd (goldens made by oracle/make_golden.py
# from the unmodified reference functions and the reference's network.py): loss and EVERY parameter gradient
# ------------------------------------------------------------------------------------------------
def _load_state(model, g):
    sd = {k[len("state__"):].replace("__", "."): torch.tensor(g[k]) for k in g.files if k.startswith("state__")}
    model.load_state_dict(sd)
    return model


def _check_param_grads(model, g):
    worst = 0.0
    for k, p in model.named_parameters():
        ref = g["grad__" + k.replace(".", "__")]
        worst = max(worst, _rel(p.grad.cpu().numpy(), ref))
    return worst


@pytest.mark.parametrize("name", golden_cases("trainstep_ns_"))
@pytest.mark.parametrize("dof_major_head", [False, True])
def test_golden_trainstep_steady_ns(feo, name, dof_major_head):
    """closure (FEONet_steady_Navier-Stokes/train_FEONet.py:334-365) + loss.backward() (:463) with the FCNN of :178-179."""
    from feonet_navier_stokes_b200 import network

    g = _load(name)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)  # noqa: E731
    N = g["A"].shape[0]
    model = _load_state(network.FCNN(6, N, [int(h) for h in g["hidden"]], dof_major_head=dof_major_head), g).to(dev).eval()
    ns = feo.SteadyNavierStokes(g["A"], g["B1"], g["B2"], _idx_sol(g), do_precond=bool(g["do_precond"]), precond=None, model_name="FCNN", device=dev)
    loss, u_pred = ns.closure(model, t(g["coeff_f"]), None, t(g["F"]), g["A"], g["B1"], g["B2"], 8)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _rel(u_pred.detach().cpu().numpy().reshape(g["u_pred"].shape), g["u_pred"]) < 1e-5
    assert _check_param_grads(model, g) < GRAD_RTOL


@pytest.mark.parametrize("name", golden_cases("trainstep_timedep_"))
def test_golden_trainstep_time_dep(feo, name):
    """closure (FEONet_time_dep_Stokes/train_FEONet.py:364-406) + loss.backward() with VectorToSequenceRNN (network.py:342-399)."""
    from feonet_navier_stokes_b200 import network

    g = _load(name)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)  # noqa: E731
    N, dt, T = g["A"].shape[0], float(g["dt"]), int(g["T"])
    model = _load_state(network.VectorToSequenceRNN(input_dim=N, hidden_dim=int(g["hidden"][0]), output_dim=N), g).to(dev).train()  # no dropout inside; cuDNN's RNN backward wants train mode
    td = feo.TimeDependentStokes(g["S"], g["A"], _idx_sol(g), dt=dt, do_precond=False, model_name="RNN", device=dev)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # the golden is torch CPU fp32; cuDNN's GRU would otherwise run its GEMMs in TF32 (1e-4)
    try:
        loss, out = td.closure(model, None, t(g["init_x"]).unsqueeze(1), t(g["init_y"]).unsqueeze(1), t(g["F"]), g["S"], g["A"], None, None, dt, T)
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _rel(out.detach().cpu().numpy(), g["u_pred"]) < 1e-5
    assert _check_param_grads(model, g) < GRAD_RTOL


# ------------------------------------------------------------------------------------------------
# oracle on seeded inputs at the reference's config sizes (cfg3 N=2178, cfg4 N=1003, cfg1 N=387)
# ------------------------------------------------------------------------------------------------
def _ns_case(feo, n, B, branch, ordering, native, seed=0):
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", n, ordering=ordering)
    rng = np.random.default_rng(seed)
    alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=bool(branch), precond=None, device=dev)
    a = torch.tensor(alpha, device=dev)
    if native:
        a = feo.to_dof_major_tensor(a)
    a.requires_grad_(True)
    loss = ns.residual_loss(a, torch.tensor(F, device=dev), fx.A, fx.B1, fx.B2, fx.idx_sol)
    (grad,) = torch.autograd.grad(loss, a)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, bool(branch), dtype=np.float64)
    return loss.item(), grad, lo, go, ns


@pytest.mark.parametrize("n,B,branch,ordering,native", [
    (15, 64, 1, "blocked", False),      # cfg3 operator, precond (identity) branch
    (15, 64, 0, "interleaved", True),   # non-precond branch, dof-major native input
    (15, 37, 1, "interleaved", False),  # ragged batch (B % 4 != 0, B % 128 != 0)
    (6, 1, 0, "blocked", False),        # single sample
    (6, 130, 1, "blocked", True),       # crosses one 128-sample block
    (24, 256, 0, "interleaved", True),
])
def test_ns_vs_oracle(feo, n, B, branch, ordering, native):
    loss, grad, lo, go, _ = _ns_case(feo, n, B, branch, ordering, native)
    assert abs(loss - lo) <= LOSS_RTOL * abs(lo)
    assert grad.shape == go.shape
    assert _rel(grad.cpu().numpy(), go) < GRAD_RTOL and _relmax(grad.cpu().numpy(), go) < GRAD_RTOL
    if native:
        assert feo.is_dof_major(grad)


def test_ns_is_deterministic_and_tile_size_independent(feo, monkeypatch):
    monkeypatch.setenv("FEO_PLAN", "tile")  # the tile plan (any CSR); the lattice plan has its own file, tests/test_gpu_lattice.py
    l1, g1, _, _, _ = _ns_case(feo, 10, 96, 1, "interleaved", True)
    l2, g2, _, _, _ = _ns_case(feo, 10, 96, 1, "interleaved", True)
    assert l1 == l2 and torch.equal(g1, g2)  # bit-reproducible: no atomics anywhere
    for lines, warps in (("96", "3"), ("160", "7")):  # many small tiles, other warp counts
        monkeypatch.setenv("FEO_TILE_LINES_FWD", lines)
        monkeypatch.setenv("FEO_TILE_LINES_BWD", lines)
        monkeypatch.setenv("FEO_TILE_WARPS_FWD", warps)
        monkeypatch.setenv("FEO_TILE_WARPS_BWD", warps)
        l3, g3, lo, go, ns = _ns_case(feo, 10, 96, 1, "interleaved", True)
        assert ns.operator.info.n_tiles_fwd > 4
        assert torch.equal(g1, g3)  # per-row / per-column arithmetic does not depend on the tiling
        assert abs(l3 - lo) <= LOSS_RTOL * abs(lo)


def test_cfg5_full_size_properties(feo):
    """BASELINE.json configs[4] at full size (n=333, N=1 001 334), a 128-sample batch: the fused kernels
    against the SAME operator applied through the generic sparse kernels (feo_spmm) composed exactly as
    SURVEY.md Appendix A.2 writes the loss and its gradient -- a size-independent consistency property --
    plus padding-sample independence and bit-reproducibility."""
    from feonet_navier_stokes_b200 import _lib as L
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 333, ordering="interleaved")
    assert fx.N == 1001334
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=False, device=dev)
    op = ns.operator
    B, ldb = 100, 128  # ragged: 100 of 128 columns used
    gen = torch.Generator(device=dev).manual_seed(5)
    aT = torch.empty(fx.N, ldb, device=dev).normal_(0.0, 0.3, generator=gen)
    fT = torch.empty(fx.N, ldb, device=dev).normal_(0.0, 1.0, generator=gen)
    loss, rT = op.residual_fwd(aT, fT, B)
    gT = op.residual_bwd(aT, rT, B)
    # reference composition on the device: r = A a + F - c (non-precond branch), c = d1*B1a + d2*B2a
    I = torch.as_tensor(np.asarray(fx.idx_u1), device=dev)
    J = torch.as_tensor(np.asarray(fx.idx_u2), device=dev)
    Aa = op.spmm(L.FEO_MAT_A, False, aT, B)
    Bu1 = op.spmm(L.FEO_MAT_B1, False, aT, B)
    Bu2 = op.spmm(L.FEO_MAT_B2, False, aT, B)
    d1 = torch.zeros_like(aT)
    d2 = torch.zeros_like(aT)
    d1[I] = aT[I]; d1[J] = aT[I]
    d2[I] = aT[J]; d2[J] = aT[J]
    r_ref = Aa + fT - (d1 * Bu1 + d2 * Bu2)
    assert _rel(rT[:, :B].cpu().numpy(), r_ref[:, :B].cpu().numpy()) < 1e-6
    l_ref = float((r_ref[:, :B].double() ** 2).sum())
    assert abs(loss.item() - l_ref) <= LOSS_RTOL * l_ref
    G = 2.0 * r_ref
    g_ref = op.spmm(L.FEO_MAT_A, True, G, B)
    g_ref -= op.spmm(L.FEO_MAT_B1, True, d1 * G, B) + op.spmm(L.FEO_MAT_B2, True, d2 * G, B)
    w1, w2 = Bu1 * G, Bu2 * G
    e = torch.zeros_like(aT)
    e[I] = w1[I] + w1[J]
    e[J] = w2[I] + w2[J]
    g_ref -= e
    assert _rel(gT[:, :B].cpu().numpy(), g_ref[:, :B].cpu().numpy()) < GRAD_RTOL
    # padding columns (b >= B) never influence the result; the kernels are bit-reproducible
    aT2 = aT.clone()
    aT2[:, B:] = float("nan")
    loss2, rT2 = op.residual_fwd(aT2, fT, B)
    gT2 = op.residual_bwd(aT2, rT2, B)
    assert loss2.item() == loss.item() and torch.equal(gT2[:, :B], gT[:, :B])


@pytest.mark.parametrize("branch", [1, 0])
def test_cfg5_full_size_vs_oracle(feo, branch):
    """BASELINE.json configs[4] -- the benchmarked operator (n=333, N=1 001 334) -- against the fp64 oracle with scipy CSR
    (oracle.ns_loss_and_grad restates FEONet_steady_Navier-Stokes/train_FEONet.py:301-365) on 6 samples, through the public
    autograd API in BOTH layouts: the reference's row-major [B,1,N] network output and the dof-major native tensor.
    Tolerances as north_star: loss 1e-5, gradient 1e-4 (2-norm and max-norm)."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 333, ordering="interleaved")
    assert fx.N == 1001334
    dev = torch.device("cuda")
    rng = np.random.default_rng(50 + branch)
    B = 6
    alpha = (0.1 * rng.standard_normal((B, fx.N))).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, bool(branch), dtype=np.float64)
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=bool(branch), device=dev)
    Fd = torch.tensor(F, device=dev)
    for native in (False, True):
        a = torch.tensor(alpha, device=dev)
        if native:
            a = feo.to_dof_major_tensor(a)
        else:
            a = a.unsqueeze(1)  # [B,1,N], as the reference networks emit it
        a.requires_grad_(True)
        loss = ns.residual_loss(a, Fd, fx.A, fx.B1, fx.B2, fx.idx_sol)
        (grad,) = torch.autograd.grad(loss, a)
        g = grad.reshape(B, fx.N).cpu().numpy()
        assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo), (native, loss.item(), lo)
        assert _rel(g, go) < GRAD_RTOL and _relmax(g, go) < GRAD_RTOL, (native, _rel(g, go), _relmax(g, go))


def test_linear_properties_at_scale(feo):
    """Size-independent properties on a larger operator (n=64, N=37 442): linearity of the residual
    in (alpha, F) and gradient = 2 A^T r checked against the generic sparse apply."""
    from feonet_navier_stokes_b200 import _lib as L
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("stokes_square", 64)
    dev = torch.device("cuda")
    st = feo.LinearStokes(fx.A, None, do_precond=False, device=dev)
    op = st.operator
    B = 256
    gen = torch.Generator(device=dev).manual_seed(1)
    a1 = torch.randn(B, fx.N, device=dev, generator=gen)
    a2 = torch.randn(B, fx.N, device=dev, generator=gen)
    zero = torch.zeros(B, fx.N, device=dev)
    Y1, _ = st.weak_form(a1, zero, fx.A, None)
    Y2, _ = st.weak_form(a2, zero, fx.A, None)
    Y12, _ = st.weak_form(a1 + 2 * a2, zero, fx.A, None)
    assert _rel((Y1 + 2 * Y2).cpu().numpy(), Y12.cpu().numpy()) < 1e-6
    F = torch.randn(B, fx.N, device=dev, generator=gen)
    a = a1.clone().requires_grad_(True)
    loss = st.residual_loss(a, F, fx.A, None)
    (g,) = torch.autograd.grad(loss, a)
    r = Y1 - F
    assert abs(loss.item() - float((r.double() ** 2).sum())) <= LOSS_RTOL * loss.item()
    rT = op.to_dof_major(r)
    gref = op.from_dof_major(op.spmm(L.FEO_MAT_A, True, rT, B, scale=2.0), B)
    assert _rel(g.cpu().numpy(), gref.cpu().numpy()) < 1e-6


def test_grad_scaling_and_loss_only(feo):
    """Upstream gradient scaling (loss * k).backward and the no-grad (validation) path."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 8)
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, device=dev)
    a = (0.2 * torch.randn(50, fx.N, device=dev)).requires_grad_(True)
    F = torch.randn(50, fx.N, device=dev)
    l1 = ns.residual_loss(a, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
    (g1,) = torch.autograd.grad(l1, a)
    l2 = ns.residual_loss(a, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
    (g2,) = torch.autograd.grad(l2 * 0.25, a)
    assert torch.allclose(g2, 0.25 * g1, rtol=1e-6, atol=0)
    with torch.no_grad():
        l3 = ns.residual_loss(a, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
    assert l3.item() == l1.item()


def test_layout_roundtrip_bit_exact(feo):
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("stokes_square", 3)
    op = feo.FEOperator(fx.N, A=fx.A)
    for B in (1, 5, 64, 131):
        x = torch.randn(B, fx.N, device="cuda")
        xT = op.to_dof_major(x)
        assert xT.shape == (fx.N, (B + 3) // 4 * 4)
        assert torch.equal(xT[:, :B].t(), x)
        assert torch.equal(op.from_dof_major(xT, B, contiguous=True), x)
        assert op.to_dof_major(xT[:, :B].t()).data_ptr() == xT.data_ptr()  # native input: zero copy


def test_dense_precond_cfg1_vs_oracle(feo):
    """cfg1: N=387 operator with the shipped preconditioner blob (from the golden file), B=1000."""
    g = _load("stokes_precond72_n6")
    dev = torch.device("cuda")
    A, P = torch.tensor(g["A"], device=dev), torch.tensor(g["P"], device=dev)
    rng = np.random.default_rng(3)
    alpha = (0.3 * rng.standard_normal((1000, 387))).astype(np.float32)
    F = rng.standard_normal((1000, 387)).astype(np.float32)
    st = feo.LinearStokes(A, P, do_precond=True, device=dev)
    a = torch.tensor(alpha, device=dev, requires_grad=True)
    loss = st.residual_loss(a, torch.tensor(F, device=dev), A, P)
    (grad,) = torch.autograd.grad(loss, a)
    lo, go, _ = orc.stokes_loss_and_grad(alpha, F, g["A"], g["P"], True, dtype=np.float64)
    assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(grad.cpu().numpy(), go) < GRAD_RTOL and _relmax(grad.cpu().numpy(), go) < GRAD_RTOL


@pytest.mark.parametrize("n,B", [(72, 5), (200, 130), (387, 1000), (813, 257), (2549, 1024)])
def test_dense_apply_tensor_core_vs_fp64(feo, n, B):
    """feo_dense_apply (tcgen05 3xTF32, feo_dense_tc.cu) against the fp64 product: ragged n (not a multiple of the
    128 x 128 x 16 tile) and B, all epilogues (scale, device scale, subtract, sum of squares).  The split must be
    fp32-grade: every element within 2e-6 of the row's |D||x| bound (plain TF32 would sit near 5e-4)."""
    from feonet_navier_stokes_b200 import _lib as L

    dev = torch.device("cuda")
    rng = np.random.default_rng(n + B)
    D = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    D[rng.random((n, n)) < 0.1] = 0.0
    op = feo.FEOperator(n, dense_m=D, device=dev)
    x = rng.standard_normal((B, n)).astype(np.float32)
    f = rng.standard_normal((B, n)).astype(np.float32)
    xT, fT = op.to_dof_major(torch.tensor(x, device=dev)), op.to_dof_major(torch.tensor(f, device=dev))
    D64, x64, f64 = D.astype(np.float64), x.astype(np.float64), f.astype(np.float64)
    bound = (np.abs(D64) @ np.abs(x64).T)  # [n, B]
    # plain product, then the transposed operator
    for which, M in ((L.FEO_DENSE_M, D64), (L.FEO_DENSE_MT, D64.T)):
        cT = op.dense_apply(which, xT, B)
        err = np.abs(cT[:, :B].cpu().numpy().astype(np.float64) - M @ x64.T)
        bnd = bound if which == L.FEO_DENSE_M else (np.abs(D64.T) @ np.abs(x64).T)
        assert (err <= 2e-6 * bnd + 1e-30).all(), (which, float((err / (bnd + 1e-30)).max()))
    # fused residual epilogue: r = 0.5 * s * D x - f, loss = sum r^2
    sdev = torch.tensor([3.0], device=dev)
    rT, loss = op.dense_apply(L.FEO_DENSE_M, xT, B, scale=0.5, scale_dev=sdev, sub=fT, want_loss=True)
    r64 = 1.5 * (D64 @ x64.T) - f64.T
    assert np.abs(rT[:, :B].cpu().numpy() - r64).max() <= 4e-6 * max(1.0, bound.max())
    assert abs(loss.item() - (r64 ** 2).sum()) <= LOSS_RTOL * (r64 ** 2).sum()


@pytest.mark.parametrize("env", [{"FEO_DENSE_CLUSTER": "2"}, {"FEO_DENSE_BN": "128"}, {"FEO_DENSE_SIMT": "1"}, {"FEO_DENSE_GEN": "1"},
                                 {"FEO_DENSE_GEN": "2", "FEO_DENSE_BN": "160"}, {"FEO_DENSE_GEN": "3", "FEO_DENSE_BN": "128"},
                                 {"FEO_DENSE_GEN": "3", "FEO_DENSE_BN": "160", "FEO_DENSE_ASTAGES": "4", "FEO_DENSE_GAP": "2"}])
def test_dense_apply_optional_paths_subprocess(feo, env):
    """The dense apply reads its tuning knobs once per process: the optional paths (cluster multicast of the operator stages,
    128- / 160-column tiles, the three kernel generations -- activations split in registers, pre-split, CTA pairs with
    cta_group::2 incl. odd row-tile counts that get a padding CTA --, the fp32 FMA comparison kernel) are exercised in a child
    process, ragged sizes, against fp64."""
    import subprocess
    import sys

    code = r"""
import numpy as np, torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200 import _lib as L
feo.load_library(build_if_missing=False)
dev = torch.device("cuda")
rng = np.random.default_rng(7)
for n, B in ((813, 257), (200, 130), (72, 5)):
    D = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    op = feo.FEOperator(n, dense_m=D, device=dev)
    x = rng.standard_normal((B, n)).astype(np.float32); f = rng.standard_normal((B, n)).astype(np.float32)
    xT, fT = op.to_dof_major(torch.tensor(x, device=dev)), op.to_dof_major(torch.tensor(f, device=dev))
    rT, loss = op.dense_apply(L.FEO_DENSE_M, xT, B, sub=fT, want_loss=True)
    r64 = D.astype(np.float64) @ x.astype(np.float64).T - f.astype(np.float64).T
    bound = np.abs(D.astype(np.float64)) @ np.abs(x.astype(np.float64)).T + np.abs(f.astype(np.float64).T)
    err = np.abs(rT[:, :B].cpu().numpy() - r64)
    assert (err <= 3e-6 * bound + 1e-30).all(), (n, B, float((err / bound).max()))
    assert abs(loss.item() - (r64 ** 2).sum()) <= 1e-5 * (r64 ** 2).sum()
print("OK")
"""
    child_env = dict(os.environ, **env)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=child_env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_time_dep_cfg4_vs_oracle(feo):
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("time_dep", 10)  # N = 1003
    dev = torch.device("cuda")
    B, T, dt = 33, 10, 0.1
    rng = np.random.default_rng(4)
    pred = (0.3 * rng.standard_normal((B, T, fx.N))).astype(np.float32)
    u0 = rng.standard_normal((B, fx.N)).astype(np.float32)
    F = np.repeat(rng.standard_normal((1, fx.N)).astype(np.float32), B, axis=0)
    td = feo.TimeDependentStokes(fx.S, fx.A, fx.idx_sol, dt=dt, do_precond=False, device=dev)
    p = torch.tensor(pred, device=dev, requires_grad=True)
    loss = td.residual_loss(p, torch.tensor(F, device=dev), fx.S, fx.A, None, dt, torch.tensor(u0, device=dev))
    (grad,) = torch.autograd.grad(loss, p)
    lo, go, _ = orc.seq_loss_and_grad(pred, F, fx.S, fx.A, None, dt, u0, False, dtype=np.float64)
    assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(grad.cpu().numpy(), go) < GRAD_RTOL


@pytest.mark.parametrize("B,T", [(1, 1), (5, 1), (7, 2), (3, 3), (37, 7), (130, 1), (65, 4), (26, 10), (1, 300)])
def test_time_dep_sequence_shapes_vs_oracle(feo, B, T):
    """The vectorised sequence kernel (four consecutive pseudo-samples per lane, neighbouring time level by shuffle): sequence
    lengths below the vector width, pseudo-sample counts that are not multiples of 4 or 128, time levels that straddle lane
    and warp-tile boundaries (FEONet_time_dep_Stokes/train_FEONet.py:343-362, :398-400)."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("time_dep", 4)
    dev = torch.device("cuda")
    dt = 0.05
    rng = np.random.default_rng(100 * B + T)
    pred = (0.3 * rng.standard_normal((B, T, fx.N))).astype(np.float32)
    u0 = rng.standard_normal((B, fx.N)).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    td = feo.TimeDependentStokes(fx.S, fx.A, fx.idx_sol, dt=dt, do_precond=False, device=dev)
    p = torch.tensor(pred, device=dev, requires_grad=True)
    loss = td.residual_loss(p, torch.tensor(F, device=dev), fx.S, fx.A, None, dt, torch.tensor(u0, device=dev))
    (grad,) = torch.autograd.grad(loss, p)
    lo, go, _ = orc.seq_loss_and_grad(pred, F, fx.S, fx.A, None, dt, u0, False, dtype=np.float64)
    assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(grad.cpu().numpy(), go) < GRAD_RTOL
    assert np.abs(grad.cpu().numpy() - go).max() <= 1e-4 * np.abs(go).max()


def test_errors_are_loud(feo):
    from feonet_navier_stokes_b200 import _lib as L
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("stokes_square", 2)
    op = feo.FEOperator(fx.N, A=fx.A)
    x = torch.zeros(fx.N, 8, device="cuda")
    with pytest.raises(L.FeoError):
        op.spmm(L.FEO_MAT_B1, False, x, 8)  # matrix not present
    with pytest.raises(L.FeoError):
        op.dense_apply(L.FEO_DENSE_M, x, 8)  # no dense operator


def test_host_batch_pipeline_matches_one_shot(feo):
    """Chunked host->device streaming (feo.HostBatchPipeline) gives the one-shot loss (sum over samples) and the
    same gradients, bit for bit per sample (per-row / per-column arithmetic does not depend on the batch split)."""
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 12, ordering="interleaved")
    dev = torch.device("cuda")
    B, N = 200, fx.N
    gen = torch.Generator().manual_seed(3)
    a_host = (0.3 * torch.randn(B, N, generator=gen)).pin_memory()
    f_host = torch.randn(B, N, generator=gen).pin_memory()
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, precond=None, device=dev)
    a = a_host.to(dev).requires_grad_(True)
    loss = ns.residual_loss(a, f_host.to(dev), fx.A, fx.B1, fx.B2, fx.idx_sol)
    (g_ref,) = torch.autograd.grad(loss, a)
    grad = torch.empty(B, N, device=dev)
    pipe = feo.HostBatchPipeline(lambda x, f: ns.residual_loss(x, f, fx.A, fx.B1, fx.B2, fx.idx_sol), N, dev, chunk=64)
    for _ in range(2):  # second pass reuses the chunk buffers
        l2 = pipe.step(a_host, f_host, grad_out=grad)
        assert abs(l2 - loss.item()) <= LOSS_RTOL * abs(loss.item())
        assert torch.equal(grad, g_ref)


@pytest.mark.parametrize("branch", [True, False])
def test_ns_noncolocated_pairs_and_cross_terms_on_device(feo, branch):
    """I, J are opaque lists: shuffled pairing plus B1/B2 entries that couple the two
    components and the pressure columns, so that the generic step kinds of the fused kernels (forward X-steps and
    row quads, backward V-, A- and X-steps) run on the device, not only in the host replay."""
    import scipy.sparse as sp
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", 7, ordering="interleaved")
    rng = np.random.default_rng(11)
    idx_i = np.asarray(fx.idx_u1).copy()
    idx_j = np.asarray(fx.idx_u2).copy()[rng.permutation(len(fx.idx_u2))]
    noise = sp.random(fx.N, fx.N, density=0.01, random_state=5, data_rvs=rng.standard_normal).tocsr()
    B1, B2 = (fx.B1 + noise).tocsr(), (fx.B2 + noise.T).tocsr()
    idx_sol = np.empty(3, dtype=object)
    idx_sol[0], idx_sol[1], idx_sol[2] = idx_i.tolist(), idx_j.tolist(), list(fx.idx_p)
    B = 70
    alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, B1, B2, idx_sol, do_precond=branch, precond=None, device=dev)
    a = torch.tensor(alpha, device=dev).requires_grad_(True)
    loss = ns.residual_loss(a, torch.tensor(F, device=dev), fx.A, B1, B2, idx_sol)
    (grad,) = torch.autograd.grad(loss, a)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, B1, B2, idx_i, idx_j, branch, dtype=np.float64)
    assert abs(loss.item() - lo) <= LOSS_RTOL * abs(lo)
    assert _rel(grad.cpu().numpy(), go) < GRAD_RTOL and _relmax(grad.cpu().numpy(), go) < GRAD_RTOL


@pytest.mark.parametrize("resol", [20, 64, 7])
def test_sincos_forcing_grid_kernel(feo, resol):
    """Fused input synthesis (feo_sincos_forcing_grid) vs the reference's eager formula in `closure`
    (FEONet_steady_Navier-Stokes/train_FEONet.py:337-345) evaluated by torch on the CPU; fp32 sin/cos of the
    two libraries differ by a few ulp."""
    gen = torch.Generator().manual_seed(resol)
    coeff = torch.rand(33, 6, generator=gen)
    coeff[:, 2:] *= np.pi
    ref = feo.sincos_forcing_grid(coeff, resol)  # CPU tensor -> the reference's torch formula
    out = feo.sincos_forcing_grid(coeff.cuda(), resol)
    assert out.shape == ref.shape == (33, 2, resol, resol)
    assert torch.allclose(out.cpu(), ref, rtol=0, atol=2e-6)
    assert np.allclose(ref.numpy(), orc.sincos_forcing_grid(coeff.numpy(), resol), atol=2e-6)


@pytest.mark.parametrize("name", golden_cases("spai_"))
def test_spai_on_device_matches_the_reference(feo, name):
    """feo.spai_device against the reference's own `spai` output (FEONet_Stokes_square/train_FEONet.py:104-121, onenormest start;
    golden made by oracle/make_golden.py from the unmodified function)."""
    g = _load(name)
    P = feo.spai_device(g["A"], int(g["m"])).cpu().numpy()
    assert _rel(P, g["M"]) < 1e-10
    assert abs(np.linalg.norm(np.eye(P.shape[0]) - g["A"] @ P) - float(g["residual"])) < 1e-9


def test_spai_on_device_matches_the_host_iteration(feo):
    """feo.spai_device runs the reference's SPAI recurrence (FEONet_Stokes_square/train_FEONet.py:104-121) on the GPU;
    same steps as the host restatement, and the residual ||I - A M||_F decreases (minimal-residual iteration)."""
    from feonet_navier_stokes_b200.fixtures import config_operators, spai

    fx = config_operators("stokes_square", 6)
    A = np.asarray(fx.A.todense())
    P_host = spai(A, 60)
    P_dev = feo.spai_device(A, 60).cpu().numpy()
    assert np.allclose(P_dev, P_host, rtol=1e-9, atol=1e-12)
    eye = np.eye(A.shape[0])
    assert np.linalg.norm(eye - A @ P_dev) < np.linalg.norm(eye - A @ feo.spai_device(A, 0).cpu().numpy())
