"""CUDA-graph capture of the launch-bound steps (feonet_navier_stokes_b200/graphs.py): a replay must give what the eager
path gives -- for new contents of the static buffers too -- and the graphed optimiser step must follow the eager trainer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ns_problem(torch, feo, n=4, B=37, seed=0):
    from feonet_navier_stokes_b200.fixtures import config_operators

    fx = config_operators("steady_ns", n)
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=False, device=dev)
    alpha = torch.tensor((0.3 * rng.standard_normal((B, fx.N))).astype(np.float32), device=dev)
    F = torch.tensor(rng.standard_normal((B, fx.N)).astype(np.float32), device=dev)
    return fx, ns, alpha, F


def test_graphed_loss_and_gradient_follow_the_static_buffers():
    import torch

    import feonet_navier_stokes_b200 as feo

    fx, ns, alpha, F = _ns_problem(torch, feo)
    loss_fn = lambda a, f: ns.residual_loss(a, f, fx.A, fx.B1, fx.B2, fx.idx_sol)  # noqa: E731

    def eager(a, f):
        a = a.detach().clone().requires_grad_(True)
        loss = loss_fn(a, f)
        (g,) = torch.autograd.grad(loss, a)
        return loss.item(), g

    a_static, f_static = alpha.clone(), F.clone()
    gl = feo.GraphedLossGrad(loss_fn, [a_static, f_static], wrt=(0,))
    l0, (g0,) = gl()
    le, ge = eager(alpha, F)
    assert l0.item() == le and torch.equal(g0, ge)  # same kernels, fixed summation order: the same bits
    # new contents, both through the call (copied into the static buffers) and written in place by the caller:
    # the load vector's layout pass must be part of the graph (the eager path caches it by tensor identity)
    a2, f2 = alpha * 0.5 + 0.1, F * -2.0
    l1, (g1,) = gl(a2, f2)
    le, ge = eager(a2, f2)
    assert l1.item() == le and torch.equal(g1, ge)
    with torch.no_grad():
        f_static.mul_(0.25)
    l2, (g2,) = gl()
    le, ge = eager(a2, f2 * 0.25)
    assert l2.item() == le and torch.equal(g2, ge)
    with pytest.raises(ValueError):
        gl(a2[:5], f2[:5])


def test_graphed_dense_and_sequence_losses():
    """The preconditioned (tensor-core) loss and the time-dependent sequence loss replay to the eager values."""
    import torch

    import feonet_navier_stokes_b200 as feo
    from feonet_navier_stokes_b200.fixtures import config_operators

    dev = torch.device("cuda")
    rng = np.random.default_rng(1)
    fx = config_operators("stokes_square", 4)
    A = torch.tensor(np.asarray(fx.A.todense(), dtype=np.float32), device=dev)
    P = torch.tensor((np.eye(fx.N) + 0.2 * rng.standard_normal((fx.N, fx.N)) / np.sqrt(fx.N)).astype(np.float32), device=dev)
    st = feo.LinearStokes(A, P, do_precond=True, device=dev)
    a = torch.tensor((0.3 * rng.standard_normal((50, fx.N))).astype(np.float32), device=dev, requires_grad=True)
    F = torch.tensor(rng.standard_normal((50, fx.N)).astype(np.float32), device=dev)
    loss = st.residual_loss(a, F, A, P)  # an eager graph on the legacy stream that stays alive across the capture
    (g,) = torch.autograd.grad(loss, a, retain_graph=True)
    gl = feo.GraphedLossGrad(lambda a_: st.residual_loss(a_, F, A, P), [a])
    lg, (gg,) = gl()
    assert lg.item() == loss.item() and torch.equal(gg, g)

    fx = config_operators("time_dep", 4)
    T, dt, B = 5, 0.1, 9
    td = feo.TimeDependentStokes(fx.S, fx.A, fx.idx_sol, dt=dt, do_precond=False, device=dev)
    pred = torch.tensor((0.3 * rng.standard_normal((B, T, fx.N))).astype(np.float32), device=dev, requires_grad=True)
    u0 = torch.tensor(rng.standard_normal((B, fx.N)).astype(np.float32), device=dev)
    Ft = torch.tensor(rng.standard_normal((B, fx.N)).astype(np.float32), device=dev)
    gl = feo.GraphedLossGrad(lambda p_: td.residual_loss(p_, Ft, fx.S, fx.A, None, dt, u0), [pred])
    loss = td.residual_loss(pred, Ft, fx.S, fx.A, None, dt, u0)
    (g,) = torch.autograd.grad(loss, pred)
    lg, (gg,) = gl()
    assert lg.item() == loss.item() and torch.equal(gg, g)


def _trainer(tmp_path, graph, extra=()):
    import torch

    from feonet_navier_stokes_b200 import train_FEONet as T

    argv = ["--variant", "steady_ns", "--train_file", "24N32", "--val_file", "4N32", "--model", "FCNN", "--optimizer", "Adam", "--do_precond", "1",
            "--epochs", "12", "--log_every", "4", "--lr", "3e-3", "--out", str(tmp_path), "--seed", "3", "--cuda_graph", str(graph), *extra]
    tr = T.Trainer(dict(T.build_parser().parse_args(argv).__dict__), device=torch.device("cuda"))
    for m in tr.model.modules():  # dropout draws differ between an eager and a replayed step: compare without it
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return tr


@pytest.mark.parametrize("batch_size", [None, 8])
def test_graphed_train_step_follows_the_eager_trainer(tmp_path, batch_size):
    """Full-batch (replay on the resident tensors) and mini-batches (copied into static buffers): the loss trajectory and the
    parameters after 12 epochs agree with the eager loop; capturing itself (three warm-up steps) does not advance training."""
    import torch

    extra = () if batch_size is None else ("--batch_size_train", str(batch_size))
    import feonet_navier_stokes_b200 as feo

    eager, graphed = _trainer(tmp_path / "e", 0, extra), _trainer(tmp_path / "g", 1, extra)
    eager.optimizer = feo.make_capturable_optimizer("Adam", eager.model.parameters(), 3e-3)  # the same fused update in both loops
    for pe, pg in zip(eager.model.parameters(), graphed.model.parameters()):
        assert torch.equal(pe, pg)
    eager.fit()
    graphed.fit()
    assert graphed._graphs and all(s.replays > 0 for s in graphed._graphs.values())
    np.testing.assert_allclose(graphed.losses, eager.losses, rtol=1e-5)
    for pe, pg in zip(eager.model.parameters(), graphed.model.parameters()):
        assert torch.allclose(pe, pg, rtol=1e-4, atol=1e-6)
    assert eager.losses[-1] < eager.losses[0]


def test_graphed_step_skips_a_non_finite_batch_on_the_device(tmp_path):
    """The bad-value guard of the reference's loop (FEONet_steady_Navier-Stokes/train_FEONet.py:434-469) inside the graph: a
    batch with a NaN load vector leaves parameters and optimiser state untouched, without a host read."""
    import torch

    tr = _trainer(tmp_path, 1)
    batch = next(tr.batches(tr.train, None, shard=True))
    loss, ok = tr.train_step(batch)
    assert bool(ok.item()) and np.isfinite(loss.item())
    before = [p.detach().clone() for p in tr.model.parameters()]
    steps = [st["step"].clone() for st in tr.optimizer.state.values()]
    good = tr.train["load_vec_f"].clone()
    tr.train["load_vec_f"][0, 0] = float("nan")
    loss, ok = tr.train_step(batch)
    assert not bool(ok.item())
    for b, p in zip(before, tr.model.parameters()):
        assert torch.equal(b, p)
    for s0, st in zip(steps, tr.optimizer.state.values()):
        assert torch.equal(s0, st["step"])
    tr.train["load_vec_f"].copy_(good)
    loss, ok = tr.train_step(batch)
    assert bool(ok.item()) and not torch.equal(before[0], next(tr.model.parameters()))
