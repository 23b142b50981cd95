"""TEST INFRASTRUCTURE -- generate `tests/golden/*.npz` from the reference's own code.

Run in the authoring container only (needs `/root/reference`):

    python oracle/make_golden.py

For every variant it executes the reference's unmodified `closure` / `weak_form` /
`weak_form_sequence` / `assemble_u_init` (AST-extracted, see `oracle/reference_extract.py`) on
seeded inputs in torch CPU fp32, with alpha fed through a leaf "model" so that
`loss.backward()` (the reference's autograd) yields d loss / d alpha.  Inputs and outputs are
stored as small compressed npz fixtures; the GPU box (no `/root/reference`) replays them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from feonet_navier_stokes_b200.fixtures import assemble_operators, config_operators, spai, structured_mesh  # noqa: E402
from oracle.reference_extract import REFERENCE_ROOT, load_reference_functions, make_idx_sol  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class Leaf(torch.nn.Module):
    """Stand-in 'network' whose output is a leaf tensor, so autograd gives d loss / d alpha."""

    def __init__(self, alpha: torch.Tensor):
        super().__init__()
        self.alpha = torch.nn.Parameter(alpha.clone())

    def forward(self, *_a, **_k):
        return self.alpha


def _dense32(K):
    return np.asarray(K.toarray(), dtype=np.float64)


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def golden_steady_ns(tag, n, B, do_precond, precond_kind, seed):
    rng = np.random.default_rng(seed)
    op = config_operators("steady_ns", n, ordering="interleaved" if seed % 2 else "blocked")
    N = op.N
    A, B1, B2 = _dense32(op.A), _dense32(op.B1), _dense32(op.B2)
    if precond_kind == "identity":  # what the shipped script does (train_FEONet.py:142)
        P = np.eye(N)
    elif precond_kind == "dense":
        P = np.eye(N) + 0.05 * rng.standard_normal((N, N)) / np.sqrt(N)
    else:
        P = np.zeros_like(A)  # train_FEONet.py:168
    alpha = (0.3 * rng.standard_normal((B, N))).astype(np.float32)
    F = rng.standard_normal((B, N)).astype(np.float32)
    idx_sol = make_idx_sol(op.idx_u1, op.idx_u2, op.idx_p)
    t = lambda a: torch.tensor(a).float()  # noqa: E731  (same cast chain as train_FEONet.py:293-296)
    ns = load_reference_functions(
        "steady_ns",
        dict(DO_PRECOND=bool(do_precond), PRECOND=t(P), IDX_SOL=idx_sol, NUM_PTS=N, FORCE="sincos",
             gparams={"model": "FCNN"}),
    )
    model = Leaf(torch.tensor(alpha))
    coeff_f = torch.zeros(B, 6)
    loss, u_pred = ns["closure"](model, coeff_f, None, t(F), t(A), t(B1), t(B2), 8)
    loss.backward()
    LHS, RHS = ns["weak_form"](model.alpha.detach().unsqueeze(1), t(F), t(A), t(B1), t(B2), idx_sol)
    _save(
        tag, variant="steady_ns", do_precond=bool(do_precond), precond_kind=precond_kind,
        A=A.astype(np.float32), B1=B1.astype(np.float32), B2=B2.astype(np.float32),
        P=P.astype(np.float32), idx_u1=op.idx_u1, idx_u2=op.idx_u2, idx_p=op.idx_p,
        alpha=alpha, F=F, LHS=LHS.numpy(), RHS=RHS.numpy(), loss=np.float32(loss.item()),
        grad=model.alpha.grad.numpy(), u_pred=u_pred.detach().numpy(),
    )


def golden_stokes(tag, variant, op, P, B, do_precond, seed):
    rng = np.random.default_rng(seed)
    N = op.N
    A = _dense32(op.A)
    alpha = (0.3 * rng.standard_normal((B, N))).astype(np.float32)
    F = rng.standard_normal((B, N)).astype(np.float32)
    t = lambda a: torch.tensor(a).float()  # noqa: E731
    ns = load_reference_functions(variant, dict(DO_PRECOND=bool(do_precond), NUM_PTS=N, gparams={"model": "FCNN"}))
    model = Leaf(torch.tensor(alpha))
    coeff_f = torch.zeros(B, 6)
    if variant == "hole":  # closure(model, coeff_f, value_f, load_vec_f, matrix, precond, resol_in)
        ns["FORCE"] = "sincos"
        loss, u_pred = ns["closure"](model, coeff_f, None, t(F), t(A), t(P), 8)
    else:
        loss, u_pred = ns["closure"](model, coeff_f, t(F), t(A), t(P), 8)
    loss.backward()
    LHS, RHS = ns["weak_form"](model.alpha.detach().unsqueeze(1), t(F), t(A), t(P))
    _save(
        tag, variant=variant, do_precond=bool(do_precond), A=A.astype(np.float32), P=np.asarray(P, np.float32),
        idx_u1=op.idx_u1, idx_u2=op.idx_u2, idx_p=op.idx_p, alpha=alpha, F=F,
        LHS=LHS.numpy(), RHS=RHS.numpy(), loss=np.float32(loss.item()),
        grad=model.alpha.grad.numpy(), u_pred=u_pred.detach().numpy(),
    )


def golden_time_dep(tag, n, B, T, dt, do_precond, seed):
    rng = np.random.default_rng(seed)
    op = config_operators("time_dep", n)
    N = op.N
    A, S = _dense32(op.A), _dense32(op.S)
    P = spai(S + dt * A, 25) if do_precond else np.zeros_like(A)
    pred = (0.3 * rng.standard_normal((B, T, N))).astype(np.float32)
    init = rng.standard_normal((B, 2, op.mesh.n_u)).astype(np.float32)
    F1 = rng.standard_normal((N,)).astype(np.float32)
    F = np.repeat(F1[None, :], B, axis=0)  # same vector replicated (train_FEONet.py:235,244)
    idx_sol = make_idx_sol(op.idx_u1, op.idx_u2, op.idx_p)
    t = lambda a: torch.tensor(a).float()  # noqa: E731
    ns = load_reference_functions(
        "time_dep",
        dict(DO_PRECOND=bool(do_precond), IDX_SOL=idx_sol, NUM_PTS=N, DT=dt, BC="lower",
             gparams={"model": "RNN"}, P=t(np.zeros((2, 2)))),
    )

    class SeqLeaf(Leaf):
        def forward(self, u_init, seq_len=None):
            return self.alpha

    model = SeqLeaf(torch.tensor(pred))
    init_x, init_y = t(init[:, 0:1, :]), t(init[:, 1:2, :])
    loss, out = ns["closure"](model, None, init_x, init_y, t(F), t(S), t(A), None, t(P), dt, T)
    loss.backward()
    u0 = ns["assemble_u_init"](init_x, init_y, idx_sol, N, torch.device("cpu"))
    LHS, RHS = ns["weak_form_sequence"](model.alpha.detach(), t(F), t(S), t(A), t(P), dt, u0, bool(do_precond))
    _save(
        tag, variant="time_dep", do_precond=bool(do_precond), A=A.astype(np.float32), S=S.astype(np.float32),
        P=np.asarray(P, np.float32), dt=np.float64(dt), idx_u1=op.idx_u1, idx_u2=op.idx_u2, idx_p=op.idx_p,
        pred=pred, init_x=init[:, 0], init_y=init[:, 1], F=F, u_init=u0.numpy(),
        LHS=LHS.numpy(), RHS=RHS.numpy(), loss=np.float32(loss.item()), grad=model.alpha.grad.numpy(),
        u_pred=out.detach().numpy(),
    )


def _reference_network(variant_dir):
    """The reference's own network.py, imported from /root/reference (plain module: no side effects at import)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(f"refnet_{abs(hash(variant_dir))}", os.path.join(REFERENCE_ROOT, variant_dir, "network.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _state_arrays(model, prefix):
    return {f"{prefix}{k.replace('.', '__')}": v.detach().numpy() for k, v in model.state_dict().items()}


def _grad_arrays(model, prefix):
    return {f"{prefix}{k.replace('.', '__')}": p.grad.detach().numpy() for k, p in model.named_parameters()}


def golden_trainstep_ns(tag, n, B, do_precond, seed):
    """One training-step evaluation of the reference: its closure (FEONet_steady_Navier-Stokes/train_FEONet.py:334-365) on ITS
    FCNN (network.py:120-138, eval mode: dropout off so that the step is reproducible) and `loss.backward()` (:463): loss and the
    gradient of every network parameter."""
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    op = config_operators("steady_ns", n, ordering="interleaved")
    N = op.N
    A, B1, B2 = _dense32(op.A), _dense32(op.B1), _dense32(op.B2)
    coeff_f = rng.uniform(0.0, 1.0, size=(B, 6)).astype(np.float32)
    F = rng.standard_normal((B, N)).astype(np.float32)
    idx_sol = make_idx_sol(op.idx_u1, op.idx_u2, op.idx_p)
    t = lambda a: torch.tensor(a).float()  # noqa: E731
    ns = load_reference_functions("steady_ns", dict(DO_PRECOND=bool(do_precond), PRECOND=torch.eye(N), IDX_SOL=idx_sol, NUM_PTS=N,
                                                    FORCE="sincos", gparams={"model": "FCNN"}))
    model = _reference_network("FEONet_steady_Navier-Stokes").FCNN(6, N, [16, 32, 64]).eval()
    state = _state_arrays(model, "state__")
    loss, u_pred = ns["closure"](model, t(coeff_f), None, t(F), t(A), t(B1), t(B2), 8)
    loss.backward()
    _save(tag, variant="steady_ns", do_precond=bool(do_precond), A=A.astype(np.float32), B1=B1.astype(np.float32), B2=B2.astype(np.float32),
          idx_u1=op.idx_u1, idx_u2=op.idx_u2, idx_p=op.idx_p, coeff_f=coeff_f, F=F, loss=np.float32(loss.item()),
          u_pred=u_pred.detach().numpy(), hidden=np.array([16, 32, 64]), **state, **_grad_arrays(model, "grad__"))


def golden_trainstep_time_dep(tag, n, B, T, dt, seed):
    """The same for the time-dependent variant: closure (FEONet_time_dep_Stokes/train_FEONet.py:364-406) on the reference's
    VectorToSequenceRNN (network.py:342-399), loss.backward() (:507)."""
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    op = config_operators("time_dep", n, ordering="interleaved")
    N = op.N
    A, S = _dense32(op.A), _dense32(op.S)
    init = (0.5 * rng.standard_normal((B, 2, op.mesh.n_u))).astype(np.float32)
    F = np.repeat(rng.standard_normal((1, N)).astype(np.float32), B, axis=0)
    idx_sol = make_idx_sol(op.idx_u1, op.idx_u2, op.idx_p)
    t = lambda a: torch.tensor(a).float()  # noqa: E731
    ns = load_reference_functions("time_dep", dict(DO_PRECOND=False, IDX_SOL=idx_sol, NUM_PTS=N, DT=dt, BC="lower",
                                                   gparams={"model": "RNN"}, P=t(np.zeros((2, 2)))))
    model = _reference_network("FEONet_time_dep_Stokes").VectorToSequenceRNN(input_dim=N, hidden_dim=24, output_dim=N, rnn_type="gru", num_layers=1)
    state = _state_arrays(model, "state__")
    loss, out = ns["closure"](model, None, t(init[:, 0:1, :]), t(init[:, 1:2, :]), t(F), t(S), t(A), None, t(np.zeros_like(A)), dt, T)
    loss.backward()
    _save(tag, variant="time_dep", do_precond=False, A=A.astype(np.float32), S=S.astype(np.float32), dt=np.float64(dt), T=np.int64(T),
          idx_u1=op.idx_u1, idx_u2=op.idx_u2, idx_p=op.idx_p, init_x=init[:, 0], init_y=init[:, 1], F=F, loss=np.float32(loss.item()),
          u_pred=out.detach().numpy(), hidden=np.array([24]), **state, **_grad_arrays(model, "grad__"))


def golden_spai(tag, variant, n, m):
    """The reference's own `spai(A, m)` (FEONet_Stokes_square/train_FEONet.py:104-121: onenormest start, dense numpy / scipy)."""
    op = config_operators(variant, n)
    A = _dense32(op.A)
    ns = load_reference_functions("stokes_square", dict(tqdm=lambda it: it))
    M = np.asarray(ns["spai"](A, m))
    _save(tag, A=A, m=np.int64(m), M=M, residual=np.float64(np.linalg.norm(np.eye(A.shape[0]) - A @ M)))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    # A.2 steady Navier-Stokes: both sign branches + a genuinely dense preconditioner
    golden_steady_ns("ns_precond_identity_n3", 3, 6, True, "identity", 10)
    golden_steady_ns("ns_noprecond_n3", 3, 6, False, "zeros", 11)
    golden_steady_ns("ns_precond_dense_n2", 2, 5, True, "dense", 12)
    golden_steady_ns("ns_noprecond_n4", 4, 7, False, "zeros", 13)
    # A.1 linear Stokes: cfg1 operator with the SHIPPED preconditioner blob
    op72 = config_operators("stokes_square", 6)
    P72 = np.load(os.path.join(REFERENCE_ROOT, "FEONet_Stokes_square", "precond_72_channel_flow.npy"))
    golden_stokes("stokes_precond72_n6", "stokes_square", op72, P72, 5, True, 20)
    op3 = config_operators("stokes_square", 3)
    golden_stokes("stokes_noprecond_n3", "stokes_square", op3, np.zeros((op3.N, op3.N)), 6, False, 21)
    oph = config_operators("hole", 5)
    golden_stokes("hole_precond_spai", "hole", oph, spai(_dense32(oph.A), 40), 5, True, 22)
    golden_stokes("hole_noprecond", "hole", oph, np.zeros((oph.N, oph.N)), 4, False, 23)
    # A.3 time-dependent Stokes
    golden_time_dep("timedep_noprecond_n3", 3, 5, 4, 0.1, False, 30)
    golden_time_dep("timedep_precond_n3", 3, 4, 3, 0.01, True, 31)
    golden_spai("spai_stokes_n3_m40", "stokes_square", 3, 40)
    golden_spai("spai_hole_n4_m25", "hole", 4, 25)
    # one training-step evaluation with the reference's own networks (loss + parameter gradients)
    golden_trainstep_ns("trainstep_ns_precond_n4", 4, 6, True, 40)
    golden_trainstep_ns("trainstep_ns_noprecond_n4", 4, 5, False, 41)
    golden_trainstep_time_dep("trainstep_timedep_n4", 4, 5, 4, 0.1, 42)


if __name__ == "__main__":
    main()
