"""Times the reference's UNMODIFIED hot-path functions (AST-extracted from /root/reference, as oracle/make_golden.py does)
beside the oracle's restatement of the same execution plan (oracle.TorchReferenceLoops), on this container's host cores, at
cfg1 (linear Stokes, N = 387, shipped-size preconditioner) and cfg3 (steady NS, N = 2178), B = 1000.
Run in the authoring container only (/root/reference does not exist on the GPU box):
    python tools/time_reference_here.py > profiles/r02_reference_cpu_container.json
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from feonet_navier_stokes_b200.fixtures import config_operators
from oracle import feonet_oracle as orc
from oracle.reference_extract import load_reference_functions, make_idx_sol

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
torch.set_num_threads(os.cpu_count())
rng = np.random.default_rng(0)
rows = []


def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


# cfg3: steady NS
fx = config_operators("steady_ns", 15)
A, B1, B2 = (torch.tensor(np.asarray(K.todense(), dtype=np.float32)) for K in (fx.A, fx.B1, fx.B2))
alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
F = rng.standard_normal((B, fx.N)).astype(np.float32)
idx_sol = make_idx_sol(fx.idx_u1, fx.idx_u2, fx.idx_p)
for precond in (True, False):
    ns = load_reference_functions("steady_ns", {"DO_PRECOND": precond, "PRECOND": torch.eye(fx.N), "IDX_SOL": idx_sol, "NUM_PTS": fx.N,
                                                "FORCE": "sincos", "gparams": {"model": "FCNN"}})

    def ref_step():
        a = torch.tensor(alpha).unsqueeze(1).requires_grad_(True)
        model = lambda c: a.squeeze(1)  # the network output is the leaf
        loss, _ = ns["closure"](model, None, None, torch.tensor(F), A, B1, B2, 64)
        loss.backward()
        return float(loss), a.grad.squeeze(1)

    loops = orc.TorchReferenceLoops(os.cpu_count())
    t_ref, (l_ref, g_ref) = timed(ref_step)
    t_port, (l_port, g_port) = timed(lambda: loops.steady_ns_step(alpha, F, A.numpy(), B1.numpy(), B2.numpy(), fx.idx_u1, fx.idx_u2, precond))
    rows.append({"config": f"cfg3 steady NS N={fx.N} B={B} do_precond={precond}", "reference_s_per_step": t_ref, "restatement_s_per_step": t_port,
                 "reference_samples_per_s": B / t_ref, "loss_equal": l_ref == l_port, "grad_equal": bool(torch.equal(g_ref, g_port))})
    print(rows[-1], file=sys.stderr, flush=True)

# cfg1: linear Stokes, preconditioned
fx = config_operators("stokes_square", 6)
M = torch.tensor(np.asarray(fx.A.todense(), dtype=np.float32))
P = torch.tensor((np.eye(fx.N) + 0.3 * rng.standard_normal((fx.N, fx.N)) / np.sqrt(fx.N)).astype(np.float32))
alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
F = rng.standard_normal((B, fx.N)).astype(np.float32)
ns = load_reference_functions("stokes_square", {"DO_PRECOND": True, "NUM_PTS": fx.N, "gparams": {"model": "FCNN"}})


def ref_step1():
    a = torch.tensor(alpha).unsqueeze(1).requires_grad_(True)
    loss, _ = ns["closure"](lambda c: a.squeeze(1), None, torch.tensor(F), M, P, 64)
    loss.backward()
    return float(loss), a.grad.squeeze(1)


loops = orc.TorchReferenceLoops(os.cpu_count())
t_ref, (l_ref, g_ref) = timed(ref_step1, reps=1)
t_port, (l_port, g_port) = timed(lambda: loops.linear_stokes_step(alpha, F, M.numpy(), P.numpy(), True), reps=1)
rows.append({"config": f"cfg1 linear Stokes N={fx.N} B={B} preconditioned", "reference_s_per_step": t_ref, "restatement_s_per_step": t_port,
             "reference_samples_per_s": B / t_ref, "loss_equal": l_ref == l_port, "grad_equal": bool(torch.equal(g_ref, g_port))})
print(rows[-1], file=sys.stderr, flush=True)
print(json.dumps({"host": "authoring container", "cores": os.cpu_count(), "torch": torch.__version__, "rows": rows}, indent=1))
