"""Developer timing of the steady Navier-Stokes loss + gradient at the reference's own size (cfg3: n = 15 mesh, N = 2178,
B = 1000) through the row-major API: eager, and as one CUDA-graph replay.  Run under `ncu --metrics gpu__time_duration.sum`
for the per-kernel list.  usage: time_cfg3.py [n] [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200.fixtures import config_operators

n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda:0")
fx = config_operators("steady_ns", n)
ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, device=dev, dof_positions=fx.pos if os.environ.get("USE_POS") else None)
a = torch.randn(B, fx.N, device=dev, requires_grad=True)
F = torch.randn(B, fx.N, device=dev)


def step():
    loss = ns.residual_loss(a, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
    torch.autograd.grad(loss, a)


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


te = timed(step)
gl = feo.GraphedLossGrad(lambda a_: ns.residual_loss(a_, F, fx.A, fx.B1, fx.B2, fx.idx_sol), [a])
tg = timed(lambda: gl())
print(f"cfg3 N={fx.N} B={B} plan={ns.operator.plan}: eager {te*1e3:.1f} us, graph replay {tg*1e3:.1f} us; ideal at 6.5 TB/s {24.0*fx.N*B/6.5e12*1e6:.1f} us")
