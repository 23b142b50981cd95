"""Developer timing of the time-dependent path at cfg4 (N = 1003, T = 10, B = 1000): the sequence kernels alone on dof-major
tensors (CUDA events) and the whole residual_loss + backward through the reference-facing API (eager and as one graph replay).

usage: time_seq.py [B] [T] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200.fixtures import config_operators

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
fx = config_operators("time_dep", 10)
dt = 0.1
td = feo.TimeDependentStokes(fx.S, fx.A, fx.idx_sol, dt=dt, do_precond=False, device=dev)
op = td.operator
J = B * T
ldj, ldb = (J + 3) // 4 * 4, (B + 3) // 4 * 4
pT = torch.randn(fx.N, ldj, device=dev)
u0T = torch.randn(fx.N, ldb, device=dev)
fT = torch.randn(fx.N, ldb, device=dev)


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


_, rT = op.seq_fwd(pT, u0T, fT, B, T)
g1 = torch.ones((), device=dev)
tf = timed(lambda: op.seq_fwd(pT, u0T, fT, B, T))
tb = timed(lambda: op.seq_bwd(rT, B, T, grad_loss=g1))
fwd_bytes = 4.0 * fx.N * (2 * J + 2 * B)
bwd_bytes = 4.0 * fx.N * 2 * J
print(f"seq kernels N={fx.N} B={B} T={T}: fwd {tf*1e3:.1f} us ({fwd_bytes/tf/1e6:.0f} GB/s algorithmic), bwd {tb*1e3:.1f} us ({bwd_bytes/tb/1e6:.0f} GB/s)")
pred = torch.randn(B, T, fx.N, device=dev, requires_grad=True)
u0 = torch.randn(B, fx.N, device=dev)
F = torch.randn(B, fx.N, device=dev)


def step():
    loss = td.residual_loss(pred, F, fx.S, fx.A, None, dt, u0)
    torch.autograd.grad(loss, pred)


te = timed(step)
gl = feo.GraphedLossGrad(lambda p_: td.residual_loss(p_, F, fx.S, fx.A, None, dt, u0), [pred])
tg = timed(lambda: gl())
print(f"API step (row-major [B,T,N] in, gradient out): eager {te*1e3:.1f} us, graph replay {tg*1e3:.1f} us")
