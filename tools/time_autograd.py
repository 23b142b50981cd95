"""Developer timing: where does the time go on the autograd path (events + CPU wall per piece)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200.fixtures import config_operators

n = int(sys.argv[1]) if len(sys.argv) > 1 else 333
B = 1024
dev = torch.device("cuda:0")
fx = config_operators("steady_ns", n, ordering="interleaved")
ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, precond=None, model_name="FCNN", device=dev)
op = ns.operator
N = fx.N
alpha = feo.dof_major_empty(B, N, dev).normal_(0, 0.1).requires_grad_(True)
F = feo.dof_major_empty(B, N, dev).normal_(0, 1.0)
ev = lambda: torch.cuda.Event(enable_timing=True)
def step(sync):
    alpha.grad = None
    e = [ev() for _ in range(3)]
    t = [time.perf_counter()]
    e[0].record()
    loss = ns.residual_loss(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
    if sync: torch.cuda.synchronize()
    t.append(time.perf_counter())
    e[1].record()
    loss.backward()
    if sync: torch.cuda.synchronize()
    t.append(time.perf_counter())
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), (t[1]-t[0])*1e3, (t[2]-t[1])*1e3
for _ in range(3): step(False)
for sync in (True, False):
    for k in range(4):
        print("sync" if sync else "async", "fwd_ev %.3f bwd_ev %.3f | cpu fwd %.3f bwd %.3f" % step(sync))
print("alpha strides", alpha.stride(), "grad strides", alpha.grad.stride(), alpha.grad.data_ptr() % 256)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(False)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
import cProfile, pstats
pr = cProfile.Profile()
alpha.grad = None
torch.cuda.synchronize()
pr.enable()
loss = ns.residual_loss(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
