"""Developer timing of the fused residual kernels through FEOperator (no autograd): CUDA events per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200.fixtures import config_operators

n = int(sys.argv[1]) if len(sys.argv) > 1 else 333
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
fx = config_operators("steady_ns", n, ordering="interleaved")
ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, precond=None, model_name="FCNN", device=dev)
op = ns.operator
N = fx.N
ldb = (B + 63) // 64 * 64
aT = torch.empty(N, ldb, device=dev).normal_(0, 0.1)
fT = torch.empty(N, ldb, device=dev).normal_(0, 1.0)
gT = torch.empty(N, ldb, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(3):
    loss, rT = op.residual_fwd(aT, fT, B)
    op.residual_bwd(aT, rT, B, out=gT)
torch.cuda.synchronize()
tf, tb = [], []
for k in range(K):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    loss, rT = op.residual_fwd(aT, fT, B)
    e1.record()
    op.residual_bwd(aT, rT, B, out=gT)
    e2.record()
    torch.cuda.synchronize()
    tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
tf.sort(); tb.sort()
print(f"n={n} N={N} B={B}: fwd median {tf[len(tf)//2]:.3f} ms (min {tf[0]:.3f}), bwd median {tb[len(tb)//2]:.3f} ms (min {tb[0]:.3f}); "
      f"samples/s {B / ((tf[len(tf)//2] + tb[len(tb)//2]) * 1e-3):.0f}; tiles {op.info.n_tiles_fwd}/{op.info.n_tiles_bwd}")
# back-to-back, no sync in between
e0, e1 = ev(), ev()
e0.record()
for k in range(K):
    loss, rT = op.residual_fwd(aT, fT, B)
    op.residual_bwd(aT, rT, B, out=gT)
e1.record()
torch.cuda.synchronize()
print(f"back-to-back: {e0.elapsed_time(e1) / K:.3f} ms per fwd+bwd")
