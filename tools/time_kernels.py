"""Developer timing of the fused residual kernels through FEOperator (no autograd): CUDA events per call.

usage: time_kernels.py [n] [B] [K] [cfg ...]   with cfg = env:K=V,... (environment knobs, sticky; K= unsets) or Wf,Lf,Wb,Lb[,gap,reserve[,Sf,Sb]] (consumer warps / staged lines forward / backward, gap filling, line stages)
The fixture is assembled once; one operator (one tile plan) is built per cfg.  A checksum of the loss
and the gradient is printed per cfg so that plans can be compared with each other.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200.fixtures import config_operators
from feonet_navier_stokes_b200.operator import FEOperator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 333
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfgs = sys.argv[4:] or ["default"]
dev = torch.device("cuda:0")
t0 = time.time()
fx = config_operators("steady_ns", n, ordering="interleaved")
print(f"fixture n={n} N={fx.N}: {time.time() - t0:.1f} s", flush=True)
N = fx.N
ldb = (B + 63) // 64 * 64
aT = torch.empty(N, ldb, device=dev).normal_(0, 0.1)
fT = torch.empty(N, ldb, device=dev).normal_(0, 1.0)
gT = torch.empty(N, ldb, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for cfg in cfgs:
    if cfg.startswith("env:"):  # env:K=V,K=V -- any FEO_* knob (plan choice, lattice widths, debug modes); "env:" alone resets nothing
        for kv in cfg[4:].split(","):
            if kv:
                k, v = kv.split("=")
                if v == "":
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    elif cfg != "default":
        wf, lf, wb, lb, *rest = cfg.split(",")
        os.environ.update(FEO_TILE_WARPS_FWD=wf, FEO_TILE_LINES_FWD=lf, FEO_TILE_WARPS_BWD=wb, FEO_TILE_LINES_BWD=lb)
        if rest:
            os.environ.update(FEO_TILE_FILL_GAP=rest[0], FEO_TILE_FILL_RESERVE=rest[1] if len(rest) > 1 else "6")
        if len(rest) > 3:
            os.environ.update(FEO_TILE_STAGES_FWD=rest[2], FEO_TILE_STAGES_BWD=rest[3])
    t0 = time.time()
    op = FEOperator(N, A=fx.A, B1=fx.B1, B2=fx.B2, idx_sol=fx.idx_sol, ns_precond_branch=True, device=dev)
    t_plan = time.time() - t0
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
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(K):
        loss, rT = op.residual_fwd(aT, fT, B)
        op.residual_bwd(aT, rT, B, out=gT)
    e1.record()
    torch.cuda.synchronize()
    print(f"cfg {cfg}: fwd {tf[len(tf)//2]:.3f} ms (min {tf[0]:.3f}), bwd {tb[len(tb)//2]:.3f} ms (min {tb[0]:.3f}); "
          f"back-to-back {e0.elapsed_time(e1) / K:.3f} ms; samples/s {B / ((tf[len(tf)//2] + tb[len(tb)//2]) * 1e-3):.0f}; "
          f"tiles {op.info.n_tiles_fwd}/{op.info.n_tiles_bwd}; plan {t_plan:.1f} s; loss {loss.item():.8e} "
          f"|g| {gT[:, :B].double().norm().item():.8e}", flush=True)
    del op, rT
