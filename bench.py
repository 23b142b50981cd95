#!/usr/bin/env python
"""bench.py -- FEM residual fwd+bwd samples/s (BASELINE.json's metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm (oracle port on the host cores)

Workload (config.workload): BASELINE.json configs[4] -- synthetic structured P2-P1 channel
mesh, n=333 cells/side => N = 1 001 334 dofs, steady Navier-Stokes residual (A, B1, B2 as CSR,
no preconditioner), batch 1024 per GPU (weak scaling: samples are independent, the operator is
replicated, no data-path collective).  A "step" = residual loss forward + backward to
d loss / d alpha for one batch, through the public autograd API.

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fem_residual_fwd_bwd_samples_per_s"
UNIT = "samples/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (e.g. NCCL's version banner) write to file descriptor 1
# behind Python's back, so fd 1 is pointed at stderr for the whole run and the line goes to a private
# duplicate of the original stdout.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n", type=int, default=333, help="mesh cells per side (333 -> N=1 001 334)")
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU")
    ap.add_argument("--ordering", default="interleaved", choices=["interleaved", "blocked"])
    ap.add_argument("--layout", default="dof_major", choices=["dof_major", "row_major"],
                    help="memory layout of alpha/F/grad for the headline value")
    ap.add_argument("--cpu-samples", type=int, default=8, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-precond-gemm", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the DP training-step arm (FCNN head + gradient all-reduce)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check against the fp64 oracle")
    ap.add_argument("--train-steps", type=int, default=8)
    ap.add_argument("--configs", action="store_true", help="time cfg1-4 (parity-test cases) on cuda:0 beside the CPU port; not the bench line")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunk", type=int, default=256, help="samples per host->device chunk of the end-to-end leg")
    return ap.parse_args()


def workload_name(args, N):
    return (f"cfg5: synthetic structured P2-P1 channel mesh n={args.n} ({N} dofs, {args.ordering} dof order), "
            f"steady Navier-Stokes residual (A,B1,B2 CSR, no preconditioner), batch {args.batch}/GPU")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception as exc:  # pragma: no cover
            log(f"[bench] nvidia-smi unavailable: {exc}")
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # pragma: no cover
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                pw.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def build_fixture(args):
    from feonet_navier_stokes_b200.fixtures import config_operators

    t0 = time.time()
    fx = config_operators("steady_ns", args.n, ordering=args.ordering)
    log(f"[bench] fixture n={args.n}: N={fx.N} nnz(A,B1,B2)=({fx.A.nnz},{fx.B1.nnz},{fx.B2.nnz}) in {time.time() - t0:.1f}s")
    return fx


def cpu_port_rate(fx, args, steps, warmup, seed=0):
    """samples/s of the oracle's multi-threaded CPU port on a bounded sample of the workload."""
    from oracle.feonet_oracle import TorchCpuSteadyNS

    port = TorchCpuSteadyNS(fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, do_precond=True, threads=os.cpu_count())
    rng = np.random.default_rng(seed)
    bs = args.cpu_samples
    alpha = (0.1 * rng.standard_normal((bs, fx.N))).astype(np.float32)
    F = rng.standard_normal((bs, fx.N)).astype(np.float32)
    for _ in range(max(1, warmup)):
        port.loss_and_grad(alpha, F)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        port.loss_and_grad(alpha, F)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    return bs * len(times) / dt, port.threads, 1e3 * dt / len(times)


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is pure
    Python/PyTorch (nothing to compile into oracle/_ref) and stores its operators dense, which is
    infeasible at this workload (4 TB per matrix), so this arm runs the oracle's sparse CPU port of the
    same formula -- a generous stand-in for the reference's dense-GEMM CPU path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fx = build_fixture(args)
    steps, warmup = max(1, args.steps), max(0, args.warmup)  # one step = one pass of the bounded sample (about 0.1 s)
    rate, threads, ms = cpu_port_rate(fx, args, steps, warmup)
    sample = f"{args.cpu_samples} of the {args.batch} samples per step, full N={fx.N} operator, fp32, torch CPU sparse-CSR"
    out = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, fx.N), "device": "host CPU"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import feonet_navier_stokes_b200 as feo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    feo.load_library(build_if_missing=False)

    fx = build_fixture(args)
    N, B = fx.N, args.batch
    t0 = time.time()
    # --ordering blocked: the operator arrives in another numbering with the dof coordinates the reference's npz carries; it is
    # renumbered at set-up (reorder.py) and the row-major layout passes apply the permutation
    ns = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=True, precond=None, model_name="FCNN", device=dev,
                                dof_positions=None if args.ordering == "interleaved" else fx.pos)
    op = ns.operator
    log(f"[bench] rank {rank}: operator on device in {time.time() - t0:.1f}s: tiles(fwd,bwd)=({op.info.n_tiles_fwd},{op.info.n_tiles_bwd}) "
        f"nnz_union={op.info.nnz_union} device_MB={op.info.device_bytes / 2**20:.0f}")

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    native = args.layout == "dof_major"

    def make(scale):
        if native:
            t = feo.dof_major_empty(B, N, dev)
            t.normal_(0.0, scale, generator=gen)
            return t
        return torch.empty(B, N, device=dev).normal_(0.0, scale, generator=gen)

    alpha = make(0.1).requires_grad_(True)  # alpha ~ N(0, 0.1^2), F ~ N(0,1)  
    F = make(1.0)
    loss_buf = torch.zeros((), device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    loss_sum = torch.zeros((), device=dev, dtype=torch.float64)

    def step(e_mid=None):
        alpha.grad = None
        loss = ns.residual_loss(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_sol)
        if e_mid is not None:
            e_mid.record()
        loss.backward()
        loss_sum.add_(loss.detach())  # logging: the summed loss is all-reduced ONCE per reporting interval, not per step
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    K = args.steps
    marks = [(ev(), ev(), ev()) for _ in range(K)]
    launches0 = op.launches
    t_wall = time.perf_counter()
    loss_sum.zero_()
    for k in range(K):
        marks[k][0].record()
        loss = step(marks[k][1])
        marks[k][2].record()
    if world > 1:  # DP bookkeeping inside the timed region: one scalar all-reduce for the K steps (the reference logs every 100 epochs)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM)
    e_end = ev()
    e_end.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = op.launches - launches0

    total_ms = marks[0][0].elapsed_time(e_end)
    fwd_ms = sum(m[0].elapsed_time(m[1]) for m in marks) / K
    bwd_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / K
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)
    loss_val = float(loss.item())

    # same path with the reference's row-major [B,N] tensors (includes the layout transposes)
    alt_value = None
    if rank == 0 and native and world == 1:
        a2 = torch.empty(B, N, device=dev).normal_(0.0, 0.1, generator=gen).requires_grad_(True)
        F2 = torch.empty(B, N, device=dev).normal_(0.0, 1.0, generator=gen)

        def step2():
            a2.grad = None
            ns.residual_loss(a2, F2, fx.A, fx.B1, fx.B2, fx.idx_sol).backward()

        for _ in range(3):
            step2()
        e0, e1 = ev(), ev()
        k2 = max(3, K // 2)
        e0.record()
        for _ in range(k2):
            step2()
        e1.record()
        torch.cuda.synchronize()
        alt_value = B * k2 / (e0.elapsed_time(e1) * 1e-3)
        del a2, F2

    # end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, torch, feo, ns, fx, dev, world, rank)
        except Exception as exc:  # pragma: no cover
            log(f"[bench] e2e leg failed: {exc!r}")
            e2e = None

    train = None
    if not args.no_train_step:
        try:
            train = run_train_step(args, torch, feo, ns, fx, dev, world, rank)
        except Exception as exc:  # pragma: no cover
            log(f"[bench] train_step leg failed: {exc!r}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    parity = None
    if not args.no_parity and world == 1:
        try:
            parity = run_parity(args, torch, feo, ns, fx, dev)
        except Exception as exc:  # pragma: no cover
            log(f"[bench] parity leg failed: {exc!r}")

    peak, peak_src = measured_peak_gbs()
    alg_fwd = 12.0 * N * B  # read alpha, F; write r
    alg_bwd = 12.0 * N * B  # read r, alpha; write grad
    lattice = op.plan == "lattice"
    names = ("residual_lattice_kernel<fwd>", "residual_lattice_kernel<bwd>") if lattice else ("residual_fwd_tiled", "residual_bwd_tiled")
    dom = names[1] if bwd_ms >= fwd_ms else names[0]
    dom_ms, dom_alg = (bwd_ms, alg_bwd) if bwd_ms >= fwd_ms else (fwd_ms, alg_fwd)
    achieved = dom_alg / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tpath):  # dram__bytes of one `ncu --set full` capture (static: ncu cannot run inside the timed bench)
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get(dom), f"profiles/traffic_r02.json ({tj.get('source', 'ncu --set full')}; static, not measured in this run)"
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(args, N), "N": N, "batch_per_gpu": B, "layout": args.layout,
            "l2": "alpha and F are 4.1 GB each per GPU: every step streams far more than the 126 MB L2",
            "parallelism": f"dp{world} (batch sharded, operator replicated; `value` = the loss path: no data-path collective, one scalar "
                           f"loss all-reduce per {K} steps; the gradient all-reduce of a training step is measured in `train_step`)",
            "plan": "lattice" if lattice else "tile",
        },
        "clocks": clocks,
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
            "algorithmic_bytes_per_launch": dom_alg, "ms_per_launch": dom_ms,
            "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
            "step_algorithmic_GBs": 24.0 * N * B / (total_ms / K * 1e-3) / 1e9,
            "step_frac_of_measured": 24.0 * N * B / (total_ms / K * 1e-3) / 1e9 / peak,
            "step_frac_of_nominal_8TBs": 24.0 * N * B / (total_ms / K * 1e-3) / 1e9 / 8000.0,
        },
        "loss": loss_val,
        "wall_ms_per_step": 1e3 * t_wall / K,
    }
    if alt_value is not None:
        out["value_row_major_layout"] = alt_value
    if parity is not None:
        out["parity"] = parity
    if train is not None:
        out["train_step"] = train
    if e2e is not None:
        out["e2e"] = e2e
    if world == 1 and not args.no_precond_gemm:
        try:
            out["precond_gemm"] = run_precond_gemm(torch, feo, dev)
        except Exception as exc:  # pragma: no cover
            log(f"[bench] precond_gemm leg failed: {exc!r}")
    if not args.no_cpu_baseline and world == 1:
        try:
            rate, threads, ms = cpu_port_rate(fx, args, 3, 1)
            out["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{args.cpu_samples} of the {B} samples, full N={N} operator, fp32, oracle torch-CPU sparse-CSR port, 3 repeats",
            }
        except Exception as exc:  # pragma: no cover
            log(f"[bench] cpu_baseline failed: {exc!r}")
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_parity(args, torch, feo, ns, fx, dev):
    """The benchmarked operator against the fp64 oracle (oracle.ns_loss_and_grad, scipy CSR) on the cpu_baseline's samples,
    through the public autograd API in both layouts (FEONet_steady_Navier-Stokes/train_FEONet.py:301-365)."""
    from oracle import feonet_oracle as orc

    rng = np.random.default_rng(0)
    bs = args.cpu_samples
    alpha = (0.1 * rng.standard_normal((bs, fx.N))).astype(np.float32)
    F = rng.standard_normal((bs, fx.N)).astype(np.float32)
    lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, True, dtype=np.float64)
    out = {"samples": bs, "against": "oracle fp64 (scipy CSR), same samples as cpu_baseline", "tolerance": {"loss": 1e-5, "grad": 1e-4}}
    Fd = torch.tensor(F, device=dev)
    for name, native in (("dof_major", True), ("row_major", False)):
        a = torch.tensor(alpha, device=dev)
        a = feo.to_dof_major_tensor(a) if native else a.unsqueeze(1)
        a.requires_grad_(True)
        loss = ns.residual_loss(a, Fd, fx.A, fx.B1, fx.B2, fx.idx_sol)
        (g,) = torch.autograd.grad(loss, a)
        g = g.reshape(bs, fx.N).double().cpu().numpy()
        out[name] = {"rel_loss": abs(loss.item() - lo) / abs(lo), "rel_grad": float(np.linalg.norm(g - go) / np.linalg.norm(go)),
                     "rel_grad_max": float(np.abs(g - go).max() / np.abs(go).max())}
    out["ok"] = all(out[k]["rel_loss"] < 1e-5 and out[k]["rel_grad"] < 1e-4 and out[k]["rel_grad_max"] < 1e-4 for k in ("dof_major", "row_major"))
    log(f"[bench] parity {out}")
    return out


def run_train_step(args, torch, feo, ns, fx, dev, world, rank):
    """The data-parallel TRAINING step the loss path sits in (FEONet_steady_Navier-Stokes/train_FEONet.py:453-473 with the only
    model that is feasible at 1M dofs, FCNN(6 -> 16 -> 32 -> 64 -> 128 -> 256 -> N), :178-179): closure() = network forward +
    residual loss, loss.backward(), NCCL all-reduce(SUM) of the ~1.03 GB of parameter gradients, fused Adam step.  The head's
    weight gradient is formed in dof-range chunks and each chunk's all-reduce is launched while the later chunks' GEMMs run
    (parallel.OverlappedLinearT).  Reported per N: step time with / without the collective, the collective alone, the exposed
    fraction, weak (1024 samples per GPU) and strong (1024 / N samples per GPU) scaling, and -- at one GPU -- closure() with the
    reference's row-major nn.Linear head beside the dof-major head."""
    import torch.distributed as dist
    from feonet_navier_stokes_b200 import network, parallel

    N = fx.N
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    K = max(2, args.train_steps)

    def build(dof_major, overlapped):
        torch.manual_seed(7)
        model = network.FCNN(6, N, [16, 32, 64, 128, 256], dof_major_head=dof_major).to(dev)
        reducer = parallel.GradientReducer(enabled=overlapped)
        if overlapped:
            model.model[-1] = parallel.OverlappedLinearT.from_linear(model.model[-1], chunks=8, reducer=reducer)
        return model, reducer, torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

    def time_steps(model, reducer, optim, B, reduce_grads):
        gen = torch.Generator(device=dev).manual_seed(99 + rank)
        coeff = torch.rand(B, 6, device=dev, generator=gen)
        F = feo.dof_major_empty(B, N, dev)
        F.normal_(0.0, 1.0, generator=gen)
        head = list(model.model[-1].parameters())

        waits = []

        def step(timed=False):
            optim.zero_grad(set_to_none=True)
            loss, _ = ns.closure(model, coeff, None, F, fx.A, fx.B1, fx.B2, 64)
            loss.backward()
            if world > 1:
                if timed:  # how long the compute stream sits between the end of backward and the last reduced gradient
                    waits.append((ev(), ev()))
                    waits[-1][0].record()
                if reduce_grads:
                    parallel.allreduce_remaining(model, head, reducer)
                reducer.finish()  # hands the head's gradients to .grad (after their reduction, when one was launched)
                if timed:
                    waits[-1][1].record()
            optim.step()
            return loss

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(K):
            loss = step(timed=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        wait_ms = sum(a.elapsed_time(b) for a, b in waits) / max(1, len(waits))
        if world > 1:
            t = torch.tensor([ms, wait_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wait_ms = float(t[0].item()), float(t[1].item())
        del coeff, F
        return ms, float(loss.item()), wait_ms

    B = args.batch
    out = {"model": "FCNN(6,[16,32,64,128,256],N) tanh MLP, dropout 0.2, fp32 (TF32 off, torch default), fused Adam",
           "grad_bytes": None, "steps": K}
    model, reducer, optim = build(True, world > 1)
    out["params"] = sum(p.numel() for p in model.parameters())
    out["grad_bytes"] = 4 * out["params"]
    ms_full, loss, wait_full = time_steps(model, reducer, optim, B, True)
    out["weak"] = {"batch_per_gpu": B, "ms_per_step": ms_full, "samples_per_s": world * B / (ms_full * 1e-3), "loss": loss}
    if world > 1:
        reducer.enabled = False
        ms_noar, _, wait_noar = time_steps(model, reducer, optim, B, False)
        reducer.enabled = True
        # the collective alone: the same chunks, nothing to overlap with
        w = model.model[-1].weight
        bufs = list(torch.empty_like(w).chunk(8, dim=0))
        rest = torch.empty(out["params"] - w.numel(), device=dev)
        for _ in range(2):
            for b_ in bufs + [rest]:
                dist.all_reduce(b_)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(3):
            for b_ in bufs + [rest]:
                dist.all_reduce(b_)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1) / 3
        t = torch.tensor([ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
        del bufs, rest
        out["weak"].update({"ms_per_step_without_allreduce": ms_noar, "allreduce_alone_ms": ar_ms,
                            "allreduce_busbw_GBs": 2.0 * (world - 1) / world * out["grad_bytes"] / (ar_ms * 1e-3) / 1e9,
                            # exposed = what the compute stream waits between the end of backward and the last reduced gradient
                            # (CUDA events inside the step), minus the same interval of the run without the collective
                            "allreduce_exposed_ms": max(0.0, wait_full - wait_noar),
                            "allreduce_exposed_frac": max(0.0, wait_full - wait_noar) / ar_ms,
                            "step_difference_ms": ms_full - ms_noar,
                            "limiting_collective": "NCCL all-reduce(SUM) of the head weight gradient (256 x N fp32), 8 dof-range chunks"})
        Bs = max(1, B // world)
        ms_strong, _, _ = time_steps(model, reducer, optim, Bs, True)
        out["strong"] = {"batch_per_gpu": Bs, "global_batch": Bs * world, "ms_per_step": ms_strong,
                         "samples_per_s": world * Bs / (ms_strong * 1e-3)}
    del model, optim
    if world == 1:  # the drop-in question: what does closure() cost with the reference's row-major nn.Linear head?
        model, reducer, optim = build(False, False)
        ms_rm, _, _ = time_steps(model, reducer, optim, B, False)
        out["closure_row_major_head"] = {"ms_per_step": ms_rm, "samples_per_s": B / (ms_rm * 1e-3),
                                         "note": "network.FCNN with the reference's nn.Linear head ([B,1,N] row-major): the loss op transposes in and out"}
        del model, optim
        # where the step's time goes: the same step with PyTorch's own TF32 switch for the network's GEMMs (NOT the reference's
        # numerics, not part of `weak`; the loss path is unchanged fp32) -- the head's three 0.54-TFLOP fp32 GEMMs are PyTorch's part
        prev = torch.backends.cuda.matmul.allow_tf32
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            model, reducer, optim = build(True, False)
            ms_tf32, _, _ = time_steps(model, reducer, optim, B, False)
            out["weak_network_tf32"] = {"ms_per_step": ms_tf32, "samples_per_s": B / (ms_tf32 * 1e-3),
                                        "note": "torch.backends.cuda.matmul.allow_tf32 = True for the network layers only; informational"}
            del model, optim
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    torch.cuda.empty_cache()
    log(f"[bench] train_step {out}")
    return out


def run_precond_gemm(torch, feo, dev, n=2549, B=1024, iters=30):
    """The dense preconditioned apply r = (A P) alpha - F with its loss (SURVEY 8a a7, cfg2b size N=2549, B=1024):
    tcgen05 3xTF32 kernel behind feo_dense_apply, CUDA events.  Tensor-pipe work = three TF32 products per fp32 product;
    the peak beside it is MEASURED_PEAKS.json's dense bf16 figure halved (TF32 runs at half the bf16 rate)."""
    from feonet_navier_stokes_b200 import _lib as L
    import numpy as np

    rng = np.random.default_rng(0)
    D = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    op = feo.FEOperator(n, dense_m=D, device=dev)
    xT = torch.randn(n, B, device=dev)
    fT = torch.randn(n, B, device=dev)
    for _ in range(3):
        op.dense_apply(L.FEO_DENSE_M, xT, B, sub=fT, want_loss=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        rT, loss = op.dense_apply(L.FEO_DENSE_M, xT, B, sub=fT, want_loss=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ref = torch.tensor(D, device=dev, dtype=torch.float64) @ xT.double() - fT.double()
    loss_ref = float((ref ** 2).sum())
    bf16_peak = None
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        bf16_peak = float(pk["bf16_tflops"])  # burst figure: the kernel is timed alone
    except Exception:
        pass
    tensor_tf = 6.0 * n * n * B / (ms * 1e-3) / 1e12
    out = {
        "kernel": "dense_split_x_kernel + dense_apply_tc3_kernel (tcgen05 kind::tf32 cta_group::2 CTA pairs, 3xTF32 split, both operands pre-split and fetched by tensor-map loads, two TMEM accumulators, tile width per problem; dense_apply_tc2_kernel when the pairs do not fit one wave) + finalize_loss_kernel", "n": n, "batch": B,
        "ms_per_apply": ms, "fp32_equivalent_tflops": 2.0 * n * n * B / (ms * 1e-3) / 1e12, "tensor_pipe_tflops": tensor_tf,
        "loss_rel_err_vs_fp64": abs(loss.item() - loss_ref) / loss_ref,
        "max_abs_err_vs_fp64": float((rT[:, :B].double() - ref).abs().max()),
    }
    if bf16_peak:
        out["tf32_peak_tflops"] = bf16_peak / 2.0
        out["tensor_pipe_frac"] = tensor_tf / (bf16_peak / 2.0)
    return out


def run_configs(args):
    """--configs: the reference's own configurations (SURVEY 8 size table cfg1-4; parity-test cases, NOT the bench
    line) through the reference-facing API on cuda:0 -- residual loss + backward to grad alpha, row-major [B, N]
    inputs as the reference passes them -- with the oracle's CPU port (numpy/scipy, fp32) timed beside it; the loss / gradient differences are taken against
    the oracle evaluated in fp64.
    Prints one JSON object; `tools/...` summarise it into profiles/."""
    import torch
    import feonet_navier_stokes_b200 as feo
    from feonet_navier_stokes_b200.fixtures import config_operators
    from oracle import feonet_oracle as orc

    feo.load_library(build_if_missing=False)
    dev = torch.device("cuda:0")
    B, T, dt = 1000, 10, 0.1
    rng = np.random.default_rng(0)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def time_gpu(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def graph_ms(loss_fn, inp, eager_loss):
        """The same loss + gradient as ONE CUDA-graph replay (feo.GraphedLossGrad): what the step costs without launch latency."""
        try:
            gl = feo.GraphedLossGrad(loss_fn, [inp])
            ms = time_gpu(lambda: gl(), 50)
            l, _ = gl()
            return {"gpu_ms_fwd_bwd_graph": ms, "graph_loss_rel_diff_vs_eager": abs(l.item() - eager_loss) / abs(eager_loss)}
        except Exception as exc:  # pragma: no cover
            log(f"[configs] graph capture failed: {exc!r}")
            return {}

    def time_cpu(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        return 1e3 * (time.perf_counter() - t0) / reps, out

    def dense_precond(n):  # a dense, well-conditioned stand-in for the missing precond_{ne}.npy blobs
        return (np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)

    rows = []
    # the reference's own execution plan (dense operators, per-sample / per-dof Python loops, autograd) on this host's cores
    loops = orc.TorchReferenceLoops(os.cpu_count())

    peak_gbs, _ = measured_peak_gbs()

    def report(name, N, gpu_ms, cpu_ms, loss_gpu, loss_cpu, grad_gpu, grad_cpu, note, alg_bytes=None, alg_flops=None):
        g, c = np.asarray(grad_gpu, dtype=np.float64), np.asarray(grad_cpu, dtype=np.float64)
        rows.append({"config": name, "N": N, "batch": B, "gpu_ms_fwd_bwd": gpu_ms, "gpu_samples_per_s": B / (gpu_ms * 1e-3),
                     "cpu_port_ms_fwd_bwd": cpu_ms, "cpu_port_samples_per_s": B / (cpu_ms * 1e-3), "speedup": cpu_ms / gpu_ms,
                     "loss_rel_diff_vs_fp64": abs(loss_gpu - loss_cpu) / abs(loss_cpu),
                     "grad_rel_diff_vs_fp64": float(np.linalg.norm(g - c) / np.linalg.norm(c)), "note": note})
        if alg_bytes is not None:  # SURVEY 8d algorithmic bytes of the step (reference-facing row-major API: layout passes included in the time)
            rows[-1].update(algorithmic_bytes=alg_bytes, algorithmic_GBs=alg_bytes / (gpu_ms * 1e-3) / 1e9,
                            hbm_roofline_frac=alg_bytes / (gpu_ms * 1e-3) / 1e9 / peak_gbs)
        if alg_flops is not None:
            rows[-1].update(algorithmic_flops=alg_flops, fp32_equivalent_TFLOPs=alg_flops / (gpu_ms * 1e-3) / 1e12)
        log(f"[configs] {rows[-1]}")

    # cfg1 / cfg2: linear Stokes, preconditioned (dense apply on the tensor cores)
    for name, fxname, n in (("cfg1 Stokes_square precond N=387", "stokes_square", 6), ("cfg2a square-with-hole precond", "hole", 10),
                             ("cfg2b square-with-hole precond", "hole", 18)):
        fx = config_operators(fxname, n)
        A = np.asarray(fx.A.todense(), dtype=np.float32)
        P = dense_precond(fx.N)
        alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
        F = rng.standard_normal((B, fx.N)).astype(np.float32)
        At, Pt = torch.tensor(A, device=dev), torch.tensor(P, device=dev)
        st = feo.LinearStokes(At, Pt, do_precond=True, device=dev)
        a = torch.tensor(alpha, device=dev, requires_grad=True)
        Ft = torch.tensor(F, device=dev)
        box = {}

        def step():
            loss = st.residual_loss(a, Ft, At, Pt)
            (box["g"],) = torch.autograd.grad(loss, a)
            box["l"] = loss

        gpu_ms = time_gpu(step)
        cpu_ms, _ = time_cpu(lambda: orc.stokes_loss_and_grad(alpha, F, A, P, True, dtype=np.float32))
        lo, go, _ = orc.stokes_loss_and_grad(alpha, F, A, P, True, dtype=np.float64)  # yardstick for the differences
        report(name + (f" N={fx.N}" if "N=" not in name else ""), fx.N, gpu_ms, cpu_ms,
               box["l"].item(), float(lo), box["g"].cpu().numpy(), go, "dense A.P folded at set-up, tcgen05 3xTF32 apply",
               alg_flops=4.0 * fx.N * fx.N * B)  # one GEMM forward, one backward
        rows[-1].update(graph_ms(lambda a_: st.residual_loss(a_, Ft, At, Pt), a, box["l"].item()))
        ref_ms, (l_ref, g_ref) = time_cpu(lambda: loops.linear_stokes_step(alpha, F, A, P, True), reps=1)
        rows[-1].update(reference_loops_ms_fwd_bwd=ref_ms, reference_loops_samples_per_s=B / (ref_ms * 1e-3),
                        speedup_vs_reference_loops=ref_ms / gpu_ms, loss_rel_diff_vs_reference_loops=abs(box["l"].item() - l_ref) / abs(l_ref))

    # cfg3: steady Navier-Stokes, both sign branches (fused sparse kernels)
    fx = config_operators("steady_ns", 15)
    alpha = (0.3 * rng.standard_normal((B, fx.N))).astype(np.float32)
    F = rng.standard_normal((B, fx.N)).astype(np.float32)
    for precond in (True, False):
        nsm = feo.SteadyNavierStokes(fx.A, fx.B1, fx.B2, fx.idx_sol, do_precond=precond, device=dev)
        a = torch.tensor(alpha, device=dev, requires_grad=True)
        Ft = torch.tensor(F, device=dev)
        box = {}

        def step():
            loss = nsm.residual_loss(a, Ft, fx.A, fx.B1, fx.B2, fx.idx_sol)
            (box["g"],) = torch.autograd.grad(loss, a)
            box["l"] = loss

        gpu_ms = time_gpu(step)
        cpu_ms, _ = time_cpu(lambda: orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, precond, dtype=np.float32))
        lo, go, _ = orc.ns_loss_and_grad(alpha, F, fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, precond, dtype=np.float64)
        report(f"cfg3 steady NS N={fx.N} ({'precond=I' if precond else 'no precond'} sign branch)", fx.N, gpu_ms, cpu_ms,
               box["l"].item(), float(lo), box["g"].cpu().numpy(), go, "fused residual kernels incl. row-major <-> dof-major transposes",
               alg_bytes=24.0 * fx.N * B)
        rows[-1].update(graph_ms(lambda a_: nsm.residual_loss(a_, Ft, fx.A, fx.B1, fx.B2, fx.idx_sol), a, box["l"].item()))
        dense = [np.asarray(K.todense(), dtype=np.float32) for K in (fx.A, fx.B1, fx.B2)]
        ref_ms, (l_ref, g_ref) = time_cpu(lambda: loops.steady_ns_step(alpha, F, *dense, fx.idx_u1, fx.idx_u2, precond), reps=1)
        rows[-1].update(reference_loops_ms_fwd_bwd=ref_ms, reference_loops_samples_per_s=B / (ref_ms * 1e-3),
                        speedup_vs_reference_loops=ref_ms / gpu_ms, loss_rel_diff_vs_reference_loops=abs(box["l"].item() - l_ref) / abs(l_ref))

    # cfg4: time-dependent Stokes, T = 10
    fx = config_operators("time_dep", 10)
    pred = (0.3 * rng.standard_normal((B, T, fx.N))).astype(np.float32)
    u0 = rng.standard_normal((B, fx.N)).astype(np.float32)
    F = np.repeat(rng.standard_normal((1, fx.N)).astype(np.float32), B, axis=0)
    td = feo.TimeDependentStokes(fx.S, fx.A, fx.idx_sol, dt=dt, do_precond=False, device=dev)
    pt = torch.tensor(pred, device=dev, requires_grad=True)
    Ft, u0t = torch.tensor(F, device=dev), torch.tensor(u0, device=dev)
    box = {}

    def step():
        loss = td.residual_loss(pt, Ft, fx.S, fx.A, None, dt, u0t)
        (box["g"],) = torch.autograd.grad(loss, pt)
        box["l"] = loss

    gpu_ms = time_gpu(step)
    cpu_ms, _ = time_cpu(lambda: orc.seq_loss_and_grad(pred, F, fx.S, fx.A, None, dt, u0, False, dtype=np.float32))
    lo, go, _ = orc.seq_loss_and_grad(pred, F, fx.S, fx.A, None, dt, u0, False, dtype=np.float64)
    report(f"cfg4 time-dependent Stokes N={fx.N} T={T}", fx.N, gpu_ms, cpu_ms, box["l"].item(), float(lo),
           box["g"].cpu().numpy(), go, "seq_kernel fwd/bwd, one sample = T rows; launch- and L2-latency-bound at this size (41 MB per tensor)",
           alg_bytes=float(B) * (T * 20.0 * fx.N + 4.0 * fx.N))
    rows[-1].update(graph_ms(lambda p_: td.residual_loss(p_, Ft, fx.S, fx.A, None, dt, u0t), pt, box["l"].item()))
    # the linear Stokes operator (no convective term: A-quads forward, 20 N B algorithmic bytes) at the cfg5 mesh size
    large = None
    try:
        from feonet_navier_stokes_b200.operator import FEOperator

        fx = config_operators("stokes_square", args.n, ordering=args.ordering)
        Bl, Nl = 1024, fx.N
        op = FEOperator(Nl, A=fx.A, device=dev)
        aT = torch.empty(Nl, Bl, device=dev).normal_(0, 0.1)
        fT = torch.empty(Nl, Bl, device=dev).normal_(0, 1.0)
        gT = torch.empty(Nl, Bl, device=dev)
        tf = time_gpu(lambda: op.residual_fwd(aT, fT, Bl), 10)
        _, rT = op.residual_fwd(aT, fT, Bl)
        tb = time_gpu(lambda: op.residual_bwd(aT, rT, Bl, out=gT), 10)
        peak, _ = measured_peak_gbs()
        large = {"config": f"linear Stokes at the cfg5 mesh (N={Nl}, B={Bl}, dof-major)", "fwd_ms": tf, "bwd_ms": tb,
                 "samples_per_s": Bl / ((tf + tb) * 1e-3), "fwd_algorithmic_GBs": 12.0 * Nl * Bl / (tf * 1e-3) / 1e9,
                 "bwd_algorithmic_GBs": 8.0 * Nl * Bl / (tb * 1e-3) / 1e9, "hbm_peak_GBs": peak}
        log(f"[configs] {large}")
    except Exception as exc:  # pragma: no cover
        log(f"[configs] large linear case failed: {exc!r}")
    for r in rows:
        if "algorithmic_bytes" in r and "gpu_ms_fwd_bwd_graph" in r:
            r["hbm_roofline_frac_graph"] = r["algorithmic_bytes"] / (r["gpu_ms_fwd_bwd_graph"] * 1e-3) / 1e9 / peak_gbs
    shell = None
    try:
        shell = run_train_shell_timing(torch, dev)
    except Exception as exc:  # pragma: no cover
        log(f"[configs] train-shell timing failed: {exc!r}")
    emit({"configs": rows, "linear_large": large, "train_shell": shell, "cores": os.cpu_count(), "cpu_kind": "port (oracle numpy/scipy, fp32)",
          "reference_loops": "oracle.TorchReferenceLoops: the reference's own execution plan (dense fp32 operators, (matrix @ precond).mm per "
                             "sample, one MSE(sum) per dof in a Python loop, autograd backward; FEONet_steady_Navier-Stokes/train_FEONet.py:301-360, "
                             "FEONet_Stokes_square/train_FEONet.py:261-296) restated and timed on this host with torch CPU, all cores; "
                             "bit-equal to the unmodified reference functions in the authoring container (profiles/r02_reference_cpu_container.json)"})


def run_train_shell_timing(torch, dev, steps=100):
    """One optimiser step of the training shell (network forward, residual loss, backward, bad-value guard, Adam) at the
    reference's sizes, eager against ONE CUDA-graph replay (`--cuda_graph 1`, graphs.GraphedTrainStep): the body of the
    reference's epoch loop, FEONet_steady_Navier-Stokes/train_FEONet.py:453-473.  Full-batch, data resident on the device."""
    import tempfile

    from feonet_navier_stokes_b200 import train_FEONet as T

    out = []
    cases = (("stokes_square FCNN preconditioned (tensor-core apply)", ["--variant", "stokes_square", "--train_file", "1000N72", "--val_file", "8N72",
                                                                         "--model", "FCNN", "--do_precond", "1", "--spai_steps", "50"]),
             ("stokes_square Net2D", ["--variant", "stokes_square", "--train_file", "1000N72", "--val_file", "8N72", "--model", "Net2D", "--do_precond", "0"]),
             ("time_dep RNN T=10", ["--variant", "time_dep", "--train_file", "256N200", "--val_file", "8N200", "--model", "RNN", "--seq_len", "10"]))
    for name, argv in cases:
        row = {"case": name}
        for graph in (0, 1):
            with tempfile.TemporaryDirectory() as tmp:
                g = dict(T.build_parser().parse_args(argv + ["--optimizer", "Adam", "--out", tmp, "--cuda_graph", str(graph)]).__dict__)
                tr = T.Trainer(g, device=dev)
                batch = next(tr.batches(tr.train, None, shard=True))
                for _ in range(5):
                    tr.train_step(batch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    loss, _ = tr.train_step(batch)
                torch.cuda.synchronize()
                ms = 1e3 * (time.perf_counter() - t0) / steps
                row.update({"N": tr.N, "batch": int(next(iter(batch.values())).shape[0]), ("graph_ms_per_step" if graph else "eager_ms_per_step"): ms,
                            ("graph_loss" if graph else "eager_loss"): float(loss.item())})
        row["speedup"] = row["eager_ms_per_step"] / row["graph_ms_per_step"]
        log(f"[configs] train shell {row}")
        out.append(row)
    return out


def run_e2e(args, torch, feo, ns, fx, dev, world, rank):
    """Public-API call with host inputs: alpha and F start in pinned host memory in the reference's
    row-major [B,N] layout, are copied to the device, go through residual_loss + backward, and the
    loss is read back to the host -- every step."""
    import torch.distributed as dist

    N, B = fx.N, args.batch
    from feonet_navier_stokes_b200.parallel import bind_to_gpu_numa_node

    numa = bind_to_gpu_numa_node(dev) if world > 1 else None  # the pinned batches below land on the GPU's own NUMA node
    a_host = torch.empty(B, N, pin_memory=True)
    f_host = torch.empty(B, N, pin_memory=True)
    chunk = min(B, args.e2e_chunk)
    for c0 in range(0, B, chunk):  # fill the host batches without holding a second full copy on the device
        c1 = min(B, c0 + chunk)
        a_host[c0:c1].copy_(torch.empty(c1 - c0, N, device=dev).normal_(0.0, 0.1))
        f_host[c0:c1].copy_(torch.empty(c1 - c0, N, device=dev).normal_(0.0, 1.0))
    torch.cuda.synchronize()
    grad_dev = torch.empty(B, N, device=dev)
    pipe = feo.HostBatchPipeline(lambda a, f: ns.residual_loss(a, f, fx.A, fx.B1, fx.B2, fx.idx_sol), N, dev, chunk=chunk)

    def step():
        return pipe.step(a_host, f_host, grad_out=grad_dev)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k = max(1, args.e2e_steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        loss_host = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * B * k / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * 4 * N * B,
            "d2h_bytes_per_step": 4, "steps": k, "ms_per_step": ms / k, "loss": loss_host,
            "note": f"feo.HostBatchPipeline: pinned host row-major alpha,F -> H2D in chunks of {chunk} samples on a copy stream, "
                    "overlapped with layout transposes + fused fwd+bwd of the previous chunk -> gradients on the device, loss D2H",
            "numa_node_rank0": numa}


def main():
    args = parse_args()
    claim_stdout()
    if args.configs:
        run_configs(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
