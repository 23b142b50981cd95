"""Training shell with the reference's command line, running on fixture data.

Mirrors what `train_FEONet.py` does around the hot path: flags
(`FEONet_steady_Navier-Stokes/train_FEONet.py:27-52`, `FEONet_Stokes_square/train_FEONet.py:27-50`), model zoo
dispatch (:176-189), optimizer factories (:247-290), epoch loop with bad-value guards (:424-478), validation
rel-L2 per component (:490-538), text log and checkpoints (:377-402, :540-569).  Differences, all outside the
loss path:

* the FEniCS npz / pkl inputs do not exist offline, so `--train_file 256N72` is read as "256 samples on the
  ne = 72 structured mesh" and the data are synthesised from the fixture assembler with the reference's
  sampling law (`assemble_fenics.py:123-124`, seeds 5 / 10 as in `create_data.py:30-33`); reference solutions
  come from a sparse direct solve (linear Stokes) or a Newton iteration on the algebraic system (steady NS,
  `compare_ordering_nonlinear.ipynb#c25`);
* the loss and its backward are `feonet_navier_stokes_b200` kernels (there is no CPU path);
* the guards read ONE device flag per step instead of one `.item()` per dof and per check; a step whose loss, prediction or
  any parameter gradient is not finite is SKIPPED (no optimizer step) -- the reference's `continue` in its gradient scan
  (:465-469) only leaves the inner loop over parameters and still steps with the bad gradient;
* `--variant time_dep` drives the time-dependent closure (`FEONet_time_dep_Stokes/train_FEONet.py:364-406`, loop :491-506) with the
  RNN model (`VectorToSequenceRNN`; flags `--seq_len --rnn_type --hidden_dim --num_layers --dt` as there); initial velocities and
  the implicit-Euler reference trajectories (`create_data.py:75-91`) are synthesised;
* `torchrun` launches shard the sample batch over ranks (parallel.py), gradients are summed with NCCL.

    python -m feonet_navier_stokes_b200.train_FEONet --variant steady_ns --bc channel_flow --forcing_term sincos \
        --train_file 256N450 --val_file 64N450 --model FCNN --optimizer Adam --do_precond 1 --epochs 200
"""
from __future__ import annotations

import argparse
import datetime
import math
import os
import time
from typing import Dict, Optional, Tuple

import numpy as np
import torch

VARIANTS = ("stokes_square", "hole", "steady_ns", "time_dep")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("SEM")
    p.add_argument("--variant", type=str, default="steady_ns", choices=VARIANTS,
                   help="which of the reference's directories this run stands for")
    # settings (reference names)
    p.add_argument("--bc", type=str, default="channel_flow", choices=["lower", "channel_flow"])
    p.add_argument("--forcing_term", type=str, default="sincos", choices=["sincos"])
    p.add_argument("--train_file", type=str, default="256N450", help="Example: --train_file 3000N18  (samples N elements)")
    p.add_argument("--val_file", type=str, default="64N450")
    p.add_argument("--domain", type=str, default="dolfin", choices=["dolfin", "square"])
    # train parameters (reference names)
    p.add_argument("--pretrained", type=str, default=None)
    p.add_argument("--model", type=str, default="FCNN", choices=["Net2D", "FCNN", "UNetWithHead", "RNN"],
                   help="RNN = VectorToSequenceRNN, the time-dependent variant's default (FEONet_time_dep_Stokes/train_FEONet.py:45-46)")
    # time-dependent variant (FEONet_time_dep_Stokes/train_FEONet.py:37-58)
    p.add_argument("--dt", type=float, default=0.1)
    p.add_argument("--seq_len", type=int, default=10)
    p.add_argument("--rnn_type", type=str, default="gru", choices=["gru", "lstm"])
    p.add_argument("--hidden_dim", type=int, default=512)
    p.add_argument("--num_layers", type=int, default=1)
    p.add_argument("--optimizer", type=str, default="Adam", choices=["LBFGS", "Adam", "SGD", "AdamW", "Adagrad"])
    p.add_argument("--do_precond", type=int, default=0)
    p.add_argument("--batch_size_train", type=int, default=None)
    p.add_argument("--batch_size_val", type=int, default=None)
    p.add_argument("--resol_in", type=int, default=20)
    p.add_argument("--blocks", type=int, default=0)
    p.add_argument("--ks", type=int, default=5)
    p.add_argument("--filters", type=int, default=32, choices=[8, 16, 32, 64])
    p.add_argument("--loss", type=str, default="MSE", choices=["MAE", "MSE", "RMSE", "RelMSE"])
    p.add_argument("--epochs", type=int, default=80000)
    p.add_argument("--pre_epochs", type=int, default=0)
    # additions
    p.add_argument("--lr", type=float, default=1e-3)
    p.add_argument("--log_every", type=int, default=100, help="validation / checkpoint period in epochs (reference: 100)")
    p.add_argument("--out", type=str, default="runs")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--spai_steps", type=int, default=200, help="minimal-residual iterations of the SPAI preconditioner")
    p.add_argument("--dof_major_head", type=int, default=1, help="let the last layer write the coefficients dof-major")
    p.add_argument("--cuda_graph", type=int, default=0,
                   help="capture the whole optimiser step (network, loss, backward, guard, update) in a CUDA graph: one GPU, Adam / AdamW / SGD")
    p.add_argument("--npz", type=str, default=None,
                   help="an assemble_fenics.py npz (data_ordered/P2x1_ne..._BC[_force].npz) to train on instead of synthesised data")
    return p


def parse_file_flag(s: str) -> Tuple[int, int]:
    """'1000N72' -> (1000 samples, ne = 72) as in FEONet_Stokes_square/train_FEONet.py:52-61."""
    num, ne = s.split("N")
    return int(num), int(ne)


def mesh_n_from_ne(ne: int, strict: bool = True) -> int:
    n = int(round(math.sqrt(ne / 2.0)))
    if strict and 2 * n * n != ne:
        raise ValueError(f"ne={ne} is not 2 n^2: the offline fixture only has structured n x n meshes")
    return n


def sample_coeff_f(num: int, seed: int) -> np.ndarray:
    """m0, m1 ~ U(0,1); n0..n3 ~ pi U(0,1)  (assemble_fenics.py:123-124)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.0, 1.0, size=(num, 6))
    c[:, 2:] *= np.pi
    return c


def newton_steady_ns(fx, L: np.ndarray, precond_branch: bool, iters: int = 12) -> np.ndarray:
    """Solves the algebraic system the loss penalises: A u + s (d1 o B1 u + d2 o B2 u) = s F, s = +1 on the
    precond branch (r = A u - F + c), -1 otherwise (r = A u + F - c); one sample."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    A, B1, B2 = fx.A.tocsr(), fx.B1.tocsr(), fx.B2.tocsr()
    N = fx.N
    I, J = np.asarray(fx.idx_u1), np.asarray(fx.idx_u2)
    rhs = L if precond_branch else -L
    s = 1.0 if precond_branch else -1.0
    u = spla.spsolve(A.tocsc(), rhs)
    for _ in range(iters):
        d1, d2 = np.zeros(N), np.zeros(N)
        d1[I] = d1[J] = u[I]
        d2[I] = d2[J] = u[J]
        bu1, bu2 = B1 @ u, B2 @ u
        res = A @ u + s * (d1 * bu1 + d2 * bu2) - rhs
        if np.linalg.norm(res) < 1e-11 * max(1.0, np.linalg.norm(rhs)):
            break
        # dc/du = diag(d1) B1 + diag(d2) B2 + diag(bu1) E1 + diag(bu2) E2, E1[r, I(r)] = 1, E2[r, J(r)] = 1 on velocity rows
        rows = np.concatenate([I, J])
        E1 = sp.csr_matrix((np.ones(2 * I.size), (rows, np.concatenate([I, I]))), shape=(N, N))
        E2 = sp.csr_matrix((np.ones(2 * I.size), (rows, np.concatenate([J, J]))), shape=(N, N))
        Jm = A + s * (sp.diags(d1) @ B1 + sp.diags(d2) @ B2 + sp.diags(bu1) @ E1 + sp.diags(bu2) @ E2)
        u = u - spla.spsolve(Jm.tocsc(), res)
    return u


def synthesize(variant: str, n: int, bc: str, num: int, seed: int, precond_branch: bool):
    """(fixture, dict of numpy arrays: coeff_f [num,6], load_vec_f [num,N], fenics_u1/u2 [num,n_u], fenics_p [num,n_p])."""
    import scipy.sparse.linalg as spla

    from .fixtures import assemble_operators, config_operators, structured_mesh

    if bc == "channel_flow":
        fx = config_operators(variant, n, ordering="interleaved")
    else:
        fx = assemble_operators(structured_mesh(n), mu=0.1, bc=bc, ordering="interleaved", with_convection=variant == "steady_ns")
    coeff = sample_coeff_f(num, seed)
    L = fx.load_vector_sincos(coeff)
    if variant == "steady_ns":
        U = np.stack([newton_steady_ns(fx, L[b], precond_branch) for b in range(num)])
    else:
        lu = spla.splu(fx.A.tocsc())
        U = np.stack([lu.solve(L[b]) for b in range(num)])
    data = {"coeff_f": coeff, "load_vec_f": L, "fenics_u1": U[:, fx.idx_u1], "fenics_u2": U[:, fx.idx_u2], "fenics_p": U[:, fx.idx_p]}
    return fx, data


def synthesize_time_dep(n: int, num: int, seed: int, dt: float, seq_len: int):
    """(fixture, dict of numpy arrays) for the time-dependent variant: initial velocities init_x, init_y [num, n_u] (zero on the
    walls, amplitudes ~ U(0,1) as `coeffs_init`), the constant-force load vector replicated [num, N]
    (`FEONet_time_dep_Stokes/train_FEONet.py:235,244`) and the implicit-Euler trajectories (S + dt A) u+ = S u + dt f
    [num, seq_len + 1, N] that `create_data.py:75-91` stores as `coeffs_u` (here by a sparse LU in fp64)."""
    import scipy.sparse.linalg as spla

    from .fixtures import config_operators

    fx = config_operators("time_dep", n, ordering="interleaved")
    rng = np.random.default_rng(seed)
    amp = rng.uniform(0.0, 1.0, size=(num, 2))
    x, y = fx.mesh.p2_xy[:, 0], fx.mesh.p2_xy[:, 1]
    bump = np.sin(np.pi * x) * y * (1.0 - y)
    init_x, init_y = amp[:, :1] * bump[None, :], 0.5 * amp[:, 1:] * bump[None, :] * np.cos(np.pi * x)[None, :]
    load = fx.load_vector_sincos(np.array([[1.0, 0.0, 0.0, 0.0, 0.0, 0.0]]))[0] * 0.0  # f = 0 in the interior ...
    load[fx.bc_dofs] = fx.bc_vals  # ... and the boundary values on the Dirichlet rows (constant in time)
    M = (fx.S + dt * fx.A).tocsc()
    lu = spla.splu(M)
    U = np.zeros((num, seq_len + 1, fx.N))
    U[:, 0, fx.idx_u1], U[:, 0, fx.idx_u2] = init_x, init_y
    for t in range(seq_len):
        U[:, t + 1] = lu.solve((fx.S @ U[:, t].T + dt * load[:, None])).T
    data = {"coeff_f": amp, "init_x": init_x, "init_y": init_y, "load_vec_f": np.repeat(load[None], num, axis=0),
            "coeffs_u": U, "fenics_u1": U[:, 1:, fx.idx_u1].reshape(num, -1), "fenics_u2": U[:, 1:, fx.idx_u2].reshape(num, -1),
            "fenics_p": U[:, 1:, fx.idx_p].reshape(num, -1)}
    return fx, data


def make_model(name: str, resol_in: int, d_out: int, filters: int, ks: int, blocks: int, dof_major_head: bool, gparams=None):
    from . import network as net

    pad = (ks - 1) // 2
    if name == "RNN":
        g = gparams or {}
        return net.VectorToSequenceRNN(input_dim=d_out, hidden_dim=g.get("hidden_dim", 512), output_dim=d_out,
                                       rnn_type=g.get("rnn_type", "gru"), num_layers=g.get("num_layers", 1))
    if name == "Net2D":
        return net.Net2D(resol_in, 2, filters, d_out, kernel_size=ks, padding=pad, blocks=blocks, dof_major_head=dof_major_head)
    if name == "FCNN":
        return net.FCNN(6, d_out, hidden_dims=[16, 32, 64, 128, 256], dof_major_head=dof_major_head)
    return net.UNetWithHead(resol_in=resol_in, in_ch=2, base_ch=32, latent_ch=64, d_out=d_out, head_filters=filters,
                            head_blocks=blocks, head_kernel_size=ks, head_padding=pad, dof_major_head=dof_major_head)


def make_optimizer(name: str, model: torch.nn.Module, lr: float) -> torch.optim.Optimizer:
    params = model.parameters()
    if name == "LBFGS":
        return torch.optim.LBFGS(params, lr=lr)
    if name == "SGD":
        return torch.optim.SGD(params, lr=lr)
    if name == "AdamW":
        return torch.optim.AdamW(params, lr=lr)
    if name == "Adagrad":
        return torch.optim.Adagrad(params, lr=lr)
    return torch.optim.Adam(params, lr=lr)


class Trainer:
    def __init__(self, gparams: Dict, device: Optional[torch.device] = None):
        import feonet_navier_stokes_b200 as feo
        from . import parallel

        self.g = gparams
        self.feo, self.parallel = feo, parallel
        self.rank, self.world, local = parallel.init_distributed()
        self.device = device if device is not None else torch.device("cuda", local)
        torch.manual_seed(gparams["seed"])
        variant = gparams["variant"]
        n_train, ne = parse_file_flag(gparams["train_file"])
        n_val, ne_val = parse_file_flag(gparams["val_file"])
        if ne != ne_val:
            raise ValueError("train and validation files must use the same mesh")
        do_precond = int(gparams["do_precond"]) > 0
        if gparams.get("npz"):
            # the reference's own file: dense operators -> CSR, the first NUM_DATA samples of each split (:69-85, :218-245)
            from types import SimpleNamespace

            from .data_io import load_reference_npz

            z = load_reference_npz(gparams["npz"])
            i_, j_, k_ = z["idx_sol"]
            self.fx = SimpleNamespace(N=z["N"], A=z["A"] if "A" in z else z["matrix"], B1=z.get("B1"), B2=z.get("B2"),
                                      idx_u1=np.asarray(i_), idx_u2=np.asarray(j_), idx_p=np.asarray(k_), idx_sol=z["idx_sol"], pos=z.get("p"))
            train = {k: v[:n_train] for k, v in z["train"].items() if k != "forcing_term"}
            val = {k: v[:n_val] for k, v in z["validate"].items() if k != "forcing_term"}
            n_train, n_val = len(train["coeff_f"]), len(val["coeff_f"])
        else:
            # data (seeds 5 / 10: create_data.py:30-33); ne fixes the synthetic mesh only here -- a reference npz carries its own
            n = mesh_n_from_ne(ne, strict=variant != "hole")  # the hole stand-in mesh is Delaunay: ne only sets its resolution
            if variant == "time_dep":
                self.fx, train = synthesize_time_dep(n, n_train, 5, gparams["dt"], gparams["seq_len"])
                _, val = synthesize_time_dep(n, n_val, 10, gparams["dt"], gparams["seq_len"])
            else:
                self.fx, train = synthesize(variant, n, gparams["bc"], n_train, 5, do_precond)
                _, val = synthesize(variant, n, gparams["bc"], n_val, 10, do_precond)
        t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)  # noqa: E731
        if n_train < self.world:
            raise ValueError(f"{n_train} training samples cannot be sharded over {self.world} ranks")
        # every rank keeps the whole (small) training set and takes ITS SHARD OF EACH GLOBAL BATCH (batches()), so that all
        # ranks run the same number of steps per epoch -- every step issues collectives
        self.train = {k: t(v).to(self.device) for k, v in train.items()}
        self.val = {k: t(v).to(self.device) for k, v in val.items()}
        self.N, self.n_u = self.fx.N, len(self.fx.idx_u1)
        self.idx = [torch.tensor(np.asarray(i), device=self.device, dtype=torch.long) for i in (self.fx.idx_u1, self.fx.idx_u2, self.fx.idx_p)]
        # operator state (the reference's module globals)
        A = self.fx.A
        self.P = None
        if variant == "time_dep":
            # FEONet_time_dep_Stokes/train_FEONet.py:364-406, loop :491-506; the sequence model sees the assembled u_init
            if gparams["model"] != "RNN":
                raise ValueError("the time-dependent shell drives the RNN model (VectorToSequenceRNN); UNet1D / UNet2D / UNetTemporal "
                                 "are image models of the reference that this repo does not mirror")
            if do_precond:
                self.P = feo.spai_device(self.fx.S + gparams["dt"] * A, gparams["spai_steps"], self.device).to(torch.float32).cpu()
            self.problem = feo.TimeDependentStokes(self.fx.S, A, self.fx.idx_sol, dt=gparams["dt"], do_precond=do_precond, precond=self.P,
                                                   model_name="RNN", device=self.device)
            S_, T_ = self.fx.S, gparams["seq_len"]
            self.closure_args = lambda b: (None, b["init_x"], b["init_y"], b["load_vec_f"], S_, A, None, self.P, gparams["dt"], T_)  # noqa: E731
        elif variant == "steady_ns":
            # PRECOND = np.eye(N) whenever --do_precond > 0 (steady NS :142): the identity is detected, never stored densely
            self.problem = feo.SteadyNavierStokes(A, self.fx.B1, self.fx.B2, self.fx.idx_sol, do_precond=do_precond, precond=None,
                                                  model_name=gparams["model"], force=gparams["forcing_term"], device=self.device,
                                                  dof_positions=getattr(self.fx, "pos", None))
            self.closure_args = lambda b: (b["coeff_f"], None, b["load_vec_f"], A, self.fx.B1, self.fx.B2, gparams["resol_in"])  # noqa: E731
        else:
            if do_precond:
                self.P = feo.spai_device(A, gparams["spai_steps"], self.device).to(torch.float32).cpu()
            hole = variant == "hole"
            self.problem = feo.LinearStokes(A, self.P, do_precond=do_precond, model_name=gparams["model"], force=gparams["forcing_term"],
                                            hole_signature=hole, device=self.device, idx_sol=self.fx.idx_sol,
                                            dof_positions=getattr(self.fx, "pos", None))
            if hole:
                self.closure_args = lambda b: (b["coeff_f"], None, b["load_vec_f"], A, self.P, gparams["resol_in"])  # noqa: E731
            else:
                self.closure_args = lambda b: (b["coeff_f"], b["load_vec_f"], A, self.P, gparams["resol_in"])  # noqa: E731
        self.model = make_model(gparams["model"], gparams["resol_in"], self.N, gparams["filters"], gparams["ks"], gparams["blocks"],
                                bool(gparams["dof_major_head"]), gparams).to(self.device)
        if gparams["pretrained"]:
            self.model.load_state_dict(torch.load(gparams["pretrained"], map_location=self.device))
        self.reducer, self.head_params = None, []
        if self.world > 1:
            parallel.broadcast_parameters(self.model)
            if gparams["model"] == "FCNN" and gparams["optimizer"] != "LBFGS":
                # the head holds almost all parameters (256 x N): its gradient all-reduce overlaps its own backward GEMMs
                self.reducer = parallel.GradientReducer()
                self.model.model[-1] = parallel.OverlappedLinearT.from_linear(self.model.model[-1], chunks=8, reducer=self.reducer)
                self.head_params = list(self.model.model[-1].parameters())
            if any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in self.model.modules()):
                self.model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(self.model)  # full-batch statistics as on one GPU
        self.graphed = bool(gparams.get("cuda_graph", 0))
        self._graphs: Dict[int, object] = {}
        if self.graphed:
            from .graphs import make_capturable_optimizer

            if self.world > 1:
                raise ValueError("--cuda_graph captures a single-GPU step (the overlapped gradient all-reduce runs on a second stream)")
            self.optimizer = make_capturable_optimizer(gparams["optimizer"], self.model.parameters(), gparams["lr"])
        else:
            self.optimizer = make_optimizer(gparams["optimizer"], self.model, gparams["lr"])
        self.losses, self.train_err, self.test_err = [], [], []
        stamp = str(datetime.datetime.now()).replace(" ", "T").replace(":", "").split(".")[0].replace("-", "")
        self.folder = os.path.join(gparams["out"], str(ne), gparams["bc"], gparams["forcing_term"], f"{gparams['model']}_epochs{gparams['epochs']}_{stamp}")
        self.log_path = None
        if self.rank == 0:
            os.makedirs(self.folder, exist_ok=True)
            self.log_path = os.path.join(self.folder, "training_log.txt")
            with open(self.log_path, "w") as f:
                f.write(f"params: {sum(p.numel() for p in self.model.parameters())}\n{self.model}\n{self.optimizer}\n{gparams}\n")

    # -- pieces of the epoch loop ---------------------------------------------------------------------
    def batches(self, data: Dict[str, torch.Tensor], batch_size: Optional[int], shard: bool = False):
        """Global batches of `batch_size` samples (the whole set when unset, as the reference's DataLoader default).  With
        `shard`, a rank gets its contiguous part of every global batch: the number of batches is the same on every rank; a
        tail smaller than the world size is merged into the batch before it so that no rank ever holds an empty shard."""
        n = data["coeff_f"].shape[0]
        bs = n if not batch_size else min(batch_size, n)
        if shard:
            bs = max(bs, self.world)  # a global batch has at least one sample per rank
        edges = list(range(0, n, bs)) + [n]
        if shard and len(edges) > 2 and edges[-1] - edges[-2] < self.world:
            del edges[-2]
        for lo, hi in zip(edges[:-1], edges[1:]):
            if shard and self.world > 1:
                a, b = self.parallel.shard_bounds(hi - lo, self.rank, self.world)
                lo, hi = lo + a, lo + b
            yield {k: v[lo:hi] for k, v in data.items()}

    def train_step_graphed(self, batch) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same step as one CUDA-graph replay (graphs.GraphedTrainStep): no host read, the guard skips the update on the
        device.  One graph per batch size; full-batch training (the reference's default) replays on the resident tensors
        themselves, mini-batches are copied into the graph's static buffers."""
        from .graphs import GraphedTrainStep

        nb = next(iter(batch.values())).shape[0]
        step = self._graphs.get(nb)
        if step is None:
            full = nb == next(iter(self.train.values())).shape[0]

            class _Seen(dict):  # which tensors of a batch the closure reads (targets such as fenics_u1 stay out of the graph)
                used = set()

                def __getitem__(self, k):
                    self.used.add(k)
                    return dict.__getitem__(self, k)

            probe = _Seen(batch)
            self.closure_args(probe)
            static = {k: (batch[k] if full else batch[k].clone()) for k in sorted(probe.used)}
            body = lambda b: self.problem.closure(self.model, *self.closure_args(b))  # noqa: E731
            step = self._graphs[nb] = GraphedTrainStep(self.model, self.optimizer, body, static)
        return step(batch)

    def train_step(self, batch) -> Tuple[torch.Tensor, bool]:
        if self.graphed:
            return self.train_step_graphed(batch)
        self.optimizer.zero_grad(set_to_none=True)
        loss, u_pred = self.problem.closure(self.model, *self.closure_args(batch))
        loss.backward()
        if self.reducer is not None:  # head gradients were reduced during backward; the small layers follow in one bucket
            self.parallel.allreduce_remaining(self.model, self.head_params, self.reducer)
            self.reducer.finish()
        # bad-value guards of the reference (:434-469) folded into one device flag, read once
        ok = torch.isfinite(loss) & torch.isfinite(u_pred).all()
        for p in self.model.parameters():
            if p.grad is not None:
                ok = ok & torch.isfinite(p.grad).all()
        if self.world > 1:
            flag = ok.to(torch.float32)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            ok = flag > 0
        if not bool(ok.item()):
            return loss.detach(), False  # skip this batch
        if self.reducer is None:
            self.parallel.allreduce_gradients(self.model)
        if isinstance(self.optimizer, torch.optim.LBFGS):
            def reeval():
                self.optimizer.zero_grad(set_to_none=True)
                l2, _ = self.problem.closure(self.model, *self.closure_args(batch))
                l2.backward()
                self.parallel.allreduce_gradients(self.model)
                return l2
            self.optimizer.step(reeval)
        else:
            self.optimizer.step()
        return self.parallel.allreduce_loss(loss), True

    @torch.no_grad()
    def evaluate(self, data: Dict[str, torch.Tensor], batch_size: Optional[int]) -> Dict[str, float]:
        self.model.eval()
        err = {"u1": 0.0, "u2": 0.0, "p": 0.0, "all": 0.0}
        n = 0
        for b in self.batches(data, batch_size):
            _, u_pred = self.problem.closure(self.model, *self.closure_args(b))
            nb = b["coeff_f"].shape[0]
            if u_pred.dim() == 3 and u_pred.shape[1] != 1:  # time-dependent: [B, T, N] against the T later slices of the trajectory
                parts = [u_pred[:, :, i].reshape(nb, -1) for i in self.idx]
                u = u_pred
            else:
                u = u_pred.reshape(nb, self.N)
                parts = [u[:, self.idx[0]], u[:, self.idx[1]], u[:, self.idx[2]]]
            true = [b["fenics_u1"], b["fenics_u2"], b["fenics_p"]]
            for key, pr, tr in zip(("u1", "u2", "p"), parts, true):
                err[key] += float(self.feo.rel_L2_error(pr, tr).sum().item())
            err["all"] += float(self.feo.rel_L2_error(torch.cat(parts, 1), torch.cat(true, 1)).sum().item())
            n += u.shape[0]
        self.model.train()
        return {k: v / max(n, 1) for k, v in err.items()}

    def log(self, msg: str):
        if self.rank == 0:
            print(msg, flush=True)
            with open(self.log_path, "a") as f:
                f.write(msg + "\n")

    def save(self):
        if self.rank != 0:
            return
        torch.save({"model_state_dict": self.model.state_dict(), "losses": self.losses, "train_rel_L2_errors": self.train_err,
                    "test_rel_L2_errors": self.test_err}, os.path.join(self.folder, "model.pt"))

    def fit(self) -> Dict[str, float]:
        g = self.g
        t0 = t_last = time.time()
        skipped = 0
        last = {}
        for epoch in range(1, g["epochs"] + 1):
            self.model.train()
            loss_total = torch.zeros((), device=self.device)
            skipped_dev = torch.zeros((), device=self.device)
            for batch in self.batches(self.train, g["batch_size_train"], shard=True):
                loss, stepped = self.train_step(batch)
                if torch.is_tensor(stepped):  # graphed step: the flag stays on the device until the next log line
                    loss_total += torch.where(stepped, loss, torch.zeros_like(loss))
                    skipped_dev += (~stepped).to(skipped_dev.dtype)
                elif stepped:
                    loss_total += loss
                else:
                    skipped += 1
            if epoch % g["log_every"] == 0 or epoch == g["epochs"]:
                skipped += int(skipped_dev.item())
                lt = float(loss_total.item())
                self.losses.append(lt)
                tr = self.evaluate(self.train, g["batch_size_val"])
                te = self.evaluate(self.val, g["batch_size_val"])
                self.train_err.append(tr)
                self.test_err.append(te)
                now = time.time()
                self.log(f"Epoch {epoch:6d} | loss {lt:.6e} | {now - t_last:.2f}s | train rel-L2 u1 {tr['u1']:.4f} u2 {tr['u2']:.4f} p {tr['p']:.4f} "
                         f"| val rel-L2 u1 {te['u1']:.4f} u2 {te['u2']:.4f} p {te['p']:.4f} all {te['all']:.4f} | skipped {skipped}")
                t_last = now
                self.save()
                last = {"loss": lt, **{f"val_{k}": v for k, v in te.items()}}
        self.log(f"Total training time {time.time() - t0:.1f}s")
        if self.rank == 0:
            torch.save(self.model.state_dict(), os.path.join(self.folder, "model.pth"))
        return last


def main(argv=None) -> Dict[str, float]:
    args = build_parser().parse_args(argv)
    return Trainer(dict(args.__dict__)).fit()


if __name__ == "__main__":
    main()
