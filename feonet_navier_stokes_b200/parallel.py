"""Sample-batch data parallelism for the FEONet training loop.

The reference is single-device.  The FEM residual shards naturally: samples are independent and
the operator (A, B1, B2, S, P, idx) is small and read-only, so every rank holds a replica of the
operator handle and a contiguous shard of the batch; the residual kernels need no communication.
The only exchange is the all-reduce of the network-parameter gradients -- SUM, not mean, because
the reference loss is a SUM over samples (steady NS train_FEONet.py:298, :360; the time-dependent
variant divides by T only, :400) -- plus an optional scalar loss all-reduce for logging.

One process per GPU, `torch.distributed` over NCCL (NVLink 5 / NVSwitch); `gloo` on CPU for tests.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise from torchrun's environment. Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa_node(device) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: /sys/bus/pci/devices/<bus id>/numa_node), so
    that pinned host batches allocated afterwards live in that node's memory: with one rank per GPU on a two-socket host, the
    host->device copies of ranks whose pages sit on the other socket cross the inter-socket link (the 8-GPU end-to-end run
    moved fewer bytes per second than the 4-GPU one).  Returns the node, or None when the topology is not visible (containers
    without sysfs, single-node hosts) -- then nothing is changed."""
    try:
        dev = torch.device(device)
        p = torch.cuda.get_device_properties(dev)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def _buckets(params: Iterable[torch.nn.Parameter], bucket_bytes: int) -> List[List[torch.nn.Parameter]]:
    out, cur, size = [], [], 0
    for p in params:
        if p.grad is None:
            continue
        nbytes = p.grad.numel() * p.grad.element_size()
        if cur and (size + nbytes > bucket_bytes or cur[0].grad.dtype != p.grad.dtype):
            out.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nbytes
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(module: torch.nn.Module, bucket_mb: float = 64.0, async_op: bool = True) -> None:
    """SUM all-reduce of every parameter gradient, flattened into buckets.

    NVSwitch gives every GPU full bandwidth to every peer, so buckets are sized for launch latency and
    overlap (tens of MB), not for link count.  Buckets are launched asynchronously back to back and
    waited on together."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    work = []
    for bucket in _buckets(reversed(list(module.parameters())), int(bucket_mb * 2 ** 20)):
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)
        work.append((h, flat, bucket))
    for h, flat, bucket in work:
        if async_op:
            h.wait()
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n


class GradientReducer:
    """Asynchronous SUM all-reduce of gradient pieces as they become available during backward.

    `launch(t)` enqueues an all-reduce of `t` (in place) behind the work already queued on the current stream and returns at
    once; the collective runs on the process group's own stream, so the GEMMs the caller queues next overlap it.
    `deposit(param, grad)` registers the tensor the pieces belong to: it is handed to `param.grad` only in `finish()`, AFTER
    every piece has been reduced -- a tensor returned to autograd while its reduction is still in flight would be cloned or
    read by AccumulateGrad before the sum has landed.  `finish()` makes the current stream wait for every launched piece and
    assigns the deposited gradients (call it after `backward()`, before the optimizer step; required at any world size)."""

    def __init__(self, enabled: bool = True):
        self.enabled = bool(enabled) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.work = []
        self.deposits = []
        self.bytes = 0

    def launch(self, t: torch.Tensor) -> None:
        if not self.enabled:
            return
        self.bytes += t.numel() * t.element_size()
        self.work.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True))

    def deposit(self, param: torch.nn.Parameter, grad: torch.Tensor) -> None:
        self.deposits.append((param, grad))

    def finish(self) -> int:
        for h in self.work:
            h.wait()  # stream-level wait for NCCL work: the host does not block
        for param, grad in self.deposits:
            param.grad = grad if param.grad is None else param.grad + grad
        n, self.work, self.deposits, self.bytes = self.bytes, [], [], 0
        return n


class _OverlappedLinearTFn(torch.autograd.Function):
    """y = x W^T + b with a dof-major result (as network.LinearT); the backward forms dW in `chunks` row blocks (dof
    ranges) and hands each block to the reducer as soon as its GEMM is queued, so the all-reduce of block k runs while the
    GEMMs of the blocks after it (and dx) execute.  dW and db reach `.grad` through the reducer (`finish()`), not through
    autograd."""

    @staticmethod
    def forward(ctx, x, weight, bias, chunks: int, reducer: GradientReducer):
        B = x.shape[0]
        pad = (-B) % 4
        xt = (torch.nn.functional.pad(x, (0, 0, 0, pad)) if pad else x).t()  # [in, ceil4(B)]
        yt = torch.addmm(bias.unsqueeze(1), weight, xt) if bias is not None else weight @ xt
        ctx.save_for_backward(x, weight)
        ctx.bias, ctx.chunks, ctx.reducer, ctx.weight = bias, int(chunks), reducer, weight
        return yt[:, :B].t()

    @staticmethod
    def backward(ctx, g):  # g: [B, out]
        x, weight = ctx.saved_tensors
        gT = g.t()  # [out, B]; a plain strided view when g is dof-major
        out_f = weight.shape[0]
        dW = torch.empty_like(weight)
        step = (out_f + ctx.chunks - 1) // ctx.chunks
        for r0 in range(0, out_f, step):
            r1 = min(out_f, r0 + step)
            torch.mm(gT[r0:r1], x, out=dW[r0:r1])
            ctx.reducer.launch(dW[r0:r1])
        ctx.reducer.deposit(ctx.weight, dW)
        if ctx.bias is not None:
            db = gT.sum(dim=1)
            ctx.reducer.launch(db)
            ctx.reducer.deposit(ctx.bias, db)
        dx = g @ weight if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None


class OverlappedLinearT(torch.nn.Linear):
    """Drop-in for the final `nn.Linear` / `network.LinearT` of a FEONet model under data parallelism (same parameters
    and state_dict keys): dof-major output, gradient all-reduce overlapped with its own backward GEMMs.  Its parameter
    gradients appear in `.grad` when `reducer.finish()` is called (after `backward()`)."""

    def __init__(self, in_features, out_features, bias=True, chunks: int = 8, reducer: Optional[GradientReducer] = None):
        super().__init__(in_features, out_features, bias=bias)
        self.chunks = chunks
        self.reducer = reducer if reducer is not None else GradientReducer()

    @classmethod
    def from_linear(cls, lin: torch.nn.Linear, chunks: int = 8, reducer: Optional[GradientReducer] = None):
        m = cls.__new__(cls)
        torch.nn.Module.__init__(m)
        m.in_features, m.out_features = lin.in_features, lin.out_features
        m.weight, m.bias = lin.weight, lin.bias
        m.chunks = chunks
        m.reducer = reducer if reducer is not None else GradientReducer()
        return m

    def forward(self, x):
        if x.dim() != 2:
            return super().forward(x)
        return _OverlappedLinearTFn.apply(x, self.weight, self.bias, self.chunks, self.reducer)


def allreduce_remaining(module: torch.nn.Module, skip: Iterable[torch.nn.Parameter], reducer: GradientReducer) -> None:
    """Gradients that were not reduced during backward (everything but the overlapped head): one flat bucket."""
    if not reducer.enabled:
        return
    skip_ids = {id(p) for p in skip}
    ps = [p for p in module.parameters() if p.grad is not None and id(p) not in skip_ids]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def allreduce_loss(loss: torch.Tensor) -> torch.Tensor:
    """Summed loss over ranks (for logging; matches the reference's full-batch SUM loss)."""
    out = loss.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def dp_step(model: torch.nn.Module, closure, optimizer: torch.optim.Optimizer, bucket_mb: float = 64.0):
    """One data-parallel training step: `closure()` returns (loss, u_pred) for THIS rank's shard (what
    the reference's epoch loop calls at steady NS :453); gradients are summed over ranks before the
    optimizer step.  Returns (global_loss, u_pred)."""
    optimizer.zero_grad(set_to_none=True)
    loss, u_pred = closure()
    loss.backward()
    allreduce_gradients(model, bucket_mb)
    optimizer.step()
    return allreduce_loss(loss), u_pred
