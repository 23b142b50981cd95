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
