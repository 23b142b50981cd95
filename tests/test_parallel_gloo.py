"""CPU, world_size 2, gloo: batch sharding + SUM gradient all-reduce reproduce the single-process
full-batch step (the N>1 path of SURVEY.md section 8e).  The loss here is plain torch -- the test
covers the data-parallel plumbing, not the CUDA kernels."""
import os
import socket
import tempfile

import torch
import torch.multiprocessing as mp

from feonet_navier_stokes_b200 import parallel as par


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(seed=0):
    torch.manual_seed(seed)
    N, B = 23, 10
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, N))
    A = torch.randn(N, N)
    x = torch.randn(B, 6)
    F = torch.randn(B, N)
    return model, A, x, F


def _loss(model, A, x, F):
    r = model(x) @ A.T - F
    return (r * r).sum(), None  # SUM over samples, like the reference


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = par.init_distributed("gloo")
    assert (r, w) == (rank, world)
    model, A, x, F = _make(seed=rank)  # different init per rank ...
    par.broadcast_parameters(model)     # ... made identical by the broadcast
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    _, A0, x0, F0 = _make(seed=0)
    xs, Fs = par.shard_batch(x0, rank, world), par.shard_batch(F0, rank, world)
    gloss, _ = par.dp_step(model, lambda: _loss(model, A0, xs, Fs), opt, bucket_mb=0.0005)
    torch.save({"loss": gloss, "params": [p.detach().clone() for p in model.parameters()]}, os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_step_matches_single_process():
    world, port = 2, _free_port()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, d), nprocs=world, join=True)
        outs = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    model, A, x, F = _make(seed=0)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    opt.zero_grad()
    loss, _ = _loss(model, A, x, F)
    loss.backward()
    opt.step()
    for o in outs:
        assert torch.allclose(o["loss"], loss.detach(), rtol=1e-5)
        for p, q in zip(o["params"], model.parameters()):
            assert torch.allclose(p, q.detach(), rtol=1e-5, atol=1e-6)


def test_shard_bounds_cover_batch():
    for n in (1, 7, 1000, 1024):
        for world in (1, 2, 3, 8):
            spans = [par.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _worker_overlapped(rank, world, port, out_dir):
    """The bench's DP training step (parallel.OverlappedLinearT + GradientReducer + allreduce_remaining): the head's weight
    gradient is all-reduced chunk by chunk DURING backward, the other layers afterwards."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    par.init_distributed("gloo")
    model, A0, x0, F0 = _make(seed=0)
    reducer = par.GradientReducer()
    assert reducer.enabled
    model[-1] = par.OverlappedLinearT.from_linear(model[-1], chunks=3, reducer=reducer)
    head = list(model[-1].parameters())
    xs, Fs = par.shard_batch(x0, rank, world), par.shard_batch(F0, rank, world)
    loss, _ = _loss(model, A0, xs, Fs)
    loss.backward()
    par.allreduce_remaining(model, head, reducer)
    nbytes = reducer.finish()
    assert nbytes == 4 * (head[0].numel() + head[1].numel())
    torch.save({"grads": [p.grad.detach().clone() for p in model.parameters()], "loss": par.allreduce_loss(loss)}, os.path.join(out_dir, f"o{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_overlapped_head_allreduce_matches_single_process():
    world, port = 2, _free_port()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker_overlapped, args=(world, port, d), nprocs=world, join=True)
        outs = [torch.load(os.path.join(d, f"o{r}.pt")) for r in range(world)]
    model, A, x, F = _make(seed=0)
    loss, _ = _loss(model, A, x, F)
    loss.backward()
    for o in outs:
        assert torch.allclose(o["loss"], loss.detach(), rtol=1e-5)
        for g, p in zip(o["grads"], model.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-4, atol=1e-5)  # SUM over ranks = the full-batch gradient
