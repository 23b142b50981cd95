"""CUDA-graph capture of the launch-bound steps.

At the reference's own sizes (N = 387 ... 2 680, B = 1000) one residual loss + backward is a handful of
kernels of a few microseconds each; the step is bound by launch latency and Python, not by the GPU (0.19 - 0.35 ms per loss
fwd + bwd through the eager API, of which the kernels are ~10 %).  The library's launches are plain stream work -- no host
synchronisation, no allocation outside torch's allocator, tensor maps passed as kernel parameters -- so a whole step can be
captured once and replayed:

* `GraphedLossGrad`  -- `loss = fn(*inputs)` and `d loss / d inputs[wrt]` as one graph (the residual loss of `closure` and its
  backward, FEONet_steady_Navier-Stokes/train_FEONet.py:351-365, :463).
* `GraphedTrainStep` -- a full optimiser step (zero_grad, network forward, loss, backward, bad-value guard, optimiser update)
  as one graph: the body of the reference's epoch loop (:453-473).  The guard of :434-469 is evaluated on the device and
  handed to the fused optimiser as `found_inf` (the mechanism GradScaler uses), so a non-finite batch is skipped without a
  host read.

Static-buffer discipline: a graph replays fixed addresses.  Inputs are copied into buffers allocated at capture time (no copy
when the caller passes the very tensors the graph was captured on -- full-batch training on resident data, the reference's
default); outputs are views of graph-owned memory, valid until the next replay.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise ValueError("CUDA graphs capture device work: the inputs must be CUDA tensors")


def _same_memory(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype)


def _side_stream_warmup(fn: Callable[[], object], warmup: int):
    """Warm-up runs on a side stream (first-call set-up -- cudaFuncSetAttribute, workspaces, cuBLAS handles, lazily created
    optimiser state -- must happen before capture)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(max(1, warmup)):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()


class GraphedLossGrad:
    """`loss, grads = graphed(*inputs)`: `fn(*inputs) -> scalar loss` and its gradients w.r.t. `inputs[i], i in wrt`.

    `inputs` at construction are the STATIC buffers (the graph reads exactly these tensors); a later call with other tensors
    of the same shapes copies them in first."""

    def __init__(self, fn: Callable[..., torch.Tensor], inputs: Sequence[torch.Tensor], wrt: Sequence[int] = (0,), warmup: int = 3):
        for t in inputs:
            _require_cuda(t)
        self.fn = fn
        self.static = [t.detach() for t in inputs]
        self.wrt = list(wrt)

        def run():
            # fresh leaves over the static memory every time: a leaf that an earlier eager graph still references keeps its
            # AccumulateGrad node, and with it the (legacy) stream that node was created on -- the engine would then make
            # that stream wait for the capturing one, which invalidates the capture
            args = [t.detach().requires_grad_(True) if i in self.wrt else t for i, t in enumerate(self.static)]
            loss = fn(*args)
            grads = torch.autograd.grad(loss, [args[i] for i in self.wrt])
            return loss.detach(), grads

        _side_stream_warmup(run, warmup)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.grads = run()
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, ...]]:
        if inputs:
            if len(inputs) != len(self.static):
                raise ValueError(f"expected {len(self.static)} inputs, got {len(inputs)}")
            with torch.no_grad():
                for dst, src in zip(self.static, inputs):
                    if not _same_memory(src, dst):
                        if src.shape != dst.shape:
                            raise ValueError(f"graph captured for shape {tuple(dst.shape)}, got {tuple(src.shape)}")
                        dst.copy_(src)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.grads


class _StateSnapshot:
    """Model parameters / buffers and optimiser state before the warm-up steps, restored IN PLACE after capture (the graph has
    recorded the addresses): capturing does not advance training.  Optimiser state that the warm-up created lazily (Adam's
    moments and step counter, SGD momentum) is reset to zero, its value before the first step."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer):
        self.model, self.optimizer = model, optimizer
        self.tensors = [(t, t.detach().clone()) for t in list(model.parameters()) + list(model.buffers())]
        self.opt_before = {id(v): v.detach().clone() for st in optimizer.state.values() for v in st.values() if torch.is_tensor(v)}
        self.rng = torch.cuda.get_rng_state()

    def restore(self):
        with torch.no_grad():
            for t, saved in self.tensors:
                t.copy_(saved)
            for st in self.optimizer.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        saved = self.opt_before.get(id(v))
                        v.copy_(saved) if saved is not None else v.zero_()
        torch.cuda.set_rng_state(self.rng)


class GraphedTrainStep:
    """One optimiser step of the reference's epoch loop as a CUDA graph.

    step_body(batch) -> (loss, u_pred): zero_grad is done here, then `step_body` (network forward + loss), backward, the
    bad-value guard and `optimizer.step()` -- all captured.  `optimizer` must keep its state on the device and accept
    `found_inf` (torch's fused Adam / AdamW / SGD: `make_capturable_optimizer`).

    `batch` at construction is the dict of STATIC tensors; `__call__(batch)` copies other tensors of the same shapes in.
    Returns (loss, ok): device scalars (views of graph memory), `ok` = 1 when the step was applied."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, step_body: Callable[[Dict[str, torch.Tensor]], Tuple[torch.Tensor, torch.Tensor]],
                 batch: Dict[str, torch.Tensor], warmup: int = 3, keep_state: bool = True):
        for t in batch.values():
            _require_cuda(t)
        self.model, self.optimizer = model, optimizer
        self.static = dict(batch)
        dev = next(iter(batch.values())).device
        # the fused optimisers skip the update on the device when found_inf != 0 and divide the gradients by grad_scale
        self.found_inf = torch.zeros((), dtype=torch.float32, device=dev)
        self.grad_scale = torch.ones((), dtype=torch.float32, device=dev)

        def run():
            optimizer.zero_grad(set_to_none=True)
            loss, u_pred = step_body(self.static)
            loss.backward()
            # guards of the reference (FEONet_steady_Navier-Stokes/train_FEONet.py:434-469) as one device flag
            ok = torch.isfinite(loss) & torch.isfinite(u_pred).all()
            for p in model.parameters():
                if p.grad is not None:
                    ok = ok & torch.isfinite(p.grad).all()
            self.found_inf.copy_((~ok).to(torch.float32))
            optimizer.grad_scale, optimizer.found_inf = self.grad_scale, self.found_inf
            try:
                optimizer.step()
            finally:
                del optimizer.grad_scale, optimizer.found_inf
            return loss.detach(), ok

        snap = _StateSnapshot(model, optimizer) if keep_state else None
        _side_stream_warmup(run, warmup)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.ok = run()
        if snap is not None:
            snap.restore()
        self.replays = 0

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if batch is not None:
            with torch.no_grad():
                for k, dst in self.static.items():
                    src = batch[k]
                    if not _same_memory(src, dst):
                        if src.shape != dst.shape:
                            raise ValueError(f"graph captured for {k} of shape {tuple(dst.shape)}, got {tuple(src.shape)}")
                        dst.copy_(src)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.ok


def make_capturable_optimizer(name: str, params, lr: float) -> torch.optim.Optimizer:
    """The optimisers of the reference's `--optimizer` flag that can run inside a graph: device-resident state, one fused
    kernel per parameter group, `found_inf` honoured on the device."""
    params = list(params)
    if name == "Adam":
        return torch.optim.Adam(params, lr=lr, fused=True, capturable=True)
    if name == "AdamW":
        return torch.optim.AdamW(params, lr=lr, fused=True, capturable=True)
    if name == "SGD":
        return torch.optim.SGD(params, lr=lr, fused=True)
    raise ValueError(f"--cuda_graph supports Adam, AdamW and SGD; {name} keeps host-side state (LBFGS line search, Adagrad step counters)")
