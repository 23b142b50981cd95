"""Host-resident batches: stream pinned host tensors through the device in chunks.

The reference moves every batch to the device inside the epoch loop (`sample_batch.to(device)`,
FEONet_steady_Navier-Stokes/train_FEONet.py:438-439).  At 1M dofs a batch of 1024 samples is 4 GB per
tensor, so the copy dominates the step; `HostBatchPipeline` splits the batch into chunks, copies chunk
c + 1 on a side stream while chunk c runs through the residual loss and its backward, and sums the loss
(the loss is a plain sum over samples, train_FEONet.py:354-360, so chunking does not change it).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class HostBatchPipeline:
    """loss_fn(alpha_chunk [n,N] requires_grad, F_chunk [n,N]) -> scalar loss (e.g. SteadyNavierStokes.residual_loss
    with the operator arguments bound).  `step` returns the summed loss as a Python float (one device->host read)."""

    def __init__(self, loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], N: int, device, chunk: int = 256):
        self.loss_fn, self.N, self.chunk = loss_fn, int(N), int(chunk)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.bufs = [(torch.empty(chunk, N, device=self.device), torch.empty(chunk, N, device=self.device)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.loss = torch.zeros((), device=self.device)

    def step(self, alpha_host: torch.Tensor, f_host: torch.Tensor, grad_out: Optional[torch.Tensor] = None) -> float:
        assert alpha_host.is_pinned() and f_host.is_pinned(), "host batches must be pinned for asynchronous copies"
        B = alpha_host.shape[0]
        cur = torch.cuda.current_stream(self.device)
        self.loss.zero_()
        for k in range(2):
            self.free[k].record(cur)
        for ci, c0 in enumerate(range(0, B, self.chunk)):
            c1 = min(B, c0 + self.chunk)
            k = ci & 1
            a_d, f_d = self.bufs[k][0][: c1 - c0], self.bufs[k][1][: c1 - c0]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.free[k])  # the chunk that used this buffer two iterations ago is done
                a_d.copy_(alpha_host[c0:c1], non_blocking=True)
                f_d.copy_(f_host[c0:c1], non_blocking=True)
                self.ready[k].record(self.copy_stream)
            cur.wait_event(self.ready[k])
            a = a_d.detach().requires_grad_(True)
            loss = self.loss_fn(a, f_d)
            loss.backward()
            self.loss += loss.detach()
            if grad_out is not None:
                grad_out[c0:c1].copy_(a.grad)
            self.free[k].record(cur)
        return float(self.loss.item())
