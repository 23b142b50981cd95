"""Developer timing of feo_dense_apply (preconditioned operator GEMM): CUDA events, fused residual epilogue.

usage: [FEO_DENSE_SIMT=1] time_dense.py [n] [B] [K]
Prints ms per apply, the fp32-equivalent TFLOP/s (2 n^2 B / t) and the tensor-pipe TFLOP/s (three TF32 products)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import feonet_navier_stokes_b200 as feo
from feonet_navier_stokes_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2549
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
D = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
op = feo.FEOperator(n, dense_m=D, device=dev)
xT = torch.randn(n, B, device=dev)
fT = torch.randn(n, B, device=dev)
for _ in range(5):
    op.dense_apply(L.FEO_DENSE_M, xT, B, sub=fT, want_loss=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    rT, loss = op.dense_apply(L.FEO_DENSE_M, xT, B, sub=fT, want_loss=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
ref = torch.tensor(D, device=dev, dtype=torch.float64) @ xT.double() - fT.double()
err = (rT.double() - ref).abs().max().item()
mode = "simt-fp32" if os.environ.get("FEO_DENSE_SIMT", "0") not in ("", "0") else "tcgen05-3xtf32"
print(f"{mode} n={n} B={B}: {ms:.4f} ms/apply (+loss finalize), {2 * n * n * B / ms / 1e9:.1f} fp32-equivalent TFLOP/s, "
      f"tensor {6 * n * n * B / ms / 1e9:.1f} TFLOP/s; max |err| vs fp64 {err:.3e}; loss {loss.item():.6e} vs {float((ref ** 2).sum()):.6e}")
