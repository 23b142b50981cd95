"""Developer timing of the layout passes at the cfg5 size: row-major [B, N] <-> dof-major [N, ldb] (feo_transpose), with and
without a dof permutation folded in.  usage: time_transpose.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from feonet_navier_stokes_b200 import _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1001334
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lib = L.load_library(build_if_missing=False)
dev = torch.device("cuda:0")
x = torch.randn(B, N, device=dev)
xT = torch.empty(N, B, device=dev)
y = torch.empty(B, N, device=dev)
perm = torch.tensor(np.random.default_rng(0).permutation(N).astype(np.int32), device=dev)
# blocked -> interleaved-like map: locally regular (stride 2-3), as reorder.lattice_permutation produces
reg = torch.tensor((np.arange(N, dtype=np.int64) * 3 % N if N % 3 else np.arange(N)).astype(np.int32), device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())


def timed(fn, k=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


gb = 8.0 * N * B / 1e9
for name, m in (("plain", None), ("regular map", reg), ("random map", perm)):
    t_in = timed(lambda: L.check(lib.feo_transpose(p(x), N, p(xT), B, B, N, p(m), st)))
    if m is None:
        t_out = timed(lambda: L.check(lib.feo_transpose(p(xT), B, p(y), N, N, B, None, st)))
    else:
        t_out = timed(lambda: L.check(lib.feo_transpose_gather(p(xT), B, p(y), N, N, B, p(m), st)))
    print(f"{name}: to dof-major {t_in:.3f} ms ({gb / t_in:.2f} TB/s), to row-major {t_out:.3f} ms ({gb / t_out:.2f} TB/s)")
t = timed(lambda: xT.copy_(y.view(N, B)))
print(f"device copy of the same bytes: {t:.3f} ms ({gb / t:.2f} TB/s)")
