"""TEST INFRASTRUCTURE -- loads the reference's own hot-path functions, unmodified.

`/root/reference/*/train_FEONet.py` are top-level scripts (argparse at import, matplotlib,
author-local data paths), so they cannot be imported.  Their hot-path functions are plain
`def`s that only read module globals; this module AST-extracts those `FunctionDef` nodes and
`exec`s them in a namespace that supplies the globals.  The code that
runs is the reference's code, byte for byte -- nothing is copied into this repository.

Only `oracle/make_golden.py` (run in the authoring container, where `/root/reference` exists)
uses this.  It is never imported by the product package, the GPU tests, `smoke()` or `bench.py`.
"""
from __future__ import annotations

import ast
import os
from typing import Dict, Iterable

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("FEONET_REFERENCE_ROOT", "/root/reference")

VARIANTS = {
    "steady_ns": "FEONet_steady_Navier-Stokes/train_FEONet.py",
    "stokes_square": "FEONet_Stokes_square/train_FEONet.py",
    "hole": "FEONet-square-with-hole/train_FEONet.py",
    "time_dep": "FEONet_time_dep_Stokes/train_FEONet.py",
}
WANTED = ("weak_form", "closure", "weak_form_sequence", "assemble_u_init", "rel_L2_error", "spai")


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(REFERENCE_ROOT, p)) for p in VARIANTS.values())


def load_reference_functions(variant: str, globals_in: Dict, names: Iterable[str] = WANTED) -> Dict:
    """Return a namespace holding the reference's functions for `variant`.

    `globals_in` supplies what the functions read as module globals, e.g. DO_PRECOND, PRECOND,
    IDX_SOL, NUM_PTS, device, FORCE, gparams, criterion_wf, DT, BC, P.
    """
    path = os.path.join(REFERENCE_ROOT, VARIANTS[variant])
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in set(names)]
    ns = {
        "torch": torch,
        "np": np,
        "device": torch.device("cpu"),
        "criterion_wf": torch.nn.MSELoss(reduction="sum"),
    }
    try:  # `spai` uses these (FEONet_Stokes_square/train_FEONet.py:11-13,104-121)
        from scipy.sparse import identity
        from scipy.sparse.linalg import onenormest
        from tqdm import tqdm

        ns.update(identity=identity, onenormest=onenormest, tqdm=tqdm)
    except Exception:  # pragma: no cover
        pass
    ns.update(globals_in)
    exec(compile(ast.Module(body=funcs, type_ignores=[]), path, "exec"), ns)
    return ns


def make_idx_sol(idx_u1, idx_u2, idx_p) -> np.ndarray:
    """What `np.load(npz, allow_pickle=True)['idx_sol']` yields: object array of 3 int lists."""
    out = np.empty(3, dtype=object)
    out[0], out[1], out[2] = list(map(int, idx_u1)), list(map(int, idx_u2)), list(map(int, idx_p))
    return out
