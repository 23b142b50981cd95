"""On-disk formats of the reference pipeline.

* `data_ordered/P2x1_ne{ne}_stokes_{bc}_BC[_{force}].npz` written by `assemble_fenics.py`
  (`FEONet_Stokes_square/assemble_fenics.py:209-222`: key `matrix`; `FEONet_steady_Navier-Stokes/assemble_fenics.py
  :350-366`: keys `A, B1, B2`, `*_forcing_term`; `FEONet_time_dep_Stokes/assemble_fenics.py:358-371`: `S, A,
  load_vector`) -- operators are DENSE float64 [N, N] with identity Dirichlet rows, `idx_sol` is an object array of
  three int lists;
* `{file}.pkl` written by `create_data.py` (`FEONet_steady_Navier-Stokes/create_data.py:77-112`): an object ndarray of
  rows `[f_value, coeff_f]` (steady NS) or `[coeff_u, f_value, coeff_f]` (linear Stokes).

`load_reference_npz` turns the dense operators into CSR with the reference's `.float()` cast semantics (an entry is
kept iff it is non-zero after the fp64 -> fp32 cast: quirk 10), which is what `FEOperator` consumes; `save_reference_npz`
writes fixture data in the same schema so that files from either source are interchangeable.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp

OPERATOR_KEYS = ("matrix", "A", "B1", "B2", "S")
SAMPLE_KEYS = ("coeff_fs", "forcing_term", "load_vectors", "fenics_u1", "fenics_u2", "fenics_p", "coeffs_init", "values_init")


def npz_name(ne: int, bc: str, force: Optional[str] = None, dt: Optional[float] = None) -> str:
    """File name convention of the four `assemble_fenics.py` variants."""
    name = f"P2x1_ne{ne}_stokes_{bc}_BC"
    if force:
        name += f"_{force}"
    if dt is not None:
        name += "_dt_" + str(dt).replace(".", "_")
    return name + ".npz"


def dense_to_csr(K: np.ndarray) -> sp.csr_matrix:
    """Dense [N,N] (any float dtype) -> CSR of the fp32-cast values, entries kept iff != 0 after the cast."""
    K32 = np.asarray(K).astype(np.float32)
    rows, cols = np.nonzero(K32)
    return sp.csr_matrix((K32[rows, cols], (rows, cols)), shape=K32.shape)


def save_reference_npz(path: str, fx, train: Dict[str, np.ndarray], validate: Dict[str, np.ndarray], variant: str = "steady_ns",
                       load_vector: Optional[np.ndarray] = None) -> str:
    """Writes fixture operators + samples in the reference's npz schema (dense float64 operators)."""
    idx_sol = np.empty(3, dtype=object)
    idx_sol[0], idx_sol[1], idx_sol[2] = [int(i) for i in fx.idx_u1], [int(i) for i in fx.idx_u2], [int(i) for i in fx.idx_p]
    pos = fx.pos if fx.pos is not None else np.zeros((fx.N, 2))
    out = dict(ne=fx.mesh.ne, ng=fx.N, p=pos, idx_sol=idx_sol, pos_u=pos[np.asarray(fx.idx_u1)], pos_p=pos[np.asarray(fx.idx_p)])
    dense = lambda K: np.asarray(K.todense(), dtype=np.float64)  # noqa: E731
    if variant == "steady_ns":
        out.update(A=dense(fx.A), B1=dense(fx.B1), B2=dense(fx.B2))
    elif variant == "time_dep":
        out.update(S=dense(fx.S), A=dense(fx.A), gfl=np.zeros((fx.N, 1)))
        if load_vector is not None:
            out["load_vector"] = np.asarray(load_vector, dtype=np.float64)
    else:
        out.update(matrix=dense(fx.A), gfl=np.zeros((fx.N, 1)))
    ren = {"coeff_f": "coeff_fs", "load_vec_f": "load_vectors"}
    for kind, data in (("train", train), ("validate", validate)):
        for k, v in data.items():
            out[f"{kind}_{ren.get(k, k)}"] = np.asarray(v)
        if variant == "steady_ns" and "forcing_term" not in data:
            out[f"{kind}_forcing_term"] = np.zeros((len(data["coeff_f"]), 2))  # sincos: placeholder zeros, as in the reference
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **out)
    return path


def load_reference_npz(path: str) -> Dict:
    """Reads an `assemble_fenics.py` npz: operators -> CSR (fp32 values), idx_sol kept as the object array the training
    scripts index (`i, j, _ = idx_sol`), sample arrays as float32 under the names the epoch loop uses."""
    z = np.load(path, allow_pickle=True)
    out: Dict = {"ne": int(z["ne"]), "N": int(z["ng"]), "idx_sol": z["idx_sol"], "p": z["p"]}
    for k in ("pos_u", "pos_p", "gfl", "load_vector"):
        if k in z.files:
            out[k] = z[k]
    for k in OPERATOR_KEYS:
        if k in z.files:
            out[k] = dense_to_csr(z[k])
    ren = {"coeff_fs": "coeff_f", "load_vectors": "load_vec_f"}
    for kind in ("train", "validate"):
        d = {}
        for k in SAMPLE_KEYS:
            key = f"{kind}_{k}"
            if key in z.files:
                d[ren.get(k, k)] = np.asarray(z[key], dtype=np.float32)
        out[kind] = d
    return out


def save_pkl(path: str, rows) -> str:
    """`save_obj` of create_data.py: pickled object ndarray [num_data, k]."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(np.array(rows, dtype=object), f, pickle.HIGHEST_PROTOCOL)
    return path


def load_pkl(path: str, variant: str = "steady_ns") -> Dict[str, np.ndarray]:
    """Rows `[f_value, coeff_f]` (steady NS) / `[coeff_u, f_value, coeff_f]` (linear Stokes) -> stacked float32 arrays."""
    with open(path, "rb") as f:
        rows = pickle.load(f)
    cols = ("f_value", "coeff_f") if variant == "steady_ns" else ("coeff_u", "f_value", "coeff_f")
    return {name: np.stack([np.asarray(r[i], dtype=np.float32) for r in rows]) for i, name in enumerate(cols)}
