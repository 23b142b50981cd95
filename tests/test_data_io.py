"""On-disk formats of the reference pipeline (feonet_navier_stokes_b200/data_io.py): npz schema round trip and the
dense -> CSR import rule (an entry survives iff it is non-zero after the fp32 cast)."""
import numpy as np

from feonet_navier_stokes_b200 import data_io as D
from feonet_navier_stokes_b200 import train_FEONet as T


def test_npz_names_follow_the_reference():
    assert D.npz_name(72, "channel_flow") == "P2x1_ne72_stokes_channel_flow_BC.npz"
    assert D.npz_name(450, "channel_flow", "sincos") == "P2x1_ne450_stokes_channel_flow_BC_sincos.npz"
    assert D.npz_name(200, "channel_flow", dt=0.01) == "P2x1_ne200_stokes_channel_flow_BC_dt_0_01.npz"


def test_dense_import_threshold_is_zero_after_the_fp32_cast():
    K = np.array([[1.0, 1e-60, 0.0], [1e-30, 0.0, -2.5], [0.0, 3.0, 1e-46]])
    C = D.dense_to_csr(K)
    assert C.dtype == np.float32 and C.nnz == 4  # 1e-60 and 1e-46 vanish in fp32, 1e-30 stays
    assert np.array_equal(C.toarray(), K.astype(np.float32))


def test_npz_round_trip_steady_ns(tmp_path):
    fx, train = T.synthesize("steady_ns", 3, "channel_flow", 4, 5, True)
    _, val = T.synthesize("steady_ns", 3, "channel_flow", 2, 10, True)
    path = D.save_reference_npz(str(tmp_path / D.npz_name(fx.mesh.ne, "channel_flow", "sincos")), fx, train, val, "steady_ns")
    z = np.load(path, allow_pickle=True)
    assert {"ne", "ng", "p", "idx_sol", "pos_u", "pos_p", "A", "B1", "B2", "train_coeff_fs", "train_forcing_term", "train_load_vectors",
            "train_fenics_u1", "train_fenics_u2", "train_fenics_p", "validate_coeff_fs", "validate_load_vectors"} <= set(z.files)
    assert z["A"].dtype == np.float64 and z["A"].shape == (fx.N, fx.N) and z["idx_sol"].dtype == object
    i, j, _ = z["idx_sol"]  # how the training scripts unpack it (steady NS train_FEONet.py:305)
    assert isinstance(i, list) and i == [int(k) for k in fx.idx_u1] and j == [int(k) for k in fx.idx_u2]
    back = D.load_reference_npz(path)
    for name in ("A", "B1", "B2"):
        ref = getattr(fx, name).astype(np.float32)
        ref.eliminate_zeros()
        assert (abs(back[name] - ref)).nnz == 0 and back[name].nnz == ref.nnz
    assert back["N"] == fx.N and np.allclose(back["train"]["load_vec_f"], train["load_vec_f"].astype(np.float32))
    assert back["validate"]["coeff_f"].shape == (2, 6)


def test_pkl_round_trip(tmp_path):
    rows = [[np.arange(8.0).reshape(2, 4) + k, np.full(6, k, dtype=float)] for k in range(3)]
    path = D.save_pkl(str(tmp_path / "data_ordered" / "channel_flow" / "sincos" / "train" / "3N8.pkl"), rows)
    back = D.load_pkl(path, "steady_ns")
    assert back["f_value"].shape == (3, 2, 4) and back["coeff_f"].shape == (3, 6) and back["coeff_f"][2, 0] == 2.0
