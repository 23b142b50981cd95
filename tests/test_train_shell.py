"""The training shell around the hot path (feonet_navier_stokes_b200/train_FEONet.py): flag parsing and data
synthesis on the CPU, a short end-to-end training run on the GPU."""
import numpy as np
import pytest

from feonet_navier_stokes_b200 import train_FEONet as T
from oracle import feonet_oracle as orc


def test_flags_follow_the_reference_cli():
    args = T.build_parser().parse_args("--bc channel_flow --forcing_term sincos --train_file 1000N450 --val_file 1000N450 "
                                       "--model UNetWithHead --optimizer Adam --do_precond 1 --resol_in 64 --blocks 4 --ks 5 "
                                       "--filters 64 --epochs 10".split())
    assert T.parse_file_flag(args.train_file) == (1000, 450) and T.mesh_n_from_ne(450) == 15 and T.mesh_n_from_ne(72) == 6
    with pytest.raises(ValueError):
        T.mesh_n_from_ne(154)
    c = T.sample_coeff_f(50, 5)
    assert c.shape == (50, 6) and (c[:, :2] <= 1).all() and (c[:, 2:] <= np.pi).all() and (c >= 0).all()


@pytest.mark.parametrize("branch", [True, False])
def test_synthetic_ns_solutions_zero_the_reference_residual(branch):
    """The Newton solutions used as validation targets solve the algebraic system the reference's loss penalises
    (both sign branches, steady NS train_FEONet.py:324-330): the oracle's residual loss vanishes on them."""
    fx, data = T.synthesize("steady_ns", 5, "channel_flow", 3, 5, branch)
    U = np.zeros((3, fx.N))
    U[:, fx.idx_u1], U[:, fx.idx_u2], U[:, fx.idx_p] = data["fenics_u1"], data["fenics_u2"], data["fenics_p"]
    loss, _, _ = orc.ns_loss_and_grad(U, data["load_vec_f"], fx.A, fx.B1, fx.B2, fx.idx_u1, fx.idx_u2, branch, dtype=np.float64)
    scale = float((data["load_vec_f"] ** 2).sum())
    assert loss < 1e-18 * scale


def test_synthetic_stokes_solutions():
    fx, data = T.synthesize("stokes_square", 4, "channel_flow", 2, 5, False)
    U = np.zeros((2, fx.N))
    U[:, fx.idx_u1], U[:, fx.idx_u2], U[:, fx.idx_p] = data["fenics_u1"], data["fenics_u2"], data["fenics_p"]
    assert np.abs(U @ fx.A.T.toarray() - data["load_vec_f"]).max() < 1e-10


def test_synthetic_time_dep_trajectories_zero_the_reference_residual():
    """The implicit-Euler trajectories the time-dependent shell validates against (create_data.py:75-91) make the reference's
    sequence residual vanish (FEONet_time_dep_Stokes/train_FEONet.py:343-362, :398-400)."""
    fx, data = T.synthesize_time_dep(4, 3, 5, 0.1, 5)
    U = data["coeffs_u"]
    assert U.shape == (3, 6, fx.N)
    u0 = orc.assemble_u_init(data["init_x"], data["init_y"], fx.idx_u1, fx.idx_u2, fx.N, dtype=np.float64)
    assert np.array_equal(u0, U[:, 0])
    loss, _, _ = orc.seq_loss_and_grad(U[:, 1:], data["load_vec_f"], fx.S, fx.A, None, 0.1, u0, False, dtype=np.float64)
    assert loss < 1e-20 * float((U ** 2).sum())
    args = T.build_parser().parse_args("--variant time_dep --model RNN --seq_len 10 --rnn_type gru --dt 0.01 --train_file 8N32 --val_file 4N32".split())
    assert args.model == "RNN" and args.seq_len == 10 and args.dt == 0.01


def test_batches_are_the_same_on_every_rank():
    """Each rank takes its shard of every GLOBAL batch: the same number of steps everywhere (every step issues collectives), no
    empty shard (a tail smaller than the world size joins the batch before it)."""
    import torch

    class Stub:
        batches = T.Trainer.batches

    from feonet_navier_stokes_b200 import parallel

    data = {"coeff_f": torch.arange(10).float().unsqueeze(1)}
    seen = []
    for rank in range(4):
        st = Stub()
        st.world, st.rank, st.parallel = 4, rank, parallel
        seen.append([b["coeff_f"].flatten().tolist() for b in st.batches(data, 3, shard=True)])
    assert len({len(x) for x in seen}) == 1 and all(len(b) > 0 for x in seen for b in x)
    flat = sorted(v for x in seen for b in x for v in b)
    assert flat == list(map(float, range(10)))  # every sample exactly once per epoch


@pytest.mark.gpu
@pytest.mark.parametrize("variant,do_precond", [("steady_ns", 1), ("stokes_square", 0), ("stokes_square", 1), ("time_dep", 0)])
def test_training_run_reduces_the_loss(tmp_path, variant, do_precond):
    import torch

    argv = ["--variant", variant, "--train_file", "32N32", "--val_file", "8N32", "--model", "RNN" if variant == "time_dep" else "FCNN",
            "--optimizer", "Adam", "--hidden_dim", "64", "--seq_len", "4",
            "--do_precond", str(do_precond), "--epochs", "60", "--log_every", "20", "--lr", "3e-3", "--out", str(tmp_path),
            "--spai_steps", "50"]
    tr = T.Trainer(dict(T.build_parser().parse_args(argv).__dict__), device=torch.device("cuda"))
    out = tr.fit()
    assert len(tr.losses) == 3 and tr.losses[-1] < tr.losses[0] and np.isfinite(out["val_all"])
    import os
    assert os.path.exists(os.path.join(tr.folder, "model.pt")) and os.path.exists(os.path.join(tr.folder, "model.pth"))
    ck = torch.load(os.path.join(tr.folder, "model.pt"), map_location="cpu")
    assert set(ck) == {"model_state_dict", "losses", "train_rel_L2_errors", "test_rel_L2_errors"}


@pytest.mark.gpu
def test_training_from_a_reference_npz(tmp_path):
    """The shell trains from an `assemble_fenics.py`-schema npz (dense float64 operators, object idx_sol)."""
    import torch

    from feonet_navier_stokes_b200 import data_io as D

    fx, train = T.synthesize("steady_ns", 4, "channel_flow", 16, 5, True)
    _, val = T.synthesize("steady_ns", 4, "channel_flow", 4, 10, True)
    path = D.save_reference_npz(str(tmp_path / D.npz_name(fx.mesh.ne, "channel_flow", "sincos")), fx, train, val, "steady_ns")
    argv = ["--variant", "steady_ns", "--train_file", "16N32", "--val_file", "4N32", "--model", "FCNN", "--optimizer", "Adam", "--do_precond", "1",
            "--epochs", "40", "--log_every", "20", "--lr", "3e-3", "--out", str(tmp_path), "--npz", path]
    tr = T.Trainer(dict(T.build_parser().parse_args(argv).__dict__), device=torch.device("cuda"))
    tr.fit()
    assert len(tr.losses) == 2 and tr.losses[-1] < tr.losses[0]
