"""CPU: the re-stated model zoo keeps the reference's public surface: class names, constructor
arguments, output shape [B,1,d_out], state_dict keys/shapes (checked against the reference tree when it
is present), and the dof-major head computes the same values as nn.Linear."""
import importlib.util
import os

import pytest
import torch

from feonet_navier_stokes_b200 import network as net

REF = "/root/reference/FEONet_steady_Navier-Stokes/network.py"


def _ref():
    spec = importlib.util.spec_from_file_location("ref_network", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


CASES = [
    ("Net2D", dict(resol_in=8, d_in=2, filters=4, d_out=23, kernel_size=5, padding=2, blocks=2), (3, 2, 8, 8)),
    ("NetA", dict(d_in=2, filters=4, d_out=16, blocks=1), (3, 2, 16)),
    ("FCNN", dict(resol_in=6, output_dim=31, hidden_dims=[16, 32]), (3, 6)),
    ("UNetWithHead", dict(resol_in=8, d_out=19, base_ch=4, latent_ch=3, head_filters=4, head_blocks=2,
                          head_kernel_size=5, head_padding=2), (3, 2, 8, 8)),
]


@pytest.mark.parametrize("name,kw,shape", CASES)
def test_output_shape_and_dof_major_head(name, kw, shape):
    torch.manual_seed(0)
    m = getattr(net, name)(**kw).eval()
    x = torch.randn(*shape)
    y = m(x)
    d_out = kw.get("d_out", kw.get("output_dim"))
    assert tuple(y.shape) == ((3, d_out) if name == "FCNN" else (3, 1, d_out))
    m2 = getattr(net, name)(**kw, dof_major_head=True).eval()
    m2.load_state_dict(m.state_dict())
    y2 = m2(x)
    assert y2.shape == y.shape and torch.allclose(y2, y, rtol=1e-5, atol=1e-6)
    flat = y2.reshape(3, d_out)
    assert flat.stride(0) == 1 and flat.stride(1) == 4  # dof-major, ld = ceil4(B)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,kw,shape", CASES)
def test_state_dict_and_values_match_reference(name, kw, shape):
    torch.manual_seed(0)
    ref = getattr(_ref(), name)(**kw).eval()
    ours = getattr(net, name)(**kw).eval()
    sd = ref.state_dict()
    assert list(sd.keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(sd)
    x = torch.randn(*shape)
    assert torch.allclose(ours(x), ref(x), rtol=1e-6, atol=1e-7)


REF_TD = "/root/reference/FEONet_time_dep_Stokes/network.py"


@pytest.mark.skipif(not os.path.exists(REF_TD), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("rnn_type,layers", [("gru", 1), ("lstm", 2)])
def test_sequence_rnn_matches_reference(rnn_type, layers):
    """VectorToSequenceRNN of the time-dependent variant (FEONet_time_dep_Stokes/network.py:342-399): same state_dict, same values."""
    spec = importlib.util.spec_from_file_location("ref_network_td", REF_TD)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    torch.manual_seed(1)
    ref = m.VectorToSequenceRNN(input_dim=37, hidden_dim=16, output_dim=37, rnn_type=rnn_type, num_layers=layers).eval()
    ours = net.VectorToSequenceRNN(input_dim=37, hidden_dim=16, output_dim=37, rnn_type=rnn_type, num_layers=layers).eval()
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(4, 37)
    assert torch.equal(ours(x, seq_len=5), ref(x, seq_len=5))


def _load_state(model, g):
    sd = {k[len("state__"):].replace("__", "."): torch.tensor(g[k]) for k in g.files if k.startswith("state__")}
    model.load_state_dict(sd)
    return model


def test_networks_reproduce_the_trainstep_goldens():
    """tests/golden/trainstep_*.npz hold one closure evaluation of the reference's OWN networks (oracle/make_golden.py): the
    restated FCNN / VectorToSequenceRNN, loaded with the stored weights, give the stored network outputs on the CPU."""
    import numpy as np
    from conftest import GOLDEN_DIR

    g = np.load(os.path.join(GOLDEN_DIR, "trainstep_ns_precond_n4.npz"))
    N = g["A"].shape[0]
    fc = _load_state(net.FCNN(6, N, [int(h) for h in g["hidden"]]).eval(), g)
    pred = fc(torch.tensor(g["coeff_f"])).unsqueeze(1)
    assert torch.allclose(pred, torch.tensor(g["u_pred"]), rtol=1e-6, atol=1e-7)  # PRECOND = I: u_pred is the prediction
    g = np.load(os.path.join(GOLDEN_DIR, "trainstep_timedep_n4.npz"))
    N = g["A"].shape[0]
    rnn = _load_state(net.VectorToSequenceRNN(input_dim=N, hidden_dim=int(g["hidden"][0]), output_dim=N).eval(), g)
    u0 = torch.zeros(g["init_x"].shape[0], N)
    u0[:, torch.tensor(g["idx_u1"])] = torch.tensor(g["init_x"])
    u0[:, torch.tensor(g["idx_u2"])] = torch.tensor(g["init_y"])
    assert torch.allclose(rnn(u0, seq_len=int(g["T"])), torch.tensor(g["u_pred"]), rtol=1e-5, atol=1e-6)
