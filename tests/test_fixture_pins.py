"""CPU: the offline P2-P1 assembler reproduces the reference notebooks' known-answer values
(SURVEY.md section 4)."""
import numpy as np

from feonet_navier_stokes_b200.fixtures import assemble_operators, config_operators, structured_mesh


def test_cond_pin_ne72():
    # FEONet_Stokes_square/test.ipynb#c3: cond(A_ne72, channel_flow) = 167.32636402645198
    for ordering in ("blocked", "interleaved"):
        op = config_operators("stokes_square", 6, ordering=ordering)
        assert op.N == 387 and op.mesh.ne == 72
        assert abs(np.linalg.cond(op.A.toarray()) - 167.32636402645198) < 1e-9


def test_minmax_pins():
    # compare_ordering_nonlinear.ipynb#c13,c15: A in [-0.13333, 1.0]; B1 (no BC) = -/+0.0066667 at n=40
    op = assemble_operators(structured_mesh(40), with_convection=True, keep_nobc=True)
    assert abs(op.A.min() + 0.4 / 3) < 1e-12 and op.A.max() == 1.0
    assert abs(op.B1_nobc.max() - 0.2 / 30) < 1e-12 and abs(op.B1_nobc.min() + 0.2 / 30) < 1e-12


def test_sizes():
    for name, n, N in (("steady_ns", 15, 2178), ("time_dep", 10, 1003)):
        assert config_operators(name, n).N == N
    op = config_operators("steady_ns", 15)
    bc = op.bc_dofs
    # bc.apply on every matrix: identity rows in A, B1, B2 (quirk 3)
    for K in (op.A, op.B1, op.B2):
        rows = K[bc].toarray()
        assert np.array_equal(rows, np.eye(op.N)[bc])
