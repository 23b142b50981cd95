"""Test/bench fixtures: offline P2-P1 operators standing in for the FEniCS-generated npz files."""
from .taylor_hood import (  # noqa: F401
    P2P1Mesh, StokesOperators, assemble_operators, config_operators, dof_numbering, spai,
    square_with_hole_mesh, structured_mesh, unstructured_mesh,
)
