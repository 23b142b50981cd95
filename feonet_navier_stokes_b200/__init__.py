"""feonet_navier_stokes_b200 -- B200-native FEM residual loss (forward + backward) for FEONet.

Drop-in for the hot path of haltmayermarc/FEONet_Navier_Stokes' `train_FEONet.py` scripts:
`weak_form` / `closure` / `weak_form_sequence` / `assemble_u_init`.  The loss path is hand-written
sm_100a CUDA behind a C ABI (include/feonet_b200.h); there is no CPU fallback.
"""
from ._lib import FeoError, load_library  # noqa: F401
from .functional import (DenseFn, DenseResidualLossFn, ResidualLossFn, SeqResidualLossFn, SpmmFn,  # noqa: F401
                         dof_major_empty, dof_major_zeros, is_dof_major, precond_output, to_dof_major_tensor)
from .graphs import GraphedLossGrad, GraphedTrainStep, make_capturable_optimizer  # noqa: F401
from .host_io import HostBatchPipeline  # noqa: F401
from .operator import FEOperator  # noqa: F401
from .precond import spai_device  # noqa: F401
from .train_api import (LinearStokes, SteadyNavierStokes, TimeDependentStokes, rel_L2_error,  # noqa: F401
                        sincos_forcing_grid)

__version__ = "0.1.0"
