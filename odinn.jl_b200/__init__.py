"""odinn.jl_b200 -- B200-native (sm_100a) SIA2D hot path of ODINN.jl behind a C ABI.

The directory name carries a dot, so the package is imported as ``odinn_b200`` through the
root-level shim ``odinn_b200.py``.  Contents: ``csrc/`` (CUDA kernels + C ABI), ``lib/``
(the built ``libodinn_b200.so``), ``_capi`` (ctypes binding), ``ensemble`` (handle wrapper),
``sia2d`` (mirror of the reference's operator interface), ``api`` (mirror of the user-facing objects and of the
two optimiser callables), ``parallel`` (ensemble sharding + the loss/gradient all-reduce).
"""
from ._capi import F32, F64, LIB_PATH, OdinnError, Phys, load  # noqa: F401
from .ensemble import Ensemble  # noqa: F401
from . import parallel  # noqa: F401
from .api import (  # noqa: F401
    ContinuousAdjoint,
    DiscreteAdjoint,
    FunctionalInversion,
    Inversion,
    LawA,
    Model,
    NeuralNetwork,
    Parameters,
    Prediction,
    SIA2D_grad_,
    SIA2Dmodel,
    SolverParameters,
    define_callback_steps,
    create_interpolation,
    is_in_glacier,
    loss_iceflow_transient,
    run_,
    train_UDE_,
)
from .sia2d import (  # noqa: F401
    AbstractVJPMethod,
    B200VJP,
    ContinuousVJP,
    DiscreteVJP,
    Glacier2D,
    SIA2D_,
    Simulation,
    VJP_λ_dSIAdH,
    VJP_λ_dSIAdθ,
)
