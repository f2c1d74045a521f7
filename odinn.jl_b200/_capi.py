"""ctypes binding of ``libodinn_b200.so`` (the C ABI declared in ``include/odinn_b200.h``).

The product path has NO CPU fallback: if the CUDA library is missing or no device is present,
loading / creating an ensemble raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ODINN_B200_LIB") or os.path.join(_HERE, "lib", "libodinn_b200.so")  # env: dev override

F32, F64 = 0, 1

FIELD_B, FIELD_H, FIELD_DH, FIELD_LAMBDA, FIELD_VJP_H, FIELD_A, FIELD_VJP_A, FIELD_H0 = range(8)
EULER, SSPRK3, BS3, RDPK3SP35 = 0, 1, 2, 3
LAW_U, LAW_Y = 1, 2
ACT = {"identity": 0, "softplus": 1, "sigmoid": 2, "tanh": 3, "relu": 4}


class Phys(C.Structure):
    """``odinn_phys``: params.physical + iceflow-cache scalars (test/params_construction.jl:24-34)."""

    _fields_ = [(k, C.c_double) for k in ("rho", "g", "eta0", "n", "p", "q", "C", "minA", "maxA")]

    def __init__(self, rho=900.0, g=9.81, eta0=1.0, n=3.0, p=3.0, q=0.0, C=0.0, minA=8.5e-20, maxA=8e-17):
        super().__init__(rho, g, eta0, n, p, q, C, minA, maxA)


class OdinnError(RuntimeError):
    pass


_lib = None

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol of include/odinn_b200.h
SIGNATURES = {
    "odinn_ensemble_create": (_i, [_i, _i, _i, _ip, _ip, _dp, _dp, C.POINTER(Phys), C.POINTER(_vp)]),
    "odinn_ensemble_destroy": (None, [_vp]),
    "odinn_last_error": (C.c_char_p, [_vp]),
    "odinn_n_glaciers": (_i, [_vp]),
    "odinn_dtype_of": (_i, [_vp]),
    "odinn_launch_count": (C.c_longlong, [_vp]),
    "odinn_synchronize": (_i, [_vp]),
    "odinn_stream": (_vp, [_vp]),
    "odinn_upload": (_i, [_vp, _i, _i, _vp, _i]),
    "odinn_download": (_i, [_vp, _i, _i, _vp, _i]),
    "odinn_set_A_scalar": (_i, [_vp, _i, _d]),
    "odinn_set_A_mode": (_i, [_vp, _i]),
    "odinn_set_temperature": (_i, [_vp, _i, _d]),
    "odinn_solve_forward": (_i, [_vp, _i, _i, _dp, _i]),
    "odinn_solve_forward_adaptive": (_i, [_vp, _i, _i, _dp, _d, _d, _d, _i, _ip, _ip]),
    "odinn_get_snapshot": (_i, [_vp, _i, _i, _vp, _i]),
    "odinn_set_snapshot": (_i, [_vp, _i, _i, _i, _vp, _i]),
    "odinn_set_reference": (_i, [_vp, _i, _i, _i, _vp, _vp, _i]),
    "odinn_loss": (_i, [_vp, _dp, _i, _dp]),
    "odinn_grad_discrete": (_i, [_vp, _dp, _i, _dp, _dp]),
    "odinn_set_mass_balance": (_i, [_vp, _i, _ip, _dp]),
    "odinn_get_mass_balance": (_i, [_vp, _i, _i, _vp, _i]),
    "odinn_surface_velocity": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _d]),
    "odinn_sia2d_vjp_surface_V": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _dp, _d]),
    "odinn_set_velocity_reference": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "odinn_set_loss_weights": (_i, [_vp, _i, _dp, _dp, _i]),
    "odinn_grad_continuous": (_i, [_vp, _dp, _i, _i, _dp, _dp, _i, _i, _i, _dp, _dp]),
    "odinn_grad_continuous_adaptive": (_i, [_vp, _dp, _i, _i, _dp, _dp, _i, _d, _d, _d, _i, _dp, _dp, _ip]),
    "odinn_set_velocity_quadrature": (_i, [_vp, _d, _i]),
    "odinn_law_A_nn_apply": (_i, [_vp, _i, _ip, _ip, _dp, _i, _dp]),
    "odinn_law_A_nn_pullback": (_i, [_vp, _dp, _dp, _i]),
    "odinn_set_phys": (_i, [_vp, C.POINTER(Phys)]),
    "odinn_sia2d_rhs": (_i, [_vp, _i, _vp, _i, _vp, _i, _d]),
    "odinn_sia2d_vjp_H": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _d]),
    "odinn_sia2d_vjp_theta": (_i, [_vp, _i, _vp, _i, _vp, _i, _dp, _d]),
    "odinn_sia2d_vjp_H_continuous": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _d]),
    "odinn_sia2d_vjp_theta_continuous": (_i, [_vp, _i, _vp, _i, _vp, _i, _dp, _d]),
    "odinn_law_cell_nn_set": (_i, [_vp, _i, _i, _ip, _ip, _dp, _i, _dp, _d, _d, _d]),
    "odinn_law_cell_clear": (_i, [_vp]),
    "odinn_law_cell_interp_set": (_i, [_vp, _i, _dp, _i, _dp]),
    "odinn_sia2d_vjp_theta_cell": (_i, [_vp, _i, _vp, _i, _vp, _i, _dp, _i, _d]),
    "odinn_law_cell_grad": (_i, [_vp, _dp, _i]),
    "odinn_rhs_resident": (_i, [_vp]),
    "odinn_vjp_resident": (_i, [_vp, _i, _dp]),
    "odinn_fwd_adj_batch_host": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _dp]),
    "odinn_set_batch_chunk": (_i, [_vp, C.c_longlong]),
    "odinn_set_cluster_mode": (_i, [_vp, _i]),
    "odinn_host_register": (_i, [_vp, _vp, C.c_size_t]),
    "odinn_host_unregister": (_i, [_vp, _vp]),
}


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OdinnError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc):
    if rc != 0:
        msg = load().odinn_last_error(handle)
        raise OdinnError(f"libodinn_b200 error {rc}: {msg.decode() if msg else '?'}")
