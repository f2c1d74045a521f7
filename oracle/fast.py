"""The NumPy oracle's time / reverse loops with the C restatement plugged in as F1 / A1 / A2 --
TEST INFRASTRUCTURE ONLY (PARITY UNPINNED BY STORED NUMBERS, see oracle/__init__.py).

The loops themselves (oracle.sia2d_numpy.solve_forward, loss_and_grad_discrete, ...) are not restated here:
`use_c_kernels()` swaps the three per-call operators those loops look up by name for the fused C versions
(oracle/sia2d_c.c, pinned to the NumPy operators at 1e-13 by tests/test_oracle_c.py and tests/test_oracle_fast.py),
so that BASELINE-size cases (500 x 500 x 61 snapshots, 64-glacier ensembles) finish in seconds on the GPU box's
host cores.  Glacier-wide `TargetA` laws only ("const", "nn", "scalar"); anything else falls through to NumPy.
"""
from __future__ import annotations

import contextlib

import numpy as np

from . import sia2d_c as oc
from . import sia2d_numpy as o


def _is_c_target(target):
    return isinstance(target, o.TargetA) and target.kind in ("const", "nn", "scalar")


def _A_of(target, theta):
    target.apply_laws(None, None, theta)  # adjoint.jl:75-76 (glacier-wide laws do not look at the node fields)
    return target.A


def SIA2D_c(H, glacier, target, theta=None):
    if not _is_c_target(target):
        return _np_SIA2D(H, glacier, target, theta)
    return oc.rhs(H, glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta))


def VJP_dSIA_dH_discrete_c(lam, H, glacier, target, theta=None):
    if not _is_c_target(target):
        return _np_VJP_H(lam, H, glacier, target, theta)
    return oc.vjp(lam, H, glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta))[0]


def VJP_dSIA_dtheta_discrete_c(lam, H, glacier, target, theta=None, dense=False):
    """adjoint.jl:250 for a glacier-wide law: (dA/dtheta) * sum_ij dA_spatial * D_adj (target_A.jl:85-87)."""
    if not _is_c_target(target) or dense or np.ndim(_A_of(target, theta)) == 2:
        return _np_VJP_theta(lam, H, glacier, target, theta, dense=dense)
    if target.vjp_theta is None:
        target.precompute_vjp(theta)
    S = oc.vjp(lam, H, glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta), want_H=False)[1]
    if target.kind == "const":
        return np.zeros(0)
    return np.atleast_1d(target.vjp_theta).reshape(-1) * S


_np_SIA2D, _np_VJP_H, _np_VJP_theta = o.SIA2D, o.VJP_dSIA_dH_discrete, o.VJP_dSIA_dtheta_discrete


@contextlib.contextmanager
def use_c_kernels(all_cores=True):
    """Inside the block every loop of oracle.sia2d_numpy evaluates F1 / A1 / A2 with the C restatement."""
    if all_cores:
        oc.use_all_cores()
    o.SIA2D, o.VJP_dSIA_dH_discrete, o.VJP_dSIA_dtheta_discrete = SIA2D_c, VJP_dSIA_dH_discrete_c, VJP_dSIA_dtheta_discrete_c
    try:
        yield
    finally:
        o.SIA2D, o.VJP_dSIA_dH_discrete, o.VJP_dSIA_dtheta_discrete = _np_SIA2D, _np_VJP_H, _np_VJP_theta


def S_of(lam, H, glacier, target, theta=None):
    """The scalar reduction of A2 (node_reduction_S) from the C kernel."""
    return oc.vjp(lam, H, glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta), want_H=False)[1]


def solve_forward_fixed(H0, glacier, target, theta, t, method="ssprk3", nsub=8):
    """oracle.sia2d_numpy.solve_forward (fixed-step schemes, no mass balance) as ONE C loop (oracle/sia2d_c_impl.h,
    sia2d_solve_fixed) -- pinned to the NumPy loop by tests/test_oracle_fast.py."""
    assert _is_c_target(target)
    return oc.solve_fixed(H0, glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta), t, nsub=nsub, method=method)


def loss_and_grad_discrete(theta, glacier, target, t, Hs, H_ref, distance=3):
    """oracle.sia2d_numpy.loss_and_grad_discrete as ONE C loop (sia2d_grad_discrete): (loss, d(theta), lambda(t_0))."""
    assert _is_c_target(target)
    target.precompute_vjp(theta)
    masks = [o.is_in_glacier(h, distance) for h in H_ref]
    ell, Ssum, lam0 = oc.grad_discrete(glacier.B, glacier.dx, glacier.dy, target.ph, _A_of(target, theta), t, Hs, H_ref, masks)
    dth = np.atleast_1d(target.vjp_theta).reshape(-1) * Ssum if target.kind != "const" else np.zeros(0)
    return ell, dth, lam0
