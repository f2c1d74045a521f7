"""Host-side mirror of the reference's operator interface for the SIA2D hot path.

Same names, argument order and meaning as ODINN.jl so that parity tests read like the
reference's own (test/SIA2D_adjoint.jl:115-137):

    SIA2D!(dH, H, simulation, t, θ)                          -> SIA2D_(dH, H, simulation, t, θ)
        [Huginn.SIA2D!, called at src/simulations/inversions/inversion_utils.jl:691-699]
    VJP_λ_∂SIA∂H(VJPMode, λ, H, θ, simulation, t)            -> VJP_λ_dSIAdH(...)
        [src/inverse/SIA2D/VJPs.jl:2-28]
    VJP_λ_∂SIA∂θ(VJPMode, λ, H, θ, dH_H, simulation, t)      -> VJP_λ_dSIAdθ(...)
        [src/inverse/SIA2D/VJPs.jl:30-59]

(``!`` and ``∂`` are not valid in Python identifiers; ``λ`` and ``θ`` are.)
Every operator runs on the GPU through ``libodinn_b200.so``; there is no CPU path here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from ._capi import Phys
from .ensemble import Ensemble


@dataclass
class Glacier2D:
    """The fields of Sleipnir.Glacier2D read on the path (src/inverse/SIA2D/adjoint.jl:47-49)."""

    B: np.ndarray
    Δx: float
    Δy: float
    H0: Optional[np.ndarray] = None
    rgi_id: str = ""

    @property
    def nx(self):
        return self.B.shape[0]

    @property
    def ny(self):
        return self.B.shape[1]


class AbstractVJPMethod:
    """src/inverse/VJPTypes.jl:10."""


class DiscreteVJP(AbstractVJPMethod):
    """src/inverse/VJPTypes.jl:29 -- the hand-derived discrete adjoint, here the A1/A2 CUDA kernels."""


class ContinuousVJP(AbstractVJPMethod):
    """src/inverse/VJPTypes.jl:57 -- the continuous adjoint (adjoint.jl:442-662), here sia2d_cont.cuh."""


class B200VJP(DiscreteVJP):
    """The plug-in flavour a maintainer adds to ODINN (``B200VJP <: AbstractVJPMethod``); numerically it
    IS the DiscreteVJP, executed by libodinn_b200."""


class Cache:
    """simulation.cache.iceflow: only ``glacier_idx`` (0-based here) is read by the operators."""

    def __init__(self):
        self.glacier_idx = 0


class Simulation:
    """What ``simulation`` carries across the boundary: glaciers, physical parameters, the A value of
    every glacier (cache.iceflow.A.value) and the device ensemble holding B.

    A: scalar per glacier, a list of scalars, or a list of dual-grid ``(nx-1, ny-1)`` matrices (gridded A,
    src/laws/Laws.jl:430-454)."""

    def __init__(self, glaciers: Sequence[Glacier2D], phys: Optional[Phys] = None, A=2.21e-18, dtype: str = "f64",
                 device: int = 0):
        self.glaciers: List[Glacier2D] = list(glaciers)
        self.phys = phys if phys is not None else Phys()
        self.dtype = dtype
        self.cache = Cache()
        self.ensemble = Ensemble([g.nx for g in glaciers], [g.ny for g in glaciers], [g.Δx for g in glaciers],
                                 [g.Δy for g in glaciers], self.phys, dtype, device)
        from . import _capi

        for k, g in enumerate(self.glaciers):
            self.ensemble.upload(k, _capi.FIELD_B, g.B)
        self.set_A(A)

    def set_A(self, A):
        if np.isscalar(A):
            A = [A] * len(self.glaciers)
        self.A = list(A)
        gridded = any(np.ndim(a) == 2 for a in self.A)
        self.ensemble.set_A_mode(gridded)
        for k, a in enumerate(self.A):
            if gridded:
                g = self.glaciers[k]
                af = np.broadcast_to(np.asarray(a, dtype=np.float64), (g.nx - 1, g.ny - 1))
                self.ensemble.set_A_field(k, af)
            else:
                self.ensemble.set_A_scalar(k, float(a))

    def close(self):
        self.ensemble.close()


def SIA2D_(dH, H, simulation: Simulation, t, θ=None):
    """In-place ODE right-hand side, mirror of ``Huginn.SIA2D!(dH, H, simulation, t, θ)``.
    Writes every entry of ``dH`` (border zeros included) and never writes ``H``."""
    g = simulation.cache.glacier_idx
    out = simulation.ensemble.sia2d_rhs(g, H, t)
    dH[...] = out
    return None


def VJP_λ_dSIAdH(VJPMode: AbstractVJPMethod, λ, H, θ, simulation: Simulation, t):
    """Mirror of ``VJP_λ_∂SIA∂H`` (src/inverse/SIA2D/VJPs.jl:2-5): returns ``(λ_∂f∂H, nothing)``."""
    if not isinstance(VJPMode, (DiscreteVJP, ContinuousVJP)):
        raise NotImplementedError(f"VJP flavour {type(VJPMode).__name__} is not provided by libodinn_b200")
    g = simulation.cache.glacier_idx
    return simulation.ensemble.sia2d_vjp_H(g, λ, H, t, continuous=isinstance(VJPMode, ContinuousVJP)), None


def VJP_λ_dSIAdθ(VJPMode: AbstractVJPMethod, λ, H, θ, dH_H, simulation: Simulation, t, vjp_θ=None):
    """Mirror of ``VJP_λ_∂SIA∂θ`` (src/inverse/SIA2D/VJPs.jl:30-33).

    For a glacier-wide law the dense ``∂D∂θ`` tensor of the reference factorises as
    ``∂A_spatial ⊗ vjp_θ`` (src/models/target/target_A.jl:85-87), so the kernel reduces the scalar
    ``Σ ∂A_spatial ∘ D†`` and the result is ``vjp_θ * scalar``.  ``vjp_θ`` is ``cache.iceflow.A.vjp_θ``
    (the law pullback ∂A/∂θ); when omitted the bare scalar is returned."""
    if not isinstance(VJPMode, (DiscreteVJP, ContinuousVJP)):
        raise NotImplementedError(f"VJP flavour {type(VJPMode).__name__} is not provided by libodinn_b200")
    g = simulation.cache.glacier_idx
    S = simulation.ensemble.sia2d_vjp_theta(g, λ, H, t, continuous=isinstance(VJPMode, ContinuousVJP))
    if vjp_θ is None:
        return S
    return np.asarray(vjp_θ, dtype=np.float64) * S
