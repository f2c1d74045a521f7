"""Host-side mirror of ODINN's user-facing API for the SIA2D path: the objects a user builds
(``NeuralNetwork``, ``LawA``, ``SIA2Dmodel``, ``Model``, ``Parameters``, ``Prediction``, ``Inversion`` /
``FunctionalInversion``) and the two callables the optimiser sees (``loss_iceflow_transient``,
``SIA2D_grad_``), all executing on the GPU through ``libodinn_b200.so``.

Reference (ODINN.jl v1.1.0):
    NeuralNetwork / build_default_NN   src/models/trainable_components/{NeuralNetwork.jl:18-73, ML_utils.jl:23-65}
    LawA                               src/laws/Laws.jl:323-460
    Model                              src/models/trainable_components/Model.jl:61-127
    Inversion                          src/simulations/inversions/Inversion.jl:16-62
    run!, train_UDE!                   src/simulations/inversions/inversion_utils.jl:21-88, 112-238
    loss_iceflow_transient             src/simulations/inversions/inversion_utils.jl:287-296
    SIA2D_grad!                        src/inverse/SIA2D/gradient.jl:6-31
    Prediction / run!(::Prediction)    Huginn.jl [not in tree]; wiring restated at inversion_utils.jl:472-539

The optimiser itself (Optimization.jl: Adam / BFGS) stays on the host, exactly like in the reference; a plain
Adam loop is provided so that the twin experiment of test/inversion_test.jl can be run end to end.
Multi-GPU: glaciers are sharded over the ranks of a torchrun job (``parallel.shard_glaciers``) and the
per-iteration ``[loss; dθ]`` is summed with one NCCL all-reduce.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _capi, parallel
from ._capi import Phys
from .ensemble import Ensemble
from .sia2d import Glacier2D


# ------------------------------------------------------------------------------------------------
# trainable components
# ------------------------------------------------------------------------------------------------
class NeuralNetwork:
    """Dense chain + flat parameter vector θ (Lux layout: per layer vec(W[out×in]) column-major, then b[out]).
    ``NeuralNetwork(params)`` of the reference builds 1→3→10→3→1 softplus×3 + sigmoid (ML_utils.jl:31-36)."""

    def __init__(self, widths: Sequence[int] = (1, 3, 10, 3, 1),
                 acts: Sequence[str] = ("softplus", "softplus", "softplus", "sigmoid"), θ=None, seed: int = 666):
        self.widths = [int(w) for w in widths]
        self.acts = list(acts)
        assert len(self.acts) == len(self.widths) - 1
        self.n_params = sum(o * i + o for i, o in zip(self.widths[:-1], self.widths[1:]))
        if θ is None:
            # Glorot-uniform weights, zero bias (Lux defaults).  The reference seeds MersenneTwister(666); that stream is
            # not reproducible outside Julia, so parity work always passes θ explicitly.
            rng = np.random.default_rng(seed)
            parts = []
            for i, o in zip(self.widths[:-1], self.widths[1:]):
                lim = np.sqrt(6.0 / (i + o))
                parts += [rng.uniform(-lim, lim, size=o * i), np.zeros(o)]
            θ = np.concatenate(parts)
        self.θ = np.asarray(θ, dtype=np.float64).copy()
        assert self.θ.size == self.n_params


class LawA:
    """``LawA(nn, params)``: A = minA + (maxA-minA)·NN([T]; θ.A), one evaluation per glacier and solve
    (callback_freq = 0).  ``LawA(params; scalar=true)``: per-glacier θ, A = minA + (maxA-minA)(tanh θ + 1)/2."""

    def __init__(self, nn: Optional[NeuralNetwork] = None, scalar: bool = True):
        self.nn = nn
        self.scalar = scalar
        self.kind = "nn" if nn is not None else ("scalar" if scalar else "gridded")


@dataclass
class SIA2Dmodel:
    A: Optional[LawA] = None
    n: float = 3.0
    C: float = 0.0


@dataclass
class Model:
    iceflow: SIA2Dmodel = field(default_factory=SIA2Dmodel)
    mass_balance: object = None
    regressors: Optional[dict] = None  # (; A = nn)

    @property
    def θ(self):
        law = self.iceflow.A
        return law.nn.θ if (law is not None and law.nn is not None) else self._θ_scalar

    @θ.setter
    def θ(self, v):
        law = self.iceflow.A
        if law is not None and law.nn is not None:
            law.nn.θ = np.asarray(v, dtype=np.float64).copy()
        else:
            self._θ_scalar = np.asarray(v, dtype=np.float64).copy()

    _θ_scalar: np.ndarray = field(default_factory=lambda: np.zeros(0))


@dataclass
class SolverParameters:
    """params.solver: ``step`` = tstops spacing (default 1/12 yr).  ``solver``: "rdpk3sp35" -- the reference's default,
    OrdinaryDiffEq's RDPK3Sp35 with its PID controller restated on the device (csrc/rdpk.cu; adaptive, reltol / abstol) --
    "bs3" (adaptive), or "euler" / "ssprk3" with ``nsub`` fixed sub-steps per tstop interval (each stage fused into the RHS
    kernel, the interval replayed as a CUDA graph: the throughput configuration)."""

    step: float = 1.0 / 12.0
    solver: str = "ssprk3"     # "rdpk3sp35" | "bs3" (adaptive) | "euler" | "ssprk3" (nsub fixed sub-steps per tstop interval)
    nsub: int = 8
    reltol: float = 1e-6
    abstol: float = 1e-6
    maxiters: int = 1_000_000


@dataclass
class DiscreteAdjoint:
    """src/inverse/AdjointTypes.jl: reverse loop over the saved steps (gradient.jl:45-274)."""

    VJP_method: str = "discrete"


@dataclass
class ContinuousAdjoint:
    """src/inverse/AdjointTypes.jl:53-66: reverse ODE on interpolated snapshots + Gauss-Legendre quadrature (gradient.jl:276-538).
    Defaults as upstream: ``solver`` RDPK3Sp35 (adaptive), reltol = abstol = 1e-8, dtmax = 1/12, n_quadrature = 200.
    ``solver`` "euler" | "ssprk3": fixed-step reverse integrator with ``nsub`` sub-steps between consecutive stops.
    ``VJP_method``: "discrete" | "continuous"."""

    VJP_method: str = "discrete"
    n_quadrature: int = 200
    solver: str = "rdpk3sp35"
    reltol: float = 1e-8
    abstol: float = 1e-8
    dtmax: float = 1.0 / 12.0
    nsub: int = 4


@dataclass
class Parameters:
    physical: Phys = field(default_factory=Phys)
    solver: SolverParameters = field(default_factory=SolverParameters)
    tspan: tuple = (2010.0, 2015.0)
    dtype: str = "f64"
    distance: int = 3          # is_in_glacier erosion distance of LossH (Losses.jl:270-291)
    grad: object = field(default_factory=DiscreteAdjoint)   # params.UDE.grad (DiscreteAdjoint | ContinuousAdjoint)
    epochs: int = 50
    lr: float = 1e-2


def define_callback_steps(tspan, step):
    """Huginn.define_callback_steps [not in tree]; call sites gradient.jl:96,131, inversion_utils.jl:487."""
    n = int(round((tspan[1] - tspan[0]) / step))
    return tspan[0] + step * np.arange(n + 1)


def create_interpolation(A, n_interp_half: int, dilation_factor: float = 1.0, minA_unif=None, minA_quantile=None,
                         maxA_unif=None, maxA_quantile=None) -> np.ndarray:
    """``create_interpolation`` (src/models/target/target_utils.jl:245-299): the 2·n_interp_half knots of the law-gradient
    interpolation -- n_interp_half equally spaced points of [minA_unif, dilation_factor·max(A)] united with n_interp_half
    interior quantiles of the values inside (minA_quantile, maxA_quantile).  Feed the result to
    ``Ensemble.law_cell_interp_set`` (what ``feed_input_cache!`` does upstream, src/laws/Cache.jl:130-154: H̄ knots with
    dilation 1.05 and quantiles from 10 m, ∇S knots with dilation 1.05, over ALL snapshots of the forward solve)."""
    A = np.asarray(A, dtype=np.float64).reshape(-1)
    minA_unif = 0.0 if minA_unif is None else minA_unif
    minA_quantile = 0.0 if minA_quantile is None else minA_quantile
    maxA_unif = dilation_factor * A.max() if maxA_unif is None else maxA_unif
    maxA_quantile = A.max() if maxA_quantile is None else maxA_quantile
    if not (minA_unif < maxA_unif and minA_quantile < maxA_quantile):
        raise ValueError("There are not enough different values of A to create a proper interpolation.")
    unif = np.linspace(minA_unif, maxA_unif, n_interp_half)
    quant = np.quantile(A[(minA_quantile < A) & (A < maxA_quantile)], np.linspace(0.0, 1.0, n_interp_half + 2)[1:-1])
    knots = np.unique(np.concatenate([unif, quant]))
    if knots.size < 2 * n_interp_half:  # coinciding knots: top up with midpoints (the reference picks the gaps at random)
        n_left = 2 * n_interp_half - knots.size
        knots = np.sort(np.concatenate([knots, 0.5 * (knots[:n_left] + knots[1:n_left + 1])]))
    return knots


def is_in_glacier(A: np.ndarray, distance: int) -> np.ndarray:
    """Sleipnir.is_in_glacier [not in tree].  Assumption: `distance` erosions of the mask A != 0 with the
    5-point cross and circular shifts."""
    m = (np.asarray(A) != 0).astype(np.float64)
    for _ in range(int(distance)):
        m = np.minimum.reduce([m, np.roll(m, 1, 0), np.roll(m, -1, 0), np.roll(m, 1, 1), np.roll(m, -1, 1)])
    return m > 0.001


# ------------------------------------------------------------------------------------------------
# simulations
# ------------------------------------------------------------------------------------------------
class _Simulation:
    def __init__(self, model: Model, glaciers: Sequence[Glacier2D], parameters: Parameters, device: Optional[int] = None,
                 temperatures: Optional[Sequence[float]] = None):
        self.model = model
        self.glaciers_all: List[Glacier2D] = list(glaciers)
        self.parameters = parameters
        rank, local_rank, ws = parallel.world()
        self.rank, self.world_size = rank, ws
        self.t = define_callback_steps(parameters.tspan, parameters.solver.step)
        costs = [g.nx * g.ny * len(self.t) for g in self.glaciers_all]
        self.my_ids = parallel.shard_glaciers(costs, ws)[rank]
        self.glaciers = [self.glaciers_all[k] for k in self.my_ids]
        self.temperatures = [(-10.0 if temperatures is None else float(temperatures[k])) for k in self.my_ids]
        ph = parameters.physical
        ph.n, ph.C = model.iceflow.n, model.iceflow.C
        self.ensemble = None
        if self.glaciers:
            self.ensemble = Ensemble([g.nx for g in self.glaciers], [g.ny for g in self.glaciers], [g.Δx for g in self.glaciers],
                                     [g.Δy for g in self.glaciers], ph, parameters.dtype, local_rank if device is None else device)
            for k, g in enumerate(self.glaciers):
                self.ensemble.upload(k, _capi.FIELD_B, g.B)
                self.ensemble.upload(k, _capi.FIELD_H0, g.H0)
                self.ensemble.set_temperature(k, self.temperatures[k])
        self.A = None

    # apply_all_non_callback_laws! for the A law (adjoint.jl:75-76): once per solve
    def apply_laws(self, θ):
        law = self.model.iceflow.A
        ens = self.ensemble
        if ens is None:
            return
        ph = self.parameters.physical
        if law is None:
            return
        if law.kind == "nn":
            self.A = ens.law_A_nn_apply(law.nn.widths, law.nn.acts, θ)
        elif law.kind == "scalar":
            th = np.asarray(θ, dtype=np.float64)[self.my_ids]
            self.A = ph.minA + (ph.maxA - ph.minA) * (np.tanh(th) + 1.0) / 2.0
            for k, a in enumerate(self.A):
                ens.set_A_scalar(k, float(a))
        else:
            raise NotImplementedError("gridded A: use the per-call VJP operators (sia2d.VJP_λ_dSIAdθ)")

    def solve(self):
        if self.ensemble is not None:
            sp = self.parameters.solver
            if sp.solver in ("bs3", "rdpk3sp35"):
                self.ensemble.solve_forward_adaptive(self.t, reltol=sp.reltol, abstol=sp.abstol, max_steps=sp.maxiters, method=sp.solver)
            else:
                self.ensemble.solve_forward(self.t, method=sp.solver, nsub=sp.nsub)

    def close(self):
        if self.ensemble is not None:
            self.ensemble.close()
            self.ensemble = None


class Prediction(_Simulation):
    """Forward simulation of an ensemble (README.md:41-79, docs/src/forward_simulation.jl:20-44)."""

    def set_A(self, A):
        A = [A] * len(self.glaciers_all) if np.isscalar(A) else list(A)
        for k, gid in enumerate(self.my_ids):
            self.ensemble.set_A_scalar(k, float(A[gid]))


class Inversion(_Simulation):
    """Functional inversion (UDE training) state: reference data + trainable θ."""

    def __init__(self, model, glaciers, parameters, H_ref: Sequence[Sequence[np.ndarray]], **kw):
        super().__init__(model, glaciers, parameters, **kw)
        n = len(self.t)
        for k, gid in enumerate(self.my_ids):
            # H_ref[gid][j] is None at tstops without thickness data (tH_ref a strict subset of the tstops): the library then weights
            # snapshot j by the time since the previous datum, 0 for the first one (diff(tH_ref), gradient.jl:79-80, 144-149)
            assert len(H_ref[gid]) == n, "one entry per tstop (None where there is no thickness data)"
            for j in range(n):
                if H_ref[gid][j] is not None:
                    self.ensemble.set_reference(k, j, n, H_ref[gid][j], is_in_glacier(H_ref[gid][j], parameters.distance))
        self.stats = {"losses": [], "grad_norms": []}


FunctionalInversion = Inversion


def run_(simulation, **kw):
    """``run!(simulation)``.  Prediction: forward solve, returns per-glacier lists of snapshots (this rank's glaciers).
    Inversion: Adam training loop (train_UDE!, inversion_utils.jl:177-238), returns θ."""
    if isinstance(simulation, Prediction):
        simulation.solve()
        ens = simulation.ensemble
        return [[ens.get_snapshot(k, j) for j in range(len(simulation.t))] for k in range(len(simulation.glaciers))]
    return train_UDE_(simulation, **kw)


def loss_iceflow_transient(θ, simulation: Inversion) -> float:
    """Σ over glaciers of the transient LossH (inversion_utils.jl:287-296): law → forward solve → loss; summed over ranks."""
    simulation.model.θ = θ
    loss = 0.0
    if simulation.ensemble is not None:
        simulation.apply_laws(θ)
        simulation.solve()
        loss = float(simulation.ensemble.loss(simulation.t).sum())
    loss, _ = parallel.allreduce_loss_grad(loss, np.zeros(0))
    return loss


def SIA2D_grad_(dθ, θ, simulation: Inversion) -> float:
    """``SIA2D_grad!(dθ, θ, simulation)`` (gradient.jl:6-31): forward solve, DiscreteAdjoint reverse loop on every
    glacier of this rank, law pullback, then ONE all-reduce of [loss; dθ].  Writes dθ in place, returns the loss."""
    simulation.model.θ = θ
    law = simulation.model.iceflow.A
    ens = simulation.ensemble
    θ = np.asarray(θ, dtype=np.float64)
    g_local = np.zeros_like(θ)
    loss = 0.0
    if ens is not None:
        simulation.apply_laws(θ)
        simulation.solve()
        gm = simulation.parameters.grad
        if isinstance(gm, ContinuousAdjoint) and gm.solver == "rdpk3sp35":
            losses, Ssum, _ = ens.grad_continuous_adaptive(simulation.t, n_quadrature=gm.n_quadrature, vjp=gm.VJP_method, reltol=gm.reltol,
                                                           abstol=gm.abstol, dtmax=gm.dtmax, max_steps=simulation.parameters.solver.maxiters)
        elif isinstance(gm, ContinuousAdjoint):
            losses, Ssum = ens.grad_continuous(simulation.t, n_quadrature=gm.n_quadrature, vjp=gm.VJP_method, method=gm.solver,
                                               nsub=gm.nsub)
        else:
            losses, Ssum = ens.grad_discrete(simulation.t)
        loss = float(losses.sum())
        if law.kind == "nn":
            g_local = ens.law_A_nn_pullback(law.nn.n_params)
        else:  # per-glacier scalar law: dA/dθ_g = (maxA-minA)/2 · (1 - tanh²θ_g), scattered by glacier id
            ph = simulation.parameters.physical
            th = θ[simulation.my_ids]
            g_local[simulation.my_ids] = (ph.maxA - ph.minA) * 0.5 * (1.0 - np.tanh(th) ** 2) * Ssum
        nrm = float(np.linalg.norm(g_local))
        if nrm > 1e7:  # gradient.jl:19-24
            import warnings

            warnings.warn(f"Potential unstable gradient: ‖dθ‖={nrm:.3e}; try reducing the reverse step Δt")
    loss, g = parallel.allreduce_loss_grad(loss, g_local)
    dθ[...] = g
    return loss


def train_UDE_(simulation: Inversion, epochs: Optional[int] = None, lr: Optional[float] = None,
               callback: Optional[Callable] = None):
    """Adam on the host (Optimisers.Adam in the reference), gradient from SIA2D_grad_."""
    p = simulation.parameters
    epochs = p.epochs if epochs is None else epochs
    lr = p.lr if lr is None else lr
    θ = np.array(simulation.model.θ, dtype=np.float64)
    m, v = np.zeros_like(θ), np.zeros_like(θ)
    b1, b2, eps = 0.9, 0.999, 1e-8
    g = np.zeros_like(θ)
    for it in range(1, epochs + 1):
        loss = SIA2D_grad_(g, θ, simulation)
        simulation.stats["losses"].append(loss)
        simulation.stats["grad_norms"].append(float(np.linalg.norm(g)))
        if callback is not None:
            callback(θ, loss)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        θ = θ - lr * (m / (1 - b1**it)) / (np.sqrt(v / (1 - b2**it)) + eps)
    simulation.model.θ = θ
    return θ
