"""Thin object wrapper over the C ABI: one ``Ensemble`` == one ``odinn_ensemble*`` handle.

An ensemble is the batch of independent glaciers one worker owns -- the unit the reference maps
with ``pmap`` (src/inverse/SIA2D/gradient.jl:9-10, src/models/trainable_components/ML_utils.jl:135-144).
Matrices cross the boundary exactly as Julia stores them: column-major ``(nx, ny)``, ``i`` fastest.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Phys


def _np_dtype(dtype):
    return np.float32 if dtype == _capi.F32 else np.float64


def _as_f(a, npdt):
    """Column-major contiguous view/copy (what a Julia ``Matrix`` pointer is)."""
    return np.asfortranarray(a, dtype=npdt)


class Ensemble:
    def __init__(self, nx: Sequence[int], ny: Sequence[int], dx: Sequence[float], dy: Sequence[float],
                 phys: Optional[Phys] = None, dtype: str = "f64", device: int = 0):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.dtype_code = {"f32": _capi.F32, "f64": _capi.F64}[dtype]
        self.np_dtype = _np_dtype(self.dtype_code)
        self.G = len(nx)
        self.nx, self.ny = [int(v) for v in nx], [int(v) for v in ny]
        self.phys = phys if phys is not None else Phys()
        ia = (C.c_int * self.G)(*self.nx)
        ja = (C.c_int * self.G)(*self.ny)
        da = (C.c_double * self.G)(*[float(v) for v in dx])
        db = (C.c_double * self.G)(*[float(v) for v in dy])
        rc = self._lib.odinn_ensemble_create(int(device), self.dtype_code, self.G, ia, ja, da, db, C.byref(self.phys),
                                             C.byref(self._h))
        _capi.check(None, rc)

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.odinn_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _capi.check(self._h, rc)

    def _mat(self, g: int, a, name: str = "matrix", dual: bool = False):
        """Column-major (nx, ny) [dual grid: (nx-1, ny-1)] matrix of glacier g in the ensemble dtype.  The C side only sees a
        pointer and a leading dimension, so a wrong shape would read or write past the host buffer: refuse it here."""
        if not (0 <= int(g) < self.G):
            raise _capi.OdinnError(f"glacier index {g} out of range (0..{self.G - 1})")
        a = _as_f(a, self.np_dtype)
        want = (self.nx[g] - 1, self.ny[g] - 1) if dual else (self.nx[g], self.ny[g])
        if a.shape != want:
            raise _capi.OdinnError(f"{name} of glacier {g} has shape {a.shape}, expected {want}")
        return a

    def _out(self, g: int, out, name: str = "out"):
        """A caller-supplied output matrix must already be what the C side writes: right shape, dtype, column-major."""
        want = (self.nx[g], self.ny[g])
        if not isinstance(out, np.ndarray) or out.shape != want or out.dtype != self.np_dtype or not out.flags.f_contiguous \
                or not out.flags.writeable:
            raise _capi.OdinnError(f"{name} of glacier {g} must be a writeable column-major {self.np_dtype.__name__} array of shape {want}")
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.odinn_launch_count(self._h))

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of this handle (record timing events on it)."""
        return int(self._lib.odinn_stream(self._h) or 0)

    def synchronize(self):
        self._ck(self._lib.odinn_synchronize(self._h))

    # -- state -----------------------------------------------------------------------------
    def upload(self, g: int, field: int, a):
        a = self._mat(g, a, "field", dual=field in (_capi.FIELD_A, _capi.FIELD_VJP_A))
        self._ck(self._lib.odinn_upload(self._h, g, field, a.ctypes.data, a.shape[0]))

    def download(self, g: int, field: int):
        if not (0 <= int(g) < self.G):
            raise _capi.OdinnError(f"glacier index {g} out of range (0..{self.G - 1})")
        dual = field in (_capi.FIELD_A, _capi.FIELD_VJP_A)
        shape = (self.nx[g] - 1, self.ny[g] - 1) if dual else (self.nx[g], self.ny[g])
        out = np.empty(shape, dtype=self.np_dtype, order="F")
        self._ck(self._lib.odinn_download(self._h, g, field, out.ctypes.data, shape[0]))
        return out

    def set_A_scalar(self, g: int, A: float):
        self._ck(self._lib.odinn_set_A_scalar(self._h, g, float(A)))

    def set_A_field(self, g: int, A):
        self._ck(self._lib.odinn_set_A_mode(self._h, 1))
        self.upload(g, _capi.FIELD_A, A)

    def set_A_mode(self, gridded: bool):
        self._ck(self._lib.odinn_set_A_mode(self._h, int(bool(gridded))))

    def set_phys(self, phys: Phys):
        self.phys = phys
        self._ck(self._lib.odinn_set_phys(self._h, C.byref(self.phys)))

    # -- reference-facing per-call operators (host in, host out) ----------------------------------
    def sia2d_rhs(self, g: int, H, t: float = 0.0, out=None):
        H = self._mat(g, H, "H")
        dH = np.empty_like(H, order="F") if out is None else self._out(g, out, "dH")
        self._ck(self._lib.odinn_sia2d_rhs(self._h, g, H.ctypes.data, H.shape[0], dH.ctypes.data, dH.shape[0], float(t)))
        return dH

    def sia2d_vjp_H(self, g: int, lam, H, t: float = 0.0, continuous: bool = False):
        H = self._mat(g, H, "H")
        lam = self._mat(g, lam, "lambda")
        out = np.empty_like(H, order="F")
        fn = self._lib.odinn_sia2d_vjp_H_continuous if continuous else self._lib.odinn_sia2d_vjp_H
        self._ck(fn(self._h, g, lam.ctypes.data, lam.shape[0], H.ctypes.data, H.shape[0],
                                             out.ctypes.data, out.shape[0], float(t)))
        return out

    def sia2d_vjp_theta(self, g: int, lam, H, t: float = 0.0, continuous: bool = False) -> float:
        H = self._mat(g, H, "H")
        lam = self._mat(g, lam, "lambda")
        S = C.c_double(0.0)
        fn = self._lib.odinn_sia2d_vjp_theta_continuous if continuous else self._lib.odinn_sia2d_vjp_theta
        self._ck(fn(self._h, g, lam.ctypes.data, lam.shape[0], H.ctypes.data, H.shape[0],
                                                 C.byref(S), float(t)))
        return S.value

    # -- device-resident ensemble operators ---------------------------------------------------
    def rhs_resident(self):
        self._ck(self._lib.odinn_rhs_resident(self._h))

    def vjp_resident(self, want_H=True, want_S=True, read_S=True, continuous=False, want_dH=False):
        """want_dH: also FIELD_DH <- SIA2D(FIELD_H) (the (λ_∂f∂H, dH) pair of VJPs.jl:12-28), fused into the same pass."""
        flags = (1 if want_H else 0) | (2 if want_S else 0) | (4 if continuous else 0) | (8 if want_dH else 0)
        if want_S and read_S:
            S = np.empty(self.G, dtype=np.float64)
            self._ck(self._lib.odinn_vjp_resident(self._h, flags, S.ctypes.data_as(C.POINTER(C.c_double))))
            return S
        self._ck(self._lib.odinn_vjp_resident(self._h, flags, None))
        return None

    def fwd_adj_batch_host(self, H_ptrs, lam_ptrs, dH_ptrs, vjpH_ptrs, S):
        """Raw-pointer batched call (bench e2e path).  Each ``*_ptrs`` is a ctypes ``c_void_p`` array or None."""
        sp = S.ctypes.data_as(C.POINTER(C.c_double)) if S is not None else None
        self._ck(self._lib.odinn_fwd_adj_batch_host(self._h, H_ptrs, lam_ptrs, dH_ptrs, vjpH_ptrs, sp))

    def set_batch_chunk(self, cells: int):
        self._ck(self._lib.odinn_set_batch_chunk(self._h, int(cells)))

    def set_cluster_mode(self, mode: int):
        """-1 automatic, 0 marching kernels only, 1/2/4/8/16: cluster size of the shared-memory-resident forward solve."""
        self._ck(self._lib.odinn_set_cluster_mode(self._h, int(mode)))

    def fwd_adj_batch(self, Hs, lams=None, want_dH=True, want_vjpH=True, want_S=True):
        """NumPy front end of ``odinn_fwd_adj_batch_host``: lists of (nx, ny) matrices in, (dH list, vjpH list, S) out."""
        if len(Hs) != self.G or (lams is not None and len(lams) != self.G):
            raise _capi.OdinnError(f"one matrix per glacier expected ({self.G})")
        Hs = [self._mat(g, h, "H") for g, h in enumerate(Hs)]
        mk = lambda arrs: (C.c_void_p * self.G)(*[a.ctypes.data for a in arrs])
        adj = want_vjpH or want_S
        ls = [self._mat(g, l, "lambda") for g, l in enumerate(lams)] if adj else None
        dH = [np.empty_like(h, order="F") for h in Hs] if want_dH else None
        vH = [np.empty_like(h, order="F") for h in Hs] if want_vjpH else None
        S = np.empty(self.G, dtype=np.float64) if want_S else None
        self.fwd_adj_batch_host(mk(Hs), mk(ls) if adj else None, mk(dH) if want_dH else None, mk(vH) if want_vjpH else None, S)
        return dH, vH, S

    # -- laws ------------------------------------------------------------------------------------
    def set_temperature(self, g: int, T: float):
        self._ck(self._lib.odinn_set_temperature(self._h, g, float(T)))

    def law_A_nn_apply(self, widths, acts, theta):
        """A_g = minA + (maxA-minA)·NN([T_g]; θ) for every glacier (LawA f!, Laws.jl:348-358); returns A (G,)."""
        widths = [int(w) for w in widths]
        codes = [_capi.ACT[a] if isinstance(a, str) else int(a) for a in acts]
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        A = np.empty(self.G, dtype=np.float64)
        self._ck(self._lib.odinn_law_A_nn_apply(self._h, len(codes), (C.c_int * len(widths))(*widths),
                                                (C.c_int * len(codes))(*codes),
                                                theta.ctypes.data_as(C.POINTER(C.c_double)), theta.size,
                                                A.ctypes.data_as(C.POINTER(C.c_double))))
        return A

    def law_cell_nn_set(self, kind: str, widths, acts, theta, prescale_bounds=None, max_NN=None, n_H=None, n_gS=None):
        """Per-cell law: kind "U" (LawU, D = H̄·U(H̄,∇S)) or "Y" (LawY hybrid, Y(T,H̄)).  prescale_bounds = ((lo0,hi0),(lo1,hi1))."""
        widths = [int(w) for w in widths]
        codes = [_capi.ACT[a] if isinstance(a, str) else int(a) for a in acts]
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        pb = None
        if prescale_bounds is not None:
            pb = (C.c_double * 4)(prescale_bounds[0][0], prescale_bounds[0][1], prescale_bounds[1][0], prescale_bounds[1][1])
        self._law_n_theta = theta.size
        self._ck(self._lib.odinn_law_cell_nn_set(self._h, {"U": _capi.LAW_U, "Y": _capi.LAW_Y}[kind], len(codes),
                                                 (C.c_int * len(widths))(*widths), (C.c_int * len(codes))(*codes),
                                                 theta.ctypes.data_as(C.POINTER(C.c_double)), theta.size, pb,
                                                 float(max_NN) if max_NN else 0.0, float(n_H) if n_H else 0.0,
                                                 float(n_gS) if n_gS else 0.0))

    def law_cell_interp_set(self, knots0=None, knots1=None):
        """interpolation = :Linear of the law pullback: knots of H̄ (and of ∇S for LawU), e.g. from ``create_interpolation``;
        ``law_cell_interp_set()`` restores the exact per-node gradient."""
        dp = C.POINTER(C.c_double)
        if knots0 is None:
            self._ck(self._lib.odinn_law_cell_interp_set(self._h, 0, None, 0, None))
            return
        k0 = np.ascontiguousarray(knots0, dtype=np.float64)
        k1 = np.ascontiguousarray(knots1, dtype=np.float64) if knots1 is not None else None
        self._ck(self._lib.odinn_law_cell_interp_set(self._h, k0.size, k0.ctypes.data_as(dp), 0 if k1 is None else k1.size,
                                                     None if k1 is None else k1.ctypes.data_as(dp)))

    def law_cell_clear(self):
        self._ck(self._lib.odinn_law_cell_clear(self._h))

    def sia2d_vjp_theta_cell(self, g: int, lam, H, t: float = 0.0):
        """∂θ (vector) of the per-cell law for one glacier: Σ (∂D/∂θ_k)·D†."""
        H = self._mat(g, H, "H")
        lam = self._mat(g, lam, "lambda")
        out = np.empty(self._law_n_theta, dtype=np.float64)
        self._ck(self._lib.odinn_sia2d_vjp_theta_cell(self._h, g, lam.ctypes.data, lam.shape[0], H.ctypes.data, H.shape[0],
                                                      out.ctypes.data_as(C.POINTER(C.c_double)), out.size, float(t)))
        return out

    def law_cell_grad(self):
        """(G, n_theta) per-glacier θ-gradients left by vjp_resident(want_S=True) or accumulated by grad_discrete."""
        out = np.empty((self.G, self._law_n_theta), dtype=np.float64)
        self._ck(self._lib.odinn_law_cell_grad(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), self._law_n_theta))
        return out

    def law_A_nn_pullback(self, n_theta: int, S=None):
        """dθ = Σ_g (∂A_g/∂θ)·S_g ; S=None uses the sums left on the device by grad_discrete."""
        out = np.empty(n_theta, dtype=np.float64)
        sp = None
        if S is not None:
            S = np.ascontiguousarray(S, dtype=np.float64)
            sp = S.ctypes.data_as(C.POINTER(C.c_double))
        self._ck(self._lib.odinn_law_A_nn_pullback(self._h, sp, out.ctypes.data_as(C.POINTER(C.c_double)), n_theta))
        return out

    # -- device-resident time loop and gradient -----------------------------------------------------
    def solve_forward(self, t, method: str = "ssprk3", nsub: int = 8):
        t = np.ascontiguousarray(t, dtype=np.float64)
        m = {"euler": _capi.EULER, "ssprk3": _capi.SSPRK3}[method]
        self._ck(self._lib.odinn_solve_forward(self._h, m, t.size, t.ctypes.data_as(C.POINTER(C.c_double)), int(nsub)))

    def solve_forward_adaptive(self, t, reltol: float = 1e-6, abstol: float = 1e-6, dt0: float = 0.0, max_steps: int = 1_000_000,
                               method: str = "bs3"):
        """Adaptive solve with per-glacier steps and tstops ``t`` (saveat = tstops); returns (steps, rejected) per glacier.
        method: "rdpk3sp35" (the reference's default solver: RDPK3Sp35 + PID controller) | "bs3"."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        steps = np.zeros(self.G, dtype=np.int32)
        rej = np.zeros(self.G, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        m = {"bs3": _capi.BS3, "rdpk3sp35": _capi.RDPK3SP35}[method]
        self._ck(self._lib.odinn_solve_forward_adaptive(self._h, m, t.size, t.ctypes.data_as(C.POINTER(C.c_double)),
                                                        float(reltol), float(abstol), float(dt0), int(max_steps),
                                                        steps.ctypes.data_as(ip), rej.ctypes.data_as(ip)))
        return steps, rej

    def get_snapshot(self, g: int, j: int):
        out = np.empty((self.nx[g], self.ny[g]), dtype=self.np_dtype, order="F")
        self._ck(self._lib.odinn_get_snapshot(self._h, g, j, out.ctypes.data, out.shape[0]))
        return out

    def set_snapshot(self, g: int, j: int, n_snap: int, H):
        H = self._mat(g, H, "snapshot")
        self._ck(self._lib.odinn_set_snapshot(self._h, g, j, n_snap, H.ctypes.data, H.shape[0]))

    def set_reference(self, g: int, j: int, n_snap: int, Href, mask):
        """mask: boolean is_in_glacier(H_ref, distance); stored as W = mask / (nx·ny) (gradient.jl:161)."""
        Href = self._mat(g, Href, "H_ref")
        W = self._mat(g, np.asarray(mask, dtype=np.float64) / float(self.nx[g] * self.ny[g]), "mask")
        self._ck(self._lib.odinn_set_reference(self._h, g, j, n_snap, Href.ctypes.data, W.ctypes.data, Href.shape[0]))

    def loss(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty(self.G, dtype=np.float64)
        self._ck(self._lib.odinn_loss(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), t.size,
                                      out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def grad_discrete(self, t):
        """Returns (loss per glacier, Ssum per glacier) of the DiscreteAdjoint reverse loop."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        loss = np.empty(self.G, dtype=np.float64)
        Ssum = np.empty(self.G, dtype=np.float64)
        self._ck(self._lib.odinn_grad_discrete(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), t.size,
                                               loss.ctypes.data_as(C.POINTER(C.c_double)),
                                               Ssum.ctypes.data_as(C.POINTER(C.c_double))))
        return loss, Ssum

    def grad_continuous(self, t, n_quadrature: int = 200, vjp: str = "discrete", method: str = "ssprk3", nsub: int = 4):
        """ContinuousAdjoint gradient (gradient.jl:276-538): returns (loss per glacier, Ssum per glacier).

        The Gauss-Legendre rule of GaussQuadrature (gradient.jl:560-566) is built here on the host (n_quadrature = 200 is the
        reference's default, src/inverse/AdjointTypes.jl)."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        x, w = np.polynomial.legendre.leggauss(int(n_quadrature))
        qn = np.ascontiguousarray(0.5 * (t[0] + t[-1]) + x * 0.5 * (t[-1] - t[0]))
        qw = np.ascontiguousarray(0.5 * (t[-1] - t[0]) * w)
        loss = np.empty(self.G, dtype=np.float64)
        Ssum = np.empty(self.G, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        m = {"euler": _capi.EULER, "ssprk3": _capi.SSPRK3}[method]
        self._ck(self._lib.odinn_grad_continuous(self._h, t.ctypes.data_as(dp), t.size, qn.size, qn.ctypes.data_as(dp),
                                                 qw.ctypes.data_as(dp), 1 if vjp == "continuous" else 0, m, int(nsub),
                                                 loss.ctypes.data_as(dp), Ssum.ctypes.data_as(dp)))
        return loss, Ssum

    def grad_continuous_adaptive(self, t, n_quadrature: int = 200, vjp: str = "discrete", reltol: float = 1e-8, abstol: float = 1e-8,
                                 dtmax: float = 1.0 / 12.0, max_steps: int = 1_000_000):
        """ContinuousAdjoint gradient with the reference's default reverse solve: adaptive RDPK3Sp35, reltol = abstol = 1e-8,
        dtmax = 1/12, n_quadrature = 200 (src/inverse/AdjointTypes.jl:53-66).  Returns (loss, Ssum, trial steps) per glacier."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        x, w = np.polynomial.legendre.leggauss(int(n_quadrature))
        qn = np.ascontiguousarray(0.5 * (t[0] + t[-1]) + x * 0.5 * (t[-1] - t[0]))
        qw = np.ascontiguousarray(0.5 * (t[-1] - t[0]) * w)
        loss = np.empty(self.G, dtype=np.float64)
        Ssum = np.empty(self.G, dtype=np.float64)
        steps = np.zeros(self.G, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        self._ck(self._lib.odinn_grad_continuous_adaptive(self._h, t.ctypes.data_as(dp), t.size, qn.size, qn.ctypes.data_as(dp),
                                                          qw.ctypes.data_as(dp), 1 if vjp == "continuous" else 0, float(reltol),
                                                          float(abstol), float(dtmax), int(max_steps), loss.ctypes.data_as(dp),
                                                          Ssum.ctypes.data_as(dp), steps.ctypes.data_as(C.POINTER(C.c_int))))
        return loss, Ssum, steps

    def set_velocity_quadrature(self, theta_scale: float, scale_loss: bool = True):
        """Continuous adjoint with a velocity loss: multiplier of the quadrature-weighted dl_V/dtheta term (1 for LossV,
        LossHV.scaling for LossHV, 0 = off) and LossV.scale_loss (gradient.jl:474-507)."""
        self._ck(self._lib.odinn_set_velocity_quadrature(self._h, float(theta_scale), int(bool(scale_loss))))

    # -- surface velocity / LossV ---------------------------------------------------------------------
    def surface_velocity(self, g: int, H, t: float = 0.0):
        """(Vx, Vy) = V_from_H(H) (Huginn.V_from_H call sites Losses.jl:314, 358)."""
        H = self._mat(g, H, "H")
        Vx = np.empty_like(H, order="F")
        Vy = np.empty_like(H, order="F")
        self._ck(self._lib.odinn_surface_velocity(self._h, g, H.ctypes.data, H.shape[0], Vx.ctypes.data, Vy.ctypes.data, Vx.shape[0], float(t)))
        return Vx, Vy

    def vjp_surface_V(self, g: int, dVx, dVy, H, t: float = 0.0):
        """(VJP_λ_∂surface_V∂H, S) with ∂θ = (∂A/∂θ)·S (adjoint.jl:268-413)."""
        H = self._mat(g, H, "H")
        dVx = self._mat(g, dVx, "dVx")
        dVy = self._mat(g, dVy, "dVy")
        out = np.empty_like(H, order="F")
        S = C.c_double(0.0)
        self._ck(self._lib.odinn_sia2d_vjp_surface_V(self._h, g, dVx.ctypes.data, dVy.ctypes.data, dVx.shape[0], H.ctypes.data, H.shape[0],
                                                     out.ctypes.data, out.shape[0], C.byref(S), float(t)))
        return out, S.value

    def set_velocity_reference(self, g: int, slot: int, n_slots: int, snapshot_index: int, Vx_ref, Vy_ref, Vabs_ref, scale_loss: bool = True):
        """Wv = mask / (nx·ny·scale), mask = Vabs_ref > 0, scale = sqrt(mean(Vx_ref² + Vy_ref² over the mask)) (Losses.jl:316-331)."""
        Vabs = np.asarray(Vabs_ref, dtype=np.float64)
        mask = Vabs > 0.0
        sc = float(np.mean(np.asarray(Vx_ref)[mask] ** 2 + np.asarray(Vy_ref)[mask] ** 2) ** 0.5) if (scale_loss and mask.any()) else 1.0
        W = mask.astype(np.float64) / (float(self.nx[g] * self.ny[g]) * sc)
        a = [self._mat(g, x, "velocity reference") for x in (Vx_ref, Vy_ref, Vabs, W)]
        self._ck(self._lib.odinn_set_velocity_reference(self._h, g, slot, n_slots, snapshot_index, a[0].ctypes.data, a[1].ctypes.data,
                                                        a[2].ctypes.data, a[3].ctypes.data, a[0].shape[0]))

    def set_loss_weights(self, wH=None, wV=None, component: str = "xy"):
        dp = C.POINTER(C.c_double)
        if wH is None:
            self._ck(self._lib.odinn_set_loss_weights(self._h, 0, None, None, 0))
            return
        wH = np.ascontiguousarray(wH, dtype=np.float64)
        wV = np.ascontiguousarray(wV if wV is not None else np.zeros_like(wH), dtype=np.float64)
        self._ck(self._lib.odinn_set_loss_weights(self._h, wH.size, wH.ctypes.data_as(dp), wV.ctypes.data_as(dp), 1 if component == "abs" else 0))

    # -- mass balance -------------------------------------------------------------------------------------
    def set_mass_balance(self, snapshot_index, params):
        """params: array [n_mb, G, 7] = (temp, gradient, ref_hgt, snow, DDF, acc_factor, scale) per MB step and glacier."""
        idx = np.ascontiguousarray(snapshot_index, dtype=np.int32)
        if idx.size == 0:
            self._ck(self._lib.odinn_set_mass_balance(self._h, 0, None, None))
            return
        par = np.ascontiguousarray(params, dtype=np.float64)
        assert par.shape == (idx.size, self.G, 7)
        self._ck(self._lib.odinn_set_mass_balance(self._h, idx.size, idx.ctypes.data_as(C.POINTER(C.c_int)),
                                                  par.ctypes.data_as(C.POINTER(C.c_double))))

    def get_mass_balance(self, g: int, m: int):
        out = np.empty((self.nx[g], self.ny[g]), dtype=self.np_dtype, order="F")
        self._ck(self._lib.odinn_get_mass_balance(self._h, g, m, out.ctypes.data, out.shape[0]))
        return out
