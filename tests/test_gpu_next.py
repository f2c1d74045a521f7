"""GPU parity of the SURVEY 8f rows built after the core path: surface velocity / LossV (N2) and the mass-balance callback
with its discrete VJP (N3), all through the C ABI against the NumPy oracle.  Tolerances as in test_gpu_parity /
test_gpu_timeloop: fp64 per-call 1e-12, loops 1e-10 / 1e-8; fp32 per-call 1e-5, loops 2e-3."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

PH = dict(minA=8e-21, maxA=8e-17)


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    return odinn_b200


def _ens(ob, gl, dtype, phys_kw=None):
    from odinn_b200 import _capi

    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl],
                      ob.Phys(**(phys_kw or PH)), dtype)
    for k, g in enumerate(gl):
        ens.upload(k, _capi.FIELD_B, g.B)
        ens.upload(k, _capi.FIELD_H0, g.H0)
    return ens


def _r(a, dtype):
    return np.asarray(a).astype(np.float32 if dtype == "f32" else np.float64).astype(np.float64)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("phys_kw", [dict(), dict(C=7e-8), dict(n=3.5)])
def test_surface_velocity_and_its_vjps(ob, dtype, phys_kw):
    gl = [o.rough_bed_glacier(33, 29), o.dome_glacier(17, 40, H0=200.0)]
    A = [3e-17, 2.21e-18]
    kw = dict(PH, **phys_kw)
    ens = _ens(ob, gl, dtype, kw)
    tol = 1e-12 if dtype == "f64" else 1e-5
    rng = np.random.default_rng(5)
    try:
        for k, g in enumerate(gl):
            ens.set_A_scalar(k, A[k])
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy)
            H = _r(g.H0, dtype)
            tg = o.TargetA(o.Phys(**kw), "const", A=A[k])
            Vx, Vy = ens.surface_velocity(k, H)
            rVx, rVy = o.surface_V(H, g2, tg)
            assert rel_l2(Vx, rVx) <= tol and rel_l2(Vy, rVy) <= tol
            dVx, dVy = _r(rng.standard_normal(H.shape), dtype), _r(rng.standard_normal(H.shape), dtype)
            out, S = ens.vjp_surface_V(k, dVx, dVy, H)
            ref = o.VJP_dsurfaceV_dH_discrete(dVx, dVy, H, g2, tg)
            assert rel_l2(out, ref) <= 10 * tol, rel_l2(out, ref)
            refS = -o.surfaceV_theta_reduction(dVx, dVy, H, g2, tg)
            assert abs(S - refS) <= 1e3 * tol * abs(refS), (S, refS)  # (a cancelling sum of signed terms)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["V", "HV"])
def test_lossV_discrete_adjoint_gradient(ob, dtype, kind):
    """LossV / LossHV (Losses.jl:293-440) in the DiscreteAdjoint reverse loop == oracle.loss_and_grad_discrete_HV."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.0 + 4.0 / 12.0), 1.0 / 12.0)
    has_V = [False, True, False, True, True]
    ph = o.Phys(**PH)
    As = [3e-17, 1.2e-17]
    wH, wV = o.loss_weights(kind, t, has_V, scaling=3.0)
    ens = _ens(ob, gl, dtype)
    try:
        refs = []
        vslots = [j for j in range(len(t)) if has_V[j]]
        for k, g in enumerate(gl):
            g32 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=g.H0)
            tref = o.TargetA(ph, "const", A=5e-17)
            Href = [_r(h, dtype) for h in o.solve_forward(g.H0, g, tref, None, t, method="ssprk3", nsub=8)]
            Hs = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=As[k]), None, t, method="ssprk3", nsub=8)]
            Vref = [None] * len(t)
            for j in range(len(t)):
                ens.set_snapshot(k, j, len(t), Hs[j])
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            for m, j in enumerate(vslots):
                vx, vy, va = (_r(v, dtype) for v in o.V_from_H(Href[j], g32, tref))
                Vref[j] = (vx, vy, va)
                ens.set_velocity_reference(k, m, len(vslots), j, vx, vy, va, scale_loss=True)
            ens.set_A_scalar(k, As[k])
            tgs = o.TargetA(ph, "scalar")
            theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
            ell, dth = o.loss_and_grad_discrete_HV(theta, g32, tgs, t, Hs, Href, Vref, wH, wV, "xy", True)
            refs.append((ell, dth[0], tgs.vjp_theta[0]))
        ens.set_loss_weights(wH, wV, "xy")
        fwd = ens.loss(t)
        loss, Ssum = ens.grad_discrete(t)
        rt_l, rt_g = (1e-10, 1e-8) if dtype == "f64" else (2e-3, 5e-3)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert fwd[k] == pytest.approx(loss[k], rel=1e-8 if dtype == "f64" else 1e-5)  # gradient.jl:259
            assert Ssum[k] * refs[k][2] == pytest.approx(refs[k][1], rel=rt_g), (k, Ssum[k] * refs[k][2], refs[k][1])
        ens.set_loss_weights(None)
        assert ens.grad_continuous(t, n_quadrature=5)[0].shape == (2,)
    finally:
        ens.close()


MB_PAR = (3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("method", ["ssprk3", "bs3"])
def test_mass_balance_forward_and_discrete_gradient(ob, dtype, method):
    """Mass-balance callback at the end of every window (inversion_utils.jl:498-517) in the forward solve, MB history, and
    VJP_λ_∂MB∂H in the reverse loop (VJPs.jl:107-151, gradient.jl:201-207) == the oracle with the same callback."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.0 + 4.0 / 12.0), 1.0 / 12.0)
    mb_idx = [2, 4]  # step_MB = 2 months
    pars = np.array([[MB_PAR, (4.0, -0.006, 2050.0, 0.7, 0.5, 1.0, 0.5)] for _ in mb_idx])  # [n_mb, G, 7]
    ph = o.Phys(**PH)
    As = [3e-17, 1.2e-17]
    ens = _ens(ob, gl, dtype)
    solve = dict(method="ssprk3", nsub=8) if method == "ssprk3" else dict(method="bs3", reltol=1e-5, abstol=1e-5)
    try:
        ens.set_mass_balance(mb_idx, pars)
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        if method == "ssprk3":
            ens.solve_forward(t, method="ssprk3", nsub=8)
        else:
            ens.solve_forward_adaptive(t, reltol=1e-5, abstol=1e-5)
        tolH = 1e-10 if dtype == "f64" else 1e-3
        for k, g in enumerate(gl):
            g32 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            mb = {j: tuple(pars[m, k]) for m, j in enumerate(mb_idx)}
            st = {}
            Hs = o.solve_forward(g32.H0, g32, o.TargetA(ph, "const", A=As[k]), None, t, mb=mb, stats=st, **solve)
            for j in range(len(t)):
                assert rel_l2(ens.get_snapshot(k, j), Hs[j]) <= tolH, (k, j)
            for m, j in enumerate(mb_idx):
                assert rel_l2(ens.get_mass_balance(k, m), st["MB"][j]) <= (1e-10 if dtype == "f64" else 1e-4), (k, m)
                assert np.abs(st["MB"][j]).max() > 0
            if method == "ssprk3":  # reverse loop on the device snapshots
                Href = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, mb=mb, **solve)]
                Hs_d = [ens.get_snapshot(k, j).astype(np.float64) for j in range(len(t))]
                MBh = {j: ens.get_mass_balance(k, m).astype(np.float64) for m, j in enumerate(mb_idx)}
                for j in range(len(t)):
                    ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
                tgs = o.TargetA(ph, "scalar")
                theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
                wH, wV = o.loss_weights("H", t)
                ell, dth = o.loss_and_grad_discrete_HV(theta, g32, tgs, t, Hs_d, Href, [None] * len(t), wH, wV, mb=mb, MB_hist=MBh)
                ell0, dth0 = o.loss_and_grad_discrete_HV(theta, g32, tgs, t, Hs_d, Href, [None] * len(t), wH, wV)
                assert abs(dth[0] - dth0[0]) > 1e-6 * abs(dth0[0])  # the MB VJP matters in this setup
                gl[k]._ref = (ell, dth[0], tgs.vjp_theta[0])
        if method == "ssprk3":
            loss, Ssum = ens.grad_discrete(t)
            rt_l, rt_g = (1e-10, 1e-8) if dtype == "f64" else (1e-4, 2e-3)
            for k, g in enumerate(gl):
                assert loss[k] == pytest.approx(g._ref[0], rel=rt_l), k
                assert Ssum[k] * g._ref[2] == pytest.approx(g._ref[1], rel=rt_g), (k, Ssum[k] * g._ref[2], g._ref[1])
    finally:
        ens.close()


def test_error_convention_of_the_new_entry_points(ob):
    """Bad arguments and call-sequence errors come back as a negative status + message (OdinnError in the mirror), never a crash."""
    g = o.rough_bed_glacier(12, 11)
    g.H0 = 0.5 * g.H0
    ens = _ens(ob, [g], "f64")
    t = np.array([2010.0, 2010.1, 2010.2])
    try:
        with pytest.raises(ob.OdinnError, match="strictly increasing"):
            ens.solve_forward_adaptive(np.array([2010.0, 2010.0, 2010.1]))
        with pytest.raises(ob.OdinnError, match="adaptive-solve"):
            ens.solve_forward_adaptive(t, reltol=-1.0)
        with pytest.raises(ob.OdinnError, match="snapshots and reference"):
            ens.grad_continuous(t, n_quadrature=4)
        ens.set_A_scalar(0, 2.21e-18)
        ens.solve_forward(t, method="ssprk3", nsub=16)
        assert np.isfinite(ens.get_snapshot(0, 2)).all()
        with pytest.raises(ob.OdinnError):  # no reference data yet
            ens.grad_continuous(t, n_quadrature=4)
        for j in range(3):
            ens.set_reference(0, j, 3, ens.get_snapshot(0, j), np.ones(g.B.shape, bool))
        with pytest.raises(ob.OdinnError, match="count"):
            ens.grad_continuous(t[:2], n_quadrature=4)
        with pytest.raises(ob.OdinnError, match="mass-balance step"):
            ens.get_mass_balance(0, 0)
        with pytest.raises(ob.OdinnError, match="velocity reference"):
            ens.set_velocity_reference(0, 3, 2, 1, g.B, g.B, g.B)
        loss, Ssum = ens.grad_continuous(t, n_quadrature=4)
        assert loss[0] == 0.0 and Ssum[0] == 0.0, (loss, Ssum)  # reference == prediction: zero loss, zero gradient
    finally:
        ens.close()
