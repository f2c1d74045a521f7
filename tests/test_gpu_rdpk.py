"""GPU parity of the reference's DEFAULT integrator on the device (RDPK3Sp35 + PID controller, csrc/rdpk.cu): the adaptive
forward solve and the adaptive reverse solve of the continuous adjoint (with the mass-balance callback and a velocity loss),
against oracle.integrate_rdpk3sp35 / loss_and_grad_continuous_adaptive glacier by glacier.

fp64: the device takes the same accept / reject sequence as the oracle (equal trial-step counts are asserted), states 1e-9 (step
sizes are functions of the error norm, so the RHS roundings feed back into dt), loss 1e-10, d(theta) 1e-7.  fp32: solver-tolerance
level agreement (the fp32 error estimate bottoms out near 1e-4)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

PH = dict(minA=8e-21, maxA=8e-17)
MB_PAR = (3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    return odinn_b200


def _r(a, dtype):
    return np.asarray(a).astype(np.float32 if dtype == "f32" else np.float64).astype(np.float64)


def _glaciers():
    gl = [o.rough_bed_glacier(40, 35), o.rough_bed_glacier(23, 50), o.dome_glacier(33, 33, H0=150.0)]
    for g in gl[:2]:
        g.H0 = 0.6 * g.H0
    return gl


def _ens(ob, gl, dtype):
    from odinn_b200 import _capi

    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl],
                      ob.Phys(**PH), dtype)
    for k, g in enumerate(gl):
        ens.upload(k, _capi.FIELD_B, g.B)
        ens.upload(k, _capi.FIELD_H0, g.H0)
    return ens


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("with_mb", [False, True])
@pytest.mark.parametrize("cluster", [0, -1, 2])   # 0: host-driven engine (rdpk.cu); -1 / 2: the cluster-resident solver (sia2d_cluster.cuh)
def test_forward_rdpk3sp35_matches_oracle(ob, dtype, with_mb, cluster):
    gl = _glaciers()
    As = [4e-17, 2.21e-18, 1.5e-17]
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    rtol = 1e-6 if dtype == "f64" else 1e-4
    mb_idx = [2, 4, 6]
    pars = np.array([[MB_PAR, (4.0, -0.006, 2050.0, 0.7, 0.5, 1.0, 0.5), MB_PAR] for _ in mb_idx])
    ens = _ens(ob, gl, dtype)
    try:
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        if with_mb:
            ens.set_mass_balance(mb_idx, pars)
        ens.set_cluster_mode(cluster)
        l0 = ens.launch_count
        steps, rej = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35")
        if cluster != 0:   # one launch per range between mass-balance callbacks (+ the callbacks)
            assert ens.launch_count - l0 == (2 * len(mb_idx) if with_mb else 1)
        assert np.all(steps > 0) and np.all(rej >= 0)
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            mb = {j: tuple(pars[m, k]) for m, j in enumerate(mb_idx)} if with_mb else None
            st = {}
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(o.Phys(**PH), "const", A=As[k]), None, t, method="rdpk3sp35", reltol=rtol, abstol=rtol,
                                 mb=mb, stats=st)
            if dtype == "f64":  # same accept / reject sequence: trial steps and RHS count (2 for the initial step + 5 per trial step + 1 per MB callback)
                assert steps[k] == st["steps"] and rej[k] == st["rejected"], (k, steps[k], rej[k], st)
            for j in (1, len(t) // 2, len(t) - 1):
                err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                assert err <= (1e-9 if dtype == "f64" else 2e-3), (k, j, err)
            if not with_mb:
                assert abs(ens.get_snapshot(k, len(t) - 1).astype(np.float64).sum() / g2.H0.sum() - 1.0) < (1e-10 if dtype == "f64" else 1e-4)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("vjp", ["discrete", "continuous"])
@pytest.mark.parametrize("cluster", [0, -1, 4])   # 0: host-driven engine; -1 / 4: the cluster-resident reverse solve (discrete flavour only)
def test_adaptive_continuous_adjoint_matches_oracle(ob, dtype, vjp, cluster):
    """ContinuousAdjoint with the reference's default reverse solve (adaptive RDPK3Sp35, reltol = abstol = 1e-8, dtmax = 1/12)."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    ph = o.Phys(**PH)
    As = [3e-17, 1.2e-17]
    tol = 1e-8 if dtype == "f64" else 1e-4
    ens = _ens(ob, gl, dtype)
    try:
        refs = []
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=g.H0)
            Href = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="ssprk3", nsub=8)]
            Hs = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=As[k]), None, t, method="ssprk3", nsub=8)]
            for j in range(len(t)):
                ens.set_snapshot(k, j, len(t), Hs[j])
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            ens.set_A_scalar(k, As[k])
            tgs = o.TargetA(ph, "scalar")
            theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
            st = {}
            ell, dth = o.loss_and_grad_continuous_adaptive(theta, g2, tgs, t, Hs, Href, n_quadrature=9, reltol=tol, abstol=tol, vjp=vjp, stats=st)
            refs.append((ell, dth[0], tgs.vjp_theta[0], st))
        ens.set_cluster_mode(cluster)
        l0 = ens.launch_count
        loss, Ssum, steps = ens.grad_continuous_adaptive(t, n_quadrature=9, vjp=vjp, reltol=tol, abstol=tol)
        if cluster != 0 and vjp == "discrete":
            assert ens.launch_count - l0 == 1   # the whole reverse solve in one launch
        rt_l, rt_g = (1e-10, 1e-7) if dtype == "f64" else (2e-4, 5e-3)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert Ssum[k] * refs[k][2] == pytest.approx(refs[k][1], rel=rt_g), (k, Ssum[k] * refs[k][2], refs[k][1])
            if dtype == "f64":
                assert steps[k] == refs[k][3]["steps"], (k, steps[k], refs[k][3])
        loss2, Ssum2, _ = ens.grad_continuous_adaptive(t, n_quadrature=9, vjp=vjp, reltol=tol, abstol=tol)
        assert np.array_equal(loss, loss2) and np.array_equal(Ssum, Ssum2)  # bit-stable run to run
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("reverse", ["adaptive", "ssprk3"])
def test_continuous_adjoint_with_mass_balance_and_velocity_loss(ob, dtype, reverse):
    """The callbacks of the reverse solve: mass-balance PeriodicCallback (MB before the loss jump, initial_affect at t_end) and a
    LossHV loss -- velocity jumps at the tstops that hold data, dl_V/dtheta quadrature-weighted over interpolated references
    (gradient.jl:289-366, 407-426, 474-507) -- for the adaptive and the fixed-step reverse integrators."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.0 + 4.0 / 12.0), 1.0 / 12.0)
    has_V = [False, True, False, True, True]
    mb_idx = [2, 4]
    pars = np.array([[MB_PAR, (4.0, -0.006, 2050.0, 0.7, 0.5, 1.0, 0.5)] for _ in mb_idx])
    ph = o.Phys(**PH)
    As = [3e-17, 1.2e-17]
    scaling = 3.0
    wH, wV = o.loss_weights("HV", t, has_V, scaling=scaling)
    tol = 1e-8 if dtype == "f64" else 1e-4
    ens = _ens(ob, gl, dtype)
    try:
        ens.set_mass_balance(mb_idx, pars)
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        ens.solve_forward(t, method="ssprk3", nsub=8)
        vslots = [j for j in range(len(t)) if has_V[j]]
        refs = []
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=g.H0)
            mb = {j: tuple(pars[m, k]) for m, j in enumerate(mb_idx)}
            tref = o.TargetA(ph, "const", A=5e-17)
            Href = [_r(h, dtype) for h in o.solve_forward(g.H0, g, tref, None, t, method="ssprk3", nsub=8, mb=mb)]
            Hs = [ens.get_snapshot(k, j).astype(np.float64) for j in range(len(t))]
            MBh = {j: ens.get_mass_balance(k, m).astype(np.float64) for m, j in enumerate(mb_idx)}
            Vref = [None] * len(t)
            for j in range(len(t)):
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            for m, j in enumerate(vslots):
                Vref[j] = tuple(_r(v, dtype) for v in o.V_from_H(Href[j], g2, tref))
                ens.set_velocity_reference(k, m, len(vslots), j, *Vref[j], scale_loss=True)
            tgs = o.TargetA(ph, "scalar")
            theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
            kw = dict(n_quadrature=9, wH=wH, wV=wV, V_ref=Vref, cV=scaling, mb=mb, MB_hist=MBh)
            if reverse == "adaptive":
                ell, dth = o.loss_and_grad_continuous_adaptive(theta, g2, tgs, t, Hs, Href, reltol=tol, abstol=tol, **kw)
            else:
                ell, dth = o.loss_and_grad_continuous_adaptive(theta, g2, tgs, t, Hs, Href, fixed=("ssprk3", 2), **kw)
            refs.append((ell, dth[0], tgs.vjp_theta[0]))
        ens.set_loss_weights(wH, wV, "xy")
        ens.set_velocity_quadrature(scaling, True)
        if reverse == "adaptive":
            loss, Ssum, _ = ens.grad_continuous_adaptive(t, n_quadrature=9, reltol=tol, abstol=tol)
        else:
            loss, Ssum = ens.grad_continuous(t, n_quadrature=9, method="ssprk3", nsub=2)
        rt_l, rt_g = (1e-10, 1e-7) if dtype == "f64" else (2e-4, 5e-3)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert Ssum[k] * refs[k][2] == pytest.approx(refs[k][1], rel=rt_g), (k, Ssum[k] * refs[k][2], refs[k][1])
        # the quadrature-weighted velocity term and the MB callback are live in this setup
        ens.set_velocity_quadrature(0.0, True)
        S0 = ens.grad_continuous_adaptive(t, n_quadrature=9, reltol=tol, abstol=tol)[1]
        assert np.all(np.abs(S0 - Ssum) > 1e-6 * np.abs(Ssum)) or reverse != "adaptive"
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("variant", ["gridded_A", "eta0", "generic_exponents"])
def test_forward_rdpk3sp35_fused_stage_other_kernel_variants(ob, dtype, variant):
    """The RDPK stage epilogue of the F1 kernels (csrc/rdpk.cu: every stage update is fused into the launch that evaluates its slope)
    exists for every template variant of the F1 dispatch: gridded A (LawA(scalar = false), Laws.jl:430-454), eta0 != 1 and generic
    exponents / sliding (n != 3, C != 0).  Host-driven engine (cluster mode 0), ragged ensemble incl. an odd-nx grid, against the oracle."""
    gl = [o.rough_bed_glacier(41, 35), o.rough_bed_glacier(23, 50), o.dome_glacier(33, 33, H0=150.0)]
    for g in gl[:2]:
        g.H0 = 0.6 * g.H0
    kw = dict(PH)
    if variant == "eta0":
        kw.update(eta0=0.6)
    if variant == "generic_exponents":
        kw.update(n=3.3, C=1e-14)   # (a sliding term that matters -- 47 instead of 22 steps -- without making the run stiff)
    ph = o.Phys(**kw)
    rng = np.random.default_rng(5)
    As = []
    for g in gl:
        a = 2e-17 if variant != "generic_exponents" else 2e-18
        As.append(a * np.exp(0.5 * rng.standard_normal((g.B.shape[0] - 1, g.B.shape[1] - 1))) if variant == "gridded_A" else a)
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    rtol = 1e-6 if dtype == "f64" else 1e-4
    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl], ob.Phys(**kw), dtype)
    try:
        from odinn_b200 import _capi

        for k, g in enumerate(gl):
            ens.upload(k, _capi.FIELD_B, g.B)
            ens.upload(k, _capi.FIELD_H0, g.H0)
            if variant == "gridded_A":
                ens.set_A_field(k, As[k])
            else:
                ens.set_A_scalar(k, As[k])
        ens.set_cluster_mode(0)
        steps, rej = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35")
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            st = {}
            Ak = _r(As[k], dtype) if variant == "gridded_A" else As[k]
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(ph, "const", A=Ak), None, t, method="rdpk3sp35", reltol=rtol, abstol=rtol, stats=st)
            if dtype == "f64":
                assert steps[k] == st["steps"] and rej[k] == st["rejected"], (k, steps[k], rej[k], st)
            for j in (1, len(t) - 1):
                err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                assert err <= (1e-9 if dtype == "f64" else 2e-3), (variant, k, j, err)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("phys_kw", [dict(eta0=0.6), dict(n=3.3, C=1e-14)])
def test_adaptive_continuous_adjoint_other_kernel_variants(ob, dtype, phys_kw):
    """The one-launch reverse-ODE stage (interpolation + A1 + RDPK stage update, RKA variants of the A1 kernels) for eta0 != 1 and for
    generic exponents / sliding (fp32: RKA variants of the generic two-column kernel; fp64: the engine falls back to the separate
    passes when n != 3 or C != 0) -- host-driven engine, against the oracle."""
    gl = [o.rough_bed_glacier(31, 30), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    kw = dict(PH)
    kw.update(phys_kw)
    ph = o.Phys(**kw)
    As = [3e-17, 1.2e-17] if "n" not in phys_kw else [3e-18, 1.2e-18]
    tol = 1e-8 if dtype == "f64" else 1e-4
    from odinn_b200 import _capi

    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl], ob.Phys(**kw), dtype)
    try:
        refs = []
        for k, g in enumerate(gl):
            ens.upload(k, _capi.FIELD_B, g.B)
            ens.upload(k, _capi.FIELD_H0, g.H0)
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=g.H0)
            # (adaptive forward runs: the explicit fixed-step loop is unstable with a sliding term)
            fw = dict(method="rdpk3sp35", reltol=1e-6, abstol=1e-6)
            Href = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=1.7 * As[k]), None, t, **fw)]
            Hs = [_r(h, dtype) for h in o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=As[k]), None, t, **fw)]
            assert all(np.isfinite(h).all() for h in Hs)
            for j in range(len(t)):
                ens.set_snapshot(k, j, len(t), Hs[j])
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            ens.set_A_scalar(k, As[k])
            tgs = o.TargetA(ph, "scalar")
            theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
            st = {}
            ell, dth = o.loss_and_grad_continuous_adaptive(theta, g2, tgs, t, Hs, Href, n_quadrature=7, reltol=tol, abstol=tol, vjp="discrete", stats=st)
            refs.append((ell, dth[0], tgs.vjp_theta[0], st))
        ens.set_cluster_mode(0)
        loss, Ssum, steps = ens.grad_continuous_adaptive(t, n_quadrature=7, vjp="discrete", reltol=tol, abstol=tol, max_steps=5000)
        # fp32: solver-tolerance level (with the sliding term the oracle's own gradient moves by 0.4 % between rtol 1e-4 and 1e-8)
        rt_l, rt_g = (1e-10, 1e-7) if dtype == "f64" else (2e-4, 5e-3 if "n" not in phys_kw else 1.5e-2)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert Ssum[k] * refs[k][2] == pytest.approx(refs[k][1], rel=rt_g), (k, Ssum[k] * refs[k][2], refs[k][1])
            if dtype == "f64":
                assert steps[k] == refs[k][3]["steps"], (k, steps[k], refs[k][3])
    finally:
        ens.close()
