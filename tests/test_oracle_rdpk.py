"""The oracle's RDPK3Sp35 + PID integrator (the reference's default solver, OrdinaryDiffEq -- not in the tree) and the adaptive
continuous adjoint built on it.

Pins available without Julia: (1) the 3S* coefficients reproduce the order conditions of a 3rd-order scheme and the published
abscissae (digit-exact main scheme); the embedded weights sum to 1 and satisfy the first-moment condition to 2.8e-7 (see the
oracle header); (2) tolerance-proportional global errors on a linear problem; (3) the continuous-adjoint gradient against
central finite differences of the loss with the reference's thresholds for ContinuousAdjoint + DiscreteVJP
[1e-3, 1e-8, 1e-3] (test/test_grad_loss.jl, runtests.jl:114-266); (4) MB / velocity callbacks against finite differences."""
import numpy as np
import pytest

from oracle import sia2d_numpy as o

PH = dict(minA=8e-21, maxA=8e-17)


def test_rdpk3sp35_order_conditions_and_abscissae():
    A, b, c = o.rdpk_butcher()
    assert abs(b.sum() - 1) < 5e-16 and abs(b @ c - 0.5) < 5e-16
    assert abs(b @ c**2 - 1.0 / 3.0) < 5e-16 and abs(b @ (A @ c) - 1.0 / 6.0) < 5e-16   # 3rd order
    assert np.abs(c[1:] - np.array(o.RDPK_C)).max() < 5e-16                              # published abscissae
    assert np.allclose(np.triu(A), 0.0)                                                   # explicit
    bh = np.array(o.RDPK_BHAT)
    assert abs(bh.sum() - 1) < 5e-16 and abs(bh @ c - 0.5) < 5e-7                        # embedded: 2nd order (to 2.8e-7)
    assert abs(sum(o.RDPK_E)) < 5e-16


def test_integrator_tolerance_proportionality_and_tstops():
    errs = []
    for tol in (1e-5, 1e-7, 1e-9):
        st = {}
        out = o.integrate_rdpk3sp35(lambda t, u: np.array([-u[0], 2.0 * np.cos(2 * t)]), np.array([1.0, 0.0]), [0.0, 0.7, 1.0, 2.0],
                                    tol, tol, stats=st)
        assert len(out) == 4 and st["nrhs"] == 2 + 5 * st["steps"]
        errs.append(max(abs(out[-1][0] - np.exp(-2.0)), abs(out[-1][1] - np.sin(4.0))))
        assert abs(out[1][0] - np.exp(-0.7)) < 50 * tol   # landed exactly on the tstop
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-7
    # dtmax and the callback (u_modified -> FSAL re-evaluation)
    st = {}
    out = o.integrate_rdpk3sp35(lambda t, u: -u, np.array([1.0]), [0.0, 1.0, 2.0], 1e-3, 1e-3, dtmax=0.05,
                                on_stop=lambda j, t, u: 2.0 * u if j == 1 else None, stats=st)
    assert st["steps"] >= 40 and abs(out[-1][0] - 2.0 * np.exp(-2.0)) < 1e-3


def _setup(n_t=4):
    g = o.rough_bed_glacier(26, 23)
    g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.0 + (n_t - 1) / 12.0), 1.0 / 12.0)
    ph = o.Phys(**PH)
    Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="ssprk3", nsub=12)
    return g, t, ph, Href


def test_adaptive_continuous_adjoint_vs_finite_differences():
    g, t, ph, Href = _setup()
    tgs = o.TargetA(ph, "scalar")
    theta = np.array([np.arctanh(2 * (3e-17 - ph.minA) / (ph.maxA - ph.minA) - 1)])
    fwd = lambda th: o.solve_forward(g.H0, g, tgs, th, t, method="rdpk3sp35", reltol=1e-9, abstol=1e-9)
    Hs = fwd(theta)
    st = {}
    ell, dth = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=40, stats=st)
    assert st["rejected"] <= st["steps"] // 4
    assert ell == pytest.approx(o.loss_forward(Hs, Href, t, g.shape), rel=1e-12)      # gradient.jl:259
    e = 1e-5
    fd = (o.loss_forward(fwd(theta + e), Href, t, g.shape) - o.loss_forward(fwd(theta - e), Href, t, g.shape)) / (2 * e)
    assert abs(dth[0] / fd - 1.0) < 5e-3, (dth, fd)   # Continuous adjoint on linearly interpolated snapshots (reference: 1e-3 with 200 nodes)
    # the fixed-step reverse solve of the same branch converges to the same gradient
    _, dth_fixed = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=40, fixed=("ssprk3", 6))
    assert dth_fixed[0] == pytest.approx(dth[0], rel=1e-5)
    # and it reproduces the older entry point exactly (same scheme, no callbacks besides the loss)
    _, dth_old = o.loss_and_grad_continuous(theta, g, tgs, t, Hs, Href, n_quadrature=40, nsub=6)
    assert dth_fixed[0] == pytest.approx(dth_old[0], rel=1e-12)


MB_PAR = (3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)


def test_adaptive_continuous_adjoint_with_mass_balance_vs_finite_differences():
    """LossH with the mass-balance callback: forward MB steps + the PeriodicCallback of the reverse solve (gradient.jl:407-426)."""
    g, t, ph, _ = _setup(n_t=5)
    mb = {2: MB_PAR, 4: MB_PAR}
    tgs = o.TargetA(ph, "scalar")
    theta = np.array([np.arctanh(2 * (2e-17 - ph.minA) / (ph.maxA - ph.minA) - 1)])
    Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="ssprk3", nsub=12, mb=mb)

    def run(th):
        st = {}
        return o.solve_forward(g.H0, g, tgs, th, t, method="rdpk3sp35", reltol=1e-9, abstol=1e-9, mb=mb, stats=st), st["MB"]

    Hs, MBh = run(theta)
    ell, dth = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=60, mb=mb, MB_hist=MBh)
    _, dth_noMB = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=60)
    assert abs(dth[0] - dth_noMB[0]) > 1e-6 * abs(dth[0])        # the MB callback matters here
    e = 1e-5
    fd = (o.loss_forward(run(theta + e)[0], Href, t, g.shape) - o.loss_forward(run(theta - e)[0], Href, t, g.shape)) / (2 * e)
    assert ell == pytest.approx(o.loss_forward(Hs, Href, t, g.shape), rel=1e-12)
    assert abs(dth[0] / fd - 1.0) < 1e-2, (dth, fd, dth_noMB)


def test_adaptive_continuous_adjoint_with_velocity_loss():
    """LossV inside the continuous adjoint (gradient.jl:289-366, 474-507): the H-mediated part (lambda driven by the velocity loss
    jumps) agrees with the discrete adjoint's; the direct part is the Gauss quadrature of dl_V/dtheta over interpolated references
    -- by construction NOT the tstop sum the loss value uses (the first datum has weight 0 there, the quadrature covers the whole
    span), so it is checked against an independent evaluation of that quadrature."""
    g, t, ph, _ = _setup(n_t=5)
    tgs = o.TargetA(ph, "scalar")
    theta = np.array([np.arctanh(2 * (2e-17 - ph.minA) / (ph.maxA - ph.minA) - 1)])
    tref = o.TargetA(ph, "const", A=5e-17)
    Href = o.solve_forward(g.H0, g, tref, None, t, method="ssprk3", nsub=12)
    has_V = [False, True, False, True, True]
    Vref = [o.V_from_H(Href[j], g, tref) if has_V[j] else None for j in range(len(t))]
    wH, wV = o.loss_weights("V", t, has_V)   # LossV: w_j = Delta t_V, quadrature multiplier 1
    Hs = o.solve_forward(g.H0, g, tgs, theta, t, method="rdpk3sp35", reltol=1e-9, abstol=1e-9)
    nq = 40
    ell, dth = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=nq, wH=wH, wV=wV, V_ref=Vref, cV=1.0)
    _, dth_lam = o.loss_and_grad_continuous_adaptive(theta, g, tgs, t, Hs, Href, n_quadrature=nq, wH=wH, wV=wV, V_ref=Vref, cV=0.0)
    elld, dthd = o.loss_and_grad_discrete_HV(theta, g, tgs, t, Hs, Href, Vref, wH, wV)
    assert ell == pytest.approx(elld, rel=1e-12)
    N = float(np.prod(g.shape))
    direct_disc = sum(wV[j] * o.backward_loss_V(Hs[j], Vref[j][2], Vref[j][0], Vref[j][1], g, tgs, theta, N)[1][0] for j in (1, 3, 4))
    assert dth_lam[0] == pytest.approx(dthd[0] - direct_disc, rel=3e-2)
    # independent quadrature of the direct term: linear interpolation of H and of the references (flat before the first datum)
    qn, qw = o.gauss_quadrature(t[0], t[-1], nq)
    tv = t[[1, 3, 4]]
    direct = 0.0
    for tq, w in zip(qn, qw):
        j = min(int(np.searchsorted(t, tq, side="right")) - 1, len(t) - 2)
        a = (tq - t[j]) / (t[j + 1] - t[j])
        Ht = (1 - a) * Hs[j] + a * Hs[j + 1]
        m = int(np.clip(np.searchsorted(tv, tq, side="right") - 1, 0, 1))
        b = float(np.clip((tq - tv[m]) / (tv[m + 1] - tv[m]), 0.0, 1.0))
        ja, jb = [1, 3, 4][m], [1, 3, 4][m + 1]
        V = [(1 - b) * Vref[ja][c] + b * Vref[jb][c] for c in range(3)]
        direct += w * o.backward_loss_V(Ht, V[2], V[0], V[1], g, tgs, theta, N)[1][0]
    assert dth[0] - dth_lam[0] == pytest.approx(direct, rel=1e-10)
