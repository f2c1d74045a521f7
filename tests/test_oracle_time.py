"""Known-answer and whole-gradient checks of the oracle's time loops.

* Halfar (1983) similarity solution -- the analytical known answer the
  reference sets up at scripts/MWEs/inversion_diffusivity/inversion_setup.jl:44-71.
* Discrete-adjoint gradient (src/inverse/SIA2D/gradient.jl:191-253) against
  central finite differences through the whole forward solve -- the protocol
  of test/test_grad_loss.jl:46-403, including the forward/reverse loss
  equality assert of gradient.jl:259 (rtol 1e-8)."""
import numpy as np
import pytest

from conftest import stats_err_arrays
from oracle import sia2d_numpy as o


def _halfar_err(n, method, **kw):
    dx, A, H0 = 3200.0 / n, 2.21e-18, 400.0
    R0 = 0.4 * n * dx
    g = o.dome_glacier(n, n, dx, H0, A)
    t0 = o.halfar_t0(R0, H0, A)
    tg = o.TargetA(o.Phys(), "const", A=A)
    ts = t0 + np.array([0.0, 1.0, 2.0])
    Hs = o.solve_forward(g.H0, g, tg, None, ts, method=method, **kw)
    xs = (np.arange(n) - n / 2) * dx
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    Ha = o.halfar(X, Y, ts[-1], R0, H0, A)
    return np.linalg.norm(Hs[-1] - Ha) / np.linalg.norm(Ha), Hs


def test_halfar_known_answer_and_convergence():
    e32, Hs = _halfar_err(32, "bs3", reltol=1e-6)
    e64, _ = _halfar_err(64, "bs3", reltol=1e-6)
    assert e64 < 0.02 and e64 < 0.7 * e32, (e32, e64)
    # flux form conserves mass to round-off while the margin stays inside the grid
    assert abs(Hs[-1].sum() / Hs[0].sum() - 1.0) < 1e-12


def test_integrators_agree():
    e_a, Ha = _halfar_err(32, "bs3", reltol=1e-8, abstol=1e-8)
    e_b, Hb = _halfar_err(32, "ssprk3", nsub=200)
    assert np.linalg.norm(Ha[-1] - Hb[-1]) / np.linalg.norm(Ha[-1]) < 1e-6


def _twin(nsteps):
    g = o.rough_bed_glacier(20, 22, dx=50.0)
    ph = o.Phys(minA=8e-21, maxA=8e-17)  # test/inversion_test.jl:59-60
    tst = o.define_callback_steps((2010.0, 2010.25), 0.25 / nsteps)
    true = o.TargetA(ph, "const", A=4e-17)
    kw = dict(method="bs3", reltol=1e-8, abstol=1e-8)
    Href = o.solve_forward(g.H0, g, true, None, tst, **kw)
    mlp = o.MLP.default(1, light=True)
    th = mlp.init(3)
    tn = o.TargetA(ph, "nn", mlp=mlp, T=-10.0)
    fwd = lambda t: o.solve_forward(g.H0, g, tn, t, tst, **kw)
    Hs = fwd(th)
    L0 = o.loss_forward(Hs, Href, tst, g.shape)
    ell, dth, _ = o.loss_and_grad_discrete(th, g, tn, tst, Hs, Href)
    assert ell == pytest.approx(L0, rel=1e-8)  # gradient.jl:259
    num = np.zeros_like(th)
    for k in range(th.size):
        tp, tm = th.copy(), th.copy()
        tp[k] += 1e-4
        tm[k] -= 1e-4
        num[k] = (o.loss_forward(fwd(tp), Href, tst, g.shape) - o.loss_forward(fwd(tm), Href, tst, g.shape)) / 2e-4
    return stats_err_arrays(dth, num)


def test_discrete_adjoint_gradient_vs_fd_first_order():
    """The reverse loop is an explicit-Euler adjoint on the saved steps
    (gradient.jl:242): first-order consistent, so the error halves with Δt."""
    r6 = np.abs(_twin(6))
    r12 = np.abs(_twin(12))
    assert r12[1] < 1e-8  # angle threshold, test/runtests.jl Discrete/Discrete
    assert r12[0] < 5e-2 and r12[2] < 5e-2, r12
    assert r12[2] < 0.65 * r6[2], (r6, r12)


def test_continuous_adjoint_gradient_agrees_with_finite_differences():
    """ContinuousAdjoint restatement (gradient.jl:276-538): with the discrete VJPs inside, the quadrature of the reverse
    solution reproduces the finite-difference gradient of the forward loss (the reference's test protocol for its
    gradient methods, test/test_grad_loss.jl: continuous adjoint vs finite differences within a few per cent)."""
    g = o.rough_bed_glacier(24, 21)
    g.H0 = 0.6 * g.H0
    ph = o.Phys(minA=8e-21, maxA=8e-17)
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    run = lambda tg, th: o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=8)
    Href = run(o.TargetA(ph, "const", A=4e-17), None)
    th = np.array([-0.4])
    tg = o.TargetA(ph, "scalar")
    Hs = run(tg, th)
    L = lambda thv: o.loss_forward(run(tg, thv), Href, t, g.shape)
    fd = (L(th + 1e-4) - L(th - 1e-4)) / 2e-4
    ell, gc = o.loss_and_grad_continuous(th, g, tg, t, Hs, Href, n_quadrature=20, nsub=4, vjp="discrete")
    assert ell == L(th) or abs(ell - L(th)) <= 1e-12 * abs(ell)  # forward / reverse loss equality (gradient.jl:259)
    assert abs(gc[0] - fd) <= 2e-3 * abs(fd), (gc, fd)
    # the differentiate-then-discretise VJP flavour ignores the flux clamp: same sign and magnitude, looser agreement
    _, gcc = o.loss_and_grad_continuous(th, g, tg, t, Hs, Href, n_quadrature=20, nsub=4, vjp="continuous")
    assert abs(gcc[0] - fd) <= 0.1 * abs(fd), (gcc, fd)
    # quadrature nodes and weights: GaussQuadrature maps Gauss-Legendre to the time span (gradient.jl:560-566)
    x, w = o.gauss_quadrature(2010.0, 2015.0, 5)
    assert abs(w.sum() - 5.0) < 1e-12 and abs((w * x**3).sum() - (2015.0**4 - 2010.0**4) / 4) < 1e-3
