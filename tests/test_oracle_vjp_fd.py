"""Pins the oracle's forward against the reference's in-tree hand adjoint.

The discrete VJPs are transcribed verbatim from src/inverse/SIA2D/adjoint.jl
(:31-255).  Upstream, test/SIA2D_adjoint.jl:139-206 requires them to agree with
finite differences of the real Huginn.SIA2D! to [5e-7, 1e-6, 5e-4]
(test/runtests.jl:89-91).  Here the same protocol (v = randn seed 1234,
minimum over eps, metrics of test/test_utils.jl:78-83) is run against the
RESTATED forward, which is what makes the restatement the operator the
reference adjoint differentiates."""
import numpy as np
import pytest

from conftest import stats_err_arrays
from oracle import sia2d_numpy as o

THRES = (5e-7, 1e-6, 5e-4)  # test/runtests.jl:89-91 (DiscreteVJP, C = 0)
THRES_C = (3e-4, 2e-4, 2e-2)  # test/runtests.jl:94 (C = 7e-8)


def _fd_H(loss, H, eps):
    num = np.zeros_like(H)
    for i in range(H.shape[0]):
        for j in range(H.shape[1]):
            if H[i, j] > 0:  # central FD is invalid exactly at the H=0 kink
                Hp, Hm = H.copy(), H.copy()
                Hp[i, j] += eps
                Hm[i, j] -= eps
                num[i, j] = (loss(Hp) - loss(Hm)) / (2 * eps)
    return num


def _fd_theta(loss, th, eps):
    num = np.zeros_like(th)
    for k in range(th.size):
        tp, tm = th.copy(), th.copy()
        tp[k] += eps
        tm[k] -= eps
        num[k] = (loss(tp) - loss(tm)) / (2 * eps)
    return num


def _min_stats(analytic, fd_fn, epss):
    s = np.array([np.abs(stats_err_arrays(analytic, fd_fn(e))) for e in epss])
    return s.min(axis=0)


def _tilted_dome(nx, ny):
    """Dome on a tilted bed.  On a FLAT bed with eta0 = 1 every margin edge sits
    exactly on the clamp bound (dS == -eta0*H/dx), where the reference's strict
    inequalities (inversion_utils.jl:24-28) return a zero sub-gradient while FD
    sees the one-sided slope; the tilt removes those ties (see test_clamp_tie)."""
    g = o.dome_glacier(nx, ny)
    X = np.arange(nx)[:, None] * g.dx
    Y = np.arange(ny)[None, :] * g.dy
    g.B = 0.03 * X + 0.011 * Y
    return g


@pytest.mark.parametrize("maker", [o.rough_bed_glacier, _tilted_dome])
@pytest.mark.parametrize("C", [0.0, 7e-8])
def test_vjp_H_matches_fd_of_forward(maker, C):
    g = maker(14, 17)
    ph = o.Phys(C=C)
    tg = o.TargetA(ph, "const", A=2.21e-18)
    H = g.H0.copy()
    lam = np.random.default_rng(1234).standard_normal(H.shape)
    dl = o.VJP_dSIA_dH_discrete(lam, H, g, tg)
    m = H > 0
    assert np.all(dl[~m] == 0.0)  # adjoint.jl:148
    loss = lambda Hx: np.sum(o.SIA2D(Hx, g, tg) * lam)
    r = _min_stats(dl[m], lambda e: _fd_H(loss, H, e)[m], [1e-3, 1e-5, 1e-7])
    thr = THRES if C == 0.0 else THRES_C
    assert r[0] < thr[0] and r[1] < thr[1] and r[2] < thr[2], r


@pytest.mark.parametrize("kind", ["nn", "scalar", "gridded"])
def test_vjp_theta_matches_fd_of_forward(kind):
    g = o.rough_bed_glacier(14, 17)
    ph = o.Phys(minA=8e-21, maxA=8e-17)  # test/SIA2D_adjoint.jl:44-46
    H = g.H0.copy()
    lam = np.random.default_rng(1234).standard_normal(H.shape)
    if kind == "nn":
        mlp = o.MLP.default(1, light=True)  # test_mode architecture, ML_utils.jl:26-29
        th = mlp.init(3)
        tg = o.TargetA(ph, "nn", mlp=mlp, T=-10.0)
    elif kind == "scalar":
        th = np.array([0.3])
        tg = o.TargetA(ph, "scalar")
    else:
        th = 0.5 * np.random.default_rng(7).standard_normal((13, 16))
        tg = o.TargetA(ph, "gridded")
    dth = o.VJP_dSIA_dtheta_discrete(lam, H, g, tg, th)

    def loss(t):
        tg.vjp_theta = None
        return np.sum(o.SIA2D(H, g, tg, t.reshape(th.shape)) * lam)

    flat = th.reshape(-1, order="F").copy()
    loss_flat = lambda t: loss(t.reshape(th.shape, order="F"))
    r = _min_stats(dth, lambda e: _fd_theta(loss_flat, flat, e), [10.0**-k for k in range(3, 8)])
    assert r[0] < THRES[0] and r[1] < THRES[1] and r[2] < THRES[2], r


def test_clamp_tie_semantics():
    """Flat bed, eta0 = 1: margin edges tie with the clamp bound.  The reference
    forward clamps to the bound and the reference adjoint passes NOTHING through
    a tied edge (strict <, > at inversion_utils.jl:24-28).  The GPU kernels must
    reproduce exactly this, so the oracle's behaviour is frozen here."""
    H = np.zeros((6, 5))
    H[2, 2] = 10.0
    dx = 2.0
    dS = o.diff_x(H[:, 1:-1]) / dx
    c = o.clamp_borders_dx(dS, H, 1.0, dx)
    assert np.array_equal(c, dS)  # bound == value on both sides of the ice cell
    ddS, dH = np.zeros_like(dS), np.zeros_like(H)
    o.clamp_borders_dx_adjoint(ddS, dH, np.ones_like(dS), 1.0, dx, H, dS)
    assert ddS[1, 1] == 0.0 and ddS[2, 1] == 0.0 and np.all(dH == 0.0)
    assert ddS[0, 1] == 0.0  # 0 < 0 is false as well: ice-free edges pass nothing


def test_dense_tensor_contraction_equals_factorised():
    """cartesian_tensor + Tullio (target_utils.jl:156-162, adjoint.jl:250) == scalar reduction x vjp."""
    g = o.rough_bed_glacier(12, 11)
    mlp = o.MLP.default(1)
    th = mlp.init(1)
    tg = o.TargetA(o.Phys(), "nn", mlp=mlp, T=-5.0)
    lam = np.random.default_rng(0).standard_normal(g.shape)
    a = o.VJP_dSIA_dtheta_discrete(lam, g.H0, g, tg, th, dense=True)
    b = o.VJP_dSIA_dtheta_discrete(lam, g.H0, g, tg, th, dense=False)
    np.testing.assert_allclose(a, b, rtol=1e-13)
    s = o.node_reduction_S(lam, g.H0, g, tg, th)
    np.testing.assert_allclose(b, tg.vjp_theta * s, rtol=1e-13)


def test_forward_properties():
    """Zero border, exact mass conservation of the flux form, no mutation of H, H<0 clipped."""
    g = o.rough_bed_glacier(20, 23)
    tg = o.TargetA(o.Phys(), "const", A=2.21e-18)
    H = g.H0.copy()
    H[5, 5] = -3.0
    H_in = H.copy()
    dH = o.SIA2D(H, g, tg)
    assert np.array_equal(H, H_in)
    assert np.all(dH[0, :] == 0) and np.all(dH[-1, :] == 0) and np.all(dH[:, 0] == 0) and np.all(dH[:, -1] == 0)
    assert abs(dH.sum()) < 1e-9 * np.abs(dH).sum()
    H2 = H.copy()
    H2[5, 5] = 0.0
    assert np.array_equal(dH, o.SIA2D(H2, g, tg))


def test_mlp_backward_matches_fd():
    mlp = o.MLP([2, 16, 16, 1], ["softplus", "softplus", "sigmoid"])
    th = mlp.init(5)
    X = np.random.default_rng(1).standard_normal((7, 2))
    gout = np.random.default_rng(2).standard_normal((7, 1))
    dth, dX = mlp.backward(th, X, gout)
    f = lambda t: np.sum(mlp.forward(t, X) * gout)
    num = _fd_theta(f, th, 1e-6)
    np.testing.assert_allclose(dth.sum(axis=0), num, rtol=1e-6, atol=1e-9)
    numX = np.zeros_like(X)
    for idx in np.ndindex(X.shape):
        Xp, Xm = X.copy(), X.copy()
        Xp[idx] += 1e-6
        Xm[idx] -= 1e-6
        numX[idx] = (np.sum(mlp.forward(th, Xp) * gout) - np.sum(mlp.forward(th, Xm) * gout)) / 2e-6
    np.testing.assert_allclose(dX, numX, rtol=1e-6, atol=1e-9)
    assert mlp.n_params == 337 and o.MLP.default(1).n_params == 83


@pytest.mark.parametrize("which", ["D", "D_hybrid"])
def test_percell_targets_theta_vjp_fd(which):
    """Per-cell MLP laws (LawU / LawY): exact θ-VJP (interpolation=:None) vs FD of the forward."""
    g = o.rough_bed_glacier(12, 13)
    ph = o.Phys()
    if which == "D":
        mlp = o.MLP.default(2, light=True)
        tg = o.TargetD(ph, mlp, prescale_bounds=[(0.0, 300.0), (0.0, 0.5)], max_NN=50.0)
    else:
        mlp = o.MLP.default(2, light=True)
        tg = o.TargetDHybrid(ph, mlp, T=-8.0)
    th = mlp.init(11, scale=0.8)
    H = g.H0.copy()
    lam = np.random.default_rng(1234).standard_normal(H.shape)
    dth = o.VJP_dSIA_dtheta_discrete(lam, H, g, tg, th)
    loss = lambda t: np.sum(o.SIA2D(H, g, tg, t) * lam)
    r = _min_stats(dth, lambda e: _fd_theta(loss, th, e), [1e-4, 1e-5, 1e-6])
    assert r[0] < 2e-4 and r[1] < 2e-4 and r[2] < 2e-2, r  # default thresholds, test/SIA2D_adjoint.jl:3


def test_continuous_vjp_is_consistent_without_clamp():
    """On a thick smooth dome the flux clamp is inactive, so the continuous VJP
    (adjoint.jl:442-555) approximates the discrete one in the interior."""
    g = o.dome_glacier(40, 40)
    tg = o.TargetA(o.Phys(), "const", A=2.21e-18)
    X, Y = np.meshgrid(np.arange(40.0), np.arange(40.0), indexing="ij")
    lam = np.sin(X / 7.0) * np.cos(Y / 9.0)
    a = o.VJP_dSIA_dH_discrete(lam, g.H0, g, tg)
    b = o.VJP_dSIA_dH_continuous(lam, g.H0, g, tg)
    core = (slice(12, 28), slice(12, 28))
    rel = np.linalg.norm(a[core] - b[core]) / np.linalg.norm(a[core])
    assert rel < 0.2, rel
