"""oracle.fast (the NumPy oracle's loops with the C kernels plugged in) == the pure NumPy oracle.

This is what lets the -m gpu parity tests run at the BASELINE sizes (500 x 500 x 61 snapshots, 64-glacier ensembles):
the loops are the NumPy oracle's own, only F1 / A1 / A2 come from the C restatement."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import fast
from oracle import sia2d_numpy as o

PH = dict(minA=8e-21, maxA=8e-17)


@pytest.mark.parametrize("kind", ["const", "nn", "scalar"])
def test_fast_loops_match_numpy(kind):
    g = o.rough_bed_glacier(37, 29)
    g.H0 = 0.6 * g.H0
    ph = o.Phys(**PH)
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    mlp = o.MLP([1, 16, 16, 1], ["softplus", "softplus", "sigmoid"])
    if kind == "const":
        tg, th = o.TargetA(ph, "const", A=3e-17), None
    elif kind == "nn":
        tg, th = o.TargetA(ph, "nn", mlp=mlp, T=-7.0), mlp.init(5, scale=0.7)
    else:
        tg, th = o.TargetA(ph, "scalar"), np.array([0.2])
    Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="ssprk3", nsub=16)
    Hs = o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=16)
    ell, dth, lam0 = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
    ell_c, dth_c = o.loss_and_grad_continuous(th, g, tg, t, Hs, Href, n_quadrature=5, nsub=2)
    with fast.use_c_kernels(all_cores=False):
        assert o.SIA2D is fast.SIA2D_c
        Hs2 = o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=16)
        ell2, dth2, lam02 = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
        ell_c2, dth_c2 = o.loss_and_grad_continuous(th, g, tg, t, Hs, Href, n_quadrature=5, nsub=2)
        Hb = o.solve_forward(g.H0, g, tg, th, t, method="bs3", reltol=1e-6, abstol=1e-6)
    assert o.SIA2D is not fast.SIA2D_c  # restored
    Hb_np = o.solve_forward(g.H0, g, tg, th, t, method="bs3", reltol=1e-6, abstol=1e-6)
    for a, b in zip(Hs2, Hs):
        assert rel_l2(a, b) < 1e-12
    assert rel_l2(Hb[-1], Hb_np[-1]) < 1e-10
    assert ell2 == pytest.approx(ell, rel=1e-12) and ell_c2 == pytest.approx(ell_c, rel=1e-12)
    assert rel_l2(lam02, lam0) < 1e-11
    if kind != "const":
        assert rel_l2(dth2, dth) < 1e-10 and rel_l2(dth_c2, dth_c) < 1e-10


def test_fast_falls_through_for_other_targets():
    g = o.rough_bed_glacier(19, 17)
    ph = o.Phys()
    Af = 2e-18 * np.ones((18, 16))
    tg = o.TargetA(ph, "gridded")
    th = np.zeros((18, 16))
    lam = np.random.default_rng(0).standard_normal(g.B.shape)
    ref = o.VJP_dSIA_dtheta_discrete(lam, g.H0, g, tg, th)
    with fast.use_c_kernels(all_cores=False):
        got = o.VJP_dSIA_dtheta_discrete(lam, g.H0, g, tg, th)
        tgc = o.TargetA(ph, "const", A=Af)  # gridded A through the C kernel
        assert rel_l2(o.SIA2D(g.H0, g, tgc), fast._np_SIA2D(g.H0, g, tgc)) < 1e-13
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("method", ["euler", "ssprk3"])
def test_c_loops_match_numpy_loops(method):
    """The C time loop / reverse loop (one call each) == the NumPy loops they restate."""
    g = o.rough_bed_glacier(41, 33)
    g.H0 = 0.6 * g.H0
    ph = o.Phys(**PH)
    t = np.array([2010.0, 2010.08, 2010.2, 2010.25])  # non-uniform tstops
    mlp = o.MLP([1, 16, 16, 1], ["softplus", "softplus", "sigmoid"])
    tg, th = o.TargetA(ph, "nn", mlp=mlp, T=-7.0), mlp.init(5, scale=0.7)
    Hs = o.solve_forward(g.H0, g, tg, th, t, method=method, nsub=24)
    Hc = fast.solve_forward_fixed(g.H0, g, tg, th, t, method=method, nsub=24)
    for a, b in zip(Hc, Hs):
        assert rel_l2(a, b) < 1e-12
    Href = [0.95 * h for h in Hs]
    ell, dth, lam0 = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
    ell2, dth2, lam02 = fast.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
    assert ell2 == pytest.approx(ell, rel=1e-12)
    assert rel_l2(dth2, dth) < 1e-10 and rel_l2(lam02, lam0) < 1e-11
