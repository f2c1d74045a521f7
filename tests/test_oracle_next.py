"""CPU checks of the oracle restatements behind SURVEY 8f N2 / N3: LossV + surface-velocity VJPs and the mass-balance VJP,
each against central finite differences of the corresponding forward (the reference's own protocol for its VJPs,
test/SIA2D_adjoint.jl:139-206)."""
import numpy as np

from oracle import sia2d_numpy as o

PH = o.Phys(minA=8e-21, maxA=8e-17)


def _setup():
    g = o.rough_bed_glacier(20, 17)
    g.H0 = 0.6 * g.H0
    tg = o.TargetA(PH, "scalar")
    th = np.array([-0.3])
    Vxr, Vyr, Var = o.V_from_H(g.H0 * 1.1, g, o.TargetA(PH, "const", A=5e-17))
    return g, tg, th, Vxr, Vyr, Var


def test_lossV_backward_matches_finite_differences():
    g, tg, th, Vxr, Vyr, Var = _setup()
    H, norm = g.H0, float(g.H0.size)
    L = lambda HH, tt: o.loss_V(HH, Var, Vxr, Vyr, g, tg, tt, norm, "xy", True)
    tg.precompute_vjp(th)
    dH, dth = o.backward_loss_V(H, Var, Vxr, Vyr, g, tg, th, norm, "xy", True)
    fd_th = (L(H, th + 1e-5) - L(H, th - 1e-5)) / 2e-5
    assert abs(dth[0] - fd_th) <= 1e-7 * abs(fd_th)
    d = np.random.default_rng(0).standard_normal(H.shape) * (H > 5)
    fd_H = (L(H + 1e-4 * d, th) - L(H - 1e-4 * d, th)) / 2e-4
    assert abs(np.sum(dH * d) - fd_H) <= 1e-6 * abs(fd_H)
    # :abs -- the cotangent written upstream (Losses.jl:369-370) equals the :xy one; kept as is
    dHa, dtha = o.backward_loss_V(H, Var, Vxr, Vyr, g, tg, th, norm, "abs", True)
    assert np.allclose(dHa, dH, rtol=1e-9, atol=1e-18) and np.allclose(dtha, dth, rtol=1e-9)


def test_surface_velocity_shape_and_sign():
    g, tg, th, *_ = _setup()
    Vx, Vy, V = o.V_from_H(g.H0, g, o.TargetA(PH, "const", A=5e-17))
    assert Vx.shape == g.H0.shape and np.all(Vx[-1, :] == 0) and np.all(Vy[:, -1] == 0)  # values live on inn1 (adjoint.jl:330-331)
    f = o._recompute_forward(g.H0, g, o.TargetA(PH, "const", A=5e-17), None)
    assert np.all(o.inn1(Vx) * f["gSx"] <= 0) and np.all(V >= 0)  # ice flows down the surface slope


def test_mass_balance_vjp_matches_finite_differences():
    g, *_ = _setup()
    par = (3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)
    H = g.H0
    MB, mask, gone, PDD = o.mb_TI1(H, g.B, par)
    assert MB.min() < 0 < MB.max() and mask.any()
    rng = np.random.default_rng(0)
    lam = rng.standard_normal(H.shape)
    F = lambda HH: np.sum(lam * o.mb_TI1(HH, g.B, par)[0])
    d = rng.standard_normal(H.shape) * (H > 20)
    fd = (F(H + 1e-3 * d) - F(H - 1e-3 * d)) / 2e-3
    assert abs(np.sum(o.VJP_MB_dH(lam, H, g.B, par) * d) - fd) <= 1e-7 * abs(fd)
    # thin ice that a strongly negative balance removes: MB is clipped to -H and the cotangent is -λ there (VJPs.jl:134-146)
    par_melt = (12.0, -0.0065, 2100.0, 0.0, 3.0, 1.0, 1.0)
    Hthin = np.where(H > 0, 5.0, 0.0)
    MB2, mask2, gone2, _ = o.mb_TI1(Hthin, g.B, par_melt)
    assert gone2.any() and np.all((Hthin + MB2)[gone2] == 0)
    assert np.all(o.VJP_MB_dH(lam, Hthin, g.B, par_melt)[gone2] == -lam[gone2])


def test_loss_weights_of_the_three_loss_types():
    t = np.array([2010.0, 2010.25, 2010.5, 2011.0])
    has_V = [False, True, False, True]
    wH, wV = o.loss_weights("H", t)
    assert np.allclose(wH, [0, .25, .25, .5]) and not wV.any()
    wH, wV = o.loss_weights("V", t, has_V)
    assert not wH.any() and np.allclose(wV, [0, 0, 0, 0.75])  # Δt.V of the first velocity datum is 0 (safe_slice)
    wH, wV = o.loss_weights("HV", t, has_V, scaling=2.0)
    assert np.allclose(wH, np.array([0, .25, .25, .5]) ** 2) and np.allclose(wV, [0, 0, 0, 2 * 0.75**2])
