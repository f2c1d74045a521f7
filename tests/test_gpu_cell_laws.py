"""Per-cell MLP laws on the GPU (LawU: SIA2D_D_target, LawY: SIA2D_D_hybrid_target) against the NumPy oracle.

Forward F1, discrete A1 (with the reference's finite-difference partials) and the θ-VJP (exact per-node network gradient,
interpolation = :None) -- test/test_grad_loss.jl:245-248 exercises the same targets upstream.
Tolerances: the law is evaluated in fp64 for the partials and the pullback in both precisions; F1 in fp32 evaluates the
network in fp32 (softplus/exp at ~1e-6), hence 5e-5 there; the fp32 A1 inherits the fp32 stencil on top of fp64 partials
stored as fp32 planes (1e-5 x a few).

fp64 A1 bound (FD_TOL = 1e-9): the reference's partials are finite differences of the network with steps 1e-4 / 1e-6
(target_D_pure.jl:105-137), which amplify rounding by 1/step.  Measured on the oracle itself: forming |∇S| in a
mathematically identical but differently ordered way moves the oracle's own VJP_H by 6e-12 .. 4e-11 (relative L2, the
four grids of the ragged test below), so two implementations -- which also use different exp / log1p -- cannot be expected
to agree below ~1e-10; observed CUDA-vs-oracle: <= 1.2e-10.  The quantities without a finite difference are held much
tighter: F1 1e-12, θ-VJP 1e-11."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

FD_TOL = 1e-9  # fp64 A1 with finite-difference partials: see the module docstring


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    return odinn_b200


def _inputs(nx, ny, dtype, seed):
    g = o.rough_bed_glacier(nx, ny)
    lam = np.random.default_rng(seed).standard_normal((nx, ny))
    npdt = np.float32 if dtype == "f32" else np.float64
    g2 = o.Glacier(B=g.B.astype(npdt).astype(np.float64), dx=g.dx, dy=g.dy)
    return g2, g.H0.astype(npdt).astype(np.float64), lam.astype(npdt).astype(np.float64)


def _mlp(widths, acts, seed):
    m = o.MLP(list(widths), list(acts))
    return m, m.init(seed=seed, scale=0.6) + 0.05 * np.random.default_rng(seed).standard_normal(m.n_params)


CASES_U = [
    dict(widths=(2, 3, 1), acts=("softplus", "sigmoid"), bounds=((0.0, 300.0), (0.0, 0.5)), max_NN=50.0),      # light net
    dict(widths=(2, 16, 16, 1), acts=("softplus", "softplus", "sigmoid"), bounds=((0.0, 300.0), (0.0, 0.5)), max_NN=50.0),
    dict(widths=(2, 3, 10, 3, 1), acts=("softplus", "softplus", "softplus", "sigmoid"), bounds=None, max_NN=None),
]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(CASES_U)))
def test_lawU_forward_vjpH_vjptheta(ob, dtype, case):
    c = CASES_U[case]
    nx, ny = 37, 29
    g, H, lam = _inputs(nx, ny, dtype, 5 + case)
    mlp, theta = _mlp(c["widths"], c["acts"], 100 + case)
    if c["bounds"] is None:  # raw inputs: keep the pre-activations in a sane range
        theta = 0.02 * theta
    tg = o.TargetD(o.Phys(), mlp, prescale_bounds=c["bounds"], max_NN=c["max_NN"])
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy)], ob.Phys(), A=1e-17, dtype=dtype)
    tf, tv = (1e-12, FD_TOL) if dtype == "f64" else (5e-5, 1e-4)
    try:
        ens = sim.ensemble
        ens.law_cell_nn_set("U", c["widths"], c["acts"], theta, prescale_bounds=c["bounds"], max_NN=c["max_NN"])
        dH = ens.sia2d_rhs(0, H)
        assert rel_l2(dH, o.SIA2D(H, g, tg, theta)) <= tf
        vH = ens.sia2d_vjp_H(0, lam, H)
        ref = o.VJP_dSIA_dH_discrete(lam, H, g, tg, theta)
        assert rel_l2(vH, ref) <= tv
        dth = ens.sia2d_vjp_theta_cell(0, lam, H)
        refth = o.VJP_dSIA_dtheta_discrete(lam, H, g, tg, theta)
        assert rel_l2(dth, refth) <= (1e-11 if dtype == "f64" else 2e-5)
        # switching the law off restores the A law
        ens.law_cell_clear()
        tgA = o.TargetA(o.Phys(), "const", A=1e-17)
        assert rel_l2(ens.sia2d_rhs(0, H), o.SIA2D(H, g, tgA)) <= (1e-12 if dtype == "f64" else 1e-5)
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_lawY_hybrid(ob, dtype):
    nx, ny = 33, 41
    g, H, lam = _inputs(nx, ny, dtype, 9)
    widths, acts = (2, 3, 10, 3, 1), ("softplus", "softplus", "softplus", "sigmoid")
    mlp, theta = _mlp(widths, acts, 77)
    T = -7.5
    ph = o.Phys()
    tg = o.TargetDHybrid(ph, mlp, T)
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy)], ob.Phys(), A=1e-17, dtype=dtype)
    tf, tv = (1e-12, FD_TOL) if dtype == "f64" else (5e-5, 1e-4)
    try:
        ens = sim.ensemble
        ens.set_temperature(0, T)
        ens.law_cell_nn_set("Y", widths, acts, theta, prescale_bounds=tg.bounds, max_NN=tg.max_NN)
        assert rel_l2(ens.sia2d_rhs(0, H), o.SIA2D(H, g, tg, theta)) <= tf
        assert rel_l2(ens.sia2d_vjp_H(0, lam, H), o.VJP_dSIA_dH_discrete(lam, H, g, tg, theta)) <= tv
        dth = ens.sia2d_vjp_theta_cell(0, lam, H)
        assert rel_l2(dth, o.VJP_dSIA_dtheta_discrete(lam, H, g, tg, theta)) <= (1e-11 if dtype == "f64" else 2e-5)
    finally:
        sim.close()


def test_lawU_ragged_resident_and_determinism(ob):
    """Ensemble-wide resident launch with a per-cell law; θ-gradients per glacier; bit-stable run to run."""
    from odinn_b200 import _capi

    rng = np.random.default_rng(3)
    shapes = [(21, 34), (40, 17), (3, 3), (33, 33)]
    c = CASES_U[1]
    mlp, theta = _mlp(c["widths"], c["acts"], 55)
    tg = o.TargetD(o.Phys(), mlp, prescale_bounds=c["bounds"], max_NN=c["max_NN"])
    gl = [o.rough_bed_glacier(nx, ny) for nx, ny in shapes]
    lams = [rng.standard_normal(s) for s in shapes]
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy) for g in gl], ob.Phys(), A=1e-17, dtype="f64")
    try:
        ens = sim.ensemble
        ens.law_cell_nn_set("U", c["widths"], c["acts"], theta, prescale_bounds=c["bounds"], max_NN=c["max_NN"])
        for k, g in enumerate(gl):
            ens.upload(k, _capi.FIELD_H, g.H0)
            ens.upload(k, _capi.FIELD_LAMBDA, lams[k])
        ens.rhs_resident()
        ens.vjp_resident(True, True, read_S=False)
        G1 = ens.law_cell_grad()
        for k, g in enumerate(gl):
            assert rel_l2(ens.download(k, _capi.FIELD_DH), o.SIA2D(g.H0, g, tg, theta)) <= 1e-12, k
            assert rel_l2(ens.download(k, _capi.FIELD_VJP_H), o.VJP_dSIA_dH_discrete(lams[k], g.H0, g, tg, theta)) <= FD_TOL, k
            ref = o.VJP_dSIA_dtheta_discrete(lams[k], g.H0, g, tg, theta)
            assert rel_l2(G1[k], ref) <= 1e-11 or not ref.any(), k
        ens.vjp_resident(True, True, read_S=False)
        assert np.array_equal(G1, ens.law_cell_grad())
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["U", "Y"])
def test_law_pullback_with_linear_interpolation(ob, dtype, kind):
    """M2 with interpolation = :Linear -- the default of SIA2D_D_hybrid_target (1-D lattice over Hbar, target_D_hybrid.jl:136-166)
    and the optional mode of SIA2D_D_target (2-D lattice over (Hbar, gradS), target_D_pure.jl:180-193, Laws.jl:140-168): the device
    reorders the contraction as a sum over knots; it must equal the oracle's per-cell interpolated gradient tensor contracted with
    D_adj, converge to the exact gradient with the knot count, and fall back to the exact gradient when the knots are cleared."""
    import odinn_b200

    nx, ny = 37, 29
    g, H, lam = _inputs(nx, ny, dtype, 21)
    widths, acts = (2, 16, 16, 1), ("softplus", "softplus", "sigmoid")
    mlp, theta = _mlp(widths, acts, 123)
    ph = o.Phys()
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy)], ob.Phys(), A=1e-17, dtype=dtype)
    tol = 1e-10 if dtype == "f64" else 5e-5
    try:
        ens = sim.ensemble
        if kind == "U":
            bounds, max_NN = ((0.0, 300.0), (0.0, 0.5)), 50.0
            exact = o.TargetD(ph, mlp, prescale_bounds=bounds, max_NN=max_NN)
            f = o._recompute_forward(H, g, exact, theta)
            ens.law_cell_nn_set("U", widths, acts, theta, prescale_bounds=bounds, max_NN=max_NN)
        else:
            T = -7.5
            exact = o.TargetDHybrid(ph, mlp, T)
            f = o._recompute_forward(H, g, exact, theta)
            ens.set_temperature(0, T)
            ens.law_cell_nn_set("Y", widths, acts, theta, prescale_bounds=exact.bounds, max_NN=exact.max_NN)
        ref_exact = o.VJP_dSIA_dtheta_discrete(lam, H, g, exact, theta)
        errs = []
        for nh in (6, 40):
            kh = o.create_interpolation(f["Hb"], nh, dilation_factor=1.05, minA_quantile=10.0 if kind == "U" else None)
            assert np.array_equal(kh, odinn_b200.create_interpolation(f["Hb"], nh, dilation_factor=1.05,
                                                                      minA_quantile=10.0 if kind == "U" else None))
            if kind == "U":
                ks = o.create_interpolation(f["gS"], nh, dilation_factor=1.05)
                tgi = o.TargetD(ph, mlp, prescale_bounds=bounds, max_NN=max_NN, interpolation="Linear", nodes_H=kh, nodes_S=ks)
                ens.law_cell_interp_set(kh, ks)
            else:
                tgi = o.TargetDHybrid(ph, mlp, T, interpolation="Linear", nodes_H=kh)
                ens.law_cell_interp_set(kh)
            ref = o.VJP_dSIA_dtheta_discrete(lam, H, g, tgi, theta)
            dth = ens.sia2d_vjp_theta_cell(0, lam, H)
            assert rel_l2(dth, ref) <= tol, (nh, rel_l2(dth, ref))
            errs.append(rel_l2(dth, ref_exact))
        assert errs[1] < 0.2 * errs[0] and errs[1] < 2e-3   # second-order convergence to the exact gradient
        ens.law_cell_interp_set()
        assert rel_l2(ens.sia2d_vjp_theta_cell(0, lam, H), ref_exact) <= (1e-11 if dtype == "f64" else 2e-5)
        with pytest.raises(ob.OdinnError):
            ens.law_cell_interp_set(np.array([1.0, 0.5]), np.array([0.0, 1.0]) if kind == "U" else None)   # knots must increase
    finally:
        sim.close()
