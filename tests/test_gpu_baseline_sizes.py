"""GPU parity AT THE SIZES BASELINE.json NAMES -- the CUDA path (through the C ABI) against the oracle, not against itself.

  config 2   500 x 500 (and the odd neighbours 499 x 501, 501 x 487): F1, A1, A2 and the fused F1+A1+A2 kernel per call;
             then the 61-snapshot forward solve + discrete adjoint at 500 x 500
  config 1   one 128 x 128 glacier, forward 2010 -> 2015 (Halfar dome and rough bed), every saved state
  config 3   64 glaciers, sizes U{100..400} (default_rng(2024)), forward 2010 -> 2015
  config 4   32 glaciers (default_rng(2025)), LawA 1 -> 16 -> 16 -> 1, loss + d(theta) of one SIA2D_grad! iteration;
             per-cell LawU 2 -> 16 -> 16 -> 1 variant on a bounded sample of the same ensemble

The oracle is the NumPy oracle's loops with the C restatement as F1 / A1 / A2 (oracle.fast; the C kernels are pinned to
the NumPy ones at 1e-13 on CPU, tests/test_oracle_c.py, tests/test_oracle_fast.py).  Tolerances: per call fp64 1e-12 /
fp32 1e-5 (the reference's operator rtol, test/SIA2D_adjoint_utils.jl:22); time loops fp64 state 1e-10, loss 1e-10,
d(theta) 1e-8 / fp32 state 1e-3, loss and d(theta) 2e-3; forward / reverse loss equality rtol 1e-8
(src/inverse/SIA2D/gradient.jl:259).  The VJP-vs-finite-difference thresholds [5e-7, 1e-6, 5e-4] of the reference
(test/SIA2D_adjoint.jl:139-206, test/runtests.jl:89-91) are applied to the CUDA VJP at 500 x 500 too."""
import numpy as np
import pytest

from conftest import rel_l2, stats_err_arrays
from oracle import fast
from oracle import sia2d_c as oc
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
A0 = 2.21e-18  # test/test_grad_loss.jl:157
PH = dict(minA=8e-21, maxA=8e-17)  # test/inversion_test.jl:59-60


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    oc.use_all_cores()
    return odinn_b200


def _noisy(nx, ny):
    g = o.rough_bed_glacier(nx, ny)
    rng = np.random.default_rng(7)
    H = g.H0 * (1.0 + 0.3 * rng.standard_normal(g.H0.shape))
    H[rng.random(H.shape) < 0.05] = 0.0
    H[rng.random(H.shape) < 0.02] = -3.0
    g.H0 = H
    return g


def _r(a, dtype):
    return np.asarray(a).astype(np.float32).astype(np.float64) if dtype == "f32" else np.asarray(a, dtype=np.float64)


def _S_close(S, ref, mag, dtype):
    return abs(S - ref) <= 10 * TOL[dtype] * abs(ref) or abs(S - ref) <= (3e-7 if dtype == "f32" else 1e-14) * mag


# ------------------------------------------------------------------------------------------------------------------
# config 2, per call
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tma_rows", [("f64", None), ("f32", None), ("f32", "100")])
@pytest.mark.parametrize("maker", ["rough", "noisy"])
def test_config2_per_call_operators_vs_oracle(ob, dtype, maker, tma_rows, monkeypatch):
    """500 x 500 = 9 strips x 8 row chunks of the marching kernels: every strip / chunk seam is compared with the oracle.
    tma_rows: the row chunks of the fused step's TMA bands (62 for a small ensemble like this one, ~100 for the bench ensemble)."""
    from odinn_b200 import _capi

    if tma_rows:
        monkeypatch.setenv("ODINN_TMA_CHUNK_ROWS", tma_rows)

    shapes = [(500, 500), (499, 501), (501, 487)]
    gl = [(o.rough_bed_glacier if maker == "rough" else _noisy)(nx, ny) for nx, ny in shapes]
    rng = np.random.default_rng(1234)
    ph = o.Phys()
    Bs = [_r(g.B, dtype) for g in gl]
    Hs = [_r(g.H0, dtype) for g in gl]
    lams = [_r(rng.standard_normal(g.B.shape), dtype) for g in gl]
    As = [A0, 2 * A0, 0.5 * A0]
    sim = ob.Simulation([ob.Glacier2D(B=B, Δx=g.dx, Δy=g.dy) for B, g in zip(Bs, gl)], ob.Phys(), A=As, dtype=dtype)
    try:
        ens = sim.ensemble
        ref = []
        for k, g in enumerate(gl):
            dH = oc.rhs(Hs[k], Bs[k], g.dx, g.dy, ph, As[k])
            vH, S, _ = oc.vjp(lams[k], Hs[k], Bs[k], g.dx, g.dy, ph, As[k])
            gk = o.Glacier(B=Bs[k], dx=g.dx, dy=g.dy)
            mag = o.node_reduction_S_terms(lams[k], Hs[k], gk, o.TargetA(ph, "const", A=As[k]))[1] if k == 0 else None
            ref.append((dH, vH, S, mag))
            ens.upload(k, _capi.FIELD_H, Hs[k])
            ens.upload(k, _capi.FIELD_LAMBDA, lams[k])
        # separate kernels
        ens.rhs_resident()
        S_sep = ens.vjp_resident(True, True)
        for k in range(3):
            assert rel_l2(ens.download(k, _capi.FIELD_DH), ref[k][0]) <= TOL[dtype], ("F1", k)
            assert rel_l2(ens.download(k, _capi.FIELD_VJP_H), ref[k][1]) <= TOL[dtype], ("A1", k)
        # fused F1 + A1 + A2 (the bench step): overwrite the outputs first so that stale planes cannot pass
        for k, g in enumerate(gl):
            ens.upload(k, _capi.FIELD_DH, np.full(g.B.shape, 7.0))
            ens.upload(k, _capi.FIELD_VJP_H, np.full(g.B.shape, -7.0))
        S_f = ens.vjp_resident(True, True, want_dH=True)
        for k, g in enumerate(gl):
            dH, vH = ens.download(k, _capi.FIELD_DH), ens.download(k, _capi.FIELD_VJP_H)
            assert rel_l2(dH, ref[k][0]) <= TOL[dtype], ("fused F1", k, rel_l2(dH, ref[k][0]))
            assert rel_l2(vH, ref[k][1]) <= TOL[dtype], ("fused A1", k, rel_l2(vH, ref[k][1]))
            assert np.all(vH[np.maximum(Hs[k], 0) <= 0] == 0)  # adjoint.jl:148
            assert not dH[0, :].any() and not dH[-1, :].any() and not dH[:, 0].any() and not dH[:, -1].any()
            mag = ref[k][3] if ref[k][3] is not None else 1e4 * abs(ref[k][2])
            assert _S_close(S_sep[k], ref[k][2], mag, dtype), ("A2", k, S_sep[k], ref[k][2])
            assert _S_close(S_f[k], ref[k][2], mag, dtype), ("fused A2", k, S_f[k], ref[k][2])
        # the reference-facing per-call entry points (host buffers) on the 500 x 500 glacier
        dH = np.zeros(shapes[0], order="F")
        ob.SIA2D_(dH, Hs[0], sim, 0.0)
        assert rel_l2(dH, ref[0][0]) <= TOL[dtype]
        vH, _ = ob.VJP_λ_dSIAdH(ob.B200VJP(), lams[0], Hs[0], None, sim, 0.0)
        assert rel_l2(vH, ref[0][1]) <= TOL[dtype]
    finally:
        sim.close()


def test_config2_vjp_vs_finite_differences_reference_thresholds(ob):
    """The reference's own acceptance test of A1 / A2 (test/SIA2D_adjoint.jl:139-206): VJP against forward differences of
    the forward operator, v = randn (seed 1234), eps in 1e-3 .. 1e-7, minimum over eps of (ratio - 1, cosine - 1, relative
    error) below [5e-7, 1e-6, 5e-4] -- here with the CUDA forward and the CUDA VJP at 500 x 500, fp64."""
    nx = ny = 500
    g = o.rough_bed_glacier(nx, ny)
    rng = np.random.default_rng(1234)
    v = rng.standard_normal((nx, ny))
    ph = o.Phys()
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy)], ob.Phys(), A=A0, dtype="f64")
    try:
        ens = sim.ensemble
        H = g.H0
        ice = H > 1.0  # directional derivatives away from the H = 0 kink
        d0 = ens.sia2d_rhs(0, H)
        vH = ens.sia2d_vjp_H(0, v, H)
        # dH-VJP: compare <v, J e_k> assembled from a directional sweep: J^T v against FD of sum(dH .* v) per direction
        dirs = [rng.standard_normal((nx, ny)) * ice for _ in range(3)]
        best = [np.inf, np.inf, np.inf]
        for eps in (1e-3, 1e-4, 1e-5, 1e-6, 1e-7):
            fd = np.array([np.sum((ens.sia2d_rhs(0, H + eps * d) - d0) * v) / eps for d in dirs])
            an = np.array([np.sum(vH * d) for d in dirs])
            r, c, e = stats_err_arrays(an, fd)
            best = [min(best[0], abs(r)), min(best[1], abs(c)), min(best[2], e)]
        assert best[0] < 5e-7 and best[1] < 1e-6 and best[2] < 5e-4, best
        # theta-VJP (glacier-wide A): d/dA sum(dH .* v) = S exactly (dH is linear in A when C = 0)
        S = ens.sia2d_vjp_theta(0, v, H)
        ens.set_A_scalar(0, 2 * A0)
        d1 = ens.sia2d_rhs(0, H)
        fdS = np.sum((d1 - d0) * v) / A0
        assert abs(S / fdS - 1.0) < 5e-7, (S, fdS)
        assert S == pytest.approx(fast.S_of(v, H, g, o.TargetA(ph, "const", A=A0)), rel=1e-11)
    finally:
        sim.close()


# ------------------------------------------------------------------------------------------------------------------
# config 2, time loop: 61 snapshots forward + discrete adjoint at 500 x 500
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_config2_61_snapshot_forward_and_discrete_adjoint(ob, dtype):
    from odinn_b200 import _capi

    nx = ny = 500
    g0 = o.rough_bed_glacier(nx, ny)
    # 0.6 x the 250 m cap: the reference's reverse loop is explicit Euler with the MONTHLY step (gradient.jl:242), stable only while
    # dt * rho(dSIA/dH) < 2.  On the full-thickness cap it is not (|lambda| reaches 1e28 and a 1e-13 perturbation of the snapshots
    # flips the sign of d(theta) in the fp64 oracle -- the situation gradient.jl:19-24 warns about), so no two implementations agree there.
    g = o.Glacier(B=_r(g0.B, dtype), dx=g0.dx, dy=g0.dy, H0=_r(0.6 * g0.H0, dtype))
    t = np.linspace(2010.0, 2015.0, 61)
    ph = o.Phys(**PH)
    A_true, A_inv, nsub = 4.0e-18, A0, 8
    tgs = o.TargetA(ph, "scalar")
    theta = np.array([np.arctanh(2 * (A_inv - ph.minA) / (ph.maxA - ph.minA) - 1)])
    Href = fast.solve_forward_fixed(g.H0, g, o.TargetA(ph, "const", A=A_true), None, t, method="ssprk3", nsub=nsub)
    Href = [_r(h, dtype) for h in Href]
    Hs = fast.solve_forward_fixed(g.H0, g, tgs, theta, t, method="ssprk3", nsub=nsub)
    ell, dth, lam0 = fast.loss_and_grad_discrete(theta, g, tgs, t, Hs, Href)
    tgs.precompute_vjp(theta)
    ens = ob.Ensemble([nx], [ny], [g.dx], [g.dy], ob.Phys(**PH), dtype)
    try:
        ens.upload(0, _capi.FIELD_B, g.B)
        ens.upload(0, _capi.FIELD_H0, g.H0)
        ens.set_A_scalar(0, float(tgs.A))
        for j in range(len(t)):
            ens.set_reference(0, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
        ens.solve_forward(t, method="ssprk3", nsub=nsub)
        st, lt, gt = (1e-10, 1e-10, 1e-8) if dtype == "f64" else (1e-3, 2e-3, 2e-3)
        for j in (1, 30, 60):
            assert rel_l2(ens.get_snapshot(0, j), Hs[j]) <= st, (j, rel_l2(ens.get_snapshot(0, j), Hs[j]))
        fwd_loss = ens.loss(t)
        loss, Ssum = ens.grad_discrete(t)
        assert loss[0] == pytest.approx(ell, rel=lt)
        assert fwd_loss[0] == pytest.approx(loss[0], rel=1e-8 if dtype == "f64" else 1e-5)  # gradient.jl:259
        assert Ssum[0] * tgs.vjp_theta[0] == pytest.approx(dth[0], rel=gt), (Ssum[0] * tgs.vjp_theta[0], dth[0])
        lam_dev = ens.download(0, _capi.FIELD_LAMBDA)
        if dtype == "f64":  # the adjoint state after the whole reverse loop (lambda at t_0)
            assert rel_l2(lam_dev, lam0) <= 1e-8
    finally:
        ens.close()


# ------------------------------------------------------------------------------------------------------------------
# config 1: one 128 x 128 glacier, 2010 -> 2015
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", ["halfar", "rough"])
@pytest.mark.parametrize("method", ["ssprk3", "bs3"])
def test_config1_forward_2010_2015(ob, dtype, case, method):
    from odinn_b200 import _capi

    n = 128
    g0 = o.dome_glacier(n, n) if case == "halfar" else o.rough_bed_glacier(n, n)
    g = o.Glacier(B=_r(g0.B, dtype), dx=g0.dx, dy=g0.dy, H0=_r(g0.H0, dtype))
    t = o.define_callback_steps((2010.0, 2015.0), 1.0 / 12.0)
    assert len(t) == 61
    ph = o.Phys()
    tg = o.TargetA(ph, "const", A=A0)
    rtol = 1e-6 if dtype == "f64" else 1e-4
    if method == "ssprk3":
        Hs = fast.solve_forward_fixed(g.H0, g, tg, None, t, method="ssprk3", nsub=8)
    else:
        with fast.use_c_kernels():
            Hs = o.solve_forward(g.H0, g, tg, None, t, method="bs3", reltol=rtol, abstol=rtol)
    ens = ob.Ensemble([n], [n], [g.dx], [g.dy], ob.Phys(), dtype)
    try:
        ens.upload(0, _capi.FIELD_B, g.B)
        ens.upload(0, _capi.FIELD_H0, g.H0)
        ens.set_A_scalar(0, A0)
        if method == "ssprk3":
            ens.solve_forward(t, method="ssprk3", nsub=8)
        else:
            ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol)
        # adaptive: every step size is a function of the error norm, so the RHS roundings (the cubic kernel groups the node products
        # differently from the oracle) feed back into dt, and over 5 years ONE accept / reject decision flips: from then on the two
        # runs are two different discretisations of the same solve and agree at the level of the solver tolerance (rtol 1e-6: 3e-6
        # measured), not of rounding.  The first half year -- before any flip -- agrees to 1e-9.
        st = 1e-10 if method == "ssprk3" else 2e-5
        if dtype == "f32":
            st = 1e-3 if method == "ssprk3" else 3e-3
        if method == "bs3" and dtype == "f64":
            assert rel_l2(ens.get_snapshot(0, 6), Hs[6]) <= 1e-9
        for j in range(0, 61, 6):
            err = rel_l2(ens.get_snapshot(0, j), Hs[j])
            assert err <= st, (j, err)
        Hend = ens.get_snapshot(0, 60).astype(np.float64)
        assert abs(Hend.sum() / g.H0.sum() - 1.0) < (1e-9 if dtype == "f64" else 1e-4)  # mass conservation (ice stays inside)
        if case == "halfar":  # known answer: the similarity solution 5 years later (discretisation error of a 128^2 grid)
            R0, H0 = 0.4 * n * g.dx, 400.0
            xs = (np.arange(n) - n / 2) * g.dx
            X, Y = np.meshgrid(xs, xs, indexing="ij")
            exact = o.halfar(X, Y, o.halfar_t0(R0, H0, A0) + 5.0, R0, H0, A0)
            assert np.abs(Hend - exact).max() < 0.05 * H0   # (15.5 m at the margin in the fp64 oracle)
            assert rel_l2(Hend, exact) < 8e-3               # (5.8e-3 in the fp64 oracle)
    finally:
        ens.close()


# ------------------------------------------------------------------------------------------------------------------
# configs 3 and 4: the ragged ensembles
# ------------------------------------------------------------------------------------------------------------------
def _ensemble_glaciers(seed, G, dtype):
    """sizes nx, ny ~ U{100..400}; dome radius proportional to the size (bench.synthetic_glacier, thinned so that
    the fixed 1/96 yr sub-step is stable for every A of the config)."""
    from bench import synthetic_glacier

    rng = np.random.default_rng(seed)
    shapes = [(int(rng.integers(100, 401)), int(rng.integers(100, 401))) for _ in range(G)]
    gl = []
    for k, (nx, ny) in enumerate(shapes):
        B, H, dx = synthetic_glacier(nx, ny, k)
        gl.append(o.Glacier(B=_r(B, dtype), dx=dx, dy=dx, H0=_r(0.4 * H, dtype)))
    return rng, gl


def _make_ens(ob, gl, dtype):
    from odinn_b200 import _capi

    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl],
                      ob.Phys(**PH), dtype)
    for k, g in enumerate(gl):
        ens.upload(k, _capi.FIELD_B, g.B)
        ens.upload(k, _capi.FIELD_H0, g.H0)
    return ens


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_config3_64_glacier_forward_ensemble(ob, dtype):
    rng, gl = _ensemble_glaciers(2024, 64, dtype)
    As = np.exp(rng.uniform(np.log(2e-18), np.log(8e-17), size=64))  # A_g ~ logU(2e-18, 8e-17)
    t = np.linspace(2010.0, 2015.0, 61)
    ph = o.Phys(**PH)
    ens = _make_ens(ob, gl, dtype)
    try:
        for k in range(64):
            ens.set_A_scalar(k, float(As[k]))
        ens.solve_forward(t, method="ssprk3", nsub=8)
        st = 1e-10 if dtype == "f64" else 1e-3
        worst = 0.0
        for k, g in enumerate(gl):
            Hend = ens.get_snapshot(k, 60).astype(np.float64)
            assert np.isfinite(Hend).all() and abs(Hend.sum() / g.H0.sum() - 1.0) < (1e-9 if dtype == "f64" else 1e-4), k
            if dtype == "f32" and k % 4:  # fp32: every 4th glacier against the oracle (bounds the CPU time of the suite)
                continue
            Hs = fast.solve_forward_fixed(g.H0, g, o.TargetA(ph, "const", A=float(As[k])), None, t, method="ssprk3", nsub=8)
            for j in (12, 60):
                err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                worst = max(worst, err)
                assert err <= st, (k, j, err)
        assert np.isfinite(worst)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_config4_32_glacier_inversion_gradient_lawA(ob, dtype):
    """One SIA2D_grad! iteration (gradient.jl:6-31) of the 32-glacier functional inversion: A = NN(T_g), 1-16-16-1."""
    rng, gl = _ensemble_glaciers(2025, 32, dtype)
    temps = rng.uniform(-20.0, 0.0, size=32)
    t = np.linspace(2010.0, 2015.0, 61)
    ph = o.Phys(**PH)
    mlp = o.MLP([1, 16, 16, 1], ["softplus", "softplus", "sigmoid"])
    # theta with the output bias at -3: A_g = minA + (maxA - minA) sigmoid(.) in 2e-18 .. 7e-18, where the reference's reverse loop
    # (explicit Euler, monthly step, gradient.jl:242) is stable on these glaciers; at sigmoid(.) ~ 0.5 (A ~ 4e-17) it is not and
    # neither the oracle's nor the device's gradient is reproducible to rounding (gradient.jl:19-24 warns about that regime)
    th = 0.3 * np.random.default_rng(1).standard_normal(mlp.n_params)
    th[-1] = -3.0
    A_true = lambda T: 4e-18 * np.exp(0.05 * (T + 10.0))  # smooth monotone ground-truth law (CuffeyPaterson table is not in the tree)
    ens = _make_ens(ob, gl, dtype)
    try:
        checked = [k for k in range(32) if dtype == "f64" or k % 4 == 0]  # fp32: every 4th glacier against the oracle
        per = {}
        for k, g in enumerate(gl):
            ens.set_temperature(k, float(temps[k]))
            Atrue = o.TargetA(ph, "const", A=float(A_true(temps[k])))
            if k in checked:
                Href = fast.solve_forward_fixed(g.H0, g, Atrue, None, t, method="ssprk3", nsub=8)
                Href = [_r(h, dtype) for h in Href]
                tg = o.TargetA(ph, "nn", mlp=mlp, T=float(temps[k]))
                Hs = fast.solve_forward_fixed(g.H0, g, tg, th, t, method="ssprk3", nsub=8)
                per[k] = fast.loss_and_grad_discrete(th, g, tg, t, Hs, Href)[:2]
            else:  # (any reference will do for the glaciers that are not compared)
                Href = [g.H0] * len(t)
            for j in range(len(t)):
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
        ens.law_A_nn_apply(mlp.widths, mlp.acts, th)
        ens.solve_forward(t, method="ssprk3", nsub=8)
        fwd_loss = ens.loss(t)
        loss, Ssum = ens.grad_discrete(t)
        dth_dev = ens.law_A_nn_pullback(mlp.n_params)
        lt, gt = (1e-10, 1e-8) if dtype == "f64" else (2e-3, 2e-3)
        assert np.allclose(fwd_loss, loss, rtol=1e-8 if dtype == "f64" else 1e-5, atol=0)  # gradient.jl:259
        for k in checked:
            assert loss[k] == pytest.approx(per[k][0], rel=lt), k
            S1 = np.zeros(32)
            S1[k] = Ssum[k]
            assert rel_l2(ens.law_A_nn_pullback(mlp.n_params, S1), per[k][1]) <= gt, k
        if dtype == "f64":  # sum(losses) and aggregate(d theta) over the whole ensemble (gradient.jl:13-17, Model.jl:208-224)
            assert loss.sum() == pytest.approx(sum(v[0] for v in per.values()), rel=lt)
            assert rel_l2(dth_dev, sum(v[1] for v in per.values())) <= gt
    finally:
        ens.close()


def test_config4_per_cell_law_variant_bounded_sample(ob):
    """Config 4's per-cell variant U = NN(Hbar, gradS), 2-16-16-1 (M1 / M3): the NumPy oracle evaluates the network at every
    node of every RHS, so the check runs on a bounded sample -- the first 3 glaciers of the seed-2025 ensemble, 3 monthly
    tstops -- forward states, loss and the per-glacier d(theta) of the discrete adjoint, fp64."""
    from odinn_b200 import _capi

    _, gl = _ensemble_glaciers(2025, 3, "f64")
    t = o.define_callback_steps((2010.0, 2010.0 + 2.0 / 12.0), 1.0 / 12.0)
    ph = o.Phys(**PH)
    mlp = o.MLP([2, 16, 16, 1], ["softplus", "softplus", "sigmoid"])
    th = 0.4 * np.random.default_rng(3).standard_normal(mlp.n_params)
    bounds = ((0.0, 300.0), (0.0, 0.6))
    max_NN = 60.0
    ens = _make_ens(ob, gl, "f64")
    try:
        ens.law_cell_nn_set("U", mlp.widths, mlp.acts, th, prescale_bounds=bounds, max_NN=max_NN)
        refs = []
        for k, g in enumerate(gl):
            tg = o.TargetD(ph, mlp, prescale_bounds=bounds, max_NN=max_NN)
            Hs = o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=4)
            Href = [0.97 * h for h in Hs]
            for j in range(len(t)):
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            ell, dth, _ = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
            refs.append((Hs, ell, dth))
        ens.solve_forward(t, method="ssprk3", nsub=4)
        loss, _ = ens.grad_discrete(t)
        grads = ens.law_cell_grad()
        for k in range(len(gl)):
            assert rel_l2(ens.get_snapshot(k, len(t) - 1), refs[k][0][-1]) <= 1e-9, k
            assert loss[k] == pytest.approx(refs[k][1], rel=1e-8), k
            assert rel_l2(grads[k], refs[k][2]) <= 1e-6, (k, rel_l2(grads[k], refs[k][2]))
    finally:
        ens.close()
