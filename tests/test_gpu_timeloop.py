"""GPU parity of the device-resident time loop, loss, discrete-adjoint reverse loop and the A law.

Everything is compared with the NumPy oracle running the SAME explicitly stated schemes
(oracle.solve_forward "euler"/"ssprk3", oracle.loss_and_grad_discrete).  Tolerances: fp64 state after
the forward solve rel-L2 <= 1e-10, loss rtol 1e-10, gradient rtol 1e-8; fp32 state <= 1e-3 after the
solve, loss/gradient rtol 2e-3.  Forward/reverse loss equality: rtol 1e-8
(src/inverse/SIA2D/gradient.jl:259)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

PH = dict(minA=8e-21, maxA=8e-17)  # test/inversion_test.jl:59-60


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    return odinn_b200


def _glaciers():
    """Thin (<= 150 m) glaciers so that the fixed 1/96 yr sub-step is stable for every A in [minA, maxA]."""
    gl = [o.rough_bed_glacier(40, 35), o.rough_bed_glacier(23, 50), o.dome_glacier(33, 33, H0=150.0)]
    for g in gl[:2]:
        g.H0 = 0.6 * g.H0
    return gl


def _ens(ob, gl, dtype):
    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl],
                      ob.Phys(**PH), dtype)
    from odinn_b200 import _capi

    for k, g in enumerate(gl):
        ens.upload(k, _capi.FIELD_B, g.B)
        ens.upload(k, _capi.FIELD_H0, g.H0)
    return ens


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_adaptive_bs3_matches_oracle(ob, dtype):
    """N1: adaptive BS3 with tstops, one step-size controller per glacier on the device == the oracle's bs3 run glacier
    by glacier (same error norm, same controller); fp64 takes the same accept/reject sequence (same RHS count)."""
    gl = _glaciers()
    As = [4e-17, 2.21e-18, 1.5e-17]
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    ens = _ens(ob, gl, dtype)
    rtol = 1e-5 if dtype == "f64" else 1e-4   # solver tolerances (the fp32 error estimate bottoms out near 1e-4)
    try:
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        steps, rej = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol)
        assert np.all(steps > 0) and np.all(rej >= 0)
        for k, g in enumerate(gl):
            if dtype == "f32":
                g = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0.astype(np.float32).astype(np.float64))
            tg = o.TargetA(o.Phys(**PH), "const", A=As[k])
            stats = {}
            Hs = o.solve_forward(g.H0, g, tg, None, t, method="bs3", reltol=rtol, abstol=rtol, stats=stats)
            if dtype == "f64":
                assert stats["nrhs"] == 1 + 3 * steps[k], (k, stats, steps[k])
            for j in (1, len(t) // 2, len(t) - 1):
                err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                assert err <= (1e-10 if dtype == "f64" else 1e-3), (k, j, err)
        # mass is conserved by every accepted step (interior ice, zero border): Σ H changes only through clipping at 0
        assert np.isfinite(ens.get_snapshot(0, len(t) - 1)).all()
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("method", ["euler", "ssprk3"])
def test_forward_solve_matches_oracle(ob, dtype, method):
    gl = _glaciers()
    As = [4e-17, 2.21e-18, 1.5e-17]
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    ens = _ens(ob, gl, dtype)
    try:
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        ens.solve_forward(t, method=method, nsub=8)
        for k, g in enumerate(gl):
            if dtype == "f32":  # oracle sees the same rounded inputs
                g = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0.astype(np.float32).astype(np.float64))
            tg = o.TargetA(o.Phys(**PH), "const", A=As[k])
            Hs = o.solve_forward(g.H0, g, tg, None, t, method=method, nsub=8)
            for j in (0, 1, len(t) - 1):
                err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                assert err <= (1e-10 if dtype == "f64" else 1e-3), (k, j, err)
            # mass is conserved while the ice stays inside the grid
            assert abs(ens.get_snapshot(k, len(t) - 1).sum() / g.H0.sum() - 1.0) < (1e-10 if dtype == "f64" else 1e-4)
    finally:
        ens.close()


def test_law_A_nn_matches_oracle(ob):
    gl = _glaciers()
    temps = [-10.0, -3.5, -18.0]
    ens = _ens(ob, gl, "f64")
    try:
        for mlp in (o.MLP.default(1), o.MLP.default(1, light=True), o.MLP([1, 16, 16, 1], ["softplus", "softplus", "sigmoid"]),
                    o.MLP([1, 5, 4, 1], ["tanh", "relu", "identity"])):
            th = mlp.init(11, scale=0.9) + 0.05
            for k, T in enumerate(temps):
                ens.set_temperature(k, T)
            A = ens.law_A_nn_apply(mlp.widths, mlp.acts, th)
            ph = o.Phys(**PH)
            for k, T in enumerate(temps):
                tg = o.TargetA(ph, "nn", mlp=mlp, T=T)
                tg.apply_laws(None, None, th)
                tg.precompute_vjp(th)
                assert A[k] == pytest.approx(tg.A, rel=1e-13)
                S = np.zeros(len(gl))
                S[k] = 1.0
                J = ens.law_A_nn_pullback(mlp.n_params, S)
                assert rel_l2(J, tg.vjp_theta) < 1e-12
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_loss_and_discrete_adjoint_gradient_match_oracle(ob, dtype):
    """Twin experiment in miniature (test/inversion_test.jl): H_ref from A_true, NN law for the inversion."""
    gl = _glaciers()
    temps = [-10.0, -3.5, -18.0]
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    ph = o.Phys(**PH)
    mlp = o.MLP.default(1)
    th = mlp.init(3, scale=0.8)
    ens = _ens(ob, gl, dtype)
    npdt = np.float32 if dtype == "f32" else np.float64
    try:
        refs, tot_loss, tot_grad = [], 0.0, np.zeros(mlp.n_params)
        for k, g in enumerate(gl):
            g = o.Glacier(B=g.B.astype(npdt).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0.astype(npdt).astype(np.float64))
            Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=4e-17), None, t, method="ssprk3", nsub=8)
            Href = [h.astype(npdt).astype(np.float64) for h in Href]
            for j in range(len(t)):
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            ens.set_temperature(k, temps[k])
            tg = o.TargetA(ph, "nn", mlp=mlp, T=temps[k])
            Hs = o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=8)
            ell, dth, _ = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
            refs.append((ell, dth))
            tot_loss += ell
            tot_grad += dth
        ens.law_A_nn_apply(mlp.widths, mlp.acts, th)
        ens.solve_forward(t, method="ssprk3", nsub=8)
        fwd_loss = ens.loss(t)
        loss, Ssum = ens.grad_discrete(t)
        dth = ens.law_A_nn_pullback(mlp.n_params)
        rt_l, rt_g = (1e-10, 1e-8) if dtype == "f64" else (2e-3, 2e-3)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert fwd_loss[k] == pytest.approx(loss[k], rel=1e-8 if dtype == "f64" else 1e-5)  # gradient.jl:259
        assert loss.sum() == pytest.approx(tot_loss, rel=rt_l)
        assert rel_l2(dth, tot_grad) <= rt_g, rel_l2(dth, tot_grad)
        # per-glacier contributions (aggregate∇θ sums them, Model.jl:208-224)
        for k in range(len(gl)):
            S1 = np.zeros(len(gl))
            S1[k] = Ssum[k]
            assert rel_l2(ens.law_A_nn_pullback(mlp.n_params, S1), refs[k][1]) <= rt_g, k
        # bit-stable run to run
        loss2, Ssum2 = ens.grad_discrete(t)
        assert np.array_equal(loss, loss2) and np.array_equal(Ssum, Ssum2)
    finally:
        ens.close()


def test_snapshots_from_host_and_state_errors(ob):
    """Snapshots saved by an external integrator (OrdinaryDiffEq in the reference) can be fed to the reverse loop;
    calling the gradient without data is a state error, not a crash."""
    g = o.rough_bed_glacier(30, 31)
    g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    ph = o.Phys(**PH)
    tg = o.TargetA(ph, "const", A=3e-17)
    ens = _ens(ob, [g], "f64")
    try:
        with pytest.raises(ob.OdinnError):
            ens.grad_discrete(t)
        Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="bs3", reltol=1e-8, abstol=1e-8)
        Hs = o.solve_forward(g.H0, g, tg, None, t, method="bs3", reltol=1e-8, abstol=1e-8)
        for j in range(len(t)):
            ens.set_snapshot(0, j, len(t), Hs[j])
            ens.set_reference(0, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
        ens.set_A_scalar(0, 3e-17)
        loss, Ssum = ens.grad_discrete(t)
        tgs = o.TargetA(ph, "scalar")
        theta = np.array([np.arctanh(2 * (3e-17 - ph.minA) / (ph.maxA - ph.minA) - 1)])
        ell, dth, _ = o.loss_and_grad_discrete(theta, g, tgs, t, Hs, Href)
        tgs.precompute_vjp(theta)
        assert loss[0] == pytest.approx(ell, rel=1e-10)
        assert Ssum[0] * tgs.vjp_theta[0] == pytest.approx(dth[0], rel=1e-8)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("vjp", ["discrete", "continuous"])
@pytest.mark.parametrize("method", ["ssprk3", "euler"])
def test_continuous_adjoint_gradient_matches_oracle(ob, dtype, vjp, method):
    """N4: ContinuousAdjoint (gradient.jl:276-538) on the device == the oracle running the same reverse scheme: loss jumps at
    the tstops, Gauss-Legendre quadrature of the theta-VJP, linear interpolation of the snapshots."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    ph = o.Phys(**PH)
    As = [3e-17, 1.2e-17]
    ens = _ens(ob, gl, dtype)
    npdt = np.float32 if dtype == "f32" else np.float64
    try:
        refs = []
        for k, g in enumerate(gl):
            g32 = o.Glacier(B=g.B.astype(npdt).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0)
            Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=5e-17), None, t, method="ssprk3", nsub=8)
            Hs = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=As[k]), None, t, method="ssprk3", nsub=8)
            Href = [h.astype(npdt).astype(np.float64) for h in Href]
            Hs = [h.astype(npdt).astype(np.float64) for h in Hs]
            for j in range(len(t)):
                ens.set_snapshot(k, j, len(t), Hs[j])
                ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
            ens.set_A_scalar(k, As[k])
            tgs = o.TargetA(ph, "scalar")
            theta = np.array([np.arctanh(2 * (As[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
            ell, dth = o.loss_and_grad_continuous(theta, g32, tgs, t, Hs, Href, n_quadrature=7, nsub=2, method=method, vjp=vjp)
            refs.append((ell, dth[0], tgs.vjp_theta[0]))
        loss, Ssum = ens.grad_continuous(t, n_quadrature=7, vjp=vjp, method=method, nsub=2)
        rt_l, rt_g = (1e-10, 1e-8) if dtype == "f64" else (2e-4, 2e-3)
        for k in range(len(gl)):
            assert loss[k] == pytest.approx(refs[k][0], rel=rt_l), k
            assert Ssum[k] * refs[k][2] == pytest.approx(refs[k][1], rel=rt_g), (k, Ssum[k] * refs[k][2], refs[k][1])
        loss2, Ssum2 = ens.grad_continuous(t, n_quadrature=7, vjp=vjp, method=method, nsub=2)
        assert np.array_equal(loss, loss2) and np.array_equal(Ssum, Ssum2)  # bit-stable run to run
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_forward_graph_replay_nonuniform_tstops_and_reuse(ob, dtype):
    """The fixed-step loop replays ONE captured CUDA graph per interval with the step sizes read from a device table: non-uniform
    tstops, a second solve on the same handle with other A values (graph reuse), and ODINN_NO_GRAPH-equivalent results."""
    gl = _glaciers()
    t = np.array([2010.0, 2010.05, 2010.2, 2010.25, 2010.45, 2010.5])
    ens = _ens(ob, gl, dtype)
    try:
        for As in ([4e-17, 2.21e-18, 1.5e-17], [1e-17, 3e-17, 5e-18]):
            for k, a in enumerate(As):
                ens.set_A_scalar(k, a)
            ens.solve_forward(t, method="ssprk3", nsub=12)
            for k, g in enumerate(gl):
                if dtype == "f32":
                    g = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0.astype(np.float32).astype(np.float64))
                Hs = o.solve_forward(g.H0, g, o.TargetA(o.Phys(**PH), "const", A=As[k]), None, t, method="ssprk3", nsub=12)
                for j in (1, 3, len(t) - 1):
                    err = rel_l2(ens.get_snapshot(k, j), Hs[j])
                    assert err <= (1e-10 if dtype == "f64" else 1e-3), (k, j, err)
        # odd Euler sub-step count: the plane rotation does not close, the loop falls back to direct launches
        ens.solve_forward(t, method="euler", nsub=13)
        g = gl[0]
        if dtype == "f32":
            g = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy, H0=g.H0.astype(np.float32).astype(np.float64))
        Hs = o.solve_forward(g.H0, g, o.TargetA(o.Phys(**PH), "const", A=1e-17), None, t, method="euler", nsub=13)
        assert rel_l2(ens.get_snapshot(0, len(t) - 1), Hs[-1]) <= (1e-10 if dtype == "f64" else 1e-3)
    finally:
        ens.close()
