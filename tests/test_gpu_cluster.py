"""GPU parity of the cluster-resident forward solve (csrc/sia2d_cluster.cuh: one thread-block cluster per glacier, the state in
shared memory, a whole range of tstop intervals per launch) against the NumPy oracle's explicitly stated schemes and against
the marching-kernel time loop (cluster mode 0), for every cluster size, both precisions, ragged ensembles with odd sizes,
non-uniform tstops, generic exponents / eta0 != 1, and with mass-balance callbacks splitting the launch ranges.

Tolerances as in test_gpu_timeloop.py: fp64 state rel-L2 <= 1e-10 against the oracle after the solve (1e-12 between the two CUDA
paths: same arithmetic, different summation order of the divergence only); fp32 <= 1e-3 against the oracle, 1e-5 between paths."""
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


def _r(a, dtype):
    return a.astype(np.float32).astype(np.float64) if dtype == "f32" else a


def _glaciers():
    gl = [o.rough_bed_glacier(40, 35), o.rough_bed_glacier(23, 50), o.dome_glacier(33, 33, H0=150.0), o.rough_bed_glacier(61, 9),
          o.rough_bed_glacier(3, 3), o.rough_bed_glacier(5, 18)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    gl[1].H0[5:9, 20:30] = -0.5   # negative input thickness is clipped inside the RHS only (adjoint.jl:52)
    return gl


def _ens(ob, gl, dtype, phys=None):
    from odinn_b200 import _capi

    ens = ob.Ensemble([g.B.shape[0] for g in gl], [g.B.shape[1] for g in gl], [g.dx for g in gl], [g.dy for g in gl],
                      phys or ob.Phys(**PH), dtype)
    for k, g in enumerate(gl):
        ens.upload(k, _capi.FIELD_B, g.B)
        ens.upload(k, _capi.FIELD_H0, g.H0)
    return ens


AS = [4e-17, 2.21e-18, 1.5e-17, 3e-17, 1e-17, 2e-17]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("method", ["euler", "ssprk3"])
@pytest.mark.parametrize("cs", [1, 2, 4, 8, 16])
def test_cluster_solve_matches_oracle_and_marching(ob, dtype, method, cs):
    from odinn_b200 import _capi

    gl = _glaciers()
    t = np.array([2010.0, 2010.0 + 1 / 12, 2010.0 + 2.5 / 12, 2010.0 + 3 / 12, 2010.0 + 5 / 12])   # non-uniform tstops
    ens = _ens(ob, gl, dtype)
    try:
        for k, a in enumerate(AS):
            ens.set_A_scalar(k, a)
        ens.set_cluster_mode(0)
        l0 = ens.launch_count
        ens.solve_forward(t, method=method, nsub=12)
        n_march = ens.launch_count - l0
        march = [[ens.get_snapshot(k, j) for j in range(len(t))] for k in range(len(gl))]
        ens.set_cluster_mode(cs)
        l0 = ens.launch_count
        ens.solve_forward(t, method=method, nsub=12)
        assert ens.launch_count - l0 == 1 < n_march      # ONE launch for the whole solve
        tol_o, tol_m = (1e-10, 1e-12) if dtype == "f64" else (1e-3, 1e-5)
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(o.Phys(**PH), "const", A=AS[k]), None, t, method=method, nsub=12)
            for j in range(len(t)):
                got = ens.get_snapshot(k, j)
                assert np.isfinite(got).all()
                assert rel_l2(got, Hs[j]) <= tol_o, (k, j, rel_l2(got, Hs[j]))
                assert rel_l2(got, march[k][j]) <= tol_m, (k, j, rel_l2(got, march[k][j]))
            assert rel_l2(ens.download(k, _capi.FIELD_H), Hs[-1]) <= tol_o   # final state left in FIELD_H
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_cluster_solve_generic_exponents(ob, dtype):
    """n != 3, sliding C != 0 and eta0 != 1 take the generic-power instantiation."""
    gl = _glaciers()[:3]
    kw = dict(n=2.6, C=3e-23, p=3.0, q=0.0, eta0=0.7, **PH)
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    ens = _ens(ob, gl, dtype, ob.Phys(**kw))
    try:
        for k in range(len(gl)):
            ens.set_A_scalar(k, AS[k])
        ens.set_cluster_mode(4)
        ens.solve_forward(t, method="ssprk3", nsub=8)
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(o.Phys(**kw), "const", A=AS[k]), None, t, method="ssprk3", nsub=8)
            assert rel_l2(ens.get_snapshot(k, len(t) - 1), Hs[-1]) <= (1e-10 if dtype == "f64" else 1e-3), k
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_cluster_solve_with_mass_balance_ranges(ob, dtype):
    """A mass-balance callback ends a launch range: apply on the global plane, next launch resumes from it."""
    gl = [o.rough_bed_glacier(30, 31), o.rough_bed_glacier(21, 26)]
    for g in gl:
        g.H0 = 0.6 * g.H0
    t = o.define_callback_steps((2010.0, 2010.0 + 5.0 / 12.0), 1.0 / 12.0)
    mb_idx = [2, 4]
    pars = np.array([[(3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0), (4.0, -0.006, 2050.0, 0.7, 0.5, 1.0, 0.5)] for _ in mb_idx])
    As = [3e-17, 1.2e-17]
    ens = _ens(ob, gl, dtype)
    try:
        ens.set_mass_balance(mb_idx, pars)
        for k, a in enumerate(As):
            ens.set_A_scalar(k, a)
        ens.set_cluster_mode(8)
        l0 = ens.launch_count
        ens.solve_forward(t, method="ssprk3", nsub=8)
        assert ens.launch_count - l0 == 3 + len(mb_idx)     # three ranges + two callbacks
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            mb = {j: tuple(pars[m, k]) for m, j in enumerate(mb_idx)}
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(o.Phys(**PH), "const", A=As[k]), None, t, mb=mb, method="ssprk3", nsub=8)
            for j in range(len(t)):
                assert rel_l2(ens.get_snapshot(k, j), Hs[j]) <= (1e-10 if dtype == "f64" else 1e-3), (k, j)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("cs", [1, 4, 16])
def test_cluster_rdpk3sp35_matches_oracle_and_engine(ob, dtype, cs):
    """The reference's default integrator (RDPK3Sp35 + PID, AdjointTypes.jl:60) with the whole adaptive loop inside the cluster ==
    the oracle's integrate_rdpk3sp35 and the host-driven device engine; fp64 takes the same accept / reject sequence."""
    gl = _glaciers()
    t = np.array([2010.0, 2010.0 + 1 / 12, 2010.0 + 2.5 / 12, 2010.0 + 3 / 12, 2010.0 + 5 / 12])
    rtol = 1e-6 if dtype == "f64" else 1e-4
    ens = _ens(ob, gl, dtype)
    try:
        for k, a in enumerate(AS):
            ens.set_A_scalar(k, a)
        ens.set_cluster_mode(0)
        steps0, rej0 = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35")
        eng = [[ens.get_snapshot(k, j) for j in range(len(t))] for k in range(len(gl))]
        ens.set_cluster_mode(cs)
        l0 = ens.launch_count
        steps, rej = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35")
        assert ens.launch_count - l0 == 1
        for k, g in enumerate(gl):
            g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
            st = {}
            Hs = o.solve_forward(g2.H0, g2, o.TargetA(o.Phys(**PH), "const", A=AS[k]), None, t, method="rdpk3sp35", reltol=rtol, abstol=rtol,
                                 stats=st)
            if dtype == "f64":
                assert steps[k] == st["steps"] == steps0[k] and rej[k] == st["rejected"] == rej0[k], (k, steps[k], rej[k], st)
            for j in range(len(t)):
                got = ens.get_snapshot(k, j)
                assert rel_l2(got, Hs[j]) <= (1e-9 if dtype == "f64" else 2e-3), (k, j, rel_l2(got, Hs[j]))
                assert rel_l2(got, eng[k][j]) <= (1e-9 if dtype == "f64" else 2e-3), (k, j)
        # a user-given first step (dt0 > 0) skips the initial-step algorithm
        s1, _ = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35", dt0=1e-3)
        ens.set_cluster_mode(0)
        s2, _ = ens.solve_forward_adaptive(t, reltol=rtol, abstol=rtol, method="rdpk3sp35", dt0=1e-3)
        if dtype == "f64":
            assert np.array_equal(s1, s2)
    finally:
        ens.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("cs", [1, 2, 8, 16])
def test_cluster_reverse_loop_matches_marching_and_oracle(ob, dtype, cs):
    """R1 / L1 cluster-resident (the whole discrete-adjoint reverse loop in one launch): loss, S = dL/dA-scalar and lambda(t0) equal the
    per-step kernel sequence (cluster mode 0) and the oracle's loss_and_grad_discrete; sparse H_ref data; mass-balance tstops split the
    launch ranges."""
    from odinn_b200 import _capi

    gl = _glaciers()
    t = np.array([2010.0, 2010.0 + 1 / 12, 2010.0 + 2.5 / 12, 2010.0 + 3 / 12, 2010.0 + 5 / 12, 2010.0 + 6 / 12])
    ph = o.Phys(**PH)
    for with_mb in (False, True):
        ens = _ens(ob, gl, dtype)
        try:
            mb_idx = [2, 4]
            pars = np.array([[(3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)] * len(gl) for _ in mb_idx])
            if with_mb:
                ens.set_mass_balance(mb_idx, pars)
            for k, a in enumerate(AS):
                ens.set_A_scalar(k, a)
            ens.set_cluster_mode(0)
            ens.solve_forward(t, method="ssprk3", nsub=12)
            refs = []
            for k, g in enumerate(gl):
                g2 = o.Glacier(B=_r(g.B, dtype), dx=g.dx, dy=g.dy, H0=_r(g.H0, dtype))
                mb = {j: tuple(pars[m, k]) for m, j in enumerate(mb_idx)} if with_mb else None
                Href = [_r(h, dtype) for h in o.solve_forward(g2.H0, g2, o.TargetA(ph, "const", A=2e-17), None, t, method="ssprk3", nsub=12, mb=mb)]
                for j in range(len(t)):
                    if j != 3:   # sparse thickness data: no H_ref at tstop 3 (gradient.jl:79-80, 144-149)
                        ens.set_reference(k, j, len(t), Href[j], o.is_in_glacier(Href[j], 3))
                refs.append(Href)
            loss0, S0 = ens.grad_discrete(t)
            lam0 = [ens.download(k, _capi.FIELD_LAMBDA) for k in range(len(gl))]
            ens.set_cluster_mode(cs)
            l0 = ens.launch_count
            loss1, S1 = ens.grad_discrete(t)
            assert ens.launch_count - l0 == (1 + 2 * len(mb_idx) if with_mb else 1), ens.launch_count - l0
            rt = 1e-10 if dtype == "f64" else 2e-4
            for k in range(len(gl)):
                assert loss1[k] == pytest.approx(loss0[k], rel=rt), (k, loss1[k], loss0[k])
                assert S1[k] == pytest.approx(S0[k], rel=1e-8 if dtype == "f64" else 5e-3, abs=1e-30), (k, S1[k], S0[k])
                lam1 = ens.download(k, _capi.FIELD_LAMBDA)
                assert rel_l2(lam1, lam0[k]) <= (1e-10 if dtype == "f64" else 1e-4), (k, rel_l2(lam1, lam0[k]))
            loss2, S2 = ens.grad_discrete(t)
            assert np.array_equal(loss1, loss2) and np.array_equal(S1, S2)   # bit-stable run to run
            if not with_mb and dtype == "f64":   # and against the oracle on the device snapshots
                for k, g in enumerate(gl[:3]):
                    g2 = o.Glacier(B=g.B, dx=g.dx, dy=g.dy, H0=g.H0)
                    Hs_d = [ens.get_snapshot(k, j).astype(np.float64) for j in range(len(t))]
                    tgs = o.TargetA(ph, "scalar")
                    theta = np.array([np.arctanh(2 * (AS[k] - ph.minA) / (ph.maxA - ph.minA) - 1)])
                    Href = list(refs[k])
                    wH = np.zeros(len(t)); last = 0
                    for j in range(1, len(t)):
                        if j != 3:
                            wH[j] = t[j] - t[last]; last = j
                    Href[3] = Hs_d[3]
                    ell, dth = o.loss_and_grad_discrete_HV(theta, g2, tgs, t, Hs_d, Href, [None] * len(t), wH, np.zeros(len(t)))
                    assert loss1[k] == pytest.approx(ell, rel=1e-10), k
                    assert S1[k] * tgs.vjp_theta[0] == pytest.approx(dth[0], rel=1e-8), (k, S1[k] * tgs.vjp_theta[0], dth[0])
        finally:
            ens.close()


def test_cluster_mode_automatic_choice_and_fallback(ob):
    """Automatic mode: a small ensemble runs cluster-resident (one launch), an ensemble holding a glacier too large for the shared
    memory of a cluster falls back to the marching kernels; results agree."""
    t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)
    g_small = o.rough_bed_glacier(128, 128)
    g_small.H0 = 0.5 * g_small.H0
    ens = _ens(ob, [g_small], "f32")
    try:
        ens.set_A_scalar(0, 2e-17)
        l0 = ens.launch_count
        ens.solve_forward(t, method="ssprk3", nsub=8)
        assert ens.launch_count - l0 == 1
        a = ens.get_snapshot(0, len(t) - 1)
        ens.set_cluster_mode(0)
        ens.solve_forward(t, method="ssprk3", nsub=8)
        assert rel_l2(ens.get_snapshot(0, len(t) - 1), a) <= 1e-5
        with pytest.raises(Exception):
            ens.set_cluster_mode(3)
    finally:
        ens.close()
    g_big = o.rough_bed_glacier(640, 600)
    g_big.H0 = 0.3 * g_big.H0
    ens = _ens(ob, [g_big, g_small], "f32")
    try:
        ens.set_A_scalar(0, 1e-17)
        ens.set_A_scalar(1, 2e-17)
        l0 = ens.launch_count
        ens.solve_forward(t[:2], method="ssprk3", nsub=16)
        assert ens.launch_count - l0 > 1
        assert np.isfinite(ens.get_snapshot(0, 1)).all()
    finally:
        ens.close()
