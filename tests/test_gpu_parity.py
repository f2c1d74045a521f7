"""GPU parity: the CUDA path (through the C ABI) against the NumPy oracle on the same inputs.

Tolerances: fp64 kernel vs fp64 oracle rel-L2 <= 1e-12 per call
(the reference's operator tests use rtol 1e-11, test/SIA2D_adjoint_utils.jl:22); fp32 kernel vs
fp64 oracle evaluated on the SAME fp32-rounded inputs rel-L2 <= 1e-5 (the reference's fp32 rtol).
"""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_numpy as o

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
A0 = 2.21e-18  # test/test_grad_loss.jl:157


def _tilted_dome(nx, ny):
    g = o.dome_glacier(nx, ny)
    X = np.arange(nx)[:, None] * g.dx
    Y = np.arange(ny)[None, :] * g.dy
    g.B = 0.03 * X + 0.011 * Y
    return g


def _noisy(nx, ny):
    """Rough bed + noisy thickness with holes: many clamp-active edges and H = 0 cells inside the ice."""
    g = o.rough_bed_glacier(nx, ny)
    rng = np.random.default_rng(7)
    H = g.H0 * (1.0 + 0.3 * rng.standard_normal(g.H0.shape))
    H[rng.random(H.shape) < 0.05] = 0.0
    H[rng.random(H.shape) < 0.02] = -3.0  # negative input must be clipped, not propagated (adjoint.jl:52)
    g.H0 = H
    return g


MAKERS = {"rough": o.rough_bed_glacier, "dome": o.dome_glacier, "tilted": _tilted_dome, "noisy": _noisy}


def _round_inputs(g, H, lam, dtype):
    """Inputs as the kernel sees them (fp32 rounding applied before the fp64 oracle runs)."""
    npdt = np.float32 if dtype == "f32" else np.float64
    g2 = o.Glacier(B=g.B.astype(npdt).astype(np.float64), dx=g.dx, dy=g.dy)
    return g2, H.astype(npdt).astype(np.float64), lam.astype(npdt).astype(np.float64)


def _S_close(S, lam, H, g, tg, dtype):
    """|S - ref| <= 10 tol |ref|, or -- when the sum cancels -- a few roundings of the accumulated magnitude Σ|v|."""
    ref, mag = o.node_reduction_S_terms(lam, H, g, tg)
    return abs(S - ref) <= 10 * TOL[dtype] * abs(ref) or abs(S - ref) <= (3e-7 if dtype == "f32" else 1e-14) * mag


def _sim(ob, g, phys_kw, A, dtype):
    return ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy)], ob.Phys(**phys_kw), A=A, dtype=dtype)


@pytest.fixture(scope="module")
def ob():
    import odinn_b200

    return odinn_b200


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("maker", list(MAKERS))
@pytest.mark.parametrize("shape", [(14, 17), (3, 3), (3, 41), (40, 3), (33, 47), (128, 128), (97, 64)])
def test_forward_and_vjps_match_oracle(ob, dtype, maker, shape):
    nx, ny = shape
    g = MAKERS[maker](nx, ny)
    lam = np.random.default_rng(1234).standard_normal((nx, ny))
    g, H, lam = _round_inputs(g, g.H0, lam, dtype)
    tg = o.TargetA(o.Phys(), "const", A=A0)
    sim = _sim(ob, g, {}, A0, dtype)
    try:
        dH = np.full((nx, ny), np.nan, order="F")
        Hin = H.copy()
        ob.SIA2D_(dH, Hin, sim, 0.0)
        assert np.array_equal(Hin, H)  # the RHS must not mutate H (docs/src/sensitivity.md:83)
        ref = o.SIA2D(H, g, tg)
        assert rel_l2(dH, ref) <= TOL[dtype], ("rhs", rel_l2(dH, ref))
        # border rows/columns are exactly zero
        assert np.all(dH[0, :] == 0) and np.all(dH[-1, :] == 0) and np.all(dH[:, 0] == 0) and np.all(dH[:, -1] == 0)

        vH, none = ob.VJP_λ_dSIAdH(ob.B200VJP(), lam, H, None, sim, 0.0)
        assert none is None
        refv = o.VJP_dSIA_dH_discrete(lam, H, g, tg)
        assert rel_l2(vH, refv) <= TOL[dtype], ("vjp_H", rel_l2(vH, refv))
        assert np.all(vH[np.maximum(H, 0) <= 0] == 0)  # adjoint.jl:148

        S = ob.VJP_λ_dSIAdθ(ob.B200VJP(), lam, H, None, None, sim, 0.0)
        assert _S_close(S, lam, H, g, tg, dtype), ("vjp_theta", S, o.node_reduction_S(lam, H, g, tg))
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("phys_kw", [dict(C=7e-8), dict(n=3.5), dict(n=4.0, C=3e-9, p=2.5, q=0.5), dict(eta0=0.3)])
def test_generic_exponents_and_sliding(ob, dtype, phys_kw):
    """The pow() path: sliding term C != 0 (test/runtests.jl:94,99) and non-cubic Glen exponents."""
    nx, ny = 37, 29
    g = o.rough_bed_glacier(nx, ny)
    lam = np.random.default_rng(1234).standard_normal((nx, ny))
    g, H, lam = _round_inputs(g, g.H0, lam, dtype)
    tg = o.TargetA(o.Phys(**phys_kw), "const", A=A0)
    sim = _sim(ob, g, phys_kw, A0, dtype)
    tol = TOL[dtype] * (4 if dtype == "f32" else 1)
    try:
        dH = np.zeros((nx, ny), order="F")
        ob.SIA2D_(dH, H, sim, 0.0)
        assert rel_l2(dH, o.SIA2D(H, g, tg)) <= tol
        vH, _ = ob.VJP_λ_dSIAdH(ob.B200VJP(), lam, H, None, sim, 0.0)
        assert rel_l2(vH, o.VJP_dSIA_dH_discrete(lam, H, g, tg)) <= tol
        S = ob.VJP_λ_dSIAdθ(ob.B200VJP(), lam, H, None, None, sim, 0.0)
        refS = o.node_reduction_S(lam, H, g, tg)
        assert abs(S - refS) <= 10 * tol * abs(refS)
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_gridded_A(ob, dtype):
    """Spatially varying A on the dual grid (MatrixCache; LawA(params; scalar=false), Laws.jl:430-454)."""
    nx, ny = 45, 38
    g = o.rough_bed_glacier(nx, ny)
    rng = np.random.default_rng(5)
    Afield = A0 * np.exp(rng.uniform(-1, 1, size=(nx - 1, ny - 1)))
    lam = rng.standard_normal((nx, ny))
    g, H, lam = _round_inputs(g, g.H0, lam, dtype)
    if dtype == "f32":
        Afield = Afield.astype(np.float32).astype(np.float64)
    tg = o.TargetA(o.Phys(), "const", A=Afield)
    sim = _sim(ob, g, {}, [Afield], dtype)
    try:
        dH = np.zeros((nx, ny), order="F")
        ob.SIA2D_(dH, H, sim, 0.0)
        assert rel_l2(dH, o.SIA2D(H, g, tg)) <= TOL[dtype]
        vH, _ = ob.VJP_λ_dSIAdH(ob.B200VJP(), lam, H, None, sim, 0.0)
        assert rel_l2(vH, o.VJP_dSIA_dH_discrete(lam, H, g, tg)) <= TOL[dtype]
        # per-node integrand ∂A_spatial ∘ D† (sparse_cartesian_tensor path, target_utils.jl:163-173)
        f = o._recompute_forward(H, g, tg, None)
        _, _, Dadj = o._D_adjoint(lam, f, g.dx, g.dy)
        ref_field = o.Gamma(tg.ph) * f["Hb"] ** 5 * f["gS"] ** 2 * Dadj
        S = ob.VJP_λ_dSIAdθ(ob.B200VJP(), lam, H, None, None, sim, 0.0)
        from odinn_b200 import _capi

        got = sim.ensemble.download(0, _capi.FIELD_VJP_A)
        assert rel_l2(got, ref_field) <= TOL[dtype]
        assert abs(S - ref_field.sum()) <= 10 * TOL[dtype] * abs(ref_field.sum())
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_mixed_size_ensemble_batch(ob, dtype):
    """One launch over a ragged ensemble (config 3 in miniature) == per-glacier oracle results."""
    rng = np.random.default_rng(2024)
    shapes = [(int(rng.integers(20, 90)), int(rng.integers(20, 90))) for _ in range(7)] + [(3, 3), (32, 16), (33, 17)]
    gl, Hs, lams, As = [], [], [], []
    for k, (nx, ny) in enumerate(shapes):
        g = (o.rough_bed_glacier if k % 2 else _tilted_dome)(nx, ny)
        lam = rng.standard_normal((nx, ny))
        g, H, lam = _round_inputs(g, g.H0, lam, dtype)
        gl.append(g), Hs.append(H), lams.append(lam), As.append(A0 * (1 + k))
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy) for g in gl], ob.Phys(), A=As, dtype=dtype)
    from odinn_b200 import _capi

    try:
        ens = sim.ensemble
        for k in range(len(gl)):
            ens.upload(k, _capi.FIELD_H, Hs[k])
            ens.upload(k, _capi.FIELD_LAMBDA, lams[k])
        ens.rhs_resident()
        S = ens.vjp_resident(True, True)
        for k, g in enumerate(gl):
            tg = o.TargetA(o.Phys(), "const", A=As[k])
            assert rel_l2(ens.download(k, _capi.FIELD_DH), o.SIA2D(Hs[k], g, tg)) <= TOL[dtype], k
            assert rel_l2(ens.download(k, _capi.FIELD_VJP_H), o.VJP_dSIA_dH_discrete(lams[k], Hs[k], g, tg)) <= TOL[dtype], k
            assert _S_close(S[k], lams[k], Hs[k], g, tg, dtype), k
        # determinism: the two-stage reduction is bit-stable run to run
        S2 = ens.vjp_resident(True, True)
        assert np.array_equal(S, S2)
    finally:
        sim.close()


def test_vjp_is_transpose_of_jacobian_at_scale(ob):
    """Size-independent property at BASELINE config-2 size (500x500): <J v, λ> == <v, Jᵀ λ> with the
    directional derivative taken by central differences of the CUDA forward (fp64)."""
    nx = ny = 500
    g = _tilted_dome(nx, ny)
    rng = np.random.default_rng(1234)
    lam = rng.standard_normal((nx, ny))
    v = rng.standard_normal((nx, ny)) * (g.H0 > 5.0)  # stay away from the H = 0 kink
    sim = _sim(ob, g, {}, A0, "f64")
    try:
        eps = 1e-4
        dp = np.zeros((nx, ny), order="F")
        dm = np.zeros((nx, ny), order="F")
        ob.SIA2D_(dp, g.H0 + eps * v, sim, 0.0)
        ob.SIA2D_(dm, g.H0 - eps * v, sim, 0.0)
        lhs = np.sum((dp - dm) / (2 * eps) * lam)
        vH, _ = ob.VJP_λ_dSIAdH(ob.B200VJP(), lam, g.H0, None, sim, 0.0)
        rhs = np.sum(vH * v)
        assert abs(lhs - rhs) <= 1e-6 * abs(rhs), (lhs, rhs)
        # mass conservation of the forward: Σ dH == 0 up to rounding when no ice touches the border
        d0 = np.zeros((nx, ny), order="F")
        ob.SIA2D_(d0, g.H0, sim, 0.0)
        assert abs(d0.sum()) <= 1e-9 * np.abs(d0).sum()
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_fused_step_at_bench_size_properties(ob, dtype):
    """BASELINE-size grids (500 x 500 and odd neighbours) through size-independent properties of the fused F1 + A1 + A2 pass:
    it reproduces the two separate kernels, the VJP is linear in λ, ∂H vanishes on ice-free cells, and dH conserves mass."""
    from odinn_b200 import _capi

    shapes = [(500, 500), (499, 501), (501, 487)]
    gl = [o.rough_bed_glacier(nx, ny) for nx, ny in shapes]
    rng = np.random.default_rng(11)
    lams = [rng.standard_normal(g.B.shape) for g in gl]
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy) for g in gl], ob.Phys(), A=[A0, 2 * A0, 3 * A0], dtype=dtype)
    try:
        ens = sim.ensemble
        for k, g in enumerate(gl):
            ens.upload(k, _capi.FIELD_H, g.H0)
            ens.upload(k, _capi.FIELD_LAMBDA, lams[k])
        ens.rhs_resident()
        S_sep = ens.vjp_resident(True, True)
        dH_sep = [ens.download(k, _capi.FIELD_DH) for k in range(3)]
        vH_sep = [ens.download(k, _capi.FIELD_VJP_H) for k in range(3)]
        S_f = ens.vjp_resident(True, True, want_dH=True)
        eps = 1e-6 if dtype == "f32" else 1e-14
        for k, g in enumerate(gl):
            dH, vH = ens.download(k, _capi.FIELD_DH), ens.download(k, _capi.FIELD_VJP_H)
            assert rel_l2(dH, dH_sep[k]) <= eps and rel_l2(vH, vH_sep[k]) <= eps, k
            assert abs(S_f[k] - S_sep[k]) <= 5e-6 * abs(S_sep[k]) + 1e-30, k   # (a cancelling fp32 sum accumulated over different row chunks)
            assert np.all(vH[np.maximum(g.H0, 0) <= 0] == 0)                      # adjoint.jl:148
            assert abs(dH.astype(np.float64).sum()) <= (1e-4 if dtype == "f32" else 1e-9) * np.abs(dH).sum()
            assert np.all(dH[0, :] == 0) and np.all(dH[-1, :] == 0) and np.all(dH[:, 0] == 0) and np.all(dH[:, -1] == 0)
        # linearity in λ: VJP(2λ) == 2 VJP(λ) exactly (scaling by 2 is exact in floating point)
        for k in range(3):
            ens.upload(k, _capi.FIELD_LAMBDA, 2.0 * lams[k])
        ens.vjp_resident(True, True, want_dH=True)
        for k in range(3):
            assert np.array_equal(ens.download(k, _capi.FIELD_VJP_H), 2.0 * vH_sep[k].astype(ens.np_dtype)) or \
                rel_l2(ens.download(k, _capi.FIELD_VJP_H), 2.0 * vH_sep[k]) <= eps, k
    finally:
        sim.close()


def test_error_convention(ob):
    """Nonzero return code -> exception carrying odinn_last_error (never exit)."""
    with pytest.raises(ob.OdinnError):
        ob.Ensemble([2], [5], [50.0], [50.0])  # grid smaller than 3x3
    sim = _sim(ob, o.rough_bed_glacier(8, 9), {}, A0, "f64")
    try:
        with pytest.raises(ob.OdinnError):
            sim.ensemble.sia2d_rhs(3, np.zeros((8, 9)))  # glacier index out of range
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("gridded", [False, True])
def test_host_batch_pipeline(ob, dtype, gridded):
    """odinn_fwd_adj_batch_host (chunked H2D -> kernels -> D2H pipeline; packed layout for scalar A, padded layout
    for gridded A) == per-glacier oracle, with several chunks and the resident planes left untouched."""
    from odinn_b200 import _capi

    rng = np.random.default_rng(99)
    shapes = [(int(rng.integers(8, 70)), int(rng.integers(8, 70))) for _ in range(9)] + [(3, 3), (64, 5)]
    gl, Hs, lams, As = [], [], [], []
    for k, (nx, ny) in enumerate(shapes):
        g = (o.rough_bed_glacier if k % 2 else _tilted_dome)(nx, ny)
        lam = rng.standard_normal((nx, ny))
        g, H, lam = _round_inputs(g, g.H0, lam, dtype)
        A = A0 * (1 + k)
        if gridded:
            A = A0 * np.exp(rng.uniform(-1, 1, size=(nx - 1, ny - 1)))
            if dtype == "f32":
                A = A.astype(np.float32).astype(np.float64)
        gl.append(g), Hs.append(H), lams.append(lam), As.append(A)
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy) for g in gl], ob.Phys(), A=As, dtype=dtype)
    try:
        ens = sim.ensemble
        ens.set_batch_chunk(3000)  # several chunks
        marker = np.full(shapes[0], 7.0)
        ens.upload(0, _capi.FIELD_H, marker)
        for rep in range(2):  # second call reuses the staging planes / events
            dH, vH, S = ens.fwd_adj_batch(Hs, lams)
            for k, g in enumerate(gl):
                tg = o.TargetA(o.Phys(), "const", A=As[k])
                assert rel_l2(dH[k], o.SIA2D(Hs[k], g, tg)) <= TOL[dtype], k
                assert rel_l2(vH[k], o.VJP_dSIA_dH_discrete(lams[k], Hs[k], g, tg)) <= TOL[dtype], k
                assert _S_close(S[k], lams[k], Hs[k], g, tg, dtype), k
        assert np.array_equal(ens.download(0, _capi.FIELD_H), marker.astype(ens.np_dtype))
        only_dH, none_v, none_S = ens.fwd_adj_batch(Hs, None, want_vjpH=False, want_S=False)
        assert none_v is None and none_S is None
        # (dH above came out of the fused F1 + A1 + A2 kernel, whose D is rounded in a different order)
        assert all(rel_l2(a, b) <= (1e-6 if dtype == "f32" else 1e-14) for a, b in zip(only_dH, dH))
    finally:
        sim.close()


def _oracle_S_continuous(lam, H, g, ph, A):
    """Σ λ ⊙ pad(∇·(avg(∂A_spatial)·clamp(∇S))) -- the scalar of adjoint.jl:582-662 for a glacier-wide law: the oracle
    returns vjp_θ·S, so evaluate it with a per-glacier scalar law and divide by its (known) vjp_θ."""
    th = np.array([0.3])
    tg = o.TargetA(ph, "scalar")
    tg.apply_laws(None, None, th)
    tg.precompute_vjp(th)
    out = o.VJP_dSIA_dtheta_continuous(lam, H, g, tg, th)
    return float(out[0] / np.atleast_1d(tg.vjp_theta)[0])


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("maker", list(MAKERS))
def test_continuous_vjps(ob, dtype, maker):
    """A1c / A2c (ContinuousVJP, adjoint.jl:442-662) against the oracle."""
    nx, ny = 61, 47
    g = MAKERS[maker](nx, ny)
    lam = np.random.default_rng(11).standard_normal((nx, ny))
    g, H, lam = _round_inputs(g, g.H0, lam, dtype)
    ph = o.Phys()
    tg = o.TargetA(ph, "const", A=A0)
    sim = _sim(ob, g, {}, A0, dtype)
    try:
        vH, _ = ob.VJP_λ_dSIAdH(ob.ContinuousVJP(), lam, H, None, sim, 0.0)
        ref = o.VJP_dSIA_dH_continuous(lam, H, g, tg)
        assert rel_l2(vH, ref) <= TOL[dtype]
        assert not vH[0, :].any() and not vH[-1, :].any() and not vH[:, 0].any() and not vH[:, -1].any()  # border 0
        S = ob.VJP_λ_dSIAdθ(ob.ContinuousVJP(), lam, H, None, None, sim, 0.0)
        refS = _oracle_S_continuous(lam, H, g, ph, A0)
        assert abs(S - refS) <= 10 * TOL[dtype] * abs(refS)
    finally:
        sim.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_continuous_vjp_generic_physics_and_gridded_A(ob, dtype):
    nx, ny = 38, 52
    rng = np.random.default_rng(21)
    g = o.rough_bed_glacier(nx, ny)
    lam = rng.standard_normal((nx, ny))
    g, H, lam = _round_inputs(g, g.H0, lam, dtype)
    tol = TOL[dtype] * (4 if dtype == "f32" else 1)
    # sliding + real Glen exponent
    kw = dict(C=7e-8, n=3.3, eta0=0.6)
    sim = _sim(ob, g, kw, A0, dtype)
    try:
        tg = o.TargetA(o.Phys(**kw), "const", A=A0)
        vH, _ = ob.VJP_λ_dSIAdH(ob.ContinuousVJP(), lam, H, None, sim, 0.0)
        assert rel_l2(vH, o.VJP_dSIA_dH_continuous(lam, H, g, tg)) <= tol
        S = ob.VJP_λ_dSIAdθ(ob.ContinuousVJP(), lam, H, None, None, sim, 0.0)
        refS = _oracle_S_continuous(lam, H, g, o.Phys(**kw), A0)
        assert abs(S - refS) <= 10 * tol * abs(refS)
    finally:
        sim.close()
    # gridded A (dual grid)
    Af = A0 * np.exp(rng.uniform(-1, 1, size=(nx - 1, ny - 1)))
    if dtype == "f32":
        Af = Af.astype(np.float32).astype(np.float64)
    sim = _sim(ob, g, {}, [Af], dtype)
    try:
        tg = o.TargetA(o.Phys(), "const", A=Af)
        vH, _ = ob.VJP_λ_dSIAdH(ob.ContinuousVJP(), lam, H, None, sim, 0.0)
        assert rel_l2(vH, o.VJP_dSIA_dH_continuous(lam, H, g, tg)) <= TOL[dtype]
        # continuous θ-VJP with a gridded A: the per-node integrand, identical to the discrete one (adjoint.jl:646-657 is the
        # transpose of the contraction at adjoint.jl:250)
        from odinn_b200 import _capi

        S = ob.VJP_λ_dSIAdθ(ob.ContinuousVJP(), lam, H, None, None, sim, 0.0)
        tgg = o.TargetA(o.Phys(), "gridded")
        thg = np.arctanh(2 * (Af - o.Phys().minA) / (o.Phys().maxA - o.Phys().minA) - 1)
        ref = np.asarray(o.VJP_dSIA_dtheta_continuous(lam, H, g, tgg, thg)).reshape(Af.shape, order="F")
        tgg.precompute_vjp(thg)
        got = sim.ensemble.download(0, _capi.FIELD_VJP_A) * tgg.vjp_theta.reshape(Af.shape)
        assert rel_l2(got, ref) <= TOL[dtype] * 5
        assert abs(S - sim.ensemble.download(0, _capi.FIELD_VJP_A).astype(np.float64).sum()) <= 1e-4 * abs(S) + 1e-30
    finally:
        sim.close()


def test_continuous_vjp_resident_ragged(ob):
    """Whole ragged ensemble in one launch (flags bit2) == per-glacier oracle."""
    from odinn_b200 import _capi

    rng = np.random.default_rng(31)
    shapes = [(int(rng.integers(10, 80)), int(rng.integers(10, 80))) for _ in range(6)] + [(3, 3), (33, 4)]
    gl, Hs, lams = [], [], []
    for k, (nx, ny) in enumerate(shapes):
        g = (o.rough_bed_glacier if k % 2 else _tilted_dome)(nx, ny)
        gl.append(g), Hs.append(g.H0), lams.append(rng.standard_normal((nx, ny)))
    sim = ob.Simulation([ob.Glacier2D(B=g.B, Δx=g.dx, Δy=g.dy) for g in gl], ob.Phys(), A=A0, dtype="f64")
    try:
        ens = sim.ensemble
        for k in range(len(gl)):
            ens.upload(k, _capi.FIELD_H, Hs[k])
            ens.upload(k, _capi.FIELD_LAMBDA, lams[k])
        S = ens.vjp_resident(True, True, continuous=True)
        tg = o.TargetA(o.Phys(), "const", A=A0)
        for k, g in enumerate(gl):
            assert rel_l2(ens.download(k, _capi.FIELD_VJP_H), o.VJP_dSIA_dH_continuous(lams[k], Hs[k], g, tg)) <= 1e-12, k
            refS = _oracle_S_continuous(lams[k], Hs[k], g, o.Phys(), A0)
            assert abs(S[k] - refS) <= 1e-11 * abs(refS) or refS == 0.0, k
    finally:
        sim.close()


def test_long_chunk_work_items_of_big_ensembles(ob):
    """Big fp32 ensembles run the whole-ensemble F1 / fused-stage launches over a second work-item table with ~125-row chunks
    (capi.cu: fewer warm-up steps and halo rows per cell).  900 glaciers of 130 x 260 (5400 long items against 13500 short ones):
    the whole-ensemble launch must equal the per-glacier call (which walks the short table) bit for bit, and the oracle within the
    fp32 bound; the SSPRK3 loop through the long table must match the oracle's."""
    from odinn_b200 import _capi

    G, nx, ny = 900, 130, 260
    gl = [o.rough_bed_glacier(nx, ny), o.dome_glacier(nx, ny, H0=180.0)]
    gl[0].H0 = 0.4 * gl[0].H0   # (thin enough for the explicit loop below)
    ens = ob.Ensemble([nx] * G, [ny] * G, [gl[0].dx] * G, [gl[0].dy] * G, ob.Phys(), "f32")
    try:
        A = 2.21e-18
        for k in range(G):
            g = gl[k % 2]
            ens.upload(k, _capi.FIELD_B, g.B)
            ens.upload(k, _capi.FIELD_H, g.H0)
            ens.upload(k, _capi.FIELD_H0, g.H0)
            ens.set_A_scalar(k, A * (1.0 + 0.001 * (k % 7)))
        ens.rhs_resident()
        for k in (0, 1, 450, 899):
            g = gl[k % 2]
            Ak = A * (1.0 + 0.001 * (k % 7))
            whole = ens.download(k, _capi.FIELD_DH)
            single = ens.sia2d_rhs(k, g.H0.astype(np.float32))
            assert np.array_equal(whole, single), k
            g32 = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy)
            ref = o.SIA2D(g.H0.astype(np.float32).astype(np.float64), g32, o.TargetA(o.Phys(), "const", A=Ak))
            assert rel_l2(whole, ref) <= 1e-5, (k, rel_l2(whole, ref))
        # A1 / A2 over the long table (vjp_resident) against the per-glacier calls (short table): A1 equal up to the last bit of a few cells
        # of the rows next to a chunk seam (measured: 10 - 16 cells per glacier, 1 ulp, rel-L2 6e-9), S to fp32 summation order
        rng = np.random.default_rng(3)
        lam = [rng.standard_normal((nx, ny)).astype(np.float32) for _ in range(2)]
        for k in range(G):
            ens.upload(k, _capi.FIELD_LAMBDA, lam[k % 2])
        S = ens.vjp_resident(True, True)
        for k in (0, 1, 899):
            g = gl[k % 2]
            whole = ens.download(k, _capi.FIELD_VJP_H)
            single = ens.sia2d_vjp_H(k, lam[k % 2], g.H0.astype(np.float32))
            assert rel_l2(whole, single) <= 1e-7 and np.count_nonzero(whole != single) <= 64, (k, rel_l2(whole, single))
            S1 = ens.sia2d_vjp_theta(k, lam[k % 2], g.H0.astype(np.float32))
            assert S[k] == pytest.approx(S1, rel=1e-5, abs=1e-6 * abs(S).max()), (k, S[k], S1)
        Sf = ens.vjp_resident(True, True, want_dH=True)   # the fused step (TMA bands) on the same inputs
        assert np.allclose(Sf, S, rtol=1e-5, atol=1e-6 * abs(S).max())
        t = np.array([2010.0, 2010.0 + 1.0 / 24.0, 2010.0 + 2.0 / 24.0])
        ens.solve_forward(t, method="ssprk3", nsub=8)
        # discrete-adjoint reverse loop over the long table against a 2-glacier ensemble (short table) with the same data
        small = ob.Ensemble([nx] * 2, [ny] * 2, [gl[0].dx] * 2, [gl[0].dy] * 2, ob.Phys(), "f32")
        try:
            snaps = [[ens.get_snapshot(k, j) for j in range(len(t))] for k in range(2)]
            for k in range(2):
                small.upload(k, _capi.FIELD_B, gl[k].B)
                small.set_A_scalar(k, A * (1.0 + 0.001 * (k % 7)))
                for j in range(len(t)):
                    small.set_snapshot(k, j, len(t), snaps[k][j])
                    small.set_reference(k, j, len(t), 0.95 * snaps[k][j], snaps[k][j] > 0)
            for k in range(G):
                for j in range(len(t)):
                    ens.set_reference(k, j, len(t), 0.95 * snaps[k % 2][j], snaps[k % 2][j] > 0)
            loss_b, S_b = ens.grad_discrete(t)
            loss_s, S_s = small.grad_discrete(t)
            for k in range(2):   # glaciers 0 and 1 of the big ensemble carry the same A as the small one's
                assert loss_b[k] == pytest.approx(loss_s[k], rel=1e-5), (k, loss_b[k], loss_s[k])
                assert S_b[k] == pytest.approx(S_s[k], rel=1e-4), (k, S_b[k], S_s[k])
        finally:
            small.close()
        t = t[:2]
        for k in (1, 898):
            g = gl[k % 2]
            Ak = A * (1.0 + 0.001 * (k % 7))
            g32 = o.Glacier(B=g.B.astype(np.float32).astype(np.float64), dx=g.dx, dy=g.dy)
            Hs = o.solve_forward(g.H0.astype(np.float32).astype(np.float64), g32, o.TargetA(o.Phys(), "const", A=Ak), None, t, method="ssprk3", nsub=8)
            assert np.isfinite(Hs[1]).all()
            assert rel_l2(ens.get_snapshot(k, 1), Hs[1]) <= 2e-5, (k, rel_l2(ens.get_snapshot(k, 1), Hs[1]))
    finally:
        ens.close()
