"""Pins the oracle's grid primitives with the reference's own operator tests.

Protocol and tolerances are those of test/SIA2D_adjoint_utils.jl:8-126 of the
reference: <A u, v> == <u, A^T v> on 10x11 random arrays, rtol 1e-11, 5 draws
each.  The transposes are transcribed verbatim from
src/inverse/SIA2D/inversion_utils.jl:3-66, so these identities determine the
out-of-tree Huginn primitives uniquely."""
import numpy as np
import pytest

from oracle import sia2d_numpy as o

SIZE = (10, 11)
FAC = SIZE[0] * SIZE[1]
RTOL = 1e-11


def _close(a, b):
    assert a == pytest.approx(b, rel=RTOL)


@pytest.mark.parametrize("seed", range(5))
def test_adjoint_diff(seed):
    rng = np.random.default_rng(seed)
    d = 2.5
    u = rng.standard_normal(SIZE)
    v = rng.standard_normal((SIZE[0] - 1, SIZE[1]))
    Au = np.zeros((SIZE[0] - 1, SIZE[1]))
    o.diff_x_inplace(Au, u, d)
    _close(np.sum(Au * v) / FAC, np.sum(u * o.diff_x_adjoint(v, d)) / FAC)
    v = rng.standard_normal((SIZE[0], SIZE[1] - 1))
    Au = np.zeros((SIZE[0], SIZE[1] - 1))
    o.diff_y_inplace(Au, u, d)
    _close(np.sum(Au * v) / FAC, np.sum(u * o.diff_y_adjoint(v, d)) / FAC)


@pytest.mark.parametrize("seed", range(5))
def test_adjoint_clamp_borders(seed):
    rng = np.random.default_rng(100 + seed)
    d, eta0 = 2.5, 1.0
    H = np.abs(rng.standard_normal(SIZE))
    dS = rng.standard_normal((SIZE[0] - 1, SIZE[1] - 2))
    v = rng.standard_normal(dS.shape)
    c = o.clamp_borders_dx(dS, H, eta0, d)
    ddS, dH = np.zeros_like(dS), np.zeros_like(H)
    o.clamp_borders_dx_adjoint(ddS, dH, v, eta0, d, H, dS)
    _close(np.sum(c * v) / FAC, np.sum(H * dH) / FAC + np.sum(dS * ddS) / FAC)
    dS = rng.standard_normal((SIZE[0] - 2, SIZE[1] - 1))
    v = rng.standard_normal(dS.shape)
    c = o.clamp_borders_dy(dS, H, eta0, d)
    ddS, dH = np.zeros_like(dS), np.zeros_like(H)
    o.clamp_borders_dy_adjoint(ddS, dH, v, eta0, d, H, dS)
    _close(np.sum(c * v) / FAC, np.sum(H * dH) / FAC + np.sum(dS * ddS) / FAC)


@pytest.mark.parametrize("seed", range(5))
def test_adjoint_avg(seed):
    rng = np.random.default_rng(200 + seed)
    u = rng.standard_normal(SIZE)
    v = rng.standard_normal((SIZE[0] - 1, SIZE[1] - 1))
    _close(np.sum(o.avg(u) * v) / FAC, np.sum(u * o.avg_adjoint(v)) / FAC)
    v = rng.standard_normal((SIZE[0] - 1, SIZE[1]))
    _close(np.sum(o.avg_x(u) * v) / FAC, np.sum(u * o.avg_x_adjoint(v)) / FAC)
    v = rng.standard_normal((SIZE[0], SIZE[1] - 1))
    _close(np.sum(o.avg_y(u) * v) / FAC, np.sum(u * o.avg_y_adjoint(v)) / FAC)


def test_shapes_follow_reference():
    """Shapes quoted in SURVEY §8(a) F1/F2 and adjoint.jl:58-97."""
    A = np.zeros(SIZE)
    assert o.diff_x(A).shape == (9, 11) and o.diff_y(A).shape == (10, 10)
    assert o.avg(A).shape == (9, 10) and o.inn(A).shape == (8, 9) and o.inn1(A).shape == (9, 10)
    assert o.clamp_borders_dx(np.zeros((9, 9)), A, 1.0, 1.0).shape == (9, 9)
    assert o.clamp_borders_dy(np.zeros((8, 10)), A, 1.0, 1.0).shape == (8, 10)


def test_continuous_theta_vjp_is_the_transpose_of_the_discrete_contraction():
    """adjoint.jl:646-657 (dense tensor dD/dtheta through avg, the CLAMPED edge slopes, the divergence, then contracted with lambda)
    == adjoint.jl:250 (sum_nodes dD/dtheta D_adj) for every law kind: glacier-wide, gridded (one parameter per node), per-cell
    networks.  The device relies on it to serve the continuous theta-VJP of gridded-A and per-cell laws with the discrete A2 kernels."""
    ph = o.Phys()
    g = o.rough_bed_glacier(23, 19)
    lam = np.random.default_rng(1).standard_normal(g.B.shape)
    mlp = o.MLP([2, 8, 8, 1], ["softplus", "softplus", "sigmoid"])
    thm = 0.4 * np.random.default_rng(3).standard_normal(mlp.n_params)
    cases = [
        (o.TargetA(ph, "scalar"), np.array([0.3])),
        (o.TargetA(ph, "gridded"), 0.1 * np.random.default_rng(2).standard_normal((22, 18))),
        (o.TargetD(ph, mlp, prescale_bounds=((0.0, 300.0), (0.0, 0.6)), max_NN=60.0), thm),
        (o.TargetDHybrid(ph, mlp, T=-7.0), thm),
        (o.TargetDHybrid(ph, mlp, T=-7.0, interpolation="Linear", n_interp_half=8), thm),
    ]
    for tg, th in cases:
        a = o.VJP_dSIA_dtheta_continuous_dense(lam, g.H0, g, tg, th)
        b = np.asarray(o.VJP_dSIA_dtheta_discrete(lam, g.H0, g, tg, th)).reshape(-1)
        assert a.shape == b.shape and np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(b), type(tg).__name__
