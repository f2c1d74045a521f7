"""The C oracle (timed CPU baseline) must agree with the NumPy oracle (semantic reference)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import sia2d_c as c
from oracle import sia2d_numpy as o

A0 = 2.21e-18


@pytest.mark.parametrize("shape", [(3, 3), (14, 17), (33, 47), (128, 96)])
@pytest.mark.parametrize("phys_kw", [{}, dict(C=7e-8), dict(n=3.5, eta0=0.4)])
def test_c_oracle_matches_numpy_f64(shape, phys_kw):
    g = o.rough_bed_glacier(*shape)
    ph = o.Phys(**phys_kw)
    tg = o.TargetA(ph, "const", A=A0)
    lam = np.random.default_rng(1234).standard_normal(shape)
    assert rel_l2(c.rhs(g.H0, g.B, g.dx, g.dy, ph, A0), o.SIA2D(g.H0, g, tg)) < 1e-13
    v, S, _ = c.vjp(lam, g.H0, g.B, g.dx, g.dy, ph, A0)
    assert rel_l2(v, o.VJP_dSIA_dH_discrete(lam, g.H0, g, tg)) < 1e-13
    ref = o.node_reduction_S(lam, g.H0, g, tg)
    assert abs(S - ref) <= 1e-12 * abs(ref)


def test_c_oracle_gridded_A_and_f32():
    shape = (41, 37)
    g = o.rough_bed_glacier(*shape)
    ph = o.Phys()
    rng = np.random.default_rng(3)
    Af = A0 * np.exp(rng.uniform(-1, 1, size=(shape[0] - 1, shape[1] - 1)))
    lam = rng.standard_normal(shape)
    tg = o.TargetA(ph, "const", A=Af)
    assert rel_l2(c.rhs(g.H0, g.B, g.dx, g.dy, ph, Af), o.SIA2D(g.H0, g, tg)) < 1e-13
    v, S, fld = c.vjp(lam, g.H0, g.B, g.dx, g.dy, ph, Af, want_field=True)
    assert rel_l2(v, o.VJP_dSIA_dH_discrete(lam, g.H0, g, tg)) < 1e-13
    assert abs(fld.sum() - S) <= 1e-12 * abs(S)
    # Float32 build of the same code (Sleipnir.doublePrec = false): loose agreement only
    v32, S32, _ = c.vjp(lam, g.H0, g.B, g.dx, g.dy, ph, A0, dtype=np.float32)
    v64, S64, _ = c.vjp(lam, g.H0, g.B, g.dx, g.dy, ph, A0)
    assert rel_l2(v32, v64) < 5e-3 and abs(S32 - S64) < 5e-3 * abs(S64)
