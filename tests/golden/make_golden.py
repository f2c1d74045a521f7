#!/usr/bin/env python
"""Generate tests/golden/*.npz from the NumPy oracle (oracle/sia2d_numpy.py).

The reference holds no golden vectors for this path and Julia is not available, so these are NOT
reference outputs: they freeze the pinned oracle (operator identities + the reference's finite-difference protocol +
Halfar, see tests/test_oracle_*.py) so that the C oracle and the CUDA path are compared against stored numbers and an
accidental change of the oracle itself is caught.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sia2d_numpy as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
A0 = 2.21e-18  # test/test_grad_loss.jl:157


def case(name, g, H, lam, ph, A):
    tg = o.TargetA(ph, "const", A=A)
    f = o._recompute_forward(H, g, tg, None)
    _, _, Dadj = o._D_adjoint(lam, f, g.dx, g.dy)
    out = dict(
        B=g.B, H=H, lam=lam, dx=g.dx, dy=g.dy, A=A,
        phys=np.array([ph.rho, ph.g, ph.eta0, ph.n, ph.p, ph.q, ph.C]),
        dH=o.SIA2D(H, g, tg),
        vjp_H=o.VJP_dSIA_dH_discrete(lam, H, g, tg),
        S=o.node_reduction_S(lam, H, g, tg),
        vjp_H_cont=o.VJP_dSIA_dH_continuous(lam, H, g, tg),
        D=f["D"], D_adjoint=Dadj,
    )
    # The same operators evaluated (in fp64) on the fp32-ROUNDED inputs: what an fp32 kernel must be compared with.
    # The clamp sub-gradient is discontinuous (strict inequalities, inversion_utils.jl:22-43), so rounding the inputs
    # can flip branches on clamp-active edges; comparing an fp32 kernel with the fp64-input vectors would be ill-posed
    # (the noisy case moves by 1.5e-2 from input rounding alone).
    r32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
    g32 = o.Glacier(B=r32(g.B), dx=g.dx, dy=g.dy)
    A32 = r32(A) if np.ndim(A) == 2 else A
    tg32 = o.TargetA(ph, "const", A=A32)
    out.update(
        dH_f32in=o.SIA2D(r32(H), g32, tg32),
        vjp_H_f32in=o.VJP_dSIA_dH_discrete(r32(lam), r32(H), g32, tg32),
        S_f32in=o.node_reduction_S(r32(lam), r32(H), g32, tg32),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, H.shape, "sum|dH| = %.6e  S = %.6e" % (np.abs(out["dH"]).sum(), out["S"]))


def main():
    rng = np.random.default_rng(1234)
    # 1. rough sloped bed (clamp active on the margin), config-1 family, small
    g = o.rough_bed_glacier(48, 37)
    case("rough_48x37", g, g.H0, rng.standard_normal(g.B.shape), o.Phys(), A0)
    # 2. Halfar dome on a flat bed
    g = o.dome_glacier(40, 40)
    case("dome_40x40", g, g.H0, rng.standard_normal(g.B.shape), o.Phys(), A0)
    # 3. noisy thickness with holes and negative input, odd sizes
    g = o.rough_bed_glacier(33, 29)
    H = g.H0 * (1.0 + 0.3 * rng.standard_normal(g.H0.shape))
    H[rng.random(H.shape) < 0.05] = 0.0
    H[rng.random(H.shape) < 0.02] = -3.0
    case("noisy_33x29", g, H, rng.standard_normal(g.B.shape), o.Phys(), A0)
    # 4. sliding + non-integer Glen exponent + eta0 != 1
    g = o.rough_bed_glacier(31, 44)
    case("generic_31x44", g, g.H0, rng.standard_normal(g.B.shape), o.Phys(C=7e-8, n=3.3, eta0=0.6), A0)
    # 5. gridded A on the dual grid
    g = o.rough_bed_glacier(36, 36)
    Af = A0 * np.exp(rng.uniform(-1, 1, size=(35, 35)))
    case("griddedA_36x36", g, g.H0, rng.standard_normal(g.B.shape), o.Phys(), Af)
    # 6. LawA(nn): A and dA/dtheta for the default 1-3-10-3-1 network at T = -12.5
    mlp = o.MLP.default(1)
    theta = mlp.init(seed=666)
    tg = o.TargetA(o.Phys(), "nn", mlp=mlp, T=-12.5)
    tg.apply_laws(None, None, theta)
    tg.precompute_vjp(theta)
    np.savez_compressed(os.path.join(HERE, "lawA_default_nn.npz"), theta=theta, T=-12.5, A=tg.A, dA_dtheta=tg.vjp_theta,
                        widths=np.array(mlp.widths))
    print("lawA_default_nn A = %.6e" % tg.A)


def next_rows():
    """SURVEY 8f rows (N1-N4) on one small glacier: adaptive BS3 run, surface velocity + its VJPs, mass balance + its VJP,
    discrete (with MB) and continuous adjoint gradients."""
    rng = np.random.default_rng(4321)
    g = o.rough_bed_glacier(24, 21)
    g.H0 = 0.6 * g.H0
    ph = o.Phys(minA=8e-21, maxA=8e-17)
    A, Aref = 1.5e-17, 4e-17
    t = o.define_callback_steps((2010.0, 2010.5), 1.0 / 12.0)
    tg, tgr = o.TargetA(ph, "const", A=A), o.TargetA(ph, "const", A=Aref)
    st = {}
    H_bs3 = o.solve_forward(g.H0, g, tg, None, t, method="bs3", reltol=1e-6, abstol=1e-6, stats=st)
    Vx, Vy = o.surface_V(g.H0, g, tg)
    dVx, dVy = rng.standard_normal(g.B.shape), rng.standard_normal(g.B.shape)
    par = (3.0, -0.0065, 2100.0, 0.9, 0.4, 1.2, 1.0)
    MB, _, _, _ = o.mb_TI1(g.H0, g.B, par)
    lam = rng.standard_normal(g.B.shape)
    mb = {j: par for j in (2, 4, 6)}
    stm = {}
    Href = o.solve_forward(g.H0, g, tgr, None, t, method="ssprk3", nsub=8, mb=mb)
    Hs = o.solve_forward(g.H0, g, tg, None, t, method="ssprk3", nsub=8, mb=mb, stats=stm)
    tgs = o.TargetA(ph, "scalar")
    theta = np.array([np.arctanh(2 * (A - ph.minA) / (ph.maxA - ph.minA) - 1)])
    wH, wV = o.loss_weights("H", t)
    ell_d, dth_d = o.loss_and_grad_discrete_HV(theta, g, tgs, t, Hs, Href, [None] * len(t), wH, wV, mb=mb, MB_hist=stm["MB"])
    Href0 = o.solve_forward(g.H0, g, tgr, None, t, method="ssprk3", nsub=8)
    Hs0 = o.solve_forward(g.H0, g, tg, None, t, method="ssprk3", nsub=8)
    ell_c, dth_c = o.loss_and_grad_continuous(theta, g, tgs, t, Hs0, Href0, n_quadrature=9, nsub=2, method="ssprk3", vjp="discrete")
    np.savez_compressed(
        os.path.join(HERE, "next_24x21.npz"), B=g.B, H0=g.H0, dx=g.dx, dy=g.dy, A=A, Aref=Aref, t=t, minA=ph.minA, maxA=ph.maxA,
        H_bs3_end=H_bs3[-1], bs3_nrhs=st["nrhs"], Vx=Vx, Vy=Vy, dVx=dVx, dVy=dVy,
        vjpV_H=o.VJP_dsurfaceV_dH_discrete(dVx, dVy, g.H0, g, tg), vjpV_S=-o.surfaceV_theta_reduction(dVx, dVy, g.H0, g, tg),
        mb_par=np.array(par), MB=MB, lam=lam, vjp_MB=o.VJP_MB_dH(lam, g.H0, g.B, par), mb_idx=np.array(sorted(mb)),
        Hs_mb_end=Hs[-1], loss_discrete_mb=ell_d, dtheta_discrete_mb=dth_d[0], vjp_theta=tgs.vjp_theta[0],
        loss_continuous=ell_c, dtheta_continuous=dth_c[0])
    print("next_24x21 bs3 nrhs = %d  loss_d = %.6e  dθ_d = %.6e  dθ_c = %.6e" % (st["nrhs"], ell_d, dth_d[0], dth_c[0]))


if __name__ == "__main__":
    main()
    next_rows()
