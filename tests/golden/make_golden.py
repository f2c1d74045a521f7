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


if __name__ == "__main__":
    main()
