"""ContinuousAdjoint gradient (the reference's default: adaptive RDPK3Sp35 reverse ODE, dtmax 1/12; gradient.jl:276-538) on a LARGE ensemble
through the host-driven engine of rdpk.cu: interpolation + A1 + stage update fused into one launch per stage (default) vs
ODINN_RK_NO_FUSE=1 (interpolation pass, A1 pass, elementwise stage pass).  usage: python tools/bench_contadj.py [f32|f64] [G] [n_quadrature]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n = 500
ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype)
for k in range(G):
    if k < 4:
        B, H, _ = synthetic_glacier(n, n, k)
        ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H0, 0.4 * H)
    else:
        ens.upload(k, _capi.FIELD_B, ens.download(k % 4, _capi.FIELD_B)); ens.upload(k, _capi.FIELD_H0, ens.download(k % 4, _capi.FIELD_H0))
    ens.set_A_scalar(k, 2.21e-18 * (1 + 0.01 * k))
ens.set_cluster_mode(0)
t = 2010.0 + np.arange(4) / 12.0
ens.solve_forward(t, method="ssprk3", nsub=8)
ref = [[None] * len(t) for _ in range(4)]
for k in range(G):
    for j in range(len(t)):
        if k < 4:
            Hj = ens.get_snapshot(k, j)
            ref[k][j] = (0.97 * Hj, (Hj > 0) / float(n * n))
        ens.set_reference(k, j, len(t), ref[k % 4][j][0], ref[k % 4][j][1] > 0)
rt = 1e-5 if dtype == "f32" else 1e-8
run = lambda: ens.grad_continuous_adaptive(t, n_quadrature=nq, reltol=rt, abstol=rt)
run(); ens.synchronize()
l0 = ens.launch_count
t0 = time.perf_counter(); loss, S, steps = run(); ens.synchronize(); s = time.perf_counter() - t0
print(json.dumps(dict(what="ContinuousAdjoint, adaptive RDPK3Sp35 reverse solve rtol %g, dtmax 1/12, %d quadrature nodes, %d tstops, %d x %dx%d" % (rt, nq, len(t), G, n, n),
                      dtype=dtype, fused=os.environ.get("ODINN_RK_NO_FUSE", "0") in ("", "0"), seconds=s, trial_steps_max=int(steps.max()),
                      trial_steps_min=int(steps.min()), launches=int(ens.launch_count - l0), loss0=float(loss[0]), S0=float(S[0]), loss_sum=float(loss.sum()), S_sum=float(S.sum()))), flush=True)
ens.close()
