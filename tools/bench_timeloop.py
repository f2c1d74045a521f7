"""Throughput of the device-resident time loop at the bench workload (500x500x256 glaciers): SSPRK3 stages fused into F1 (4 words/cell),
the discrete-adjoint reverse step, the continuous-adjoint stage.  usage: python tools/bench_timeloop.py [f32|f64]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G, n = 256, 500
w = 4 if dtype == "f32" else 8
ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(minA=8e-21, maxA=8e-17), dtype)
for k in range(G):
    if k < 4:
        B, H, _ = synthetic_glacier(n, n, k)
        ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H0, 0.4 * H)
    else:
        ens.upload(k, _capi.FIELD_B, ens.download(k % 4, _capi.FIELD_B)); ens.upload(k, _capi.FIELD_H0, ens.download(k % 4, _capi.FIELD_H0))
    ens.set_A_scalar(k, 2.21e-18 * (1 + 0.01 * k))
cells = G * n * n
t = 2010.0 + np.arange(5) / 12.0
nsub = 8
def fwd():
    ens.solve_forward(t, method="ssprk3", nsub=nsub); ens.synchronize()
fwd()
t0 = time.perf_counter(); fwd(); s = time.perf_counter() - t0
rhs = (len(t) - 1) * nsub * 3
print(json.dumps(dict(what="forward SSPRK3, stage fused into F1", dtype=dtype, ms_per_rhs=1e3 * s / rhs, cell_steps_per_s=cells * rhs / s,
                      frac_of_hbm_peak_4w=cells * rhs * 4 * w / s / 6550.1e9)), flush=True)
for k in range(G):
    for j in range(len(t)):
        pass
# reference data = the snapshots themselves shifted (device-side copies would do; host round trip keeps the script short for 4 glaciers only)
for j in range(len(t)):
    for k in range(4):
        Hj = ens.get_snapshot(k, j)
        for kk in range(k, G, 4):
            ens.set_reference(kk, j, len(t), 0.98 * Hj, (Hj > 0))
def grad():
    return ens.grad_discrete(t)
grad()
t0 = time.perf_counter(); grad(); s = time.perf_counter() - t0
steps = len(t) - 1
print(json.dumps(dict(what="discrete-adjoint reverse loop (A1 + loss/seed + A2 per saved step)", dtype=dtype, ms_per_step=1e3 * s / steps,
                      cell_steps_per_s=cells * steps / s)), flush=True)
def gradc():
    return ens.grad_continuous(t, n_quadrature=8, vjp="discrete", method="ssprk3", nsub=1)
gradc()
t0 = time.perf_counter(); gradc(); s = time.perf_counter() - t0
print(json.dumps(dict(what="continuous adjoint, 8 nodes + 5 tstops, SSPRK3 nsub 1", dtype=dtype, seconds=s)), flush=True)
ens.close()
