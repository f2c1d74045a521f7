"""Continuous-adjoint VJP kernels (A1c, A2c) at the bench workload.  usage: python tools/bench_cont_vjp.py [f32|f64]"""
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
ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype)
rng = np.random.default_rng(0)
lam = rng.standard_normal((n, n))
for k in range(G):
    if k < 4:
        B, H, _ = synthetic_glacier(n, n, k)
        ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H, H); ens.upload(k, _capi.FIELD_LAMBDA, lam)
    else:
        for f in (_capi.FIELD_B, _capi.FIELD_H, _capi.FIELD_LAMBDA):
            ens.upload(k, f, ens.download(k % 4, f))
    ens.set_A_scalar(k, 2.21e-18)
cells = G * n * n
def timed(fn, reps=20):
    fn(); ens.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    ens.synchronize(); return (time.perf_counter() - t0) / reps
for name, fn, words in (("A1c continuous VJP_H", lambda: ens.vjp_resident(True, False, read_S=False, continuous=True), 4),
                        ("A2c continuous VJP_theta (unit-A F1 + dot)", lambda: ens.vjp_resident(False, True, read_S=False, continuous=True), 6),
                        ("A1 discrete VJP_H alone", lambda: ens.vjp_resident(True, False, read_S=False), 4),
                        ("A2 discrete VJP_theta alone", lambda: ens.vjp_resident(False, True, read_S=False), 3)):
    s = timed(fn)
    print(json.dumps(dict(what=name, dtype=dtype, ms=1e3 * s, frac_of_hbm_peak=cells * words * w / s / 6550.1e9)), flush=True)
ens.close()
