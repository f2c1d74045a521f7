"""Adaptive BS3 forward solve at the bench workload: time per accepted/trial step and its split.  usage: python tools/bench_adaptive.py [f32|f64]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G, n = 256, 500
ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype)
for k in range(G):
    if k < 4:
        B, H, _ = synthetic_glacier(n, n, k)
        ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H0, 0.4 * H)
    else:
        ens.upload(k, _capi.FIELD_B, ens.download(k % 4, _capi.FIELD_B)); ens.upload(k, _capi.FIELD_H0, ens.download(k % 4, _capi.FIELD_H0))
    ens.set_A_scalar(k, 2.21e-18 * (1 + 0.01 * k))
t = 2010.0 + np.arange(5) / 12.0
rt = 1e-4
ens.solve_forward_adaptive(t, reltol=rt, abstol=rt); ens.synchronize()
l0 = ens.launch_count
t0 = time.perf_counter(); steps, rej = ens.solve_forward_adaptive(t, reltol=rt, abstol=rt); ens.synchronize(); s = time.perf_counter() - t0
print(json.dumps(dict(what="adaptive BS3, 4 monthly intervals, rtol 1e-4", dtype=dtype, seconds=s, trial_steps_max=int(steps.max()), trial_steps_min=int(steps.min()),
                      rejected_max=int(rej.max()), launches=int(ens.launch_count - l0), ms_per_ensemble_step=1e3 * s / max(int(steps.max()), 1))), flush=True)
ens.close()
