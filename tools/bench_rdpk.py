"""RDPK3Sp35 + PID forward solve (the reference's default integrator) at the bench workload, host-driven large-ensemble engine:
time per ensemble-wide trial step, fused stage epilogues (default) vs ODINN_RK_NO_FUSE=1.  usage: python tools/bench_rdpk.py [f32|f64] [G]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 256
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
t = 2010.0 + np.arange(5) / 12.0
rt = 1e-4 if dtype == "f32" else 1e-6
ens.solve_forward_adaptive(t, reltol=rt, abstol=rt, method="rdpk3sp35"); ens.synchronize()
l0 = ens.launch_count
t0 = time.perf_counter(); steps, rej = ens.solve_forward_adaptive(t, reltol=rt, abstol=rt, method="rdpk3sp35"); ens.synchronize(); s = time.perf_counter() - t0
H = ens.get_snapshot(3, len(t) - 1).astype(np.float64)
print(json.dumps(dict(what="RDPK3Sp35 + PID, 4 monthly intervals, rtol %g, %d x %dx%d" % (rt, G, n, n), dtype=dtype, fused=os.environ.get("ODINN_RK_NO_FUSE", "0") in ("", "0"),
                      seconds=s, trial_steps_max=int(steps.max()), trial_steps_min=int(steps.min()), rejected_max=int(rej.max()),
                      launches=int(ens.launch_count - l0),
                      checksum=float(H.sum()), hmax=float(H.max()))), flush=True)
ens.close()
