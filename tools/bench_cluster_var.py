"""Config 1 (one 128x128 glacier, 2010-2015, SSPRK3 nsub 8) through the cluster-resident solver for a sweep of cluster sizes.
usage: [ODINN_B200_LIB=...] python tools/bench_cluster_var.py [f32|f64]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
t5 = np.linspace(2010.0, 2015.0, 61)
for n in (128, 64, 256):
    ens = ob.Ensemble([n], [n], [50.0], [50.0], ob.Phys(minA=8e-21, maxA=8e-17), dtype)
    B, H, _ = synthetic_glacier(n, n, 0)
    ens.upload(0, _capi.FIELD_B, B); ens.upload(0, _capi.FIELD_H0, 0.5 * H); ens.set_A_scalar(0, 5e-18)
    for cs in (4, 8, 16):
        try:
            ens.set_cluster_mode(cs)
            ens.solve_forward(t5, method="ssprk3", nsub=8); ens.synchronize()
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); ens.solve_forward(t5, method="ssprk3", nsub=8); ens.synchronize(); best = min(best, time.perf_counter() - t0)
            print(json.dumps(dict(lib=os.path.basename(os.environ.get("ODINN_B200_LIB", "default")), dtype=dtype, n=n, cs=cs, ms=1e3 * best, us_per_stage=1e6 * best / 1440)), flush=True)
        except Exception as ex:
            print(json.dumps(dict(n=n, cs=cs, error=str(ex)[:100])), flush=True)
    ens.close()
