"""Throughput of the per-cell MLP law path (LawU, 2-16-16-1: BASELINE config 4's per-cell variant) at 500x500 x G glaciers:
F1 (node pass + stencil in D-field mode), A1 (node pass with the reference's finite-difference partials + stencil), A2 (per-node back-propagation).
usage: python tools/bench_cell_law.py [f32|f64] [G]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = 500
ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype)
rng = np.random.default_rng(0)
for k in range(G):
    B, H, _ = synthetic_glacier(n, n, k % 4)
    ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H, H); ens.upload(k, _capi.FIELD_LAMBDA, rng.standard_normal((n, n)))
widths, acts = [2, 16, 16, 1], ["softplus", "softplus", "sigmoid"]
nth = sum(o * i + o for i, o in zip(widths[:-1], widths[1:]))
ens.law_cell_nn_set("U", widths, acts, 0.4 * rng.standard_normal(nth), prescale_bounds=((0.0, 300.0), (0.0, 0.5)), max_NN=50.0)
cells = G * n * n
def timed(fn, reps=5):
    fn(); ens.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    ens.synchronize()
    return (time.perf_counter() - t0) / reps
for name, fn in (("F1 (law nodes + stencil)", lambda: ens.rhs_resident()),
                 ("A1 (law nodes with FD partials + stencil)", lambda: ens.vjp_resident(True, False, read_S=False)),
                 ("A1 + A2 (+ per-node back-propagation)", lambda: ens.vjp_resident(True, True, read_S=False))):
    s = timed(fn)
    print(json.dumps(dict(what=name, law="LawU 2-16-16-1", dtype=dtype, glaciers=G, ms=1e3 * s, cell_evals_per_s=cells / s)), flush=True)
ens.close()
