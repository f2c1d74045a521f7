"""Config 4 / config 3 forward solve only (marching kernels + CUDA graph), short horizon: for per-kernel durations under ncu and
chunk-row sweeps.  usage: python tools/bench_config4_fwd.py [f32|f64] [n_glaciers] [n_intervals]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nint = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rng = np.random.default_rng(2025 if G == 32 else 2024)
shapes = [(int(rng.integers(100, 401)), int(rng.integers(100, 401))) for _ in range(G)]
cells = sum(a * b for a, b in shapes)
ens = ob.Ensemble([s[0] for s in shapes], [s[1] for s in shapes], [50.0] * G, [50.0] * G, ob.Phys(minA=8e-21, maxA=8e-17), dtype)
for k, (nx, ny) in enumerate(shapes):
    B, H, _ = synthetic_glacier(nx, ny, k)
    ens.upload(k, _capi.FIELD_B, B); ens.upload(k, _capi.FIELD_H0, 0.4 * H); ens.set_A_scalar(k, 5e-18)
t = 2010.0 + np.arange(nint + 1) / 12.0
run = lambda: (ens.solve_forward(t, method="ssprk3", nsub=8), ens.synchronize())
run()
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); run(); best = min(best, time.perf_counter() - t0)
rhs = nint * 24
print(json.dumps(dict(G=G, dtype=dtype, cells=cells, seconds=best, us_per_stage=1e6 * best / rhs, cell_steps_per_s=cells * rhs / best,
                      chunk_rows2=os.environ.get("ODINN_CHUNK_ROWS2", "auto"))), flush=True)
ens.close()
