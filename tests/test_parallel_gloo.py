"""Host-side multi-GPU logic on CPU: ensemble sharding and the [loss; dθ] all-reduce, world_size 2 over gloo.

The per-rank "compute" is the NumPy oracle (test infrastructure) so that the test needs no GPU; what is under test
is odinn_b200.parallel: the shards partition the ensemble, and the reduced loss/gradient equal the serial sums
(the reference's sum(losses) + aggregate∇θ, src/inverse/SIA2D/gradient.jl:13-30, Model.jl:208-224)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["ODINN_ROOT"])
import importlib.util
spec = importlib.util.spec_from_file_location("odinn_parallel", os.path.join(os.environ["ODINN_ROOT"], "odinn.jl_b200", "parallel.py"))
par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
from oracle import sia2d_numpy as o

dist = par.init_process_group("gloo")
rank, _, ws = par.world()
shapes = [(14, 17), (20, 12), (9, 25), (16, 16), (11, 13)]
temps = [-12.0, -3.0, -8.0, -15.0, -1.0]
ph = o.Phys(minA=8e-21, maxA=8e-17)
mlp = o.MLP.default(1, light=True)
th = mlp.init(5, scale=0.7)
t = o.define_callback_steps((2010.0, 2010.25), 1.0 / 12.0)

def one(k):
    g = o.rough_bed_glacier(*shapes[k]); g.H0 = 0.5 * g.H0
    Href = o.solve_forward(g.H0, g, o.TargetA(ph, "const", A=4e-17), None, t, method="ssprk3", nsub=8)
    tg = o.TargetA(ph, "nn", mlp=mlp, T=temps[k])
    Hs = o.solve_forward(g.H0, g, tg, th, t, method="ssprk3", nsub=8)
    ell, dth, _ = o.loss_and_grad_discrete(th, g, tg, t, Hs, Href)
    return ell, dth

costs = [a * b for a, b in shapes]
mine = par.shard_glaciers(costs, ws)[rank]
loss, grad = 0.0, np.zeros(mlp.n_params)
for k in mine:
    l, d = one(k); loss += l; grad += d
loss, grad = par.allreduce_loss_grad(loss, grad)
per_glacier = par.scatter_per_glacier(np.array([float(k + 1) for k in mine]), mine, len(shapes))
if rank == 0:
    sl, sg = 0.0, np.zeros(mlp.n_params)
    for k in range(len(shapes)):
        l, d = one(k); sl += l; sg += d
    print(json.dumps({"loss": loss, "serial_loss": sl, "grad_err": float(np.linalg.norm(grad - sg) / np.linalg.norm(sg)),
                      "per_glacier": per_glacier.tolist()}))
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_glaciers_partitions_and_balances():
    sys.path.insert(0, ROOT)
    import importlib.util

    spec = importlib.util.spec_from_file_location("odinn_parallel", os.path.join(ROOT, "odinn.jl_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    rng = np.random.default_rng(2024)
    costs = (rng.integers(100, 401, size=64) * rng.integers(100, 401, size=64)).tolist()  # config 3 sizes
    for ws in (1, 2, 4, 8):
        bins = par.shard_glaciers(costs, ws)
        assert sorted(sum(bins, [])) == list(range(64))
        loads = [sum(costs[k] for k in b) for b in bins]
        assert max(loads) <= 1.1 * (sum(costs) / ws)
        assert bins == par.shard_glaciers(costs, ws)  # deterministic
    assert par.shard_glaciers([5.0], 4) == [[0], [], [], []]
    assert par.allreduce_loss_grad(1.5, np.arange(3.0))[0] == 1.5  # identity without a process group


def test_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ODINN_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert abs(d["loss"] - d["serial_loss"]) <= 1e-12 * abs(d["serial_loss"])
    assert d["grad_err"] < 1e-12
    assert d["per_glacier"] == [1.0, 2.0, 3.0, 4.0, 5.0]
