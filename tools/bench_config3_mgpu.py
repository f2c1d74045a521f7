"""BASELINE config 3 / 4 across GPUs: the 64-glacier ensemble (mixed 100-400 px grids) sharded over the ranks of a torchrun job
(`Prediction`, forward run, no collective) and the 32-glacier training iteration (`SIA2D_grad_`, one all-reduce of [loss; dθ]).
usage: [torchrun --nproc-per-node N] python tools/bench_config3_mgpu.py      -> one JSON line per measurement on rank 0."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import odinn_b200 as ob
from odinn_b200 import parallel
from bench import synthetic_glacier

rank, local_rank, ws = parallel.world()
torch.cuda.set_device(local_rank)
dist = parallel.init_process_group()


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


def glaciers(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        nx, ny = int(rng.integers(100, 401)), int(rng.integers(100, 401))
        B, H, dx = synthetic_glacier(nx, ny, k)
        out.append(ob.Glacier2D(B=B, Δx=dx, Δy=dx, H0=0.4 * H))
    return out


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        barrier(); t0 = time.perf_counter(); fn(); barrier(); best = min(best, time.perf_counter() - t0)
    return best


ph = ob.Phys(minA=8e-21, maxA=8e-17)
params = ob.Parameters(physical=ph, tspan=(2010.0, 2015.0), dtype="f32", solver=ob.SolverParameters(solver="ssprk3", nsub=8))
# config 3: forward Prediction run
gl = glaciers(64, 2024)
pred = ob.Prediction(ob.Model(ob.SIA2Dmodel(A=None)), gl, params)
rngA = np.random.default_rng(5)
pred.set_A(list(np.exp(rngA.uniform(np.log(2e-18), np.log(2e-17), size=64))))
s = timed(lambda: (pred.solve(), pred.ensemble.synchronize()))
cells = sum(g.nx * g.ny for g in gl)
rhs = 60 * 8 * 3
if rank == 0:
    print(json.dumps(dict(config="3: 64 glaciers 100-400 px, forward 2010-2015 (Prediction), SSPRK3 nsub 8", n_gpus=ws, seconds=s,
                          cell_steps_per_s=cells * rhs / s, glaciers_on_rank0=len(pred.glaciers))), flush=True)
pred.close()
# config 4: one training iteration (law + forward + discrete adjoint + pullback + all-reduce)
gl = glaciers(32, 2025)
temps = list(np.random.default_rng(6).uniform(-20, 0, size=32))
pred = ob.Prediction(ob.Model(ob.SIA2Dmodel(A=None)), gl, params)
pred.set_A([ph.minA + (ph.maxA - ph.minA) * (0.15 + 0.02 * (T + 20.0)) for T in temps])
href_local = ob.run_(pred)
pred.close()
# every rank needs the reference list indexed by global glacier id: fill the others with placeholders (never read on this rank)
H_ref = [None] * 32
for k, gid in enumerate(pred.my_ids):
    H_ref[gid] = href_local[k]
nn = ob.NeuralNetwork(widths=(1, 16, 16, 1), acts=("softplus", "softplus", "sigmoid"), seed=3)
inv = ob.Inversion(ob.Model(ob.SIA2Dmodel(A=ob.LawA(nn))), gl, params, H_ref, temperatures=temps)
θ = np.array(inv.model.θ)
θ[-1] = -3.0   # A_g within 2e-18 .. 7e-18: the monthly reverse Euler step of the discrete adjoint is stable there (gradient.jl:19-24)
g = np.zeros_like(θ)
s = timed(lambda: ob.SIA2D_grad_(g, θ, inv), reps=2)
if rank == 0:
    print(json.dumps(dict(config="4: 32 glaciers, LawA(1-16-16-1): SIA2D_grad! (forward + discrete adjoint + all-reduce)", n_gpus=ws, seconds=s,
                          grad_norm=float(np.linalg.norm(g)))), flush=True)
inv.close()
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
