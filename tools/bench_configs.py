"""Wall-clock numbers for the BASELINE.json configs that bench.py's headline line does not cover (configs 1, 3, 4):
forward time loops and full gradient iterations on device-resident ensembles.  usage: python tools/bench_configs.py [f32|f64]
Prints one JSON line per measurement.  (Synthetic glaciers: SURVEY.md 8d.)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odinn_b200 as ob
from odinn_b200 import _capi
from bench import synthetic_glacier

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
PH = dict(minA=8e-21, maxA=8e-17)


def make(shapes, seed=0, thin=1.0):
    G = len(shapes)
    ens = ob.Ensemble([s[0] for s in shapes], [s[1] for s in shapes], [50.0] * G, [50.0] * G, ob.Phys(**PH), dtype)
    rng = np.random.default_rng(seed)
    for k, (nx, ny) in enumerate(shapes):
        B, H, _ = synthetic_glacier(nx, ny, k)
        ens.upload(k, _capi.FIELD_B, B)
        ens.upload(k, _capi.FIELD_H0, thin * H)
        ens.set_A_scalar(k, float(np.exp(rng.uniform(np.log(2e-18), np.log(2e-17)))))
    return ens


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best


def report(name, **kw):
    print(json.dumps(dict(config=name, dtype=dtype, **kw)), flush=True)


t5 = np.linspace(2010.0, 2015.0, 61)   # 5 years, monthly tstops

# config 1: single 128x128 glacier, forward 2010-2015
ens = make([(128, 128)], thin=0.5)
for nsub in (8,):
    s = timed(lambda: (ens.solve_forward(t5, method="ssprk3", nsub=nsub), ens.synchronize()))
    rhs = 60 * nsub * 3
    report("1: 128x128 forward 2010-2015, SSPRK3", nsub=nsub, seconds=s, rhs_evals=rhs, us_per_rhs=1e6 * s / rhs, cell_steps_per_s=128 * 128 * rhs / s)
st = None
def run_bs3():
    global st
    st = ens.solve_forward_adaptive(t5, reltol=1e-4, abstol=1e-4)
    ens.synchronize()
s = timed(run_bs3)
report("1: 128x128 forward 2010-2015, adaptive BS3 rtol 1e-4", seconds=s, steps=int(st[0][0]), rejected=int(st[1][0]))
# the reference's default integrator at the tolerance its tests construct it with (test/params_construction.jl:7; fp32: the error estimate bottoms out near 1e-4)
rt = 1e-8 if dtype == "f64" else 1e-4
def run_rdpk():
    global st
    st = ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35")
    ens.synchronize()
s = timed(run_rdpk)
report(f"1: 128x128 forward 2010-2015, adaptive RDPK3Sp35 + PID rtol {rt:g} (the reference's default solver)", seconds=s, steps=int(st[0][0]),
       rejected=int(st[1][0]), us_per_trial_step=1e6 * s / max(int(st[0][0]), 1))
# config 1 as an inversion: one optimiser iteration = forward solve + DiscreteAdjoint reverse loop (61 monthly snapshots), the twin
# experiment of test/inversion_test.jl on one small glacier
ens.solve_forward(t5, method="ssprk3", nsub=8)
for j in range(len(t5)):
    Hj = ens.get_snapshot(0, j)
    ens.set_reference(0, j, len(t5), 0.97 * Hj, (Hj > 0))
ens.set_A_scalar(0, 7e-18)
def it_ssprk3():
    ens.solve_forward(t5, method="ssprk3", nsub=8); return ens.grad_discrete(t5)
def it_rdpk():
    ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35"); return ens.grad_discrete(t5)
def it_adj():
    return ens.grad_discrete(t5)
l0 = ens.launch_count; it_ssprk3(); n_launch = ens.launch_count - l0
report("1: 128x128 inversion iteration, SSPRK3 forward + discrete adjoint", seconds=timed(it_ssprk3), launches=int(n_launch))
report("1: 128x128 inversion iteration, RDPK3Sp35 forward + discrete adjoint", seconds=timed(it_rdpk))
report("1: 128x128 discrete-adjoint reverse loop alone (60 saved steps)", seconds=timed(it_adj), us_per_saved_step=1e6 * timed(it_adj) / 60)
# the reference's DEFAULT gradient: ContinuousAdjoint, adaptive RDPK3Sp35 reverse solve, dtmax 1/12, 200 quadrature nodes (AdjointTypes.jl:53-66)
rta = 1e-8 if dtype == "f64" else 1e-5
res = None
def it_cont():
    global res
    res = ens.grad_continuous_adaptive(t5, n_quadrature=200, reltol=rta, abstol=rta)
s = timed(it_cont)
report(f"1: 128x128 continuous adjoint (adaptive reverse RDPK3Sp35 rtol {rta:g}, dtmax 1/12, 200 quadrature nodes) alone", seconds=s,
       trial_steps=int(res[2][0]), us_per_trial_step=1e6 * s / max(int(res[2][0]), 1))
def it_default():
    ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35"); return ens.grad_continuous_adaptive(t5, n_quadrature=200, reltol=rta, abstol=rta)
l0 = ens.launch_count; it_default(); n_launch = ens.launch_count - l0
report("1: 128x128 inversion iteration in the reference's DEFAULT configuration (RDPK3Sp35 forward + ContinuousAdjoint)", seconds=timed(it_default), launches=int(n_launch))
ens.close()

# config 3: 64 glaciers, sizes U{100..400}, forward Prediction run
rng = np.random.default_rng(2024)
shapes = [(int(rng.integers(100, 401)), int(rng.integers(100, 401))) for _ in range(64)]
cells = sum(a * b for a, b in shapes)
ens = make(shapes, thin=0.4)
nsub = 8
s = timed(lambda: (ens.solve_forward(t5, method="ssprk3", nsub=nsub), ens.synchronize()))
rhs = 60 * nsub * 3
report("3: 64 glaciers 100-400 px, forward 2010-2015, SSPRK3 nsub 8", seconds=s, cells=cells, rhs_evals=rhs, cell_steps_per_s=cells * rhs / s,
       frac_of_hbm_peak=cells * rhs * 4 * (4 if dtype == "f32" else 8) / s / 6550.1e9)
# the same run with the reference's default solver (host-driven engine: stage updates fused into F1, landed glaciers skipped)
st3 = None
def run3_rdpk():
    global st3
    st3 = ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35")
    ens.synchronize()
s = timed(run3_rdpk)
report(f"3: 64 glaciers 100-400 px, forward 2010-2015, adaptive RDPK3Sp35 + PID rtol {rt:g} (the reference's default solver)", seconds=s, cells=cells,
       trial_steps_min=int(st3[0].min()), trial_steps_max=int(st3[0].max()), rejected_max=int(st3[1].max()))
ens.close()

# config 4: 32 glaciers, LawA(nn 1-16-16-1), one optimiser iteration = law + forward solve + adjoint + pullback
rng = np.random.default_rng(2025)
shapes = [(int(rng.integers(100, 401)), int(rng.integers(100, 401))) for _ in range(32)]
cells = sum(a * b for a, b in shapes)
ens = make(shapes, thin=0.4)
widths, acts = [1, 16, 16, 1], ["softplus", "softplus", "sigmoid"]
nth = sum(o * i + o for i, o in zip(widths[:-1], widths[1:]))
theta = 0.3 * np.random.default_rng(1).standard_normal(nth)
for k in range(32):
    ens.set_temperature(k, float(rng.uniform(-20, 0)))
ens.law_A_nn_apply(widths, acts, theta - 1.0)
ens.solve_forward(t5, method="ssprk3", nsub=nsub)
for k in range(32):
    for j in range(len(t5)):
        Hj = ens.get_snapshot(k, j)
        ens.set_reference(k, j, len(t5), Hj, ob.is_in_glacier(Hj, 3))

def iteration(mode):
    ens.law_A_nn_apply(widths, acts, theta)
    ens.solve_forward(t5, method="ssprk3", nsub=nsub)
    if mode == "discrete":
        ens.grad_discrete(t5)
    else:
        ens.grad_continuous(t5, n_quadrature=200, vjp="discrete", method="ssprk3", nsub=1)
    return ens.law_A_nn_pullback(nth)

for mode in ("discrete", "continuous"):
    s = timed(lambda: iteration(mode), reps=2)
    report(f"4: 32 glaciers, LawA(1-16-16-1), one iteration: law + forward (SSPRK3 nsub 8) + {mode} adjoint + pullback", seconds=s, cells=cells,
           n_theta=nth)
def iteration_default():   # the reference's default configuration: RDPK3Sp35 forward + ContinuousAdjoint (adaptive reverse solve, 200 nodes)
    ens.law_A_nn_apply(widths, acts, theta)
    ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35")
    ens.grad_continuous_adaptive(t5, n_quadrature=200, reltol=rta, abstol=rta)
    return ens.law_A_nn_pullback(nth)
def iteration_rdpk_discrete():
    ens.law_A_nn_apply(widths, acts, theta)
    ens.solve_forward_adaptive(t5, reltol=rt, abstol=rt, method="rdpk3sp35")
    ens.grad_discrete(t5)
    return ens.law_A_nn_pullback(nth)
s = timed(iteration_rdpk_discrete, reps=2)
report("4: 32 glaciers, LawA(1-16-16-1), one iteration: law + forward (RDPK3Sp35) + discrete adjoint + pullback", seconds=s, cells=cells, n_theta=nth)
s = timed(iteration_default, reps=2)
report("4: 32 glaciers, LawA(1-16-16-1), one iteration in the reference's DEFAULT configuration (RDPK3Sp35 forward + ContinuousAdjoint, 200 nodes)",
       seconds=s, cells=cells, n_theta=nth)
s_f = timed(lambda: (ens.solve_forward(t5, method="ssprk3", nsub=nsub), ens.synchronize()), reps=2)
s_g = timed(lambda: ens.grad_discrete(t5), reps=2)
report("4: split", forward_seconds=s_f, discrete_adjoint_seconds=s_g, adjoint_cell_steps_per_s=cells * 60 / s_g)
ens.close()
