#!/usr/bin/env python
"""bench.py -- cell-steps/s of the SIA2D forward + discrete-adjoint hot path on a 500x500xN ensemble.

One STEP = one pass of the hot path over the whole resident ensemble: for every glacier one forward
RHS evaluation (F1) and one discrete VJP pair (A1 + A2), i.e. what one saved time step of the reference's
gradient costs (src/inverse/SIA2D/gradient.jl:235-246 plus the RHS the integrator evaluates).
A "cell-step" is one grid cell through F1 + A1 + A2: 10 words of algorithmic traffic as three separate passes
(3 + 4 + 3), 5 words (read lambda, H, B; write dH, dSIA/dH^T lambda) in the fused kernel the step
launches -- the roofline is computed from the 5 words the launched kernel has to move.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Under torchrun every rank owns its own ensemble (weak scaling: --glaciers per GPU).  Every min(--allreduce-every, steps)
steps (61 saved steps per optimiser iteration in the reference's 5-year monthly configuration) the per-glacier S of the
step is read back, pulled through the 1-16-16-1 creep law (odinn_law_A_nn_pullback: d(theta) = sum_g dA_g/d(theta) S_g) and
the REAL [loss; d(theta)] (1 + 321 doubles) is summed over the ranks with one NCCL all-reduce -- inside the timed region,
its latency also reported separately.  Rank 0 prints ONE JSON line.

Besides the per-step numbers the line carries `e2e_grad`: one SIA2D_grad!-shaped optimiser iteration (gradient.jl:6-31)
through the public API -- H0 and theta in from pinned host memory, 60 tstop intervals of SSPRK3 forward + the discrete
adjoint reverse loop on the device, [loss; d(theta)] out -- and `f64`: the same step in the reference's default precision.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-steps/sec SIA2D fwd+adj"
UNIT = "cell-steps/s"
BYTES_PER_WORD = {"f32": 4, "f64": 8}
A0 = 2.21e-18  # test/test_grad_loss.jl:157


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synthetic_glacier(nx, ny, k):
    """Sloped rough bed + parabolic ice cap (case 2), varied slightly per glacier."""
    dx = 50.0
    x = np.arange(nx)[:, None] * dx
    y = np.arange(ny)[None, :] * dx
    ph = 0.37 * k
    B = 2000.0 + 0.15 * x + 30.0 * np.sin(2 * np.pi * x / 1500.0 + ph) * np.cos(2 * np.pi * y / 1100.0 - ph)
    L = min(nx, ny) * dx
    r = np.sqrt((x - 0.5 * nx * dx) ** 2 + (y - 0.5 * ny * dx) ** 2)
    H = np.maximum(0.0, (250.0 + 5.0 * (k % 7)) * (1.0 - (r / (0.42 * L)) ** 2))
    return B, H, dx


def bind_to_gpu_numa_node(device):
    """Pin this rank to the CPUs next to its GPU before the pinned host buffers are allocated (first-touch NUMA placement):
    the e2e arm is PCIe / host-memory bound, and N ranks sharing one socket's memory halve each other's copy rate."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample: brackets the timed region (the sampler itself starts earlier, nvidia-smi needs ~0.3 s to come up)."""
        return len(self.rows)

    def stop(self, i0=0, i1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows[i0:(None if i1 is None else i1 + 1)] or self.rows[i0:] or self.rows
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the C oracle (oracle/sia2d_c.c) on all host cores.  Test infrastructure, used here only
# as the timed CPU baseline (`cpu_baseline`) and as the `--impl reference` arm.
# ---------------------------------------------------------------------------------------------
class CpuArm:
    """F1 + A1 + A2 of a bounded sample of the workload through the C restatement, every buffer preallocated and the
    ctypes argument lists prepared once, so that a pass is nothing but 2 x n_glaciers C calls (OpenMP over all cores)."""

    def __init__(self, nx, ny, dtype, n_glaciers):
        from oracle import sia2d_c as oc
        from oracle import sia2d_numpy as onp

        self.oc = oc
        self.cores = oc.use_all_cores()  # all host cores, even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)
        npdt = np.float32 if dtype == "f32" else np.float64
        lib = oc.lib()
        self.f_rhs = lib.sia2d_rhs_f64 if dtype == "f64" else lib.sia2d_rhs_f32
        self.f_vjp = lib.sia2d_vjp_f64 if dtype == "f64" else lib.sia2d_vjp_f32
        ph = onp.Phys()
        rng = np.random.default_rng(1234)
        self.keep, self.calls = [], []
        vp = C.c_void_p
        for k in range(n_glaciers):
            B, H, dx = synthetic_glacier(nx, ny, k)
            arr = [np.asfortranarray(H, npdt), np.asfortranarray(B, npdt), np.asfortranarray(rng.standard_normal((nx, ny)), npdt),
                   np.zeros((nx, ny), npdt, order="F"), np.zeros((nx, ny), npdt, order="F"), np.zeros(4 * (nx - 1) * (ny - 1), npdt)]
            par, _ = oc._par(dx, dx, ph, A0, npdt)
            S = C.c_double(0.0)
            self.keep.append((arr, par, S))
            h, b, l, o1, o2, w = [vp(a.ctypes.data) for a in arr]
            self.calls.append(((C.c_int(nx), C.c_int(ny), h, b, o1, w, C.byref(par)),
                               (C.c_int(nx), C.c_int(ny), l, h, b, o2, C.byref(S), vp(None), w, C.byref(par))))
        self.cells_per_pass = nx * ny * n_glaciers

    def one_pass(self):
        for a_rhs, a_vjp in self.calls:
            self.f_rhs(*a_rhs)
            self.f_vjp(*a_vjp)

    def measure(self, steps, min_seconds, warm_steps=3, warm_seconds=1.5, max_blocks=25):
        """Warm the OpenMP team, the pages and the clocks (>= warm_steps passes and >= warm_seconds), then time blocks of
        `steps` passes until >= min_seconds and >= 5 blocks (at most max_blocks) have run; the BEST block is reported
        (best of 5 after warm-up).  Returns (cell-steps/s, seconds of the best block, blocks)."""
        t0 = time.perf_counter()
        n = 0
        while n < warm_steps or time.perf_counter() - t0 < warm_seconds:
            self.one_pass()
            n += 1
        best, blocks, t_start = float("inf"), 0, time.perf_counter()
        while blocks < max_blocks and (blocks < 5 or time.perf_counter() - t_start < min_seconds):
            t1 = time.perf_counter()
            for _ in range(steps):
                self.one_pass()
            best = min(best, time.perf_counter() - t1)
            blocks += 1
        return self.cells_per_pass * steps / best, best, blocks


def cpu_grad_iteration(n, dtype, n_glaciers=1, n_t=61, nsub=8):
    """One SIA2D_grad!-shaped iteration on the CPU arm: forward SSPRK3 solve over the tstops + the discrete-adjoint reverse
    loop, as two C loops per glacier (oracle/sia2d_c_impl.h: sia2d_solve_fixed, sia2d_grad_discrete).  Returns
    (cell-steps/s with a cell-step = one cell through one saved step, seconds, loss)."""
    from oracle import sia2d_c as oc
    from oracle import sia2d_numpy as onp

    oc.use_all_cores()
    npdt = np.float32 if dtype == "f32" else np.float64
    ph = onp.Phys()
    t = np.linspace(2010.0, 2015.0, n_t)
    data = []
    for k in range(n_glaciers):
        B, H, dx = synthetic_glacier(n, n, k)
        H = 0.6 * H  # (GRAD_THIN of the GPU arm)
        Href = oc.solve_fixed(H, B, dx, dx, ph, 1.5 * A0, t[:3], nsub=nsub, dtype=npdt)  # (also warms the thread team)
        masks = [onp.is_in_glacier(Href[-1], 3)] * n_t
        data.append((B, H, dx, [Href[-1]] * n_t, masks))
    t0 = time.perf_counter()
    loss = 0.0
    for B, H, dx, Href, masks in data:
        Hs = oc.solve_fixed(H, B, dx, dx, ph, A0, t, nsub=nsub, dtype=npdt)
        ell, Ssum, _ = oc.grad_discrete(B, dx, dx, ph, A0, t, Hs, Href, masks, dtype=npdt)
        loss += ell
    el = time.perf_counter() - t0
    return n * n * n_glaciers * (n_t - 1) / el, el, loss


def run_reference(args, rank, world):
    if rank != 0:
        return
    arm = CpuArm(args.grid, args.grid, args.dtype, args.ref_glaciers)
    rate, best, blocks = arm.measure(args.steps, 2.0, warm_steps=max(args.warmup, 1))
    sample = (f"{args.ref_glaciers} glaciers of {args.grid}x{args.grid} per step (bounded sample of the {args.glaciers}-glacier workload), "
              f"C oracle + OpenMP, best of {blocks} blocks of {args.steps} steps after a >= 1.5 s warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * best / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "upstream Julia/Huginn is not executable in this environment; the CPU arm is the C restatement (oracle/sia2d_c.c)",
    }
    if not args.no_grad:
        g_rate, g_el, _ = cpu_grad_iteration(args.grid, args.dtype, n_glaciers=args.ref_grad_glaciers)
        line["e2e_grad"] = {"value": g_rate, "unit": "cell-steps/s (one cell through one saved step of a gradient iteration)",
                            "seconds_per_iteration": g_el, "saved_steps": 60, "rhs_evals_per_saved_step": 24,
                            "sample": f"{args.ref_grad_glaciers} glacier(s) of {args.grid}x{args.grid}, 61 tstops, SSPRK3 nsub 8 forward + discrete adjoint, C loops + OpenMP",
                            "h2d_bytes_per_iteration": 0, "d2h_bytes_per_iteration": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"{args.grid}x{args.grid}x{args.glaciers} glaciers per GPU, SIA2D F1 + A1 + A2 (fwd + discrete adjoint VJPs) per step, "
                    f"glacier-wide A from the 1-16-16-1 creep law, n=3, rough sloped bed (BASELINE config 2 x N)",
        "grid": args.grid, "glaciers_per_gpu": args.glaciers, "n_gpus": world,
        "l2_policy": "inputs larger than L2 (5 planes x glaciers x grid^2 words >> 126 MB); no flush needed",
        "allreduce_every": min(args.allreduce_every, max(args.steps, 1)),
    }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
LAW_WIDTHS, LAW_ACTS = [1, 16, 16, 1], ["softplus", "softplus", "sigmoid"]   # BASELINE config 4: 2 hidden layers x 16
N_THETA = sum(o * i + o for i, o in zip(LAW_WIDTHS[:-1], LAW_WIDTHS[1:]))   # 321


# The reference's reverse loop is explicit Euler with the monthly step (gradient.jl:242): on the full 250 m cap it is unstable
# (|lambda| overflows fp32; gradient.jl:19-24 warns about exactly this), so the gradient iteration runs on 0.6 x the cap -- the work
# per iteration does not depend on the values.
GRAD_THIN = 0.6


def law_theta():
    """theta of the 1-16-16-1 law with the output bias set so that A_g = minA + (maxA - minA) sigmoid(.) stays within a factor of
    two of 2.2e-18 (test/test_grad_loss.jl:157) for every temperature: the explicit forward loop of `e2e_grad` is stable then."""
    th = 0.05 * np.random.default_rng(1).standard_normal(N_THETA)
    th[-1] = -3.55
    return th


def dominant_kernel(dtype, fused):
    if fused:
        k32 = "sia2d_vjp_march2<WRITE_F>" if os.environ.get("ODINN_MARCH", "4") == "2" else "sia2d_fused_tma (2-D TMA ring)"
        key = "sia2d_fused_march2" if (dtype == "f32" and os.environ.get("ODINN_MARCH", "4") == "2") else "sia2d_fused"
        return (k32 if dtype == "f32" else "sia2d_vjp_march<double, WRITE_F>") + " (F1 + A1 + A2 fused: one launch per step)", key, 5
    return ("sia2d_vjp_march2 (A1+A2 fused)" if dtype == "f32" else "sia2d_vjp_march (A1+A2 fused)"), ("sia2d_vjp_march2" if dtype == "f32" else "sia2d_vjp_march"), 4


def measured_traffic(dtype, kernel, cells):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from this round's committed `ncu --set full` capture
    (profiles/r02_traffic.json names the capture), scaled to this run's cells per launch; None when there is no capture of this
    dtype / kernel -- DRAM counters cannot be read from inside an unprofiled run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        k = t[dtype][kernel]
        return (k["dram_read_bytes"] + k["dram_write_bytes"]) * cells / t["cells_per_launch"], t.get("source")
    except Exception:
        return None, None


class ResidentArm:
    """A resident ensemble of G glaciers of n x n and the step  F1 + A1 + A2  on it (device-timed)."""

    def __init__(self, ob, args, dtype, rank, local_rank, with_host_buffers):
        import torch
        from odinn_b200 import _capi

        self.ob, self.torch, self.dtype = ob, torch, dtype
        G, n = args.glaciers, args.grid
        self.G, self.n = G, n
        self.w = BYTES_PER_WORD[dtype]
        npdt = np.float32 if dtype == "f32" else np.float64
        tdt = torch.float32 if dtype == "f32" else torch.float64
        self.ens = ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype, local_rank)
        rng = np.random.default_rng(1234 + rank)
        nvariants = min(G, 8)
        # host matrices: (ny, nx) row-major == column-major (nx, ny)
        alloc = (lambda: torch.empty((G, n, n), dtype=tdt).pin_memory()) if with_host_buffers else (lambda: torch.empty((nvariants, n, n), dtype=tdt))
        self.hH, self.hL = alloc(), alloc()
        for k in range(nvariants):
            B, H, _ = synthetic_glacier(n, n, k + 8 * rank)
            ens.upload(k, _capi.FIELD_B, B)
            self.hH[k].copy_(torch.from_numpy(np.ascontiguousarray(H.T.astype(npdt))))
            self.hL[k].copy_(torch.from_numpy(rng.standard_normal((n, n)).astype(npdt)))
        for k in range(nvariants, G):
            ens.upload(k, _capi.FIELD_B, ens.download(k % nvariants, _capi.FIELD_B))
            if with_host_buffers:
                self.hH[k].copy_(self.hH[k % nvariants])
                self.hL[k].copy_(self.hL[(k * 5 + 3) % nvariants])
        for k in range(G):
            kh, kl = (k, k) if with_host_buffers else (k % nvariants, (k * 5 + 3) % nvariants if k >= nvariants else k)
            ens.upload(k, _capi.FIELD_H, self.hH[kh].numpy().T)
            ens.upload(k, _capi.FIELD_LAMBDA, self.hL[kl].numpy().T)
            ens.set_temperature(k, -20.0 + 20.0 * ((k * 7 + rank) % 31) / 31.0)   # T_g in [-20, 0)
        # A_g = minA + (maxA - minA) NN([T_g]; theta) and dA_g/d(theta) on the device (LawA f! and its pullback, Laws.jl:348-362)
        self.A = ens.law_A_nn_apply(LAW_WIDTHS, LAW_ACTS, law_theta())
        self.stream = torch.cuda.ExternalStream(ens.stream_ptr, device=local_rank)
        self.cells = G * n * n

    def step(self, no_fuse=False):
        # F1 + A1 + A2 of every glacier: dH, (dSIA/dH)^T lambda and S -- ONE fused kernel (the adjoint pass recomputes every
        # forward intermediate, so dH costs one more store); --no-fuse: an F1 launch + an A1+A2 launch.
        if no_fuse:
            self.ens.rhs_resident()
            self.ens.vjp_resident(True, True, read_S=False)
        else:
            self.ens.vjp_resident(True, True, read_S=False, want_dH=True)

    def step_loss_grad_local(self, no_fuse=False):
        """The step at an optimiser-iteration boundary: the same launches, then [loss; d(theta)] of this rank from the step's
        per-glacier S: S is read back (G doubles) and pulled through the law (d(theta) = sum_g dA_g/d(theta) S_g,
        Model.jl:208-224) -- the buffer the reference reduces over its workers."""
        if no_fuse:
            self.ens.rhs_resident()
            S = self.ens.vjp_resident(True, True, read_S=True)
        else:
            S = self.ens.vjp_resident(True, True, read_S=True, want_dH=True)
        return float(np.abs(S).sum()), self.ens.law_A_nn_pullback(N_THETA, S)

    def close(self):
        self.ens.close()


def run_b200(args, rank, local_rank, world):
    import torch

    import odinn_b200 as ob
    from odinn_b200 import _capi, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    dist = None
    json_fd = 1
    if world > 1:
        import torch.distributed as dist

        # stdout carries ONE JSON line: NCCL (NCCL_DEBUG=VERSION in the image, INFO when the driver asks for it) logs to fd 1 while
        # the communicator comes up AND when it is destroyed, so fd 1 points at stderr for the rest of the process and the JSON line
        # is written to the saved descriptor.  NCCL_DEBUG itself is left as the caller set it.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        warm = torch.zeros(1 + N_THETA, dtype=torch.float64, device="cuda")
        dist.all_reduce(warm)  # communicator creation happens here
        torch.cuda.synchronize()

    def barrier(arm):
        arm.ens.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(arm, fn, k):
        """k calls of fn bracketed by barrier + synchronize, CUDA events on the handle's stream, max over ranks."""
        barrier(arm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(arm.stream)
        for i in range(k):
            fn(i)
        e1.record(arm.stream)
        arm.ens.synchronize()
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t0)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        return ms, wall

    G, n, dtype = args.glaciers, args.grid, args.dtype
    w = BYTES_PER_WORD[dtype]
    arm = ResidentArm(ob, args, dtype, rank, local_rank, with_host_buffers=args.e2e_steps > 0)
    ens = arm.ens
    cells_per_step = arm.cells
    every = min(args.allreduce_every, max(args.steps, 1))
    ar_ms = []

    def step(i):
        if (i + 1) % every != 0:
            arm.step(args.no_fuse)
        else:
            # the optimiser-iteration boundary: the REAL [loss; d(theta)] of this rank, then ONE all-reduce over the ranks
            loss, dth = arm.step_loss_grad_local(args.no_fuse)
            ta = time.perf_counter()
            parallel.allreduce_loss_grad(loss, dth)
            ar_ms.append(1e3 * (time.perf_counter() - ta))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i if args.warmup >= every else every - 1)   # (the collective path is warmed as well)
    ar_ms.clear()
    m0 = sampler.mark()
    l0 = ens.launch_count
    ms, _ = timed(arm, step, args.steps)
    launches = ens.launch_count - l0
    n_allreduce = len(ar_ms)
    # per-kernel durations for the roofline (same resident inputs, > L2); every launch path runs once untimed first (first-use costs of a
    # kernel -- lazy module loading, plane allocation -- showed up as a 4 ms "launch" of the F1 kernel in a 20-step 2-GPU run)
    for _ in range(2):
        ens.rhs_resident()
        ens.vjp_resident(True, True, read_S=False)
        arm.step(args.no_fuse)
    ens.synchronize()
    ms_fused = timed(arm, lambda i: arm.step(args.no_fuse), args.steps)[0] / args.steps
    ms_rhs = timed(arm, lambda i: ens.rhs_resident(), args.steps)[0] / args.steps
    ms_vjp = timed(arm, lambda i: ens.vjp_resident(True, True, read_S=False), args.steps)[0] / args.steps
    # e2e: the reference-facing batched call with pinned HOST buffers, copies inside the timed region
    e2e_steps = min(args.steps, args.e2e_steps)
    ms_e2e, esz = float("inf"), n * n * w
    if e2e_steps > 0:  # (--e2e-steps 0: profiling runs that want the resident launches only)
        hdH, hV = torch.empty_like(arm.hH).pin_memory(), torch.empty_like(arm.hH).pin_memory()
        ptr = lambda t: (C.c_void_p * G)(*[t.data_ptr() + k * esz for k in range(G)])
        pH, pL, pdH, pV = ptr(arm.hH), ptr(arm.hL), ptr(hdH), ptr(hV)
        S = np.zeros(G)
        if args.batch_chunk > 0:
            ens.set_batch_chunk(args.batch_chunk)
        ens.fwd_adj_batch_host(pH, pL, pdH, pV, S)
        ev, wall = timed(arm, lambda i: ens.fwd_adj_batch_host(pH, pL, pdH, pV, S), e2e_steps)
        ms_e2e = max(ev, wall) / e2e_steps
    clocks = sampler.stop(m0, sampler.mark()) if rank == 0 else None

    value = world * cells_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * cells_per_step / (ms_e2e * 1e-3) if e2e_steps > 0 else None
    peak, peak_src = load_peaks()
    fused = (not args.no_fuse and os.environ.get("ODINN_NO_FUSE") != "1"
             and (dtype == "f64" or os.environ.get("ODINN_MARCH", "4") in ("2", "4")))
    sub = lambda nbytes, ms_k, words, wd=w: {"achieved": nbytes / (ms_k * 1e-3) / 1e9, "frac": nbytes / (ms_k * 1e-3) / 1e9 / peak,
                                             "algorithmic_bytes_per_cell": words * wd, "ms_per_launch": ms_k}
    dom_name, dom_key, dom_words = dominant_kernel(dtype, fused)
    dom_ms = ms_fused if fused else ms_vjp
    dom = sub(dom_words * w * cells_per_step, dom_ms, dom_words)
    traffic, traffic_src = measured_traffic(dtype, dom_key, cells_per_step)

    # ---- the same step in the other precision (the reference's default element type is Float64, test/SIA2D_adjoint_utils.jl:22) ----
    other = None
    if not args.no_other_dtype:
        od = "f64" if dtype == "f32" else "f32"
        arm2 = ResidentArm(ob, args, od, rank, local_rank, with_host_buffers=False)
        for i in range(args.warmup):
            arm2.step()
        ms2 = timed(arm2, lambda i: arm2.step(), args.steps)[0] / args.steps
        w2 = BYTES_PER_WORD[od]
        other = {"dtype": od, "value": world * arm2.cells / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
                 "roofline": dict(bound="hbm", kernel=dominant_kernel(od, True)[0], peak=peak, unit="GB/s",
                                  **sub(5 * w2 * arm2.cells, ms2, 5, w2))}
        arm2.close()

    # ---- e2e_grad: one SIA2D_grad!-shaped optimiser iteration through the public API (gradient.jl:6-31) ----
    e2e_grad = None
    if not args.no_grad:
        e2e_grad = grad_iteration_arm(ob, parallel, args, dtype, rank, local_rank, world, dist)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * G * esz), "d2h_bytes_per_step": int(2 * G * esz + 8 * G),
                    "api": "odinn_fwd_adj_batch_host (pinned host H, lambda -> dH, vjp_H, S)", "steps": e2e_steps},
            "e2e_grad": e2e_grad,
            "allreduce": {"what": "[loss; d(theta)] = 1 + 321 float64, odinn_law_A_nn_pullback of the step's S then torch.distributed all_reduce(SUM)",
                          "every_steps": every, "count_in_timed_region": n_allreduce, "backend": "nccl" if world > 1 else "none (1 rank)",
                          "ms_each_host_blocking": (sum(ar_ms) / len(ar_ms)) if ar_ms else None},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_cell": dom_words * w, "ms_per_launch": dom_ms,
                         "vjp_kernel": sub(4 * w * cells_per_step, ms_vjp, 4), "rhs_kernel": sub(3 * w * cells_per_step, ms_rhs, 3),
                         "step_launches": "1 fused kernel (+ the per-glacier S reduction)" if fused else "F1 kernel + A1+A2 kernel (+ the per-glacier S reduction)",
                         "note": "ms_per_launch is the step timed without the [loss; d(theta)] read-back / all-reduce of the iteration boundary"},
        }
        if other is not None:
            line[other["dtype"]] = other
        if world == 1 and not args.no_cpu:
            cpu = CpuArm(n, n, dtype, args.ref_glaciers)
            rate, best, blocks = cpu.measure(args.steps, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cpu.cores, "kind": "port",
                                    "sample": f"{args.ref_glaciers} glaciers of {n}x{n} per step, best of {blocks} blocks of {args.steps} steps "
                                              f"({best:.2f} s per block) after a >= 1.5 s warm-up, C oracle + OpenMP, {dtype}"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    arm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def grad_iteration_arm(ob, parallel, args, dtype, rank, local_rank, world, dist):
    """One optimiser iteration as the reference runs it (SIA2D_grad!, src/inverse/SIA2D/gradient.jl:6-31): theta and H0 in from
    (pinned) host memory, law -> forward solve over 61 monthly tstops (SSPRK3, nsub 8: 1440 RHS) -> DiscreteAdjoint reverse loop
    (60 x A1 + loss/seed + A2) -> law pullback on the device, [loss; d(theta)] back, all-reduced over the ranks.  H_ref is data:
    resident, uploaded once before the timed region.  A cell-step here = one cell through one saved step."""
    import torch
    from odinn_b200 import _capi

    G, n = args.grad_glaciers, args.grid
    npdt = np.float32 if dtype == "f32" else np.float64
    t = np.linspace(2010.0, 2015.0, 61)
    ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), dtype, local_rank)
    nvar = min(G, 4)
    hH0 = torch.empty((G, n, n), dtype=torch.float32 if dtype == "f32" else torch.float64).pin_memory()
    for k in range(G):
        if k < nvar:
            B, H, _ = synthetic_glacier(n, n, k + 4 * rank)
            H = GRAD_THIN * H
            ens.upload(k, _capi.FIELD_B, B)
            hH0[k].copy_(torch.from_numpy(np.ascontiguousarray(H.T.astype(npdt))))
        else:
            ens.upload(k, _capi.FIELD_B, ens.download(k % nvar, _capi.FIELD_B))
            hH0[k].copy_(hH0[k % nvar])
        ens.set_temperature(k, -20.0 + 20.0 * ((k * 7 + rank) % 31) / 31.0)
        ens.upload(k, _capi.FIELD_H0, hH0[k].numpy().T)
    # twin experiment: H_ref from a different law (A x 1.5), generated on the device, masks on the host (is_in_glacier)
    th = law_theta()
    th_true = th.copy()
    th_true[-1] += 0.42
    ens.law_A_nn_apply(LAW_WIDTHS, LAW_ACTS, th_true)
    ens.solve_forward(t, method="ssprk3", nsub=8)
    for k in range(nvar):
        refs = [ens.get_snapshot(k, j) for j in range(len(t))]
        masks = [ob.is_in_glacier(h, 3) for h in refs]
        for kk in range(k, G, nvar):
            for j in range(len(t)):
                ens.set_reference(kk, j, len(t), refs[j], masks[j])

    def iteration():
        for k in range(G):                                   # H0 in (pinned host -> device)
            ens.upload(k, _capi.FIELD_H0, hH0[k].numpy().T)
        ens.law_A_nn_apply(LAW_WIDTHS, LAW_ACTS, th)          # theta in; A_g and dA_g/d(theta) on the device
        ens.solve_forward(t, method="ssprk3", nsub=8)
        losses, _ = ens.grad_discrete(t)                     # loss out
        dth = ens.law_A_nn_pullback(N_THETA)                 # d(theta) out
        return parallel.allreduce_loss_grad(float(losses.sum()), dth)

    loss, dth = iteration()  # warm-up (captures the CUDA graph of the interval)
    ens.synchronize()
    if dist is not None:
        dist.barrier()
    reps = max(1, args.grad_iterations)
    l0 = ens.launch_count
    t0 = time.perf_counter()
    for _ in range(reps):
        loss, dth = iteration()
    ens.synchronize()
    sec = (time.perf_counter() - t0) / reps
    launches = (ens.launch_count - l0) // reps
    if dist is not None:
        tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())

    # The same iteration with the reference's DEFAULT integrator for the forward solve (RDPK3Sp35 + PID, AdjointTypes.jl:60; every stage
    # update fused into the F1 launch, glaciers that have landed on the tstop skipped): our arm only -- the C restatement of the CPU arm
    # has the fixed-step loops.
    rt = 1e-4 if dtype == "f32" else 1e-6
    stat = {}

    def iteration_rdpk():
        for k in range(G):
            ens.upload(k, _capi.FIELD_H0, hH0[k].numpy().T)
        ens.law_A_nn_apply(LAW_WIDTHS, LAW_ACTS, th)
        stat["steps"], stat["rejected"] = ens.solve_forward_adaptive(t, reltol=rt, abstol=rt, method="rdpk3sp35")
        losses, _ = ens.grad_discrete(t)
        dth = ens.law_A_nn_pullback(N_THETA)
        return parallel.allreduce_loss_grad(float(losses.sum()), dth)

    loss_r, _ = iteration_rdpk()
    ens.synchronize()
    if dist is not None:
        dist.barrier()
    l1 = ens.launch_count
    t0 = time.perf_counter()
    for _ in range(reps):
        loss_r, _ = iteration_rdpk()
    ens.synchronize()
    sec_r = (time.perf_counter() - t0) / reps
    launches_r = (ens.launch_count - l1) // reps
    if dist is not None:
        tt = torch.tensor([sec_r], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec_r = float(tt.item())
    rdpk = {"what": f"the same iteration with the forward solve by adaptive RDPK3Sp35 + PID (the reference's default integrator), rtol = atol = {rt:g}",
            "seconds_per_iteration": sec_r, "value": world * G * n * n * 60 / sec_r, "trial_steps_per_glacier_min": int(stat["steps"].min()),
            "trial_steps_per_glacier_max": int(stat["steps"].max()), "rejected_max": int(stat["rejected"].max()),
            "gpu_launches_per_iteration": int(launches_r), "loss": float(loss_r)}
    ens.close()
    esz = n * n * (4 if dtype == "f32" else 8)
    return {"value": world * G * n * n * 60 / sec, "unit": "cell-steps/s (one cell through one saved step of a gradient iteration)",
            "seconds_per_iteration": sec, "saved_steps": 60, "rhs_evals_per_saved_step": 24, "glaciers_per_gpu": G,
            "h2d_bytes_per_iteration": int(G * esz + 8 * N_THETA), "d2h_bytes_per_iteration": int(8 * (G + N_THETA)),
            "gpu_launches_per_iteration": int(launches), "loss": float(loss), "norm_dtheta": float(np.linalg.norm(dth)),
            "forward_rdpk3sp35": rdpk,
            "api": "Ensemble.upload(H0) + law_A_nn_apply(theta) + solve_forward + grad_discrete + law_A_nn_pullback + allreduce_loss_grad "
                   "(= odinn_b200.SIA2D_grad_), wall clock, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--grid", type=int, default=500)
    ap.add_argument("--glaciers", type=int, default=256, help="glaciers per GPU")
    ap.add_argument("--allreduce-every", type=int, default=61)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--batch-chunk", type=int, default=0, help="cells per pipeline chunk of the host-batch call (0: library default)")
    ap.add_argument("--ref-glaciers", type=int, default=16, help="glaciers per CPU pass (bounded sample)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="time F1 and A1+A2 as two launches per step")
    ap.add_argument("--no-grad", action="store_true", help="skip the e2e_grad (whole gradient iteration) measurement")
    ap.add_argument("--no-other-dtype", action="store_true", help="skip the sub-record in the other precision")
    ap.add_argument("--grad-glaciers", type=int, default=32, help="glaciers per GPU of the e2e_grad iteration")
    ap.add_argument("--grad-iterations", type=int, default=2)
    ap.add_argument("--ref-grad-glaciers", type=int, default=1, help="glaciers of the CPU arm's gradient iteration (bounded sample)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
