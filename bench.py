#!/usr/bin/env python
"""bench.py -- cell-steps/s of the SIA2D forward + discrete-adjoint hot path on a 500x500xN ensemble.

One STEP = one pass of the hot path over the whole resident ensemble: for every glacier one forward
RHS evaluation (F1) and one discrete VJP pair (A1 + A2), i.e. what one saved time step of the reference's
gradient costs (src/inverse/SIA2D/gradient.jl:235-246 plus the RHS the integrator evaluates).
A "cell-step" is one grid cell through F1 + A1 + A2: 10 words of algorithmic traffic.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Under torchrun every rank owns its own ensemble (weak scaling: --glaciers per GPU); the loss/gradient
all-reduce (NCCL) runs once every --allreduce-every steps (61 saved steps per optimiser iteration in the
reference's 5-year monthly configuration).  Rank 0 prints ONE JSON line.
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


def measured_traffic(dtype, kernel, cells):
    """dram read+write bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json), scaled to
    this run's cells per launch; None when no capture exists for this dtype / kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        k = t[dtype][kernel]
        return (k["dram_read_bytes"] + k["dram_write_bytes"]) * cells / t["cells_per_launch"]
    except Exception:
        return None


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
# as the timed CPU baseline.
# ---------------------------------------------------------------------------------------------
def cpu_fwd_adj(nx, ny, dtype, n_glaciers, min_seconds, reps_max=10**9):
    from oracle import sia2d_c as oc
    from oracle import sia2d_numpy as onp

    oc.use_all_cores()  # all host cores, even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)
    npdt = np.float32 if dtype == "f32" else np.float64
    ph = onp.Phys()
    data = []
    rng = np.random.default_rng(1234)
    for k in range(n_glaciers):
        B, H, dx = synthetic_glacier(nx, ny, k)
        lam = rng.standard_normal((nx, ny))
        data.append((np.asfortranarray(H, npdt), np.asfortranarray(B, npdt), np.asfortranarray(lam, npdt), dx))

    def one_pass():
        for H, B, lam, dx in data:
            oc.rhs(H, B, dx, dx, ph, A0, dtype=npdt)
            oc.vjp(lam, H, B, dx, dx, ph, A0, dtype=npdt)

    one_pass()  # warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        one_pass()
        reps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or reps >= reps_max:
            break
    cells = nx * ny * n_glaciers * reps
    return cells / el, el, reps, oc.threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step_glaciers = args.ref_glaciers
    # warm-up passes then K timed steps; each step = one pass over a bounded sample of the same workload
    rate_w, _, _, cores = cpu_fwd_adj(args.grid, args.grid, args.dtype, per_step_glaciers, 0.0, reps_max=max(args.warmup, 1))
    t0 = time.perf_counter()
    rate, el, reps, cores = cpu_fwd_adj(args.grid, args.grid, args.dtype, per_step_glaciers, 1e9, reps_max=args.steps)
    sample = f"{per_step_glaciers} glaciers of {args.grid}x{args.grid} per step (bounded sample of the {args.glaciers}-glacier workload), C oracle + OpenMP"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / max(reps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "upstream Julia/Huginn is not executable in this environment; the CPU arm is the C restatement (oracle/sia2d_c.c)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"{args.grid}x{args.grid}x{args.glaciers} glaciers per GPU, SIA2D F1 + A1 + A2 (fwd + discrete adjoint VJPs) per step, "
                    f"scalar A, n=3, rough sloped bed",
        "grid": args.grid, "glaciers_per_gpu": args.glaciers, "n_gpus": world,
        "l2_policy": "inputs larger than L2 (5 planes x glaciers x grid^2 words >> 126 MB); no flush needed",
        "allreduce_every": args.allreduce_every,
    }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch

    import odinn_b200 as ob
    from odinn_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    # stdout carries ONE JSON line.  The image exports NCCL_DEBUG=VERSION, which makes NCCL print "NCCL version ..." on stdout when
    # the communicator is created: drop it (ODINN_NCCL_DEBUG re-enables NCCL logging) and create the communicator with fd 1 -> stderr.
    os.environ.pop("NCCL_DEBUG", None)
    if "ODINN_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["ODINN_NCCL_DEBUG"]
    dist = None
    if world > 1:
        import torch.distributed as dist

        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)  # communicator creation happens here
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    G, n = args.glaciers, args.grid
    npdt = np.float32 if args.dtype == "f32" else np.float64
    w = BYTES_PER_WORD[args.dtype]
    ens = ob.Ensemble([n] * G, [n] * G, [50.0] * G, [50.0] * G, ob.Phys(), args.dtype, local_rank)
    rng = np.random.default_rng(1234 + rank)
    # pinned host buffers: inputs H, lambda; outputs dH, vjpH (the reference-facing host arrays)
    hH = torch.empty((G, n, n), dtype=torch.float32 if w == 4 else torch.float64).pin_memory()
    hL = torch.empty_like(hH).pin_memory()
    hdH = torch.empty_like(hH).pin_memory()
    hV = torch.empty_like(hH).pin_memory()
    nvariants = min(G, 8)
    for k in range(G):
        if k < nvariants:
            B, H, _ = synthetic_glacier(n, n, k + 8 * rank)
            ens.upload(k, _capi.FIELD_B, B)
            hH[k].copy_(torch.from_numpy(np.ascontiguousarray(H.T.astype(npdt))))  # (ny, nx) row-major == column-major (nx, ny)
            hL[k].copy_(torch.from_numpy(rng.standard_normal((n, n)).astype(npdt)))
        else:
            ens.upload(k, _capi.FIELD_B, ens.download(k % nvariants, _capi.FIELD_B))
            hH[k].copy_(hH[k % nvariants])
            hL[k].copy_(hL[(k * 5 + 3) % nvariants])
        ens.set_A_scalar(k, A0 * (1.0 + 0.01 * k))
    esz = hH.element_size() * n * n
    ptr = lambda t: (C.c_void_p * G)(*[t.data_ptr() + k * esz for k in range(G)])
    pH, pL, pdH, pV = ptr(hH), ptr(hL), ptr(hdH), ptr(hV)
    S = np.zeros(G)
    # resident planes for the device-timed arm (the host-batch call works on its own staging planes)
    for k in range(G):
        ens.upload(k, _capi.FIELD_H, hH[k].numpy().T)
        ens.upload(k, _capi.FIELD_LAMBDA, hL[k].numpy().T)
    if args.batch_chunk > 0:
        ens.set_batch_chunk(args.batch_chunk)
    if args.e2e_steps > 0:
        ens.fwd_adj_batch_host(pH, pL, pdH, pV, S)

    stream = torch.cuda.ExternalStream(ens.stream_ptr, device=local_rank)
    cells_per_step = G * n * n

    def barrier():
        ens.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    vjp_theta = torch.zeros(322, dtype=torch.float64, device="cuda")  # [loss; dθ] of the 1-16-16-1 law (321 params)

    def step(i):
        # F1 + A1 + A2 of every glacier: dH, (dSIA/dH)^T lambda and S -- ONE fused kernel (the adjoint pass recomputes
        # every forward intermediate, so dH costs one more store); --no-fuse: an F1 launch + an A1+A2 launch.
        if args.no_fuse:
            ens.rhs_resident()
            ens.vjp_resident(True, True, read_S=False)
        else:
            ens.vjp_resident(True, True, read_S=False, want_dH=True)
        if dist is not None and (i + 1) % args.allreduce_every == 0:
            ens.synchronize()
            dist.all_reduce(vjp_theta)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(k):
            fn(i)
        e1.record(stream)
        ens.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    m0 = sampler.mark()
    l0 = ens.launch_count
    ms = timed(step, args.steps)
    launches = ens.launch_count - l0
    # per-kernel durations for the roofline (same resident inputs, > L2)
    ms_rhs = timed(lambda i: ens.rhs_resident(), args.steps) / args.steps
    ms_vjp = timed(lambda i: ens.vjp_resident(True, True, read_S=False), args.steps) / args.steps
    # e2e: the reference-facing batched call with pinned HOST buffers, copies inside the timed region
    e2e_steps = min(args.steps, args.e2e_steps)
    ms_e2e = float("inf")
    if e2e_steps > 0:  # (--e2e-steps 0: profiling runs that want the resident launches only)
        ens.fwd_adj_batch_host(pH, pL, pdH, pV, S)
        ms_e2e = timed(lambda i: ens.fwd_adj_batch_host(pH, pL, pdH, pV, S), e2e_steps) / e2e_steps
    clocks = sampler.stop(m0, sampler.mark()) if rank == 0 else None

    value = world * cells_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * cells_per_step / (ms_e2e * 1e-3) if e2e_steps > 0 else None
    peak, peak_src = load_peaks()
    # Roofline of the dominant kernel: the fused F1 + A1 + A2 kernel -- 5 words/cell (read λ, H, B; write dH, ∂H);
    # --no-fuse: the A1+A2 kernel -- 4 words/cell.  The F1 and A1+A2 kernels timed alone are reported beside it.
    fused = (not args.no_fuse and os.environ.get("ODINN_NO_FUSE") != "1"
             and (args.dtype == "f64" or os.environ.get("ODINN_MARCH", "2") == "2"))
    vjp_bytes = 4 * w * cells_per_step
    rhs_bytes = 3 * w * cells_per_step
    sub = lambda nbytes, ms_k, words: {"achieved": nbytes / (ms_k * 1e-3) / 1e9, "frac": nbytes / (ms_k * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_cell": words * w, "ms_per_launch": ms_k}
    if fused:
        dom_name = ("sia2d_vjp_march2<WRITE_F>" if args.dtype == "f32" else "sia2d_vjp_march<double, WRITE_F>") + " (F1 + A1 + A2 fused: one launch per step)"
        dom_key, dom_words, dom_ms = "sia2d_fused", 5, ms / args.steps
    else:
        dom_name = "sia2d_vjp_march2 (A1+A2 fused)" if args.dtype == "f32" else "sia2d_vjp_march (A1+A2 fused)"
        dom_key, dom_words, dom_ms = "sia2d_vjp_march2" if args.dtype == "f32" else "sia2d_vjp_march", 4, ms_vjp
    dom = sub(dom_words * w * cells_per_step, dom_ms, dom_words)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * G * esz), "d2h_bytes_per_step": int(2 * G * esz + 8 * G),
                    "api": "odinn_fwd_adj_batch_host (pinned host H, lambda -> dH, vjp_H, S)", "steps": e2e_steps},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": measured_traffic(args.dtype, dom_key, cells_per_step),
                         "traffic_note": "dram read+write bytes per launch from the committed ncu capture (profiles/r01_traffic.json)",
                         "peak_source": peak_src, "algorithmic_bytes_per_cell": dom_words * w, "ms_per_launch": dom_ms,
                         "vjp_kernel": sub(vjp_bytes, ms_vjp, 4), "rhs_kernel": sub(rhs_bytes, ms_rhs, 3),
                         "step_launches": "1 fused kernel (+ the per-glacier S reduction)" if fused else "F1 kernel + A1+A2 kernel (+ the per-glacier S reduction)"},
        }
        if world == 1 and not args.no_cpu:
            rate, el, reps, cores = cpu_fwd_adj(n, n, args.dtype, args.ref_glaciers, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_glaciers} glaciers of {n}x{n} x {reps} passes ({el:.1f} s), C oracle + OpenMP, {args.dtype}"}
        print(json.dumps(line), flush=True)
    ens.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
