"""Ensemble sharding across the GPUs of one box and the loss/gradient all-reduce.

The reference maps one glacier per ``pmap`` task over local worker processes and reduces with
``sum(losses)`` / ``aggregate∇θ`` (src/inverse/SIA2D/gradient.jl:9-30, src/models/trainable_components/
Model.jl:208-224, worker setup src/setup/config.jl:97-139).  Here: one process per GPU (torchrun), glaciers
packed onto ranks by cost, θ replicated, and ONE all-reduce of ``[loss; dθ]`` (|θ|+1 doubles) per optimiser
iteration -- NCCL over NVLink on GPUs; ``gloo`` exercises the same host logic on CPU in the tests.
The message is latency-bound (a few hundred doubles), so it is not fused with compute.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np


def shard_glaciers(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Static greedy bin packing (largest first) of glacier indices onto ranks by cost (nx·ny·n_steps).
    Deterministic; every rank's list is sorted so that per-rank results keep the reference's glacier order."""
    order = sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k))
    loads = [0.0] * world_size
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        bins[r].append(k)
        loads[r] += float(costs[k])
    return [sorted(b) for b in bins]


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process == 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment when WORLD_SIZE > 1.  Returns the module or None."""
    rank, local_rank, ws = world()
    if ws <= 1:
        return None
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return dist


def allreduce_loss_grad(loss: float, dθ: np.ndarray):
    """sum over ranks of [loss; dθ] in one collective (the reference's ``sum(losses)`` + ``aggregate∇θ``).
    Identity when no process group is initialised."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return loss, dθ
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return loss, dθ
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    host = np.empty(1 + dθ.size, dtype=np.float64)
    host[0] = float(loss)
    host[1:] = np.asarray(dθ, dtype=np.float64).ravel()
    buf = torch.from_numpy(host).to(dev)   # one copy in, one collective, one copy out
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    return float(out[0]), out[1:].copy()


def scatter_per_glacier(values_local: np.ndarray, my_ids: Sequence[int], n_total: int) -> np.ndarray:
    """Per-glacier parameter blocks (classical per-glacier A, PerGlacierModel in Model.jl:214-216) are scattered by
    glacier id, not summed: place this rank's values at their global ids, zeros elsewhere, then sum-reduce."""
    full = np.zeros(n_total, dtype=np.float64)
    full[np.asarray(list(my_ids), dtype=np.int64)] = np.asarray(values_local, dtype=np.float64)
    _, full = allreduce_loss_grad(0.0, full)
    return full
