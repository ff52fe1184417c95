"""Sharding of a batch of independent QPs across the GPUs of one box (BASELINE config C5).

A single solve does not shard; a batch of independent instances does: rank r of R takes the contiguous
slice [r*ceil(n/R), (r+1)*ceil(n/R)) and solves it with one batched MPCSolver on its own GPU and stream.
There is no collective on the data path; the only exchange is one all_gather of the per-instance result
record (status, iterations, objective, inf_pr, inf_du, seconds = 48 bytes), over NCCL on GPUs or gloo on CPU.
"""
from __future__ import annotations

import numpy as np

RECORD_FIELDS = ("status", "iter", "objective", "inf_pr", "inf_du", "seconds")
STATUS_CODES = {"SOLVE_SUCCEEDED": 1, "MAXIMUM_ITERATIONS_EXCEEDED": 2, "ERROR_IN_STEP_COMPUTATION": 3,
                "INTERNAL_ERROR": 4, "INITIAL": 0}
STATUS_NAMES = {v: k for k, v in STATUS_CODES.items()}


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice of rank `rank`; every rank gets ceil(n/R) items except possibly the last ones."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def make_records(status, iters, objective, inf_pr, inf_du, seconds):
    """[n_local, 6] float64 record array for the gather."""
    n = len(status)
    rec = np.zeros((n, len(RECORD_FIELDS)))
    rec[:, 0] = [STATUS_CODES.get(s, 4) for s in status]
    rec[:, 1] = iters
    rec[:, 2] = objective
    rec[:, 3] = inf_pr
    rec[:, 4] = inf_du
    rec[:, 5] = seconds
    return rec


def gather_records(local: np.ndarray, n_items: int, device=None):
    """All-gather the per-instance records of every rank; returns the [n_items, 6] table in instance order.
    Uses the default torch.distributed process group (nccl on GPUs, gloo on CPU); with no process group it
    is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[:n_items].copy()
    world = dist.get_world_size()
    per = (n_items + world - 1) // world
    buf = torch.zeros((per, len(RECORD_FIELDS)), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.as_tensor(local, dtype=torch.float64, device=device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    table = torch.cat(out).cpu().numpy()
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        rows.append(table[r * per:r * per + (hi - lo)])
    return np.concatenate(rows) if rows else table[:0]
