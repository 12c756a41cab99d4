"""Multi-GPU sharding of a batched evaluate (SURVEY §8 e): the network replicated, the points partitioned into
contiguous blocks, no exchange step during evaluation.  Two drivers of the same partition:

  * ONE process, G GPUs — `evaluate_multi` / `evaluate(f, pts, ngpus=G)`: a thin call into the library's
    multi-device plan (ttn_plan_create_multi: one host thread and three streams per GPU, every GPU copies its block in
    and its values out of the caller's arrays, per-GPU sums added in device order).  This is what a Julia caller gets.
  * one process PER GPU (torch.distributed, the launch contract of bench.py) — `evaluate_sharded`: every rank evaluates
    its block through a single-device plan; collectives only AFTER the kernels:
  * reduce="sum"  -> all_reduce(SUM) of one (real) or two (complex) doubles (grid quadrature),
  * gather=True   -> all_gather of the per-rank value blocks (otherwise every rank keeps its block).
On GPUs the process group is NCCL (NVLink/NVSwitch); the CPU tests drive the same code with gloo
and an injected evaluator.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block of rank `rank`: [rank*ceil(n/world), min(n, (rank+1)*ceil(n/world)))."""
    per = -(-int(n) // int(world))
    lo = min(int(n), rank * per)
    return lo, min(int(n), lo + per)


def _default_evaluator(fitn, dims, device):
    plan = fitn.plan(dims, device=device)

    def run(coords, want_sum):
        out, o = plan.evaluate_host(coords, reduce_sum=bool(want_sum), want_values=not want_sum)
        return out, complex(o.sum_out[0], o.sum_out[1])

    return run, plan.packed.is_complex


def evaluate_multi(fitn, points, dims=None, *, ngpus=None, reduce=None, **kw):
    """One process driving `ngpus` GPUs (default: all of them) through a multi-device plan."""
    from . import _capi
    from .itensornetworkfunction import evaluate
    if ngpus is None:
        ngpus = _capi.lib().ttn_device_count()
    return evaluate(fitn, points, dims, ngpus=int(ngpus), reduce=reduce, **kw)


def evaluate_sharded(fitn, points, dims=None, *, reduce=None, gather=True, group=None, device=None,
                     evaluator=None):
    """Evaluate `points` (the SAME full array on every rank) with the ranks of `group` each taking
    one contiguous block.  Returns the full value vector (gather=True), this rank's block
    (gather=False) or the global sum (reduce="sum")."""
    import torch
    import torch.distributed as dist
    from .itensornetworkfunction import _points_to_coords

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    coords, dims, _ = _points_to_coords(fitn, points, dims)
    n = coords.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    if evaluator is None:
        if device is None:
            device = torch.cuda.current_device()
        run, is_complex = _default_evaluator(fitn, dims, device)
    else:
        run, is_complex = evaluator
    backend = dist.get_backend(group) if dist.is_initialized() else None
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    vals, s = run(coords[lo:hi], reduce == "sum")
    if reduce == "sum":
        t = torch.tensor([s.real, s.imag], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
        return complex(t[0].item(), t[1].item()) if is_complex else t[0].item()
    if not gather or world == 1:
        return vals
    per = -(-n // world)
    dt = np.complex128 if is_complex else np.float64
    block = np.zeros(per, dtype=dt)
    block[: hi - lo] = vals
    tb = torch.from_numpy(block.view(np.float64)).to(dev)
    full = torch.empty(world * tb.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full, tb, group=group)
    return full.cpu().numpy().view(dt)[:n].copy() if per * world != n else full.cpu().numpy().view(dt)
