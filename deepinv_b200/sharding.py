"""Batch sharding over the GPUs of one box.

Every sample is independent in every operator, in CG (per-sample dots), in the denoisers and in DDRM, so the
path shards over dim 0 with NO collective inside the iteration loop: one process per GPU (torchrun), operator
parameters replicated, per-sample parameters (masks, coil maps, filters with a batch dimension) sharded with the
batch.  The single collective is the all-gather of the final reconstructions.

The reference's own `deepinv.distributed` is operator-/tile-parallel and orthogonal to this (not mirrored).
Two reference semantics that involve the WHOLE batch are preserved explicitly: the `early_stop` criterion is a batch
mean (optimizers.py:722-724) and CG stops when ALL samples converged (conjugate_gradient.py:61) — with sharding, run
fixed iteration counts (the benchmark does) or all-reduce the one scalar (`allreduce_mean` / `allreduce_all`).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """contiguous split of range(n) into world_size shards, remainder spread over the first ranks"""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int | None = None, world_size: int | None = None) -> torch.Tensor:
    """this rank's slice of a batched tensor; tensors with batch dimension 1 (shared parameters) are replicated"""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if t.shape[0] == 1:
        return t
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def gather_batch(x_local: torch.Tensor, n_total: int | None = None) -> torch.Tensor:
    """all-gather the per-rank reconstructions into the full batch (the path's only collective)"""
    rank, w = world()
    if w == 1:
        return x_local
    x_local = x_local.contiguous()
    sizes = [shard_bounds(n_total, r, w)[1] - shard_bounds(n_total, r, w)[0] for r in range(w)] if n_total is not None else None
    if sizes is None or len(set(sizes)) == 1:
        out = torch.empty((w * x_local.shape[0],) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local)
        return out
    # uneven shards: pad every shard to the largest one (collectives need equal message sizes), gather, strip
    mx = max(sizes)
    padded = torch.zeros((mx,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    padded[: x_local.shape[0]] = x_local
    out = torch.empty((w * mx,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(w)], 0)


def allreduce_mean(v: torch.Tensor, n_local: int, n_total: int) -> torch.Tensor:
    """batch mean of a per-rank mean (the early-stop criterion)"""
    _, w = world()
    if w == 1:
        return v
    s = v * n_local
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s / n_total


def allreduce_all(flag: torch.Tensor) -> torch.Tensor:
    """logical AND over ranks of an int/bool flag (CG's all-samples-converged test)"""
    _, w = world()
    if w == 1:
        return flag
    f = flag.to(torch.int32)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return f


def reconstruct_sharded(algo, y: torch.Tensor, make_physics, n_total: int | None = None, **kwargs) -> torch.Tensor:
    """run `algo(y_shard, physics_shard)` on this rank's slice of the measurements and gather the result;
    `make_physics(lo, hi)` builds the operator for samples [lo, hi) (sharding per-sample parameters)"""
    rank, w = world()
    n_total = y.shape[0] if n_total is None else n_total
    lo, hi = shard_bounds(n_total, rank, w)
    x = algo(y[lo:hi], make_physics(lo, hi), **kwargs)
    return gather_batch(x, n_total)
