"""Multi-GPU plumbing: one process per GPU, the independent-filter batch is cut into contiguous shards, the time axis
is never split and there is NO collective on the data path.  ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests) is used only after the scan: to gather ``nell`` / per-filter summaries and for the argmin over a
theta grid.

The reference's equivalent is process fan-out in shell scripts (``dardel/run_benes_bernoulli_mf.sh:26-45``, Slurm
arrays in ``dardel/run_prey_predator_mf.sh:6``) followed by post-processing of ``.npz`` files
(``dardel/benes_bernoulli/post_processing_mf.py:41-66``).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'shard', 'gather_filters', 'argmin_over_shards', 'local_argmin', 'running_argmin', 'world',
           'max_over_ranks']


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``total`` filters: the first ``total % world_size`` ranks get one more."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(x, rank: Optional[int] = None, world_size: Optional[int] = None, axis: int = 0):
    """This rank's contiguous slice of ``x`` along the batch ``axis`` (NumPy array or torch tensor)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[axis], rank, world_size)
    idx = [slice(None)] * x.ndim
    idx[axis] = slice(lo, hi)
    return x[tuple(idx)]


def gather_filters(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather per-filter results (leading axis = this rank's shard, made by ``shard_bounds``) into the full
    ``(total, ...)`` tensor on every rank.  Shards may differ by one row, so they are padded to the largest."""
    rank, w = world()
    if w == 1:
        return local
    sizes = [shard_bounds(total, r, w)[1] - shard_bounds(total, r, w)[0] for r in range(w)]
    mx = max(sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def argmin_over_shards(local_min: torch.Tensor, local_arg: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global (min, argmin) per trajectory when the theta grid is sharded over ranks.

    ``local_min[j]`` is this rank's smallest nell for trajectory ``j`` and ``local_arg[j]`` its GLOBAL theta index.
    NCCL has no MINLOC, so the (value, index) pairs are all-gathered (16 B per trajectory per rank) and reduced
    locally; NaN objectives (diverged filters) never win; ties go to the smallest index.
    """
    rank, w = world()
    if w == 1:
        return local_min, local_arg
    mins = [torch.empty_like(local_min) for _ in range(w)]
    args = [torch.empty_like(local_arg) for _ in range(w)]
    dist.all_gather(mins, local_min.contiguous())
    dist.all_gather(args, local_arg.contiguous())
    mins, args = torch.stack(mins), torch.stack(args)          # (world, n_traj)
    inf = torch.full_like(mins, float('inf'))
    clean = torch.where(torch.isnan(mins), inf, mins)
    best = clean.min(dim=0).values
    is_best = clean == best.unsqueeze(0)
    big = torch.full_like(args, torch.iinfo(args.dtype).max)
    best_arg = torch.where(is_best, args, big).min(dim=0).values
    return best, best_arg


def local_argmin(nell: torch.Tensor, theta_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(min, global argmin) over axis 0 of a ``(n_theta_local, n_traj)`` nell table, ignoring NaN."""
    clean = torch.where(torch.isnan(nell), torch.full_like(nell, float('inf')), nell)
    val, idx = clean.min(dim=0)
    return val, idx.to(torch.int64) + int(theta_offset)


def running_argmin(best: Optional[Tuple[torch.Tensor, torch.Tensor]], nell: torch.Tensor, theta_offset: int = 0):
    """Fold one chunk of a theta grid, a ``(n_theta_chunk, n_traj)`` nell table whose first row is global theta index
    ``theta_offset``, into the running per-trajectory ``(min, argmin)``; ``best=None`` starts.  Chunks must arrive in
    ascending theta order for ties to go to the smallest index (strict ``<``).  This is how a rank walks its slice of a
    grid that does not fit one launch (BASELINE configs[3]: 512 x 512 theta x 10^3 records)."""
    v, a = local_argmin(nell, theta_offset)
    if best is None:
        return v, a
    upd = v < best[0]
    return torch.where(upd, v, best[0]), torch.where(upd, a, best[1])


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (used for device-timed durations)."""
    rank, w = world()
    if w == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
