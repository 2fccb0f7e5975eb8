"""Env-range sharding across ranks: one process per GPU, contiguous env ranges, no
data-path collective; one small all-reduce of the final cost statistics.

Host logic only (works with any ``torch.distributed`` backend: NCCL on the box, gloo in the CPU
tests).  Episodes never interact (the reference batches them with ``jax.vmap``,
``deluca/training/apg.py:91``), so nothing inside the T-step loop communicates.
"""
import torch
import torch.distributed as dist


def shard_range(N_total, rank, world):
    """Contiguous range [lo, hi) of the N_total global envs owned by ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(N_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_stats(cost_sum, status):
    """[sum cost, sum cost^2, finite envs, non-finite envs] of this rank's shard (fp64, on device)."""
    ok = status == 0
    cs = torch.where(ok, cost_sum.double(), torch.zeros_like(cost_sum, dtype=torch.float64))
    return torch.stack([cs.sum(), (cs * cs).sum(), ok.sum().double(), (~ok).sum().double()])


def reduce_stats(stats):
    """Sum the per-rank statistics vector over all ranks (a <= 8 KB all-reduce; NCCL over NVLink on
    the box).  No-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats, T):
    s = stats.tolist()
    n = max(s[2], 1.0)
    mean = s[0] / n
    var = max(s[1] / n - mean * mean, 0.0)
    return {"mean_cost_per_env_step": mean / T, "std_episode_cost": var ** 0.5, "finite_envs": s[2],
            "nonfinite_envs": s[3]}
