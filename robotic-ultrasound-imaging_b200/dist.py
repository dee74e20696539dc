"""Multi-GPU plumbing: env sharding and rollout-statistic reduction.

Envs never interact (the reference runs them as isolated processes, rl.py:130), so the env step itself has
NO collective: rank r of G owns the contiguous global env ids ``[r*N/G, (r+1)*N/G)`` and per-env Philox
streams are keyed by the GLOBAL id, which makes results independent of G.  Collectives (NCCL on GPUs, gloo
in the CPU tests) are used only for rollout statistics (running normaliser moments, episode statistics) and,
in the PPO driver, for the gradient all-reduce.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_id_offset, num_envs) of ``rank``; the first ``total % world`` ranks take one extra env."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total_envs, world)
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def merge_moments(count_a, mean_a, m2_a, count_b, mean_b, m2_b):
    """Chan et al. parallel merge of (count, mean, M2)."""
    n = count_a + count_b
    delta = mean_b - mean_a
    safe = torch.clamp(n, min=1)
    mean = mean_a + delta * (count_b / safe)
    m2 = m2_a + m2_b + delta * delta * (count_a * count_b / safe)
    return n, mean, m2


def allreduce_moments(count: torch.Tensor, mean: torch.Tensor, m2: torch.Tensor):
    """Exact global (count, mean, M2) from per-rank moments with ONE all_reduce(SUM).

    Uses sums of (n, n*mean, M2 + n*mean^2) in float64, which is algebraically the Chan merge over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return count, mean, m2
    c = count.to(torch.float64).reshape(1)
    mu = mean.to(torch.float64)
    buf = torch.cat([c, (c * mu).reshape(-1), (m2.to(torch.float64) + c * mu * mu).reshape(-1)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    k = mu.numel()
    n = buf[0]
    gmean = buf[1 : 1 + k] / torch.clamp(n, min=1)
    gm2 = buf[1 + k :] - n * gmean * gmean
    return n.to(count.dtype).reshape(count.shape), gmean.to(mean.dtype).reshape(mean.shape), gm2.to(m2.dtype).reshape(m2.shape)


def allreduce_episode_stats(ret_sum: torch.Tensor, len_sum: torch.Tensor, n_ep: torch.Tensor):
    """Sum of episode returns / lengths / counts over ranks (one flat all_reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ret_sum, len_sum, n_ep
    buf = torch.stack([ret_sum.to(torch.float64), len_sum.to(torch.float64), n_ep.to(torch.float64)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[0], buf[1], buf[2]
