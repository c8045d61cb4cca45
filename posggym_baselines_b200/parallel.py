"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards without data-path collectives: particles / rollouts / envs are split
into contiguous index ranges and every random draw is a Philox counter of the GLOBAL index,
so results do not depend on the number of GPUs.  The one exchange
step is root-parallel POMCP: every rank grows its own tree from the same root belief and
the root statistics -- ``ActionNode.visits`` and ``ActionNode.total_value``
(/root/reference/posggym_baselines/planning/node.py:138-140,173-174) -- are summed with ONE
allreduce per decision (|A| x 2 doubles = 80 B for 5 actions: latency-bound, NCCL over
NVLink; on CPU test runs the same code path uses gloo).
"""
from __future__ import annotations

import random
from typing import List, Sequence, Tuple

import numpy as np
import torch


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of ``n_total`` units owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_root_statistics(visits: np.ndarray, total_value: np.ndarray, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """Sum of the per-rank root statistics: one allreduce of ``2 * |A|`` float64."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return np.asarray(visits, np.float64), np.asarray(total_value, np.float64)
    buf = torch.tensor(np.concatenate([visits, total_value]), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    a = len(visits)
    return out[:a], out[a:]


def merge_root_statistics_device(buf: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """``buf``: device tensor of ``2 * |A|`` float64 (visits then value sums, filled by
    ``posg_tree_root_stats_device``).  ONE allreduce on the device buffer, ONE read-back."""
    import torch.distributed as dist

    if world()[1] > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    a = out.shape[0] // 2
    return out[:a].copy(), out[a:].copy()


def select_root_action(visits: np.ndarray, total_value: np.ndarray, actions: Sequence[int], mode: str,
                       tie_seed: int) -> int:
    """Final action from (merged) root statistics: max value for ucb / uniform, max visits for
    pucb (mcts.py:66-74,566-600); ties broken by a generator seeded identically on every rank."""
    rng = random.Random(tie_seed)
    visits, total_value = np.asarray(visits, np.float64), np.asarray(total_value, np.float64)
    if mode == "pucb":
        if visits.sum() == 0:
            return rng.choice(list(actions))
        best = visits.max()
        return rng.choice([a for a in actions if visits[a] == best])
    values = np.where(visits > 0, total_value / np.maximum(visits, 1), 0.0)
    best = values.max()
    return rng.choice([a for a in actions if values[a] == best])
