"""Multi-GPU host logic: environments shard across ranks with no data-path collective; NCCL is used
only for PPO gradient averaging and rollout statistics.  Mirrors what the reference does with forked workers and a
multiprocessing.Queue (uhc/agents/agent_handmimic.py:503-535, uhc/khrylib/rl/core/logger_rl.py:51-67)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global env ids owned by `rank`: env i -> rank i mod world."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def mixed_model_ids(n_total: int, n_models: int, rank: int, world: int) -> np.ndarray:
    """Model id per local env for mixed-object batches: global env id mod n_models, so every rank sees the same mix."""
    return (shard(n_total, rank, world) % n_models).astype(np.int32)


def _on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class RolloutStats:
    """Device-side sums of LoggerRL (logger_rl.py:28-67): steps, episodes, reward sum / min / max, 9-vector info sum."""

    def __init__(self, device):
        self.device = device
        self.sums = torch.zeros(2 + 1 + 9, dtype=torch.float64, device=device)   # steps, episodes, reward, info[9]
        self.minmax = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device=device)

    def update(self, reward, info, done):
        r = reward.to(torch.float64)
        self.sums[0] += r.numel()
        self.sums[1] += done.to(torch.float64).sum()
        self.sums[2] += r.sum()
        self.sums[3:] += info.to(torch.float64).sum(0)
        self.minmax[0] = torch.minimum(self.minmax[0], r.min())
        self.minmax[1] = torch.maximum(self.minmax[1], r.max())

    def allreduce(self) -> dict:
        s, mn, mx = self.sums.clone(), self.minmax[0:1].clone(), self.minmax[1:2].clone()
        if _on():
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        s = s.cpu().numpy()
        return dict(num_steps=int(s[0]), num_episodes=int(s[1]), reward_sum=float(s[2]), info_sum=s[3:],
                    reward_min=float(mn.item()), reward_max=float(mx.item()),
                    avg_reward=float(s[2] / max(s[0], 1)))


def global_mean_std(x: torch.Tensor):
    """Mean and unbiased std over all ranks from (sum, sum of squares, n)."""
    x = x.to(torch.float64)
    m = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if _on():
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
    n = m[2]
    mean = m[0] / n
    var = (m[1] - n * mean * mean) / (n - 1)
    return float(mean), float(torch.sqrt(torch.clamp(var, min=0)))


def allreduce_grads(params):
    """Average gradients across ranks (one flat buffer, one collective)."""
    params = [p for p in params if p.grad is not None]
    if not _on() or not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    o = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
