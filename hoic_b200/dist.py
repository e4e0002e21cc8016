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
    """Device-side LoggerRL (uhc/khrylib/rl/core/logger_rl.py:28-67) for a batch of envs stepping together: step / episode
    counts, c_reward sum / min / max and the 9-vector info sum per step, and per EPISODE the reward sum with its min / max
    (`episode_reward`, kept per env across rollouts so that an episode spanning two iterations is counted whole).
    `finished` collects (motion_id, start_idx, percent) of the episodes that ended, the reference's `freq_dict` rows
    (agent_handmimic.py:487-489)."""

    def __init__(self, device, episode_reward=None):
        self.device = device
        self.sums = torch.zeros(2 + 1 + 9 + 1, dtype=torch.float64, device=device)   # steps, episodes, reward, info[9], episode reward
        self.minmax = torch.tensor([float("inf"), float("-inf"), float("inf"), float("-inf")], dtype=torch.float64, device=device)
        self.episode_reward = episode_reward      # [N] running sums (float64), owned by the caller
        self.finished = []

    def reset(self):
        """Zero in place (the tensors may be baked into a captured CUDA graph)."""
        self.sums.zero_()
        self.minmax.copy_(torch.tensor([float("inf"), float("-inf"), float("inf"), float("-inf")], dtype=torch.float64))
        self.finished = []

    def update(self, reward, info, done, motion_id=None, start_idx=None, percent=None):
        if reward.is_cuda and reward.dtype == torch.float32 and done.dtype == torch.uint8 and info.dtype == torch.float32 \
                and reward.is_contiguous() and info.is_contiguous():
            # one library kernel (hoic_logger_step) instead of ~15 small torch reductions per sampler step
            import ctypes as C
            from . import capi
            L = capi.lib()
            capi.check(L.hoic_logger_step(capi.ptr(reward), capi.ptr(info), capi.ptr(done), reward.numel(), capi.ptr(self.sums),
                                          capi.ptr(self.minmax), capi.ptr(self.episode_reward),
                                          C.c_void_p(torch.cuda.current_stream(reward.device).cuda_stream)), "hoic_logger_step")
            if motion_id is not None:
                self.finished.append((done.to(torch.bool), motion_id.clone(), start_idx.clone(), percent.clone()))
            return
        r = reward.to(torch.float64)
        d = done.to(torch.bool)
        self.sums[0] += r.numel()
        self.sums[1] += d.to(torch.float64).sum()
        self.sums[2] += r.sum()
        self.sums[3:12] += info.to(torch.float64).sum(0)
        self.minmax[0] = torch.minimum(self.minmax[0], r.min())
        self.minmax[1] = torch.maximum(self.minmax[1], r.max())
        if self.episode_reward is not None:
            self.episode_reward += r
            ep = self.episode_reward
            self.sums[12] += torch.where(d, ep, torch.zeros_like(ep)).sum()
            self.minmax[2] = torch.minimum(self.minmax[2], torch.where(d, ep, torch.full_like(ep, float("inf"))).min())
            self.minmax[3] = torch.maximum(self.minmax[3], torch.where(d, ep, torch.full_like(ep, float("-inf"))).max())
            self.episode_reward.masked_fill_(d, 0.0)
        if motion_id is not None:
            self.finished.append((d.clone(), motion_id.clone(), start_idx.clone(), percent.clone()))

    def freq_rows(self, extra=()):
        """[(motion_id, percent, start_idx)] of the finished episodes of this rank, on the host.  `extra`: more
        (done, motion_id, start_idx, percent) tuples, e.g. whole [T, N] rollout arrays."""
        rows = []
        for d, mid, st, pc in list(self.finished) + list(extra):
            d, mid, st, pc = d.reshape(-1).to(torch.bool), mid.reshape(-1), st.reshape(-1), pc.reshape(-1)
            idx = d.nonzero().squeeze(1)
            if idx.numel():
                rows.append(torch.stack([mid[idx].double(), pc[idx].double(), st[idx].double()], 1).cpu().numpy())
        return np.concatenate(rows, 0) if rows else np.zeros((0, 3))

    def allreduce(self) -> dict:
        s, mn, mx = self.sums.clone(), self.minmax[0::2].clone(), self.minmax[1::2].clone()
        if _on():
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        s, mn, mx = s.cpu().numpy(), mn.cpu().numpy(), mx.cpu().numpy()
        steps, eps = max(s[0], 1), max(s[1], 1)
        return dict(num_steps=int(s[0]), num_episodes=int(s[1]), reward_sum=float(s[2]), info_sum=s[3:12],
                    reward_min=float(mn[0]), reward_max=float(mx[0]), avg_reward=float(s[2] / steps),
                    # LoggerRL names
                    avg_c_reward=float(s[2] / steps), avg_c_info=s[3:12] / steps, avg_episode_len=float(s[0] / eps),
                    total_reward=float(s[12]), avg_episode_reward=float(s[12] / eps), avg_episode_c_reward=float(s[2] / eps),
                    min_episode_reward=float(mn[1]), max_episode_reward=float(mx[1]))


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


_flat_cache = {}


def allreduce_grads(params):
    """Average gradients across ranks with ONE collective per network: the gradients of a network live as views into one
    flat buffer (set up on first use, `p.grad` is re-pointed into it), so there is no gather / scatter copy around the
    all-reduce."""
    params = [p for p in params if p.requires_grad]
    if not _on() or not params:
        return
    key = tuple(id(p) for p in params)
    ent = _flat_cache.get(key)
    if ent is None or any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, ent[1])):
        flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
        views, o = [], 0
        for p in params:
            v = flat[o:o + p.numel()].view_as(p)
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
            views.append(v)
            o += p.numel()
        ent = (flat, views)
        _flat_cache[key] = ent
    flat = ent[0]
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
