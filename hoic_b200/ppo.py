"""Batched PPO pieces around the rollout engine, in plain PyTorch (policy / value MLPs and the update only):
`MLP` / `PolicyGaussian` / `Value` (uhc/khrylib/models/mlp.py:5-27, rl/core/policy_gaussian.py:9-47, critic.py:5-18,
state_dict keys kept: net.affine_layers.N.*, action_mean.*, action_log_std, value_head.*), a batched `ZFilter`
(utils/zfilter.py:7-73), GAE (rl/core/common.py:5-25) as a device-side reverse scan over [T, N], and the clipped
surrogate update (rl/agents/agent_ppo.py:16-64, agent_pg.py:18-56).
"""
from __future__ import annotations

import math
import pickle

import torch
import torch.nn as nn

from . import dist as hdist


class MLP(nn.Module):
    """uhc/khrylib/models/mlp.py:5-27.  The first layer's input width (617 observation entries) is not a multiple of 8, which
    sends its GEMM to a legacy tensor-op kernel five times slower than the native one (profiles/README.md, r02): input and
    weight are zero-padded to a multiple of 64 columns on the fly (exact: the padded products are zeros; parameters keep their
    checkpoint shapes)."""

    def __init__(self, input_dim, hidden_dims=(2048, 1024, 512), activation="gelu"):
        super().__init__()
        self.activation = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid, "gelu": nn.GELU()}[activation]
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh
        self.in_pad = (-input_dim) % 64 if input_dim % 8 else 0

    def forward(self, x):
        for i, affine in enumerate(self.affine_layers):
            if i == 0 and self.in_pad and x.is_cuda:
                x = nn.functional.linear(nn.functional.pad(x, (0, self.in_pad)), nn.functional.pad(affine.weight, (0, self.in_pad)), affine.bias)
            else:
                x = affine(x)
            x = self.activation(x)
        return x


class PolicyGaussian(nn.Module):
    def __init__(self, net, action_dim, log_std=0.0, fix_std=False):
        super().__init__()
        self.type = "gaussian"
        self.net = net
        self.action_mean = nn.Linear(net.out_dim, action_dim)
        self.action_mean.weight.data.mul_(0.1)
        self.action_mean.bias.data.mul_(0.0)
        self.action_log_std = nn.Parameter(torch.ones(1, action_dim) * log_std, requires_grad=not fix_std)

    def forward(self, x):
        mean = self.action_mean(self.net(x))
        log_std = self.action_log_std.expand_as(mean)
        return mean, log_std

    @torch.no_grad()
    def select_action(self, x, mean_action=False, generator=None):
        mean, log_std = self.forward(x)
        if mean_action:
            return mean
        return mean + torch.exp(log_std) * torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)

    def get_log_prob(self, x, action):
        mean, log_std = self.forward(x)
        var = torch.exp(2 * log_std)
        lp = -(action - mean) ** 2 / (2 * var) - 0.5 * math.log(2 * math.pi) - log_std
        return lp.sum(1, keepdim=True)


class Value(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net
        self.value_head = nn.Linear(net.out_dim, 1)
        self.value_head.weight.data.mul_(0.1)
        self.value_head.bias.data.mul_(0.0)

    def forward(self, x):
        return self.value_head(self.net(x))


class ZFilter:
    """y = clip((x - mean) / (std + 1e-8), -clip, clip) with running statistics (uhc/khrylib/utils/zfilter.py:7-73), batched:
    a [N, dim] batch is merged with Chan's parallel Welford update.  The accumulators are float64 like the reference's
    numpy RunningStat, whatever dtype the observations have.  The statistics see every env (the reference keeps worker
    0's, SURVEY.md section 5 -- documented difference).

    Multi-rank: each rank also accumulates what it saw SINCE the last `allreduce()` in a separate (n, mean, M2) triple;
    `allreduce()` pools those per-rank deltas and merges them into the pooled state of the previous sync.  (Pooling the
    ranks' full statistics again and again would count the shared history `world` times per iteration.)"""

    def __init__(self, shape, clip=5.0, device="cpu", dtype=torch.float64):
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=device)
        self.n, self.mean, self.m2 = z(()), z(shape), z(shape)
        self._dn, self._dmean, self._dm2 = z(()), z(shape), z(shape)     # since the last allreduce
        self._bn, self._bmean, self._bm2 = z(()), z(shape), z(shape)     # pooled state at the last allreduce
        self.clip = clip
        self.out_dtype = dtype

    @property
    def std(self):
        var = torch.where(self.n > 1, self.m2 / torch.clamp(self.n - 1, min=1), self.mean ** 2)
        return torch.sqrt(torch.clamp(var, min=0))

    @staticmethod
    def _merge(n, mean, m2, nb, bmean, bm2):
        tot = n + nb
        delta = bmean - mean
        w = nb / torch.clamp(tot, min=1)
        return tot, mean + delta * w, m2 + bm2 + delta ** 2 * (n * w)

    def _assign(self, names, values):
        # in place: the tensors keep their addresses, so an update may be captured in a CUDA graph and replayed
        for name, v in zip(names, values):
            getattr(self, name).copy_(v)

    def update(self, x):
        x = x.to(torch.float64)
        nb = float(x.shape[0])
        bmean = x.mean(0)
        bm2 = ((x - bmean) ** 2).sum(0)
        self._assign(("n", "mean", "m2"), self._merge(self.n, self.mean, self.m2, nb, bmean, bm2))
        self._assign(("_dn", "_dmean", "_dm2"), self._merge(self._dn, self._dmean, self._dm2, nb, bmean, bm2))

    def __call__(self, x, update=True, out=None, out16=None):
        """`out` (float32 [N, dim]) / `out16` (bfloat16 [N, ld >= dim], padding columns zeroed): device buffers the library
        path writes the normalised batch into (the sampler step keeps both at fixed addresses for its CUDA graph)."""
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous():
            return self._call_library(x, update, out, out16)
        if update:
            self.update(x)
        y = (x.to(torch.float64) - self.mean) / (self.std + 1e-8)
        return torch.clamp(y, -self.clip, self.clip).to(x.dtype)

    def _call_library(self, x, update, out=None, out16=None):
        """Device batches go through libhoic_b200.so (hoic_zfilter: one pass for the batch moments, one to normalise) instead
        of a dozen torch kernels; same arithmetic (double accumulation, Chan merge)."""
        from . import capi
        import ctypes as C
        L = capi.lib()
        if getattr(self, "_lib_buf", None) is None or self._lib_buf[0].device != x.device:
            dim = self.mean.numel()
            self._lib_buf = (torch.zeros(64 * dim + 1, dtype=torch.float64, device=x.device),
                             torch.zeros(2 * dim, dtype=torch.float32, device=x.device), {})
        scratch, shift_scale, outs = self._lib_buf
        key = (x.shape[0], torch.cuda.is_current_stream_capturing())
        y = out if out is not None else outs.get(key)
        if y is None:
            y = outs[key] = torch.empty_like(x)
        if out16 is not None and not (out16.dtype == torch.bfloat16 and out16.is_contiguous() and out16.shape[0] == x.shape[0]
                                      and out16.shape[1] >= x.shape[1]):
            raise ValueError("ZFilter: out16 must be a contiguous bfloat16 [N, ld >= dim] tensor")
        capi.check(L.hoic_zfilter16(capi.ptr(x), x.shape[0], x.shape[1], capi.ptr(self.n), capi.ptr(self.mean), capi.ptr(self.m2),
                                    capi.ptr(self._dn), capi.ptr(self._dmean), capi.ptr(self._dm2), capi.ptr(scratch),
                                    capi.ptr(shift_scale), float(self.clip), 1 if update else 0, capi.ptr(y),
                                    capi.ptr(out16) if out16 is not None else None, out16.shape[1] if out16 is not None else 0,
                                    C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "hoic_zfilter")
        return y

    def set_state(self, n, mean, m2):
        """Adopt pooled statistics (checkpoint load): they become the synced base, nothing is pending."""
        dev = self.mean.device
        self._assign(("n", "mean", "m2"), (torch.as_tensor(float(n), dtype=torch.float64, device=dev),
                                            torch.as_tensor(mean, dtype=torch.float64, device=dev),
                                            torch.as_tensor(m2, dtype=torch.float64, device=dev)))
        self._assign(("_bn", "_bmean", "_bm2"), (self.n, self.mean, self.m2))
        for name in ("_dn", "_dmean", "_dm2"):
            getattr(self, name).zero_()

    @staticmethod
    def merge_moments(parts):
        """Pooled (n, mean, M2) of (n, mean, M2) triples: the arithmetic `allreduce` performs with collectives."""
        n = sum(p[0].double() for p in parts)
        mean = sum(p[1].double() * p[0].double() for p in parts) / torch.clamp(n, min=1)
        m2 = sum(p[2].double() + p[0].double() * (p[1].double() - mean) ** 2 for p in parts)
        return n, mean, torch.clamp(m2, min=0)

    def allreduce(self):
        """Pool what the ranks saw since the last call (counts add, means pool, centred second moments combine as
        M2 = sum_r [M2_r + n_r (mean_r - mean)^2], Chan et al.) and merge it into the previously pooled state."""
        if not hdist._on():
            return
        import torch.distributed as dist
        n_r, mean_r, m2_r = self._dn, self._dmean, self._dm2
        head = torch.cat([n_r.reshape(1), (mean_r * n_r).reshape(-1)])
        dist.all_reduce(head)
        dn = head[0]
        dmean = head[1:].reshape(mean_r.shape) / torch.clamp(dn, min=1)
        dm2 = m2_r + n_r * (mean_r - dmean) ** 2
        dist.all_reduce(dm2)
        self._assign(("n", "mean", "m2"), self._merge(self._bn, self._bmean, self._bm2, dn, dmean, torch.clamp(dm2, min=0)))
        self._assign(("_bn", "_bmean", "_bm2"), (self.n, self.mean, self.m2))
        for name in ("_dn", "_dmean", "_dm2"):
            getattr(self, name).zero_()


def estimate_advantages(rewards, masks, values, gamma, tau):
    """GAE over time-major [T, N] tensors (common.py:5-25 applied per environment column).  `values` has T+1 rows: the
    last row bootstraps envs whose episode is still running when the fixed-horizon rollout stops (the reference only
    stores complete episodes; SURVEY.md R21 documents this).  A terminal transition (mask 0) SELECTS zero for the tail
    terms instead of multiplying by the mask, so a non-finite value behind an episode boundary cannot cross it."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    zero = torch.zeros_like(last)
    for t in range(T - 1, -1, -1):
        live = masks[t] > 0
        delta = rewards[t] + torch.where(live, gamma * values[t + 1], zero) - values[t]
        last = delta + torch.where(live, gamma * tau * last, zero)
        adv[t] = last
    returns = values[:-1] + adv
    mean, std = hdist.global_mean_std(adv.reshape(-1))
    return (adv - mean) / std, returns


def ppo_update(policy, value, opt_policy, opt_value, states, actions, returns, advantages, fixed_log_probs, clip_epsilon=0.2,
               num_epoch=5, grad_clip=40.0, ind=None):
    """agent_ppo.py:16-64 with full-batch steps (use_mini_batch=False, agent_handmimic.py:66); `ind` = indices of the
    exploratory samples (exps.nonzero(), :44) the surrogate is taken over.  Gradients are averaged across ranks before
    the grad-norm clip, so N ranks with equal shards take the single-process full-batch step."""
    vloss = surr = None
    if ind is not None and ind.numel() == states.shape[0]:
        ind = None
    for _ in range(num_epoch):
        vloss = (value(states) - returns).pow(2).mean()
        opt_value.zero_grad(set_to_none=False)
        vloss.backward()
        hdist.allreduce_grads(value.parameters())
        opt_value.step()
        if ind is None:
            log_probs = policy.get_log_prob(states, actions)
            ratio = torch.exp(log_probs - fixed_log_probs)
            adv = advantages
        else:
            log_probs = policy.get_log_prob(states[ind], actions[ind])
            ratio = torch.exp(log_probs - fixed_log_probs[ind])
            adv = advantages[ind]
        surr = -torch.min(ratio * adv, torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon) * adv).mean()
        opt_policy.zero_grad(set_to_none=False)
        surr.backward()
        hdist.allreduce_grads(policy.parameters())
        torch.nn.utils.clip_grad_norm_(policy.parameters(), grad_clip)
        opt_policy.step()
    if vloss is None:
        return {}
    return dict(value_loss=float(vloss.detach()), surrogate=float(surr.detach()))


def make_policy_value(cfg, obs_dim=617, act_dim=32, device="cuda", dtype=torch.float32, seed=1):
    torch.manual_seed(seed)
    policy = PolicyGaussian(MLP(obs_dim, cfg.policy_hsize, cfg.policy_htype), act_dim, log_std=cfg.log_std, fix_std=cfg.fix_std)
    value = Value(MLP(obs_dim, cfg.value_hsize, cfg.value_htype))
    return policy.to(device=device, dtype=dtype), value.to(device=device, dtype=dtype)


# ------------------------------------------------------------------------------------------------ checkpoints
# The reference pickles {"policy_dict", "value_dict", "running_state"} with `running_state` an instance of
# uhc.khrylib.utils.zfilter.ZFilter holding a RunningStat (_n, _M, _S) (agent_handmimic.py:175-186; loaded through the
# CustomUnpickler of uhc/utils/tools.py:7-18).  The two classes below carry the same attribute names and are pickled under
# the reference's module path, so a file written here opens in the reference and a released `iter_XXXX.p` opens here
# without the reference being importable.
class _RefRunningStat:
    def __init__(self, n, M, S):
        self._n, self._M, self._S = n, M, S


class _RefZFilter:
    def __init__(self, rs, clip):
        self.demean, self.destd, self.clip, self.rs = True, True, clip, rs


_RefRunningStat.__module__ = _RefZFilter.__module__ = "uhc.khrylib.utils.zfilter"
_RefRunningStat.__qualname__ = _RefRunningStat.__name__ = "RunningStat"
_RefZFilter.__qualname__ = _RefZFilter.__name__ = "ZFilter"


class _ref_modules:
    """While saving, the reference's module path resolves to the two stand-ins, so that pickle writes plain class
    references (`uhc.khrylib.utils.zfilter.ZFilter`) + attribute dicts: exactly what the reference's own dump contains."""
    NAMES = ("uhc", "uhc.khrylib", "uhc.khrylib.utils", "uhc.khrylib.utils.zfilter")

    def __enter__(self):
        import sys, types
        self.saved = {n: sys.modules.get(n) for n in self.NAMES}
        for n in self.NAMES:
            if n not in sys.modules:
                sys.modules[n] = types.ModuleType(n)
        self.prev = {k: getattr(sys.modules[self.NAMES[-1]], k, None) for k in ("ZFilter", "RunningStat")}
        sys.modules[self.NAMES[-1]].ZFilter = _RefZFilter
        sys.modules[self.NAMES[-1]].RunningStat = _RefRunningStat

    def __exit__(self, *exc):
        import sys
        for k, v in self.prev.items():
            if v is not None:
                setattr(sys.modules[self.NAMES[-1]], k, v)
        for n, m in self.saved.items():
            if m is None:
                sys.modules.pop(n, None)


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == "ZFilter" and "zfilter" in module:
            return _RefZFilter
        if name == "RunningStat" and "zfilter" in module:
            return _RefRunningStat
        return super().find_class(module, name)


def save_checkpoint(path, policy_net, value_net, running_state):
    """`AgentHandMimic.save_checkpoint` (agent_handmimic.py:175-186): state dicts on the CPU + the observation normaliser."""
    import numpy as np
    rs = _RefRunningStat(int(running_state.n.item()), running_state.mean.detach().double().cpu().numpy().copy(),
                         running_state.m2.detach().double().cpu().numpy().copy())
    cp = {"policy_dict": {k: v.detach().cpu() for k, v in policy_net.state_dict().items()},
          "value_dict": {k: v.detach().cpu() for k, v in value_net.state_dict().items()},
          "running_state": _RefZFilter(rs, running_state.clip)}
    with open(path, "wb") as f, _ref_modules():
        pickle.dump(cp, f, protocol=2)


def load_checkpoint(path, policy_net, value_net, running_state):
    """`load_checkpoint` (agent_handmimic.py:135-169): restores the networks (cast to their dtype) and the normaliser."""
    with open(path, "rb") as f:
        cp = _RefUnpickler(f).load()
    policy_net.load_state_dict({k: torch.as_tensor(v).to(next(policy_net.parameters())) for k, v in cp["policy_dict"].items()})
    value_net.load_state_dict({k: torch.as_tensor(v).to(next(value_net.parameters())) for k, v in cp["value_dict"].items()})
    z = cp["running_state"]
    running_state.set_state(z.rs._n, z.rs._M, z.rs._S)
    if z.clip:
        running_state.clip = float(z.clip)
    return cp
