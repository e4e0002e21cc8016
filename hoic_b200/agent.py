"""One PPO iteration on the batched rollout engine: the host-side mirror of `AgentHandMimic`
(uhc/agents/agent_handmimic.py:26-74, 264-337, 430-535) + `AgentPPO.update_params` (uhc/khrylib/rl/agents/agent_pg.py:40-56,
agent_ppo.py:16-64) with the same method names (`per_epoch_update`, `sample`, `update_params`, `optimize_policy`).

What differs from the reference, by construction of a batched collector:
* N environments step in lockstep for a fixed horizon T = ceil(min_batch_size / (N * world)) instead of forked workers
  running whole episodes until `min_batch_size` samples exist; an episode that is still running at the horizon is
  bootstrapped with V(s_T) in the GAE (the reference only stores complete episodes, so its last mask is always 0);
* the observation normaliser sees every environment (the reference keeps worker 0's statistics);
* float32 networks (the reference runs torch.float64), Adam steps are full-batch over the rank's shard with gradients
  averaged across ranks (NCCL) before the grad-norm clip, which equals the single-process full-batch step.
PyTorch does the MLPs and the update only; every env number comes from libhoic_b200.so through `HandMimicVecEnv`.
"""
from __future__ import annotations

import math
import time

import torch

from . import dist as hdist
from . import ppo
from .sampler import SamplerStep, matmul_context
from .vecenv import HandMimicVecEnv


class TrajBatch:
    """Time-major rollout storage on the device (uhc/khrylib/rl/core/trajbatch.py:5-15 fields, [T, N, ...])."""

    def __init__(self, T, N, obs_dim, act_dim, device):
        f = dict(dtype=torch.float32, device=device)
        self.states = torch.empty(T, N, obs_dim, **f)
        self.actions = torch.empty(T, N, act_dim, **f)
        self.rewards = torch.empty(T, N, **f)
        self.masks = torch.empty(T, N, **f)
        self.exps = torch.empty(T, N, **f)
        self.last_state = torch.empty(N, obs_dim, **f)   # normalised s_T of the envs (bootstrap)
        self.T, self.N = T, N


class AgentHandMimic:
    def __init__(self, models, n_envs, motions=None, motion_model=None, env_model_id=None, device=0, seed=None, cfg=None,
                 policy_grad_clip=40.0, policy_dtype="bf16", update_dtype="tf32", graph=True):
        models = [models] if not isinstance(models, (list, tuple)) else list(models)
        self.cfg = cfg or models[0].cfg
        cfg = self.cfg
        self.device = torch.device("cuda", device)
        self.rank = torch.distributed.get_rank() if hdist._on() else 0
        self.world = torch.distributed.get_world_size() if hdist._on() else 1
        seed = cfg.seed if seed is None else seed
        if motions is None:
            # the sampler never draws a data set's LAST sequence (randint(0, seq_num - 1), agent_handmimic.py:444): it is
            # the held-out sequence eval_policy runs (:344).  Default data set per model: synthetic motions 1..9, of
            # which 1..8 train and 9 evaluates.
            motions, motion_model = [], []
            for i, m in enumerate(models):
                for s in range(1, 9):
                    motions.append(m.motion(s))
                    motion_model.append(i)
            self.eval_motions = [[m.motion(9)] for m in models]
        else:
            self.eval_motions = None
        self.env = HandMimicVecEnv(models, n_envs, motions, motion_model=motion_model, env_model_id=env_model_id, device=device,
                                   seed=seed + 1000 * self.rank, auto_reset=True, random_start=True)
        self.n_envs = n_envs
        # identical initial weights on every rank (same seed), as a forked worker would inherit them
        self.policy_net, self.value_net = ppo.make_policy_value(cfg, self.env.obs_dim, self.env.act_dim, device=self.device,
                                                                dtype=torch.float32, seed=seed)
        self.running_state = ppo.ZFilter((self.env.obs_dim,), clip=5.0, device=self.device)   # float64 accumulators
        self.optimizer_policy = torch.optim.Adam(self.policy_net.parameters(), lr=cfg.policy_lr, weight_decay=cfg.policy_weightdecay)
        self.optimizer_value = torch.optim.Adam(self.value_net.parameters(), lr=cfg.value_lr, weight_decay=cfg.value_weightdecay)
        self.policy_grad_clip = policy_grad_clip
        self.gen = torch.Generator(device=self.device).manual_seed(seed + 7919 * (self.rank + 1))
        self.noise_rate = 1.0
        self.epoch = 0
        self._obs = None
        # env.end_reward of the reference (ho_im4.py:117; set per iteration in optimize_policy, agent_handmimic.py:317-319)
        self.end_reward = 0.0
        self._episode_reward = torch.zeros(n_envs, dtype=torch.float64, device=self.device)
        # freq_dict of the reference (agent_handmimic.py:487-489, 523-533): per training motion the [percent, start] rows of
        # its finished episodes, the newest `max_freq` kept
        self.freq_dict = {k: [] for k in range(len(motions))}
        self.max_freq = 50
        # the sampler's inner loop as one CUDA graph; MLPs on the tensor cores (bf16 operands, fp32 accumulate and master weights)
        # rollout inference: bf16 operands (the action-mean error, < 1e-3 sigma at init, only perturbs the exploration noise);
        # GAE values, log-probabilities and the PPO update: TF32 tensor cores -- with bf16 there the rounding noise of the mean
        # enters the importance ratio exp(logp - logp_old) through (a - mu) dmu / sigma^2 with sigma = 0.1 and training
        # became unstable (profiles/r02_train_box_2048x25_*.jsonl)
        self.policy_dtype, self.update_dtype = policy_dtype, update_dtype
        self.stats = hdist.RolloutStats(self.device, episode_reward=self._episode_reward)
        self.sampler = SamplerStep(self.env, self.policy_net, self.running_state, self.gen, stats=self.stats,
                                   policy_dtype=policy_dtype, graph=graph)

    # ------------------------------------------------------------------ agent_handmimic.py:264-282
    def per_epoch_update(self, epoch):
        cfg = self.cfg
        cfg.update_adaptive_params(epoch)
        self.noise_rate = float(cfg.adp_noise_rate)
        for g in self.optimizer_policy.param_groups:
            g["lr"] = float(cfg.adp_policy_lr)
        if cfg.fix_std:
            with torch.no_grad():
                self.policy_net.action_log_std.fill_(float(cfg.adp_log_std))
        self.env.set_reward_weights(cfg.reward_vector())   # epoch-adaptive reward weights (handmimic_config.py:175-195)

    # ------------------------------------------------------------------ agent_handmimic.py:430-535
    @torch.no_grad()
    def sample(self, min_batch_size):
        env, N = self.env, self.n_envs
        env.set_mode("train")
        T = max(1, math.ceil(min_batch_size / (N * self.world)))
        batch = TrajBatch(T, N, env.obs_dim, env.act_dim, self.device)
        s = self.sampler
        if self._obs is None:
            self._obs = env.reset()
            s.warmup()                     # (three policy steps whose transitions are dropped, then the graph exists)
        s.set_noise_rate(self.noise_rate)
        stats = self.stats
        stats.reset()
        use_end = bool(self.cfg.end_reward) and self.end_reward != 0.0
        ep = dict(done=torch.empty(T, N, dtype=torch.uint8, device=self.device), mid=torch.empty(T, N, dtype=torch.int32, device=self.device),
                  start=torch.empty(T, N, dtype=torch.int32, device=self.device), percent=torch.empty(T, N, device=self.device))
        for t in range(T):
            s.step()
            batch.states[t].copy_(s.x)
            batch.actions[t].copy_(s.act)
            batch.exps[t].copy_(s.exp)
            # `reward += env.end_reward` on info["end"] (agent_handmimic.py:478-480); the logger sees c_reward without it
            batch.rewards[t] = env.reward + self.end_reward * env.end.to(torch.float32) if use_end else env.reward
            batch.masks[t] = 1.0 - env.done.to(torch.float32)
            ep["done"][t].copy_(env.done); ep["mid"][t].copy_(s.motion_id); ep["start"][t].copy_(s.start_idx)
            ep["percent"][t].copy_(env.percent)
        self._obs = env.obs
        batch.last_state = self.running_state(env.obs, update=False).to(torch.float32)
        self.running_state.allreduce()
        for k, pc, st in stats.freq_rows([(ep["done"], ep["mid"], ep["start"], ep["percent"])]):
            self.freq_dict.setdefault(int(k), []).append([float(pc), int(st)])
        self.freq_dict = {k: v[-self.max_freq:] for k, v in self.freq_dict.items()}
        return batch, stats

    # ------------------------------------------------------------------ agent_pg.py:40-56 + agent_ppo.py:16-64
    def update_params(self, batch):
        cfg = self.cfg
        T, N = batch.T, batch.N
        states = batch.states.reshape(T * N, -1)
        actions = batch.actions.reshape(T * N, -1)
        with torch.no_grad(), matmul_context(self.update_dtype):
            values = self.value_net(torch.cat([states, batch.last_state], 0)).float().reshape(T + 1, N)
            adv, ret = ppo.estimate_advantages(batch.rewards, batch.masks, values, cfg.gamma, cfg.tau)
            fixed_log_probs = self.policy_net.get_log_prob(states, actions).float()
        ind = batch.exps.reshape(-1).nonzero().squeeze(1)
        with matmul_context(self.update_dtype):
            return self._ppo_update(states, actions, ret, adv, fixed_log_probs, ind)

    def _ppo_update(self, states, actions, ret, adv, fixed_log_probs, ind):
        cfg = self.cfg
        return ppo.ppo_update(self.policy_net, self.value_net, self.optimizer_policy, self.optimizer_value, states, actions,
                              ret.reshape(-1, 1), adv.reshape(-1, 1), fixed_log_probs, clip_epsilon=cfg.clip_epsilon,
                              num_epoch=cfg.num_optim_epoch, grad_clip=self.policy_grad_clip, ind=ind)

    # ------------------------------------------------------------------ agent_handmimic.py:311-337
    def optimize_policy(self, epoch):
        self.epoch = epoch
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        self.per_epoch_update(epoch)
        batch, stats = self.sample(self.cfg.min_batch_size)
        torch.cuda.synchronize(self.device)
        t1 = time.perf_counter()
        upd = self.update_params(batch)
        torch.cuda.synchronize(self.device)
        t2 = time.perf_counter()
        log = stats.allreduce()
        if self.cfg.end_reward:
            # env.end_reward = avg_c_reward * gamma / (1 - gamma): used by the NEXT iteration's rollout (:317-319)
            self.end_reward = log["avg_c_reward"] * self.cfg.gamma / (1.0 - self.cfg.gamma)
        return dict(log=log, end_reward=self.end_reward, T_sample=t1 - t0, T_update=t2 - t1, T_total=t2 - t0, samples=batch.T * batch.N * self.world,
                    horizon=batch.T, **upd)

    # ------------------------------------------------------------------ agent_handmimic.py:339-403
    @torch.no_grad()
    def eval_policy(self, motions=None, model_index=0, record=False, max_steps=10000):
        """`eval_policy` of the reference, batched: every test motion runs in its own environment, all at once, in
        mode 'test' (no tracking-error termination, ho_im4.py:655-656) from frame 0 with the mean action and the frozen
        observation normaliser.  Returns the reference's metrics dict averaged over the motions (`per_motion` holds
        them one by one) and, with `record`, the `MotionData` tensors of every step ([step][env, substep, ...])."""
        import numpy as np
        mdl = self.env.models[model_index]
        if motions is None:
            # the held-out last sequence of the model's data set (seq_idx = seq_num - 1, agent_handmimic.py:344)
            motions = self.eval_motions[model_index] if self.eval_motions else [mdl.motion(9)]
        n = len(motions)
        env = HandMimicVecEnv(mdl, n, motions, device=self.device.index, seed=0, auto_reset=False, random_start=False, balance=False)
        env.set_mode("test")
        env.set_reward_weights(self.cfg.reward_vector())
        if record:
            env.set_record(True)
        hb = torch.as_tensor(np.asarray(mdl.arrays["env_hand_body_idx"]).astype(np.int64), device=self.device)
        T = max(int(m["seq_len"]) for m in motions)
        gt = torch.zeros(n, T, 26, device=self.device)
        gtj = torch.zeros(n, T, len(hb), 3, device=self.device)
        for i, m in enumerate(motions):
            L = int(m["seq_len"])
            gt[i, :L] = torch.as_tensor(m["hand_dof_seq"], dtype=torch.float32)
            gtj[i, :L] = torch.as_tensor(m["body_pos_seq"], dtype=torch.float32)
            gt[i, L:], gtj[i, L:] = gt[i, L - 1], gtj[i, L - 1]
        lens = torch.as_tensor([int(m["seq_len"]) for m in motions], device=self.device)
        obs = env.reset(motion_id=list(range(n)), start_idx=[0] * n)
        alive = torch.ones(n, dtype=torch.bool, device=self.device)
        acc = {k: torch.zeros(n, device=self.device) for k in ("pose_err", "pose_steps", "mpjpe", "reward", "steps")}
        info_sum = torch.zeros(n, 9, device=self.device)
        percent = torch.zeros(n, device=self.device)
        frames = []
        ar = torch.arange(n, device=self.device)
        for t in range(max_steps):
            q, _ = env.get_state()
            xpos = env.get_derived()[0]
            ti = torch.remainder(torch.full_like(lens, t), lens)
            w = alive.to(torch.float32)
            # the reference's `res["gt"][6:] - res["pred"][6:]` slices the stacked [T, 26] arrays along TIME: frames 6.. of
            # all 26 hand dofs (agent_handmimic.py:386)
            if t >= 6:
                acc["pose_err"] += w * torch.linalg.norm(gt[ar, ti] - q[:, :26], dim=-1)
                acc["pose_steps"] += w
            acc["mpjpe"] += w * torch.linalg.norm(gtj[ar, ti] - xpos[:, hb], dim=-1).mean(-1)
            x = self.running_state(obs, update=False)
            act = self.policy_net(x)[0]                      # select_action(mean_action=True)
            obs, r, done, info = env.step(act)
            acc["reward"] += w * r
            acc["steps"] += w
            info_sum += w[:, None] * info["reward_info"]
            percent = torch.where(alive, info["percent"], percent)
            if record:
                frames.append({k: v.clone() for k, v in env.motion_data().items()})
            alive = alive & ~done.bool()
            if not bool(alive.any()):
                break
        steps = acc["steps"].clamp(min=1)
        names = ("pose_reward", "wpose_reward", "jpos_reward", "vel_reward", "obj_pos_reward", "obj_rot_reward", "obj_vel_reward",
                 "obj_rfc_reward")
        per = dict(pose_err=acc["pose_err"] / acc["pose_steps"].clamp(min=1), mpjpe=acc["mpjpe"] / steps, avg_reward=acc["reward"] / steps,
                   total_reward=acc["reward"], percent=percent, **{k: info_sum[:, i] / steps for i, k in enumerate(names)})
        per = {k: v.cpu().numpy() for k, v in per.items()}
        metrics = {k: float(v.mean()) for k, v in per.items()}
        metrics["per_motion"] = per
        if record:
            metrics["motion_data"] = frames
        env.close()
        return metrics

    # ------------------------------------------------------------------ agent_handmimic.py:135-186
    def save_checkpoint(self, path):
        ppo.save_checkpoint(path, self.policy_net, self.value_net, self.running_state)

    def load_checkpoint(self, path):
        """Opens a checkpoint written here or a released `iter_XXXX.p` of the reference (same pickle layout)."""
        return ppo.load_checkpoint(path, self.policy_net, self.value_net, self.running_state)

    def close(self):
        self.env.close()
