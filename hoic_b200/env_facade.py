"""`HandObjMimic4`-shaped single-environment façade over the batched engine: keeps alive the attribute
surface the reference's callers read (`AgentHandMimic.setup_env / sample_process / eval_policy`, `ho_mimic_reward_9`,
scripts/eval_handmimic.py) — `observation_space.shape[0]`, `action_space.shape[0]`, `cur_t`, `start_ind`, `expert_len`,
`ndof`, `vf_dim`, `rfc_score`, `data.qpos`, `get_hand_qpos / qvel()`, `get_obj_qpos / qvel()`, `get_wbody_pos / quat()`,
`get_expert_attr(name, ind)`, `set_expert`, `set_mode`, `seed`, `reset()`, `step(a) -> obs, 1.0, done, info` — with every
number produced by `HandMimicVecEnv(n_envs=1)`, i.e. by libhoic_b200.so.  It exists for drop-in use and debugging; throughput
comes from the batched env.  (uhc/envs/ho_im4.py:46-167, 611-662, 690-716, 739-858.)
"""
from __future__ import annotations

import numpy as np
import torch

from . import motion as hmotion


class _Box:
    """The two members of gym.spaces.Box the reference reads (`shape`, `low` / `high`; mujoco_env.py:66-68, ho_im4.py:157-167)."""

    def __init__(self, n, bound):
        self.shape = (n,)
        self.low, self.high = -bound * np.ones(n), bound * np.ones(n)


class _Data:
    """`env.data.qpos / qvel` and the poses of the last forward (`body_xpos / body_xquat`), read through the C ABI."""

    def __init__(self, owner):
        self._o = owner

    @property
    def qpos(self):
        return self._o._state()[0]

    @property
    def qvel(self):
        return self._o._state()[1]

    @property
    def body_xpos(self):
        return self._o._derived()[0]

    @property
    def body_xquat(self):
        return self._o._derived()[1]


class HandObjMimic4:
    def __init__(self, cfg, expert_seq, model, data_specs=None, mode="train", device=0, seed=1, _vecenv=None):
        """`model` is a compiled `hoic_b200.assets.Model` (the reference passes the merged MJCF here)."""
        self.cc_cfg, self.model_handle, self.data_specs = cfg, model, data_specs
        self.sim_step, self.w_size = int(cfg.sim_step), int(cfg.future_w_size)
        if _vecenv is None:
            from .vecenv import HandMimicVecEnv
            _vecenv = HandMimicVecEnv(model, 1, [expert_seq], device=device, seed=seed, auto_reset=False, random_start=False, balance=False)
        self._env = _vecenv
        A = model.arrays
        self.hand_body_idx = [int(i) for i in np.asarray(A["env_hand_body_idx"])]
        ids = A["env_ids"]
        self.hand_geom_range, self.obj_geom_range = [int(ids[2]), int(ids[3])], [int(ids[4]), int(ids[5])]
        self.ndof, self.vf_dim = 26, 6 if cfg.residual_force else 0
        self.hand_qpos_dim = self.hand_qvel_dim = 26
        self.qpos_dim, self.qvel_dim = 33, 32
        self.observation_space, self.action_space = _Box(617, np.inf), _Box(self.ndof + self.vf_dim, 1.0)
        self.data = _Data(self)
        self.np_random = np.random.RandomState(seed)
        self.mode = mode
        self._env.set_mode(mode)
        self.rfc_score, self.end_reward = 0.0, 0.0
        self._reward, self._reward_info = 0.0, np.zeros(9)
        self.set_expert(expert_seq, _upload=False)

    # ---- ho_im4.py:132-137, mujoco_env.py:70-72
    def set_mode(self, mode):
        self.mode = mode
        self._env.set_mode(mode)

    def set_expert(self, expert_seq, _upload=True):
        self.expert = expert_seq
        self.expert_len = int(expert_seq["hand_dof_seq"].shape[0])
        self.cur_t, self.start_ind = 0, 0
        if _upload:
            self._env.set_motions([expert_seq])

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    # ---- mujoco_env.py:95-105, ho_im4.py:690-716
    def reset(self, tracking=False):
        self.cur_t, self.rfc_score = 0, 0.0
        obs = self._env.reset(motion_id=[0], start_idx=[0])
        return obs[0].double().cpu().numpy()

    # ---- ho_im4.py:611-662; the reward of ho_mimic_reward_9 is computed by the same kernel and kept for `custom_reward`
    def step(self, a):
        act = torch.as_tensor(np.asarray(a, dtype=np.float32)[None, :self.ndof + self.vf_dim])
        if act.shape[1] < 32:
            act = torch.cat([act, torch.zeros(1, 32 - act.shape[1])], 1)
        obs, reward, done, info = self._env.step(act)
        self.cur_t += 1
        self._reward = float(reward[0])
        self._reward_info = info["reward_info"][0].double().cpu().numpy()
        self.rfc_score = float(self._env.get_rfc()[0, 0])
        out = {"fail": bool(info["fail"][0]), "end": bool(info["end"][0]), "percent": float(info["percent"][0])}
        return obs[0].double().cpu().numpy(), 1.0, bool(done[0]), out

    # ---- state accessors (ho_im4.py:794-858)
    def _state(self):
        q, v = self._env.get_state()
        return q[0].double().cpu().numpy(), v[0].double().cpu().numpy()

    def _derived(self):
        xpos, xquat = self._env.get_derived()[:2]
        return xpos[0].double().cpu().numpy(), xquat[0].double().cpu().numpy()

    def get_hand_qpos(self):
        return self.data.qpos[:self.hand_qpos_dim].copy()

    def get_hand_qvel(self):
        return self.data.qvel[:self.hand_qvel_dim].copy()

    def get_obj_qpos(self):
        return self.data.qpos[self.hand_qpos_dim:].copy()

    def get_obj_qvel(self):
        return self.data.qvel[self.hand_qvel_dim:].copy()

    def get_wbody_pos(self, selectList=None):
        return self.data.body_xpos[self.hand_body_idx].copy().ravel()

    def get_wbody_quat(self, selectList=None):
        return self.data.body_xquat[self.hand_body_idx].copy().ravel()

    def get_hand_root_pose(self):
        xpos, xquat = self._derived()
        b = self.hand_body_idx[0]
        return np.concatenate((xpos[b], xquat[b]))

    # ---- expert accessors (ho_im4.py:739-792)
    def get_expert_attr(self, attr, ind):
        ind = min(ind, self.expert_len - 1)
        return np.asarray(self.expert[attr][ind]).copy()

    def get_expert_hand_qpos(self, delta_t=0):
        return self.get_expert_attr("hand_dof_seq", self.cur_t + delta_t + self.start_ind)

    def get_expert_hand_qvel(self, delta_t=0):
        return self.get_expert_attr("hand_dof_vel_seq", self.cur_t + delta_t + self.start_ind)

    def get_expert_obj_pose(self, delta_t=0):
        return self.get_expert_attr("obj_pose_seq", self.cur_t + delta_t + self.start_ind)

    def get_expert_obj_vel(self, delta_t=0):
        ind = self.cur_t + delta_t + self.start_ind
        return np.concatenate([self.get_expert_attr("obj_vel_seq", ind), self.get_expert_attr("obj_angle_vel_seq", ind)])

    def get_expert_wbody_pos(self, delta_t=0):
        return self.get_expert_attr("body_pos_seq", self.cur_t + delta_t + self.start_ind).reshape(-1, 3).flatten()

    def get_expert_wbody_quat(self, delta_t=0):
        return self.get_expert_attr("body_quat_seq", self.cur_t + delta_t + self.start_ind).reshape(-1, 4).flatten()

    def close(self):
        self._env.close()


def ho_mimic_reward_9(env, state, action, info):
    """`custom_reward(env, state, action, info) -> (reward, info[9])` (agent_handmimic.py:171-173, 471-473): the kernel evaluated
    ho_reward.py:943-1046 inside `env.step`, with the epoch-adaptive weights last passed to `set_reward_weights`."""
    return env._reward, env._reward_info.copy()
