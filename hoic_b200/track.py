"""Online tracking environment: the batched mirror of `HandObjMimicTest` (uhc/envs/ho_im_test.py:26-91), the env the
reference's inference server drives (InferenceServer/RLTest.py): the expert is not a recorded sequence but a sliding window
of `future_w_size + 1` frames that the tracker refills one frame per step (`insert_new_frame`), and every expert lookup is
relative to the window (`get_expert_attr(attr, delta_t)`), whatever `cur_t` is.

It runs on the same kernels as training.  The kernels index the expert as `cur_t + delta_t` (ho_im4.py:739-792), use
`cur_t` for the PD target during the substeps and `cur_t + 1` for termination / reward / observation after them, so the
window `w0 .. w5` is laid out per env as a 7-frame motion

    after reset  (cur_t = 0):  [w0, w1, w2, w3, w4, w5, w5]   state <- w0, observation targets w1 .. w5
    for a step   (cur_t = 0):  [w0, w0, w1, w2, w3, w4, w5]   PD tracks w0; afterwards (cur_t = 1) the tracking error and the
                                                             reward are against w0, the observation targets are w1 .. w5

which is exactly what the window-relative accessors of ho_im_test.py:40-75 return.  `end` is meaningless here (the window
never ends) and is not reported; `fail` keeps the tracking-error thresholds (the reference builds this env in mode 'train').
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi, motion as hmotion
from .assets import Model
from .vecenv import HandMimicVecEnv


def _pack(frames):
    """[..., 212] array, or an expert dict / list of single-frame dicts with the reference's keys -> float32 [..., 212]."""
    if isinstance(frames, dict):
        return hmotion.pack_frames(frames).astype(np.float32)
    if isinstance(frames, (list, tuple)) and frames and isinstance(frames[0], dict):
        return np.stack([np.concatenate([np.asarray(f[k], dtype=np.float64).reshape(-1) for k in hmotion.EXPERT_KEYS]) for f in frames]).astype(np.float32)
    return np.asarray(frames, dtype=np.float32)


class HandMimicTrackEnv:
    def __init__(self, model: Model, n_envs, init_window, device=0, seed=1):
        """`init_window`: [n_envs, w_size + 1, 212] packed frames (or one expert dict of w_size + 1 frames shared by all envs)."""
        self.model, self.n_envs = model, int(n_envs)
        self.w_size = int(model.cfg.future_w_size) if hasattr(model.cfg, "future_w_size") else 5
        W = self.w_size + 1
        w = _pack(init_window)
        if w.ndim == 2:
            w = np.broadcast_to(w[None], (self.n_envs,) + w.shape)
        assert w.shape == (self.n_envs, W, capi.FRAME_DIM), w.shape
        self.expert_window = np.array(w, dtype=np.float32)          # [N, W, 212], slot order = age (index 0 = current frame)
        motions = [self._as_expert(self._reset_layout()[i]) for i in range(self.n_envs)]
        self.env = HandMimicVecEnv(model, self.n_envs, motions, motion_model=[0] * self.n_envs, device=device, seed=seed,
                                   auto_reset=False, random_start=False, balance=False)
        self._ids = list(range(self.n_envs))
        self.obs_dim, self.act_dim = self.env.obs_dim, self.env.act_dim

    # ---- window bookkeeping (ho_im_test.py:34-43)
    def insert_new_frame(self, frame):
        """Drop the oldest frame of every env's window and append `frame` ([n_envs, 212], [212], or a frame dict)."""
        f = _pack(frame)
        if f.ndim == 1:
            f = np.broadcast_to(f[None], (self.n_envs, capi.FRAME_DIM))
        self.expert_window = np.concatenate([self.expert_window[:, 1:], f[:, None]], axis=1)

    def get_expert_attr(self, attr, ind, env=0):
        assert ind <= self.w_size
        ex = self._as_expert(self.expert_window[env])
        return ex[attr][ind].copy()

    @staticmethod
    def _as_expert(frames):
        from .dataset import unpack_frames
        return unpack_frames(np.asarray(frames, dtype=np.float64))

    def _reset_layout(self):
        w = self.expert_window
        return np.concatenate([w, w[:, -1:]], axis=1)            # [w0 .. w5, w5]

    def _step_layout(self):
        w = self.expert_window
        return np.concatenate([w[:, :1], w], axis=1)             # [w0, w0 .. w5]

    # ---- env surface
    def reset(self):
        """ho_im_test.py:77-91: state <- the window's current frame, observation against the rest of the window."""
        self.env.set_motion_frames(self._reset_layout())
        return self.env.reset(motion_id=self._ids, start_idx=[0] * self.n_envs)

    def step(self, actions):
        self.env.set_motion_frames(self._step_layout())
        self.env.set_cur_t([0] * self.n_envs)
        obs, reward, done, info = self.env.step(actions)
        return obs, reward, info["fail"], {"fail": info["fail"], "reward_info": info["reward_info"]}

    def set_mode(self, mode):
        self.env.set_mode(mode)

    def get_state(self):
        return self.env.get_state()

    def close(self):
        self.env.close()
