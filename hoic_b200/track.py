"""Online tracking environment: the batched mirror of `HandObjMimicTest` (uhc/envs/ho_im_test.py:26-91), the env the
reference's inference server drives (InferenceServer/RLTest.py): the expert is not a recorded sequence but a sliding window
of `future_w_size + 1` frames that the tracker refills one frame per step (`insert_new_frame`), and every expert lookup is
relative to the window (`get_expert_attr(attr, delta_t)`), whatever `cur_t` is.

It runs on the same kernels as training.  The kernels index the expert as `start_idx + cur_t + delta_t` (ho_im4.py:739-792),
use `cur_t` for the PD target during the substeps and `cur_t + 1` for termination / reward / observation after them.  The
window `w0 .. w5` lives ON THE DEVICE as one 8-frame motion per env, `[w0, w0, w1, w2, w3, w4, w5, w5]`:

    reset        (start_idx 1, cur_t 0):  sees [w0, w1, w2, w3, w4, w5, w5]   state <- w0, observation targets w1 .. w5
    step         (start_idx 0, cur_t 0):  sees [w0, w0, w1, w2, w3, w4, w5]   PD tracks w0; afterwards (cur_t = 1) the tracking
                                                                             error and the reward are against w0, the
                                                                             observation targets are w1 .. w5

which is exactly what the window-relative accessors of ho_im_test.py:40-75 return.  `insert_new_frame` is `hoic_push_frame`
(a device-side shift of every window: only the new frame crosses the PCIe bus), `step(actions, values=...)` adds the inference
server's value-threshold reset (InferenceServer/RLTest.py:283-299).  `end` is meaningless here (the window never ends) and is
not reported; `fail` keeps the tracking-error thresholds (the reference builds this env in mode 'train').
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
        table = np.concatenate([w[:, :1], w, w[:, -1:]], axis=1)        # [w0, w0, w1 .. w5, w5] per env
        motions = [self._as_expert(table[i]) for i in range(self.n_envs)]
        self.env = HandMimicVecEnv(model, self.n_envs, motions, motion_model=[0] * self.n_envs, device=device, seed=seed,
                                   auto_reset=False, random_start=False, balance=False)
        dev = self.env.device
        self.expert_window = torch.as_tensor(np.ascontiguousarray(w), dtype=torch.float32, device=dev)   # [N, W, 212], index 0 = current frame
        self._ids = torch.arange(self.n_envs, dtype=torch.int32, device=dev)
        self._zero = torch.zeros(self.n_envs, dtype=torch.int32, device=dev)
        self._one = torch.ones(self.n_envs, dtype=torch.int32, device=dev)
        self.obs_dim, self.act_dim = self.env.obs_dim, self.env.act_dim

    # ---- window bookkeeping (ho_im_test.py:34-43)
    def insert_new_frame(self, frame):
        """Drop the oldest frame of every env's window and append `frame` ([n_envs, 212], [212], a frame dict, or a device tensor)."""
        dev = self.env.device
        f = frame if isinstance(frame, torch.Tensor) else torch.as_tensor(_pack(frame))
        f = f.to(device=dev, dtype=torch.float32)
        per = f.dim() == 2
        f = f.contiguous()
        assert f.shape[-1] == capi.FRAME_DIM and (not per or f.shape[0] == self.n_envs)
        capi.check(self.env.L.hoic_push_frame(self.env._hb, capi.ptr(f), 1 if per else 0, self.env._stream()), "hoic_push_frame")
        self.expert_window = torch.cat([self.expert_window[:, 1:], (f if per else f.expand(self.n_envs, -1))[:, None]], dim=1)

    def get_expert_attr(self, attr, ind, env=0):
        assert ind <= self.w_size
        ex = self._as_expert(self.expert_window[env].cpu().numpy())
        return ex[attr][ind].copy()

    @staticmethod
    def _as_expert(frames):
        from .dataset import unpack_frames
        return unpack_frames(np.asarray(frames, dtype=np.float64))

    # ---- env surface
    def reset(self, mask=None):
        """ho_im_test.py:77-91 (`env.reset(True)`): state <- the window's current frame, observation against the rest of the
        window; `mask` restricts it to some envs (the others keep their state and their row of the returned observation)."""
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.env.device).to(torch.uint8).contiguous()
        e = self.env
        capi.check(e.L.hoic_reset(e._hb, capi.ptr(mask), capi.ptr(self._ids), capi.ptr(self._one), capi.ptr(e.obs), e._stream()), "hoic_reset")
        return e.obs

    def step(self, actions, values=None, reset_threshold=12.0):
        """One tracking step.  `values`: V(s) of the observation the actions were computed from; envs whose value is below
        `reset_threshold` are re-initialised on the window's current frame after the step (RLTest.py:283-299) and
        `info["reset"]` marks them."""
        e = self.env
        capi.check(e.L.hoic_set_episode(e._hb, None, capi.ptr(self._zero), capi.ptr(self._zero), e._stream()), "hoic_set_episode")
        obs, reward, done, info = e.step(actions)
        out = {"fail": info["fail"], "reward_info": info["reward_info"]}
        if values is not None:
            lost = torch.as_tensor(values, device=e.device).reshape(-1) < reset_threshold
            out["reset"] = lost
            obs = self.reset(mask=lost)
        return obs, reward, info["fail"], out

    def set_mode(self, mode):
        self.env.set_mode(mode)

    def get_state(self):
        return self.env.get_state()

    def close(self):
        self.env.close()
