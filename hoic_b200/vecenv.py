"""Batched HandMimic environments on one GPU: the host-side mirror of the reference's env surface
(`HandObjMimic4.reset/step -> obs, reward, done, info`, uhc/envs/ho_im4.py:611-662, 690-716) above the C ABI.

PyTorch is used only for device memory and streams; every number is produced by libhoic_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi, motion as hmotion
from .assets import Model


class HandMimicVecEnv:
    """N independent environments; env i uses models[env_model_id[i]].

    `motions` is a list of expert dicts (hoic_b200.motion) and `motion_model[k]` the index of the model motion k was
    made for; an env only ever draws motions of its own model.  With `auto_reset`, an env that finishes (`done`) is
    reset inside `step` to a freshly drawn (motion, start_idx) exactly as the sampler does between episodes
    (uhc/agents/agent_handmimic.py:437-454) and the returned observation is the reset observation.
    """

    obs_dim, act_dim = capi.OBS_DIM, capi.ACT_DIM

    def __init__(self, models, n_envs, motions, motion_model=None, env_model_id=None, device=0, seed=1,
                 auto_reset=True, random_start=True, min_episode=200, balance=True):
        if isinstance(models, Model):
            models = [models]
        if not torch.cuda.is_available():
            raise capi.HoicError("HandMimicVecEnv needs a CUDA device (there is no CPU path)")
        self.L = capi.lib()
        self.models, self.n_envs, self.device_index = models, int(n_envs), int(device)
        self.device = torch.device("cuda", device)
        self.auto_reset, self.random_start, self.min_episode = auto_reset, random_start, min_episode
        torch.cuda.set_device(self.device)
        self._hm = []
        for m in models:
            h = self.L.hoic_model_create(m.blob, len(m.blob), device)
            if not h:
                raise capi.HoicError("hoic_model_create: " + self.L.hoic_last_error().decode())
            self._hm.append(C.c_void_p(h))
        if env_model_id is None:
            env_model_id = np.arange(n_envs, dtype=np.int32) % len(models)
        self.env_model_id = np.ascontiguousarray(env_model_id, dtype=np.int32)
        arr = (C.c_void_p * len(models))(*[h.value for h in self._hm])
        hb = self.L.hoic_batch_create(arr, len(models), capi.ptr(self.env_model_id), n_envs, device)
        if not hb:
            raise capi.HoicError("hoic_batch_create: " + self.L.hoic_last_error().decode())
        self._hb = C.c_void_p(hb)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        N = self.n_envs
        f32 = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros(N, capi.OBS_DIM, **f32)
        self.reward = torch.zeros(N, **f32)
        self.reward_info = torch.zeros(N, capi.INFO_DIM, **f32)
        self.percent = torch.zeros(N, **f32)
        self.done = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.fail = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.end = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._seed = int(seed)
        self._hs = None
        self._mode = "train"     # the library's default (hoic_batch_create)
        self.graph_epoch = 0     # bumped whenever something a captured CUDA graph bakes in changes (mode, recording, tables)
        self._ep_mid = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._ep_start = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._ep_t = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.set_motions(motions, motion_model)
        self.set_reward_weights(models[0].cfg.reward_vector())
        self.balance = balance

    # ------------------------------------------------------------------ configuration
    def set_motions(self, motions, motion_model=None):
        if motion_model is None:
            motion_model = [0] * len(motions)
        frames = [hmotion.pack_frames(e).astype(np.float32) for e in motions]
        lens = np.array([f.shape[0] for f in frames], dtype=np.int32)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        allf = np.ascontiguousarray(np.concatenate(frames, 0))
        capi.check(self.L.hoic_set_motions(self._hb, capi.ptr(allf), capi.ptr(offs), capi.ptr(lens), len(frames)), "hoic_set_motions")
        self.motion_len = torch.as_tensor(lens, device=self.device)
        # per-model motion tables for on-device sampling
        nm = len(self.models)
        per = [[k for k, mm in enumerate(motion_model) if mm == i] for i in range(nm)]
        if any(len(p) == 0 for p in per):
            raise ValueError("every model needs at least one motion")
        width = max(len(p) for p in per)
        table = np.zeros((nm, width), dtype=np.int64)
        for i, p in enumerate(per):
            table[i, :len(p)] = p
            table[i, len(p):] = p[0]
        self._motion_table = torch.as_tensor(table, device=self.device)
        self._motion_count = torch.as_tensor([len(p) for p in per], device=self.device)
        self._env_model = torch.as_tensor(self.env_model_id.astype(np.int64), device=self.device)
        # the same tables for the library's on-device sampler (auto-reset inside hoic_step, hoic_reset_sampled)
        offs_m = np.concatenate([[0], np.cumsum([len(p) for p in per])]).astype(np.int32)
        ids = np.ascontiguousarray(np.concatenate(per), dtype=np.int32)
        capi.check(self.L.hoic_set_sampler(self._hb, capi.ptr(offs_m), capi.ptr(ids), 1 if self.random_start else 0,
                                           int(self.min_episode), self._seed), "hoic_set_sampler")
        self._hs = None     # the table may have been reallocated: a captured host-step graph holds stale pointers
        self.graph_epoch += 1

    def set_motion_frames(self, frames):
        """Replace the frames of ALL motions in place: `frames` [n_motions, T, 212] with the same n_motions / T as the
        current table (one host->device copy; used by the sliding-window tracking env, hoic_b200/track.py)."""
        f = np.ascontiguousarray(frames, dtype=np.float32)
        nm, T = f.shape[0], f.shape[1]
        assert f.shape[2] == capi.FRAME_DIM and nm == int(self.motion_len.numel()) and bool((self.motion_len == T).all())
        lens = np.full(nm, T, dtype=np.int32)
        offs = (np.arange(nm, dtype=np.int32) * T).astype(np.int32)
        capi.check(self.L.hoic_set_motions(self._hb, capi.ptr(f.reshape(nm * T, capi.FRAME_DIM)), capi.ptr(offs), capi.ptr(lens), nm), "hoic_set_motions")

    def set_reward_weights(self, w24):
        w = np.ascontiguousarray(w24, dtype=np.float32)
        assert w.size == 24
        capi.check(self.L.hoic_set_reward_weights(self._hb, capi.ptr(w)), "hoic_set_reward_weights")

    def set_mode(self, mode):
        capi.check(self.L.hoic_set_mode(self._hb, 1 if mode == "train" else 0), "hoic_set_mode")
        if getattr(self, "_mode", None) != mode:      # (the sampler sets the mode at the start of every rollout: no change, no recapture)
            self._mode = mode
            self._hs = None     # kernel arguments are baked into a captured graph: recapture
            self.graph_epoch += 1

    @property
    def balance(self):
        return self._balance

    @balance.setter
    def balance(self, on):
        self._balance = bool(on)
        capi.check(self.L.hoic_set_balance(self._hb, 1 if on else 0), "hoic_set_balance")
        self._hs = None
        self.graph_epoch += 1

    def set_record(self, on=True):
        """Test-mode recording of `MotionData` (ho_im4.py:27-41): with it on, every `step` also fills the per-substep
        frames of `compute_data_frame` (:599-609) and the surface-contact record of `solve_rfc` (:1072-1081)."""
        if on:
            sim_step = int(self.models[0].cfg.sim_step)
            self._rec_frames = torch.zeros(self.n_envs, sim_step, capi.REC_STRIDE, dtype=torch.float32, device=self.device)
            self._rec_surface = torch.zeros(self.n_envs, capi.SURF_STRIDE, dtype=torch.float32, device=self.device)
        else:
            self._rec_frames = self._rec_surface = None
        capi.check(self.L.hoic_set_record(self._hb, capi.ptr(self._rec_frames), capi.ptr(self._rec_surface)), "hoic_set_record")
        self._hs = None
        self.graph_epoch += 1

    def motion_data(self):
        """Record of the last `step` as a dict of device tensors named like the reference's `MotionData` lists; the
        leading dims are [env, substep].  `valid` marks substeps that ran, `n_contact` / `n_surface` the filled rows."""
        if getattr(self, "_rec_frames", None) is None:
            raise capi.HoicError("motion_data: call set_record(True) first")
        f, u = self._rec_frames, self._rec_surface
        N, S = f.shape[0], f.shape[1]
        con = f[:, :, capi.REC_CON:].reshape(N, S, capi.REC_MAXCON, 19)
        pts = u[:, capi.SURF_PT:capi.SURF_PT + 15 * capi.SURF_MAXPT].reshape(N, capi.SURF_MAXPT, 15)
        return dict(hand_kps=f[:, :, 0:63].reshape(N, S, 21, 3), hand_qpos=f[:, :, 63:89], compensate_ft=f[:, :, 89:95],
                    obj_pose=f[:, :, 95:102], obj_vel=f[:, :, 102:108], obj_acc=f[:, :, 108:114],
                    n_contact=f[:, :, 114].to(torch.int32), valid=f[:, :, 115] > 0,
                    contact_frames=con[..., 0:12], contact_force=con[..., 12:18], contact_body=con[..., 18].to(torch.int32),
                    n_surface=u[:, 0].to(torch.int32), rest_obj_ft=u[:, 1:7], surface_valid=u[:, 7] > 0,
                    surface_contact_frames=pts[..., 0:12], surface_contact_force=pts[..., 12:15])

    # ------------------------------------------------------------------ episodes
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sample_assignments(self):
        """Random (motion_id, start_idx) per env, as agent_handmimic.py:439-448, drawn with torch's generator (explicit
        resets; the auto-reset inside `step` and `reset()` without arguments use the library's on-device sampler)."""
        N = self.n_envs
        cnt = self._motion_count[self._env_model]
        u = torch.rand(N, device=self.device, generator=self.gen)
        pick = torch.minimum((u * cnt).long(), cnt - 1)
        mid = self._motion_table[self._env_model, pick]
        if self.random_start:
            span = torch.clamp(self.motion_len[mid] - self.min_episode, min=1)
            u2 = torch.rand(N, device=self.device, generator=self.gen)
            start = torch.minimum((u2 * span).long(), span - 1)
        else:
            start = torch.zeros(N, dtype=torch.long, device=self.device)
        return mid.to(torch.int32).contiguous(), start.to(torch.int32).contiguous()

    def _episode(self):
        capi.check(self.L.hoic_get_episode(self._hb, capi.ptr(self._ep_mid), capi.ptr(self._ep_start), capi.ptr(self._ep_t),
                                           self._stream()), "hoic_get_episode")

    @property
    def motion_id(self):
        """Motion index of every env's current episode ([N] int32, device)."""
        self._episode()
        return self._ep_mid

    @property
    def start_idx(self):
        self._episode()
        return self._ep_start

    @property
    def cur_t(self):
        self._episode()
        return self._ep_t

    def reset(self, mask=None, motion_id=None, start_idx=None):
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        if motion_id is None and start_idx is None:
            capi.check(self.L.hoic_reset_sampled(self._hb, capi.ptr(mask), capi.ptr(self.obs), self._stream()), "hoic_reset_sampled")
            return self.obs
        if motion_id is not None:
            motion_id = torch.as_tensor(motion_id, dtype=torch.int32, device=self.device).contiguous()
        if start_idx is not None:
            start_idx = torch.as_tensor(start_idx, dtype=torch.int32, device=self.device).contiguous()
        capi.check(self.L.hoic_reset(self._hb, capi.ptr(mask), capi.ptr(motion_id), capi.ptr(start_idx), capi.ptr(self.obs),
                                     self._stream()), "hoic_reset")
        return self.obs

    def step(self, actions):
        """`env.step` for the whole batch: ONE library call (env ordering, the step pipeline and, with `auto_reset`, the
        restart of finished envs on a freshly drawn episode all happen inside hoic_step)."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        assert a.shape == (self.n_envs, capi.ACT_DIM)
        capi.check(self.L.hoic_step(self._hb, capi.ptr(a), capi.ptr(self.obs), capi.ptr(self.reward), capi.ptr(self.reward_info),
                                    capi.ptr(self.done), capi.ptr(self.fail), capi.ptr(self.end), capi.ptr(self.percent),
                                    1 if self.auto_reset else 0, self._stream()), "hoic_step")
        info = {"fail": self.fail, "end": self.end, "percent": self.percent, "reward_info": self.reward_info}
        return self.obs, self.reward, self.done, info

    # ------------------------------------------------------------------ host-buffer step (the reference-facing call)
    def step_host(self, actions_host):
        """`env.step` for callers that live on the host, like the reference's sampler: `actions_host` is a [N, 32] float32
        array / CPU tensor; returns (obs, reward, done, info) as pinned CPU tensors that stay valid until the next call.
        The H2D copy of the actions, env ordering, the step kernel, the in-place auto-reset and the D2H copies of the
        results are captured once into a CUDA graph and replayed as ONE launch per step."""
        if self._hs is None:
            self._build_host_step()
        hs = self._hs
        src = torch.as_tensor(actions_host, dtype=torch.float32)
        if src.data_ptr() != hs["act"].data_ptr():
            hs["act"].copy_(src.reshape(self.n_envs, capi.ACT_DIM))
        if hs["graph"] is not None:
            hs["graph"].replay()
        else:
            self._host_step_body()
        torch.cuda.current_stream(self.device).synchronize()
        info = {"fail": hs["fail"], "end": hs["end"], "percent": hs["percent"], "reward_info": hs["reward_info"]}
        return hs["obs"], hs["reward"], hs["done"], info

    def _host_step_body(self):
        hs = self._hs
        hs["d_act"].copy_(hs["act"], non_blocking=True)
        obs, rew, done, info = self.step(hs["d_act"])
        hs["obs"].copy_(obs, non_blocking=True)
        hs["reward"].copy_(rew, non_blocking=True)
        hs["done"].copy_(done, non_blocking=True)
        hs["fail"].copy_(info["fail"], non_blocking=True)
        hs["end"].copy_(info["end"], non_blocking=True)
        hs["percent"].copy_(info["percent"], non_blocking=True)
        hs["reward_info"].copy_(info["reward_info"], non_blocking=True)

    def _build_host_step(self):
        N = self.n_envs
        pin = lambda *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype).pin_memory()
        self._hs = dict(act=pin(N, capi.ACT_DIM), obs=pin(N, capi.OBS_DIM), reward=pin(N), percent=pin(N),
                        reward_info=pin(N, capi.INFO_DIM), done=pin(N, dtype=torch.uint8), fail=pin(N, dtype=torch.uint8),
                        end=pin(N, dtype=torch.uint8), d_act=torch.zeros(N, capi.ACT_DIM, device=self.device), graph=None)
        self._hs["act"].zero_()
        self.h2d_bytes = N * capi.ACT_DIM * 4
        self.d2h_bytes = N * (capi.OBS_DIM * 4 + 4 + 4 + capi.INFO_DIM * 4 + 3)
        try:
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):      # warm up the copies of the body without touching the environments
                self._hs["d_act"].copy_(self._hs["act"], non_blocking=True)
                self._hs["obs"].copy_(self.obs, non_blocking=True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._host_step_body()
            torch.cuda.synchronize(self.device)
            self._hs["graph"] = g
        except Exception as e:                     # no graph support in this torch build: plain launches
            import warnings
            warnings.warn(f"step_host: CUDA graph capture unavailable ({e}); using eager launches")
            self._hs["graph"] = None

    def get_state(self):
        q = torch.empty(self.n_envs, capi.NQ, dtype=torch.float32, device=self.device)
        v = torch.empty(self.n_envs, capi.NV, dtype=torch.float32, device=self.device)
        capi.check(self.L.hoic_get_state(self._hb, capi.ptr(q), capi.ptr(v), self._stream()), "hoic_get_state")
        return q, v

    def set_state(self, qpos, qvel):
        q = torch.as_tensor(qpos, dtype=torch.float32, device=self.device).contiguous()
        v = torch.as_tensor(qvel, dtype=torch.float32, device=self.device).contiguous()
        assert q.shape == (self.n_envs, capi.NQ) and v.shape == (self.n_envs, capi.NV)
        capi.check(self.L.hoic_set_state(self._hb, capi.ptr(q), capi.ptr(v), self._stream()), "hoic_set_state")

    def set_cur_t(self, t):
        t = torch.as_tensor(t, dtype=torch.int32, device=self.device).contiguous()
        capi.check(self.L.hoic_set_cur_t(self._hb, capi.ptr(t), self._stream()), "hoic_set_cur_t")

    def get_warmstart(self):
        w = torch.empty(self.n_envs, capi.NV, dtype=torch.float32, device=self.device)
        capi.check(self.L.hoic_get_warmstart(self._hb, capi.ptr(w), self._stream()), "hoic_get_warmstart")
        return w

    def set_warmstart(self, w):
        w = torch.as_tensor(w, dtype=torch.float32, device=self.device).contiguous()
        capi.check(self.L.hoic_set_warmstart(self._hb, capi.ptr(w), self._stream()), "hoic_set_warmstart")

    def get_contacts(self):
        N, K = self.n_envs, capi.MAX_CONTACTS
        ncon = torch.empty(N, dtype=torch.int32, device=self.device)
        pairs = torch.empty(N, K, 2, dtype=torch.int32, device=self.device)
        data = torch.empty(N, K, 13, dtype=torch.float32, device=self.device)
        capi.check(self.L.hoic_get_contacts(self._hb, capi.ptr(ncon), capi.ptr(pairs), capi.ptr(data), self._stream()), "hoic_get_contacts")
        return ncon, pairs, data

    def get_derived(self):
        N = self.n_envs
        nb = int(self.models[0].arrays["nbody"][0])
        f32 = dict(dtype=torch.float32, device=self.device)
        xpos, xquat = torch.empty(N, nb, 3, **f32), torch.empty(N, nb, 4, **f32)
        qM, bias = torch.empty(N, 32, 32, **f32), torch.empty(N, 32, **f32)
        capi.check(self.L.hoic_get_derived(self._hb, capi.ptr(xpos), capi.ptr(xquat), capi.ptr(qM), capi.ptr(bias), self._stream()), "hoic_get_derived")
        return xpos, xquat, qM, bias

    def substep_debug(self, ctrl, qfrc_applied):
        N = self.n_envs
        c = torch.as_tensor(ctrl, dtype=torch.float32, device=self.device).contiguous()
        q = torch.as_tensor(qfrc_applied, dtype=torch.float32, device=self.device).contiguous()
        qacc = torch.empty(N, 32, dtype=torch.float32, device=self.device)
        fcon = torch.empty(N, 32, dtype=torch.float32, device=self.device)
        stats = torch.empty(N, 4, dtype=torch.int32, device=self.device)
        capi.check(self.L.hoic_substep_debug(self._hb, capi.ptr(c), capi.ptr(q), capi.ptr(qacc), capi.ptr(fcon), capi.ptr(stats),
                                             self._stream()), "hoic_substep_debug")
        return qacc, fcon, stats

    def get_rfc(self):
        out = torch.empty(self.n_envs, 8, dtype=torch.float32, device=self.device)
        capi.check(self.L.hoic_get_rfc(self._hb, capi.ptr(out), self._stream()), "hoic_get_rfc")
        return out

    def counters(self):
        c = capi.Counters()
        capi.check(self.L.hoic_get_counters(self._hb, C.byref(c)), "hoic_get_counters")
        return c.as_dict()

    def close(self):
        self._hs = None
        if getattr(self, "_hb", None):
            self.L.hoic_batch_destroy(self._hb)
            self._hb = None
            for h in self._hm:
                self.L.hoic_model_destroy(h)
            self._hm = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
