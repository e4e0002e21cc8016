"""Real-data ingest: the host-side mirror of `DatasetSingleDepth` (uhc/data_loaders/dataset_singledepth.py:20-330) with
the per-sequence preprocessing (`preprocess_seq` :78-142: clipping, finite-difference velocities, FK of the hand bodies,
hand-object contact pre-pass) done on the GPU by `hoic_preprocess_seq` (include/hoic.h), one warp per frame.

The expert pickle has the reference's schema: `{seq_name: [ {"hand_pose_seq": [T,26], "obj_pose_seq": [T,7],
"conf": [T] (optional), "obj_init_pos": [3] (optional), "gt_hand_pose_seq" / "gt_obj_pose_seq" (optional)}, ... ]}`.
There is no CPU path: the derivations need the CUDA library and a device.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np
import torch

from . import capi, motion as hmotion
from .assets import Model


class DevicePreprocessor:
    """`hoic_preprocess_seq` for one model: raw (hand dofs, object pose) sequence -> expert dict on the device."""

    def __init__(self, model: Model, device=0):
        if not torch.cuda.is_available():
            raise capi.HoicError("DevicePreprocessor needs a CUDA device (there is no CPU path)")
        self.L = capi.lib()
        self.model, self.device = model, torch.device("cuda", device)
        h = self.L.hoic_model_create(model.blob, len(model.blob), device)
        if not h:
            raise capi.HoicError("hoic_model_create: " + self.L.hoic_last_error().decode())
        self._hm = C.c_void_p(h)

    def __call__(self, hand_pose_seq, obj_pose_seq, motion_freq=30.0, contacts=True, clip=True):
        hand = torch.as_tensor(np.asarray(hand_pose_seq), dtype=torch.float32, device=self.device).contiguous()
        obj = torch.as_tensor(np.asarray(obj_pose_seq), dtype=torch.float32, device=self.device).contiguous()
        T = hand.shape[0]
        assert hand.shape == (T, 26) and obj.shape == (T, 7)
        frames = torch.empty(T, capi.FRAME_DIM, dtype=torch.float32, device=self.device)
        cmask = torch.zeros(T, capi.NHG, dtype=torch.int32, device=self.device) if contacts else None
        cpos = torch.zeros(T, capi.NHG, 3, dtype=torch.float32, device=self.device) if contacts else None
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        capi.check(self.L.hoic_preprocess_seq(self._hm, capi.ptr(hand), capi.ptr(obj), T, C.c_float(motion_freq), 1 if clip else 0, capi.ptr(frames),
                                              capi.ptr(cmask), capi.ptr(cpos), stream), "hoic_preprocess_seq")
        return frames, cmask, cpos

    def close(self):
        if getattr(self, "_hm", None):
            self.L.hoic_model_destroy(self._hm)
            self._hm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_frames(frames: np.ndarray) -> dict:
    """[T,212] frame table -> expert dict with the reference's keys (inverse of motion.pack_frames)."""
    T = frames.shape[0]
    shapes = {"hand_dof_seq": (26,), "hand_dof_vel_seq": (26,), "obj_pose_seq": (7,), "obj_vel_seq": (3,),
              "obj_angle_vel_seq": (3,), "body_pos_seq": (21, 3), "body_quat_seq": (21, 4)}
    out, k = {}, 0
    for key in hmotion.EXPERT_KEYS:
        n = int(np.prod(shapes[key]))
        out[key] = frames[:, k:k + n].reshape((T,) + shapes[key])
        k += n
    assert k == frames.shape[1]
    out["seq_len"] = T
    return out


class DatasetSingleDepth:
    """Same constructor arguments, attributes and `load_seq` / `load_frame` / `get_len` / `get_seq_num` /
    `get_obj_init_pos` surface as the reference class; `expert_seq[i]` holds numpy arrays ready for
    `HandMimicVecEnv.set_motions`."""

    def __init__(self, model: Model, data_spec, noise=0.0, trans=(0, 0), mode="train", idx=None, step=1, device=0, seed=0):
        expert_fn = data_spec["expert_fn"] if mode == "train" else data_spec["test_expert_fn"]
        with open(expert_fn, "rb") as f:
            self.expert_data = pickle.load(f)
        self.seq_name = data_spec["seq_name"]
        self.all_seqs = self.expert_data[self.seq_name]
        self.seq_num = len(self.all_seqs)
        self.single_clip_s = data_spec.get("single_clip_s")
        self.single_clip_e = data_spec.get("single_clip_e")
        self.noise, self.trans = noise, trans
        self.max_len = data_spec.get("max_len", 600)
        self.motion_freq, self.sample_freq = data_spec["motion_freq"], data_spec["sample_freq"]
        self.hand_dof = np.asarray(self.all_seqs[0]["hand_pose_seq"]).shape[-1]
        self.model = model
        A = model.arrays
        self.joint_lower_limit, self.joint_upper_limit = A["jnt_range"][:-1, 0].copy(), A["jnt_range"][:-1, 1].copy()
        self.has_gt = False
        self.expert_seq = []
        self._pre = DevicePreprocessor(model, device)
        self._rng = np.random.RandomState(seed)
        if idx is None:
            self.all_seqs = [self.all_seqs[i] for i in range(0, len(self.all_seqs), step)]
            for seq in self.all_seqs:
                self.preprocess_seq(seq)
        else:
            self.preprocess_seq(self.all_seqs[idx])
        self.data_keys, self.curr_key, self.fr_start = [self.seq_name], self.seq_name, 0

    def preprocess_seq(self, seq):
        """dataset_singledepth.py:78-142.  Noise and the table-plane translation are applied to the raw arrays on the host
        (as the reference does, in place), everything derived comes from the device kernel."""
        hand = np.array(seq["hand_pose_seq"], dtype=np.float64)
        raw_hand = hand.copy()
        obj = np.array(seq["obj_pose_seq"], dtype=np.float64)
        conf = seq.get("conf")
        gt_hand, gt_obj = seq.get("gt_hand_pose_seq"), seq.get("gt_obj_pose_seq")
        if gt_hand is not None and gt_obj is not None:
            self.has_gt = True
        if self.noise > 0.0:
            scale = self.joint_upper_limit - self.joint_lower_limit
            scale[:6] = 0
            hand += 2 * self.noise * scale * (self._rng.random_sample(hand.shape) - 0.5)
        hand[:, 0] += self.trans[0]; hand[:, 1] += self.trans[1]
        obj[:, 0] += self.trans[0]; obj[:, 1] += self.trans[1]
        frames, cmask, cpos = self._pre(hand, obj, self.motion_freq)
        raw_frames, _, _ = self._pre(raw_hand, obj, self.motion_freq, contacts=False, clip=False)
        ex = unpack_frames(frames.cpu().numpy().astype(np.float64))
        ex["raw_hand_dof_seq"] = raw_hand
        ex["raw_body_pos_seq"] = unpack_frames(raw_frames.cpu().numpy().astype(np.float64))["body_pos_seq"]
        cmask, cpos = cmask.cpu().numpy(), cpos.cpu().numpy().astype(np.float64)
        g0 = int(self.model.arrays["env_ids"][2])     # hand_geom_range[0]
        ex["contact_info_seq"] = np.array([{g0 + int(k): cpos[t, k].copy() for k in np.nonzero(cmask[t])[0]} for t in range(len(cmask))],
                                          dtype=object)
        ex["conf_seq"] = None if conf is None else np.asarray(conf)
        ex["obj_init_pos"] = seq.get("obj_init_pos")
        ex["gt_hand_dof_seq"], ex["gt_obj_pose_seq"] = gt_hand, gt_obj
        ex["gt_body_pos_seq"] = ex["gt_body_quat_seq"] = None
        if self.has_gt:
            gtf, _, _ = self._pre(np.asarray(gt_hand), obj, self.motion_freq, contacts=False, clip=False)
            g = unpack_frames(gtf.cpu().numpy().astype(np.float64))
            ex["gt_body_pos_seq"], ex["gt_body_quat_seq"] = g["body_pos_seq"], g["body_quat_seq"]
        if self.single_clip_s is not None:
            s, e = self.single_clip_s, self.single_clip_e
            for k, v in list(ex.items()):
                if isinstance(v, np.ndarray) and v.shape[:1] == (ex["seq_len"],):
                    ex[k] = v[s:e]
            ex["seq_len"] = e - s
        self.expert_seq.append(ex)

    # ---- accessors (dataset_singledepth.py:239-311)
    _SEQ_KEYS = ("hand_dof_seq", "raw_hand_dof_seq", "hand_dof_vel_seq", "obj_pose_seq", "obj_vel_seq", "obj_angle_vel_seq",
                 "body_pos_seq", "raw_body_pos_seq", "body_quat_seq", "contact_info_seq")

    def load_seq(self, seq_idx, start_idx=0, full_seq=False, end_idx=-1):
        ex = self.expert_seq[seq_idx]
        n = ex["seq_len"]
        clip_s = min(start_idx, n)
        clip_e = min(clip_s + self.max_len, n)
        gt = {k: None for k in ("gt_hand_dof_seq", "gt_obj_pose_seq", "gt_body_pos_seq", "gt_body_quat_seq")}
        if self.has_gt:
            sel = np.arange(clip_s, clip_e, self.sample_freq)
            gt = {k: np.asarray(ex[k])[sel] for k in gt}
        if full_seq:
            clip_e = n
        if end_idx > 0:
            clip_e = end_idx
        sel = np.arange(clip_s, clip_e, self.sample_freq)
        out = {k: ex[k][sel] for k in self._SEQ_KEYS}
        out.update(gt)
        out["conf_seq"] = ex["conf_seq"][sel] if ex["conf_seq"] is not None else None
        out["seq_len"] = len(sel)
        return out

    def load_frame(self, seq_idx, ind):
        ex = self.expert_seq[seq_idx]
        out = {k: ex[k][ind] for k in self._SEQ_KEYS}
        out["conf_seq"] = ex["conf_seq"][ind] if ex["conf_seq"] is not None else None
        return out

    def get_len(self, seq_idx):
        return self.expert_seq[seq_idx]["seq_len"]

    def get_seq_num(self):
        return len(self.expert_seq)

    def get_obj_init_pos(self, seq_idx):
        p = self.expert_seq[seq_idx].get("obj_init_pos")
        return np.zeros(3) if p is None else p
