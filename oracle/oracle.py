"""ctypes front for the CPU ORACLE (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; product code must never import it.
PARITY UNPINNED (see oracle/orc_model.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
OBS_DIM, FRAME_DIM = 617, 212
REC_STRIDE, SURF_STRIDE = 116 + 19 * 32, 8 + 15 * 95 + 3


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_physics.c", "orc_collision.c", "orc_env.c", "orc_math.h", "orc_model.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_env_create.restype = C.c_void_p
        L.orc_env_create.argtypes = [C.c_void_p]
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_env_data.restype = C.c_void_p
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One float64 environment.  Mirrors the HandObjMimic4 call surface needed by the parity tests."""

    def __init__(self, blob: bytes):
        L = lib()
        self._blob = blob
        self.model = L.orc_model_create(blob, len(blob))
        if not self.model:
            raise RuntimeError("orc_model_create failed")
        self.env = C.c_void_p(L.orc_env_create(C.c_void_p(self.model)))
        from hoic_b200.blob import unpack  # format helper only
        d = unpack(blob)
        self.nq, self.nv, self.nu = int(d["nq"][0]), int(d["nv"][0]), int(d["nu"][0])
        self.nbody, self.ngeom = int(d["nbody"][0]), int(d["ngeom"][0])
        self.arrays = d

    def close(self):
        L = lib()
        if self.env:
            L.orc_env_destroy(self.env)
            L.orc_model_destroy(C.c_void_p(self.model))
            self.env = None

    # ---- env API
    def set_motion(self, frames):
        f = np.ascontiguousarray(frames, dtype=np.float64)
        assert f.ndim == 2 and f.shape[1] == FRAME_DIM
        lib().orc_env_set_motion(self.env, _p(f), C.c_int(f.shape[0]))

    def set_reward_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.size == 24
        lib().orc_env_set_reward_weights(self.env, _p(w))

    def set_mode(self, mode):
        lib().orc_env_set_mode(self.env, C.c_int(1 if mode == "train" else 0))

    def reset(self):
        obs = np.zeros(OBS_DIM)
        lib().orc_env_reset(self.env, _p(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs, scal, info = np.zeros(OBS_DIM), np.zeros(3), np.zeros(9)
        flags = np.zeros(3, dtype=np.int32)
        lib().orc_env_step(self.env, _p(a), _p(obs), _p(scal), _p(info), _p(flags))
        return obs, float(scal[0]), bool(flags[0]), dict(fail=bool(flags[1]), end=bool(flags[2]), percent=float(scal[1]),
                                                          rfc_score=float(scal[2]), reward_info=info)

    def obs(self):
        o = np.zeros(OBS_DIM)
        lib().orc_env_obs(self.env, _p(o))
        return o

    def set_state(self, qpos, qvel):
        q = np.ascontiguousarray(qpos, dtype=np.float64)
        v = np.ascontiguousarray(qvel, dtype=np.float64)
        lib().orc_env_set_state(self.env, _p(q), _p(v))

    def substep(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        return int(lib().orc_env_substep(self.env, _p(a)))

    def mj_step(self):
        lib().orc_env_mj_step(self.env)

    def forward(self):
        lib().orc_env_forward(self.env)

    def set_ctrl_applied(self, ctrl, qfrc):
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        q = np.ascontiguousarray(qfrc, dtype=np.float64)
        lib().orc_set_ctrl_applied(self.env, _p(c), _p(q))

    def set_warmstart(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        lib().orc_set_warmstart(self.env, _p(w))

    def compute_torque(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        t = np.zeros(32)
        lib().orc_env_compute_torque(self.env, _p(a), _p(t))
        return t[:self.nu]

    @property
    def cur_t(self):
        return int(lib().orc_env_cur_t(self.env))

    @cur_t.setter
    def cur_t(self, t):
        lib().orc_env_set_cur_t(self.env, C.c_int(t))

    def clear_warnings(self):
        lib().orc_env_clear_warnings(self.env)

    def get(self, name, n=None):
        buf = np.zeros(4096 if name != "efc_J" else 500 * 32)
        k = getattr(lib(), "orc_get_" + name)(self.env, _p(buf))
        return buf[:k].copy()

    def qM(self):
        return self.get("qM").reshape(32, 32)[:self.nv, :self.nv]

    def efc_J(self):
        buf = np.zeros(500 * self.nv)
        k = lib().orc_get_efc_J(self.env, _p(buf))
        return buf[:k * self.nv].reshape(k, self.nv).copy()

    def contacts(self):
        buf = np.zeros(100 * 16)
        k = lib().orc_get_contacts(self.env, _p(buf))
        return buf[:16 * k].reshape(k, 16).copy()

    def counters(self):
        c = np.zeros(16, dtype=np.int32)
        lib().orc_get_counters(self.env, _p(c))
        keys = ["ncon", "nefc", "solver_iter", "warn_bad", "warn_contactfull", "warn_cnstrfull", "mpr_calls",
                "mpr_iters", "qp_n", "qp_iters", "n_avg_cp"]
        return dict(zip(keys, c.tolist()))

    def contact_forces(self):
        """mj_contactForce of every contact of the last forward: [ncon, 6]."""
        buf = np.zeros(100 * 6)
        k = lib().orc_env_ncontact_force(self.env, _p(buf))
        return buf[:6 * k].reshape(k, 6).copy()

    def set_record(self, on=True):
        lib().orc_env_set_record(self.env, C.c_int(1 if on else 0))

    def record(self):
        """(frames [sim_step, REC_STRIDE], surface [SURF_STRIDE]) of the last step (layout: include/hoic.h HOIC_REC_*)."""
        frames, surface = np.zeros(15 * REC_STRIDE), np.zeros(SURF_STRIDE)
        k = lib().orc_get_record(self.env, _p(frames), _p(surface))
        return frames[:k * REC_STRIDE].reshape(k, REC_STRIDE).copy(), surface

    def rfc(self):
        r = np.zeros(7)
        lib().orc_get_rfc(self.env, _p(r))
        return r

    def ho_diff(self):
        r = np.zeros(5)
        lib().orc_env_ho_diff(self.env, _p(r))
        return r


def preprocess_seq(blob: bytes, hand_pose_seq, obj_pose_seq, motion_freq=30.0, clip=True):
    """Oracle of hoic_preprocess_seq: (frames [T,212], contact_mask [T,19], contact_pos [T,19,3])."""
    L = lib()
    mdl = L.orc_model_create(blob, len(blob))
    hand = np.ascontiguousarray(hand_pose_seq, dtype=np.float64)
    obj = np.ascontiguousarray(obj_pose_seq, dtype=np.float64)
    T = hand.shape[0]
    frames, cmask, cpos = np.zeros((T, FRAME_DIM)), np.zeros((T, 24), dtype=np.int32), np.zeros((T, 24, 3))
    L.orc_preprocess_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.orc_preprocess_seq(C.c_void_p(mdl), _p(hand), _p(obj), T, C.c_double(motion_freq), 1 if clip else 0, _p(frames), _p(cmask), _p(cpos))
    L.orc_model_destroy(C.c_void_p(mdl))
    if rc != 0:
        raise ValueError("orc_preprocess_seq: need at least 2 frames")
    return frames, cmask[:, :19].copy(), cpos[:, :19].copy()


def solve_wrench_qp(contacts, target, center, mu=0.75, w_t=1e4, weights=None, penalties=None, surface_contact=True):
    """Oracle of hoic_solve_wrench_qp for ONE problem: contacts [nc, 12] -> (rest[6], score, point_force [nc * points, 3])."""
    L = lib()
    c = np.ascontiguousarray(contacts, dtype=np.float64).reshape(-1, 12)
    nc, ec = c.shape[0], 5 if surface_contact else 1
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    p = None if penalties is None else np.ascontiguousarray(penalties, dtype=np.float64)
    t, ce = np.ascontiguousarray(target, dtype=np.float64), np.ascontiguousarray(center, dtype=np.float64)
    rest, score, pf = np.zeros(6), np.zeros(1), np.zeros((max(nc * ec, 1), 3))
    L.orc_solve_wrench_qp.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_double] + [C.c_void_p] * 3
    L.orc_solve_wrench_qp(nc, ec, _p(c), None if w is None else _p(w), None if p is None else _p(p), _p(t), _p(ce),
                          C.c_double(mu), C.c_double(w_t), _p(rest), _p(score), _p(pf))
    return rest, float(score[0]), pf[:nc * ec]
