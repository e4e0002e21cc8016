"""Generate tests/golden/<obj>_states.npz: single-mj_step parity vectors at SURVEY 8(d)'s size from the CPU oracle (float64):
>= 1024 states sampled from noisy-policy rollouts that START in frames 60..210 of the synthetic motions (approach end,
grasp, lift: the contact-rich part of the task) + 256 hand-crafted contact-rich states (expert grasp frames with the
fingers closed further and the object pushed a few millimetres into them).  Per state: the inputs of one mj_step
(qpos, qvel, ctrl, qfrc_applied, qacc_warmstart) and its outputs (qpos, qvel, contact geom pairs, ncon, nefc).
The reference's MuJoCo cannot run here or on the GPU box (profiles/r02_probe_mujoco_gpu_box.txt), so these pin the
ORACLE (regression) and give the GPU tests fixed vectors; parity stays "unpinned" in the sense of DESIGN.md section 2.
    python tools/make_golden_states.py [obj ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import OBJECTS, model, oracle_env  # noqa: E402

MAXC = 48            # contact-pair slots per state in the fixture
GPU_NCON, GPU_NBASIS, GPU_NEDGE = 32, 80, 128


class Rec:
    keys = ("q0", "v0", "w0", "ctrl", "qapp", "q1", "v1")

    def __init__(self):
        self.S = {k: [] for k in self.keys}
        self.ncon, self.nefc, self.pairs, self.kind = [], [], [], []

    def add(self, env, q0, v0, w0, kind):
        con = env.contacts()
        c = env.counters()
        if len(con) > GPU_NCON:
            return False
        dims = con[:, 15].astype(int) if con.shape[1] > 15 else None
        vals = dict(q0=q0, v0=v0, w0=w0, ctrl=env.get("ctrl"), qapp=env.get("qfrc_applied"), q1=env.get("qpos"), v1=env.get("qvel"))
        if not all(np.isfinite(v).all() for v in vals.values()):
            return False
        for k in self.keys:
            self.S[k].append(np.array(vals[k], dtype=np.float64))
        p = -np.ones((MAXC, 2), dtype=np.int16)
        p[:len(con)] = con[:, 13:15].astype(np.int16)
        self.pairs.append(p); self.ncon.append(c["ncon"]); self.nefc.append(c["nefc"]); self.kind.append(kind)
        return True

    def out(self):
        d = {"sub_" + k: np.array(v) for k, v in self.S.items()}
        d.update(sub_ncon=np.array(self.ncon, dtype=np.int16), sub_nefc=np.array(self.nefc, dtype=np.int16),
                 sub_pairs=np.array(self.pairs), sub_kind=np.array(self.kind, dtype=np.int8))
        return d


def rollout_states(obj, rec, n_states, seed=0):
    """States every 7th substep of rollouts under N(0, std) actions, std in {0.1, 0.2, 0.3}, starting in frames 60..210."""
    rng = np.random.RandomState(seed)
    sub = 0
    while len(rec.ncon) < n_states:
        env = oracle_env(obj, seed=int(rng.randint(1, 9)), start_idx=int(rng.randint(60, 211)))
        env.reset()
        std = (0.1, 0.2, 0.3)[rng.randint(3)]
        for step in range(int(rng.randint(4, 12))):
            a = np.clip(rng.randn(32) * std, -1, 1)
            fail = 0
            for k in range(15):
                q0, v0, w0 = env.get("qpos"), env.get("qvel"), env.get("qacc_warmstart")
                fail = env.substep(a)
                if fail:
                    break
                sub += 1
                if sub % 7 == 0 and len(rec.ncon) < n_states:
                    rec.add(env, q0, v0, w0, 0)
            env.cur_t = env.cur_t + 1
            d = env.ho_diff()
            if fail or d[3] > 0.1 or d[0] > 0.1 or env.cur_t > 80:
                break
        env.close()


def crafted_states(obj, rec, n_states, seed=1):
    """Expert frames of the grasp / lift with the fingers closed further, the object pushed towards the palm by a few mm and
    random joint torques: deep multi-contact states (many hand-object contacts on top of the object-table ones)."""
    rng = np.random.RandomState(seed)
    mdl = model(obj)
    lo, hi = np.asarray(mdl.arrays["jnt_range"]).reshape(-1, 2)[:26].T
    tl = np.asarray(mdl.arrays["env_torque_lim"])
    tries = 0
    env = oracle_env(obj, seed=1, start_idx=0)
    target = len(rec.ncon) + n_states
    while len(rec.ncon) < target and tries < 40 * n_states:
        tries += 1
        ex = mdl.motion(int(rng.randint(1, 9)))
        t = int(rng.randint(95, 260))
        q = np.zeros(33); v = np.zeros(32)
        hand = np.array(ex["hand_dof_seq"][t], dtype=np.float64)
        hand[6:] += rng.uniform(0.0, 0.35, 20)                       # close the fingers further
        hand[:3] += rng.randn(3) * 0.002
        hand[3:6] += rng.randn(3) * 0.02
        q[:26] = np.clip(hand, lo, hi)
        pose = np.array(ex["obj_pose_seq"][t], dtype=np.float64)
        pose[:3] += rng.randn(3) * 0.003
        dq = np.concatenate([[1.0], rng.randn(3) * 0.03]); dq /= np.linalg.norm(dq)
        w0, x0, y0, z0 = pose[3:]; w1, x1, y1, z1 = dq
        pose[3:] = [w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                    w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1]
        q[26:] = pose
        v[:26] = np.array(ex["hand_dof_vel_seq"][t]) + rng.randn(26) * 0.05
        v[26:29] = np.array(ex["obj_vel_seq"][t]) + rng.randn(3) * 0.01
        v[29:] = np.array(ex["obj_angle_vel_seq"][t]) + rng.randn(3) * 0.05
        env.clear_warnings()
        env.set_state(q, v)
        ncon_now = len(env.contacts())
        if ncon_now < 10 and rng.rand() < 0.9:                       # keep the set contact-rich
            continue
        ctrl = rng.uniform(-0.5, 0.5, 26) * tl
        qapp = np.zeros(32); qapp[2] = 0.635 * 9.8; qapp[26:29] = rng.randn(3) * 0.5
        w = rng.randn(32) * (0.0 if rng.rand() < 0.5 else 5.0)
        env.set_ctrl_applied(ctrl, qapp)
        env.set_warmstart(w)
        q0, v0 = env.get("qpos"), env.get("qvel")
        env.mj_step()
        c = env.counters()
        if c["warn_bad"] or c["warn_contactfull"] or c["warn_cnstrfull"]:
            continue
        if c["nefc"] - 32 > 32 + GPU_NEDGE:      # (limit rows + edges must fit the GPU's row capacity)
            continue
        rec.add(env, q0, v0, w, 1)
    env.close()


if __name__ == "__main__":
    objs = sys.argv[1:] or OBJECTS
    for obj in objs:
        rec = Rec()
        rollout_states(obj, rec, 1024)
        crafted_states(obj, rec, 256)
        d = rec.out()
        p = os.path.join(ROOT, "tests", "golden", f"{obj}_states.npz")
        np.savez_compressed(p, **d)
        nc, ne = d["sub_ncon"], d["sub_nefc"]
        print(obj, len(nc), "states; ncon hist", np.bincount(nc.clip(0, 40))[:34].tolist(), "max nefc", int(ne.max()),
              "nefc>=90:", int((ne >= 90).sum()), "ncon>=12:", int((nc >= 12).sum()), os.path.getsize(p), "bytes")
