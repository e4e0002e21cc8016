"""Generate tests/golden/<obj>.npz from the CPU oracle (float64).  The reference itself cannot run here (MuJoCo is an
absent third-party binary), so these vectors pin the ORACLE's behaviour (regression) and give the GPU
tests fixed inputs/outputs that do not depend on the oracle being rebuilt on the GPU box.
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import OBJECTS, model, oracle_env  # noqa: E402


def rollout_fixture(obj, n_steps=8, seed=3):
    env = oracle_env(obj)
    out = dict(reset_obs=env.reset())
    rng = np.random.RandomState(seed)
    acts, obs, rew, info, rfc, qpos, qvel, flags = [], [], [], [], [], [], [], []
    for t in range(n_steps):
        a = np.clip(rng.randn(32) * 0.2, -1.2, 1.2)
        o, r, d, i = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); info.append(i["reward_info"]); rfc.append(i["rfc_score"])
        qpos.append(env.get("qpos")); qvel.append(env.get("qvel")); flags.append([d, i["fail"], i["end"]])
        if d:
            break
    out.update(actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), info=np.array(info), rfc=np.array(rfc),
               qpos=np.array(qpos), qvel=np.array(qvel), flags=np.array(flags, dtype=np.int32))
    return out


def substep_fixture(obj, n_states=48, seed=0):
    env = oracle_env(obj)
    env.reset()
    rng = np.random.RandomState(seed)
    keys = ("q0", "v0", "w0", "ctrl", "qapp", "q1", "v1", "qacc", "fcon")
    S = {k: [] for k in keys}
    ncon, nefc, pairs, cdist = [], [], [], []
    steps = 0
    while len(S["q0"]) < n_states and steps < 400:
        a = np.clip(rng.randn(32) * 0.3, -1, 1)
        fail = 0
        for k in range(15):
            q0, v0, w0 = env.get("qpos"), env.get("qvel"), env.get("qacc_warmstart")
            fail = env.substep(a)
            if fail:
                break
            if k % 5 == 2 and len(S["q0"]) < n_states:
                c = env.counters()
                con = env.contacts()
                vals = dict(q0=q0, v0=v0, w0=w0, ctrl=env.get("ctrl"), qapp=env.get("qfrc_applied"), q1=env.get("qpos"),
                            v1=env.get("qvel"), qacc=env.get("qacc"), fcon=env.get("qfrc_constraint"))
                for kk in keys:
                    S[kk].append(vals[kk])
                ncon.append(c["ncon"]); nefc.append(c["nefc"])
                p = -np.ones((32, 2), dtype=np.int32); dd = np.zeros(32)
                p[:len(con)] = con[:, 13:15].astype(np.int32); dd[:len(con)] = con[:, 0]
                pairs.append(p); cdist.append(dd)
        env.cur_t = env.cur_t + 1
        steps += 1
        d = env.ho_diff()
        if fail or d[3] > 0.1 or d[0] > 0.1 or env.cur_t > 280:
            env.clear_warnings()
            env.reset()
    out = {"sub_" + k: np.array(v) for k, v in S.items()}
    out.update(sub_ncon=np.array(ncon, dtype=np.int32), sub_nefc=np.array(nefc, dtype=np.int32), sub_pairs=np.array(pairs),
               sub_cdist=np.array(cdist))
    return out


def extras_fixture():
    """Fixtures of the widened rows: the test-mode record of box steps in the grasp, random wrench QPs, a contact pre-pass."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    out = {}
    env = oracle_env("box", start_idx=100)
    env.set_mode("test"); env.set_record(True); env.reset()
    rng = np.random.RandomState(21)
    acts, frames, surf = [], [], []
    for t in range(4):
        a = np.clip(rng.randn(32) * 0.1, -1, 1)
        env.step(a)
        fr, su = env.record()
        acts.append(a); frames.append(fr); surf.append(su)
    out.update(rec_actions=np.array(acts), rec_frames=np.array(frames), rec_surface=np.array(surf))
    qc, qt, qce, qn, qrest, qscore = [], [], [], [], [], []
    for nc in (1, 2, 5, 19):
        c = np.zeros((19, 12)); center = rng.randn(3) * 0.1 + [0.4, 0.4, 0.6]
        for k in range(nc):
            n = rng.randn(3); n /= np.linalg.norm(n)
            t = np.cross(n, [1.0, 0, 0]) if abs(n[0]) < 0.9 else np.cross(n, [0, 1.0, 0]); t /= np.linalg.norm(t)
            c[k] = np.concatenate([center + rng.randn(3) * 0.03, n, t, np.cross(n, t)])
        target = np.concatenate([rng.randn(3) * 0.5 + [0, 0, 0.6], rng.randn(3) * 0.005])
        rest, score, _ = oracle.solve_wrench_qp(c[:nc], target, center, mu=0.75, w_t=1e4)
        qc.append(c); qt.append(target); qce.append(center); qn.append(nc); qrest.append(rest); qscore.append(score)
    out.update(qp_contacts=np.array(qc), qp_target=np.array(qt), qp_center=np.array(qce), qp_n=np.array(qn, dtype=np.int32),
               qp_rest=np.array(qrest), qp_score=np.array(qscore))
    mdl = model("bottle")
    ex = mdl.motion(2)
    fr, cm, cp = oracle.preprocess_seq(mdl.blob, ex["hand_dof_seq"][::10], ex["obj_pose_seq"][::10], 3.0)
    out.update(pre_frames=fr, pre_cmask=cm, pre_cpos=cp)
    return out


if __name__ == "__main__":
    d = extras_fixture()
    p = os.path.join(ROOT, "tests", "golden", "extras.npz")
    np.savez_compressed(p, **d)
    print("extras", {k: v.shape for k, v in d.items()}, os.path.getsize(p), "bytes")
    if "--extras-only" in sys.argv:
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for obj in OBJECTS:
        d = rollout_fixture(obj)
        d.update(substep_fixture(obj))
        p = os.path.join(ROOT, "tests", "golden", f"{obj}.npz")
        np.savez_compressed(p, **d)
        print(obj, {k: v.shape for k, v in d.items()}, os.path.getsize(p), "bytes")
