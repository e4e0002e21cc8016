import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from common import model, oracle_env
import oracle
from hoic_b200.metrics import PhysMetrics
np.set_printoptions(precision=5, suppress=True, linewidth=200)
mdl = model("box"); ex = mdl.motion(1)
qpos = np.concatenate([ex["hand_dof_seq"], ex["obj_pose_seq"]], 1)
pm = PhysMetrics(mdl, qpos)
stable, rest_ft = [x.cpu().numpy() for x in pm.eval_stable(return_rest=True)]
force, torque = [x.cpu().numpy() for x in pm.obtain_target_ft()]
orc = oracle_env("box"); orc.reset()
m = mdl.arrays["body_mass"][-1]
bad = 0
for t in range(0, 300):
    orc.set_state(qpos[t], np.zeros(32))
    c = orc.contacts()
    ho = c[(c[:, 13] >= 2) & (c[:, 13] <= 20) & (c[:, 14] >= 21)]
    _, o_score, _ = oracle.solve_wrench_qp(ho[:, 1:13], np.concatenate([force[t], torque[t]]), qpos[t, 26:29], mu=1.0, w_t=1.0)
    if abs(rest_ft[t] - o_score / m) > 5e-3 * max(1.0, rest_ft[t]):
        bad += 1
        if bad <= 2:
            mask = pm.ho_mask[t].cpu().numpy()
            gp = pm.contact_pos[t].cpu().numpy()[mask]; gf = pm.contact_frame[t].cpu().numpy()[mask]
            print("frame", t, "gpu", rest_ft[t], "oracle", o_score / m, "n", len(ho), mask.sum())
            print(" oracle geoms", ho[:, 13], "\n oracle pos\n", ho[:, 1:4], "\n gpu pos\n", gp, "\n oracle normal\n", ho[:, 4:7], "\n gpu normal\n", gf[:, :3])
            # oracle QP on the GPU's contacts
            _, s2, _ = oracle.solve_wrench_qp(np.concatenate([gp, gf], 1), np.concatenate([force[t], torque[t]]), qpos[t, 26:29], mu=1.0, w_t=1.0)
            print(" oracle QP on gpu contacts:", s2 / m)
print("mismatching frames", bad, "of 300")
