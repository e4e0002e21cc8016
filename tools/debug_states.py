"""Worst states of the large single-mj_step parity sets: python tools/debug_states.py [obj]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
G = np.load(os.path.join(ROOT, "tests", "golden", f"{obj}_states.npz"))
n = len(G["sub_q0"])
mdl = model(obj)
env = HandMimicVecEnv(mdl, n, [mdl.motion(1)], auto_reset=False, random_start=False)
env.reset(motion_id=[0] * n, start_idx=[0] * n)
env.set_state(G["sub_q0"], G["sub_v0"]); env.set_warmstart(G["sub_w0"])
qacc, fcon, stats = env.substep_debug(G["sub_ctrl"], G["sub_qapp"])
q1, v1 = [t.cpu().numpy() for t in env.get_state()]
stats = stats.cpu().numpy()
ev = np.abs(v1 - G["sub_v1"])
rv = ev / (1e-4 * np.abs(G["sub_v1"]) + 1e-5)
dv = np.abs(G["sub_v1"] - G["sub_v0"])
worst = rv.max(1)
print("fraction within tol", (worst <= 1).mean(), "ncon mismatch", int((stats[:, 1] != G["sub_ncon"]).sum()))
for i in np.argsort(-worst)[:12]:
    j = rv[i].argmax()
    print(f"state {i} kind {G['sub_kind'][i]} ncon {G['sub_ncon'][i]}/{stats[i,1]} nefc {G['sub_nefc'][i]}/{stats[i,0]} iters {stats[i,2]} flags {stats[i,3]} "
          f"worst ratio {worst[i]:.1f} at dof {j}: err {ev[i,j]:.2e} v1 {G['sub_v1'][i,j]:.3e} |dv| dof {dv[i,j]:.2e} max|v1| {np.abs(G['sub_v1'][i]).max():.2e} max|dv| {dv[i].max():.2e} "
          f"err/max|dv| {ev[i].max()/dv[i].max():.1e}")
# norm-wise metric
nw = ev.max(1) / (1e-4 * np.maximum(np.abs(G["sub_v1"]).max(1), dv.max(1)) + 1e-5)
print("norm-wise (err_inf <= 1e-4 max(|v1|_inf, |dv|_inf) + 1e-5): within", (nw <= 1).mean(), "worst", np.sort(nw)[-5:])
