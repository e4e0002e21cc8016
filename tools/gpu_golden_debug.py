import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model, oracle_env
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1]
G = np.load(os.path.join(ROOT, "tests", "golden", f"{obj}.npz"))
mdl = model(obj)
env = HandMimicVecEnv(mdl, 2, [mdl.motion(1)], auto_reset=False, random_start=False)
orc = oracle_env(obj); orc.reset()
env.reset(motion_id=[0, 0], start_idx=[0, 0])
for t in range(len(G["actions"])):
    a = G["actions"][t]
    oo, orw, od, oi = orc.step(a)
    obs, rew, done, info = env.step(torch.as_tensor(np.tile(a, (2, 1)), dtype=torch.float32))
    q, v = [x.cpu().numpy() for x in env.get_state()]
    e = np.abs(obs.cpu().numpy()[0] - G["obs"][t])
    print(t, "obs err", e.max(), "idx", e.argmax(), "| vs live oracle", np.abs(obs.cpu().numpy()[0] - oo).max(), "qpos err", np.abs(q[0] - G["qpos"][t]).max(), np.abs(q[0] - G["qpos"][t]).argmax(),
          "qvel err", np.abs(v[0] - G["qvel"][t]).max(), np.abs(v[0] - G["qvel"][t]).argmax(), "rfc", oi["rfc_score"], float(env.get_rfc()[0, 0]), "ncg", float(env.get_rfc()[0, 7]), orc.counters()["n_avg_cp"], "ncon", orc.counters()["ncon"])
print(env.counters())
