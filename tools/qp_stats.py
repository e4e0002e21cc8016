"""QP iteration statistics: python tools/qp_stats.py [obj] [n_envs]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=1, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for _ in range(30):
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
c0 = env.counters()
hist = torch.zeros(20, dtype=torch.long)
for _ in range(20):
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
    ncg = env.get_rfc()[:, 7].long().cpu()
    hist += torch.bincount(ncg.clamp(max=19), minlength=20)
c1 = env.counters()
steps = c1["env_steps"] - c0["env_steps"]
print(f"{obj}: qp newton iterations per env step {(c1['qp_iters'] - c0['qp_iters']) / steps:.2f}; contact-geom histogram (envs with k hand geoms in contact): {hist.tolist()}")
withqp = int(hist[1:].sum())
print(f"envs with a QP: {withqp / max(int(hist.sum()), 1):.2%}; iterations per QP {(c1['qp_iters'] - c0['qp_iters']) / max(withqp, 1):.2f}")
