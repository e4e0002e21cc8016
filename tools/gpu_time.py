"""Kernel-only timing of hoic_step for A/B experiments: python tools/gpu_time.py [obj] [n_envs] [steps]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=1, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for _ in range(30):   # spread the envs over the episode
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
torch.cuda.synchronize()
c0 = env.counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acts = [torch.randn(N, 32, device="cuda", generator=g) * 0.1 for _ in range(K)]
e0.record()
for a in acts:
    env.step(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
c1 = env.counters()
d = {k: (c1[k] - c0[k]) / max(c1["substeps"] - c0["substeps"], 1) for k in ("ncon", "nefc", "newton_iters", "ls_evals", "mpr_calls", "mpr_iters")}
print(f"{os.environ.get('HOIC_LIB','default')} {obj} N={N}: {ms:.3f} ms/step  {N/ms*1e3:.0f} env-steps/s  per-substep {d}")
# same measurement without load balancing
env.balance = False
env.L.hoic_set_env_order(env._hb, None)
e0.record()
for a in acts:
    env.step(a)
e1.record(); torch.cuda.synchronize()
print(f"  without env ordering: {e0.elapsed_time(e1)/K:.3f} ms/step")
