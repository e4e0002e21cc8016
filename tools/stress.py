"""Robustness sweep: large random actions, every object; reports failures, capacity overflows and non-finite outputs.
python tools/stress.py [n_envs] [steps] [action_std]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
std = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
for obj in ("box", "bottle", "banana"):
    mdl = model(obj)
    env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=3, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    nonfinite = 0
    maxcon = 0
    eps = 0
    for t in range(K):
        obs, r, d, info = env.step(torch.randn(N, 32, device="cuda", generator=g) * std)
        nonfinite += int((~torch.isfinite(obs)).any(1).sum()) + int((~torch.isfinite(r)).sum()) + int((~torch.isfinite(info["reward_info"])).any(1).sum())
        eps += int(d.sum())
        if t % 10 == 0:
            maxcon = max(maxcon, int(env.get_contacts()[0].max()))
    c = env.counters()
    print(f"{obj}: {N}x{K} steps, action std {std}: episodes ended {eps}, fail_physics {c['fail_physics']}, fail_overflow {c['fail_overflow']}, "
          f"non-finite outputs {nonfinite}, max contacts seen {maxcon}, contacts/substep {c['ncon'] / max(c['substeps'], 1):.2f}, "
          f"newton/substep {c['newton_iters'] / max(c['substeps'], 1):.2f}, high-water marks: contacts {c['max_ncon']} basis rows {c['max_nbasis']} edge rows {c['max_nedge']}")
    env.close()
