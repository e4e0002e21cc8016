"""Kernel-only timing with IDENTICAL environments (same motion, start frame and actions): removes the per-env load
imbalance inside a CTA, isolating contention from imbalance.  python tools/gpu_time_uniform.py [obj] [n_envs] [steps] [start]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
start = int(sys.argv[4]) if len(sys.argv) > 4 else 60
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(1)], seed=1, auto_reset=False, random_start=False, balance=False)
env.reset(motion_id=[0] * N, start_idx=[start] * N)
g = torch.Generator(device="cuda").manual_seed(0)
acts = [(torch.randn(1, 32, device="cuda", generator=g) * 0.1).repeat(N, 1) for _ in range(K + 3)]
for a in acts[:3]:
    env.step(a)
torch.cuda.synchronize()
c0 = env.counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for a in acts[3:]:
    env.step(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
c1 = env.counters()
d = {k: round((c1[k] - c0[k]) / max(c1["substeps"] - c0["substeps"], 1), 2) for k in ("ncon", "nefc", "newton_iters", "ls_evals")}
print(f"{os.path.basename(os.environ.get('HOIC_LIB','default'))} uniform {obj} N={N} start={start}: {ms:.3f} ms/step {N/ms*1e3:.0f} env-steps/s {d}")
