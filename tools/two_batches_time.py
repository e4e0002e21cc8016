"""Does a previously created box batch change the speed of a banana batch (shared-memory carve-out)?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
first = sys.argv[1] if len(sys.argv) > 1 else "none"
keep = None
if first != "none":
    m0 = model(first); keep = HandMimicVecEnv(m0, 64, [m0.motion(1)], seed=1); keep.reset(); keep.step(torch.zeros(64, 32, device="cuda"))
mdl = model("banana"); N = 4096
env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=1, auto_reset=True); env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for _ in range(30): env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acts = [torch.randn(N, 32, device="cuda", generator=g) * 0.1 for _ in range(10)]
e0.record()
for a in acts: env.step(a)
e1.record(); torch.cuda.synchronize()
print(f"banana 4096 after creating '{first}' first: {e0.elapsed_time(e1) / 10:.3f} ms/step")
