"""Where the host-buffer step (HandMimicVecEnv.step_host) spends its time: python tools/e2e_breakdown.py [obj] [n_envs] [steps]
Alternates device-tensor steps and host-buffer steps on the same env batch so both see the same mix of states."""
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
for _ in range(30):
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
h_act = (torch.randn(N, 32) * 0.1).pin_memory()
d_act = h_act.cuda()
for _ in range(3):
    env.step_host(h_act)
torch.cuda.synchronize()
t_dev = t_host = t_gpu_host = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(K):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    env.step(d_act); torch.cuda.synchronize()
    t_dev += time.perf_counter() - t0
    t0 = time.perf_counter(); e0.record()
    env.step_host(h_act)
    e1.record(); torch.cuda.synchronize()
    t_host += time.perf_counter() - t0
    t_gpu_host += e0.elapsed_time(e1) * 1e-3
print(f"{obj} N={N}: device-tensor step {t_dev/K*1e3:.3f} ms (wall, synced) | step_host {t_host/K*1e3:.3f} ms wall, {t_gpu_host/K*1e3:.3f} ms on the stream"
      f" | h2d {env.h2d_bytes/1e6:.2f} MB d2h {env.d2h_bytes/1e6:.2f} MB")
