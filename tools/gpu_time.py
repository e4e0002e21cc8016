"""Kernel-only timing of hoic_step for A/B experiments: python tools/gpu_time.py [obj] [n_envs] [steps] [graph]
`graph`: the step (library call incl. auto-reset) is captured once into a CUDA graph and replayed."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
use_graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=1, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for _ in range(30):   # spread the envs over the episode
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
torch.cuda.synchronize()
act = torch.zeros(N, 32, device="cuda")
graph = None
if use_graph:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        env.step(act)
    torch.cuda.synchronize()
c0 = env.counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acts = [torch.randn(N, 32, device="cuda", generator=g) * 0.1 for _ in range(K)]
e0.record()
for a in acts:
    if graph is not None:
        act.copy_(a)
        graph.replay()
    else:
        env.step(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
c1 = env.counters()
d = {k: round((c1[k] - c0[k]) / max(c1["substeps"] - c0["substeps"], 1), 2) for k in ("ncon", "nefc", "newton_iters", "ls_evals", "mpr_calls", "mpr_iters")}
print(f"{obj} N={N} parts={os.environ.get('HOIC_STEP_PARTS', 'default')} graph={use_graph}: {ms:.3f} ms/step  {N/ms*1e3:.0f} env-steps/s  {d}", flush=True)
