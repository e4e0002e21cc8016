"""Per-env work metric statistics (hoic_get_work: newton*8 + ncon*2 + mpr + qp term of the last step)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200 import capi
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = 4096
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], seed=1, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
W = []
for t in range(40):
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
    capi.check(env.L.hoic_get_work(env._hb, capi.ptr(env._work), env._stream()), "w")
    W.append(env._work.cpu().numpy().copy())
W = np.array(W[10:], dtype=float)
print("work mean %.1f std %.1f min %d max %d" % (W.mean(), W.std(), W.min(), W.max()))
print("percentiles", np.percentile(W, [5, 25, 50, 75, 95, 99]).round(0))
c = [np.corrcoef(W[i], W[i + 1])[0, 1] for i in range(len(W) - 1)]
print("step-to-step correlation of work per env: %.3f" % np.mean(c))
# loss of grouping by 7: sum of group max / sum of work, random order vs sorted by previous step vs oracle sort
def loss(order, w):
    g = w[order][: (N // 7) * 7].reshape(-1, 7)
    return g.max(1).sum() * 7 / g.sum()
r = np.random.RandomState(0)
print("group-of-7 max/mean: random %.3f  sorted-by-prev %.3f  sorted-by-self %.3f" % (
    np.mean([loss(r.permutation(N), W[i]) for i in range(1, len(W))]),
    np.mean([loss(np.argsort(W[i - 1], kind="stable"), W[i]) for i in range(1, len(W))]),
    np.mean([loss(np.argsort(W[i], kind="stable"), W[i]) for i in range(1, len(W))])))
