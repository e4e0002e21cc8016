"""Cycles of the Newton solve (k_p3s / k_p3) by iteration count, measured in situ during hoic_step (analysis build:
tools/build_variant.sh stats -DH_SOLVE_STATS; HOIC_LIB=variants/lib_stats.so python tools/solve_stats.py [obj] [n_envs])."""
import ctypes as C, os, sys
import numpy as np, torch
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
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
for name in ("hoic_debug_solve_stats_small", "hoic_debug_solve_stats_full"):
    getattr(env.L, name)(buf, 1)
for _ in range(20):
    env.step(torch.randn(N, 32, device="cuda", generator=g) * 0.1)
torch.cuda.synchronize()
for name in ("hoic_debug_solve_stats_small", "hoic_debug_solve_stats_full"):
    getattr(env.L, name)(buf, 0)
    a = np.array(list(buf), dtype=np.float64)
    cnt, cyc, ls = a[:16], a[16:32], a[32:48]
    print(name, "solves", int(cnt.sum()))
    for it in range(16):
        if cnt[it]:
            print(f"  iters {it:2d}: {100 * cnt[it] / cnt.sum():6.2f} %  mean {cyc[it] / cnt[it] / 1.965e3:8.1f} us  ({cyc[it] / cnt[it]:9.0f} cycles)  ls evals {ls[it] / cnt[it]:5.1f}")
    tot_it = float((cnt * np.arange(16)).sum())
    names = ["init", "gradient", "weights", "direction (assembly + Cholesky)", "line search + move", "  of direction: Hessian assembly", "  of line search: mul_M + mul_J"]
    for i, nm in enumerate(names):
        print(f"  {nm:36s} {a[48 + i] / max(cnt.sum(), 1):9.0f} cycles per solve   {a[48 + i] / max(tot_it, 1):9.0f} per iteration")
