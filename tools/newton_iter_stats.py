"""Per-substep Newton iteration counts of the oracle under the bench's sampler (random-init policy, auto-reset): how the long
solves are distributed over envs and substeps (do the same envs stay slow through a step?).  Analysis tool, CPU only.

  python tools/newton_iter_stats.py [obj] [envs] [steps]
"""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def worker(args):
    obj, seed, steps = args
    import torch
    torch.set_num_threads(1)
    from common import model, oracle_env
    from oracle import lib
    from hoic_b200 import ppo
    mdl = model(obj)
    rng = np.random.default_rng(seed)
    env = oracle_env(obj, seed=1 + seed % 8, start_idx=int(rng.integers(0, 200)))
    policy, _ = ppo.make_policy_value(mdl.cfg, device="cpu", dtype=torch.float64, seed=1)
    zf = ppo.ZFilter((617,), clip=5.0, device="cpu")
    gen = torch.Generator().manual_seed(seed)
    obs = env.reset()
    rows = []
    it, nc = np.zeros(32, dtype=np.int32), np.zeros(32, dtype=np.int32)
    for _ in range(steps):
        x = zf(torch.as_tensor(obs)[None])
        act = policy.select_action(x, generator=gen)[0].numpy()
        obs, r, done, info = env.step(act)
        n = lib().orc_get_substep_log(env.env, it.ctypes.data_as(C.c_void_p), nc.ctypes.data_as(C.c_void_p))
        rows.append((it[:n].copy(), nc[:n].copy()))
        if done:
            env.clear_warnings()
            obs = env.reset()
    return rows


if __name__ == "__main__":
    obj = sys.argv[1] if len(sys.argv) > 1 else "box"
    envs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(worker, [(obj, 100 + i, steps) for i in range(envs)])
    iters = np.array([[r[0] for r in env_rows] for env_rows in res])     # [env, step, substep]
    ncon = np.array([[r[1] for r in env_rows] for env_rows in res])
    np.save("/tmp/newton_iters.npy", iters)
    np.save("/tmp/newton_ncon.npy", ncon)
    print("iters histogram", np.bincount(iters.ravel())[:20] / iters.size)
    print("mean iters", iters.mean(), "mean ncon", ncon.mean())
    # per step (axis 1): sum over substeps of the max over envs vs max over envs of the sum over substeps
    som = iters.max(axis=0).sum(axis=1)          # [step]
    mos = iters.sum(axis=2).max(axis=0)          # [step]
    avg = iters.sum(axis=2).mean(axis=0)
    print("per step: sum_sub max_env", som.mean(), " max_env sum_sub", mos.mean(), " mean_env sum_sub", avg.mean())
    print("p99 / p999 of single solves:", np.percentile(iters, 99), np.percentile(iters, 99.9), "max", iters.max())
