"""Look for non-finite numbers coming out of the rollout: python tools/debug_nan.py [obj] [n_envs] [iters] [n_seeds]"""
import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.agent import AgentHandMimic
obj = sys.argv[1] if len(sys.argv) > 1 else "banana"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nseed = int(sys.argv[4]) if len(sys.argv) > 4 else 8
mdl = model(obj)
for r in range(nseed):
    cfg = copy.deepcopy(mdl.cfg)
    cfg.min_batch_size = N * 4
    agent = AgentHandMimic(mdl, N, cfg=cfg, seed=cfg.seed + 1000 * r)
    env = agent.env
    orig_step = env.step
    def step(a, _o=orig_step, _e=env):
        obs, rew, done, info = _o(a)
        for name, t in (("obs", obs), ("reward", rew), ("info", info["reward_info"]), ("act", a)):
            badrows = (~torch.isfinite(t.reshape(t.shape[0], -1))).any(1).nonzero().flatten()
            if badrows.numel():
                i = int(badrows[0])
                q, v = _e.get_state()
                print(f"seed {r}: non-finite {name} in {badrows.numel()} envs, first env {i}: done={int(done[i])} fail={int(info['fail'][i])} "
                      f"reward={float(rew[i])} |q|max={float(q[i].abs().max()):.3g} |v|max={float(v[i].abs().max()):.3g} "
                      f"bad obs idx={(~torch.isfinite(obs[i])).nonzero().flatten()[:8].tolist()} counters={_e.counters()}")
        big = obs.abs().max()
        if big > 1e4:
            print(f"seed {r}: |obs| max {float(big):.3g}")
        return obs, rew, done, info
    env.step = step
    for i in range(iters):
        info = agent.optimize_policy(i)
        print(f"seed {r} iter {i}: value_loss {info['value_loss']:.4g} surrogate {info['surrogate']:.4g} episodes {info['log']['num_episodes']} "
              f"zf mean finite {bool(torch.isfinite(agent.running_state.mean).all())} std finite {bool(torch.isfinite(agent.running_state.std).all())}")
    agent.close()
