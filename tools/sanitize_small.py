"""Tiny workload for compute-sanitizer (memcheck / racecheck): python tools/sanitize_small.py [obj] [n_envs] [steps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import model
from hoic_b200.vecenv import HandMimicVecEnv
obj = sys.argv[1] if len(sys.argv) > 1 else "box"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mdl = model(obj)
env = HandMimicVecEnv(mdl, N, [mdl.motion(1)], seed=1, auto_reset=True, random_start=False)
env.reset(motion_id=[0] * N, start_idx=[100] * N)
if os.environ.get("HOIC_SAN_RECORD", "1") == "1":
    env.set_record(True)
if os.environ.get("HOIC_SAN_SAMPLER") == "1":
    # the sampler step's library kernels (hoic_zfilter16, hoic_sample_action, hoic_logger_step) around hoic_step, eagerly
    from hoic_b200 import ppo, dist as hdist
    from hoic_b200.sampler import SamplerStep
    dev = torch.device("cuda", 0)
    pol, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    smp = SamplerStep(env, pol, ppo.ZFilter((617,), clip=5.0, device=dev), torch.Generator(device=dev).manual_seed(1),
                      stats=hdist.RolloutStats(dev), policy_dtype="bf16", graph=False)
    for _ in range(K):
        smp.step()
    r = env.reward
for _ in range(0 if os.environ.get("HOIC_SAN_SAMPLER") == "1" else K):
    o, r, d, i = env.step(torch.zeros(N, 32, device="cuda"))
torch.cuda.synchronize()
print(obj, "ok", float(r.mean()), env.counters()["ncon"])
