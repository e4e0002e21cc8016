"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding and rollout-statistics reduction."""
import os
import subprocess
import sys

import numpy as np

from common import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from hoic_b200 import dist as hd
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
sh = hd.shard(10, rank, 2)
assert sh.tolist() == list(range(rank, 10, 2)), sh
mid = hd.mixed_model_ids(10, 3, rank, 2)
assert mid.tolist() == [(i % 3) for i in range(rank, 10, 2)]
# rollout statistics: sums / min / max / moments must equal the single-process values
rng = np.random.RandomState(0)
rew = rng.rand(10); info = rng.rand(10, 9); done = rng.rand(10) < 0.4
st = hd.RolloutStats(device="cpu")
st.update(torch.as_tensor(rew[sh]), torch.as_tensor(info[sh]), torch.as_tensor(done[sh]))
tot = st.allreduce()
assert abs(tot["reward_sum"] - rew.sum()) < 1e-12 and tot["num_steps"] == 10 and tot["num_episodes"] == int(done.sum())
assert abs(tot["reward_min"] - rew.min()) < 1e-12 and abs(tot["reward_max"] - rew.max()) < 1e-12
assert np.allclose(tot["info_sum"], info.sum(0))
# global advantage normalisation from sharded moments (uhc/khrylib/rl/core/common.py:22)
adv = rng.randn(10)
mean, std = hd.global_mean_std(torch.as_tensor(adv[sh]))
assert abs(mean - adv.mean()) < 1e-12 and abs(std - torch.as_tensor(adv).std().item()) < 1e-12
# gradient averaging keeps replicas identical
p = torch.nn.Linear(4, 2)
torch.manual_seed(rank); 
for q in p.parameters(): q.grad = torch.randn_like(q)
ref = [q.grad.clone() for q in p.parameters()]
hd.allreduce_grads(p.parameters())
g = [torch.zeros_like(r) for r in ref]
for q, r in zip(p.parameters(), ref):
    both = [torch.zeros_like(r), torch.zeros_like(r)]
    dist.all_gather(both, r)
    assert torch.allclose(q.grad, (both[0] + both[1]) / 2)
# observation normaliser merged across ranks (float32, as the agent runs it): a constant entry, a large-mean / tiny-spread
# entry and ordinary ones must come out finite and equal to the single-process statistics
from hoic_b200 import ppo
x = rng.randn(4000, 6).astype(np.float32)
x[:, 0] = 1.0
x[:, 1] = 1000.0 + 1e-3 * x[:, 1]
x[:, 2] = 0.0
zf = ppo.ZFilter((6,), clip=5.0, dtype=torch.float32)
for chunk in np.array_split(x[rank::2], 5):
    zf(torch.as_tensor(chunk))
zf.allreduce()
one = ppo.ZFilter((6,), clip=5.0, dtype=torch.float64)
one(torch.as_tensor(x, dtype=torch.float64))
assert float(zf.n) == 4000 and torch.isfinite(zf.std).all() and (zf.m2 >= 0).all()
assert torch.allclose(zf.mean.double(), one.mean, rtol=1e-6, atol=1e-6)
assert torch.allclose(zf.std.double()[3:], one.std[3:], rtol=1e-4) and float(zf.std[0]) < 1e-3 and float(zf.std[2]) == 0.0
assert abs(float(zf.std[1]) - float(one.std[1])) < 1e-3
y = zf(torch.as_tensor(x[:8]), update=False)
assert torch.isfinite(y).all()
# repeated rollouts: update / allreduce every iteration (agent.sample).  Each rank must end up with the statistics of
# ALL samples seen so far counted ONCE (pooling full per-rank statistics again would count the shared history `world`
# times per iteration and freeze the normaliser).
zr = ppo.ZFilter((6,), clip=5.0)
single = ppo.ZFilter((6,), clip=5.0)
seen = 0
for it in range(5):
    xi = (rng.randn(300, 6) * (1 + it) + it).astype(np.float32)
    for chunk in np.array_split(xi[rank::2], 3):
        zr(torch.as_tensor(chunk))
    zr.allreduce()
    single(torch.as_tensor(xi))
    seen += 300
    assert float(zr.n) == seen, (float(zr.n), seen)
    assert torch.allclose(zr.mean, single.mean, rtol=1e-10, atol=1e-10) and torch.allclose(zr.std, single.std, rtol=1e-10)
# flat-buffer gradient averaging survives a second backward pass (the grads stay views of the flat buffer)
lin = torch.nn.Linear(3, 2)
torch.manual_seed(7)
for q in lin.parameters(): torch.nn.init.constant_(q, 0.5)
for rep in range(2):
    lin.zero_grad(set_to_none=False)
    (lin(torch.full((1, 3), float(rank + 1 + rep))).sum()).backward()
    hd.allreduce_grads(lin.parameters())
    assert torch.allclose(lin.weight.grad, torch.full((2, 3), 1.5 + rep)), lin.weight.grad
dist.destroy_process_group()
print("OK", rank)
'''


def test_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 500)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o
