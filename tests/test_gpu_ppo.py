"""GPU test of a complete PPO iteration (rollout -> GAE -> 5-epoch clipped-surrogate update) on the batched engine:
BASELINE config 4 shape at reduced size (uhc/agents/agent_handmimic.py:311-337, scripts/train_hand_mimic.py:77-80)."""
import copy

import numpy as np
import pytest
import torch

from common import model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("obj", ["box", "banana"])
def test_ppo_iteration_end_to_end(obj):
    from hoic_b200.agent import AgentHandMimic
    mdl = model(obj)
    cfg = copy.deepcopy(mdl.cfg)
    cfg.min_batch_size = 512 * 6
    agent = AgentHandMimic(mdl, 512, cfg=cfg)
    p0 = [p.detach().clone() for p in agent.policy_net.parameters()]
    v0 = [p.detach().clone() for p in agent.value_net.parameters()]
    infos = [agent.optimize_policy(i) for i in range(2)]
    for info in infos:
        assert info["samples"] == 512 * 6 and info["horizon"] == 6
        assert np.isfinite(info["value_loss"]) and np.isfinite(info["surrogate"])
        assert info["log"]["num_steps"] == 512 * 6 and 0.0 <= info["log"]["avg_reward"] <= 1.0
    assert infos[1]["value_loss"] < infos[0]["value_loss"] * 1.5
    assert any(not torch.equal(a, b) for a, b in zip(p0, agent.policy_net.parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(v0, agent.value_net.parameters()))
    assert all(torch.isfinite(p).all() for p in agent.policy_net.parameters())
    # the epoch-adaptive reward weights reached the kernel (handmimic_config.py:175-195)
    w1 = cfg.reward_vector().copy()
    agent.per_epoch_update(500)
    assert not np.allclose(w1, cfg.reward_vector())
    c = agent.env.counters()
    assert c["env_steps"] == 2 * 512 * 6 and c["fail_overflow"] == 0
    agent.close()


def test_noise_rate_masks_exploration():
    from hoic_b200.agent import AgentHandMimic
    mdl = model("box")
    cfg = copy.deepcopy(mdl.cfg)
    cfg.min_batch_size = 256 * 3
    agent = AgentHandMimic(mdl, 256, cfg=cfg)
    agent.per_epoch_update(0)
    agent.noise_rate = 0.5
    batch, _ = agent.sample(cfg.min_batch_size)
    frac = float(batch.exps.mean())
    assert 0.35 < frac < 0.65
    # non-exploratory samples carry the mean action
    mean, _ = agent.policy_net(batch.states[0])
    det = batch.exps[0] == 0
    assert torch.allclose(batch.actions[0][det], mean[det], atol=1e-6)
    info = agent.update_params(batch)
    assert np.isfinite(info["surrogate"])
    agent.close()
