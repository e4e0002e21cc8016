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
    assert c["env_steps"] == (2 * 6 + 2) * 512 and c["fail_overflow"] == 0     # + the 2 warm-up steps of the sampler graph
    # end_reward of the next iteration = avg_c_reward gamma / (1 - gamma) (agent_handmimic.py:317-319)
    assert abs(infos[1]["end_reward"] - infos[1]["log"]["avg_c_reward"] * cfg.gamma / (1 - cfg.gamma)) < 1e-9
    assert infos[1]["log"]["num_episodes"] >= 0 and len(agent.freq_dict) >= 8
    agent.close()


def test_noise_rate_masks_exploration():
    from hoic_b200.agent import AgentHandMimic
    mdl = model("box")
    cfg = copy.deepcopy(mdl.cfg)
    cfg.min_batch_size = 256 * 3
    agent = AgentHandMimic(mdl, 256, cfg=cfg, policy_dtype="fp32")
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


def test_eval_policy_batched_test_mode():
    """eval_policy (agent_handmimic.py:339-403) with every test motion in its own env: metrics of the reference, test mode
    (no early termination), optional MotionData recording; the oracle replays motion 0 with the same mean actions."""
    import sys, os
    from common import oracle_env
    from hoic_b200.agent import AgentHandMimic
    mdl = model("box")
    agent = AgentHandMimic(mdl, 64)
    agent.per_epoch_update(0)
    agent.sample(64 * 4)                                      # fills the observation normaliser
    motions = [mdl.motion(s) for s in (1, 2, 3)]
    m = agent.eval_policy(motions=motions, record=True, max_steps=12)
    per = m["per_motion"]
    assert set(("pose_err", "mpjpe", "avg_reward", "total_reward", "pose_reward", "obj_rfc_reward", "percent")) <= set(m)
    assert per["pose_err"].shape == (3,) and np.isfinite(per["mpjpe"]).all() and (per["avg_reward"] > 0).all()
    assert abs(m["percent"] - 12.0 / 299.0) < 1e-6           # 12 steps of a 300-frame motion, nobody terminated early
    assert len(m["motion_data"]) == 12 and m["motion_data"][0]["hand_kps"].shape == (3, 15, 21, 3)
    # oracle replay of motion 0: same policy (float64 copy), frozen normaliser, mean action
    orc = oracle_env("box", seed=1)
    orc.set_mode("test")
    orc.set_reward_weights(agent.cfg.reward_vector())
    obs = orc.reset()
    pol = copy.deepcopy(agent.policy_net).double().cpu()
    mean, std = agent.running_state.mean.double().cpu(), agent.running_state.std.double().cpu()
    pe = mp = rw = 0.0
    hb = np.asarray(mdl.arrays["env_hand_body_idx"]).astype(int)
    for t in range(12):
        if t >= 6:   # res["gt"][6:] - res["pred"][6:] slices the stacked [T, 26] arrays along time (agent_handmimic.py:386)
            pe += np.linalg.norm(motions[0]["hand_dof_seq"][t] - orc.get("qpos")[:26])
        mp += np.linalg.norm(motions[0]["body_pos_seq"][t] - orc.get("xpos").reshape(-1, 3)[hb], axis=-1).mean()
        x = torch.clamp((torch.as_tensor(obs) - mean) / (std + 1e-8), -5, 5)
        a = pol(x[None])[0][0].detach().numpy()
        obs, r, d, info = orc.step(a)
        rw += r
    assert abs(per["pose_err"][0] - pe / 6) < 1e-3 and abs(per["mpjpe"][0] - mp / 12) < 1e-4
    assert abs(per["avg_reward"][0] - rw / 12) < 2e-3
    agent.close()


def test_bf16_policy_matches_fp32_within_a_thousandth_of_sigma_and_graph_equals_eager():
    """Rollout inference runs on the tensor cores (bf16 operands, fp32 accumulate, fp32 master weights): the action mean must
    stay within 1e-3 sigma of the float32 network's; and the CUDA-graph sampler step must produce exactly what the eager
    step produces from the same state and RNG stream."""
    from hoic_b200 import ppo, dist as hdist
    from hoic_b200.sampler import SamplerStep, matmul_context
    from hoic_b200.vecenv import HandMimicVecEnv
    mdl = model("box")
    dev = torch.device("cuda", 0)
    pol, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    x = torch.clamp(torch.randn(2048, 617, device=dev), -5, 5)
    with torch.no_grad():
        ref = pol(x)[0]
        with matmul_context("bf16"):
            lo = pol(x)[0].float()
    sigma = float(torch.exp(pol.action_log_std).mean())
    assert float((lo - ref).abs().max()) < 1e-3 * sigma, float((lo - ref).abs().max()) / sigma
    outs = []
    for graph in (False, True):
        env = HandMimicVecEnv(mdl, 64, [mdl.motion(s) for s in range(1, 9)], device=0, seed=3, auto_reset=True)
        zf = ppo.ZFilter((617,), clip=5.0, device=dev)
        gen = torch.Generator(device=dev).manual_seed(11)
        st = hdist.RolloutStats(dev)
        s = SamplerStep(env, pol, zf, gen, stats=st, policy_dtype="bf16", graph=graph)
        env.reset()
        if graph:
            s.warmup()
        else:
            for _ in range(2):
                s.step()        # the graph's two eager warm-up steps
        st.reset()
        for _ in range(4):
            s.step()
        torch.cuda.synchronize()
        outs.append((s.act.clone(), env.reward.clone(), env.done.clone(), zf.mean.clone(), st.sums.clone()))
        env.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_library_normaliser_and_logger_match_the_reference_arithmetic():
    """hoic_zfilter / hoic_logger_step (device batches) against the float64 torch implementation that the CPU tests pin to
    zfilter.py / logger_rl.py: statistics, normalised output, since-last-sync accumulators, per-episode statistics."""
    from hoic_b200 import ppo, dist as hdist
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(2)
    gz, cz = ppo.ZFilter((617,), clip=5.0, device=dev), ppo.ZFilter((617,), clip=5.0, device="cpu")
    for it in range(4):
        x = (rng.randn(1000 + 37 * it, 617) * (1 + it) + 3 * it).astype(np.float32)
        x[:, 5] = 0.25                      # a constant observation entry
        x[:, 7] = 1000.0 + 1e-3 * x[:, 7]   # large mean, tiny spread
        yg = gz(torch.as_tensor(x, device=dev), update=it != 2)
        yc = cz(torch.as_tensor(x), update=it != 2)
        assert float(gz.n) == float(cz.n)
        assert torch.allclose(gz.mean.cpu(), cz.mean, rtol=1e-12, atol=1e-12) and torch.allclose(gz.m2.cpu(), cz.m2, rtol=1e-9, atol=1e-9)
        assert torch.allclose(gz._dmean.cpu(), cz._dmean, rtol=1e-12, atol=1e-12) and float(gz._dn) == float(cz._dn)
        assert torch.allclose(yg.cpu(), yc, rtol=1e-5, atol=2e-6)
    N = 300
    gs = hdist.RolloutStats(dev, episode_reward=torch.zeros(N, dtype=torch.float64, device=dev))
    cs = hdist.RolloutStats("cpu", episode_reward=torch.zeros(N, dtype=torch.float64))
    for t in range(12):
        r = rng.rand(N).astype(np.float32); info = rng.rand(N, 9).astype(np.float32); d = (rng.rand(N) < 0.2).astype(np.uint8)
        gs.update(torch.as_tensor(r, device=dev), torch.as_tensor(info, device=dev), torch.as_tensor(d, device=dev))
        cs.update(torch.as_tensor(r).double(), torch.as_tensor(info).double(), torch.as_tensor(d))
    a, b = gs.allreduce(), cs.allreduce()
    for k in ("num_steps", "num_episodes"):
        assert a[k] == b[k]
    for k in ("avg_c_reward", "avg_episode_len", "avg_episode_reward", "min_episode_reward", "max_episode_reward", "reward_min", "reward_max"):
        assert abs(a[k] - b[k]) < 1e-9, k
    assert np.allclose(a["info_sum"], b["info_sum"]) and torch.allclose(gs.episode_reward.cpu(), cs.episode_reward)


def test_sampler_fast_path_equals_the_autocast_path_and_host_step_returns_the_transition():
    """The sampler step's bf16 fast path (hoic_zfilter16 -> MLP on cached bfloat16 weights -> hoic_sample_action) against
    the torch.autocast path it replaces, from the same observations / noise / parameters; the weight copies follow an
    in-place parameter update; step_host (early D2H copies on a side stream) returns the device transition."""
    from hoic_b200 import ppo
    from hoic_b200.sampler import SamplerStep
    from hoic_b200.vecenv import HandMimicVecEnv
    mdl = model("box")
    dev = torch.device("cuda", 0)
    pol, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    acts, xs = [], []
    for fast in (True, False):
        env = HandMimicVecEnv(mdl, 96, [mdl.motion(s) for s in range(1, 9)], device=0, seed=3, auto_reset=True)
        zf = ppo.ZFilter((617,), clip=5.0, device=dev)
        gen = torch.Generator(device=dev).manual_seed(11)
        s = SamplerStep(env, pol, zf, gen, policy_dtype="bf16", graph=False, host_noise=True)
        assert s._fast
        s._fast = fast
        env.reset()
        s.noise.copy_(torch.randn(96, 32, generator=torch.Generator().manual_seed(5)).to(dev))
        with torch.no_grad():
            (s._act_fast if fast else s._act_torch)()
        torch.cuda.synchronize()
        acts.append(s.act.clone()); xs.append(s.x.clone())
        if fast:
            assert torch.equal(s.x16[:, :617].float(), s.x.to(torch.bfloat16).float()) and float(s.x16[:, 617:].abs().max()) == 0.0
            # parameters change in place -> the bf16 copies follow at the next step
            with torch.no_grad():
                pol.action_mean.bias.add_(0.25)
            s._refresh_weights()
            assert torch.equal(s._w16[-1].float(), pol.action_mean.bias.to(torch.bfloat16).float())
            with torch.no_grad():
                pol.action_mean.bias.sub_(0.25)
            s._refresh_weights()
        env.close()
    assert torch.equal(xs[0], xs[1])
    sigma = float(torch.exp(pol.action_log_std).mean())
    assert float((acts[0] - acts[1]).abs().max()) < 2e-3 * sigma, float((acts[0] - acts[1]).abs().max())
    # host-resident step: pinned buffers hold exactly the device transition
    env = HandMimicVecEnv(mdl, 96, [mdl.motion(s) for s in range(1, 9)], device=0, seed=3, auto_reset=True)
    zf = ppo.ZFilter((617,), clip=5.0, device=dev)
    s = SamplerStep(env, pol, zf, torch.Generator(device=dev).manual_seed(11), policy_dtype="bf16", graph=True, host_noise=True)
    env.reset()
    noise = s.host_noise_buffers()
    for b in noise:
        b.normal_()
    for k in range(3):
        s.step_host_begin(noise[k & 1])
        noise[(k + 1) & 1].normal_()
        out = s.step_host_end()
        assert torch.equal(out["state"], s.x.cpu()) and torch.equal(out["action"], s.act.cpu()) and torch.equal(out["exp"], s.exp.cpu())
        assert torch.equal(out["reward"], env.reward.cpu()) and torch.equal(out["done"], env.done.cpu())
        assert torch.equal(out["reward_info"], env.reward_info.cpu())
        # (noise[(k + 1) & 1] was redrawn while step k ran; noise[k & 1] is what step k used)
    assert torch.equal(s.noise.cpu(), noise[0])                 # step 2 read staging buffer 0
    ext = torch.randn(96, 32).pin_memory()                      # a caller's own pinned buffer goes through staging buffer 0
    s.step_host(ext)
    assert torch.equal(s.noise.cpu(), ext)
    env.close()


def test_sampler_graph_is_recaptured_when_the_env_changes_what_the_graph_bakes_in():
    """set_mode / set_record / set_motions change kernel arguments (and may reallocate tables) that a captured sampler graph
    holds by value: the sampler must drop its graph (HandMimicVecEnv.graph_epoch) instead of replaying stale arguments.
    Train mode fails an env whose hand drifts from the expert (body_fail), test mode does not: after set_mode("test") the
    replayed step must follow the test-mode rule."""
    from hoic_b200 import ppo
    from hoic_b200.sampler import SamplerStep
    from hoic_b200.vecenv import HandMimicVecEnv
    mdl = model("box")
    dev = torch.device("cuda", 0)
    pol, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    with torch.no_grad():
        pol.action_log_std.fill_(1.5)            # wild actions: body_fail fires within a few steps in train mode
    env = HandMimicVecEnv(mdl, 64, [mdl.motion(s) for s in range(1, 9)], device=0, seed=3, auto_reset=True)
    s = SamplerStep(env, pol, ppo.ZFilter((617,), clip=5.0, device=dev), torch.Generator(device=dev).manual_seed(1), policy_dtype="bf16", graph=True)
    env.reset()
    fails = 0
    for _ in range(12):
        s.step()
        fails += int((env.fail & ~env.end.bool()).sum())
    assert fails > 0
    g0 = s._graph
    env.set_mode("test")
    env.set_record(True)
    env.reset()
    fails = 0
    for _ in range(12):
        s.step()
        fails += int(env.fail.sum())
    assert s._graph is not g0
    assert fails == 0 or int(env.counters()["fail_physics"]) + int(env.counters()["fail_overflow"]) > 0
    assert float(env._rec_frames.abs().sum()) > 0          # the recapture wrote into the new recording buffers
    env.close()
