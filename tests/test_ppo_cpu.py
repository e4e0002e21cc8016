"""CPU tests of the PPO host logic against literal transcriptions of the reference's pure-Python code:
GAE (uhc/khrylib/rl/core/common.py:5-25), ZFilter / RunningStat (uhc/khrylib/utils/zfilter.py:7-73), the clipped
surrogate step (uhc/khrylib/rl/agents/agent_ppo.py:16-64) and the network layout (models/mlp.py, policy_gaussian.py)."""
import numpy as np
import pytest
import torch

from common import model
from hoic_b200 import ppo


def ref_estimate_advantages(rewards, masks, values, gamma, tau):
    """common.py:5-25 on one concatenated [n, 1] batch (float64)."""
    n = rewards.shape[0]
    deltas, adv = np.zeros((n, 1)), np.zeros((n, 1))
    prev_value = prev_adv = 0.0
    for i in reversed(range(n)):
        deltas[i] = rewards[i] + gamma * prev_value * masks[i] - values[i]
        adv[i] = deltas[i] + gamma * tau * prev_adv * masks[i]
        prev_value, prev_adv = values[i, 0], adv[i, 0]
    returns = values + adv
    a = torch.as_tensor(adv)
    return ((a - a.mean()) / a.std()).numpy(), returns


def test_gae_matches_reference_on_complete_episodes():
    rng = np.random.RandomState(0)
    T, N = 40, 6
    rewards, values = rng.rand(T, N), rng.randn(T + 1, N)
    masks = (rng.rand(T, N) > 0.15).astype(float)
    masks[-1] = 0.0      # the reference only stores complete episodes: every column ends with mask 0
    adv, ret = ppo.estimate_advantages(torch.as_tensor(rewards), torch.as_tensor(masks), torch.as_tensor(values), 0.95, 0.95)
    # reference layout: episodes concatenated env after env
    r = rewards.T.reshape(-1, 1); m = masks.T.reshape(-1, 1); v = values[:-1].T.reshape(-1, 1)
    radv, rret = ref_estimate_advantages(r, m, v, 0.95, 0.95)
    assert np.allclose(adv.numpy().T.reshape(-1, 1), radv, atol=1e-12)
    assert np.allclose(ret.numpy().T.reshape(-1, 1), rret, atol=1e-12)


def test_gae_bootstraps_truncated_tails():
    # an episode still running at the horizon uses V(s_T); with tau = 1 the return is the discounted sum + gamma^k V
    T = 5
    rewards = torch.ones(T, 1, dtype=torch.float64)
    masks = torch.ones(T, 1, dtype=torch.float64)
    values = torch.zeros(T + 1, 1, dtype=torch.float64)
    values[-1] = 10.0
    _, ret = ppo.estimate_advantages(rewards, masks, values, 0.9, 1.0)
    expect = sum(0.9 ** k for k in range(T)) + 0.9 ** T * 10.0
    assert abs(float(ret[0, 0]) - expect) < 1e-12


def test_zfilter_batched_equals_sequential_welford():
    from_ref_n, M, S = 0, np.zeros(5), np.zeros(5)
    rng = np.random.RandomState(1)
    xs = rng.randn(7, 4, 5) * 3 + 1
    zf = ppo.ZFilter((5,), clip=5.0, dtype=torch.float64)
    for b in xs:
        for x in b:             # RunningStat.push, zfilter.py:14-24
            from_ref_n += 1
            if from_ref_n == 1:
                M[...] = x
            else:
                old = M.copy()
                M[...] = old + (x - old) / from_ref_n
                S[...] = S + (x - old) * (x - M)
        y = zf(torch.as_tensor(b))
    std = np.sqrt(S / (from_ref_n - 1))
    assert np.allclose(zf.mean.numpy(), M, atol=1e-12) and np.allclose(zf.std.numpy(), std, atol=1e-12)
    expect = np.clip((xs[-1] - M) / (std + 1e-8), -5, 5)
    assert np.allclose(y.numpy(), expect, atol=1e-12)
    # update=False leaves the statistics alone
    n = float(zf.n)
    zf(torch.as_tensor(xs[0]), update=False)
    assert float(zf.n) == n


def test_network_layout_and_state_dict_keys():
    cfg = model("box").cfg
    pol, val = ppo.make_policy_value(cfg, device="cpu", dtype=torch.float64, seed=1)
    assert [l.out_features for l in pol.net.affine_layers] == [2048, 1024, 512] and pol.action_mean.out_features == 32
    keys = set(pol.state_dict().keys())
    assert {"net.affine_layers.0.weight", "action_mean.weight", "action_mean.bias", "action_log_std"} <= keys
    assert {"net.affine_layers.2.bias", "value_head.weight"} <= set(val.state_dict().keys())
    assert torch.allclose(pol.action_log_std, torch.tensor(cfg.log_std, dtype=torch.float64), atol=1e-6) and torch.all(pol.action_mean.bias == 0)
    x = torch.randn(3, 617, dtype=torch.float64)
    a = pol.select_action(x, mean_action=True)
    lp = pol.get_log_prob(x, a)     # log-density of the mean: -sum(log_std) - d/2 log(2 pi)
    assert np.allclose(lp.detach().numpy(), -32 * cfg.log_std - 16 * np.log(2 * np.pi), atol=1e-4)


def test_ppo_update_is_the_reference_step():
    """One epoch of ppo_update == value MSE Adam step then clipped-surrogate Adam step with grad-norm clip (agent_ppo.py:44-64)."""
    torch.manual_seed(0)
    mk = lambda: (ppo.PolicyGaussian(ppo.MLP(8, (16, 8), "gelu"), 3, log_std=-1.0).double(), ppo.Value(ppo.MLP(8, (16, 8), "gelu")).double())
    pol, val = mk()
    pol2, val2 = mk()
    pol2.load_state_dict(pol.state_dict()); val2.load_state_dict(val.state_dict())
    n = 64
    s, a = torch.randn(n, 8, dtype=torch.float64), torch.randn(n, 3, dtype=torch.float64) * 0.3
    ret, adv = torch.randn(n, 1, dtype=torch.float64), torch.randn(n, 1, dtype=torch.float64)
    exps = (torch.rand(n) < 0.8).double()
    ind = exps.nonzero().squeeze(1)
    with torch.no_grad():
        flp = pol.get_log_prob(s, a)
    op, ov = torch.optim.Adam(pol.parameters(), lr=5e-3), torch.optim.Adam(val.parameters(), lr=3e-3)
    st = ppo.ppo_update(pol, val, op, ov, s, a, ret, adv, flp, clip_epsilon=0.2, num_epoch=1, grad_clip=0.05, ind=ind)
    # transcription
    op2, ov2 = torch.optim.Adam(pol2.parameters(), lr=5e-3), torch.optim.Adam(val2.parameters(), lr=3e-3)
    vl = (val2(s) - ret).pow(2).mean(); ov2.zero_grad(); vl.backward(); ov2.step()
    lp = pol2.get_log_prob(s[ind], a[ind]); ratio = torch.exp(lp - flp[ind])
    sl = -torch.min(ratio * adv[ind], torch.clamp(ratio, 0.8, 1.2) * adv[ind]).mean()
    op2.zero_grad(); sl.backward(); torch.nn.utils.clip_grad_norm_(pol2.parameters(), 0.05); op2.step()
    for p, q in zip(list(pol.parameters()) + list(val.parameters()), list(pol2.parameters()) + list(val2.parameters())):
        assert torch.allclose(p, q, atol=1e-14)
    assert abs(st["value_loss"] - float(vl)) < 1e-12 and abs(st["surrogate"] - float(sl)) < 1e-12
    # more epochs reduce the value loss
    st5 = ppo.ppo_update(pol, val, op, ov, s, a, ret, adv, flp, num_epoch=20, ind=ind)
    assert st5["value_loss"] < st["value_loss"]


def test_checkpoint_roundtrip_in_the_reference_format(tmp_path):
    """iter_XXXX.p layout of agent_handmimic.py:175-186: state dicts + a pickled uhc.khrylib.utils.zfilter.ZFilter whose
    RunningStat carries (_n, _M, _S); readable without the reference through the name-patching unpickler (tools.py:7-18)."""
    import pickle
    import pickletools
    from common import model
    cfg = model("box").cfg
    pol, val = ppo.make_policy_value(cfg, device="cpu", dtype=torch.float32, seed=3)
    zf = ppo.ZFilter((617,), clip=5.0)
    x = torch.randn(50, 617, dtype=torch.float64) * 3 + 1
    zf(x)
    fn = str(tmp_path / "iter_0001.p")
    ppo.save_checkpoint(fn, pol, val, zf)
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(open(fn, "rb").read())]
    globals_ = {arg for name, arg in ops if name == "GLOBAL"}
    assert "uhc.khrylib.utils.zfilter ZFilter" in globals_ and "uhc.khrylib.utils.zfilter RunningStat" in globals_
    assert not any("hoic_b200" in str(arg) for _, arg in ops)          # nothing of this repo is needed to open the file
    with pytest.raises(ModuleNotFoundError):
        pickle.load(open(fn, "rb"))                                      # plain pickle needs the reference's module ...
    pol2, val2 = ppo.make_policy_value(cfg, device="cpu", dtype=torch.float32, seed=4)
    zf2 = ppo.ZFilter((617,), clip=10.0)
    cp = ppo.load_checkpoint(fn, pol2, val2, zf2)                        # ... the patched unpickler does not
    assert set(cp) == {"policy_dict", "value_dict", "running_state"}
    for a, b in zip(list(pol.parameters()) + list(val.parameters()), list(pol2.parameters()) + list(val2.parameters())):
        assert torch.equal(a, b)
    assert float(zf2.n) == 50 and torch.allclose(zf2.mean, zf.mean) and torch.allclose(zf2.std, zf.std) and zf2.clip == 5.0
    # the reference's RunningStat.var = _S / (_n - 1): same statistic as ours
    rs = cp["running_state"].rs
    assert np.allclose(rs._S / (rs._n - 1), x.var(0, unbiased=True).numpy())
    y = torch.randn(4, 617, dtype=torch.float64)
    assert torch.equal(zf(y, update=False), zf2(y, update=False))


def test_zfilter_merge_of_eight_float32_ranks_stays_finite():
    """The pooled second moment is combined from centred moments in float64.  The earlier raw-moment form
    (sum of squares - n mean^2 in float32) went negative on near-constant observation entries once 8 ranks were merged
    and turned the normaliser's std, and with it every network input, into NaN."""
    rng = np.random.RandomState(0)
    x = rng.randn(8, 96, 5).astype(np.float32)
    x[..., 0] = 0.8660254
    x[..., 1] = 0.5 + 1e-6 * x[..., 1]
    x[..., 2] = 0.0
    x[..., 3] = 3.0 + 1e-4 * x[..., 3]
    parts = []
    for r in range(8):
        zf = ppo.ZFilter((5,), dtype=torch.float32)
        for chunk in np.array_split(x[r], 4):
            zf(torch.as_tensor(chunk))
        parts.append((zf.n, zf.mean, zf.m2))
    n, mean, m2 = ppo.ZFilter.merge_moments(parts)
    std = torch.sqrt(m2 / (n - 1))
    ref = torch.as_tensor(x.reshape(-1, 5), dtype=torch.float64)
    assert float(n) == 768 and torch.isfinite(std).all() and (m2 >= 0).all()
    assert torch.allclose(mean, ref.mean(0), atol=1e-6) and torch.allclose(std[3:], ref.std(0)[3:], rtol=2e-3)
    assert float(std[0]) < 1e-6 and float(std[2]) == 0.0 and abs(float(std[1]) - float(ref.std(0)[1])) < 1e-6
    # accumulators are float64 whatever dtype the observations have (zfilter.py:7-24 is numpy float64)
    assert parts[0][1].dtype == torch.float64 and parts[0][0].dtype == torch.float64


def test_gae_terminal_transition_isolates_non_finite_tail():
    """A NaN / inf value or reward BEHIND an episode boundary must not leak across it (mask 0 selects zero, it does not
    multiply): one blown-up env may not poison the batch (common.py:5-25 semantics otherwise unchanged)."""
    T, N = 6, 3
    rng = np.random.RandomState(3)
    rewards = torch.as_tensor(rng.rand(T, N))
    values = torch.as_tensor(rng.rand(T + 1, N))
    masks = torch.ones(T, N, dtype=torch.float64)
    masks[2, 1] = 0.0                       # env 1: episode ends at t = 2, the next one is garbage from t = 3 on
    clean_adv, clean_ret = ppo.estimate_advantages(rewards, masks, values, 0.95, 0.95)
    values2 = values.clone()
    values2[3, 1] = float("nan")            # V(s_3) of the NEXT episode
    adv, ret = ppo.estimate_advantages(rewards, masks, values2, 0.95, 0.95)
    # the transitions of env 1 up to its terminal step see none of it (returns are un-normalised: exact)
    assert torch.isfinite(ret[:3, 1]).all() and torch.allclose(ret[:3, 1], clean_ret[:3, 1])
    assert torch.isfinite(ret[:, 0]).all() and torch.isfinite(ret[:, 2]).all()


def test_end_reward_and_logger_statistics_follow_the_reference():
    """`reward += env.end_reward` on info["end"] with env.end_reward = avg_c_reward gamma / (1 - gamma) of the PREVIOUS
    iteration (agent_handmimic.py:317-319, 478-480), and LoggerRL's per-episode statistics (logger_rl.py:28-67), against a
    transcription of the reference's sequential loop on a scripted environment."""
    from hoic_b200 import dist as hd
    rng = np.random.RandomState(5)
    T, N, gamma = 40, 4, 0.95
    rew = rng.rand(T, N)
    done = rng.rand(T, N) < 0.15
    end = done & (rng.rand(T, N) < 0.5)
    info = rng.rand(T, N, 9)
    end_reward = 0.7 * gamma / (1 - gamma)
    # --- reference: per env, sequential episodes
    ref_train = np.zeros((T, N))
    n_steps = n_eps = 0
    total_c = total_ep = 0.0
    mn_ep, mx_ep = np.inf, -np.inf
    for e in range(N):
        ep = 0.0
        for t in range(T):
            c_reward = rew[t, e]
            reward = c_reward + (end_reward if end[t, e] else 0.0)      # :478-480
            ep += c_reward                                               # LoggerRL.step sees env_reward = c_reward here
            total_c += c_reward
            n_steps += 1
            ref_train[t, e] = reward
            if done[t, e]:
                n_eps += 1
                total_ep += ep
                mn_ep, mx_ep = min(mn_ep, ep), max(mx_ep, ep)
                ep = 0.0
    # --- batched
    ep_acc = torch.zeros(N, dtype=torch.float64)
    st = hd.RolloutStats("cpu", episode_reward=ep_acc)
    train = np.zeros((T, N))
    for t in range(T):
        r = torch.as_tensor(rew[t])
        train[t] = (r + end_reward * torch.as_tensor(end[t]).to(r.dtype)).numpy()
        st.update(r, torch.as_tensor(info[t]), torch.as_tensor(done[t]), torch.zeros(N, dtype=torch.int32),
                  torch.zeros(N, dtype=torch.int32), torch.as_tensor(rng.rand(N)))
    log = st.allreduce()
    assert np.allclose(train, ref_train)
    assert log["num_steps"] == n_steps and log["num_episodes"] == n_eps
    assert abs(log["avg_c_reward"] - total_c / n_steps) < 1e-12 and abs(log["avg_episode_len"] - n_steps / n_eps) < 1e-12
    assert abs(log["avg_episode_reward"] - total_ep / n_eps) < 1e-12
    assert abs(log["min_episode_reward"] - mn_ep) < 1e-12 and abs(log["max_episode_reward"] - mx_ep) < 1e-12
    assert st.freq_rows().shape == (n_eps, 3)
    # next iteration's end reward
    assert abs(log["avg_c_reward"] * gamma / (1 - gamma) - (total_c / n_steps) * gamma / (1 - gamma)) < 1e-12
