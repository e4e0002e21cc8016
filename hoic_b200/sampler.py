"""One step of the sampler's inner loop for the whole batch (uhc/agents/agent_handmimic.py:460-490):
observation normaliser -> policy MLP -> exploration noise -> env.step (+ custom reward, auto-reset) -> logger,
captured once into ONE CUDA graph and replayed per step.

The MLP runs on the tensor cores: bfloat16 operands with float32 accumulation under `torch.autocast`, float32 master
weights (`policy_dtype="tf32"` / `"fp32"` select TF32 tensor cores / the plain float32 pipe; the reference runs float64
on the CPU).  Everything else of the step is libhoic_b200.so (`HandMimicVecEnv.step` = one `hoic_step` call).
"""
from __future__ import annotations

import contextlib

import torch

from . import capi


def matmul_context(policy_dtype):
    """Context in which the policy / value MLPs run: bf16 autocast, TF32 tensor cores, or strict float32."""
    if policy_dtype == "bf16":
        return torch.autocast(device_type="cuda", dtype=torch.bfloat16)
    if policy_dtype == "tf32":
        return _tf32(True)
    if policy_dtype == "fp32":
        return _tf32(False)
    raise ValueError(policy_dtype)


@contextlib.contextmanager
def _tf32(on):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = on
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


class SamplerStep:
    """`step()` advances every env by one policy step and leaves the transition in fixed device buffers:
    `x` (normalised state the action was computed from), `act`, `exp`, `reward` (c_reward), `done`, `end`, `fail`,
    `percent`, `reward_info`, `motion_id` / `start_idx` (the episode the transition belongs to), `obs` (next state, raw;
    the reset observation where the episode ended).  `stats` (dist.RolloutStats) and `running_state` (ppo.ZFilter) are
    updated in place inside the graph."""

    def __init__(self, env, policy, running_state, gen, stats=None, noise_rate=1.0, policy_dtype="bf16", graph=True,
                 update_filter=True, host_noise=False):
        self.env, self.policy, self.zf, self.gen, self.stats = env, policy, running_state, gen, stats
        self.policy_dtype, self.update_filter = policy_dtype, update_filter
        N, dev = env.n_envs, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.noise_rate = torch.full((), float(noise_rate), **f32)      # device scalar: changing it needs no recapture
        self.x = torch.zeros(N, env.obs_dim, **f32)
        self.act = torch.zeros(N, env.act_dim, **f32)
        self.exp = torch.ones(N, **f32)
        self.motion_id = torch.zeros(N, dtype=torch.int32, device=dev)
        self.start_idx = torch.zeros(N, dtype=torch.int32, device=dev)
        self.noise = torch.zeros(N, env.act_dim, **f32) if host_noise else None   # filled by the caller (step_host)
        self.obs = env.obs
        self._graph = None
        self._use_graph = graph
        self._host = None
        self._warm_done = self._host_warm_done = False
        # bf16 fast path: the normaliser writes a bfloat16 copy of the state (padded to the first layer's GEMM width) next to
        # the float32 one, the MLP runs on bfloat16 copies of the weights that are refreshed only when the parameters change
        # (torch.autocast re-casts every weight on every call: 9 cast / pad kernels per step), and the draw of the action is
        # one library kernel.  Same arithmetic as the autocast path (bf16 operands, fp32 accumulation, bf16 activations).
        self._fast = policy_dtype == "bf16" and hasattr(policy, "net") and hasattr(policy.net, "affine_layers") and env.device.type == "cuda"
        if self._fast:
            in_dim = policy.net.affine_layers[0].in_features
            self._ld16 = in_dim + ((-in_dim) % 64 if in_dim % 8 else 0)
            self.x16 = torch.zeros(N, self._ld16, dtype=torch.bfloat16, device=dev)
            self._u = torch.zeros(N, **f32)
            self._rnd = torch.zeros(N, env.act_dim, **f32)
            self._w16, self._w16_version = None, None
            self._refresh_weights()

    # ------------------------------------------------------------------ bf16 copies of the policy's parameters
    def _params(self):
        ps = []
        for a in self.policy.net.affine_layers:
            ps += [a.weight, a.bias]
        return ps + [self.policy.action_mean.weight, self.policy.action_mean.bias]

    def _refresh_weights(self):
        """In-place refresh of the bfloat16 weight copies when a parameter changed (optimizer step, checkpoint load): the
        buffers keep their addresses, so a captured graph keeps reading the right memory."""
        ps = self._params()
        version = tuple((p.data_ptr(), p._version) for p in ps)
        if version == self._w16_version:
            return
        with torch.no_grad():
            if self._w16 is None:
                self._w16 = [torch.zeros(p.shape[0], self._ld16, dtype=torch.bfloat16, device=p.device) if i == 0
                             else torch.empty_like(p, dtype=torch.bfloat16) for i, p in enumerate(ps)]
            for i, (w, p) in enumerate(zip(self._w16, ps)):
                if i == 0:
                    w[:, :p.shape[1]].copy_(p)
                else:
                    w.copy_(p)
        self._w16_version = version

    def set_noise_rate(self, r):
        self.noise_rate.fill_(float(r))

    # ------------------------------------------------------------------ the step body (eager or under capture)
    def _act_fast(self):
        env = self.env
        self.zf(env.obs, update=self.update_filter, out=self.x, out16=self.x16)
        h = self.x16
        w = self._w16
        for i in range(len(self.policy.net.affine_layers)):
            h = self.policy.net.activation(torch.nn.functional.linear(h, w[2 * i], w[2 * i + 1]))
        mean = torch.nn.functional.linear(h, w[-2], w[-1])
        if self.noise is None:
            noise = self._rnd.normal_(generator=self.gen)
        else:
            noise = self.noise
        self._u.uniform_(generator=self.gen)
        capi.check(env.L.hoic_sample_action(capi.ptr(mean), 1, capi.ptr(self.policy.action_log_std), capi.ptr(noise), capi.ptr(self._u),
                                            capi.ptr(self.noise_rate), mean.shape[0], mean.shape[1], capi.ptr(self.act), capi.ptr(self.exp),
                                            env._stream()), "hoic_sample_action")

    def _act_torch(self):
        env = self.env
        x = self.zf(env.obs, update=self.update_filter)
        with matmul_context(self.policy_dtype):
            mean, log_std = self.policy(x)
        mean, log_std = mean.float(), log_std.float()
        if self.noise is None:
            noise = torch.randn(mean.shape, device=mean.device, dtype=torch.float32, generator=self.gen)
        else:
            noise = self.noise
        # mean_action = binomial(1, 1 - noise_rate) per (env, step); exp = 1 - mean_action (agent_handmimic.py:464-465, :484)
        u = torch.rand(mean.shape[0], device=mean.device, generator=self.gen)
        self.exp.copy_((u < self.noise_rate).to(torch.float32))
        self.x.copy_(x)
        self.act.copy_(mean + torch.exp(log_std) * noise * self.exp[:, None])

    def compute_action(self):
        """The first half of the step alone, eagerly: normalise `env.obs`, run the policy, draw `self.act` (bench.py times
        `hoic_step` by itself on the actions the sampler would feed it)."""
        if self._fast:
            self._refresh_weights()
        with torch.no_grad():
            if self._fast:
                self._act_fast()
            else:
                self._act_torch()

    def _body(self, after_action=None):
        env = self.env
        with torch.no_grad():
            if self._fast:
                self._act_fast()
            else:
                self._act_torch()
        if after_action is not None:
            after_action()      # step_host: the state / action of the transition start their way to the host here
        capi.check(env.L.hoic_get_episode(env._hb, capi.ptr(self.motion_id), capi.ptr(self.start_idx), None, env._stream()),
                   "hoic_get_episode")
        obs, r, done, info = env.step(self.act)
        if self.stats is not None:
            self.stats.update(r, info["reward_info"], done)

    def warmup(self):
        """Two eager steps (cuBLAS handles, autocast caches, lazily loaded kernels), then the capture.  The eager steps are
        ordinary rollout steps whose transitions are dropped: call this before the rollouts that count (and reset `stats`)."""
        if not self._use_graph or self._graph is not None:
            return
        dev = self.env.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(0 if self._warm_done else 2):     # a recapture (env.graph_epoch changed) needs no warm-up steps
                self._body()
        self._warm_done = True
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        with torch.cuda.graph(g, stream=side):
            self._body()
        torch.cuda.synchronize(dev)
        self._graph = g

    def _check_epoch(self):
        """The env's mode / recording buffers / motion tables are kernel arguments baked into a captured graph: drop the
        graphs when the env says one of them changed (HandMimicVecEnv.graph_epoch) and capture again on the next step."""
        ep = getattr(self.env, "graph_epoch", 0)
        if ep != getattr(self, "_epoch", ep):
            self._graph = None
            if self._host is not None:
                self._host["graph"] = None
        self._epoch = ep

    def step(self):
        self._check_epoch()
        if self._fast:
            self._refresh_weights()
        if self._use_graph:
            if self._graph is None:
                self.warmup()
            self._graph.replay()
        else:
            self._body()

    # ------------------------------------------------------------------ host-resident rollout storage (Memory.push on the host)
    def step_host(self, noise_host):
        """The same step for a sampler whose RNG and rollout storage live on the HOST, like the reference's
        (agent_handmimic.py:463-485, khrylib/rl/core/memory.py): `noise_host` [N, 32] pinned float32 standard-normal draws
        go H2D, the transition (normalised state, action, exp, reward, mask, info, percent) comes back D2H into pinned
        buffers.  One graph launch + one stream synchronise per step.  Returns the dict of pinned host tensors."""
        self.step_host_begin(noise_host)
        return self.step_host_end()

    def step_host_begin(self, noise_host):
        """Launch the host-resident step and return at once: the caller may draw the next step's noise (into ANOTHER pinned
        buffer) while the device works; `step_host_end()` waits for the transition."""
        if self.noise is None:
            raise capi.HoicError("construct SamplerStep(host_noise=True) for step_host")
        self._check_epoch()
        if self._fast:
            self._refresh_weights()
        if self._host is None or self._host["graph"] is None:
            self._build_host()
        h = self._host
        ptrs = [b.data_ptr() for b in h["noise"]]
        if noise_host.data_ptr() in ptrs:
            idx = ptrs.index(noise_host.data_ptr())      # one of the sampler's own staging buffers: no host copy
        else:
            idx = 0
            h["noise"][0].copy_(noise_host)
        h["graph"][idx].replay()

    def host_noise_buffers(self):
        """The two pinned [N, act_dim] staging buffers of the host-resident step (each has its own captured graph): a caller
        that draws its noise straight into them, alternating, saves the host-side copy of `step_host_begin`."""
        if self.noise is None:
            raise capi.HoicError("construct SamplerStep(host_noise=True) for step_host")
        if self._host is None or self._host["graph"] is None:
            self._check_epoch()
            self._build_host()
        return self._host["noise"]

    def step_host_end(self):
        torch.cuda.current_stream(self.env.device).synchronize()
        return self._host["out"]

    def _host_body(self, idx=0):
        h = self._host
        env = self.env
        dev = env.device
        self.noise.copy_(h["noise"][idx], non_blocking=True)
        main = torch.cuda.current_stream(dev)

        def early():
            # the normalised state (95 % of the transition's bytes) and the action are final before the physics starts:
            # their D2H copies run on a side stream underneath hoic_step
            h["ev0"].record(main)
            h["side"].wait_event(h["ev0"])
            with torch.cuda.stream(h["side"]):
                h["out"]["state"].copy_(self.x, non_blocking=True)
                h["out"]["action"].copy_(self.act, non_blocking=True)
                h["out"]["exp"].copy_(self.exp, non_blocking=True)
                h["ev1"].record(h["side"])

        self._body(after_action=early)
        for k, src in (("reward", env.reward), ("done", env.done), ("end", env.end), ("fail", env.fail), ("percent", env.percent),
                       ("reward_info", env.reward_info)):
            h["out"][k].copy_(src, non_blocking=True)
        main.wait_event(h["ev1"])

    def _build_host(self):
        env, N = self.env, self.env.n_envs
        pin = lambda *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype).pin_memory()
        out = dict(state=pin(N, env.obs_dim), action=pin(N, env.act_dim), exp=pin(N), reward=pin(N), percent=pin(N),
                   reward_info=pin(N, capi.INFO_DIM), done=pin(N, dtype=torch.uint8), end=pin(N, dtype=torch.uint8),
                   fail=pin(N, dtype=torch.uint8))
        keep = self._host["noise"] if self._host is not None else None      # a recapture keeps the buffers callers may hold
        self._host = dict(noise=keep or [pin(N, env.act_dim), pin(N, env.act_dim)], out=out, graph=None,
                          side=torch.cuda.Stream(env.device), ev0=torch.cuda.Event(), ev1=torch.cuda.Event())
        if keep is None:
            for b in self._host["noise"]:
                b.zero_()
        self.h2d_bytes = N * env.act_dim * 4
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in out.values())
        dev = env.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(0 if self._host_warm_done else 2):
                self._host_body()
        self._host_warm_done = True
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graphs = []
        for idx in range(2):       # one graph per staging buffer (the H2D copy node reads a fixed host address)
            g = torch.cuda.CUDAGraph()
            g.register_generator_state(self.gen)
            with torch.cuda.graph(g, stream=side):
                self._host_body(idx)
            torch.cuda.synchronize(dev)
            graphs.append(g)
        self._host["graph"] = graphs
