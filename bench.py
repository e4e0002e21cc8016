"""bench.py -- HandMimic rollout throughput (env-steps/s), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--obj box] [--no-extras]

Our arm, one step = one pass of the sampler's inner loop over the batch (uhc/agents/agent_handmimic.py:460-490):
observation normaliser -> policy MLP (PyTorch, bf16 tensor cores) -> hoic_step (15 physics substeps + control + contact
bookkeeping + RFC QP + termination + reward + observation + auto-reset of finished episodes: the kernel pipeline of
libhoic_b200.so) -> rollout statistics, replayed as ONE CUDA graph (hoic_b200/sampler.py).
Workload at N=1 is BASELINE config 2: box_future5_light_add_geom, 4096 batched envs per GPU (weak scaling: 4096 envs on
every rank).  `value`: transition buffers stay on the device.  `e2e`: the same step for a sampler whose RNG and rollout
storage live on the HOST like the reference's (exploration noise H2D, the whole transition D2H into pinned memory every
step).  Extra keys carry the other BASELINE configs (config3, config5, ppo_iteration = config 4) and the grasp phase.
The reference arm times the CPU restatement of the same path (oracle/, float64; MuJoCo itself cannot be installed
offline, profiles/r02_probe_mujoco_gpu_box.txt) under a multiprocessing sampler that mirrors agent_handmimic.py:430-535.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec at 1/2/4/8 B200 vs mujoco_py sampler on N host cores"
UNIT = "env-steps/s"
ALGO_BYTES_PER_ENV_STEP = 6000.0          # ~1500 float32 of state / action / obs traffic per env step
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")   # ncu dram__bytes per hoic_step, keyed "obj:envs"
TIMING_CLASSES = ("k_p1", "k_p2", "k_p3s", "k_p3", "k_tail", "k_reset", "k_order", "hoic_step")


# ---------------------------------------------------------------------------------------------- CPU baseline
def _cpu_worker(obj, seed, seconds, q):
    """One sampler process: batch-1 float64 policy -> oracle env.step -> reward (agent_handmimic.py:430-501)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from common import model, oracle_env
    from hoic_b200 import ppo
    mdl = model(obj)
    env = oracle_env(obj, seed=1)
    policy, _ = ppo.make_policy_value(mdl.cfg, device="cpu", dtype=torch.float64, seed=1)
    zf = ppo.ZFilter((617,), clip=5.0, device="cpu")
    gen = torch.Generator().manual_seed(seed)
    steps = 0
    t_pol = t_env = 0.0
    t0 = time.perf_counter()
    obs = env.reset()
    while time.perf_counter() - t0 < seconds:
        a0 = time.perf_counter()
        x = zf(torch.as_tensor(obs)[None])
        act = policy.select_action(x, generator=gen)[0].numpy()
        a1 = time.perf_counter()
        obs, r, done, info = env.step(act)
        a2 = time.perf_counter()
        t_pol += a1 - a0
        t_env += a2 - a1
        steps += 1
        if done:
            env.clear_warnings()
            obs = env.reset()
    q.put((steps, time.perf_counter() - t0, t_pol, t_env))


def cpu_baseline(obj, seconds, procs):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_cpu_worker, args=(obj, 100 + i, seconds, q)) for i in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    rate = sum(r[0] / r[1] for r in res)
    pol, envt = sum(r[2] for r in res), sum(r[3] for r in res)
    return dict(value=rate, unit=UNIT, cores=procs, kind="port",
                sample=f"{steps} env steps of {obj} in {procs} processes x {seconds:.0f} s (oracle/ float64 restatement, not mujoco_py; "
                       f"policy {100 * pol / max(pol + envt, 1e-9):.0f}% / env {100 * envt / max(pol + envt, 1e-9):.0f}% of step time)",
                wall_s=wall)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """Keep the samples taken inside [t0, t1] (the sampler is started early: nvidia-smi needs ~0.1 s to produce its
        first line, longer than a short timed region)."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for t, r in self.rows if t0 <= t <= t1]
        for r in (inside if inside else [r for _, r in self.rows[-3:]]):
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def flop_model(c0, c1):
    """Algorithmic FLOPs from the kernels' integer counters, total and per kernel of the pipeline."""
    d = {k: c1[k] - c0[k] for k in c1}
    sub, nefc, ncon = d["substeps"], d["nefc"], d["ncon"]
    nefc_avg = nefc / max(sub, 1)
    nv = 32
    per = {
        # FK 8.0k, comPos 3.5k, CRB 3.6k, L'DL 3.5k, RNE + passive 5.3k, M^-1 0.9k, Euler 4.5k, PD 8.7k, applyFT x2 1.2k
        "k_p1": sub * 39000.0,
        "k_p2": sub * 900.0 + 480.0 * d["mpr_iters"] + 700.0 * ncon + 30.0 * nefc,
        "k_p3": d["newton_iters"] * (4 * nefc_avg * nv + 0.7 * nefc_avg * nv * (nv + 1) + nv ** 3 / 3 + 2 * nv * nv) + 12 * nefc_avg * d["ls_evals"],
        "k_tail": d["env_steps"] * 12000.0 + 60.0 * 60.0 * d["qp_iters"],
    }
    return float(sum(per.values())), per, d


def load_traffic(obj, n):
    try:
        return json.load(open(TRAFFIC_FILE)).get(f"{obj}:{n}")
    except Exception:
        return None


def build_env(obj, N, local, rank, world, start_lo=None):
    from hoic_b200 import dist as hdist
    from hoic_b200.assets import load_model
    from hoic_b200.vecenv import HandMimicVecEnv
    if obj == "mixed":
        # BASELINE config 5: env i uses model (global env id) mod 3, every rank sees the same mix
        models = [load_model(o) for o in ("box", "bottle", "banana")]
        motions = [mm.motion(s) for mm in models for s in range(1, 9)]
        motion_model = [k for k in range(3) for _ in range(8)]
        ids = hdist.mixed_model_ids(N * world, 3, rank, world)
        return models[0], HandMimicVecEnv(models, N, motions, motion_model=motion_model, env_model_id=ids, device=local, seed=1 + rank, auto_reset=True)
    mdl = load_model(obj)
    return mdl, HandMimicVecEnv(mdl, N, [mdl.motion(s) for s in range(1, 9)], device=local, seed=1 + rank, auto_reset=True)


def timed_rollout(sampler, steps, flush, world, dist, dev):
    """K timed sampler steps, L2 flushed (untimed) between steps, CUDA events on the launching stream; max over ranks."""
    import torch
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.time()
    for k in range(steps):
        if flush is not None:
            flush.fill_(k & 0xff)
        ev[k][0].record()
        sampler.step()
        ev[k][1].record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return float(tmax.item()), t0, t1


def kernel_timing(env, sampler, steps):
    """Duration of hoic_step (CUDA events on the launching stream around replays of a graph that holds only that call, the
    way the sampler step runs it; a few extra rollout steps, not part of the timed region) and, from a second pass of eager
    calls with the library's per-launch event timing, how the stream time splits over the pipeline's kernels."""
    import ctypes as C
    import torch
    L = env.L
    actions = sampler.act
    l0 = L.hoic_launch_count()
    for _ in range(2):
        sampler.compute_action()
        env.step(actions)
    torch.cuda.synchronize()
    launches = (L.hoic_launch_count() - l0) // 2
    # hoic_step alone as the sampler runs it: captured into a CUDA graph, replayed, CUDA events around every replay
    dev = actions.device
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        env.step(actions)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        sampler.compute_action()        # untimed: the rollout goes on under the policy, so the states stay those of the timed region
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    del g
    L.hoic_set_timing(env._hb, 1)
    for _ in range(steps):
        sampler.compute_action()
        env.step(actions)
    ms = (C.c_double * 8)()
    n = C.c_int(0)
    L.hoic_get_timing(env._hb, ms, C.byref(n))
    L.hoic_set_timing(env._hb, 0)
    kt = {k: ms[i] / max(n.value, 1) for i, k in enumerate(TIMING_CLASSES)}
    kt["hoic_step_with_event_pairs"] = kt["hoic_step"]
    kt["hoic_step"] = step_ms
    kt["launches_per_hoic_step"] = int(launches)
    return kt


def short_config(obj, N, local, rank, world, dist, dev, policy_seed=1, steps=8, preroll=40, grasp=False):
    """value / kernel time of another BASELINE config (or of the grasp phase) on this rank's GPU: same sampler step."""
    import torch
    from hoic_b200 import dist as hdist, ppo
    from hoic_b200.sampler import SamplerStep
    mdl, env = build_env(obj, N, local, rank, world)
    policy, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=policy_seed)
    zf = ppo.ZFilter((617,), clip=5.0, device=dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    sampler = SamplerStep(env, policy, zf, gen, stats=None, policy_dtype="bf16", graph=True)
    if grasp:
        # grasp / lift phase of the synthetic motions: every episode starts in frames 100..179 and is timed over frames
        # ~105..125 (fingers closed on the object, lift under way): contact-rich substeps
        start = torch.randint(100, 180, (N,), device=dev, dtype=torch.int32)
        mid = torch.randint(0, 8, (N,), device=dev, dtype=torch.int32)
        env.reset(motion_id=mid, start_idx=start)
        env.auto_reset = False
        preroll = 5
    else:
        env.reset()
    sampler.warmup()
    for _ in range(preroll):
        sampler.step()
    c0 = env.counters()
    total_ms, _, _ = timed_rollout(sampler, steps, None, world, dist, dev)
    c1 = env.counters()
    flops, per, dc = flop_model(c0, c1)
    kt = kernel_timing(env, sampler, 3)
    out = dict(workload=f"{obj}, {N} envs per GPU" + (", episodes started in frames 100-179 (grasp / lift)" if grasp else ""),
               value=world * N * steps / (total_ms * 1e-3), unit=UNIT, ms_per_step=total_ms / steps, kernel_ms=kt["hoic_step"],
               flops_per_env_step=flops / max(dc["env_steps"], 1),
               counters_per_substep={k: dc[k] / max(dc["substeps"], 1) for k in ("ncon", "nefc", "newton_iters", "mpr_iters")},
               max_ncon=c1["max_ncon"], max_nbasis=c1["max_nbasis"], max_nedge=c1["max_nedge"],
               fail_overflow=dc["fail_overflow"], fail_physics=dc["fail_physics"])
    env.close()
    del sampler, env
    torch.cuda.empty_cache()
    return out


def ppo_iteration(local, rank, world, dev, iters=3):
    """BASELINE config 4: banana_future5_light_add_geom, one full PPO iteration (>= 50k-sample rollout + GAE + 5-epoch update)
    across the run's GPUs (envs sharded, gradients / statistics all-reduced)."""
    import copy
    import torch
    from hoic_b200.agent import AgentHandMimic
    from hoic_b200.assets import load_model
    mdl = load_model("banana")
    cfg = copy.deepcopy(mdl.cfg)
    n_envs = 2048
    cfg.min_batch_size = 51200
    agent = AgentHandMimic(mdl, n_envs, device=local, cfg=cfg)
    infos = [agent.optimize_policy(i) for i in range(iters)]
    last = infos[-1]
    out = dict(workload=f"banana, {n_envs} envs per GPU x {last['horizon']} steps = {last['samples']} samples per iteration, GAE + "
                        f"{cfg.num_optim_epoch} epochs (rollout MLP bf16, GAE + update TF32 tensor cores, fp32 master weights)",
               T_sample=last["T_sample"], T_update=last["T_update"], T_total=last["T_total"], samples=last["samples"],
               samples_per_s=last["samples"] / last["T_total"], avg_reward=last["log"]["avg_reward"],
               value_loss=last["value_loss"], iterations_timed=iters)
    agent.close()
    del agent
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from hoic_b200 import capi, dist as hdist, ppo
    from hoic_b200.sampler import SamplerStep

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs
    mdl, env = build_env(args.obj, N, local, rank, world)
    policy, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    zf = ppo.ZFilter((617,), clip=5.0, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    stats = hdist.RolloutStats(dev)
    sampler = SamplerStep(env, policy, zf, gen, stats=stats, policy_dtype=args.policy_dtype, graph=True)
    host_sampler = SamplerStep(env, policy, zf, gen, stats=stats, policy_dtype=args.policy_dtype, graph=True, host_noise=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = capi.lib()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    env.reset()
    sampler.warmup()
    # untimed pre-roll: episodes start at random frames of the 300-frame motions and last up to ~200 steps; after 150
    # steps the mix of episode phases (approach / grasp / lift / hold) over the batch is stationary
    for _ in range(max(args.warmup, 3) + args.preroll):
        sampler.step()
    torch.cuda.synchronize()
    stats.reset()

    c0 = env.counters()
    l0 = L.hoic_launch_count()
    total_ms_max, t_wall0, t_wall1 = timed_rollout(sampler, args.steps, flush, world, dist, dev)
    clocks.window(t_wall0, t_wall1)
    c1 = env.counters()
    launches = L.hoic_launch_count() - l0      # 0 inside a replayed graph: counted from one eager step below
    value = world * N * args.steps / (total_ms_max * 1e-3)
    clk = clocks.stop() if rank == 0 else None
    tot = stats.allreduce()

    # ---- per-kernel durations and the launch count of one step (eager launches with the library's event timing)
    kt = kernel_timing(env, sampler, 10)
    launches_per_step = kt["launches_per_hoic_step"] + 5       # + k_zf_moments, k_zf_apply16, k_sample_action, k_get_episode, k_logger

    # ---- e2e: host-resident sampler (noise H2D, transition D2H, one graph launch + stream sync per step)
    h_noise = host_sampler.host_noise_buffers()       # the sampler's two pinned staging buffers: the host RNG draws straight into them
    for b in h_noise:
        b.normal_()
    for _ in range(3):
        host_sampler.step_host(h_noise[0])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        # the host's RNG draw is part of the step, as in the reference's sampler; the draw for step k + 1 runs on the host
        # while the device works on step k (two pinned buffers), then the host waits for step k's transition
        host_sampler.step_host_begin(h_noise[k & 1])
        h_noise[(k + 1) & 1].normal_()
        host_sampler.step_host_end()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * args.steps / float(te.item())
    host_sampler_bytes = (host_sampler.h2d_bytes, host_sampler.d2h_bytes)

    extras = {}
    if not args.no_extras:
        def guarded(name, fn):
            try:
                extras[name] = fn()
            except Exception as e:  # noqa: BLE001  (an extra must never take the headline line down)
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        env.close()
        del sampler, host_sampler, env, flush
        torch.cuda.empty_cache()
        guarded("grasp_phase", lambda: short_config(args.obj, N, local, rank, world, dist, dev, grasp=True))
        guarded("config3", lambda: short_config("bottle", 16384, local, rank, world, dist, dev))
        guarded("config5", lambda: short_config("mixed", 65536, local, rank, world, dist, dev, steps=4, preroll=20))
        guarded("ppo_iteration", lambda: ppo_iteration(local, rank, world, dev))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cfg_name = {("box", 4096): "BASELINE config 2", ("bottle", 16384): "BASELINE config 3", ("mixed", 65536): "BASELINE config 5"}.get(
        (args.obj, N), "non-headline size")
    flops, per_flops, dc = flop_model(c0, c1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    tf = C.c_double(0)
    L.hoic_measure_fp32_peak(local, C.byref(tf))
    step_ms = kt["hoic_step"]
    flops_step = flops / args.steps
    fp32_ach = flops_step / (step_ms * 1e-3) / 1e12
    # dominant kernel: the Newton solve (both capacity tiers; they run side by side)
    p3_ms = kt["k_p3s"]
    p3_ach = per_flops["k_p3"] / args.steps / max(p3_ms * 1e-3, 1e-12) / 1e12
    traffic = load_traffic(args.obj, N)
    ach_gbs = ALGO_BYTES_PER_ENV_STEP * N / (step_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (300-frame approach-grasp-lift motions, random-init policy)",
        "config": {"workload": f"{args.obj}_future5_light_add_geom, {N} batched envs per GPU ({cfg_name}); sampler step = normaliser + "
                               f"policy MLP ({args.policy_dtype}) + env.step + reward + auto-reset + statistics, one CUDA graph per step",
                   "envs_per_gpu": N, "sim_step": 15, "preroll_steps": args.preroll,
                   "l2": "flushed between steps (256 MiB write, untimed)", "parallelism": f"env-sharded x{world}"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_sampler_bytes[0]), "d2h_bytes_per_step": int(host_sampler_bytes[1]),
                "api": "SamplerStep.step_host_begin / step_host_end: exploration noise drawn by the host RNG (the draw for the next step "
                       "overlaps the device's work on this one) from pinned HOST memory in, normalised state / action / reward / mask / info "
                       "of the transition into pinned HOST memory out (state / action copies run underneath hoic_step), one stream "
                       "synchronise per step; normaliser + policy + hoic_step + auto-reset in between"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clk,
        "roofline": {"bound": "fp32", "achieved": fp32_ach, "peak": tf.value, "unit": "TFLOP/s", "frac": fp32_ach / max(tf.value, 1e-9),
                     "traffic": traffic, "kernel": "hoic_step pipeline (k_p1 + k_p2 + k_p3s || k_p3 per substep, k_tail)",
                     "kernel_ms": step_ms, "flops_per_env_step": flops / max(dc["env_steps"], 1),
                     "peak_source": "hoic_measure_fp32_peak (FP32 FMA micro-benchmark, this run): the path is FP32-pipe / latency "
                                    "bound, neither HBM nor tensor cores bind it (SURVEY.md 8d)",
                     "traffic_source": "profiles/r02_dram_traffic.json (ncu dram__bytes_read + write summed over the launches of one hoic_step)",
                     "dominant_kernel": {"name": "k_p3s (Newton solve, small capacity tier)", "ms_per_step": p3_ms,
                                         "share_of_stream_time": p3_ms / max(sum(kt[k] for k in TIMING_CLASSES[:7]), 1e-9),
                                         "note": "ms_per_step sums the launches of all parallel streams (the batch runs as interleaved "
                                                 "parts, the solver's two tiers side by side): concurrent time counts more than once",
                                         "achieved": p3_ach, "frac": p3_ach / max(tf.value, 1e-9)},
                     "kernel_ms_per_step": kt},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "algorithmic_bytes": ALGO_BYTES_PER_ENV_STEP * N, "traffic": traffic, "peak_source": peak_src,
                         "note": "does not bind; the working set of the pipeline travels between kernels through L2"},
        "counters_per_substep": {k: dc[k] / max(dc["substeps"], 1) for k in ("ncon", "nefc", "newton_iters", "ls_evals", "mpr_iters")},
        "capacity_high_water": {"max_ncon": c1["max_ncon"], "max_nbasis": c1["max_nbasis"], "max_nedge": c1["max_nedge"],
                                "capacity": [32, 80, 128]},
        "rollout": {"avg_reward": tot["avg_reward"], "episodes": tot["num_episodes"], "avg_episode_len": tot["avg_episode_len"],
                    "fail_overflow": dc["fail_overflow"], "fail_physics": dc["fail_physics"]},
    }
    out.update(extras)
    # CPU baseline: bounded sample on the host cores (N=1 only)
    if world == 1 and not args.no_cpu:
        procs = args.cpu_procs or (os.cpu_count() or 1)
        out["cpu_baseline"] = cpu_baseline(args.obj, args.cpu_seconds, procs)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    procs = args.cpu_procs or (os.cpu_count() or 1)
    per = max(2.0, args.cpu_seconds)
    vals, last = [], None
    for k in range(max(args.warmup, 0) and 1):
        cpu_baseline(args.obj, 2.0, procs)
    t0 = time.perf_counter()
    for k in range(max(1, min(args.steps, 3))):
        last = cpu_baseline(args.obj, per, procs)
        vals.append(last["value"])
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / max(v, 1e-9), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic (same motions / random-init policy as the GPU arm)",
           "config": {"workload": f"{args.obj}_future5_light_add_geom, reference-style multiprocessing sampler, 1 env per process"},
           "cpu_baseline": dict(last, value=v), "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--obj", default="box")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip grasp_phase / config3 / config5 / ppo_iteration")
    ap.add_argument("--preroll", type=int, default=150)
    ap.add_argument("--policy-dtype", default="bf16", choices=["bf16", "tf32", "fp32"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
