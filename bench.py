"""bench.py -- HandMimic rollout throughput (env-steps/s), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--obj box]

Our arm, one step = one pass of the sampler hot path over the batch: observation normaliser -> random-init policy MLP
(PyTorch) -> hoic_step (15 physics substeps + control + contact bookkeeping + RFC QP + termination + reward +
observation, one fused CUDA kernel) -> in-place auto-reset of finished episodes.  Workload at N=1 is BASELINE config 2:
box_future5_light_add_geom, 4096 batched envs per GPU (weak scaling: 4096 envs on every rank).
The reference arm times the CPU restatement of the same path (oracle/, float64; MuJoCo itself cannot be installed
offline) under a multiprocessing sampler that mirrors uhc/agents/agent_handmimic.py:430-535.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch from the `ncu --set full` capture of this workload
# (profiles/r01_v14_k_step_raw.csv: 16.49 MB read + 8.27 MB write at box / 4096 envs); other workloads: not captured -> null
NCU_TRAFFIC_BYTES = {("box", 4096): 16.493312e6 + 8.271104e6}


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
    """Algorithmic FLOPs from the kernel's integer counters."""
    d = {k: c1[k] - c0[k] for k in c1}
    sub, nefc, ncon = d["substeps"], d["nefc"], d["ncon"]
    nefc_avg = nefc / max(sub, 1)
    nv = 32
    f = sub * (39000 + 900) + 480 * d["mpr_iters"] + 700 * ncon + 30 * nefc
    f += d["newton_iters"] * (4 * nefc_avg * nv + 0.7 * nefc_avg * nv * (nv + 1) + nv ** 3 / 3 + 2 * nv * nv) + 12 * nefc_avg * d["ls_evals"]
    f += d["env_steps"] * 12000 + 60 * 60 * d["qp_iters"]
    return float(f), d


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hoic_b200 import capi, dist as hdist, ppo
    from hoic_b200.assets import load_model
    from hoic_b200.vecenv import HandMimicVecEnv

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs
    if args.obj == "mixed":
        # BASELINE config 5: env i uses model (global env id) mod 3, every rank sees the same mix
        models = [load_model(o) for o in ("box", "bottle", "banana")]
        motions = [mm.motion(s) for mm in models for s in range(1, 9)]
        motion_model = [k for k in range(3) for _ in range(8)]
        ids = hdist.mixed_model_ids(N * world, 3, rank, world)
        mdl = models[0]
        env = HandMimicVecEnv(models, N, motions, motion_model=motion_model, env_model_id=ids, device=local, seed=1 + rank, auto_reset=True)
    else:
        mdl = load_model(args.obj)
        motions = [mdl.motion(s) for s in range(1, 9)]
        env = HandMimicVecEnv(mdl, N, motions, device=local, seed=1 + rank, auto_reset=True)
    policy, _ = ppo.make_policy_value(mdl.cfg, device=dev, dtype=torch.float32, seed=1)
    zf = ppo.ZFilter((617,), clip=5.0, device=dev, dtype=torch.float32)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    stats = hdist.RolloutStats(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = capi.lib()

    def one_step(obs):
        x = zf(obs)
        act = policy.select_action(x, generator=gen)
        o, r, d, info = env.step(act)
        stats.update(r, info["reward_info"], d)
        return o

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    obs = env.reset()
    for _ in range(max(args.warmup, 3)):
        obs = one_step(obs)
    torch.cuda.synchronize()

    # ---- timed region: K steps, L2 flushed (untimed) between steps, CUDA events on the launching stream
    c0 = env.counters()
    l0 = L.hoic_launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        e0, e1, e2 = ev[k]
        e0.record()
        x = zf(obs)
        act = policy.select_action(x, generator=gen)
        e1.record()                     # kernel-only window: hoic_step (k_step) + the auto-reset kernel
        obs, r, d, info = env.step(act)
        e2.record()
        stats.update(r, info["reward_info"], d)
    torch.cuda.synchronize()
    clocks.window(t_wall0, time.time())
    if world > 1:
        dist.barrier()
    launches = L.hoic_launch_count() - l0
    c1 = env.counters()
    ms_steps = [e0.elapsed_time(e2) for e0, e1, e2 in ev]
    ms_kernel = [e1.elapsed_time(e2) for e0, e1, e2 in ev]
    total_ms = float(sum(ms_steps))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = world * N * args.steps / (total_ms_max * 1e-3)
    clk = clocks.stop() if rank == 0 else None

    # ---- e2e: the reference-facing call with HOST buffers (HandMimicVecEnv.step_host): H2D of the actions, env ordering,
    # step kernel, auto-reset and D2H of obs / reward / done / info inside the timed region, one CUDA-graph launch per step
    h_act = torch.randn(N, 32).mul_(0.1).pin_memory()

    def e2e_step():
        env.step_host(h_act)

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * args.steps / float(te.item())
    tot = stats.allreduce()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cfg_name = {("box", 4096): "BASELINE config 2", ("bottle", 16384): "BASELINE config 3", ("mixed", 65536): "BASELINE config 5"}.get(
        (args.obj, N), "non-headline size")
    flops, dc = flop_model(c0, c1)
    kernel_ms = float(np.mean(ms_kernel))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    import ctypes as C
    tf = C.c_double(0)
    L.hoic_measure_fp32_peak(local, C.byref(tf))
    ach_gbs = ALGO_BYTES_PER_ENV_STEP * N / (kernel_ms * 1e-3) / 1e9
    fp32_ach = flops / args.steps / (kernel_ms * 1e-3) / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (300-frame approach-grasp-lift motions, random-init policy)",
        "config": {"workload": f"{args.obj}_future5_light_add_geom, {N} batched envs per GPU ({cfg_name}), policy MLP + "
                               "env.step + reward + auto-reset per step", "envs_per_gpu": N, "sim_step": 15,
                   "l2": "flushed between steps (256 MiB write, untimed)", "parallelism": f"env-sharded x{world}"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(env.h2d_bytes), "d2h_bytes_per_step": int(env.d2h_bytes),
                "api": "HandMimicVecEnv.step_host (pinned host buffers in and out, CUDA graph of copies + hoic_step + auto-reset)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": NCU_TRAFFIC_BYTES.get((args.obj, N)),
                     "algorithmic_bytes": ALGO_BYTES_PER_ENV_STEP * N, "traffic_source": "profiles/r01_v14_k_step_raw.csv (ncu --set full, per launch)",
                     "peak_source": peak_src, "kernel": "k_step", "kernel_ms": kernel_ms,
                     "note": "neither HBM nor tensor cores bind this kernel: it is FP32-pipe / latency bound, see roofline_fp32"},
        "roofline_fp32": {"bound": "fp32", "achieved": fp32_ach, "peak": tf.value, "unit": "TFLOP/s", "frac": fp32_ach / max(tf.value, 1e-9),
                          "flops_per_env_step": flops / max(dc["env_steps"], 1), "peak_source": "hoic_measure_fp32_peak (FMA micro-benchmark, this run)"},
        "counters_per_substep": {k: dc[k] / max(dc["substeps"], 1) for k in ("ncon", "nefc", "newton_iters", "ls_evals", "mpr_iters")},
        "rollout": {"avg_reward": tot["avg_reward"], "episodes": tot["num_episodes"], "fail_overflow": dc["fail_overflow"], "fail_physics": dc["fail_physics"]},
    }
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
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
