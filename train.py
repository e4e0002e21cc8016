"""Train HandMimic with PPO on the batched B200 engine (mirrors scripts/train_hand_mimic.py:60-82 of the reference).

  python train.py --obj box --envs 4096 --iters 10
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py --obj banana --envs 2048
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obj", default="box", help="box | bottle | banana | mixed")
    ap.add_argument("--envs", type=int, default=2048, help="environments per GPU")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--epoch", type=int, default=0, help="first epoch (adaptive schedules start here)")
    ap.add_argument("--min-batch-size", type=int, default=0, help="override cfg.min_batch_size (total over all GPUs)")
    ap.add_argument("--policy-dtype", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="arithmetic of the policy / value MLPs in rollout and update: bf16 tensor cores with fp32 accumulation and "
                         "fp32 master weights (default), TF32 tensor cores, or the plain fp32 pipe (the reference runs fp64)")
    ap.add_argument("--update-dtype", default="tf32", choices=["bf16", "tf32", "fp32"], help="arithmetic of GAE values / log-probabilities / the PPO update")
    ap.add_argument("--load", default="", help="checkpoint to start from (iter_XXXX.p layout of the reference)")
    ap.add_argument("--save", default="", help="write the final checkpoint here (rank 0)")
    ap.add_argument("--eval", action="store_true", help="run eval_policy (test mode, mean action) after training")
    a = ap.parse_args()
    import torch.distributed as dist
    from hoic_b200.agent import AgentHandMimic
    from hoic_b200.assets import load_model
    from hoic_b200 import dist as hdist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if a.obj == "mixed":
        models = [load_model(o) for o in ("box", "bottle", "banana")]
        ids = hdist.mixed_model_ids(a.envs * world, 3, rank, world)
    else:
        models, ids = [load_model(a.obj)], None
    agent = AgentHandMimic(models, a.envs, env_model_id=ids, device=local, policy_dtype=a.policy_dtype, update_dtype=a.update_dtype)
    if a.min_batch_size:
        agent.cfg.min_batch_size = a.min_batch_size
    if a.load:
        agent.load_checkpoint(a.load)
    for i in range(a.epoch, a.epoch + a.iters):
        info = agent.optimize_policy(i)
        if rank == 0:
            log = info.pop("log")
            print(json.dumps(dict(iter=i, policy_dtype=a.policy_dtype, update_dtype=a.update_dtype, avg_reward=log["avg_reward"], episodes=log["num_episodes"],
                                  avg_episode_len=log["avg_episode_len"], avg_episode_reward=log["avg_episode_reward"],
                                  samples_per_s=info["samples"] / info["T_total"], **{k: v for k, v in info.items()})), flush=True)
    if rank == 0 and a.save:
        agent.save_checkpoint(a.save)
    if rank == 0 and a.eval:
        m = agent.eval_policy()
        print(json.dumps({"eval": {k: v for k, v in m.items() if k != "per_motion"}}))
    agent.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
