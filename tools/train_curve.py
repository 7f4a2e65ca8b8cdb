"""python tools/train_curve.py [task] [algo] [n_envs] [episode_times] [out.json]
The reference's `train_reach_with_TD3` loop (main.py:165-231) on the batched engine: N envs in lockstep, the reference's
update cadence (n_train updates per episode-time), HER, save-on-best bookkeeping.  Prints / stores the success-rate
series (one point per 25 episode-times, main.py:222) next to the reference's learning-curve pins."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import drl_on_robot_arm_b200 as pkg  # noqa: F401
from drl_on_robot_arm_b200 import distributed as D
from drl_on_robot_arm_b200 import metrics, train

# torchrun --nproc-per-node G tools/train_curve.py ... : n_envs is PER GPU (BASELINE config 5 = 8 x 4096, DARC)
rank, world, local = D.init_from_env("nccl")

task = sys.argv[1] if len(sys.argv) > 1 else "reach"
algo = sys.argv[2] if len(sys.argv) > 2 else "TD3_MLP"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
episode_times = int(sys.argv[4]) if len(sys.argv) > 4 else 100
out = sys.argv[5] if len(sys.argv) > 5 else None
budget_s = float(os.environ.get("TRAIN_BUDGET_S", "600"))

t_init = time.time()
if world > 1:   # first collective = communicator set-up (seconds with 8 ranks): keep it out of the training clock
    w = torch.ones(1, device="cuda:%d" % local)
    torch.distributed.all_reduce(w)
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"nccl_first_collective_s": time.time() - t_init}), flush=True)
sink = metrics.MetricsSink()
# WINDOW = replay ring length in lockstep rows (default 1024 = the last ~2 episode-times; the reference keeps every
# trajectory of a run: WINDOW >= episode_times * 501 does the same).  UTD = updates per finished env-episode (default:
# 40 per episode-TIME; UTD=40 is the reference's update-to-data ratio, main.py:209-212).  RATE_WINDOW = episodes per
# success-rate point (reference: 25).
window = int(os.environ.get("WINDOW", "1024"))
utd = float(os.environ["UTD"]) if "UTD" in os.environ else None
rate_window = int(os.environ["RATE_WINDOW"]) if "RATE_WINDOW" in os.environ else 5 * n * world
tr = train.make_trainer(task=task, algo=algo, n_envs=n, device="cuda:%d" % local, seed=int(os.environ.get("SEED", "0")), window=window,
                        metrics=sink, sync_every=int(os.environ.get("SYNC_EVERY", "32")),
                        window_episodes=rate_window, clip_actions=(os.environ.get("CLIP", "0") == "1"),
                        noise_std=float(os.environ["NOISE"]) if "NOISE" in os.environ else None, updates_per_episode=utd)
t0 = time.time()
chunk = 501
log = []
for k in range(episode_times):
    res = tr.run(chunk)
    torch.cuda.synchronize()
    el = time.time() - t0
    log.append({"episode_time": k + 1, "wall_s": el, "env_steps": res["env_steps"], "updates": res["updates"], "episodes": res["episodes"],
                "success_rate": res["success_rate"], "avg_return": res["avg_return"], "her_ratio": res["her_ratio"]})
    if rank == 0 and ((k + 1) % 5 == 0 or k == 0):
        print(json.dumps(log[-1]), flush=True)
    stop = torch.tensor([1.0 if el > budget_s else 0.0], device="cuda:%d" % local)
    if world > 1:
        torch.distributed.all_reduce(stop, op=torch.distributed.ReduceOp.MAX)      # every rank leaves together
    if float(stop) > 0:
        break
summary = {"task": task, "algo": algo, "n_envs": n, "n_gpus": world, "replay_window_rows": window, "updates_per_episode": utd,
           "noise_std": tr.noise_std, "episodes": log[-1]["episodes"], "updates": log[-1]["updates"], "n_envs_total": n * world, "episode_times": len(log), "wall_s": time.time() - t0,
           "env_steps_per_s_incl_learning": log[-1]["env_steps"] / (time.time() - t0),
           "success_rate_series": sink.series.get("success_rate", []), "log": log}
if rank == 0:
    print(json.dumps({k: v for k, v in summary.items() if k != "log"}))
    if out:
        json.dump(summary, open(out, "w"), indent=1)
if world > 1:
    D.shutdown()
