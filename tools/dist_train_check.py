"""torchrun --nproc-per-node N tools/dist_train_check.py : N-GPU training check.  Every rank owns an env + replay
shard; after K lockstep steps with learning the agent replicas must be bit-identical on all ranks (the only exchange
is the flat gradient all-reduce) while the env shards differ."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler

import torch
import torch.distributed as dist

if os.environ.get("HANG_DUMP_S"):          # diagnosis: print every thread's stack and exit if the check has not finished by then
    faulthandler.dump_traceback_later(float(os.environ["HANG_DUMP_S"]), exit=True)

import drl_on_robot_arm_b200 as pkg  # noqa: F401
from drl_on_robot_arm_b200 import distributed as D
from drl_on_robot_arm_b200 import train

rank, world, local = D.init_from_env("nccl")
dev = torch.device("cuda", local)
tr = train.make_trainer(task="reach", algo=os.environ.get("ALGO", "DARC_MLP"), n_envs=512, device=dev, seed=5, window=256,
                        sync_every=8, window_episodes=4 * 512 * world)
tr.env.close()
tr.env = D.make_sharded_env("reach", 512 * world, device=dev, seed=5, auto_reset=True, max_steps=20)
out = tr.run(160)
flat = torch.cat([p.detach().reshape(-1) for l, _ in tr.agent._learners() for p in l.net.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
goal = torch.from_numpy(tr.env.get_state(2)).to(dev)
goals = [torch.empty_like(goal) for _ in range(world)]
dist.all_gather(goals, goal)
differ = not torch.equal(goals[0], goals[-1]) if world > 1 else True
if rank == 0:
    print("updates", out["updates"], "episodes", out["episodes"], "replicas identical:", same, "shards differ:", differ)
    if same and differ and out["updates"] > 0:
        print("DIST_TRAIN_OK")
D.shutdown()
sys.exit(0 if (same and differ) else 1)
