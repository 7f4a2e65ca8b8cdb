"""kernel list of one lockstep rollout step (eager, for `ncu --metrics gpu__time_duration.sum`)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drl_on_robot_arm_b200 import train
tr = train.make_trainer(task="reach", algo="TD3_MLP", n_envs=4096, device="cuda:0", seed=0, minimal_episodes=10 ** 12,
                        sync_every=10 ** 9, use_cuda_graph=False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    tr.rollout_step()
torch.cuda.synchronize()
print("ok")
