"""Tiny driver for ncu: eager (non-graph) launches of the fused step kernel.
   python tools/profile_step.py <task>[_torque] <n_envs> <pool> <steps>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import drl_on_robot_arm_b200 as pkg

task, n, pool, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torque = task.endswith("_torque")
task = task.split("_")[0]
envs = [pkg.BatchedArmEnv(task, n_envs=n, device="cuda:0", seed=0, auto_reset=True, env_id_offset=b * n,
                          mode="torque" if torque else "ik_teleport") for b in range(pool)]
NA = 7 if n >= (1 << 20) else 61   # action sets in rotation: every env sees a different action at each step
if torque:
    acts = (torch.rand((NA, n, 7), device="cuda") * 2.0 - 1.0) * 30.0      # joint torques, N m
else:
    acts = torch.rand((NA, n, 3), device="cuda") * 1.4 - 0.7
    if task != "reach":
        acts *= 0.4 / 0.7
for k in range(steps):
    envs[k % pool].step(acts[k % NA])
torch.cuda.synchronize()
print("done", task, "torque" if torque else "ik", n, pool, steps)
