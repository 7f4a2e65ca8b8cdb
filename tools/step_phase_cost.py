"""per-launch time of the reach step kernel as a function of the episode step (1M envs, events per launch)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200 import _lib as L
dev = torch.device('cuda:0')
n = 1 << 20
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
env = pkg.BatchedArmEnv(task, n_envs=n, device=dev, seed=0, auto_reset=True)
env.reset()
T = 560
acts = torch.rand((8, n, 3), device=dev) * 1.4 - 0.7
if task != "reach": acts *= 0.4 / 0.7
ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
its = []
torch.cuda.synchronize()
ev[0].record()
for k in range(T):
    env.step(acts[k % 8])
    ev[k + 1].record()
    if k in (0, 1, 2, 5, 20, 100, 300):
        its.append((k, float(torch.from_numpy(env.get_state(L.F_IK_ITERS)).float().mean())))
torch.cuda.synchronize()
us = np.array([ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(T)])
print("step: us/launch", [(k, round(us[k], 1)) for k in (0, 1, 2, 3, 5, 10, 20, 50, 100, 200, 300, 400, 499, 500, 501, 502, 510, 550)])
print("mean ik iters at step:", its)
print("mean over steps 100..500: %.1f us ; steps 1..20: %.1f us" % (us[100:500].mean(), us[1:20].mean()))
