"""python tools/sanitize_smoke.py [what] : small runs of the round-2 kernels for compute-sanitizer (memcheck / racecheck).
   what = policy | server | tracked | cube   (default: all).  Ragged sizes on purpose."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200.algo.nets import PolicyNet

what = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(0)
if what in ("all", "policy"):
    for task, n in (("reach", 1000), ("push", 77)):
        env = pkg.BatchedArmEnv(task, n_envs=n, device="cuda:0", seed=1)
        net = PolicyNet(env.obs_dim, 256, 3, 0.7).to("cuda:0")
        obs = torch.rand((n, env.obs_dim), device="cuda")
        for _ in range(3):
            env.policy_act(net, obs, noise_std=0.5, clip=0.7)
            env.policy_act(net, obs)
        torch.cuda.synchronize()
        env.close()
    print("policy ok")
if what in ("all", "tracked", "cube"):
    for task, n in (("reach", 1000), ("push", 333), ("pick", 130)):
        env = pkg.BatchedArmEnv(task, n_envs=n, device="cuda:0", seed=2, auto_reset=True, max_steps=5)
        env.reset()
        for k in range(8):
            a = (torch.rand((n, 3), device="cuda") * 2 - 1) * 0.5
            env.step(a, final_obs=True, track=True)
        torch.cuda.synchronize()
        env.close()
    print("tracked / cube ok")
if what in ("all", "server"):
    for task, n in (("reach", 1000), ("push", 200)):
        env = pkg.ArmSimHandle(task, n_envs=n, seed=3, auto_reset=True, max_steps=5)
        env.host_server(200000)
        env.reset_host()
        for k in range(6):
            env.step_host(rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32))
        env.get_state(0)
        env.step_host(rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32))
        env.close()
    print("server ok")
