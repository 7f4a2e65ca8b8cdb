#!/usr/bin/env python
"""System-level check of the ORACLE's physics, build container only (reads /root/reference, never runs on the GPU box).

The REFERENCE's own learner -- `TD3_MLP` / `DATD3_MLP` ... (algo/*_mlp.py) and `ReplayBuffer_Trajectory_push` /
`_reach` (utils/rl_utils.py), imported unmodified from /root/reference -- is driven through the reference's own
training loop (main.py:449-515 `train_push_with_TD3`, :518-584 `train_pick_with_TD3`, :165-231 reach) against ONE
oracle env (oracle/armsim_oracle.c).  The only thing that is not the reference's code is the env: if the oracle's
cube / gripper model behaves like Bullet's, the learning curve must look like `visdata/push/updata_TD3/*.csv`
(success 0.5 after ~550 episodes, 0.9 after ~725; first-episode returns around -320..-504, successful returns ~ +115).

    python tests/system/ref_learner_on_oracle.py push TD3_MLP 1000 out.json
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ARMSIM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import algo as RA  # noqa: E402  (the reference's agents)
from utils import rl_utils as RU  # noqa: E402  (the reference's replay)
from config import opt  # noqa: E402  (the reference's hyper-parameters)

from oracle import oracle as O  # noqa: E402

TASKS = {"reach": O.TASK_REACH, "push": O.TASK_PUSH, "pick": O.TASK_PICK}


class OracleEnv:
    """gym-style single env over the oracle (what envs/rl_push_env.py is over pybullet)."""

    def __init__(self, task, seed=0):
        self.task = task
        kw = {}
        if os.environ.get("REACH_DIS"):          # the reference's "harder" reach runs: opt.reach_dis = 0.005 (visdata/reach/*_0.005)
            kw["reach_dis"] = float(os.environ["REACH_DIS"])
        self.sim = O.OracleSim(O.default_config(TASKS[task], 1, seed=seed, auto_reset=0, **kw))

    def reset(self):
        return self.sim.reset()[0].astype(np.float64)

    def step(self, action):
        obs, r, d, s = self.sim.step(np.asarray(action, np.float32).reshape(1, 3))
        return obs[0].astype(np.float64), float(r[0]), bool(d[0]), bool(s[0])


def main():
    task = sys.argv[1] if len(sys.argv) > 1 else "push"
    algo = sys.argv[2] if len(sys.argv) > 2 else "TD3_MLP"
    episodes = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    out = sys.argv[4] if len(sys.argv) > 4 else None
    seed = int(os.environ.get("SEED", opt.random_seed))
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    torch.set_num_threads(int(os.environ.get("THREADS", "2")))
    env = OracleEnv(task, seed)
    if task == "reach":
        state_dim, action_bound, noise = 6, 0.4 + 0.3, 1.0 * opt.gamma           # main.py:173,200
        buf = RU.ReplayBuffer_Trajectory_reach(opt.buffer_size)
        ok_reward = 0.0
    else:
        state_dim, action_bound = 9, 0.4                                          # main.py:455-457
        noise = action_bound * opt.gamma                                          # main.py:484
        buf = RU.ReplayBuffer_Trajectory_push(opt.buffer_size)
        ok_reward = 100.0                                                         # main.py:486
    agent = getattr(RA, algo)(state_dim=state_dim, action_dim=3, action_bound=action_bound, hidden_dim=opt.hidden_dim,
                              device=torch.device("cpu"))
    her_ratio = opt.her_ratio
    returns, rates, lens = [], [], []
    rate, max_rate = 0.0, 0.0
    t0 = time.time()
    for ep in range(episodes):
        state = env.reset()
        traj = RU.Trajectory(state)
        done, ret, n = False, 0.0, 0
        while not done:
            action = agent.take_action(state)
            action = action + np.random.normal(0, noise, size=3)
            state, reward, done, _ = env.step(action)
            if reward == ok_reward:
                rate += 1
            ret += reward
            n += 1
            traj.store_step(action, state, reward, done)
        buf.add_trajectory(traj)
        returns.append(ret); lens.append(n)
        if buf.size() >= opt.minimal_episodes:
            for _ in range(opt.n_train):
                agent.train(buf.sample(opt.batch_size, use_her=True, her_ratio=her_ratio))
        if (ep + 1) % 25 == 0:
            rate /= 25.0
            rates.append(rate)
            print(json.dumps({"episode": ep + 1, "success_rate": rate, "avg_return_25": float(np.mean(returns[-25:])),
                              "avg_len_25": float(np.mean(lens[-25:])), "her_ratio": her_ratio, "wall_s": time.time() - t0}), flush=True)
            if rate >= max_rate:
                max_rate = rate
                her_ratio *= 0.75
            rate = 0.0
    if out:
        json.dump({"task": task, "algo": algo, "episodes": episodes, "seed": seed, "returns": returns, "lengths": lens,
                   "success_rate_per_25": rates, "wall_s": time.time() - t0,
                   "what": "the reference's learner + replay + loop (imported from /root/reference) on the oracle env"},
                  open(out, "w"))


if __name__ == "__main__":
    main()
