#!/usr/bin/env python
"""Generate tests/golden/her_{reach,push}.npz by running the REFERENCE's replay buffer
(/root/reference/utils/rl_utils.py: Trajectory :91-105, ReplayBuffer_Trajectory_reach.sample :119-152,
ReplayBuffer_Trajectory_push.sample :165-199) on seeded synthetic trajectories, recording which
(trajectory, step, goal step) each sample drew.  Run in the BUILD container only (needs /root/reference); the
fixtures are committed so nothing on the GPU box reads the reference.

The fixtures pin the relabelling arithmetic of oracle/replay_oracle.py (CPU test) and of the CUDA gather kernel
(GPU test): given the same picks they must reproduce the reference's batches.
"""
import os
import random
import sys

import numpy as np

REF = os.environ.get("ARMSIM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from utils import rl_utils as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_trajs(rng, n_traj, obs_dim, dtype):
    trajs = []
    for _ in range(n_traj):
        L = int(rng.integers(1, 41))
        s = np.zeros(obs_dim, dtype)
        s[:3] = rng.uniform(0.2, 0.6, 3)
        s[3:6] = rng.uniform(0.2, 0.6, 3)
        if obs_dim > 6:
            s[6:] = rng.uniform(0.2, 0.6, obs_dim - 6)
        t = R.Trajectory(s.copy())
        for k in range(L):
            a = rng.uniform(-0.7, 0.7, 3).astype(np.float32)
            s = s.copy()
            s[:3] += (0.12 * a).astype(dtype)              # random walk: future states straddle the 0.1 m threshold
            if obs_dim > 6:
                s[3:6] += (0.01 * rng.uniform(-1, 1, 3)).astype(dtype)
            t.store_step(a, s.copy(), float(-10 * np.linalg.norm(s[:3] - s[3:6])), bool(k == L - 1))
        trajs.append(t)
    return trajs


def run(kind, seed):
    rng = np.random.default_rng(seed)
    obs_dim, dtype = (6, np.float32) if kind == "reach" else (9, np.float64)
    buf = (R.ReplayBuffer_Trajectory_reach if kind == "reach" else R.ReplayBuffer_Trajectory_push)(1000)
    trajs = make_trajs(rng, 37, obs_dim, dtype)
    for t in trajs:
        buf.add_trajectory(t)
    index = {id(t): i for i, t in enumerate(trajs)}
    picks, cur = [], {}
    o_sample, o_randint, o_uniform = random.sample, np.random.randint, np.random.uniform

    def rec_sample(pop, k):
        out = o_sample(pop, k)
        if cur:
            picks.append((cur["traj"], cur["step"], cur.get("goal", -1)))
            cur.clear()
        cur["traj"] = index[id(out[0])]
        return out

    def rec_randint(*a):
        v = o_randint(*a)
        if "step" not in cur:
            cur["step"] = int(v)
        else:
            cur["goal"] = int(v)
        return v
    random.seed(seed)
    np.random.seed(seed)
    random.sample, np.random.randint = rec_sample, rec_randint
    try:
        B = 96
        batch = buf.sample(B, True, dis_threshold=0.1, her_ratio=0.8)
    finally:
        random.sample, np.random.randint, np.random.uniform = o_sample, o_randint, o_uniform
    picks.append((cur["traj"], cur["step"], cur.get("goal", -1)))
    assert len(picks) == B
    Lmax = max(t.length for t in trajs)
    states = np.zeros((len(trajs), Lmax + 1, obs_dim), np.float64)
    actions = np.zeros((len(trajs), Lmax, 3), np.float32)
    rewards = np.zeros((len(trajs), Lmax), np.float64)
    dones = np.zeros((len(trajs), Lmax), np.uint8)
    lengths = np.array([t.length for t in trajs], np.int32)
    for i, t in enumerate(trajs):
        states[i, :t.length + 1] = np.array(t.states)
        actions[i, :t.length] = np.array(t.actions)
        rewards[i, :t.length] = t.rewards
        dones[i, :t.length] = t.dones
    out = os.path.join(ROOT, "tests", "golden", "her_%s.npz" % kind)
    np.savez_compressed(out, states=states, actions=actions, rewards=rewards, dones=dones, lengths=lengths,
                        picks=np.array(picks, np.int32), out_states=np.array(batch["states"], np.float64),
                        out_next_states=np.array(batch["next_states"], np.float64),
                        out_actions=np.array(batch["actions"], np.float32),
                        out_rewards=np.array(batch["rewards"], np.float64),
                        out_dones=np.array(batch["dones"], np.uint8), dis_threshold=0.1, her_ratio=0.8)
    her = sum(1 for p in picks if p[2] >= 0)
    print(kind, "->", out, "B", B, "HER samples", her, "relabelled successes", int(np.sum(np.array(batch["rewards"]) == 1.0)))


if __name__ == "__main__":
    run("reach", 11)
    run("push", 12)
