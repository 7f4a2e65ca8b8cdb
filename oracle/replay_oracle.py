"""numpy restatement of the reference's trajectory replay relabelling -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows /root/reference utils/rl_utils.py: ReplayBuffer_Trajectory_reach.sample :119-152 and
ReplayBuffer_Trajectory_push.sample :165-199, with the random draws taken as INPUT (traj, step, goal step or -1) so
that the CUDA gather kernel, this restatement and the reference's own output (tests/golden/her_*.npz, produced by
tools/gen_her_golden.py from the real reference class) can be compared on identical picks.
Parity status: PINNED by those fixtures.
"""
import numpy as np


def relabel(states, actions, rewards, dones, lengths, picks, kind="reach", dis_threshold=0.1):
    """states [T, Lmax+1, O], actions [T, Lmax, A], rewards/dones [T, Lmax], lengths [T]; picks [B,3] = (traj, step, goal|-1).
    Returns dict(states, actions, next_states, rewards, dones) like the reference's batch."""
    out = dict(states=[], actions=[], next_states=[], rewards=[], dones=[])
    for traj, step, goal_step in np.asarray(picks):
        assert 0 <= step < lengths[traj]
        state = states[traj, step]                       # rl_utils.py:128
        next_state = states[traj, step + 1]              # :129
        action = actions[traj, step]                     # :130
        reward = rewards[traj, step]                     # :131
        done = bool(dones[traj, step])                   # :132
        if goal_step >= 0:                               # :134 use_her and uniform() <= her_ratio
            assert step + 1 <= goal_step <= lengths[traj]        # :135 randint(step_state + 1, length + 1)
            goal = states[traj, goal_step][:3]           # :136
            dis = np.sqrt(np.sum(np.square(next_state[:3] - goal)))   # :137
            reward = -0.1 if dis > dis_threshold else 1.0             # :138
            done = False if dis > dis_threshold else True             # :139
            if kind == "reach":
                state = np.hstack((state[:3], goal))                  # :140
                next_state = np.hstack((next_state[:3], goal))        # :141
            else:
                tail = state[6:10]                                    # :187 (state is already relabelled at :188)
                state = np.hstack((state[:3], goal, tail))
                next_state = np.hstack((next_state[:3], goal, state[6:10]))
        out["states"].append(state)
        out["next_states"].append(next_state)
        out["actions"].append(action)
        out["rewards"].append(reward)
        out["dones"].append(done)
    return {k: np.array(v) for k, v in out.items()}
