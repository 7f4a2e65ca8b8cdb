"""CPU: the five MLP agents against fixtures produced by the REFERENCE's own agents (tools/gen_algo_golden.py runs
/root/reference/algo/*_mlp.py): same initial weights + same batches + same noise seeds -> same losses, same chosen
actions, same final weights of every network and target.  Tolerance 2e-6 absolute (fp32, different op fusion in
the Polyak update)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["DDPG_MLP", "TD3_MLP", "DADDPG_MLP", "DATD3_MLP", "DARC_MLP"]


def _load(agent, nets):
    for n, sd in nets.items():
        getattr(agent, n).load_state_dict(sd)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("bucket", [False, True])
def test_agent_reproduces_reference_updates(pkg, name, bucket):
    from drl_on_robot_arm_b200 import algo
    g = torch.load(os.path.join(GOLD, "algo_%s.pt" % name), weights_only=False)
    S, A, H, B, K = g["dims"]
    agent = getattr(algo, name)(state_dim=S, action_dim=A, action_bound=0.7, hidden_dim=H, device="cpu", distributed=bucket)
    _load(agent, g["init"])
    for k, b in enumerate(g["batches"]):
        torch.manual_seed(100 + k)
        out = agent.train({kk: vv.copy() for kk, vv in b.items()})
        if g["losses"][k] is not None:
            assert abs(float(out) - g["losses"][k]) <= 2e-5 * max(1.0, abs(g["losses"][k]))
    if name in ("TD3_MLP", "DADDPG_MLP", "DARC_MLP"):          # the reference DDPG / DATD3 never advance their counter
        assert agent.total_it == g["total_it"]
    for n, sd in g["final"].items():
        mine = getattr(agent, n).state_dict()
        for key, ref in sd.items():
            assert torch.allclose(mine[key], ref, atol=2e-6, rtol=0), (n, key, (mine[key] - ref).abs().max())
    acts = np.stack([agent.take_action(p) for p in g["probe"]])
    assert np.abs(acts - g["actions"]).max() <= 2e-6
    # the batched act() is the same policy
    assert np.abs(agent.act(torch.from_numpy(g["probe"])).numpy() - g["actions"]).max() <= 2e-6


@pytest.mark.parametrize("name", NAMES)
def test_save_load_roundtrip_and_reference_file_names(pkg, name, tmp_path):
    from drl_on_robot_arm_b200 import algo
    a = getattr(algo, name)(6, 3, 0.7, hidden_dim=16, device="cpu")
    prefix = str(tmp_path / "ck")
    a.save(prefix)
    files = sorted(os.listdir(tmp_path))
    expect = {"DDPG_MLP": ["ck_actor.pt", "ck_critic.pt"], "TD3_MLP": ["ck_actor.pt", "ck_critic.pt"],
              "DADDPG_MLP": ["ck_actor1.pt", "ck_actor2.pt", "ck_critic.pt"],
              "DATD3_MLP": ["ck_actor1.pt", "ck_actor2.pt", "ck_critic1.pt", "ck_critic2.pt"],
              "DARC_MLP": ["ck_actor1.pt", "ck_actor2.pt", "ck_critic1.pt", "ck_critic2.pt"]}[name]
    assert files == expect                                     # TD3_mlp.py:163-168, DARC_mlp.py:221-230
    b = getattr(algo, name)(6, 3, 0.7, hidden_dim=16, device="cpu")
    b.load(prefix)
    s = torch.rand(4, 6)
    assert torch.equal(a.act(s), b.act(s))
    for (la, _), (lb, _) in zip(a._learners(), b._learners()):
        for pa, pb in zip(lb.net.parameters(), lb.target.parameters()):
            assert torch.equal(pa, pb)                         # targets follow the loaded nets
    # full resume state
    b2 = getattr(algo, name)(6, 3, 0.7, hidden_dim=16, device="cpu")
    b2.load_state_dict(a.state_dict())
    assert torch.equal(a.act(s), b2.act(s))


def test_registry_lookup_like_main_py(pkg):
    from drl_on_robot_arm_b200 import algo
    from drl_on_robot_arm_b200.config import opt
    agent = getattr(algo, opt.algo)(state_dim=6, action_dim=3, action_bound=0.7, device="cpu")     # main.py:95
    assert type(agent).__name__ == "DADDPG_MLP" and agent.take_action(np.zeros(6, np.float32)).shape == (3,)


@pytest.mark.parametrize("name", NAMES)
def test_update_cycle_describes_the_host_control_flow(pkg, name):
    """update_cycle() = (trains per control-flow cycle, phase): which networks a train() call steps depends only on the
    phase, and repeats every `cycle` calls -- what VectorTrainer keys its CUDA graphs of the updates by"""
    from drl_on_robot_arm_b200 import algo
    torch.manual_seed(0)
    agent = getattr(algo, name)(state_dim=6, action_dim=3, action_bound=0.7, hidden_dim=8, device="cpu")
    batch = {"states": np.random.rand(16, 6).astype(np.float32), "actions": np.random.rand(16, 3).astype(np.float32),
             "rewards": np.random.rand(16).astype(np.float32), "next_states": np.random.rand(16, 6).astype(np.float32),
             "dones": np.zeros(16, np.float32)}
    cycle, phase0 = agent.update_cycle()
    assert cycle == {"TD3_MLP": 3, "DADDPG_MLP": 2}.get(name, 1) and phase0 == 0
    pattern = {}
    for k in range(3 * cycle):
        _, phase = agent.update_cycle()
        before = [[p.detach().clone() for p in l.net.parameters()] for l, _ in agent._learners()]
        agent.train({kk: vv.copy() for kk, vv in batch.items()})
        stepped = tuple(any(not torch.equal(a, b) for a, b in zip(bf, l.net.parameters()))
                        for bf, (l, _) in zip(before, agent._learners()))
        assert pattern.setdefault(phase, stepped) == stepped          # same phase -> same networks stepped
    assert len(pattern) == cycle
