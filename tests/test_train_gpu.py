"""GPU: the vectorised rollout + training loop (train.py, N1), checkpoint / resume (N3) and the metrics sink."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _mk(pkg, tmp_path=None, graph=True, algo="TD3_MLP", task="reach", n=256, **kw):
    from drl_on_robot_arm_b200 import metrics, train
    sink = metrics.MetricsSink(str(tmp_path) if tmp_path else None, "t")
    return train.make_trainer(task=task, algo=algo, n_envs=n, device="cuda:0", seed=3, window=256, metrics=sink,
                              use_cuda_graph=graph, window_episodes=4 * n, sync_every=8, **kw)


@pytest.mark.parametrize("algo,task", [("TD3_MLP", "reach"), ("DARC_MLP", "reach"), ("DADDPG_MLP", "reach"), ("DATD3_MLP", "pick"),
                                       ("DDPG_MLP", "push")])
def test_trainer_runs_the_reference_cadence(pkg, torch_cuda, tmp_path, algo, task):
    torch = torch_cuda
    tr = _mk(pkg, tmp_path, algo=algo, task=task, n=128)
    tr.env.close()
    # short episodes so that windows close quickly
    from drl_on_robot_arm_b200.distributed import make_sharded_env
    tr.env = make_sharded_env(task, 128, device="cuda:0", seed=3, auto_reset=True, max_steps=20)
    out = tr.run(200)
    assert out["steps"] == 200 and out["env_steps"] == 200 * 128
    ep = out["episodes"]
    assert ep >= 128 * 9                                       # >= 9 timeouts of 21 steps per env
    # n_train (40) updates per n_envs finished episodes once minimal_episodes exist (main.py:209-212)
    assert abs(out["updates"] - 40 * ep / 128) <= 40 + 1
    assert out["success_rate"] is not None and 0.0 <= out["success_rate"] <= 1.0
    assert tr.her_ratio < 0.8                                   # at least one window closed with rate >= best (0): x0.75
    for name in ("return", "avg_return", "success_rate"):
        assert len(tr.metrics.series[name]) >= 1
    paths = tr.metrics.export_csv()
    assert open(paths[0]).readline().strip() == "xData,yData"   # main.py:615-621
    assert all(torch.isfinite(p).all() for l, _ in tr.agent._learners() for p in l.net.parameters())
    info = tr.replay.info()
    assert info["rows"] == 200 and info["trajectories"] == int(ep)
    tr.env.close(); tr.replay.close()


def test_graph_and_eager_rollouts_agree(pkg, torch_cuda):
    """the CUDA-graph replay of the rollout step does exactly what the eager step does (no learning: pure rollout)"""
    torch = torch_cuda
    res = []
    for graph in (True, False):
        tr = _mk(pkg, graph=graph, n=200, minimal_episodes=10 ** 9)
        torch.manual_seed(11)
        tr.run(40)
        res.append((tr.env.get_state(0).copy(), tr.stats.cpu().numpy().copy(), tr.replay.info()))
        tr.env.close(); tr.replay.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]


def test_checkpoint_resume_continues_identically(pkg, torch_cuda, tmp_path):
    """save at step 48, keep going to 80; a fresh trainer resumed from the file reaches the same env state, agent
    weights, replay cursors and statistics (eager rollouts on both sides: a resumed run re-captures its graph)"""
    torch = torch_cuda
    from drl_on_robot_arm_b200 import train
    a = _mk(pkg, graph=False, n=128)
    a.env.close()
    from drl_on_robot_arm_b200.distributed import make_sharded_env
    a.env = make_sharded_env("reach", 128, device="cuda:0", seed=3, auto_reset=True, max_steps=15)
    a.run(48)
    ck = str(tmp_path / "run.ckpt")
    train.save_checkpoint(ck, a)
    a.run(32)
    b = _mk(pkg, graph=False, n=128)
    b.env.close()
    b.env = make_sharded_env("reach", 128, device="cuda:0", seed=3, auto_reset=True, max_steps=15)    # same config (the Philox key
    b.run(21)                                                  # is configuration); scramble every piece of state before loading
    train.load_checkpoint(ck, b)
    assert b.steps == 48
    b.run(32)
    assert b.steps == a.steps == 80 and b.updates == a.updates and b.her_ratio == a.her_ratio
    assert np.array_equal(a.env.get_state(0), b.env.get_state(0))                 # joint angles of every env
    assert np.array_equal(a.env.get_state(2), b.env.get_state(2))                 # goals
    assert a.replay.info() == b.replay.info()
    assert np.array_equal(a.stats.cpu().numpy(), b.stats.cpu().numpy())
    for (la, _), (lb, _) in zip(a.agent._learners(), b.agent._learners()):
        for pa, pb in zip(la.net.parameters(), lb.net.parameters()):
            assert torch.equal(pa, pb)
    for t in (a, b):
        t.env.close(); t.replay.close()


def test_two_gpu_replicas_stay_identical(pkg, torch_cuda):
    """torchrun x2 (NCCL): env shards differ, gradient buckets are all-reduced, replicas end bit-identical"""
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "dist_train_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DIST_TRAIN_OK" in res.stdout
