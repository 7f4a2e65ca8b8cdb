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


def test_explore_noise_kernel(pkg, torch_cuda):
    """armsim_explore (main.py:200, :116-117): exact pass-through at sigma = 0, N(0, sigma) otherwise, fresh draws on every
    call (also from a CUDA-graph replay), clip honoured, and a noise stream keyed by the GLOBAL env id (shard-invariant)"""
    torch = torch_cuda
    n = 65536
    env = pkg.BatchedArmEnv("reach", n_envs=n, device="cuda:0", seed=5)
    mu = torch.rand((n, 3), device="cuda") - 0.5
    assert torch.equal(env.explore(mu, 0.0), mu)
    a1 = env.explore(mu, 0.98)
    a2 = env.explore(mu, 0.98)
    z = ((a1 - mu) / 0.98).double().cpu().numpy().ravel()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((z ** 4).mean() - 3.0) < 0.1 and abs((z ** 3).mean()) < 0.05          # normal: kurtosis 3, no skew
    assert abs(np.corrcoef(z[0::3], z[1::3])[0, 1]) < 0.01                           # components independent
    assert not torch.equal(a1, a2)
    assert abs(np.corrcoef(z, ((a2 - mu) / 0.98).double().cpu().numpy().ravel())[0, 1]) < 0.01
    c = env.explore(mu, 5.0, clip=0.7)
    assert float(c.abs().max()) <= 0.7 and float((c.abs() == 0.7).float().mean()) > 0.5
    # graph replay: the draw counter lives on the device, so every replay is a new draw
    out = torch.empty_like(mu)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        env.explore(mu, 1.0, out=out)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            env.explore(mu, 1.0, out=out)
        g.replay(); st.synchronize(); r1 = out.clone()
        g.replay(); st.synchronize(); r2 = out.clone()
    assert not torch.equal(r1, r2)
    # shard invariance: envs [n/2, n) of this handle == a handle created with env_id_offset = n/2 (same draw count)
    lo = pkg.BatchedArmEnv("reach", n_envs=n // 2, device="cuda:0", seed=5, env_id_offset=n // 2)
    whole = pkg.BatchedArmEnv("reach", n_envs=n, device="cuda:0", seed=5)
    assert torch.equal(whole.explore(mu, 1.0)[n // 2:], lo.explore(mu[n // 2:].contiguous(), 1.0))
    for e in (env, lo, whole):
        e.close()


@pytest.mark.parametrize("task,n", [("reach", 4096), ("push", 1000), ("pick", 33)])
def test_policy_act_matches_the_torch_actor(pkg, torch_cuda, task, n):
    """armsim_policy_act (PolicyNet.forward of algo/TD3/net_mlp.py:29-40 as one launch, fp32 FFMA): equals the PyTorch
    module on the same parameters to 5e-6 (summation order only), ragged batch sizes included; with exploration its noise
    is bit-identical to armsim_explore applied to the bare output (same Philox stream and draw counter); picks up
    in-place parameter updates (it reads the nn.Linear storage) and replays from a CUDA graph with fresh noise."""
    torch = torch_cuda
    from drl_on_robot_arm_b200.algo.nets import PolicyNet
    torch.manual_seed(11)
    env = pkg.BatchedArmEnv(task, n_envs=n, device="cuda:0", seed=5)
    twin = pkg.BatchedArmEnv(task, n_envs=n, device="cuda:0", seed=5)
    S = env.obs_dim
    net = PolicyNet(S, 256, 3, 0.7).to("cuda:0")
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(3.0)                                  # spread the pre-activations over tanh's range
    assert env.policy_supported(net, S, 3)
    obs = (torch.rand((n, S), device="cuda") - 0.5) * 2.0
    with torch.no_grad():
        want = net(obs)
    got = env.policy_act(net, obs)
    assert float((got - want).abs().max()) <= 5e-6, float((got - want).abs().max())
    assert float(want.abs().max()) > 0.5 and float(want.abs().min()) < 0.05      # saturated and linear outputs both present
    noisy = env.policy_act(net, obs, noise_std=0.4, clip=0.7)
    assert torch.equal(noisy, twin.explore(got, 0.4, clip=0.7))
    noisy2 = env.policy_act(net, obs, noise_std=0.4, clip=0.7)
    assert not torch.equal(noisy, noisy2) and torch.equal(noisy2, twin.explore(got, 0.4, clip=0.7))
    with torch.no_grad():
        net.fc3.bias.add_(0.25)
        want2 = net(obs)
    assert float((env.policy_act(net, obs) - want2).abs().max()) <= 5e-6
    out = torch.empty_like(got)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        env.policy_act(net, obs, noise_std=1.0, out=out)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            env.policy_act(net, obs, noise_std=1.0, out=out)
        g.replay(); st.synchronize(); r1 = out.clone()
        g.replay(); st.synchronize(); r2 = out.clone()
    assert not torch.equal(r1, r2)
    z = ((r1 - want2) / 1.0).double().cpu().numpy().ravel()
    if n >= 1000:
        assert abs(z.mean()) < 0.08 and abs(z.std() - 1.0) < 0.08
    bad = PolicyNet(S, 128, 3, 0.7).to("cuda:0")
    assert not env.policy_supported(bad, S, 3)
    with pytest.raises(Exception):
        env.policy_act(bad, obs)
    for e in (env, twin):
        e.close()


def test_fused_policy_rollout_matches_torch_actor_rollout(pkg, torch_cuda):
    """the trainer's rollout with the one-launch policy == the rollout through agent.act() + explore (noise off: the
    action differs by fp32 summation order only, so joint angles agree to 1e-4 rad over 40 steps and the episode
    bookkeeping is identical)"""
    res = []
    for fused in (True, False):
        tr = _mk(pkg, n=256, minimal_episodes=10 ** 9, noise_std=0.0, fused_policy=fused)
        assert (tr._policy is not None) == fused
        tr.run(40)
        res.append((tr.env.get_state(0).copy(), tr.stats.cpu().numpy().copy()))
        tr.env.close(); tr.replay.close()
    assert np.abs(res[0][0] - res[1][0]).max() <= 1e-4
    assert np.array_equal(res[0][1][:2], res[1][1][:2])


@pytest.mark.parametrize("fused", [False, True])
def test_track_episodes_matches_host_bookkeeping(pkg, torch_cuda, fused):
    """armsim_track_episodes (main.py:202-207, :222-229), as its own launch and folded into the step launch
    (armsim_step_tracked), against the same bookkeeping done in numpy float64"""
    torch = torch_cuda
    n = 1000
    env = pkg.BatchedArmEnv("reach", n_envs=n, device="cuda:0", seed=2, auto_reset=True, max_steps=17, reach_dis=0.05)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    ret = np.zeros(n); want = np.zeros(3)
    for k in range(120):
        a = (torch.rand((n, 3), device="cuda", generator=gen) * 2 - 1) * 0.7
        if fused:
            obs, rew, done, succ = env.step(a, final_obs=True if k % 2 else None, track=True)
        else:
            obs, rew, done, succ = env.step(a)
            env.track_episodes()
        r, d, s = rew.double().cpu().numpy(), done.cpu().numpy().astype(bool), succ.cpu().numpy().astype(bool)
        ret += r
        want += [d.sum(), (d & s).sum(), ret[d].sum()]
        ret[d] = 0.0
    got = env.episode_stats()
    assert got[0] == want[0] and got[1] == want[1] and got[0] >= 6 * n
    assert abs(got[2] - want[2]) <= 1e-4 * got[0]                  # 2^-16 fixed point + f32 running returns
    assert np.allclose(env.get_state(pkg._lib.F_EP_RETURN), ret, atol=1e-3)
    env.set_episode_stats([3, 1, -2.5])
    assert list(env.episode_stats()) == [3.0, 1.0, -2.5]
    env.close()


def test_fused_and_torch_bookkeeping_agree(pkg, torch_cuda):
    """VectorTrainer with the engine's explore / track_episodes kernels == the same loop in elementwise torch ops when
    the exploration noise is off (the two paths draw their noise from different generators)"""
    torch = torch_cuda
    res = []
    for fused in (True, False):
        tr = _mk(pkg, n=200, minimal_episodes=10 ** 9, noise_std=0.0, fused_bookkeeping=fused, fused_policy=False)
        tr.env.close()
        from drl_on_robot_arm_b200.distributed import make_sharded_env
        tr.env = make_sharded_env("reach", 200, device="cuda:0", seed=3, auto_reset=True, max_steps=12)
        tr.run(60)
        res.append((tr.env.get_state(0).copy(), tr.stats.cpu().numpy().copy(), tr.replay.info()))
        tr.env.close(); tr.replay.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][2] == res[1][2]
    assert np.array_equal(res[0][1][:2], res[1][1][:2]) and res[0][1][0] >= 200 * 4
    assert abs(res[0][1][2] - res[1][1][2]) <= 1e-4 * res[0][1][0]


@pytest.mark.parametrize("algo", ["DDPG_MLP", "DADDPG_MLP"])
def test_graphed_updates_equal_eager_updates(pkg, torch_cuda, algo):
    """train_updates as CUDA-graph replays of one control-flow cycle == the same updates issued one by one (agents without
    random draws in train(): bit-identical weights, targets and update counters)"""
    torch = torch_cuda
    res = []
    for graphed in (True, False):
        tr = _mk(pkg, n=128, algo=algo, noise_std=0.3, graph_updates=graphed)
        tr.env.close()
        from drl_on_robot_arm_b200.distributed import make_sharded_env
        tr.env = make_sharded_env("reach", 128, device="cuda:0", seed=3, auto_reset=True, max_steps=10)
        tr.run(120)
        assert tr.updates >= 200
        assert bool(tr._update_graphs) == graphed
        res.append(([p.detach().clone() for l, _ in tr.agent._learners() for p in list(l.net.parameters()) + list(l.target.parameters())],
                    tr.updates, tr.agent.total_it))
        tr.env.close(); tr.replay.close()
    assert res[0][1:] == res[1][1:]
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)


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
