"""GPU: the HBM trajectory replay (csrc/armsim_replay.cu, through the C-ABI) against the reference's own batches
(tests/golden/her_*.npz, made from /root/reference/utils/rl_utils.py by tools/gen_her_golden.py) and against a host
shadow of a real rollout.  Integer / index work bit-exact; relabelled f32 data bit-exact for reach (the reference's
reach trajectories are float32); push compared at f32 resolution (its reference obs are float64)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _load_fixture_into_ring(pkg, torch, g, kind, window):
    """env e replays fixture trajectory e in lockstep; afterwards a filler episode that never terminates"""
    states, actions, rewards, dones, lengths = (g[k] for k in ("states", "actions", "rewards", "dones", "lengths"))
    T, Lmax, O = states.shape[0], actions.shape[1], states.shape[2]
    rep = pkg.TrajectoryReplay(n_envs=T, obs_dim=O, act_dim=3, window=window, table_cap=64, kind=kind, seed=1)
    dev = rep.device
    rep.begin(torch.tensor(states[:, 0], dtype=torch.float32, device=dev))
    for k in range(Lmax + 3):
        live = k < lengths
        a = np.where(live[:, None], actions[:, min(k, Lmax - 1)], 0.25).astype(np.float32)
        r = np.where(live, rewards[:, min(k, Lmax - 1)], -7.0).astype(np.float32)
        d = np.where(live, dones[:, min(k, Lmax - 1)], 0).astype(np.uint8)
        nxt = np.where(live[:, None], states[:, min(k + 1, Lmax)], 9.0).astype(np.float32)
        out = nxt.copy()
        out[d != 0] = -5.0                                     # "post-reset" obs of the filler episode
        rep.store(torch.from_numpy(a).to(dev), torch.from_numpy(r).to(dev), torch.from_numpy(d).to(dev),
                  torch.from_numpy(nxt).to(dev), torch.from_numpy(out).to(dev))
    return rep


@pytest.mark.parametrize("kind", ["reach", "push"])
def test_gather_reproduces_reference_batches(pkg, torch_cuda, kind):
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "her_%s.npz" % kind))
    rep = _load_fixture_into_ring(pkg, torch, g, kind, window=64)
    T = len(g["lengths"])
    info = rep.info()
    assert info["trajectories"] == T and rep.size() == T            # every fixture episode committed, fillers not
    env, start, ln = rep.table(T)
    assert sorted(env.tolist()) == list(range(T)) and np.array_equal(ln[np.argsort(env)], g["lengths"])
    assert np.all(start == 0)
    slot_of = np.argsort(env)                                       # trajectory (= env) id -> table slot
    picks = g["picks"]
    out = rep.gather(slot_of[picks[:, 0]], picks[:, 1], picks[:, 2], float(g["dis_threshold"]))
    st, nx = out["states"].cpu().numpy(), out["next_states"].cpu().numpy()
    assert np.array_equal(out["actions"].cpu().numpy(), g["out_actions"])
    if kind == "reach":
        assert np.array_equal(st, g["out_states"].astype(np.float32)) and np.array_equal(nx, g["out_next_states"].astype(np.float32))
        assert np.array_equal(out["rewards"].cpu().numpy(), g["out_rewards"].astype(np.float32))
        assert np.array_equal(out["dones"].cpu().numpy(), g["out_dones"].astype(np.float32))
    else:
        assert np.abs(st - g["out_states"]).max() <= 1e-6 and np.abs(nx - g["out_next_states"]).max() <= 1e-6
        dis = np.linalg.norm(g["out_next_states"][:, :3] - g["out_states"][:, 3:6], axis=1)
        clear = (picks[:, 2] < 0) | (np.abs(dis - 0.1) > 1e-6)
        assert np.abs(out["rewards"].cpu().numpy() - g["out_rewards"])[clear].max() <= 1e-5      # env rewards stored as f32
        assert np.array_equal(out["dones"].cpu().numpy()[clear], g["out_dones"].astype(np.float32)[clear])
    rep.close()


def test_sampler_follows_the_reference_distribution(pkg, torch_cuda):
    """uniform over committed trajectories, uniform step, P(HER) = her_ratio, goal uniform in (step, len];
    Philox: same seed -> same picks, consecutive calls differ"""
    torch = torch_cuda
    g = np.load(os.path.join(GOLD, "her_reach.npz"))
    reps = [_load_fixture_into_ring(pkg, torch, g, "reach", window=64) for _ in range(2)]
    T, lengths = len(g["lengths"]), g["lengths"]
    B = 1 << 17
    outs = [r.sample(B, True, 0.1, 0.8, return_picks=True) for r in reps]
    p0, p1 = outs[0]["picks"].cpu().numpy(), outs[1]["picks"].cpu().numpy()
    assert np.array_equal(p0, p1)                                    # same seed, same call number
    p2 = reps[0].sample(B, True, 0.1, 0.8, return_picks=True)["picks"].cpu().numpy()
    assert not np.array_equal(p0, p2)                                # the device-side call counter advanced
    assert reps[0].info()["sample_calls"] == 2
    env, _, ln = reps[0].table(T)
    slot, step, goal = p0[:, 0], p0[:, 1], p0[:, 2]
    assert slot.min() >= 0 and slot.max() < T
    L = ln[slot]
    assert np.all(step >= 0) and np.all(step < L)
    her = goal >= 0
    assert abs(her.mean() - 0.8) < 0.01
    assert np.all(goal[her] > step[her]) and np.all(goal[her] <= L[her])
    cnt = np.bincount(slot, minlength=T)
    assert np.abs(cnt / B - 1.0 / T).max() < 0.15 / T                # uniform over trajectories (B/T ~ 3500 per bin)
    # step uniform given the trajectory: mean relative position ~ (L-1)/(2L)
    long = L >= 20
    assert abs((step[long] / L[long]).mean() - ((L[long] - 1) / (2.0 * L[long])).mean()) < 0.01
    # goal uniform in (step, L]
    frac = (goal[her] - step[her] - 1) / np.maximum(L[her] - step[her], 1)
    assert abs(frac.mean() - ((L[her] - step[her] - 1) / (2.0 * np.maximum(L[her] - step[her], 1))).mean()) < 0.01
    no_her = reps[0].sample(4096, False, 0.1, 0.8, return_picks=True)
    assert (no_her["picks"][:, 2] < 0).all()
    for r in reps:
        r.close()


def test_rollout_store_sample_under_cuda_graph(pkg, torch_cuda):
    """{env step (with final_obs) -> replay store -> replay sample} captured ONCE in a CUDA graph and replayed: the
    device-side cursors advance, committed trajectories match the env's own done flags, and non-HER samples are real
    consecutive transitions of the rollout"""
    torch = torch_cuda
    n, T = 512, 90
    env = pkg.BatchedArmEnv("reach", n_envs=n, device="cuda:0", seed=5, auto_reset=True, max_steps=24)
    rep = pkg.TrajectoryReplay(n_envs=n, obs_dim=6, act_dim=3, window=128, table_cap=8192, kind="reach", device="cuda:0", seed=3)
    obs = env.reset()
    rep.begin(obs)
    acts = (torch.rand((n, 3), device="cuda:0") * 1.4 - 0.7)
    side = torch.cuda.Stream()
    hist = []
    with torch.cuda.stream(side):
        def body():
            o, r, d, s = env.step(acts, final_obs=True)
            rep.store(acts, r, d, env.final_obs, o)
            return rep.sample(256, True, 0.1, 0.8, return_picks=True)
        for _ in range(3):
            body()                                                   # warm-up (eager); rows 0..2
            hist.append((env.final_obs.cpu().numpy().copy(), env.obs.cpu().numpy().copy(), env.done.cpu().numpy().copy(), env.reward.cpu().numpy().copy()))
        side.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=side):
            batch = body()
        for _ in range(T):
            gph.replay()
            side.synchronize()
            hist.append((env.final_obs.cpu().numpy().copy(), env.obs.cpu().numpy().copy(), env.done.cpu().numpy().copy(), env.reward.cpu().numpy().copy()))
    torch.cuda.synchronize()
    info = rep.info()
    n_done = int(sum(h[2].sum() for h in hist))
    assert info["rows"] == T + 3 and info["trajectories"] == n_done and info["sample_calls"] == T + 3
    assert n_done >= 3 * n                                            # 93 steps, episodes <= 25 steps
    # last captured sample: check the non-HER transitions against the host shadow of the rollout
    picks = batch["picks"].cpu().numpy()
    env_id, start, ln = rep.table(min(n_done, 8192))
    st, nx, rw = batch["states"].cpu().numpy(), batch["next_states"].cpu().numpy(), batch["rewards"].cpu().numpy()
    checked = 0
    for b in range(256):
        slot, step, goal = picks[b]
        if goal >= 0:
            assert np.array_equal(st[b, 3:], nx[b, 3:])               # both carry the relabelled goal
            continue
        e, row = env_id[slot], start[slot] + step
        prev_obs = hist[row - 1][1][e] if row > 0 else None
        if prev_obs is not None:
            assert np.array_equal(st[b], prev_obs)                    # state = what the env showed before the step
        assert np.array_equal(nx[b], hist[row][0][e]) and rw[b] == hist[row][3][e]
        checked += 1
    assert checked > 20
    env.close(); rep.close()


def test_sample_from_an_empty_replay_is_flagged_not_garbage(pkg, torch_cuda):
    """the reference raises on an empty buffer (random.sample on an empty deque, rl_utils.py:126); the device sampler
    cannot raise from inside a CUDA graph: it emits a zero batch with done = 1 (no bootstrap) and counts the call"""
    torch = torch_cuda
    rep = pkg.TrajectoryReplay(n_envs=8, obs_dim=6, act_dim=3, window=16, table_cap=32, kind="reach", seed=1)
    rep.begin(torch.zeros((8, 6), device=rep.device))
    for k in range(3):                                                # rows stored, nothing committed
        z = torch.full((8, 6), float(k), device=rep.device)
        rep.store(torch.ones((8, 3), device=rep.device), torch.ones(8, device=rep.device),
                  torch.zeros(8, dtype=torch.uint8, device=rep.device), z, z)
    out = rep.sample(64, True, 0.1, 0.8, return_picks=True)
    assert float(out["states"].abs().max()) == 0 and float(out["next_states"].abs().max()) == 0
    assert float(out["rewards"].abs().max()) == 0 and bool((out["dones"] == 1).all()) and bool((out["picks"] == -1).all())
    info = rep.info()
    assert info["trajectories"] == 0 and info["empty_samples"] == 1 and info["sample_calls"] == 1 and rep.size() == 0
    rep.close()


def test_sampling_stays_uniform_when_most_of_the_table_is_stale(pkg, torch_cuda):
    """with a short ring and a long table most table entries point at overwritten rows; draws must still be
    uniform over the INTACT trajectories (rl_utils.py:126 draws uniformly over what the buffer holds) -- no pile-up on
    the newest entry"""
    torch = torch_cuda
    n, W, ep_len, T = 64, 48, 10, 400
    rep = pkg.TrajectoryReplay(n_envs=n, obs_dim=6, act_dim=3, window=W, table_cap=8192, kind="reach", seed=5)
    dev = rep.device
    rep.begin(torch.zeros((n, 6), device=dev))
    phase = np.arange(n) % ep_len                                     # envs finish on different rows
    for k in range(T):
        d = torch.from_numpy((((k + phase) % ep_len) == ep_len - 1).astype(np.uint8)).to(dev)
        o = torch.full((n, 6), float(k), device=dev)
        rep.store(torch.zeros((n, 3), device=dev), torch.zeros(n, device=dev), d, o, o)
    info = rep.info()
    ntraj = info["trajectories"]
    env, start, ln = rep.table(ntraj)
    now = info["rows"]
    intact = (start - 1 >= now - W) & (ln > 0)
    assert intact.sum() >= n * 2 and intact.sum() < 0.2 * ntraj      # most of the table is stale
    B = 1 << 16
    picks = rep.sample(B, False, 0.1, 0.8, return_picks=True)["picks"].cpu().numpy()
    slot = picks[:, 0]
    assert intact[slot].all()                                         # never a trajectory the ring has overwritten
    cnt = np.bincount(slot, minlength=ntraj)[intact]
    mean = B / intact.sum()
    assert cnt.min() > 0.7 * mean and cnt.max() < 1.3 * mean, (cnt.min(), cnt.max(), mean)
    assert rep.info()["empty_samples"] == 0
    rep.close()
