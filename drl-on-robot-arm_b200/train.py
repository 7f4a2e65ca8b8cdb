"""VectorTrainer -- the reference's training loops (main.py:77-162 `run`, :165-231 `train_reach_with_TD3`, :449-584
push / pick) re-expressed for N_envs lockstep environments on one GPU (or one shard per GPU).

Reference, per episode of ONE env                         here, per lockstep step of N envs
--------------------------------------------------------  ---------------------------------------------------------
action = agent.take_action(state) + N(0, sigma)  :196-200  actions = agent.act(obs) + N(0, sigma)          (device)
state, reward, done, ok = env.step(action)        :201      env.step(actions, final_obs=True)      (ONE kernel launch)
traj.store_step(...); buffer.add_trajectory(traj) :205-206  replay.store(...)  (an env's done commits its trajectory)
if buffer.size() >= minimal_episodes:             :209      the same gate, on committed trajectories
    n_train x agent.train(buffer.sample(B, her))  :210-212  n_train updates per N finished episodes (one "episode-time");
                                                            `updates_per_episode` restores the reference's ratio, below
every 25 episodes: success_rate; if >= best:      :222-229  every 25 * N finished episodes: the same bookkeeping --
    agent.save(prefix); her_ratio *= 0.75                    save-on-best, her_ratio decay x0.75

The rollout step {actor forward, exploration noise, fused env step, replay store, episode statistics} is captured in
ONE CUDA graph; statistics stay on the device and are read back every `sync_every` steps (one small copy).
Multi-GPU: one process per GPU, each with its env + replay shard; the agents are replicas kept identical by the flat
gradient all-reduce inside agent.train (distributed.GradBucket); logged statistics are summed over ranks.

Two deliberate differences from the reference's single-env loop, both with an option that removes them:
  * update-to-data ratio.  The reference runs n_train = 40 updates after EVERY episode of its one env (main.py:209-212).
    The default here is 40 updates per episode-TIME (N x world episodes), i.e. 1/N of the reference's ratio: the same
    number of updates per second of simulated time, N times the data per update.  `updates_per_episode=k` schedules k
    updates per finished env-episode instead (k = opt.n_train reproduces the reference's ratio exactly; the learning
    curve is then comparable with visdata/** episode for episode).
  * replay horizon.  The reference keeps the last 1,000,000 TRAJECTORIES (rl_utils.py:110-111, main.py:179), i.e.
    everything a run ever produces.  Here the replay is a ring of `window` lockstep rows in HBM (make_trainer(window=));
    window = ceil(total_steps) keeps everything too: at 91 B per env-row for push a 180 GB B200 holds 1.9e9 env-steps.
"""
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .config import opt
from .distributed import allreduce_scalars, register_graph_holder
from .metrics import MetricsSink


class VectorTrainer:
    def __init__(self, env, agent, replay, noise_std=None, clip_actions=False, batch_size=None, n_train=None,
                 minimal_episodes=None, use_her=True, her_ratio=None, dis_threshold=0.1, window_episodes=None, avg_window=10,
                 sync_every=16, metrics=None, save_prefix=None, use_cuda_graph=True, max_updates_per_sync=None,
                 fused_bookkeeping=True, graph_updates=None, updates_per_episode=None, fused_policy=True):
        self.env, self.agent, self.replay = env, agent, replay
        self.device = env.device
        self.n = env.n
        self.action_bound = float(agent.action_bound)
        # main.py:200 adds N(0, 1 * opt.gamma) (the DISCOUNT doubles as the noise scale -- reference quirk, kept as default)
        self.noise_std = float(opt.gamma if noise_std is None else noise_std)
        self.clip_actions = bool(clip_actions)                  # run() clips to +-action_bound (main.py:116-117)
        self.batch_size = int(batch_size or opt.batch_size)
        self.n_train = int(n_train or opt.n_train)
        # None: n_train per episode-time (N x world episodes); k: k updates per finished env-episode (the reference's
        # update-to-data ratio for k = opt.n_train, main.py:209-212)
        self.updates_per_episode = None if updates_per_episode is None else float(updates_per_episode)
        self.minimal_episodes = int(opt.minimal_episodes if minimal_episodes is None else minimal_episodes)
        self.use_her = bool(use_her)
        self.her_ratio = float(opt.her_ratio if her_ratio is None else her_ratio)
        self.dis_threshold = float(dis_threshold)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.window_episodes = int(window_episodes or 25 * self.n * self.world)      # main.py:222 `% 25`, per env
        self.avg_window = int(avg_window)
        self.sync_every = int(sync_every)
        self.metrics = metrics if metrics is not None else MetricsSink()
        self.save_prefix = save_prefix
        self.use_cuda_graph = bool(use_cuda_graph)
        self.max_updates_per_sync = max_updates_per_sync
        self.fused_bookkeeping = bool(fused_bookkeeping)       # armsim_explore / armsim_track_episodes vs torch elementwise ops
        # single-MLP actors (DDPG, TD3): forward + exploration noise as ONE launch reading the nn.Linear parameters in place
        # (armsim_policy_act) instead of ~10 PyTorch kernels; double-actor agents keep agent.act()
        pol = agent.acting_policy() if hasattr(agent, "acting_policy") else None
        self._policy = pol if (fused_policy and self.fused_bookkeeping and pol is not None and
                               env.policy_supported(pol, env.obs_dim, env.act_dim)) else None
        # CUDA-graph the learning updates (single GPU; the multi-GPU path keeps eager updates around its all-reduce)
        if graph_updates is None:
            # multi-GPU: the NCCL all-reduce is recorded inside the update graph; train_updates() never leaves a replay in
            # flight (see there), which is what keeps the communicator's collectives in one order on every rank.
            # ARMSIM_GRAPH_NCCL_UPDATES=0 falls back to eager updates around the all-reduce.
            graph_updates = self.use_cuda_graph and (self.world == 1 or os.environ.get("ARMSIM_GRAPH_NCCL_UPDATES", "1") != "0")
        self.graph_updates = bool(graph_updates)
        self._update_graphs, self._eager_updates = {}, 0
        if self.world > 1:
            register_graph_holder(self)
        dev = self.device
        # device-side episode statistics: [episodes finished, successes, sum of finished returns]
        self._stats = torch.zeros(3, device=dev, dtype=torch.float64)
        self.ep_return = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.actions = torch.zeros((self.n, env.act_dim), device=dev)
        self.obs = None
        self._graph = None
        self._chunk_graph, self._chunk_len = None, 0           # `chunk` consecutive rollout steps as ONE graph launch
        self._side = torch.cuda.Stream(device=dev)             # replay stores of a chunk graph (overlap the next policy launch)
        self._actions2 = (self.actions, torch.zeros_like(self.actions))
        self._stream = torch.cuda.Stream(device=dev)
        # host-side bookkeeping
        self.steps = 0
        self.updates = 0
        self.episodes_seen = 0.0          # global (all ranks), as of the last sync
        self.update_credit = 0.0
        self.window_acc = np.zeros(3)     # episodes, successes, return sum of the open success-rate window
        self.best_rate = 0.0
        self.returns_log = []
        self._last_stats = np.zeros(3)
        self._replay_ready = False        # every rank holds at least one committed trajectory

    # ------------------------------------------------------------------ rollout
    def reset(self):
        with torch.cuda.stream(self._stream):
            self.obs = self.env.reset()
            self.replay.begin(self.obs)
            self.ep_return.zero_()
            if self.fused_bookkeeping:
                self.env.set_state(L.F_EP_RETURN, np.zeros(self.n, np.float32))
        self._stream.synchronize()

    def _rollout_body(self, actions=None, side=None, first=True):
        """one lockstep step.  `side` (chunk graphs only): the replay store of THIS step is issued on that stream so that
        it overlaps the NEXT step's policy launch (which only reads obs and writes the other action buffer); the caller
        alternates `actions` between two buffers and joins the side stream at the end of the chunk."""
        env = self.env
        actions = self.actions if actions is None else actions
        main = torch.cuda.current_stream(self.device)
        track = self.fused_bookkeeping and int(env.cfg.mode) == 0    # IK-teleport mode: episode statistics ride in the step launch

        def store(obs, rew, done):
            if side is None:
                self.replay.store(actions, rew, done, env.final_obs, obs)
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.replay.store(actions, rew, done, env.final_obs, obs)

        if self._policy is not None or self.fused_bookkeeping:
            clip = self.action_bound if self.clip_actions else 0.0
            if self._policy is not None:
                # 3 launches per lockstep step: {actor MLP + exploration noise}, {fused env step + episode statistics}, replay store
                env.policy_act(self._policy, env.obs, self.noise_std, clip, out=actions)
            else:
                # main.py:200 (+ :117 clip) and :202-207 in the engine's kernels instead of ~16 elementwise torch ops
                env.explore(self.agent.act(env.obs), self.noise_std, clip, out=actions)
            if side is not None and not first:
                main.wait_stream(side)          # the previous step's store has read the buffers this step overwrites
            obs, rew, done, succ = env.step(actions, final_obs=True, track=track)
            store(obs, rew, done)
            if not track:
                env.track_episodes(rew, done, succ)
            return
        a = self.agent.act(env.obs)
        a = a + torch.randn_like(a) * self.noise_std                              # main.py:200
        if self.clip_actions:
            a = a.clamp(-self.action_bound, self.action_bound)                    # main.py:117
        self.actions.copy_(a)
        obs, rew, done, succ = env.step(self.actions, final_obs=True)
        self.replay.store(self.actions, rew, done, env.final_obs, obs)
        self.ep_return += rew
        d = done.to(torch.float32)
        fin = torch.stack([d.sum(), succ.to(torch.float32).sum(), (self.ep_return * d).sum()]).to(torch.float64)
        self._stats += fin
        self.ep_return *= (1.0 - d)

    @property
    def stats(self):
        """[episodes finished, successes, sum of finished returns] so far (float64 tensor on the host side of a sync)"""
        if self.fused_bookkeeping:
            self._stream.synchronize()
            return torch.from_numpy(self.env.episode_stats())
        return self._stats

    def rollout_step(self):
        """one lockstep step of all envs (one CUDA-graph replay once captured)"""
        if self.obs is None:
            self.reset()
        with torch.cuda.stream(self._stream):
            if not self.use_cuda_graph:
                self._rollout_body()
            elif self._graph is None:
                for _ in range(3):                                                # warm-up outside capture (these steps count)
                    self._rollout_body()
                    self.steps += 1
                self._stream.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=self._stream):
                    self._rollout_body()
                self._graph.replay()                                              # capture does not execute: run it once
            else:
                self._graph.replay()
        self.steps += 1

    def rollout_chunk(self, max_steps):
        """as many lockstep steps as fit before the next statistics read-back (at most `max_steps`), issued as ONE graph
        launch when a whole chunk of min(sync_every, 64) steps fits: the host then pays one cudaGraphLaunch per chunk
        instead of one per step (a 4096-env reach step is ~20 us of GPU work, about what a Python-issued launch costs).
        Same kernels in the same order as rollout_step(), so the results are identical."""
        chunk = min(self.sync_every, 64)
        if (not self.use_cuda_graph or self._graph is None or chunk < 2 or max_steps < chunk or
                self.steps % self.sync_every != 0):
            self.rollout_step()
            return 1
        with torch.cuda.stream(self._stream):
            if self._chunk_graph is None or self._chunk_len != chunk:
                self._stream.synchronize()
                g = torch.cuda.CUDAGraph()
                overlap = self.fused_bookkeeping or self._policy is not None
                with torch.cuda.graph(g, stream=self._stream):
                    for k in range(chunk):
                        if overlap:
                            self._rollout_body(self._actions2[k % 2], self._side, first=(k == 0))
                        else:
                            self._rollout_body()
                    if overlap:
                        self._stream.wait_stream(self._side)
                self._chunk_graph, self._chunk_len = g, chunk
            self._chunk_graph.replay()
        self.steps += chunk
        return chunk

    # ------------------------------------------------------------------ learning
    def train_updates(self, k):
        """k updates = k x {replay.sample, agent.train}.  On one GPU the updates run as replays of a CUDA graph that
        holds one control-flow cycle of the agent (TD3: policy_freq = 3 updates, the third with the actor step; DADDPG:
        2) -- about a hundred small kernels per update leave one launch each instead of one Python call each.  A graph
        is keyed by the agent's phase and by the sampler arguments (her_ratio decays during a run)."""
        k = int(k)
        self.updates += k
        with torch.cuda.stream(self._stream):
            while k > 0:
                cyc, phase = self.agent.update_cycle()
                if not self.graph_updates or k < cyc or self._eager_updates < 3 * cyc:
                    self._one_update()
                    self._eager_updates += 1
                    k -= 1
                    continue
                key = (cyc, phase, self.batch_size, self.use_her, self.dis_threshold, self.her_ratio)
                g = self._update_graphs.get(key)
                if g is None:
                    if len(self._update_graphs) >= 64:
                        self.release_graphs()
                    self._stream.synchronize()
                    it0 = self.agent.total_it
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._stream):
                        for _ in range(cyc):
                            self._one_update()
                    self._update_graphs[key] = (g, self.agent.total_it - it0)
                    self.agent.total_it = it0                       # capture records, it does not run
                    g = self._update_graphs[key]
                g[0].replay()
                self.agent.total_it += g[1]
                k -= cyc
            if self.world > 1 and self.graph_updates:
                # a recorded all-reduce runs on THIS stream, eager collectives on the process group's own one: never leave
                # a replay in flight where the caller may issue another collective (NCCL needs one order on every rank)
                self._stream.synchronize()

    def release_graphs(self):
        """drop the recorded update cycles (they hold NCCL work when world > 1; see distributed.shutdown)"""
        self._stream.synchronize()
        for g, _ in self._update_graphs.values():
            g.reset()
        self._update_graphs.clear()
        self._eager_updates = 0

    def _one_update(self):
        batch = self.replay.sample(self.batch_size, self.use_her, self.dis_threshold, self.her_ratio)
        self.agent.train(batch, sync=False)

    def _sync(self):
        """read the device statistics, all-reduce them over ranks, run the updates that became due, log windows"""
        self._stream.synchronize()
        local = self.stats.cpu().numpy().copy()
        delta = local - self._last_stats
        self._last_stats = local
        if self.world > 1:
            g = allreduce_scalars({"e": delta[0], "s": delta[1], "r": delta[2]})
            delta = np.array([g["e"], g["s"], g["r"]])
        self.episodes_seen += delta[0]
        self.window_acc += delta
        if delta[0] > 0:
            mean_ret = delta[2] / delta[0]
            self.returns_log.append(mean_ret)
            self.metrics.plot("return", mean_ret, x=self.episodes_seen)                               # main.py:207
            self.metrics.plot("avg_return", float(np.mean(self.returns_log[-self.avg_window:])), x=self.episodes_seen)  # :220
        # n_train updates per N finished episodes, once minimal_episodes trajectories exist (main.py:209-212) -- and once
        # EVERY rank can sample: each rank draws from its own replay shard, so a rank that has committed nothing yet
        # (its envs' first episodes all still running, or all longer than the ring) must hold everybody back
        if not self._replay_ready and self.episodes_seen >= self.minimal_episodes:
            have = float(self.replay.size())
            if self.world > 1:
                have = allreduce_scalars({"t": have}, op="min")["t"]
            self._replay_ready = have >= 1.0
        if self.episodes_seen >= self.minimal_episodes and self._replay_ready:
            if self.updates_per_episode is not None:
                self.update_credit += delta[0] * self.updates_per_episode / float(self.n_train)
            else:
                self.update_credit += delta[0] / float(self.n * self.world)
            k = int(self.update_credit * self.n_train)
            if self.max_updates_per_sync is not None:
                k = min(k, int(self.max_updates_per_sync))
            k -= k % self.agent.update_cycle()[0]          # whole control-flow cycles (= graph replays); the rest stays credited
            if k > 0:
                self.update_credit -= k / float(self.n_train)
                self.train_updates(k)
        if self.window_acc[0] >= self.window_episodes:                                               # main.py:222-229
            rate = self.window_acc[1] / self.window_acc[0]
            self.metrics.plot("success_rate", rate, x=self.episodes_seen)
            if rate >= self.best_rate:
                if self.save_prefix and self.rank == 0:
                    os.makedirs(os.path.dirname(os.path.abspath(self.save_prefix)), exist_ok=True)
                    self.agent.save("%s%s" % (self.save_prefix, rate))
                self.best_rate = rate
                self.her_ratio *= 0.75
            self.window_acc[:] = 0.0

    def run(self, total_steps):
        """advance every env by total_steps lockstep steps, learning as the reference's cadence dictates"""
        end = self.steps + int(total_steps)
        while self.steps < end:
            self.rollout_chunk(end - self.steps)
            if self.steps % self.sync_every == 0 or self.steps >= end:
                self._sync()
        return {"steps": self.steps, "env_steps": self.steps * self.n * self.world, "updates": self.updates,
                "episodes": self.episodes_seen, "success_rate": self.metrics.last("success_rate"),
                "avg_return": self.metrics.last("avg_return"), "her_ratio": self.her_ratio}

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self, include_replay=True):
        env_state = {f: self.env.get_state(f) for f in L.STATE_FIELDS}
        self._stream.synchronize()
        return {"agent": self.agent.state_dict(), "env": env_state, "env_obs": self.env.obs.cpu(),
                "replay": self.replay.state_blob() if include_replay else None,
                "stats": self.stats.cpu(), "ep_return": self.ep_return.cpu(),
                "host": {k: getattr(self, k) for k in ("steps", "updates", "episodes_seen", "update_credit", "best_rate", "her_ratio")},
                "window_acc": self.window_acc.copy(), "last_stats": self._last_stats.copy(), "returns_log": list(self.returns_log),
                "torch_rng": torch.get_rng_state(), "cuda_rng": torch.cuda.get_rng_state(self.device)}

    def load_state_dict(self, sd):
        self._update_graphs.clear()                 # recorded updates point at the optimizer state being replaced
        self._eager_updates = 0
        self.agent.load_state_dict(sd["agent"])
        for f, v in sd["env"].items():
            self.env.set_state(f, v)
        self.env.obs.copy_(sd["env_obs"].to(self.device))
        self.obs = self.env.obs
        if sd["replay"] is not None:
            self.replay.load_state_blob(sd["replay"])
        self._stats.copy_(sd["stats"].to(self.device))
        if self.fused_bookkeeping:
            self.env.set_episode_stats(sd["stats"].numpy())
        self.ep_return.copy_(sd["ep_return"].to(self.device))
        for k, v in sd["host"].items():
            setattr(self, k, v)
        self.window_acc, self._last_stats = sd["window_acc"].copy(), sd["last_stats"].copy()
        self.returns_log = list(sd["returns_log"])
        torch.set_rng_state(sd["torch_rng"])
        torch.cuda.set_rng_state(sd["cuda_rng"], self.device)


def save_checkpoint(path, trainer, include_replay=True):
    """everything needed to resume a run (the reference only ever saves actor/critic weights, TD3_mlp.py:163-168)"""
    tmp = path + ".tmp"
    torch.save(trainer.state_dict(include_replay), tmp)
    os.replace(tmp, path)


def load_checkpoint(path, trainer):
    trainer.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))


TASK_DEFAULTS = {   # state_dim, action_bound (main.py:87 high[0]+0.3 for reach; :457 / :526 0.4 for push / pick), replay kind,
    # exploration noise std: N(0, 1 * opt.gamma) for reach (main.py:200), N(0, action_bound * opt.gamma) = 0.392 for
    # push / pick (main.py:484, :553) -- the DISCOUNT doubles as the noise scale in the reference (quirk, kept)
    "reach": (6, 0.7, "reach", 1.0 * opt.gamma), "push": (9, 0.4, "push", 0.4 * opt.gamma), "pick": (9, 0.4, "push", 0.4 * opt.gamma),
}


def make_trainer(task="reach", algo="TD3_MLP", n_envs=4096, device=None, seed=None, window=1024, distributed=None, **kw):
    """getattr(envs, opt.env) / getattr(algo, opt.algo) of main.py:83,95 for the batched engine"""
    from . import algo as A
    from .distributed import make_sharded_env
    from .replay import TrajectoryReplay
    seed = opt.random_seed if seed is None else seed
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    distributed = (world > 1) if distributed is None else distributed
    state_dim, bound, kind, noise = TASK_DEFAULTS[task]
    if kw.get("noise_std") is None:
        kw["noise_std"] = noise
    torch.manual_seed(seed)                                                       # main.py:175-177 (same init on every rank)
    env = make_sharded_env(task, n_envs * world, rank=rank, world=world, device=device, seed=seed, auto_reset=True)
    agent = getattr(A, algo)(state_dim=state_dim, action_dim=3, action_bound=bound, device=env.device, distributed=distributed)
    if distributed:
        agent.broadcast_parameters(0)
    torch.manual_seed(seed + 1000 * (rank + 1))                                   # exploration noise differs per rank
    # trajectory-table slots: every episode the ring can still hold, down to 32-step episodes
    table_cap = int(min(max(1024, 4 * env.n, env.n * (int(window) // 32)), 1 << 26))
    replay = TrajectoryReplay(n_envs=env.n, obs_dim=env.obs_dim, act_dim=3, window=window, table_cap=table_cap, kind=kind,
                              device=env.device, seed=seed + rank)
    return VectorTrainer(env, agent, replay, **kw)
