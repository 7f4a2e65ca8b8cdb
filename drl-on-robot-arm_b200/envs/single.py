"""Drop-in single-env classes with the reference's names, constructor, reset/step/close signatures and return
conventions (envs/__init__.py:1-3; main.py:83 `getattr(envs, opt.env)(is_render=True, is_good_view=False)`).

Each instance is an N_envs = 1 view of the batched CUDA engine going through the host-buffer C-ABI call
(armsim_step_host), so the reference training loops (main.py:111-128) run unchanged.
"""
import time

import numpy as np

from .. import _lib as L
from ..config import opt
from ..spaces import Box, Env
from .batched import ArmSimHandle


class _ArmEnvBase(Env):
    _task = "reach"
    _obs_low_high = None

    def __init__(self, is_render=False, is_good_view=False, device=0, robot="kuka_iiwa", seed=None):
        # is_render: the reference opens the Bullet GUI (rl_reach_env.py:59-62); there is no renderer here.
        self.is_render = is_render
        self.is_good_view = is_good_view
        self.max_steps_one_episode = opt.max_steps_one_episode
        self.x_low_obs, self.x_high_obs = 0.2, 0.7          # rl_reach_env.py:65-70
        self.y_low_obs, self.y_high_obs = -0.3, 0.3
        self.z_low_obs, self.z_high_obs = 0, 0.55
        self.x_low_action, self.x_high_action = -0.4, 0.4   # rl_reach_env.py:73-78
        self.y_low_action, self.y_high_action = -0.4, 0.4
        self.z_low_action, self.z_high_action = -0.6, 0.3
        self.action_space = Box(low=np.array([self.x_low_action, self.y_low_action, self.z_low_action]),
                                high=np.array([self.x_high_action, self.y_high_action, self.z_high_action]),
                                dtype=np.float32)
        self.observation_space = self._make_observation_space()
        self.step_counter = 0
        self.init_joint_positions = [0.006418, 0.413184, -0.011401, -1.589317, 0.005379, 1.137684, -0.006539]
        self._device, self._robot = device, robot
        self._sim = None
        self._seed = opt.random_seed if seed is None else seed
        self.seed(seed)
        self.reset()

    # -- overridden per task
    def _make_observation_space(self):
        lo = [self.x_low_obs, self.y_low_obs, self.z_low_obs]
        hi = [self.x_high_obs, self.y_high_obs, self.z_high_obs]
        return Box(low=np.array(lo), high=np.array(hi), dtype=np.float32)

    def _overrides(self):
        return {}

    def _make_sim(self):
        if self._sim is not None:
            self._sim.close()
        self._sim = ArmSimHandle(task=self._task, n_envs=1, device=self._device, robot=self._robot, seed=self._seed,
                                 auto_reset=False, **self._overrides())

    def seed(self, seed=None):
        """rl_reach_env.py:127-130.  A new seed re-keys the Philox stream of the goal / cube sampler."""
        if seed is not None:
            self._seed = int(seed)
            self._sim = None
        return [seed]

    def reset(self):
        self.step_counter = 0
        self.terminated = False
        if self._sim is None:
            self._make_sim()       # armsim_create performs the first reset (episode 0)
            obs = self._first_obs()
        else:
            obs = self._sim.reset_host()
        return self._format_obs(obs[0])

    def _first_obs(self):
        # observation of the state armsim_create left behind, without consuming another episode
        sim = self._sim
        q = sim.get_state(L.F_Q)
        pos, _ = sim.fk(q)
        goal = sim.get_state(L.F_GOAL)
        if sim.obs_dim == 6:
            return np.hstack([pos, goal])
        if sim.obs_dim == 3:
            return pos
        return np.hstack([pos, sim.get_state(L.F_CUBE_POS), goal])

    def _format_obs(self, o):
        return np.asarray(o, dtype=np.float32).copy()

    def _step_raw(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 3)
        obs, rew, done, succ = self._sim.step_host(a)
        if self.is_good_view:
            time.sleep(0.05)                                # rl_reach_env.py:261-262
        self.step_counter += 1
        self.terminated = bool(done[0])
        return obs[0], float(rew[0]), bool(done[0]), bool(succ[0])

    def close(self):
        if self._sim is not None:
            self._sim.close()
            self._sim = None


class RLReachEnv(_ArmEnvBase):
    """envs/rl_reach_env.py:38.  obs f32[6] = [ee, goal]; step -> (obs, reward, done, is_success: bool)."""
    _task = "reach"

    def _make_observation_space(self):
        lo = [self.x_low_obs, self.y_low_obs, self.z_low_obs] * 2          # rl_reach_env.py:93-96
        hi = [self.x_high_obs, self.y_high_obs, self.z_high_obs] * 2
        return Box(low=np.array(lo), high=np.array(hi), dtype=np.float32)

    def _overrides(self):
        return dict(dv=opt.reach_ctr, reach_dis=opt.reach_dis, max_steps=opt.max_steps_one_episode)

    def step(self, action):
        obs, r, done, succ = self._step_raw(action)
        self.is_success = succ
        return self._format_obs(obs), r, done, succ                         # rl_reach_env.py:319


class KukaReachEnv(_ArmEnvBase):
    """envs/kuka_reach_env.py:54.  obs f32[3] = ee; step -> (obs, reward, done, distance: float)."""
    _task = "kuka_reach"
    max_steps_one_episode = 1000

    def step(self, action):
        obs, r, done, _ = self._step_raw(action)
        goal = self._sim.get_state(L.F_GOAL)[0]
        self.distance = float(np.linalg.norm(obs.astype(np.float64) - goal.astype(np.float64)))
        return self._format_obs(obs), r, done, self.distance                 # kuka_reach_env.py:304-305


class RLPushEnv(_ArmEnvBase):
    """envs/rl_push_env.py:43.  obs f64[9] = [ee, cube, target]; step -> (obs, reward, done, {'is_success': f32})."""
    _task = "push"
    distance_threshold = 0.05

    def _overrides(self):
        return dict(max_steps=opt.max_steps_one_episode)

    def _format_obs(self, o):
        return np.asarray(o, dtype=np.float64).copy()                        # hstack of f32 + f64 -> f64 (:308)

    def step(self, action):
        obs, r, done, succ = self._step_raw(action)
        info = {'is_success': np.float32(1.0 if succ else 0.0)}             # rl_push_env.py:430-432
        return self._format_obs(obs), r, done, info


class RLPickEnv(RLPushEnv):
    """envs/rl_pick_env.py:43."""
    _task = "pick"
    gripper_length = 0.257

    def _make_observation_space(self):
        lo = [self.x_low_obs, self.y_low_obs, self.z_low_obs + self.gripper_length]    # rl_pick_env.py:95-96
        hi = [self.x_high_obs, self.y_high_obs, self.z_high_obs + self.gripper_length]
        return Box(low=np.array(lo), high=np.array(hi), dtype=np.float32)
