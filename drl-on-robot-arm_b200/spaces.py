"""Minimal stand-ins for gym.spaces.Box / gym.Env (gym is not a dependency; the reference only reads
`.shape`, `.high`, `.low` and `.sample()` -- main.py:85-87, rl_reach_env.py:334-347)."""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.dtype = np.dtype(dtype)
        self.shape = self.low.shape
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s)" % (self.low, self.high, self.dtype)


class Env:
    """gym.Env surface the reference envs expose."""
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return [seed]

    def render(self, mode="human"):
        return None

    def close(self):
        pass
