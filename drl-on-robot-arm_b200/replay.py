"""TrajectoryReplay -- the reference's trajectory replay buffers with HER "future" relabelling
(utils/rl_utils.py:91-199), resident in HBM and driven by CUDA kernels (csrc/armsim_replay.cu) through the C-ABI.

Reference usage (main.py:107-138):          here, for the whole lockstep batch:
    traj = Trajectory(state)                    rep.begin(obs)                       # after env.reset()
    traj.store_step(a, s', r, done)             rep.store(a, r, done, final_obs, obs)   # every env.step
    buffer.add_trajectory(traj)                 (implicit: an env's done commits its trajectory)
    buffer.size()                               rep.size()
    b = buffer.sample(B, use_her, thr, ratio)   b = rep.sample(B, use_her, thr, ratio)  -> dict of CUDA tensors
"""
import ctypes as C

import numpy as np

from . import _lib as L

KIND = {"reach": 0, "push": 1, "pick": 1}


class TrajectoryReplay:
    def __init__(self, n_envs, obs_dim, act_dim=3, window=2048, table_cap=None, kind="reach", device=None, seed=0):
        import torch
        if not torch.cuda.is_available():
            raise L.ArmsimError("TrajectoryReplay needs a CUDA device (no CPU fallback)")
        self.torch = torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        cfg = L.ArmReplayConfig()
        cfg.struct_size = C.sizeof(L.ArmReplayConfig)
        cfg.n_envs, cfg.obs_dim, cfg.act_dim, cfg.window = int(n_envs), int(obs_dim), int(act_dim), int(window)
        cfg.table_cap = int(table_cap if table_cap is not None else max(1024, 4 * n_envs))
        cfg.kind = KIND[kind] if isinstance(kind, str) else int(kind)
        cfg.device = self.device.index
        cfg.seed = int(seed)
        self.cfg = cfg
        h = C.c_void_p()
        L.check_replay(L.lib().armsim_replay_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.n, self.obs_dim, self.act_dim = int(n_envs), int(obs_dim), int(act_dim)
        self._out = {}

    def close(self):
        if getattr(self, "h", None):
            L.lib().armsim_replay_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, t, shape):
        tt = self.torch
        if t.dtype != tt.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shape:
            t = t.to(device=self.device, dtype=tt.float32).reshape(shape).contiguous()
        return t

    def begin(self, obs0):
        obs0 = self._f32(obs0, (self.n, self.obs_dim))
        L.check_replay(L.lib().armsim_replay_begin(self.h, obs0.data_ptr(), self._stream()))

    def store(self, action, reward, done, final_obs, obs_out):
        tt = self.torch
        a = self._f32(action, (self.n, self.act_dim))
        r = self._f32(reward, (self.n,))
        fo = self._f32(final_obs, (self.n, self.obs_dim))
        oo = self._f32(obs_out, (self.n, self.obs_dim))
        if done.dtype != tt.uint8 or not done.is_contiguous():
            done = done.to(tt.uint8).contiguous()
        L.check_replay(L.lib().armsim_replay_store(self.h, a.data_ptr(), r.data_ptr(), done.data_ptr(), fo.data_ptr(),
                                                   oo.data_ptr(), self._stream()))

    def _buffers(self, B, with_picks):
        key = (B, with_picks)
        if key not in self._out:
            tt, dev = self.torch, self.device
            self._out[key] = dict(states=tt.empty((B, self.obs_dim), device=dev), actions=tt.empty((B, self.act_dim), device=dev),
                                  next_states=tt.empty((B, self.obs_dim), device=dev), rewards=tt.empty((B,), device=dev),
                                  dones=tt.empty((B,), device=dev),
                                  picks=tt.empty((B, 3), device=dev, dtype=tt.int32) if with_picks else None)
        return self._out[key]

    def sample(self, batch_size, use_her=True, dis_threshold=0.1, her_ratio=0.8, return_picks=False):
        """rl_utils.py:119-152.  Returns a dict of CUDA tensors (states, actions, next_states, rewards, dones[, picks]);
        the tensors are reused by the next call with the same batch size."""
        o = self._buffers(int(batch_size), return_picks)
        L.check_replay(L.lib().armsim_replay_sample(
            self.h, int(batch_size), 1 if use_her else 0, float(dis_threshold), float(her_ratio), o["states"].data_ptr(),
            o["actions"].data_ptr(), o["next_states"].data_ptr(), o["rewards"].data_ptr(), o["dones"].data_ptr(),
            o["picks"].data_ptr() if return_picks else None, self._stream()))
        return o

    def gather(self, slots, steps, goal_steps, dis_threshold=0.1):
        """deterministic variant of sample(): explicit (table slot, step, goal step or -1) picks"""
        tt = self.torch
        B = len(slots)
        ix = [tt.as_tensor(np.asarray(x, np.int32), device=self.device).contiguous() for x in (slots, steps, goal_steps)]
        o = self._buffers(B, False)
        L.check_replay(L.lib().armsim_replay_gather(self.h, B, ix[0].data_ptr(), ix[1].data_ptr(), ix[2].data_ptr(),
                                                    float(dis_threshold), o["states"].data_ptr(), o["actions"].data_ptr(),
                                                    o["next_states"].data_ptr(), o["rewards"].data_ptr(),
                                                    o["dones"].data_ptr(), self._stream()))
        return o

    def info(self):
        v = (C.c_int64 * 4)()
        L.check_replay(L.lib().armsim_replay_info(self.h, v))
        return {"rows": int(v[0]), "trajectories": int(v[1]), "sample_calls": int(v[2]), "empty_samples": int(v[3])}

    def size(self):
        """number of committed trajectories still addressable (rl_utils.py:115); synchronises"""
        return min(self.info()["trajectories"], self.cfg.table_cap)

    def state_blob(self):
        """the whole replay (ring, trajectory table, cursors) as a uint8 array, for checkpoints"""
        nb = int(L.lib().armsim_replay_state_bytes(self.h))
        buf = np.empty(nb, np.uint8)
        L.check_replay(L.lib().armsim_replay_get_state(self.h, buf.ctypes.data, nb))
        return buf

    def load_state_blob(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        L.check_replay(L.lib().armsim_replay_set_state(self.h, buf.ctypes.data, buf.nbytes))

    def table(self, count=None):
        count = self.size() if count is None else int(count)
        env, start, ln = np.zeros(count, np.int32), np.zeros(count, np.int64), np.zeros(count, np.int32)
        L.check_replay(L.lib().armsim_replay_table(self.h, env.ctypes.data, start.ctypes.data, ln.ctypes.data, count))
        return env, start, ln
