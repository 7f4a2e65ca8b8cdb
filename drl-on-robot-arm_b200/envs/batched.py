"""BatchedArmEnv -- N_envs reference environments stepped by ONE fused CUDA launch.

Host-side mirror of the reference Env API for a whole batch: reset() -> obs [N, obs_dim], step(action [N, 3]) ->
(obs, reward [N], done [N], success [N]).  Tensors live on the GPU; the call goes straight through the C-ABI
(armsim_step) with raw device pointers on torch's current stream -- PyTorch is only the allocator / stream owner.
"""
import ctypes as C

import numpy as np

from .. import _lib as L

_TASKS = {"reach": L.TASK_REACH, "push": L.TASK_PUSH, "pick": L.TASK_PICK, "kuka_reach": L.TASK_KUKA_REACH}
_ROBOTS = {"kuka_iiwa": L.ROBOT_KUKA_IIWA, "diana_s1": L.ROBOT_DIANA_S1}
_MODES = {"ik_teleport": L.MODE_IK_TELEPORT, "torque": L.MODE_TORQUE}


class ArmSimHandle:
    """Thin RAII wrapper of an ArmSim* (create / destroy / state io); no torch needed."""

    def __init__(self, task="reach", n_envs=1, device=0, robot="kuka_iiwa", seed=0, env_id_offset=0, auto_reset=False,
                 chain=None, mode="ik_teleport", **overrides):
        """mode = "ik_teleport": the reference's step (action [n,3] EE servo).  mode = "torque": joint torques
        [n,7] through articulated-body forward dynamics (obs = task obs + q[7] + qd[7])."""
        task_id = _TASKS[task] if isinstance(task, str) else int(task)
        cfg = L.default_config(task_id, **overrides)
        cfg.mode = _MODES[mode] if isinstance(mode, str) else int(mode)
        cfg.n_envs = int(n_envs)
        cfg.device = int(device)
        cfg.seed = int(seed)
        cfg.env_id_offset = int(env_id_offset)
        cfg.auto_reset = 1 if auto_reset else 0
        self._chain = None
        if chain is not None:
            self._chain = chain          # keep alive: cfg holds a raw pointer
            cfg.robot = L.ROBOT_CUSTOM
            cfg.custom_chain = C.pointer(chain)
        else:
            cfg.robot = _ROBOTS[robot] if isinstance(robot, str) else int(robot)
        self.cfg = cfg
        h = C.c_void_p()
        L.check(L.lib().armsim_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.n = int(n_envs)
        self.obs_dim = L.lib().armsim_obs_dim(h)
        self.act_dim = L.lib().armsim_action_dim(h)
        self.task = task_id

    def host_server(self, idle_us=20000):
        """keep ONE kernel resident for the host-buffer step (armsim_host_server): step_host / step_pinned / step_async then
        cost a release store + a doorbell poll instead of a launch.  idle_us = 0 turns it off.  Any other call on this
        handle makes the resident kernel leave first; it also leaves by itself after idle_us without a step."""
        L.check(L.lib().armsim_host_server(self.h, int(idle_us)))

    @property
    def mapping(self):
        """resolved thread mapping of the step kernel ("lane": one CUDA lane per arm; see include/armsim.h)"""
        return {L.MAP_LANE: "lane"}[L.lib().armsim_mapping(self.h)]

    def close(self):
        if getattr(self, "h", None):
            self._hb = None              # the views die with the pinned block
            self._pinned_call = None
            L.lib().armsim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state injection / read-back (parity tests, checkpointing)
    def set_state(self, field, arr):
        a = np.ascontiguousarray(arr, np.int32 if field in L.INT_FIELDS else np.float32)
        L.check(L.lib().armsim_set_state(self.h, field, a.ctypes.data, a.nbytes))

    def get_state(self, field):
        w = L.FIELD_WIDTH[field]
        a = np.zeros((self.n, w) if w > 1 else (self.n,), np.int32 if field in L.INT_FIELDS else np.float32)
        L.check(L.lib().armsim_get_state(self.h, field, a.ctypes.data, a.nbytes))
        return a

    def fk(self, q):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 7)
        pos = np.zeros((q.shape[0], 3), np.float32)
        rot = np.zeros((q.shape[0], 9), np.float32)
        L.check(L.lib().armsim_fk_host(self.h, q.ctypes.data, q.shape[0], pos.ctypes.data, rot.ctypes.data))
        return pos, rot.reshape(-1, 3, 3)

    @property
    def launch_count(self):
        return int(L.lib().armsim_launch_count(self.h))

    # ---- host-buffer path (numpy in / numpy out; H2D + launch + D2H inside the call)
    def host_buffers(self):
        """numpy views (action [n,3], obs [n,obs_dim], reward [n], done [n], success [n]) of the handle's pinned,
        device-mapped I/O block (armsim_host_buffers).  step_host(action_view, out=the four output views) is the
        copy-free form of the call: write the actions into `action`, call, read the results in place."""
        if getattr(self, "_hb", None) is None:
            ptr = [C.c_void_p() for _ in range(5)]
            L.check(L.lib().armsim_host_buffers(self.h, *[C.byref(p) for p in ptr]))
            n = self.n

            def view(p, ctype, dtype, shape):
                cnt = int(np.prod(shape))
                return np.frombuffer((ctype * cnt).from_address(p.value), dtype=dtype).reshape(shape)
            self._hb = (view(ptr[0], C.c_float, np.float32, (n, self.act_dim)),
                        view(ptr[1], C.c_float, np.float32, (n, self.obs_dim)),
                        view(ptr[2], C.c_float, np.float32, (n,)),
                        view(ptr[3], C.c_uint8, np.uint8, (n,)),
                        view(ptr[4], C.c_uint8, np.uint8, (n,)))
        return self._hb

    def step_pinned(self):
        """armsim_step_host on the handle's own pinned block with every argument pre-bound: the leanest host-side call
        (no per-call ctypes conversions).  Actions are read from host_buffers()[0]; results land in host_buffers()[1:]."""
        call = getattr(self, "_pinned_call", None)
        if call is None:
            hb = self.host_buffers()
            fn = L.lib().armsim_step_host
            args = (self.h,) + tuple(C.c_void_p(b.ctypes.data) for b in hb)
            chk = L.check

            def call():
                rc = fn(*args)
                if rc:
                    chk(rc)
            self._pinned_call = call
        call()
        return self._hb[1:]

    def episode_stats(self):
        """{episodes finished, successes, sum of finished returns} accumulated by track_episodes (float64[3], synchronous)"""
        out = (C.c_double * 3)()
        L.check(L.lib().armsim_episode_stats(self.h, out))
        return np.array(out[:], np.float64)

    def set_episode_stats(self, stats):
        a = (C.c_double * 3)(*[float(x) for x in stats])
        L.check(L.lib().armsim_set_episode_stats(self.h, a))

    def step_async(self):
        """gym.vector's step_async on the pinned block: actions are read from host_buffers()[0]; returns at once"""
        calls = getattr(self, "_async_calls", None)
        if calls is None:
            hb = self.host_buffers()
            lib = L.lib()
            h, a = self.h, C.c_void_p(hb[0].ctypes.data)
            outs = tuple(C.c_void_p(b.ctypes.data) for b in hb[1:])
            fa, fw, chk = lib.armsim_step_host_async, lib.armsim_step_host_wait, L.check

            def submit():
                rc = fa(h, a)
                if rc:
                    chk(rc)

            def wait():
                rc = fw(h, *outs)
                if rc:
                    chk(rc)
            calls = self._async_calls = (submit, wait)
        calls[0]()

    def step_wait(self):
        """gym.vector's step_wait: blocks until the step issued by step_async has landed in host_buffers()[1:]"""
        self._async_calls[1]()
        return self._hb[1:]

    def step_host(self, action, out=None):
        a = action if (isinstance(action, np.ndarray) and action.dtype == np.float32 and action.flags.c_contiguous) \
            else np.ascontiguousarray(action, np.float32)
        if a.shape != (self.n, self.act_dim):
            raise ValueError("action must be [%d, %d], got %s" % (self.n, self.act_dim, a.shape))
        if out is None:
            out = (np.empty((self.n, self.obs_dim), np.float32), np.empty(self.n, np.float32),
                   np.empty(self.n, np.uint8), np.empty(self.n, np.uint8))
        obs, rew, done, succ = out
        L.check(L.lib().armsim_step_host(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                         succ.ctypes.data))
        return obs, rew, done, succ

    def reset_host(self, mask=None, obs=None):
        if obs is None:
            obs = np.zeros((self.n, self.obs_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        L.check(L.lib().armsim_reset_host(self.h, None if m is None else m.ctypes.data, obs.ctypes.data))
        return obs


class BatchedArmEnv(ArmSimHandle):
    """Vectorised env on one GPU: torch CUDA tensors in, torch CUDA tensors out, zero host round trips."""

    def __init__(self, task="reach", n_envs=4096, device=None, **kw):
        import torch
        if not torch.cuda.is_available():
            raise L.ArmsimError("BatchedArmEnv needs a CUDA device (no CPU fallback)")
        self.torch = torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        super().__init__(task=task, n_envs=n_envs, device=idx, **kw)
        self.device = torch.device("cuda", idx)
        n, od = self.n, self.obs_dim
        self.obs = torch.empty((n, od), dtype=torch.float32, device=self.device)
        self.reward = torch.empty((n,), dtype=torch.float32, device=self.device)
        self.done = torch.empty((n,), dtype=torch.uint8, device=self.device)
        self.success = torch.empty((n,), dtype=torch.uint8, device=self.device)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask=None):
        m = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
            m = mask.data_ptr()
        L.check(L.lib().armsim_reset(self.h, m, self.obs.data_ptr(), self._stream()))
        return self.obs

    def explore(self, actor_out, noise_std, clip=0.0, out=None):
        """main.py:200 / :116-117 on the device: out = actor_out + noise_std * N(0,1) (clipped to +-clip when clip > 0);
        in-kernel Philox noise keyed by (seed, global env id, draws so far), so CUDA-graph replays draw fresh noise"""
        t = self.torch
        if actor_out.dtype != t.float32 or not actor_out.is_contiguous():
            actor_out = actor_out.to(t.float32).contiguous()
        if out is None:
            out = t.empty_like(actor_out)
        L.check(L.lib().armsim_explore(self.h, actor_out.data_ptr(), float(noise_std), float(clip), out.data_ptr(), self._stream()))
        return out

    def policy_act(self, policy, obs=None, noise_std=None, clip=0.0, out=None):
        """The acting policy in ONE launch (armsim_policy_act): out = policy(obs) [+ noise_std * N(0,1), clipped], read
        straight from the nn.Linear parameters of `policy` (algo.nets.PolicyNet: fc1, fc2, fc3, action_bound; hidden 256,
        fp32, on this device).  noise_std=None returns the bare policy output.  Replaces policy(obs) + explore()."""
        t = self.torch
        obs = self.obs if obs is None else obs
        if obs.dtype != t.float32 or not obs.is_contiguous():
            obs = obs.to(t.float32).contiguous()
        if out is None:
            out = t.empty((self.n, self.act_dim), device=self.device, dtype=t.float32)
        ps = [policy.fc1.weight, policy.fc1.bias, policy.fc2.weight, policy.fc2.bias, policy.fc3.weight, policy.fc3.bias]
        for p_ in ps:
            if p_.dtype != t.float32 or not p_.is_contiguous() or p_.device != self.device:
                raise ValueError("policy_act: parameters must be contiguous fp32 tensors on %s" % self.device)
        L.check(L.lib().armsim_policy_act(self.h, obs.data_ptr(), *[p_.data_ptr() for p_ in ps], int(policy.fc2.weight.shape[0]),
                                          float(policy.action_bound), -1.0 if noise_std is None else float(noise_std), float(clip),
                                          out.data_ptr(), self._stream()))
        return out

    @staticmethod
    def policy_supported(policy, obs_dim, act_dim):
        """can armsim_policy_act run this module? (3-layer PolicyNet, hidden 256, obs <= 16, actions <= 4, fp32)"""
        try:
            return (policy.fc1.weight.shape == (256, obs_dim) and policy.fc2.weight.shape == (256, 256) and
                    policy.fc3.weight.shape == (act_dim, 256) and obs_dim <= 16 and act_dim <= 4 and
                    str(policy.fc1.weight.dtype) == "torch.float32")
        except AttributeError:
            return False

    def track_episodes(self, reward=None, done=None, success=None):
        """main.py:202-207, :222-229 on the device: running returns + {episodes, successes, return sum} (episode_stats())"""
        r = self.reward if reward is None else reward
        d = self.done if done is None else done
        s = self.success if success is None else success
        L.check(L.lib().armsim_track_episodes(self.h, r.data_ptr(), d.data_ptr(), s.data_ptr(), self._stream()))

    def step(self, action, out=None, final_obs=None, track=False):
        """action: float32 CUDA tensor [N, act_dim].  Returns (obs, reward, done, success) views of the env's own
        output buffers (overwritten by the next step) unless `out` = (obs, reward, done, success) tensors is given.
        final_obs (optional f32 [N, obs_dim] tensor, or True for the env's own buffer `self.final_obs`) receives this
        step's observation before any in-kernel auto-reset (what a replay buffer stores as next_state).
        track=True folds track_episodes() of this step into the same launch (armsim_step_tracked; IK-teleport mode)."""
        t = self.torch
        if action.dtype != t.float32 or not action.is_cuda or not action.is_contiguous() or \
                tuple(action.shape) != (self.n, self.act_dim):
            action = action.to(device=self.device, dtype=t.float32).reshape(self.n, self.act_dim).contiguous()
        obs, rew, done, succ = out if out is not None else (self.obs, self.reward, self.done, self.success)
        if track:
            fo = None
            if final_obs is True:
                if getattr(self, "final_obs", None) is None:
                    self.final_obs = t.empty_like(self.obs)
                fo = self.final_obs
            elif final_obs is not None:
                fo = final_obs
            L.check(L.lib().armsim_step_tracked(self.h, action.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                                succ.data_ptr(), fo.data_ptr() if fo is not None else None, self._stream()))
        elif final_obs is None:
            L.check(L.lib().armsim_step(self.h, action.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                        succ.data_ptr(), self._stream()))
        else:
            if final_obs is True:
                if getattr(self, "final_obs", None) is None:
                    self.final_obs = t.empty_like(self.obs)
                final_obs = self.final_obs
            L.check(L.lib().armsim_step_ex(self.h, action.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                           succ.data_ptr(), final_obs.data_ptr(), self._stream()))
        return obs, rew, done, succ
