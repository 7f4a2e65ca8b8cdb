"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under drl-on-robot-arm_b200/ does.
Parity status: FK / clip / reward / done pinned by the reference's golden artefacts; the Bullet-internal IK and
stepSimulation arithmetic is restated and PARITY UNPINNED (see armsim_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "liboracle.so")

NJ = 7
TASK_REACH, TASK_PUSH, TASK_PICK, TASK_KUKA_REACH = 0, 1, 2, 3
ROBOT_KUKA, ROBOT_DIANA, ROBOT_CUSTOM = 0, 1, 2
MODE_IK_TELEPORT, MODE_TORQUE = 0, 1
(F_Q, F_QD, F_GOAL, F_STEP, F_EPISODE, F_CUBE_POS, F_CUBE_QUAT, F_CUBE_LINVEL, F_CUBE_ANGVEL, F_LAST_DIST, F_GRIP,
 F_IK_ITERS) = range(12)
FIELD_WIDTH = {F_Q: 7, F_QD: 7, F_GOAL: 3, F_STEP: 1, F_EPISODE: 1, F_CUBE_POS: 3, F_CUBE_QUAT: 4, F_CUBE_LINVEL: 3,
               F_CUBE_ANGVEL: 3, F_LAST_DIST: 1, F_GRIP: 1, F_IK_ITERS: 1}
INT_FIELDS = (F_STEP, F_EPISODE, F_IK_ITERS)


class ArmsimChain(C.Structure):
    _fields_ = [("base_xyz", C.c_double * 3), ("base_rpy", C.c_double * 3),
                ("xyz", (C.c_double * 3) * NJ), ("rpy", (C.c_double * 3) * NJ),
                ("lower", C.c_double * NJ), ("upper", C.c_double * NJ), ("effort", C.c_double * NJ),
                ("velocity", C.c_double * NJ), ("damping", C.c_double * NJ),
                ("mass", C.c_double * NJ), ("com", (C.c_double * 3) * NJ), ("inertia", (C.c_double * 6) * NJ)]


class ArmsimConfig(C.Structure):
    """Mirror of include/armsim.h ArmsimConfig (kept in sync by tests/test_abi.py)."""
    _fields_ = [("struct_size", C.c_int32), ("task", C.c_int32), ("robot", C.c_int32), ("mode", C.c_int32),
                ("mapping", C.c_int32), ("n_envs", C.c_int32), ("device", C.c_int32), ("auto_reset", C.c_int32),
                ("seed", C.c_uint64), ("env_id_offset", C.c_uint64),
                ("dv", C.c_double), ("reach_dis", C.c_double), ("max_steps", C.c_int32),
                ("ws_lo", C.c_double * 3), ("ws_hi", C.c_double * 3),
                ("goal_lo", C.c_double * 3), ("goal_hi", C.c_double * 3),
                ("target_rpy", C.c_double * 3), ("init_q", C.c_double * NJ),
                ("ik_damping", C.c_double), ("ik_max_iters", C.c_int32), ("ik_residual", C.c_double),
                ("clamp_joint_limits", C.c_int32), ("reserved", C.c_int32 * 7),
                ("sim_dt", C.c_double), ("gravity", C.c_double * 3),
                ("custom_chain", C.POINTER(ArmsimChain))]


def build(force=False):
    """gcc-compile the oracle (plain C).  Building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("armsim_oracle.c", "armsim_oracle.h", "cube_model.h", "aba_model.h")]
    srcs += [os.path.join(ROOT, "include", f) for f in ("armsim.h", "armsim_defaults.h", "armsim_robot_models.h")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_reset_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        L.orc_fk.argtypes = [C.c_int32, dp, dp, dp, dp, dp]
        L.orc_jacobian.argtypes = [C.c_int32, dp, dp]
        L.orc_ik.argtypes = [C.c_int32, dp, dp, dp, C.c_double, C.c_int, C.c_double, dp, dp]
        L.orc_quat_from_euler.argtypes = [dp, dp]
        L.orc_create.argtypes = [C.POINTER(ArmsimConfig)]
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_step_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5
        L.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        L.orc_get_state_f64.argtypes = [C.c_void_p, C.c_int32, dp, C.c_size_t]
        L.orc_grip_distance.argtypes = [C.c_void_p, dp]
        L.orc_pgs_stats.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.orc_pool_create.argtypes = [C.c_void_p, C.c_int]
        L.orc_pool_create.restype = C.c_void_p
        L.orc_pool_destroy.argtypes = [C.c_void_p]
        L.orc_step_mt.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_run_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_obs_dim.argtypes = [C.c_void_p]
        L.orc_obs_dim.restype = C.c_int32
        L.orc_action_dim.argtypes = [C.c_void_p]
        L.orc_action_dim.restype = C.c_int32
        L.orc_aba.argtypes = [C.c_int32, dp, dp, dp, dp]
        L.orc_rnea.argtypes = [C.c_int32, dp, dp, dp, C.c_int, dp]
        L.orc_dense_fd.argtypes = [C.c_int32, dp, dp, dp, dp, dp]
        if hasattr(L, "orc_default_config"):
            L.orc_default_config.argtypes = [C.c_int32, C.POINTER(ArmsimConfig)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_config(task, n_envs=1, **kw):
    cfg = ArmsimConfig()
    rc = lib().orc_default_config(task, C.byref(cfg))
    assert rc == 0
    cfg.n_envs = n_envs
    for k, v in kw.items():
        cur = getattr(cfg, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(cfg, k, v)
    return cfg


def fk(q, robot=ROBOT_KUKA):
    q = np.ascontiguousarray(q, dtype=np.float64)
    pos, rot, org, axes = np.zeros(3), np.zeros(9), np.zeros(21), np.zeros(21)
    rc = lib().orc_fk(robot, _dp(q), _dp(pos), _dp(rot), _dp(org), _dp(axes))
    assert rc == 0
    return pos, rot.reshape(3, 3), org.reshape(7, 3), axes.reshape(7, 3)


def jacobian(q, robot=ROBOT_KUKA):
    q = np.ascontiguousarray(q, dtype=np.float64)
    J = np.zeros(42)
    assert lib().orc_jacobian(robot, _dp(q), _dp(J)) == 0
    return J.reshape(6, 7)


def quat_from_euler(rpy):
    rpy = np.ascontiguousarray(rpy, dtype=np.float64)
    q = np.zeros(4)
    lib().orc_quat_from_euler(_dp(rpy), _dp(q))
    return q


def ik(q, target_pos, target_quat, damping=1e-5, max_iters=20, residual=1e-4, robot=ROBOT_KUKA):
    q = np.ascontiguousarray(q, dtype=np.float64)
    tp = np.ascontiguousarray(target_pos, dtype=np.float64)
    tq = np.ascontiguousarray(target_quat, dtype=np.float64)
    out, diff = np.zeros(7), np.zeros(1)
    its = lib().orc_ik(robot, _dp(q), _dp(tp), _dp(tq), damping, max_iters, residual, _dp(out), _dp(diff))
    return out, its, float(diff[0])


def aba(q, qd, tau, robot=ROBOT_KUKA):
    """articulated-body forward dynamics (gravity (0,0,-10)): qdd"""
    q, qd, tau = (np.ascontiguousarray(x, dtype=np.float64) for x in (q, qd, tau))
    out = np.zeros(7)
    assert lib().orc_aba(robot, _dp(q), _dp(qd), _dp(tau), _dp(out)) == 0
    return out


def rnea(q, qd, qdd, robot=ROBOT_KUKA, gravity=True):
    """recursive Newton-Euler inverse dynamics: tau"""
    q, qd, qdd = (np.ascontiguousarray(x, dtype=np.float64) for x in (q, qd, qdd))
    out = np.zeros(7)
    assert lib().orc_rnea(robot, _dp(q), _dp(qd), _dp(qdd), 1 if gravity else 0, _dp(out)) == 0
    return out


def dense_fd(q, qd, tau, robot=ROBOT_KUKA):
    """qdd = M^-1 (tau - h) with M, h from RNEA; returns (qdd, M)"""
    q, qd, tau = (np.ascontiguousarray(x, dtype=np.float64) for x in (q, qd, tau))
    out, M = np.zeros(7), np.zeros(49)
    assert lib().orc_dense_fd(robot, _dp(q), _dp(qd), _dp(tau), _dp(out), _dp(M)) == 0
    return out, M.reshape(7, 7)


def pgs_stats(reset=True):
    """(contact-solver sweeps, cube sim steps) accumulated by this process"""
    out = (C.c_uint64 * 2)()
    lib().orc_pgs_stats(out, 1 if reset else 0)
    return int(out[0]), int(out[1])


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def reset_uniforms(seed, gid, episode, block):
    u = (C.c_float * 4)()
    lib().orc_reset_uniforms(seed, gid, episode, block, u)
    return np.array(list(u), dtype=np.float32)


class OracleSim:
    """Batch of reference-restated envs on the CPU (fp64)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._keep = cfg
        self.h = lib().orc_create(C.byref(cfg))
        if not self.h:
            raise ValueError("orc_create failed (bad config)")
        self.n = cfg.n_envs
        self.obs_dim = lib().orc_obs_dim(self.h)
        self.act_dim = lib().orc_action_dim(self.h)
        self.obs = np.zeros((self.n, self.obs_dim), np.float32)
        self.reward = np.zeros(self.n, np.float64)
        self.done = np.zeros(self.n, np.uint8)
        self.success = np.zeros(self.n, np.uint8)

    def close(self):
        if getattr(self, "_pool", None) is not None:
            lib().orc_pool_destroy(self._pool[0])
            self._pool = None
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data
        lib().orc_reset(self.h, m, self.obs.ctypes.data)
        return self.obs.copy()

    def step(self, action, lo=0, hi=None):
        a = np.ascontiguousarray(action, np.float32)
        assert a.shape == (self.n, self.act_dim)
        lib().orc_step_range(self.h, lo, self.n if hi is None else hi, a.ctypes.data, self.obs.ctypes.data,
                             self.reward.ctypes.data, self.done.ctypes.data, self.success.ctypes.data)
        return self.obs.copy(), self.reward.copy(), self.done.copy(), self.success.copy()

    def grip_distance(self):
        """pick: the distance the last step compared with the 6 mm finger-closing threshold (rl_pick_env.py:412)"""
        a = np.zeros(self.n, np.float64)
        lib().orc_grip_distance(self.h, _dp(a))
        return a

    def run_mt(self, actions, steps, nthreads):
        """`steps` Env.steps of the whole batch inside one C call, split over `nthreads` persistent worker threads;
        actions = a ring [n_sets, n, act_dim] (step k uses set k % n_sets).  The CPU-baseline leg of bench.py."""
        a = np.ascontiguousarray(actions, np.float32)
        assert a.ndim == 3 and a.shape[1:] == (self.n, self.act_dim)
        pool = getattr(self, "_pool", None)
        if pool is None or pool[1] != nthreads:
            if pool is not None:
                lib().orc_pool_destroy(pool[0])
            pool = self._pool = (lib().orc_pool_create(self.h, int(nthreads)), int(nthreads))
        lib().orc_run_mt(pool[0], a.ctypes.data, a.shape[0], int(steps), self.obs.ctypes.data, self.reward.ctypes.data,
                         self.done.ctypes.data, self.success.ctypes.data)
        return self.obs, self.reward, self.done, self.success

    def set_state(self, field, arr):
        dt = np.int32 if field in INT_FIELDS else np.float32
        a = np.ascontiguousarray(arr, dt)
        rc = lib().orc_set_state(self.h, field, a.ctypes.data, a.nbytes)
        assert rc == 0, rc

    def get_state(self, field):
        dt = np.int32 if field in INT_FIELDS else np.float32
        w = FIELD_WIDTH[field]
        a = np.zeros((self.n, w) if w > 1 else (self.n,), dt)
        rc = lib().orc_get_state(self.h, field, a.ctypes.data, a.nbytes)
        assert rc == 0, rc
        return a

    def get_state_f64(self, field):
        w = FIELD_WIDTH[field]
        a = np.zeros((self.n, w) if w > 1 else (self.n,), np.float64)
        rc = lib().orc_get_state_f64(self.h, field, _dp(a), a.size)
        assert rc == 0, rc
        return a
