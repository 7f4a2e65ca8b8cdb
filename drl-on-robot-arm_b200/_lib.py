"""ctypes binding of the C-ABI in include/armsim.h (libarmsim.so).

This is the ONLY compute path of the package: if the CUDA library is missing or no CUDA device is usable the
calls raise -- there is no CPU or PyTorch fallback (and nothing here touches oracle/).
"""
import ctypes as C
import os

from . import _build

NJ = 7
ABI_VERSION = 5
TASK_REACH, TASK_PUSH, TASK_PICK, TASK_KUKA_REACH = 0, 1, 2, 3
ROBOT_KUKA_IIWA, ROBOT_DIANA_S1, ROBOT_CUSTOM = 0, 1, 2
MODE_IK_TELEPORT, MODE_TORQUE = 0, 1
MAP_AUTO, MAP_LANE = 0, 1
(F_Q, F_QD, F_GOAL, F_STEP, F_EPISODE, F_CUBE_POS, F_CUBE_QUAT, F_CUBE_LINVEL, F_CUBE_ANGVEL, F_LAST_DIST, F_GRIP,
 F_IK_ITERS, F_EP_RETURN, F_EXPLORE_COUNT) = range(14)
STATE_FIELDS = tuple(f for f in range(14) if f != F_IK_ITERS)      # everything a checkpoint has to carry
FIELD_WIDTH = {F_Q: 7, F_QD: 7, F_GOAL: 3, F_STEP: 1, F_EPISODE: 1, F_CUBE_POS: 3, F_CUBE_QUAT: 4, F_CUBE_LINVEL: 3,
               F_CUBE_ANGVEL: 3, F_LAST_DIST: 1, F_GRIP: 1, F_IK_ITERS: 1, F_EP_RETURN: 1, F_EXPLORE_COUNT: 1}
INT_FIELDS = (F_STEP, F_EPISODE, F_IK_ITERS, F_EXPLORE_COUNT)

EXPORTS = [
    "armsim_abi_version", "armsim_last_error", "armsim_default_config", "armsim_create", "armsim_destroy",
    "armsim_reset", "armsim_step", "armsim_step_host", "armsim_reset_host", "armsim_set_state", "armsim_get_state",
    "armsim_obs_dim", "armsim_action_dim", "armsim_n_envs", "armsim_mapping", "armsim_launch_count", "armsim_fk_host",
    "armsim_host_buffers", "armsim_host_server", "armsim_step_ex", "armsim_step_tracked", "armsim_step_host_async", "armsim_step_host_wait",
    "armsim_explore", "armsim_policy_act", "armsim_track_episodes", "armsim_episode_stats", "armsim_set_episode_stats",
    "armsim_replay_create", "armsim_replay_destroy", "armsim_replay_begin", "armsim_replay_store", "armsim_replay_sample",
    "armsim_replay_gather", "armsim_replay_info", "armsim_replay_table", "armsim_replay_last_error",
    "armsim_replay_state_bytes", "armsim_replay_get_state", "armsim_replay_set_state",
]


class ArmsimChain(C.Structure):
    _fields_ = [("base_xyz", C.c_double * 3), ("base_rpy", C.c_double * 3),
                ("xyz", (C.c_double * 3) * NJ), ("rpy", (C.c_double * 3) * NJ),
                ("lower", C.c_double * NJ), ("upper", C.c_double * NJ), ("effort", C.c_double * NJ),
                ("velocity", C.c_double * NJ), ("damping", C.c_double * NJ),
                ("mass", C.c_double * NJ), ("com", (C.c_double * 3) * NJ), ("inertia", (C.c_double * 6) * NJ)]


class ArmsimConfig(C.Structure):
    """include/armsim.h ArmsimConfig."""
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


class ArmReplayConfig(C.Structure):
    """include/armsim.h ArmReplayConfig."""
    _fields_ = [("struct_size", C.c_int32), ("n_envs", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32),
                ("window", C.c_int32), ("table_cap", C.c_int32), ("kind", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64)]


class ArmsimError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libarmsim.so (building it first if the sources are newer).  Raises when it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if os.environ.get("ARMSIM_LIB"):         # A/B experiments: another build of the SAME ABI (nvcc ... -D<variant> -o <path>)
        path = os.environ["ARMSIM_LIB"]
    elif not os.path.exists(path) or (_build.needs_build() and os.environ.get("ARMSIM_AUTO_REBUILD", "1") != "0"):
        path = _build.build_libarmsim()      # missing, or older than csrc/ / include/: never load a stale library
    try:
        L = C.CDLL(path)
    except OSError as e:  # loud: no fallback
        raise ArmsimError("cannot load %s: %s (the package has no CPU fallback)" % (path, e)) from e
    vp, i32 = C.c_void_p, C.c_int32
    L.armsim_abi_version.restype = i32
    L.armsim_last_error.restype = C.c_char_p
    L.armsim_default_config.argtypes = [i32, C.POINTER(ArmsimConfig)]
    L.armsim_create.argtypes = [C.POINTER(ArmsimConfig), C.POINTER(vp)]
    L.armsim_destroy.argtypes = [vp]
    L.armsim_destroy.restype = None
    L.armsim_reset.argtypes = [vp, vp, vp, vp]
    L.armsim_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.armsim_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.armsim_step_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.armsim_step_tracked.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.armsim_step_host_async.argtypes = [vp, vp]
    L.armsim_host_server.argtypes = [vp, i32]
    L.armsim_explore.argtypes = [vp, vp, C.c_float, C.c_float, vp, vp]
    L.armsim_policy_act.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, C.c_float, C.c_float, C.c_float, vp, vp]
    L.armsim_track_episodes.argtypes = [vp, vp, vp, vp, vp]
    L.armsim_episode_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.armsim_set_episode_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.armsim_step_host_wait.argtypes = [vp, vp, vp, vp, vp]
    L.armsim_reset_host.argtypes = [vp, vp, vp]
    L.armsim_host_buffers.argtypes = [vp] + [C.POINTER(vp)] * 5
    L.armsim_set_state.argtypes = [vp, i32, vp, C.c_size_t]
    L.armsim_get_state.argtypes = [vp, i32, vp, C.c_size_t]
    for f in ("armsim_obs_dim", "armsim_action_dim", "armsim_n_envs", "armsim_mapping"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = i32
    L.armsim_launch_count.argtypes = [vp]
    L.armsim_launch_count.restype = C.c_int64
    L.armsim_fk_host.argtypes = [vp, vp, i32, vp, vp]
    L.armsim_replay_create.argtypes = [C.POINTER(ArmReplayConfig), C.POINTER(vp)]
    L.armsim_replay_destroy.argtypes = [vp]
    L.armsim_replay_destroy.restype = None
    L.armsim_replay_begin.argtypes = [vp, vp, vp]
    L.armsim_replay_store.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.armsim_replay_sample.argtypes = [vp, i32, i32, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp]
    L.armsim_replay_gather.argtypes = [vp, i32, vp, vp, vp, C.c_float, vp, vp, vp, vp, vp, vp]
    L.armsim_replay_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.armsim_replay_table.argtypes = [vp, vp, vp, vp, i32]
    L.armsim_replay_last_error.restype = C.c_char_p
    L.armsim_replay_state_bytes.argtypes = [vp]
    L.armsim_replay_state_bytes.restype = C.c_int64
    L.armsim_replay_get_state.argtypes = [vp, vp, C.c_int64]
    L.armsim_replay_set_state.argtypes = [vp, vp, C.c_int64]
    if L.armsim_abi_version() != ABI_VERSION:
        raise ArmsimError("libarmsim ABI version %d != %d" % (L.armsim_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise ArmsimError("libarmsim error %d: %s" % (rc, lib().armsim_last_error().decode()))


def check_replay(rc):
    if rc != 0:
        raise ArmsimError("libarmsim replay error %d: %s" % (rc, lib().armsim_replay_last_error().decode()))


def default_config(task, **overrides):
    cfg = ArmsimConfig()
    check(lib().armsim_default_config(task, C.byref(cfg)))
    for k, v in overrides.items():
        cur = getattr(cfg, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(cfg, k, v)
    return cfg
