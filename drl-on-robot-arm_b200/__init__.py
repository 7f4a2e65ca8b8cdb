"""drl-on-robot-arm_b200: B200-native batched replacement for the env.step() hot path of Shimly-2/DRL-on-robot-arm.

Import name: `drl_on_robot_arm_b200` (see drl_on_robot_arm_b200.py at the repo root; the directory keeps the
hyphenated project name).  Compute happens only in libarmsim.so (hand-written sm_100a CUDA behind include/armsim.h).
"""
from . import _build, _lib, config, envs, metrics, replay, spaces
from ._lib import ArmsimError
from .config import opt
from .replay import TrajectoryReplay
from .envs import ArmSimHandle, BatchedArmEnv, KukaReachEnv, RLPickEnv, RLPushEnv, RLReachEnv

__all__ = ["envs", "config", "opt", "spaces", "ArmsimError", "ArmSimHandle", "BatchedArmEnv", "RLReachEnv",
           "RLPushEnv", "RLPickEnv", "KukaReachEnv", "TrajectoryReplay", "replay"]
__version__ = "0.1.0"
