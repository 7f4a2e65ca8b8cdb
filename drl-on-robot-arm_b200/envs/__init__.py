"""Env registry, looked up by name like the reference's envs/__init__.py:1-3 (`getattr(envs, opt.env)`)."""
from .batched import ArmSimHandle, BatchedArmEnv
from .single import KukaReachEnv, RLPickEnv, RLPushEnv, RLReachEnv

__all__ = ["RLReachEnv", "RLPushEnv", "RLPickEnv", "KukaReachEnv", "BatchedArmEnv", "ArmSimHandle"]
