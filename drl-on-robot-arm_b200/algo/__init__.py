"""Agent registry with the reference's names (algo/__init__.py:26-37: `getattr(algo, opt.algo)(state_dim=, action_dim=,
action_bound=)`, main.py:95).  MLP agents only: the *_CNN variants belong to the camera envs, which are out of scope."""
from .agents import DADDPG_MLP, DARC_MLP, DATD3_MLP, DDPG_MLP, TD3_MLP
from .nets import PolicyNet, QValueNet, TwinQValueNet

__all__ = ["DDPG_MLP", "TD3_MLP", "DADDPG_MLP", "DATD3_MLP", "DARC_MLP", "PolicyNet", "QValueNet", "TwinQValueNet"]
