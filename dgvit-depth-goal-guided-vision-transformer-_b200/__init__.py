"""dgvit_b200 — B200-native (sm_100a) implementation of the DGViT actor-critic hot path.

Import as ``dgvit_b200`` (the shim package at the repo root points here; the directory
name carries the reference's repository name and is not a valid Python identifier).

Public surface (mirrors the reference):
    GoTPolicy, GoTQNetwork, GoT          nn.Module drop-ins (vn/got_sac_network.py, vn/GoalFormer.py)
    QNetwork                             CNN twin-Q critic drop-in (vn/got_sac_network.py:125-170)
    SAC, ReplayStore                     agent drop-in (vn/DRL.py)
    soft_update, hard_update             vn/utils.py:31-37
    depth_augment                        vn/env_lab.py:420-434,78-90,69-76,295-299
"""
from .modules import GoT, GoTPolicy, GoTQNetwork, QNetwork, set_seed, weights_init_   # noqa: F401
from .agent import SAC, ReplayStore                                        # noqa: F401
from .ops import soft_update, hard_update, depth_augment                   # noqa: F401
from . import parallel                                                     # noqa: F401
from . import _lib                                                         # noqa: F401

__all__ = ["GoT", "GoTPolicy", "GoTQNetwork", "QNetwork", "SAC", "ReplayStore", "soft_update", "hard_update",
           "depth_augment", "set_seed", "weights_init_"]
