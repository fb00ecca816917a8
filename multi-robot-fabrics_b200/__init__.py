"""multi-robot-fabrics_b200 -- B200-native hot path of tud-amr/multi-robot-fabrics.

Batched multi-robot dynamic-fabric actions and Rollout-Fabrics (RF / RF-CV) horizons as hand-written sm_100a
CUDA kernels behind a C-ABI (include/mrf_b200.h), with the reference's planner call surface on top.
"""
from . import _lib, scenarios  # noqa: F401
from ._lib import Handle, MrfConfig, MrfError, default_config  # noqa: F401

__all__ = ["Handle", "MrfConfig", "MrfError", "default_config", "scenarios"]
