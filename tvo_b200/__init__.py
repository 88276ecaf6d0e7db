"""tvo_b200 — B200-native truncated E-step / M-step for TVO (drop-in for the hot path of tvlearn/tvo).

Mirrors the reference's public names for this path:
    tvo_b200.get_device / _set_device / get_run_policy        (tvo/__init__.py)
    tvo_b200.variational.{EVOVariationalStates, RandomSampledVarStates, FullEM}
    tvo_b200.models.{BSC, NoisyOR, SSSC}
    tvo_b200.trainer.Trainer
    tvo_b200.utils.{parallel, data.TVODataLoader}
Underneath, every hot op is a hand-written sm_100a kernel reached through the C ABI in
include/tvo_b200.h (libtvo_b200.so, loaded by tvo_b200._lib).  There is no CPU fallback.
"""
from .utils.global_settings import get_device, get_run_policy, _set_device

__all__ = ["get_device", "_set_device", "get_run_policy"]
__version__ = "0.1.0"
