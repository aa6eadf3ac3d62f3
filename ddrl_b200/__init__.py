"""ddrl_b200 — B200-native learner hot path of DDRL (decentralized per-leg PPO controllers).

Product code only: hand-written sm_100a kernels behind the C ABI in ``include/ddrl_b200.h`` plus the host-side
mirror of the reference's RLlib ModelV2 / multi-agent interface.  Nothing here imports ``oracle/``."""
from ._lib import DDRLError, LIB_PATH  # noqa: F401
from .catalog import ModelCatalog, register_custom_model  # noqa: F401
from .config import PPOConfig  # noqa: F401

__all__ = ["DDRLError", "LIB_PATH", "ModelCatalog", "register_custom_model", "PPOConfig"]
