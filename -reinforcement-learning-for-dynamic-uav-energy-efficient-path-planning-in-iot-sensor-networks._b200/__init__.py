"""uavsim-b200: B200-native batched UAV-LoRa environment step (one hot path of the reference,
`UAVEnvironment.reset/step`), behind a C ABI (include/uavsim.h) and an SB3-VecEnv-shaped host class.

Importing the package loads libuavsim.so (building it with nvcc when stale); there is no CPU
implementation to fall back to.
"""
from . import _native
from ._native import UavsimError  # noqa: F401

_native.load()  # fail loudly at import time if the CUDA library is missing

from .sharding import allreduce_episode_stats, make_sharded_env, shard_range  # noqa: E402,F401
from .vec_env import CURRICULUM_STAGES, MAX_SENSORS_LIMIT, UAVVecEnv  # noqa: E402,F401

__all__ = ["UAVVecEnv", "UavsimError", "CURRICULUM_STAGES", "MAX_SENSORS_LIMIT", "shard_range",
           "make_sharded_env", "allreduce_episode_stats"]
