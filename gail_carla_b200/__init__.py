"""gail_carla_b200 - B200-native learning hot path of gustavokcouto/gail-carla.

Drop-in classes (same names and call signatures as the reference modules they replace):
    RolloutStorage  <- tools/storage.py          Policy          <- tools/model.py
    PPO             <- algo/ppo.py               Discriminator   <- algo/wdgail.py
    RunningMeanStd  <- common/running_mean_std.py
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/gail_carla_b200.h; there is no
CPU fallback - importing works anywhere, calling needs the built library and a CUDA device.
The training iteration around them (tools/learn.py) is `gail_carla_b200.learn.gail_learning`.
"""
from .storage import RolloutStorage
from .model import Policy
from .ppo import PPO
from .wdgail import Discriminator
from .running_mean_std import RunningMeanStd, update_mean_var_count_from_moments

__all__ = ["RolloutStorage", "Policy", "PPO", "Discriminator", "RunningMeanStd", "update_mean_var_count_from_moments"]
