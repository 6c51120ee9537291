"""amp_extensions_b200 — B200-native batched learned-dynamics env step of gym-simenv / MILO.

Host-side mirror of the reference's plugin surface for this one hot path:

  DynamicsEnsemble / DynamicsModel   (reference milo/milo/dynamics.py)
  RBFLinearCost / MLPCost            (reference milo/milo/linear_cost.py)
  GAILCost (evaluation side)         (reference milo/milo/gail_cost.py)
  AmpDataset                         (reference milo/milo/datasets.py)
  SimEnv / VecSimEnv                 (reference gym-simenv/gym_simenv/envs/sim_env.py)
  ImitationReward                    (reference DeepMimicCore/scenes/SceneImitate.cpp)

All compute goes through libsimstep.so (include/simstep.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from .datasets import AmpDataset  # noqa: F401
from .dynamics import DynamicsEnsemble, DynamicsModel  # noqa: F401
from .engine import Engine, HumanoidTermination  # noqa: F401
from .linear_cost import MLPCost, RBFLinearCost  # noqa: F401
from .gail_cost import GAILCost  # noqa: F401
from .sim_env import SimEnv, VecSimEnv  # noqa: F401
from .character import Character, humanoid3d  # noqa: F401
from .imitation import ImitationReward  # noqa: F401
from .motion import MotionClip  # noqa: F401
