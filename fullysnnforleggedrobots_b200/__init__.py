"""B200-native spiking actor-critic hot path.

Public surface mirrors rsl_rl's for this path only:
    ActorCritic, PopSpikeActor      modules/actor_critic.py of the four reference forks
    PPO, RolloutStorage             algorithms/ppo.py, storage/rollout_storage.py
    AMPPPO (+ discriminator, replay) algorithms/amp_ppo.py of the AMP fork
"""
from . import _lib  # noqa: F401
from .actor import PopSpikeActor  # noqa: F401
from .actor_critic import (ActorCritic, ActorCriticAMP, ActorCriticCassie, ActorCriticHumanoid,  # noqa: F401
                           ActorCriticRMA, VARIANTS)
from .engine import CriticEngine, MlpEngine, SnnEngine, gae  # noqa: F401
from .ppo import PPO, PPORMA  # noqa: F401
from .amp import AMPPPO, AMPDiscriminator, Normalizer, ReplayBuffer  # noqa: F401
from .storage import RolloutStorage, RolloutStorageRMA  # noqa: F401
from .export import export_policy, load_checkpoint, load_policy, save_checkpoint  # noqa: F401

__all__ = ["PopSpikeActor", "SnnEngine", "CriticEngine", "MlpEngine", "gae", "ActorCritic", "ActorCriticAMP", "ActorCriticCassie",
           "ActorCriticHumanoid", "ActorCriticRMA", "VARIANTS", "PPO", "PPORMA", "AMPPPO", "AMPDiscriminator", "ReplayBuffer", "Normalizer", "RolloutStorage", "RolloutStorageRMA",
           "export_policy", "load_policy", "save_checkpoint", "load_checkpoint"]
