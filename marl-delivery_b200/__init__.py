"""marl-delivery_b200: B200-native batched environment transition / reward / featurisation path of
nguyenduchuyiu/marl-delivery behind the reference's own API."""
from .batch import MAPPO_SHAPING, QMIX_SHAPING, STAT_NAMES, DeliveryBatch, compute_gae, encode_action  # noqa: F401
from .env import Environment  # noqa: F401
from .env_vectorized import VectorizedEnv  # noqa: F401
from . import greedyagent, helper, idq, maps, qmix_helper, rollout, sharding  # noqa: F401
