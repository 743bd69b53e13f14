"""diy_gym_b200: B200-native, lockstep-batched simulation backend behind the DIYGym API."""
from .addons.addon import Addon, AddonFactory, Receptor  # noqa: F401
from .config import Configuration  # noqa: F401


def __getattr__(name):   # lazy: importing the package must not need torch / CUDA
    if name == 'DIYGym':
        from .diy_gym import DIYGym
        return DIYGym
    if name == 'Model':
        from .model import Model
        return Model
    raise AttributeError(name)
