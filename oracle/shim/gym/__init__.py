"""TEST INFRASTRUCTURE: the sliver of `gym` the reference imports (`gym.Env`, `gym.spaces`, `gym.utils.seeding`),
so that /root/reference/diy_gym can be imported unmodified on top of the oracle (see oracle/shim/pybullet.py)."""
from . import spaces  # noqa: F401
from . import utils  # noqa: F401


class Env:
    metadata = {}

    def __init__(self):
        pass
