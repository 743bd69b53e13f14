"""Add-on plugin surface, same shape as the reference's (`diy_gym/addons/addon.py:5-210`):

    class MyAddon(Addon):
        def __init__(self, parent, config): ...      # parent is a Model or the DIYGym environment
        def update(self, action): ...                 # action: tensor / dict of tensors with leading dim num_envs
        def reset(self, mask=None): ...               # mask: [num_envs] bool tensor of environments being reset
        def observe(self): ...                        # -> tensor / dict of tensors [num_envs, ...] or None
        def reward(self): ...                         # -> [num_envs] tensor or None
        def is_terminal(self): ...                    # -> [num_envs] bool tensor or None
    AddonFactory.register_addon('my_addon', MyAddon)

Built-in add-ons are *compiled*: instead of running Python per step they emit an op into the scene program
(`compile(builder)`), which the fused CUDA step executes for every environment; their hooks then only hand
out views of the device buffers.  User add-ons run as batched PyTorch code on the state views the parent
`Model` exposes (`diy_gym_b200/model.py`).
"""
from collections import OrderedDict

from .. import spaces


class AddonFactory:
    """Name -> class registry (singleton, lazily filled with the built-ins; `addon.py:5-81` of the reference)."""
    instance = None

    class _Registry:
        def __init__(self):
            from . import builtin
            self.addons = dict(builtin.BUILTIN_ADDONS)

    @staticmethod
    def get():
        if AddonFactory.instance is None:
            AddonFactory.instance = AddonFactory._Registry()
        return AddonFactory.instance

    @staticmethod
    def build(name, parent, config):
        return AddonFactory.get().addons[name](parent, config)

    @staticmethod
    def register_addon(name, cls):
        AddonFactory.get().addons[name] = cls


class Addon:
    def __init__(self, parent, config):
        self.parent = parent
        self.name = getattr(config, 'name', None)
        self.action_space = None
        self.observation_space = None
        self.hide = config.get('hide', False)

    # ---- compile-time hook (built-ins only) ----
    def compile(self, builder):
        """Emit scene ops.  Called once, after every model exists, in sorted receptor / add-on order."""

    def bind(self, env):
        """Called once the device world exists (buffers can be sliced here)."""

    # ---- run-time hooks, as in the reference ----
    def update(self, action):
        pass

    def reset(self, mask=None):
        pass

    def observe(self):
        pass

    def reward(self):
        pass

    def is_terminal(self):
        pass


class Receptor:
    """Anything an add-on can be attached to (a Model or the environment); `addon.py:189-210` of the reference."""
    def __init__(self):
        self.addons = OrderedDict()
        self.models = OrderedDict()

    def build_spaces(self):
        obs_space, act_space = spaces.Dict({}), spaces.Dict({})
        for name, addon in self.addons.items():
            if addon.hide:
                continue
            if addon.observation_space is not None:
                obs_space.spaces[name] = addon.observation_space
            if addon.action_space is not None:
                act_space.spaces[name] = addon.action_space
        return obs_space, act_space
