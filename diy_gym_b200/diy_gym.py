"""DIYGym: the gym-style environment facade of the reference (`diy_gym/diy_gym.py:53-225`), batched.

    env = DIYGym('config.yaml', num_envs=4096, device=0)
    obs = env.reset()
    obs, reward, terminal, info = env.step(action)      # every leaf has a leading num_envs dimension

Same constructor argument, YAML schema, nested-dict structure (receptors and add-ons in sorted-name order), same
`observation_space` / `action_space` trees (per-environment shapes) and the same options (`sum_rewards`,
`terminal_if_any`, `terminal_if_all`, `flatten_observations`, `flatten_actions`, `max_episode_steps`, `hot_start`,
`timestep`, `update_freq`, `solver_iterations`, `gravity`; extension keys `max_contacts`, `engine_semantics`).  New keyword arguments, never required in the YAML:
`num_envs`, `device`, `seed`, `auto_reset`, `team`, `env_id_offset` (global id of environment 0, for multi-GPU).

Construction compiles the whole scene (models + add-on program) into flat buffers (`compiler/scene.py`) and
creates one device world (`backend.World`); `step()` is then: built-in add-ons copy their action slices, user
add-ons run their batched torch code, ONE fused kernel launch advances all environments and evaluates every
built-in sensor / reward / terminal, and the returned dict is assembled from views of the output buffers.
"""
from collections import OrderedDict

import torch

from . import spaces
from .addons.addon import AddonFactory, Receptor
from .compiler.scene import SceneBuilder
from .config import Configuration
from .model import Model
from .utils import flatten, get_bounds_for_space, unflatten, walk_dict


class DIYGym(Receptor):
    metadata = {'render.modes': []}

    def __init__(self, config_file, num_envs=1, device=0, seed=1234, auto_reset=False, team=0, env_id_offset=0, world_factory=None,
                 compile_only=False):
        Receptor.__init__(self)
        config = config_file if isinstance(config_file, Configuration) else Configuration.from_file(config_file)
        self.config = config
        self.name = config.name
        self.env = self
        self.num_envs = int(num_envs)
        self._env_id_offset = int(env_id_offset)
        self.auto_reset = bool(auto_reset)
        self._max_episode_steps = config.get('max_episode_steps') if 'max_episode_steps' in config else None
        self.hot_start = int(config.get('hot_start', 1))
        # `render` (GUI) and the camera_* debug-visualiser keys of the reference are accepted and ignored: headless only.
        timestep = config.get('timestep', 1 / 240.)
        sub_steps = int(1. / config.get('update_freq', 100) / timestep)   # diy_gym.py:77 -> 2 with the defaults
        iterations = config.get('solver_iterations', 150)
        gravity = config.get('gravity', [0.0, 0.0, -9.81])
        self.builder = SceneBuilder(timestep=timestep, substeps=max(sub_steps, 1), iterations=iterations, gravity=gravity,
                                    hot_start=self.hot_start, max_contacts=int(config.get('max_contacts', 0)),
                                    semantics=tuple(config.get('engine_semantics', ())))   # compiler/scene.py SEMANTICS (extension key)
        self.world = None

        self.models = OrderedDict(sorted({child.name: Model(child, env=self) for child in config.find_all('model')}.items(),
                                         key=lambda t: t[0]))
        self.addons = OrderedDict(sorted({child.name: AddonFactory.build(child.get('addon'), self, child)
                                          for child in config.find_all('addon')}.items(), key=lambda t: t[0]))
        self.receptors = OrderedDict(sorted({**self.models, self.name: self}.items(), key=lambda t: t[0]))

        self.collapse_rewards_func = sum if config.get('sum_rewards', False) else None
        self.collapse_terminals_func = any if config.get('terminal_if_any', False) else all if config.get('terminal_if_all', False) else None
        self.flatten_observations = config.get('flatten_observations', False)
        self.flatten_actions = config.get('flatten_actions', False)

        # ---- compile the add-on program in the order the reference walks add-ons (sorted receptors, sorted add-ons)
        for receptor in self.receptors.values():
            for addon in receptor.addons.values():
                addon.compile(self.builder)
        self._timer_op = None
        if self._max_episode_steps is not None:
            self._timer_op = self.builder.add_op('EPISODE_TIMER', [], [float(self._max_episode_steps)], n_term=1)
        self.scene = self.builder.finalize()

        if compile_only:   # host-side compilation only (scene + spaces); nothing can be stepped
            self._build_spaces()
            return
        # ---- device world (no CPU fallback: backend.World raises without CUDA / the built library)
        if world_factory is None:
            from .backend import World
            self.world = World(self.scene, self.num_envs, device=device, team=team, seed=seed, env_id_offset=env_id_offset)
        else:
            self.world = world_factory(self.scene, self.num_envs, seed, env_id_offset)   # test hook (tests/emul)
        for receptor in self.receptors.values():
            for addon in receptor.addons.values():
                addon.bind(self)
        self._timer = self.world.term[:, self._timer_op['term_off']] if self._timer_op is not None else None

        self.seed(seed)
        self.reset()
        self._build_spaces()

    def _build_spaces(self):
        self.observation_space, self.action_space = spaces.Dict({}), spaces.Dict({})
        for name, receptor in self.receptors.items():
            obs_space, act_space = receptor.build_spaces()
            if len(obs_space.spaces):
                self.observation_space.spaces[name] = obs_space
            if len(act_space.spaces):
                self.action_space.spaces[name] = act_space
        if self.flatten_observations:
            lows, highs = [flatten(get_bounds_for_space(self.observation_space, opt), batched=False) for opt in [True, False]]
            self.original_observation_space = self.observation_space
            self.observation_space = spaces.Box(low=lows, high=highs)
        if self.flatten_actions:
            lows, highs = [flatten(get_bounds_for_space(self.action_space, opt), batched=False) for opt in [True, False]]
            self.original_action_space = self.action_space
            self.action_space = spaces.Box(low=lows, high=highs)

    # --------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        """Seeds the per-environment random streams of respawn / dynamics_randomizer (keyed by seed, global environment id and
        reset count).  Unlike the reference (diy_gym.py:124-128, which seeds an np_random nothing reads - SURVEY App. D.7) it
        takes effect: from the next reset on."""
        self._seed = seed
        if seed is not None and self.world is not None and hasattr(self.world, 'set_seed'):
            self.world.set_seed(int(seed), self._env_id_offset)
        return [seed]

    def reset(self, mask=None):
        """Reset every environment (mask=None) or the masked ones ([num_envs] bool tensor).  Add-on reset hooks
        run first (built-ins inside the kernel, user add-ons in Python), then `hot_start` physics steps."""
        for addon, takes_mask in self._reset_hooks():
            if takes_mask:
                addon.reset(mask)
            else:
                addon.reset()
        self.world.reset(mask)
        return self.observe()

    def _reset_hooks(self):
        """(add-on, accepts a mask argument) for every add-on that overrides reset(); a user add-on written for the reference
        has `def reset(self)` - it is then reset for every environment.  Decided from the signature, once: a TypeError raised
        INSIDE a user hook is the user's to see."""
        if not hasattr(self, '_reset_hook_list'):
            import inspect
            from .addons.addon import Addon
            hooks = []
            for receptor in self.receptors.values():
                for addon in receptor.addons.values():
                    if type(addon).reset is Addon.reset:
                        continue
                    params = [p for p in inspect.signature(addon.reset).parameters.values()
                              if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD, p.VAR_POSITIONAL)]
                    hooks.append((addon, len(params) > 0))
            self._reset_hook_list = hooks
        return self._reset_hook_list

    def observe(self):
        ret = self.walk_addons(lambda addon: addon.observe())
        return flatten(ret) if self.flatten_observations else ret

    def reward(self):
        ret = self.walk_addons(lambda addon: addon.reward())
        return walk_dict(ret, self.collapse_rewards_func) if self.collapse_rewards_func is not None else ret

    def is_terminal(self):
        ret = self.walk_addons(lambda addon: addon.is_terminal())
        if self._timer is not None:
            if self.name not in ret:
                ret[self.name] = OrderedDict()
            ret[self.name]['episode_timer'] = self._timer.bool()
        return walk_dict(ret, self.collapse_terminals_func) if self.collapse_terminals_func is not None else ret

    @property
    def step_counter(self):
        """[num_envs] steps since each environment's last reset."""
        return self.world.state[:, self.scene.hdr['S_STEP']]

    def step(self, action):
        if self.flatten_actions:
            action = unflatten(torch.as_tensor(action, device=self.world.action.device, dtype=torch.float32).reshape(self.num_envs, -1),
                               self.original_action_space)
        present = set()
        for receptor_name, receptor_action in action.items():
            for addon_name, addon_action in receptor_action.items():
                addon = self.receptors[receptor_name].addons[addon_name]
                addon.update(addon_action)
                present.add(id(addon))
        self._apply_action_mask(present)
        self.world.step()
        obs, rew, term = self.observe(), self.reward(), self.is_terminal()
        info = {}
        if self.auto_reset:
            done = term if isinstance(term, torch.Tensor) else walk_dict(term, any)
            if isinstance(done, torch.Tensor):
                # The returned leaves are VIEWS of the world's output buffers and the masked reset rewrites the rows of the
                # finished environments: the terminal step's reward and observation are copied out first (the reward a learner
                # gets for the last transition must be that transition's).  No host synchronisation: the masked reset is launched
                # unconditionally, a block whose mask bytes are all zero returns at once.
                def clone(t):
                    if isinstance(t, dict):
                        return OrderedDict((k, clone(v)) for k, v in t.items())
                    return t.clone() if isinstance(t, torch.Tensor) else t
                rew, info = clone(rew), {'terminal_observation': clone(obs), 'done': done}
                obs = self.reset(done)
        return obs, rew, term, info

    def _apply_action_mask(self, present):
        """Only add-ons present in the action dict are updated (diy_gym.py:202-204): switch the others' ops off."""
        if not hasattr(self, '_action_ops'):
            self._action_ops = [(id(a), a.op) for r in self.receptors.values() for a in r.addons.values()
                                if getattr(a, 'op', None) is not None and a.op['n_act']]
            self._op_index = {id(o): k for k, o in enumerate(self.builder.ops)}
            self._mask_key = None
        key = tuple(aid in present for aid, _ in self._action_ops)
        if key != self._mask_key:
            enabled = [1] * len(self.builder.ops)
            for (aid, op), on in zip(self._action_ops, key):
                enabled[self._op_index[id(op)]] = int(on)
            self.world.set_action_mask(enabled)
            self._mask_key = key

    def walk_addons(self, func):
        ret = OrderedDict()
        for receptor_name, receptor in self.receptors.items():
            receptor_ret = OrderedDict()
            for addon_name, addon in receptor.addons.items():
                addon_ret = func(addon)
                if addon_ret is not None:
                    receptor_ret[addon_name] = addon_ret
            if len(receptor_ret):
                ret[receptor_name] = receptor_ret
        return ret

    def sample_action(self, generator=None):
        """Uniform random action batch with the structure of `action_space` (device tensors)."""
        dev = self.world.action.device

        def rec(sp):
            if isinstance(sp, spaces.Dict):
                return OrderedDict((k, rec(v)) for k, v in sp.spaces.items())
            lo = torch.as_tensor(sp.low, device=dev, dtype=torch.float32)
            hi = torch.as_tensor(sp.high, device=dev, dtype=torch.float32)
            u = torch.rand((self.num_envs, ) + tuple(sp.shape), device=dev, generator=generator)
            return lo + (hi - lo) * u
        return rec(self.action_space)

    def close(self):
        if self.world is not None:
            self.world.close()
            self.world = None
