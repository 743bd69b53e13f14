"""Minimal `gym.spaces` work-alike (gym is not a dependency of this backend).

Mirrors the subset of the gym space API the reference relies on
(`diy_gym/addons/addon.py:196-210`, `diy_gym/utils.py:6-28`): `Box`, `Dict`, `Discrete`,
`MultiDiscrete`, `MultiBinary`, `Tuple`, each with `sample()` / `contains()`.  Spaces describe the
*per-environment* shape; the batched backend adds a leading `num_envs` dimension to every leaf.
"""
from collections import OrderedDict

import numpy as np


class Space:
    shape = None
    dtype = None
    _rng = np.random.default_rng(0)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype='float32'):
        self.dtype = np.dtype(dtype)
        if shape is None:
            low = np.asarray(low, dtype=self.dtype)
            high = np.asarray(high, dtype=self.dtype)
            shape = np.broadcast(low, high).shape
        shape = tuple(int(s) for s in shape)
        self.shape = shape
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return 'Box(%s, %s, %s, %s)' % (self.low.min(initial=0), self.high.max(initial=0), self.shape, self.dtype)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) \
            and np.allclose(self.high, other.high)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype('int64')

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype('int64')

    def sample(self):
        return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))


class MultiBinary(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = (self.n, )
        self.dtype = np.dtype('int8')

    def sample(self):
        return self._rng.integers(0, 2, size=self.n).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))


class Dict(Space):
    """Ordered mapping of sub-spaces; `.spaces` is mutable exactly as the reference uses it."""
    def __init__(self, spaces=None):
        self.spaces = OrderedDict(spaces or {})

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys()) and all(
            self.spaces[k].contains(v) for k, v in x.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def seed(self, seed=None):
        for i, s in enumerate(self.spaces.values()):
            s.seed(None if seed is None else seed + i)
        return [seed]

    def __repr__(self):
        return 'Dict(' + ', '.join('%s:%r' % kv for kv in self.spaces.items()) + ')'
